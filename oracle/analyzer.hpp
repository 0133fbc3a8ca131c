// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference's pkg/analyzer (Go) in C++17.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may link or call this.  The product (libwva_b200.so) never does.
//
// Parity pin: checked against every golden vector the reference's own tests
// hold for this package (tests/test_oracle_kat.py).  The Go
// toolchain is absent, so the reference itself cannot be run here.
//
// Float rules followed (Go on amd64): every float32 expression rounds per
// operation, no FMA contraction (build with -ffp-contract=off, no -ffast-math,
// gradual underflow kept), float64(x) widening is exact, float32(x) rounds.
//
// All citations are relative to /root/reference/pkg/analyzer/.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <functional>
#include <vector>
#include <algorithm>

namespace wva_oracle {

// queueanalyzer.go:8-14
static const float Epsilon = 0.001f;
static const float StabilitySafetyFraction = 0.1f;
static const int DefaultMaxNumTokens = 8192;
// utils.go:8-9
static const float bs_epsilon = 1e-6f;
static const int maxIterations = 100;

// queueanalyzer.go:38-42
struct ServiceParms { float Alpha, Beta, Gamma; };
// queueanalyzer.go:45-48
struct RequestSize { float AvgInputTokens, AvgOutputTokens; };
// queueanalyzer.go:51-54
struct RateRange { float Min, Max; };
// queueanalyzer.go:57-67
struct AnalysisMetrics {
  float Throughput, AvgRespTime, AvgWaitTime, AvgNumInServ, AvgPrefillTime,
      AvgTokenTime, AvgTTFT, MaxRate, Rho;
};
// queueanalyzer.go:70-74
struct TargetPerf { float TargetTTFT, TargetITL, TargetTPS; };
// queueanalyzer.go:77-81
struct TargetRate { float RateTargetTTFT, RateTargetITL, RateTargetTPS; };

// queueanalyzer.go:261-265
inline float IterationTime(const ServiceParms& p, const RequestSize& r, float batchSize) {
  float tokensCompute = (r.AvgInputTokens + r.AvgOutputTokens) / (r.AvgOutputTokens + 1);
  float tokensMemory = r.AvgInputTokens + r.AvgOutputTokens / 2;
  return p.Alpha + batchSize * (p.Beta * tokensCompute + p.Gamma * tokensMemory);
}
// queueanalyzer.go:268-273
inline float PrefillTime(const ServiceParms& p, const RequestSize& r, float batchSize) {
  if (r.AvgInputTokens == 0) return 0;
  return IterationTime(p, r, batchSize) + (p.Beta + p.Gamma) * r.AvgInputTokens;
}
// queueanalyzer.go:276-279
inline float DecodeTime(const ServiceParms& p, const RequestSize& r, float batchSize) {
  return IterationTime(p, r, batchSize) + p.Beta +
         p.Gamma * (r.AvgInputTokens + r.AvgOutputTokens / 2);
}

// utils.go:12-23
inline bool WithinTolerance(float x, float value, float tolerance) {
  if (x == value) return true;
  if (value == 0 || tolerance < 0) return false;
  return std::fabs((double)((x - value) / value)) <= (double)tolerance;
}

// utils.go:26-70.  eval returns false on error.  Return code: 0 ok, 1 invalid
// range, 2 eval error.
inline int BinarySearch(float xMin, float xMax, float yTarget,
                        const std::function<bool(float, float*)>& eval,
                        float* xOut, int* indOut) {
  *xOut = 0; *indOut = 0;
  if (xMin > xMax) return 1;
  float yBounds[2];
  const float xs[2] = {xMin, xMax};
  for (int i = 0; i < 2; i++) {
    if (!eval(xs[i], &yBounds[i])) return 2;
    if (WithinTolerance(yBounds[i], yTarget, bs_epsilon)) { *xOut = xs[i]; *indOut = 0; return 0; }
  }
  bool increasing = yBounds[0] < yBounds[1];
  if ((increasing && yTarget < yBounds[0]) || (!increasing && yTarget > yBounds[0])) {
    *xOut = xMin; *indOut = -1; return 0;
  }
  if ((increasing && yTarget > yBounds[1]) || (!increasing && yTarget < yBounds[1])) {
    *xOut = xMax; *indOut = +1; return 0;
  }
  float xStar = 0, yStar = 0;
  for (int it = 0; it < maxIterations; it++) {
    xStar = 0.5f * (xMin + xMax);
    if (!eval(xStar, &yStar)) return 2;
    if (WithinTolerance(yStar, yTarget, bs_epsilon)) break;
    if ((increasing && yTarget < yStar) || (!increasing && yTarget > yStar)) xMax = xStar;
    else xMin = xStar;
  }
  *xOut = xStar; *indOut = 0;
  return 0;
}

// Go's math.Pow special cases are irrelevant here (finite positive base);
// glibc pow is within 1 ulp like Go's pure-Go Pow — MM1K outputs are float32
// and compared at 1e-6 relative, never bit-exactly.

// queuemodel.go:9-84 + mm1kmodel.go:10-108: the base class and M/M/1/K.
struct MM1KModel {
  // QueueModel fields (queuemodel.go:10-19)
  float lambda = 0, mu = 0, rho = 0;
  float avgRespTime = 0, avgWaitTime = 0, avgServTime = 0, avgNumInSystem = 0, avgQueueLength = 0;
  bool isValid = false;
  // MM1KModel fields (mm1kmodel.go:11-15)
  int K = 0;
  std::vector<double> p;
  double sumP = 0;
  float throughput = 0;

  explicit MM1KModel(int K_) : K(K_), p((size_t)K_ + 1, 0.0) {}
  virtual ~MM1KModel() {}

  // mm1kmodel.go:36-42
  virtual float ComputeRho() { return (lambda == mu) ? 1.0f : lambda / mu; }
  // mm1kmodel.go:45-47
  float GetRhoMax() const { return (float)K; }

  // queuemodel.go:27-37
  void Solve(float lambda_, float mu_) {
    lambda = lambda_;
    mu = mu_;
    rho = ComputeRho();
    if ((rho < 0) || (rho >= GetRhoMax()) || (lambda_ < 0) || (mu_ <= 0)) {
      isValid = false;
    } else {
      isValid = true;
      computeStatistics();
    }
  }

  // mm1kmodel.go:50-71
  virtual void computeProbabilities() {
    for (int i = 0; i <= K; i++) p[i] = 0;
    sumP = 1;
    if (!isValid) p[0] = 1;
    if (rho == 1) p[0] = 1 / (double)(K + 1);
    else p[0] = (1 - (double)rho) / (1 - std::pow((double)rho, (double)(K + 1)));
    sumP = 0;
    const double p0 = p[0];
    for (int i = 0; i <= K; i++) {
      // the reference reads m.p[0] each iteration; at i==0 it overwrites p[0]
      // with p0*Pow(rho,0) == p0*1 == p0, so hoisting is exact.
      p[i] = p0 * std::pow((double)rho, (double)i);
      sumP += p[i];
    }
  }

  // mm1kmodel.go:74-92
  virtual void computeStatistics() {
    if (!isValid) return;
    computeProbabilities();
    double temp = 0;
    for (int i = 0; i <= K; i++) temp += (double)i * p[i];
    avgNumInSystem = (float)temp;
    throughput = lambda * (1 - (float)p[K]);
    avgRespTime = avgNumInSystem / throughput;
    avgServTime = 1 / mu;
    avgWaitTime = avgRespTime - avgServTime;
    if (avgWaitTime < 0) avgWaitTime = 0;
    avgQueueLength = throughput * avgWaitTime;
  }
};

// mm1modelstatedependent.go:9-128
struct MM1ModelStateDependent : MM1KModel {
  std::vector<float> servRate;
  float avgNumInServers = 0;
  bool pathological = false;  // set when the reference would loop forever (guard only)
  long statesVisited = 0;     // instrumentation: chain states per lifetime
  long solves = 0;

  MM1ModelStateDependent(int K_, const std::vector<float>& sr) : MM1KModel(K_), servRate(sr) {}

  // mm1modelstatedependent.go:33-35  (quirk Q1: reads p[0] left by the previous solve)
  float ComputeRho() override { return 1 - (float)p[0]; }

  // mm1modelstatedependent.go:70-116
  void computeProbabilities() override {
    p[0] = 1;
    const double scale = DBL_MAX / (double)K;
    double sRate = 0;
    const int num = (int)servRate.size();
    for (int n = 0; n < K; n++) {
      if (n < num) sRate = (double)servRate[n];
      else sRate = (double)servRate[num - 1];
      p[n + 1] = p[n] * (double)lambda / sRate;
      int guard = 0;
      while (p[n + 1] < 0 || std::isinf(p[n + 1]) || std::isnan(p[n + 1])) {
        for (int i = 0; i <= n; i++) p[i] /= scale;
        p[n + 1] = p[n] * (double)lambda / sRate;
        if (++guard > 64) { pathological = true; break; }  // reference: infinite loop
      }
    }
    double sum = 0;
    for (int n = 0; n <= K; n++) {
      sum += p[n];
      if (sum < 0 || std::isinf(sum)) {
        sum = 0;
        for (int i = 0; i <= K; i++) {
          p[i] /= scale;
          if (i <= n) sum += p[i];
        }
      }
    }
    sumP = 0;
    for (int n = 0; n <= K; n++) {
      p[n] /= sum;
      sumP += p[n];
    }
    rho = ComputeRho();
    statesVisited += K + 1;
    solves += 1;
  }

  // mm1modelstatedependent.go:38-67
  void computeStatistics() override {
    if (!isValid) return;
    computeProbabilities();
    const int num = (int)servRate.size();
    double avgNumInServersD = 0;
    double avgNumInSystemD = 0;
    double sumPl = p[0];
    for (int i = 1; i <= K; i++) {
      avgNumInSystemD += (double)i * p[i];
      sumPl += p[i];
      if (i == num) avgNumInServersD = avgNumInSystemD + (1 - sumPl) * (double)num;
    }
    avgNumInServers = (float)avgNumInServersD;
    avgNumInSystem = (float)avgNumInSystemD;
    throughput = lambda * (1 - (float)p[K]);
    avgRespTime = avgNumInSystem / throughput;
    avgServTime = avgNumInServers / throughput;
    avgWaitTime = avgRespTime - avgServTime;
    if (avgWaitTime < 0) avgWaitTime = 0;
    avgQueueLength = throughput * avgWaitTime;
  }
};

// queueanalyzer.go:17-25, 28-33
struct Configuration {
  int MaxBatchSize = 0, MaxNumTokens = 0, MaxQueueSize = 0;
  ServiceParms parms{0, 0, 0};
  bool hasParms = true;
};

struct QueueAnalyzer {
  int MaxBatchSize, MaxNumTokens, MaxQueueSize;
  ServiceParms parms;
  RequestSize req;
  MM1ModelStateDependent model;
  RateRange rateRange;

  // queueanalyzer.go:95-124 BuildModel
  static std::vector<float> buildServRate(const Configuration& c, const RequestSize& r) {
    std::vector<float> servRate((size_t)c.MaxBatchSize);
    for (int n = 1; n <= c.MaxBatchSize; n++) {
      float prefillTime = PrefillTime(c.parms, r, (float)n);
      float decodeTime = r.AvgOutputTokens * DecodeTime(c.parms, r, (float)n);
      servRate[n - 1] = (float)n / (prefillTime + decodeTime);
    }
    return servRate;
  }

  QueueAnalyzer(const Configuration& c, const RequestSize& r)
      : MaxBatchSize(c.MaxBatchSize), MaxNumTokens(c.MaxNumTokens), MaxQueueSize(c.MaxQueueSize),
        parms(c.parms), req(r), model(c.MaxQueueSize + c.MaxBatchSize, buildServRate(c, r)) {
    float lambdaMin = model.servRate[0] * Epsilon;
    float lambdaMax = model.servRate[(size_t)c.MaxBatchSize - 1] * (1 - Epsilon);
    rateRange = RateRange{lambdaMin * 1000, lambdaMax * 1000};
  }

  // utils.go:95-118 (Configuration.check, RequestSize.check); queueanalyzer.go:83-92
  static bool checkConfig(Configuration& c) {
    if (c.MaxBatchSize <= 0 || c.MaxQueueSize < 0 || c.MaxNumTokens < 0 || !c.hasParms) return false;
    if (c.MaxNumTokens == 0) c.MaxNumTokens = DefaultMaxNumTokens;
    return true;
  }
  static bool checkRequest(const RequestSize& r) {
    return !(r.AvgInputTokens < 0 || r.AvgOutputTokens < 1);
  }

  // queueanalyzer.go:127-167.  Returns false on error.
  bool Analyze(float requestRate, AnalysisMetrics* out) {
    if (requestRate <= 0) return false;
    if (requestRate > rateRange.Max) return false;
    model.Solve(requestRate / 1000, 1);
    if (!model.isValid) return false;
    float avgNumInServ = model.avgNumInServers;
    float avgPrefillTime = PrefillTime(parms, req, avgNumInServ);
    float avgDecodeTime = (model.avgServTime - avgPrefillTime) / req.AvgOutputTokens;
    float avgTTFT = model.avgWaitTime + avgPrefillTime + avgDecodeTime;
    float rho = avgNumInServ / (float)MaxBatchSize;
    rho = std::min(std::max(rho, 0.0f), 1.0f);
    out->Throughput = model.throughput * 1000;
    out->AvgRespTime = model.avgRespTime;
    out->AvgWaitTime = model.avgWaitTime;
    out->AvgNumInServ = avgNumInServ;
    out->AvgPrefillTime = avgPrefillTime;
    out->AvgTokenTime = avgDecodeTime;
    out->AvgTTFT = avgTTFT;
    out->MaxRate = rateRange.Max;
    out->Rho = rho;
    return true;
  }

  // queueanalyzer.go:283-296 EvalTTFT
  bool evalTTFT(float x, float* y) {
    model.Solve(x, 1);
    if (!model.isValid) return false;
    float avgPrefillTime = PrefillTime(parms, req, model.avgNumInServers);
    float avgDecodeTime = (model.avgServTime - avgPrefillTime) / req.AvgOutputTokens;
    *y = model.avgWaitTime + avgPrefillTime + avgDecodeTime;
    return true;
  }
  // queueanalyzer.go:298-308 EvalITL
  bool evalITL(float x, float* y) {
    model.Solve(x, 1);
    if (!model.isValid) return false;
    float avgPrefillTime = PrefillTime(parms, req, model.avgNumInServers);
    *y = (model.avgServTime - avgPrefillTime) / req.AvgOutputTokens;
    return true;
  }

  // queueanalyzer.go:181-258.  Returns false on error.
  bool Size(const TargetPerf& tp, TargetRate* targetRate, AnalysisMetrics* metrics, TargetPerf* achieved) {
    // utils.go:121-128 TargetPerf.check
    if (tp.TargetITL < 0 || tp.TargetTTFT < 0 || tp.TargetTPS < 0) return false;
    float lambdaMin = rateRange.Min / 1000;
    float lambdaMax = rateRange.Max / 1000;
    int ind = 0;
    float lambdaStarTTFT = lambdaMax;
    if (tp.TargetTTFT > 0) {
      int rc = BinarySearch(lambdaMin, lambdaMax, tp.TargetTTFT,
                            [this](float x, float* y) { return evalTTFT(x, y); }, &lambdaStarTTFT, &ind);
      if (ind < 0 || rc != 0) return false;
    }
    float lambdaStarITL = lambdaMax;
    if (tp.TargetITL > 0) {
      int rc = BinarySearch(lambdaMin, lambdaMax, tp.TargetITL,
                            [this](float x, float* y) { return evalITL(x, y); }, &lambdaStarITL, &ind);
      if (ind < 0 || rc != 0) return false;
    }
    float lambdaStarTPS = lambdaMax;
    if (tp.TargetTPS > 0) lambdaStarTPS = lambdaMax * (1 - StabilitySafetyFraction);
    float lambda = std::min(std::min(lambdaStarTTFT, lambdaStarITL), lambdaStarTPS);
    float requestRate = lambda * 1000;
    if (!Analyze(requestRate, metrics)) return false;
    if (targetRate) *targetRate = TargetRate{lambdaStarTTFT * 1000, lambdaStarITL * 1000, lambdaStarTPS * 1000};
    if (achieved) *achieved = TargetPerf{metrics->AvgTTFT, metrics->AvgTokenTime, metrics->Throughput * req.AvgOutputTokens};
    return true;
  }
};

}  // namespace wva_oracle
