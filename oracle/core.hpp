// ORACLE — TEST INFRASTRUCTURE ONLY (see analyzer.hpp header).
// CPU restatement of the reference's pkg/core (allocation.go, server.go,
// system.go, model.go, serviceclass.go) over the index-keyed SoA of
// include/wva_b200.h.  Go maps keyed by name become index loops in ascending
// index order — the canonical order (the shim indexes SORTED names), which is
// one of the orders Go's random map iteration can produce.
//
// Citations are relative to /root/reference/pkg/core/ unless noted.
#pragma once
#include "../include/wva_b200.h"
#include "analyzer.hpp"
#include <cmath>
#include <vector>

namespace wva_oracle {

// pkg/config/defaults.go:18,21
static const int MaxQueueToBatchRatio = 10;
static const float AccelPenaltyFactor = 0.1f;

// core.Allocation (allocation.go:13-24); state encodes nil / "" accelerator.
struct Allocation {
  uint8_t state = WVA_ALLOC_NONE;
  int acc = -1;
  long long numReplicas = 0;  // Go int is 64-bit
  int batchSize = 0;
  float cost = 0, value = 0, itl = 0, ttft = 0, rho = 0, maxArrvRatePerReplica = 0;
  int nSolves = 0;
  long nStates = 0;
};

// Go int(math.Ceil(x)) on amd64 (CVTTSD2SQ): NaN / out of range -> MinInt64
inline long long goIntCeil(double x) {
  double c = std::ceil(x);
  if (!(c < 9223372036854775808.0) || !(c >= -9223372036854775808.0)) return (long long)0x8000000000000000ULL;
  return (long long)c;
}

// model.go:40-42 + 52-55: NumInstances
inline int NumInstances(const wva_system& s, int model, int acc) {
  int c = s.perf_acc_count[(size_t)model * s.n_acc + acc];
  return c <= 0 ? 1 : c;
}

// allocation.go:251-280
inline Allocation zeroLoadAllocation(const wva_system& s, int srv, int model, int acc) {
  Allocation a;
  int numReplicas = s.srv_min_replicas[srv];
  if (numReplicas == 0) {
    a.state = WVA_ALLOC_EMPTY;  // accelerator "", all zero, value 0
    return a;
  }
  size_t pi = (size_t)model * s.n_acc + acc;
  int maxBatchSize = s.perf_max_batch[pi];
  if (s.srv_max_batch[srv] > 0) maxBatchSize = s.srv_max_batch[srv];
  long long totalNumInstances = (long long)NumInstances(s, model, acc) * numReplicas;
  float cost = s.acc_cost[acc] * (float)totalNumInstances;
  float alpha = s.perf_alpha[pi], beta = s.perf_beta[pi];
  float decodeTime = alpha + beta;
  float maxDecodeTime = alpha + beta * (float)maxBatchSize;
  float prefillTime = alpha + beta;
  float maxServTime = prefillTime + maxDecodeTime;
  float maxArrvRatePerReplica = (float)maxBatchSize / maxServTime;
  a.state = WVA_ALLOC_ACC;
  a.acc = acc;
  a.numReplicas = numReplicas;
  a.batchSize = maxBatchSize;
  a.cost = cost;
  a.itl = decodeTime;
  a.ttft = prefillTime;
  a.rho = 0;
  a.maxArrvRatePerReplica = maxArrvRatePerReplica;
  a.value = a.cost;
  return a;
}

// allocation.go:27-155.  Returns state NONE for Go's nil.
inline Allocation CreateAllocation(const wva_system& s, int srv, int acc) {
  Allocation none;
  if (acc < 0 || acc >= s.n_acc) return none;                       // :42-44
  // load checks :51-54
  float arrival = s.srv_arrival[srv];
  int inTok = s.srv_in_tokens[srv], outTok = s.srv_out_tokens[srv];
  if (arrival < 0 || inTok < 0 || outTok < 0) return none;
  int model = s.srv_model[srv];
  if (model < 0 || model >= s.n_models) return none;                // :57-59
  size_t pi = (size_t)model * s.n_acc + acc;
  if (!s.perf_present[pi]) return none;                             // :60-62
  if (!s.srv_target_present[srv]) return none;                      // :65-70
  if (arrival == 0 || outTok == 0) return zeroLoadAllocation(s, srv, model, acc);  // :73-75

  int K = outTok;                                                   // :78
  int N;
  if (s.srv_max_batch[srv] > 0) N = s.srv_max_batch[srv];           // :82-83
  else N = (int)std::max<int64_t>((int64_t)s.perf_max_batch[pi] * s.perf_at_tokens[pi] / K, 1);  // :85 (Go int is 64-bit; division truncates)
  int maxQueue = N * MaxQueueToBatchRatio;                          // :87

  Configuration qc;
  qc.MaxBatchSize = N;
  qc.MaxQueueSize = maxQueue;
  qc.parms = ServiceParms{s.perf_alpha[pi], s.perf_beta[pi], s.perf_gamma[pi]};
  RequestSize rq{(float)inTok, (float)K};
  if (!QueueAnalyzer::checkConfig(qc) || !QueueAnalyzer::checkRequest(rq)) return none;  // :105-109
  QueueAnalyzer qa(qc, rq);

  TargetPerf tp{s.srv_slo_ttft[srv], s.srv_slo_itl[srv], s.srv_slo_tps[srv]};  // :111-115
  AnalysisMetrics metrics;
  if (!qa.Size(tp, nullptr, &metrics, nullptr)) {                   // :118-122
    none.nSolves = (int)qa.model.solves;
    none.nStates = qa.model.statesVisited;
    return none;
  }
  float rateStar = metrics.Throughput;                              // :123

  float totalRate;                                                  // :126-131
  if (tp.TargetTPS == 0) totalRate = arrival / 60;
  else totalRate = tp.TargetTPS / (float)K;
  long long numReplicas = goIntCeil((double)totalRate / (double)rateStar);  // :132
  numReplicas = std::max<long long>(numReplicas, s.srv_min_replicas[srv]);  // :133

  long long totalNumInstances =                                     // :136 (Go int multiply wraps)
      (long long)((unsigned long long)NumInstances(s, model, acc) * (unsigned long long)numReplicas);
  float cost = s.acc_cost[acc] * (float)totalNumInstances;          // :137

  float rate = totalRate / (float)numReplicas;                      // :140
  if (!qa.Analyze(rate, &metrics)) {                                // :141-145
    none.nSolves = (int)qa.model.solves;
    none.nStates = qa.model.statesVisited;
    return none;
  }
  Allocation a;
  a.state = WVA_ALLOC_ACC;
  a.acc = acc;
  a.numReplicas = numReplicas;
  a.batchSize = N;
  a.cost = cost;
  a.itl = metrics.AvgTokenTime;                                     // :147
  a.ttft = metrics.AvgWaitTime + metrics.AvgPrefillTime;            // :148 (quirk Q2)
  a.rho = metrics.Rho;
  a.maxArrvRatePerReplica = rateStar / 1000;                        // :152
  a.value = a.cost;                                                 // :153
  a.nSolves = (int)qa.model.solves;
  a.nStates = qa.model.statesVisited;
  return a;
}

// allocation.go:283-292.  `a` is the server's current allocation
// (AllocationFromData, allocation.go:324-333: accelerator, numReplicas, cost).
inline float TransitionPenalty(int curAcc, int curReplicas, float curCost, const Allocation& b) {
  bool sameAcc = (b.state == WVA_ALLOC_EMPTY) ? (curAcc == WVA_CUR_ACC_EMPTY)
                                              : (curAcc == b.acc);
  if (sameAcc) {
    if (curReplicas == b.numReplicas) return 0;
    return b.cost - curCost;
  }
  return AccelPenaltyFactor * (curCost + b.cost) + (b.cost - curCost);
}

// server.go:55-82 Server.Calculate + GetCandidateAccelerators, for all servers
// (system.go:258-268 System.Calculate).  out is row-major [S][A].
inline void SystemCalculate(const wva_system& s, std::vector<Allocation>& out) {
  const int S = s.n_servers, A = s.n_acc;
  out.assign((size_t)S * A, Allocation{});
#pragma omp parallel for schedule(dynamic, 4)
  for (int srv = 0; srv < S; srv++) {
    int curAcc = s.srv_cur_acc[srv];
    bool restrict_ = s.srv_keep_acc[srv] && curAcc != WVA_CUR_ACC_EMPTY;  // server.go:71-72
    for (int g = 0; g < A; g++) {
      if (restrict_ && g != curAcc) continue;  // unknown cur acc (-2) => no candidates (server.go:75-77)
      Allocation a = CreateAllocation(s, srv, g);
      if (a.state != WVA_ALLOC_NONE) {
        // curAllocation is never nil (server.go:49, allocation.go:324)
        a.value = TransitionPenalty(curAcc, s.srv_cur_replicas[srv], s.srv_cur_cost[srv], a);  // server.go:60-63
      }
      out[(size_t)srv * A + g] = a;
    }
  }
}

}  // namespace wva_oracle
