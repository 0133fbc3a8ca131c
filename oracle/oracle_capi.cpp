// ORACLE — TEST INFRASTRUCTURE ONLY (see analyzer.hpp header).
// C entry points so tests/ (ctypes) and bench.py's cpu_baseline / --impl
// reference legs can drive the CPU restatement with the same SoA buffers the
// product's C-ABI takes.  Built by oracle/Makefile into oracle/liboracle.so.
#include "analyzer.hpp"
#include "core.hpp"
#include "saturation.hpp"
#include "solver.hpp"
#include "pipeline_v2.hpp"
#include <cstring>
#include <climits>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace wva_oracle;

static void store_cand(const Allocation& a, wva_candidates* out, size_t i) {
  out->state[i] = a.state;
  out->num_replicas[i] = (int32_t)std::min<long long>(std::max<long long>(a.numReplicas, INT32_MIN), INT32_MAX);
  out->batch_size[i] = a.batchSize;
  out->cost[i] = a.cost;
  out->value[i] = a.value;
  out->itl[i] = a.itl;
  out->ttft[i] = a.ttft;
  out->rho[i] = a.rho;
  out->max_arrv_rate[i] = a.maxArrvRatePerReplica;
  if (out->n_solves) out->n_solves[i] = a.nSolves;
}

static Allocation load_cand(const wva_candidates* in, size_t i, int acc) {
  Allocation a;
  a.state = in->state[i];
  a.acc = (a.state == WVA_ALLOC_ACC) ? acc : -1;
  a.numReplicas = in->num_replicas[i];
  a.batchSize = in->batch_size[i];
  a.cost = in->cost[i];
  a.value = in->value[i];
  a.itl = in->itl[i];
  a.ttft = in->ttft[i];
  a.rho = in->rho[i];
  a.maxArrvRatePerReplica = in->max_arrv_rate[i];
  return a;
}

extern "C" {

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// System.Calculate.  nthreads <= 0: all cores.  Returns total chain solves via *solves.
int oracle_calculate(const wva_system* sys, wva_candidates* out, int nthreads, int64_t* solves, int64_t* states) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<Allocation> cand;
  SystemCalculate(*sys, cand);
  int64_t ns = 0, nst = 0;
  for (size_t i = 0; i < cand.size(); i++) {
    store_cand(cand[i], out, i);
    ns += cand[i].nSolves;
    nst += cand[i].nStates;
  }
  if (solves) *solves = ns;
  if (states) *states = nst;
  return 0;
}

// Manager.Optimize on candidates produced by oracle_calculate (or by the GPU).
int oracle_solve(const wva_system* sys, const wva_candidates* in, wva_solution* out) {
  const int S = sys->n_servers, A = sys->n_acc, T = sys->n_types;
  std::vector<Allocation> cand((size_t)S * A);
  for (int s = 0; s < S; s++)
    for (int g = 0; g < A; g++) cand[(size_t)s * A + g] = load_cand(in, (size_t)s * A + g, g);
  Solution sol;
  ManagerOptimize(*sys, cand, sol);
  for (int s = 0; s < S; s++) {
    const Allocation& a = sol.alloc[(size_t)s];
    out->state[s] = a.state;
    out->acc[s] = (a.state == WVA_ALLOC_ACC) ? a.acc : -1;
    out->num_replicas[s] = (int32_t)std::min<long long>(std::max<long long>(a.numReplicas, INT32_MIN), INT32_MAX);
    out->batch_size[s] = a.batchSize;
    out->cost[s] = a.cost;
    out->value[s] = a.value;
    out->itl[s] = a.itl;
    out->ttft[s] = a.ttft;
    out->rho[s] = a.rho;
    out->max_arrv_rate[s] = a.maxArrvRatePerReplica;
  }
  for (int t = 0; t < T; t++) {
    out->type_count[t] = sol.typeCount[(size_t)t];
    out->type_cost[t] = sol.typeCost[(size_t)t];
  }
  return 0;
}

// float32 sequential by-type cost, exactly as System.AllocateByType accumulates it
int oracle_type_cost_f32(const wva_system* sys, const wva_solution* solin, float* type_cost_f32) {
  for (int t = 0; t < sys->n_types; t++) type_cost_f32[t] = 0;
  for (int s = 0; s < sys->n_servers; s++) {
    if (solin->state[s] != WVA_ALLOC_ACC || sys->srv_model[s] < 0) continue;
    type_cost_f32[sys->acc_type[solin->acc[s]]] += solin->cost[s];
  }
  return 0;
}

// Replica-grid evaluator (see wva_analyze_grid in include/wva_b200.h).
int oracle_analyze_grid(const wva_system* sys, int R, uint8_t* ok, float* ttft, float* itl, float* rho,
                        float* tput, int32_t* frontier, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const wva_system& s = *sys;
  const int S = s.n_servers, A = s.n_acc;
#pragma omp parallel for schedule(dynamic, 1)
  for (int srv = 0; srv < S; srv++) {
    for (int g = 0; g < A; g++) {
      size_t pair = (size_t)srv * A + g;
      size_t base = pair * (size_t)R;
      auto blank = [&]() {
        for (int r = 0; r < R; r++) {
          if (ok) ok[base + r] = 0;
          if (ttft) ttft[base + r] = 0;
          if (itl) itl[base + r] = 0;
          if (rho) rho[base + r] = 0;
          if (tput) tput[base + r] = 0;
        }
        if (frontier) frontier[pair] = 0;
      };
      float arrival = s.srv_arrival[srv];
      int inTok = s.srv_in_tokens[srv], outTok = s.srv_out_tokens[srv];
      int model = s.srv_model[srv];
      if (arrival < 0 || inTok < 0 || outTok < 0 || model < 0 || model >= s.n_models) { blank(); continue; }
      size_t pi = (size_t)model * A + g;
      if (!s.perf_present[pi] || !s.srv_target_present[srv] || arrival == 0 || outTok == 0) { blank(); continue; }
      int K = outTok;
      int N;
      if (s.srv_max_batch[srv] > 0) N = s.srv_max_batch[srv];
      else N = (int)std::max<int64_t>((int64_t)s.perf_max_batch[pi] * s.perf_at_tokens[pi] / K, 1);
      Configuration qc;
      qc.MaxBatchSize = N;
      qc.MaxQueueSize = N * MaxQueueToBatchRatio;
      qc.parms = ServiceParms{s.perf_alpha[pi], s.perf_beta[pi], s.perf_gamma[pi]};
      RequestSize rq{(float)inTok, (float)K};
      if (!QueueAnalyzer::checkConfig(qc) || !QueueAnalyzer::checkRequest(rq)) { blank(); continue; }
      QueueAnalyzer qa(qc, rq);
      float sloTTFT = s.srv_slo_ttft[srv], sloITL = s.srv_slo_itl[srv], sloTPS = s.srv_slo_tps[srv];
      float totalRate = (sloTPS == 0) ? arrival / 60 : sloTPS / (float)K;
      float lambdaMax = qa.rateRange.Max / 1000;
      float lambdaTPS = lambdaMax * (1 - StabilitySafetyFraction);
      int front = 0;
      for (int r = 1; r <= R; r++) {
        AnalysisMetrics m;
        float rate = totalRate / (float)r;
        bool good = qa.Analyze(rate, &m);
        size_t o = base + (size_t)(r - 1);
        if (ok) ok[o] = good ? 1 : 0;
        if (ttft) ttft[o] = good ? m.AvgWaitTime + m.AvgPrefillTime : 0;
        if (itl) itl[o] = good ? m.AvgTokenTime : 0;
        if (rho) rho[o] = good ? m.Rho : 0;
        if (tput) tput[o] = good ? m.Throughput : 0;
        if (good && front == 0) {
          bool meets = (sloTTFT <= 0 || m.AvgTTFT <= sloTTFT) && (sloITL <= 0 || m.AvgTokenTime <= sloITL) &&
                       (sloTPS <= 0 || rate / 1000 <= lambdaTPS);
          if (meets) front = r;
        }
      }
      if (frontier) frontier[pair] = front;
    }
  }
  return 0;
}

// MM1KModel.Solve for n independent triples
int oracle_mm1k_eval(int64_t n, const float* lambda, const float* mu, const int32_t* K, uint8_t* valid,
                     float* avg_resp, float* avg_wait, float* avg_serv, float* avg_num, float* avg_queue,
                     float* throughput, float* rho) {
  for (int64_t i = 0; i < n; i++) {
    MM1KModel m(K[i]);
    m.Solve(lambda[i], mu[i]);
    valid[i] = m.isValid ? 1 : 0;
    avg_resp[i] = m.avgRespTime;
    avg_wait[i] = m.avgWaitTime;
    avg_serv[i] = m.avgServTime;
    avg_num[i] = m.avgNumInSystem;
    avg_queue[i] = m.avgQueueLength;
    throughput[i] = m.throughput;
    rho[i] = m.rho;
  }
  return 0;
}

int oracle_saturation_v1(const wva_saturation_in* in, const wva_saturation_out* out) {
  SaturationV1(*in, *out);
  return 0;
}

int oracle_limit(int64_t D, int T, const int32_t* acc_type, const int32_t* current, const int32_t* target,
                 const int32_t* gpr, const double* spare, const double* cost, const int32_t* type_limit,
                 int32_t* out_target, int32_t* out_gpus, uint8_t* out_limited) {
  Limit(D, T, acc_type, current, target, gpr, spare, cost, type_limit, out_target, out_gpus, out_limited);
  return 0;
}

// ---- V2 pipeline (pipeline_v2.hpp) ----
int oracle_saturation_v2(const wva_saturation_v2_in* in, const wva_saturation_v2_out* out) {
  using namespace oracle_v2;
  std::vector<int> ready; std::vector<double> a, b, c, d;
  for (int64_t m = 0; m < in->n_models; m++) {
    const int v0 = in->model_variant_off[m], v1 = in->model_variant_off[m + 1], V = v1 - v0;
    V2ModelIn mi;
    mi.V = V; mi.vro = in->variant_replica_off + v0;
    mi.rep_total_kv = (const long long*)in->rep_total_kv_tokens; mi.rep_tokens_in_use = (const long long*)in->rep_tokens_in_use;
    mi.rep_queue_len = (const long long*)in->rep_queue_length; mi.rep_k2 = (const long long*)in->rep_k2;
    mi.rep_avg_in = in->rep_avg_input_tokens; mi.rep_avg_out = in->rep_avg_output_tokens; mi.rep_hit = in->rep_prefix_hit_rate;
    mi.slice_order = in->rep_slice_order;
    mi.var_current = in->var_current + v0; mi.var_pending = in->var_pending + v0; mi.var_fallback = in->var_fallback_capacity + v0;
    mi.kv_threshold = in->cfg_kv_threshold[m]; mi.scale_up_threshold = in->cfg_scale_up_threshold[m];
    mi.scale_down_boundary = in->cfg_scale_down_boundary[m];
    mi.has_queue = in->sched_queue_size != nullptr;
    mi.queue_size = mi.has_queue ? in->sched_queue_size[m] : 0; mi.queue_bytes = mi.has_queue ? in->sched_queue_bytes[m] : 0;
    ready.assign(V + 1, 0); a.assign(V + 1, 0); b.assign(V + 1, 0); c.assign(V + 1, 0); d.assign(V + 1, 0);
    V2ModelOut mo;
    mo.rep_k1 = (long long*)out->rep_k1; mo.rep_effective = (long long*)out->rep_effective; mo.rep_demand = (long long*)out->rep_demand;
    mo.rep_saturated = out->rep_saturated;
    mo.var_ready = ready.data(); mo.var_cap = a.data(); mo.var_total_cap = b.data(); mo.var_total_demand = c.data(); mo.var_util = d.data();
    saturation_v2_model(mi, mo);
    for (int v = 0; v < V; v++) {
      if (out->var_ready) out->var_ready[v0 + v] = ready[v];
      if (out->var_per_replica_capacity) out->var_per_replica_capacity[v0 + v] = a[v];
      if (out->var_total_capacity) out->var_total_capacity[v0 + v] = b[v];
      if (out->var_total_demand) out->var_total_demand[v0 + v] = c[v];
      if (out->var_utilization) out->var_utilization[v0 + v] = d[v];
    }
    if (out->mod_total_supply) out->mod_total_supply[m] = mo.total_supply;
    if (out->mod_total_demand) out->mod_total_demand[m] = mo.total_demand;
    if (out->mod_utilization) out->mod_utilization[m] = mo.utilization;
    if (out->mod_required_capacity) out->mod_required_capacity[m] = mo.required;
    if (out->mod_spare_capacity) out->mod_spare_capacity[m] = mo.spare;
  }
  return 0;
}

int oracle_cost_aware_optimize(int64_t M, int64_t V, const int32_t* mvo, const double* required, const double* spare,
                               const uint8_t* has_result, const int32_t* current, const double* cost, const double* cap,
                               int32_t* target) {
  (void)V;
  for (int64_t m = 0; m < M; m++) {
    const int v0 = mvo[m], v1 = mvo[m + 1];
    if (has_result && !has_result[m]) { for (int v = v0; v < v1; v++) target[v] = -1; continue; }
    oracle_v2::cost_aware_model(v1 - v0, required[m], spare[m], current + v0, cost + v0, cap + v0, target + v0);
  }
  return 0;
}

int oracle_enforce_ranked(int64_t M, int64_t V, const int32_t* mvo, const uint8_t* s2z, const double* request_count,
                          const uint8_t* request_error, const double* cost, const uint8_t* has_cost, const int32_t* name_rank,
                          int32_t* target, uint8_t* applied) {
  (void)V;
  for (int64_t m = 0; m < M; m++) {
    const int v0 = mvo[m], v1 = mvo[m + 1];
    const bool app = oracle_v2::enforce_model(v1 - v0, target + v0, cost + v0, has_cost ? has_cost + v0 : nullptr, s2z[m] != 0,
                                              request_count[m], request_error && request_error[m], name_rank ? name_rank + v0 : nullptr);
    if (applied) applied[m] = app ? 1 : 0;
  }
  return 0;
}

int oracle_enforce(int64_t M, int64_t V, const int32_t* mvo, const uint8_t* s2z, const double* request_count,
                   const uint8_t* request_error, const double* cost, const uint8_t* has_cost, int32_t* target, uint8_t* applied) {
  (void)V;
  for (int64_t m = 0; m < M; m++) {
    const int v0 = mvo[m], v1 = mvo[m + 1];
    const bool app = oracle_v2::enforce_model(v1 - v0, target + v0, cost + v0, has_cost ? has_cost + v0 : nullptr, s2z[m] != 0,
                                              request_count[m], request_error && request_error[m]);
    if (applied) applied[m] = app ? 1 : 0;
  }
  return 0;
}

long long oracle_estimate_capacity_from_params(long long max_batched_tokens, long long max_num_seqs, double avg_in, double avg_out) {
  return oracle_v2::estimate_capacity_from_params(max_batched_tokens, max_num_seqs, avg_in, avg_out);
}

// ---- known-answer-test helpers (pin the oracle to the reference's own tests) ----
float oracle_prefill_time(float a, float b, float g, float in, float out, float n) {
  return PrefillTime(ServiceParms{a, b, g}, RequestSize{in, out}, n);
}
float oracle_decode_time(float a, float b, float g, float in, float out, float n) {
  return DecodeTime(ServiceParms{a, b, g}, RequestSize{in, out}, n);
}
float oracle_iteration_time(float a, float b, float g, float in, float out, float n) {
  return IterationTime(ServiceParms{a, b, g}, RequestSize{in, out}, n);
}
int oracle_within_tolerance(float x, float v, float tol) { return WithinTolerance(x, v, tol) ? 1 : 0; }

// BinarySearch over f(x) = c2*x*x + c1*x + c0 (the shapes the reference's utils_test.go uses);
// fail_at >= 0 makes the eval function return an error for x >= fail_at.
int oracle_binary_search_poly(float xMin, float xMax, float yTarget, float c2, float c1, float c0, int use_fail,
                              float fail_at, float* x, int* ind) {
  return BinarySearch(xMin, xMax, yTarget,
                      [=](float xx, float* y) {
                        if (use_fail && xx >= fail_at) return false;
                        *y = c2 * xx * xx + c1 * xx + c0;
                        return true;
                      },
                      x, ind);
}

// MM1ModelStateDependent: solve a sequence of lambdas on ONE model object (quirk Q1/Q5).
// stats rows: [valid, lambda, rho, avgRespTime, avgWaitTime, avgServTime, avgNumInSystem,
//              avgQueueLength, throughput, avgNumInServers]; p_out (may be NULL) gets the last p[0..K].
int oracle_statedep_solve(int K, const float* servRate, int n, const float* lambdas, int nl, float* stats,
                          double* p_out) {
  MM1ModelStateDependent m(K, std::vector<float>(servRate, servRate + n));
  for (int i = 0; i < nl; i++) {
    m.Solve(lambdas[i], 1);
    float* r = stats + (size_t)i * 10;
    r[0] = m.isValid ? 1.0f : 0.0f;
    r[1] = m.lambda; r[2] = m.rho; r[3] = m.avgRespTime; r[4] = m.avgWaitTime; r[5] = m.avgServTime;
    r[6] = m.avgNumInSystem; r[7] = m.avgQueueLength; r[8] = m.throughput; r[9] = m.avgNumInServers;
  }
  if (p_out) std::memcpy(p_out, m.p.data(), sizeof(double) * (size_t)(K + 1));
  return m.pathological ? 1 : 0;
}

// MM1KModel.Solve + probabilities
int oracle_mm1k_solve(int K, float lambda, float mu, float* stats, double* p_out) {
  MM1KModel m(K);
  m.Solve(lambda, mu);
  stats[0] = m.isValid ? 1.0f : 0.0f;
  stats[1] = m.lambda; stats[2] = m.rho; stats[3] = m.avgRespTime; stats[4] = m.avgWaitTime;
  stats[5] = m.avgServTime; stats[6] = m.avgNumInSystem; stats[7] = m.avgQueueLength; stats[8] = m.throughput;
  if (p_out) std::memcpy(p_out, m.p.data(), sizeof(double) * (size_t)(K + 1));
  return 0;
}

// QueueAnalyzer: NewQueueAnalyzer + Analyze.  Returns 0 ok, 1 config error, 2 analyze error.
// metrics: 9 floats in AnalysisMetrics order; range: [Min, Max].
int oracle_queue_analyze(int maxBatch, int maxQueue, float a, float b, float g, float in, float out, float rate,
                         float* metrics, float* range) {
  Configuration qc;
  qc.MaxBatchSize = maxBatch; qc.MaxQueueSize = maxQueue; qc.parms = ServiceParms{a, b, g};
  RequestSize rq{in, out};
  if (!QueueAnalyzer::checkConfig(qc) || !QueueAnalyzer::checkRequest(rq)) return 1;
  QueueAnalyzer qa(qc, rq);
  if (range) { range[0] = qa.rateRange.Min; range[1] = qa.rateRange.Max; }
  AnalysisMetrics m;
  if (!qa.Analyze(rate, &m)) return 2;
  std::memcpy(metrics, &m, sizeof(m));
  return 0;
}

// QueueAnalyzer.Size.  Returns 0 ok, 1 config error, 2 size error.
// rates: 3 floats (TargetRate), metrics: 9 floats, achieved: 3 floats; *solves = chain solves used.
int oracle_queue_size(int maxBatch, int maxQueue, float a, float b, float g, float in, float out, float ttft,
                      float itl, float tps, float* rates, float* metrics, float* achieved, int* solves) {
  Configuration qc;
  qc.MaxBatchSize = maxBatch; qc.MaxQueueSize = maxQueue; qc.parms = ServiceParms{a, b, g};
  RequestSize rq{in, out};
  if (!QueueAnalyzer::checkConfig(qc) || !QueueAnalyzer::checkRequest(rq)) return 1;
  QueueAnalyzer qa(qc, rq);
  TargetRate tr; AnalysisMetrics m; TargetPerf ach;
  bool ok = qa.Size(TargetPerf{ttft, itl, tps}, &tr, &m, &ach);
  if (solves) *solves = (int)qa.model.solves;
  if (!ok) return 2;
  std::memcpy(rates, &tr, sizeof(tr));
  std::memcpy(metrics, &m, sizeof(m));
  std::memcpy(achieved, &ach, sizeof(ach));
  return 0;
}

float oracle_transition_penalty(int curAcc, int curReplicas, float curCost, int bState, int bAcc, int bReplicas,
                                float bCost) {
  Allocation b;
  b.state = (uint8_t)bState; b.acc = bAcc; b.numReplicas = bReplicas; b.cost = bCost;
  return TransitionPenalty(curAcc, curReplicas, curCost, b);
}

}  // extern "C"
