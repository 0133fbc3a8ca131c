// saturation_kernel.cuh — V1 saturation capacity model for a batch of models:
// saturation.Analyzer.AnalyzeModelSaturation / analyzeVariant / shouldScaleUp /
// isScaleDownSafe (internal/saturation/analyzer.go:31-280) and
// CalculateSaturationTargets (analyzer.go:290-439).
//
// One warp per model, one lane per variant (chunks of 32).  A lane streams its
// variant's replicas in slice order (the per-variant float64 sums are order
// dependent); the variant -> model accumulation runs in ascending variant index on
// every lane redundantly (canonical order, see oracle/saturation.hpp).  HBM-bound:
// 16 B per replica (kv float64 + queue int64) is the only large stream.
#pragma once
#include "wva_core.cuh"

namespace wva {

struct SatIn {
  long long n_models, n_variants, n_replicas;
  const int *model_variant_off, *variant_replica_off;
  const double* rep_kv; const long long* rep_queue;
  const double* var_cost; const int *var_current, *var_desired, *var_pending;
  const unsigned char* var_has_state;
  const double *cfg_kv_threshold, *cfg_queue_threshold, *cfg_kv_trigger, *cfg_queue_trigger;
};
struct SatOut {
  int *var_target, *var_replica_count, *var_non_saturated;
  double* var_max_kv; long long* var_max_queue; double *var_avg_spare_kv, *var_avg_spare_queue;
  unsigned char* rep_saturated;
  int *mod_total_replicas, *mod_non_saturated; double *mod_avg_spare_kv, *mod_avg_spare_queue;
  unsigned char* mod_flags;
  long long* partials;
};

#define SAT_FLAG_UP 1
#define SAT_FLAG_DOWN 2
#define SAT_FLAG_TRANS 4
#define SAT_FLAG_KV 8
#define SAT_FLAG_Q 16

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src) {
  int lo = __shfl_sync(mask, __double2loint(v), src), hi = __shfl_sync(mask, __double2hiint(v), src);
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(256) saturation_kernel(SatIn in, SatOut out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  long long n_up = 0, n_down = 0, n_trans = 0, sum_targets = 0;
  const double* __restrict__ in_kv = in.rep_kv;
  const long long* __restrict__ in_q = in.rep_queue;
  __shared__ double2 terms[8][32];
  double2* my_terms = terms[threadIdx.x >> 5];

  for (long long m = warp0; m < in.n_models; m += nwarps) {
    const int v0 = in.model_variant_off[m], v1 = in.model_variant_off[m + 1];
    const double kvThr = in.cfg_kv_threshold[m], qThr = in.cfg_queue_threshold[m];
    const double kvTrig = in.cfg_kv_trigger[m], qTrig = in.cfg_queue_trigger[m];
    double totalSpareKv = 0.0, totalSpareQueue = 0.0;
    int nonSaturated = 0, totalReplicas = 0, nAnalysed = 0;
    bool inTransition = false;
    int cheap_v = -1, exp_v = -1;
    double cheap_c = 0.0, exp_c = 0.0;

    // ---- phase A: analyzeVariant per lane, ordered combine ---------------------------------
    for (int c0 = v0; c0 < v1; c0 += 32) {
      const int v = c0 + lane;
      const bool act = v < v1;
      int cnt = 0, ns = 0;
      double sumKv = 0.0, sumQ = 0.0, maxKv = 0.0, avgKv = 0.0, avgQ = 0.0;
      long long maxQ = 0;
      if (act) {
        const int lo = in.variant_replica_off[v], hi = in.variant_replica_off[v + 1];
        cnt = hi - lo;
        // replicas are streamed 4 at a time: the 8 loads of a batch are independent (one memory
        // round trip), the accumulation below stays in slice order (float64 sums are order dependent)
        for (int base = lo; base < hi; base += 4) {
          double kvv[4]; long long qq[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const bool in = base + j < hi;
            kvv[j] = in ? __ldg(in_kv + base + j) : 0.0;
            qq[j] = in ? __ldg(in_q + base + j) : 0;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (base + j < hi) {
              const double kv = kvv[j];
              const long long q = qq[j];
              const bool sat = kv >= kvThr || (double)q >= qThr;               // analyzer.go:160-161
              if (out.rep_saturated) out.rep_saturated[base + j] = sat ? 1 : 0;
              if (!sat) {
                sumKv = d_add(sumKv, d_sub(kvThr, kv));                        // :167-171
                sumQ = d_add(sumQ, d_sub(qThr, (double)q));
                ns++;
              }
              if (kv > maxKv) maxKv = kv;                                      // :177-182
              if (q > maxQ) maxQ = q;
            }
          }
        }
        if (ns > 0) { avgKv = d_div(sumKv, (double)ns); avgQ = d_div(sumQ, (double)ns); }  // :188-191
        if (out.var_replica_count) out.var_replica_count[v] = cnt;
        if (out.var_non_saturated) out.var_non_saturated[v] = ns;
        if (out.var_max_kv) out.var_max_kv[v] = maxKv;
        if (out.var_max_queue) out.var_max_queue[v] = maxQ;
        if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = avgKv;
        if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = avgQ;
      }
      const bool analysed = act && cnt > 0;   // only variants with metrics enter VariantAnalyses
      // ordered accumulation over the chunk (analyzer.go:86-94)
      const double termKv = d_mul(avgKv, (double)ns), termQ = d_mul(avgQ, (double)ns);
      const unsigned amask = __ballot_sync(full, analysed);
      // in ascending variant order, through shared memory (broadcast reads; exact sequential sum)
      __syncwarp();
      my_terms[lane] = make_double2(termKv, termQ);
      __syncwarp();
      for (unsigned rest = amask; rest; rest &= rest - 1) {
        const double2 t2 = my_terms[__ffs(rest) - 1];
        totalSpareKv = d_add(totalSpareKv, t2.x);
        totalSpareQueue = d_add(totalSpareQueue, t2.y);
      }
      int t;
      t = analysed ? ns : 0;  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(full, t, o);  nonSaturated += t;
      t = act ? cnt : 0;      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(full, t, o);  totalReplicas += t;
      nAnalysed += __popc(amask);
      // transition checks (analyzer.go:322-341); a variant without state reads the zero value
      const bool hs = act && (!in.var_has_state || in.var_has_state[v]);
      const int cur = hs ? in.var_current[v] : 0, des = hs ? in.var_desired[v] : 0, pen = hs ? in.var_pending[v] : 0;
      const bool trans = analysed && ((des != 0 && des != cur) || (cnt != cur));
      if (__any_sync(full, trans)) inTransition = true;
      // scale-up candidate: cheapest without pending, tie -> lower index (:378-395)
      {
        double c = (analysed && pen <= 0) ? in.var_cost[v] : 0.0;
        int idx = (analysed && pen <= 0) ? v : -1;
        for (int o = 16; o; o >>= 1) {
          double oc = shfl_d(full, c, lane ^ o); int oi = __shfl_xor_sync(full, idx, o);
          bool take = oi >= 0 && (idx < 0 || oc < c || (oc == c && oi < idx));
          if (take) { c = oc; idx = oi; }
        }
        if (idx >= 0 && (cheap_v < 0 || c < cheap_c)) { cheap_v = idx; cheap_c = c; }
      }
      // scale-down candidate: most expensive with base target > 1, tie -> higher index (:407-425)
      {
        double c = (analysed && cnt > 1) ? in.var_cost[v] : 0.0;
        int idx = (analysed && cnt > 1) ? v : -1;
        for (int o = 16; o; o >>= 1) {
          double oc = shfl_d(full, c, lane ^ o); int oi = __shfl_xor_sync(full, idx, o);
          bool take = oi >= 0 && (idx < 0 || oc > c || (oc == c && oi > idx));
          if (take) { c = oc; idx = oi; }
        }
        if (idx >= 0 && (exp_v < 0 || c >= exp_c)) { exp_v = idx; exp_c = c; }
      }
    }

    // ---- model level (analyzer.go:96-121, 199-280) ---------------------------------------------
    double avgSpareKv = 0.0, avgSpareQueue = 0.0;
    bool up = false, downSafe = false, kvT = false, qT = false;
    if (totalReplicas > 0) {
      if (nonSaturated > 0) {
        avgSpareKv = d_div(totalSpareKv, (double)nonSaturated);
        avgSpareQueue = d_div(totalSpareQueue, (double)nonSaturated);
      }
      kvT = avgSpareKv < kvTrig;
      qT = avgSpareQueue < qTrig;
      up = kvT || qT;
      if (nonSaturated >= 2) {
        const double avgKvLoad = d_sub(kvThr, avgSpareKv), avgQLoad = d_sub(qThr, avgSpareQueue);
        const double scale = d_div((double)nonSaturated, (double)(nonSaturated - 1));
        const double remKv = d_sub(kvThr, d_mul(avgKvLoad, scale)), remQ = d_sub(qThr, d_mul(avgQLoad, scale));
        downSafe = (remKv >= kvTrig) && (remQ >= qTrig);
      }
    }
    if (lane == 0) {
      if (out.mod_total_replicas) out.mod_total_replicas[m] = totalReplicas;
      if (out.mod_non_saturated) out.mod_non_saturated[m] = nonSaturated;
      if (out.mod_avg_spare_kv) out.mod_avg_spare_kv[m] = avgSpareKv;
      if (out.mod_avg_spare_queue) out.mod_avg_spare_queue[m] = avgSpareQueue;
      if (out.mod_flags)
        out.mod_flags[m] = (up ? SAT_FLAG_UP : 0) | (downSafe ? SAT_FLAG_DOWN : 0) | (inTransition ? SAT_FLAG_TRANS : 0) |
                           (kvT ? SAT_FLAG_KV : 0) | (qT ? SAT_FLAG_Q : 0);
    }
    // ---- targets (analyzer.go:303-436) ------------------------------------------------------------
    int plus_v = -1, minus_v = -1;
    if (nAnalysed > 0 && !inTransition) {
      if (up) plus_v = cheap_v;
      else if (downSafe) minus_v = exp_v;
    }
    if (lane == 0) {
      if (nAnalysed > 0 && inTransition) n_trans++;
      if (plus_v >= 0) n_up++;
      if (minus_v >= 0) n_down++;
    }
    for (int c0 = v0; c0 < v1; c0 += 32) {
      const int v = c0 + lane;
      if (v >= v1) continue;
      const int cnt = in.variant_replica_off[v + 1] - in.variant_replica_off[v];
      const bool hs = !in.var_has_state || in.var_has_state[v];
      const int cur = hs ? in.var_current[v] : 0, des = hs ? in.var_desired[v] : 0;
      int tgt;
      if (nAnalysed == 0) tgt = hs ? in.var_current[v] : -1;          // nil safety :303-309
      else if (cnt == 0) tgt = -1;                                      // not in VariantAnalyses
      else if (inTransition) tgt = (des != 0 && des != cur) ? des : cur;  // :350-359
      else tgt = cnt + (v == plus_v ? 1 : 0) - (v == minus_v ? 1 : 0);   // :362, :399, :428
      if (out.var_target) out.var_target[v] = tgt;
      if (tgt >= 0) sum_targets += tgt;
    }
  }
  if (out.partials) {
    for (int o = 16; o; o >>= 1) sum_targets += __shfl_down_sync(full, sum_targets, o);
    if (lane == 0) {
      if (n_up) atomicAdd((unsigned long long*)&out.partials[0], (unsigned long long)n_up);
      if (n_down) atomicAdd((unsigned long long*)&out.partials[1], (unsigned long long)n_down);
      if (n_trans) atomicAdd((unsigned long long*)&out.partials[2], (unsigned long long)n_trans);
      if (sum_targets) atomicAdd((unsigned long long*)&out.partials[3], (unsigned long long)sum_targets);
    }
  }
}

}  // namespace wva
