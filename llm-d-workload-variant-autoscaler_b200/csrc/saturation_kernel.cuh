// saturation_kernel.cuh — V1 saturation capacity model for a batch of models:
// saturation.Analyzer.AnalyzeModelSaturation / analyzeVariant / shouldScaleUp /
// isScaleDownSafe (internal/saturation/analyzer.go:31-280) and
// CalculateSaturationTargets (analyzer.go:290-439).
//
// One warp per model, one lane per variant (chunks of 32).  A lane streams its
// variant's replicas in slice order, 4 independent loads at a time (the per-variant
// float64 sums are order dependent); the variant -> model accumulation runs in
// ascending variant index through shared memory (broadcast reads, fixed 32-step
// unrolled chain; lanes without metrics contribute an exact +0.0).  The cheapest /
// most-expensive variant search runs only for the models that scale.  HBM-bound by
// design: 16 B per replica (kv float64 + queue int64) is the only large stream.
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

__device__ __forceinline__ double shfl_xor_d(unsigned mask, double v, int lanemask) {
  int lo = __shfl_xor_sync(mask, __double2loint(v), lanemask), hi = __shfl_xor_sync(mask, __double2hiint(v), lanemask);
  return __hiloint2double(hi, lo);
}

// x / n for an integer 0 < n < 2^24, correctly rounded: n is float32-valued, so (E1) of wva_core.cuh applies
// (3 FP64 ops + the shared reciprocal instead of the ~20-instruction div.rn.f64 sequence)
__device__ __forceinline__ double div_small_int(double x, int n) {
  const float nf = (float)n;
  const double nd = (double)n;
  if (n < (1 << 24) && (in_window(x) || x == 0.0)) return div_f32den(x, nd, rcp_f32den(nf, nd));
  return d_div(x, nd);
}

template <bool DETAIL>
// 4 blocks per SM = the kernel's natural 64 registers; 5 / 6 blocks (48 / 40 registers, spills) measured 0.230 / 0.267 ms
// against 0.219 ms on 28.8 M replicas
#ifndef SAT_MINB
#define SAT_MINB 4
#endif
__global__ void __launch_bounds__(256, SAT_MINB) saturation_kernel(SatIn in, SatOut out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  long long n_up = 0, n_down = 0, n_trans = 0, sum_targets = 0;
  const double* __restrict__ in_kv = in.rep_kv;
  const long long* __restrict__ in_q = in.rep_queue;
  const int* __restrict__ vro = in.variant_replica_off;
  __shared__ double2 terms[8][32];
  double2* my_terms = terms[threadIdx.x >> 5];

  for (long long m = warp0; m < in.n_models; m += nwarps) {
    const int v0 = in.model_variant_off[m], v1 = in.model_variant_off[m + 1];
    const double kvThr = in.cfg_kv_threshold[m], qThr = in.cfg_queue_threshold[m];
    const double kvTrig = in.cfg_kv_trigger[m], qTrig = in.cfg_queue_trigger[m];
    double totalSpareKv = 0.0, totalSpareQueue = 0.0;
    int nonSaturated = 0, totalReplicas = 0, nAnalysed = 0;
    bool inTransition = false;
    const bool single = v1 - v0 <= 32;          // the usual case: the lane's variant data stays in registers
    int r_cur = 0, r_des = 0, r_pen = 0, r_cnt = 0;
    double r_cost = 0.0;

    // ---- phase A: analyzeVariant per lane, ordered combine ---------------------------------
    for (int c0 = v0; c0 < v1; c0 += 32) {
      const int v = c0 + lane;
      const bool act = v < v1;
      int cnt = 0, ns = 0;
      double sumKv = 0.0, sumQ = 0.0, maxKv = 0.0, avgKv = 0.0, avgQ = 0.0;
      long long maxQ = 0;
      // the variant's state is requested NOW, together with its replica range, so these loads overlap the
      // replica stream instead of adding two more dependent memory round trips per model
      const bool hs = act && (!in.var_has_state || in.var_has_state[v]);
      const int cur = hs ? in.var_current[v] : 0, des = hs ? in.var_desired[v] : 0;
      r_cur = cur; r_des = des; r_pen = hs ? in.var_pending[v] : 0; r_cost = act ? in.var_cost[v] : 0.0;
      if (act) {
        const int lo = vro[v], hi = vro[v + 1];
        cnt = hi - lo;
        for (int base = lo; base < hi; base += 4) {
          double kvv[4]; long long qq[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const bool inb = base + j < hi;
            kvv[j] = inb ? __ldg(in_kv + base + j) : 0.0;
            qq[j] = inb ? __ldg(in_q + base + j) : 0;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (base + j < hi) {
              const double kv = kvv[j];
              const long long q = qq[j];
              const double qd = (double)q;
              const bool sat = kv >= kvThr || qd >= qThr;                       // analyzer.go:163-164
              if (DETAIL) { if (out.rep_saturated) out.rep_saturated[base + j] = sat ? 1 : 0; }
              if (!sat) {
                sumKv = d_add(sumKv, d_sub(kvThr, kv));                        // :170-175
                sumQ = d_add(sumQ, d_sub(qThr, qd));
                ns++;
              }
              if (DETAIL) {
                if (kv > maxKv) maxKv = kv;                                    // :179-184
                if (q > maxQ) maxQ = q;
              }
            }
          }
        }
        if (ns > 0) { avgKv = div_small_int(sumKv, ns); avgQ = div_small_int(sumQ, ns); }        // :190-193
        if (DETAIL) {
          if (out.var_replica_count) out.var_replica_count[v] = cnt;
          if (out.var_non_saturated) out.var_non_saturated[v] = ns;
          if (out.var_max_kv) out.var_max_kv[v] = maxKv;
          if (out.var_max_queue) out.var_max_queue[v] = maxQ;
          if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = avgKv;
          if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = avgQ;
        }
      }
      const bool analysed = act && cnt > 0;   // only variants with metrics enter VariantAnalyses
      // ordered accumulation over the chunk (analyzer.go:86-94).  A variant without metrics has
      // ns == 0 -> term +0.0, and x + 0.0 == x exactly, so all 32 slots are added unconditionally.
      const double termKv = analysed ? d_mul(avgKv, (double)ns) : 0.0, termQ = analysed ? d_mul(avgQ, (double)ns) : 0.0;
      __syncwarp();
      my_terms[lane] = make_double2(termKv, termQ);
      __syncwarp();
      {
        // lanes 0-15 run the KV chain, lanes 16-31 the queue chain (each exact and sequential)
        const double* col = reinterpret_cast<const double*>(my_terms) + (lane >> 4);
        double acc = (lane < 16) ? totalSpareKv : totalSpareQueue;
#pragma unroll
        for (int l = 0; l < 32; l++) acc = d_add(acc, col[2 * l]);
        const double other = shfl_xor_d(full, acc, 16);
        totalSpareKv = (lane < 16) ? acc : other;
        totalSpareQueue = (lane < 16) ? other : acc;
      }
      nonSaturated += __reduce_add_sync(full, analysed ? ns : 0);
      totalReplicas += __reduce_add_sync(full, act ? cnt : 0);
      nAnalysed += __popc(__ballot_sync(full, analysed));
      // transition checks (analyzer.go:322-341); a variant without state reads the zero value
      r_cnt = cnt;
      const bool trans = analysed && ((des != 0 && des != cur) || (cnt != cur));
      if (__any_sync(full, trans)) inTransition = true;
    }

    // ---- model level (analyzer.go:96-121, 199-280) ---------------------------------------------
    double avgSpareKv = 0.0, avgSpareQueue = 0.0;
    bool up = false, downSafe = false, kvT = false, qT = false;
    if (totalReplicas > 0) {
      if (nonSaturated > 0) {
        avgSpareKv = div_small_int(totalSpareKv, nonSaturated);
        avgSpareQueue = div_small_int(totalSpareQueue, nonSaturated);
      }
      kvT = avgSpareKv < kvTrig;
      qT = avgSpareQueue < qTrig;
      up = kvT || qT;
      if (nonSaturated >= 2) {
        const double avgKvLoad = d_sub(kvThr, avgSpareKv), avgQLoad = d_sub(qThr, avgSpareQueue);
        const double scale = div_small_int((double)nonSaturated, nonSaturated - 1);
        const double remKv = d_sub(kvThr, d_mul(avgKvLoad, scale)), remQ = d_sub(qThr, d_mul(avgQLoad, scale));
        downSafe = (remKv >= kvTrig) && (remQ >= qTrig);
      }
    }
    if (lane == 0) {
      if (DETAIL) {
        if (out.mod_total_replicas) out.mod_total_replicas[m] = totalReplicas;
        if (out.mod_non_saturated) out.mod_non_saturated[m] = nonSaturated;
        if (out.mod_avg_spare_kv) out.mod_avg_spare_kv[m] = avgSpareKv;
        if (out.mod_avg_spare_queue) out.mod_avg_spare_queue[m] = avgSpareQueue;
      }
      if (out.mod_flags)
        out.mod_flags[m] = (up ? SAT_FLAG_UP : 0) | (downSafe ? SAT_FLAG_DOWN : 0) | (inTransition ? SAT_FLAG_TRANS : 0) |
                           (kvT ? SAT_FLAG_KV : 0) | (qT ? SAT_FLAG_Q : 0);
    }
    // ---- scaling candidate, only for the models that scale (analyzer.go:376-433) -------------------
    int plus_v = -1, minus_v = -1;
    const bool stable = nAnalysed > 0 && !inTransition;
    if (stable && (up || downSafe)) {
      const bool want_min = up;    // cheapest without pending, tie -> lower index (:378-395)
                                   // else most expensive with base target > 1, tie -> higher index (:407-425)
      int best_v = -1;
      double best_c = 0.0;
      for (int c0 = v0; c0 < v1; c0 += 32) {
        const int v = c0 + lane;
        bool cand = false;
        double vcost = 0.0;
        if (v < v1) {
          int cnt, pen;
          if (single) { cnt = r_cnt; pen = r_pen; vcost = r_cost; }
          else {
            cnt = vro[v + 1] - vro[v];
            const bool hs2 = !in.var_has_state || in.var_has_state[v];
            pen = hs2 ? in.var_pending[v] : 0;
            vcost = in.var_cost[v];
          }
          cand = cnt > 0 && (want_min ? (pen <= 0) : (cnt > 1));
        }
        double c = cand ? vcost : 0.0;
        int idx = cand ? v : -1;
        for (int o = 16; o; o >>= 1) {
          const double oc = shfl_xor_d(full, c, o);
          const int oi = __shfl_xor_sync(full, idx, o);
          const bool take = oi >= 0 && (idx < 0 || (want_min ? (oc < c || (oc == c && oi < idx))
                                                              : (oc > c || (oc == c && oi > idx))));
          if (take) { c = oc; idx = oi; }
        }
        if (idx >= 0 && (best_v < 0 || (want_min ? (c < best_c) : (c >= best_c)))) { best_v = idx; best_c = c; }
      }
      if (want_min) plus_v = best_v; else minus_v = best_v;
    }
    if (lane == 0) {
      if (nAnalysed > 0 && inTransition) n_trans++;
      if (plus_v >= 0) n_up++;
      if (minus_v >= 0) n_down++;
    }
    // ---- targets (analyzer.go:303-436) ------------------------------------------------------------
    for (int c0 = v0; c0 < v1; c0 += 32) {
      const int v = c0 + lane;
      if (v >= v1) continue;
      const int cnt = single ? r_cnt : vro[v + 1] - vro[v];
      const bool hs = !in.var_has_state || in.var_has_state[v];
      int tgt;
      if (nAnalysed == 0) tgt = hs ? (single ? r_cur : in.var_current[v]) : -1;   // nil safety :303-309
      else if (cnt == 0) tgt = -1;                                      // not in VariantAnalyses
      else if (inTransition) {                                          // :350-359
        const int cur = single ? r_cur : (hs ? in.var_current[v] : 0), des = single ? r_des : (hs ? in.var_desired[v] : 0);
        tgt = (des != 0 && des != cur) ? des : cur;
      } else tgt = cnt + (v == plus_v ? 1 : 0) - (v == minus_v ? 1 : 0);   // :362, :399, :428
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
