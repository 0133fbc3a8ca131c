// saturation_kernel.cuh — V1 saturation capacity model for a batch of models:
// saturation.Analyzer.AnalyzeModelSaturation / analyzeVariant / shouldScaleUp /
// isScaleDownSafe (internal/saturation/analyzer.go:31-280) and
// CalculateSaturationTargets (analyzer.go:290-439).
//
// HBM-bound by design: 16 B per replica (kv float64 + queue int64) + 32 B per variant + 40 B per model, each read once.
// One warp per model, every warp its own software pipeline — no block-level synchronisation anywhere:
//   * the replica stream (the two big arrays, one contiguous range per model thanks to the CSR layout) is moved by the
//     TMA unit: lane 0 issues two `cp.async.bulk` (1-D TMA) copies global -> shared memory per model, completing on the
//     warp's own mbarrier, SAT_NS models ahead of the one being analysed (24 warps x ~2.3 KB in flight per SM);
//   * the per-variant arrays (24 B per variant) are read directly, one lane per variant: perfectly coalesced;
//   * a lane then streams its variant's replicas out of shared memory in slice order (the per-variant float64 sums are
//     order dependent), the variant -> model accumulation runs in ascending variant index (broadcast reads, fixed
//     32-step unrolled chain; lanes without metrics contribute an exact +0.0), and the cheapest / most-expensive
//     variant is a two-word warp arg-min on the order-preserving bit pattern of the cost.
// A model with more replicas than a stage holds (SAT_CAP) takes the same code over global memory instead.
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

// ---- per-warp staging geometry -----------------------------------------------------------------------------------------
constexpr int SAT_WARPS = 8;         // warps (= models in flight) per CTA; three CTAs per SM
constexpr int SAT_NS = 2;            // stages per warp
constexpr int SAT_CAP = 224;         // replicas a stage holds (32 variants x 7 replicas)
constexpr int SAT_VCAP = 32;         // variants a stage holds (one per lane)

struct SatDesc {                     // per model, 48 bytes (written once per batch by saturation_desc_kernel)
  int v0, v1, r0, r1;                // variants [v0, v1), replicas [r0, r1)
  double kvThr, qThr, kvTrig, qTrig; // SaturationScalingConfig of the model
};
struct alignas(16) SatStage {
  double kv[SAT_CAP + 2];            // cp.async.bulk destinations: 16-byte aligned
  long long q[SAT_CAP + 2];
  double cost[SAT_VCAP];             // per-lane columns, filled by each lane's own cp.async
  int lo[SAT_VCAP], cur[SAT_VCAP], des[SAT_VCAP], pen[SAT_VCAP];
  SatDesc desc;
};
struct alignas(16) SatWarpSmem {
  SatStage stage[SAT_NS];
  double termsKv[32], termsQ[32];   // the variants' terms of the two ordered sums, one column each
  unsigned long long bar[SAT_NS];
};
static_assert(sizeof(SatDesc) == 48 && sizeof(SatStage) % 16 == 0 && offsetof(SatStage, q) % 16 == 0 &&
              offsetof(SatStage, cost) % 16 == 0 && offsetof(SatStage, desc) % 16 == 0 && offsetof(SatWarpSmem, termsKv) % 16 == 0 && offsetof(SatWarpSmem, termsQ) % 16 == 0,
              "alignment");
constexpr int SAT_WARPS_PER_SM = 24;   // resident warps per SM: 24 x 9.2 KB of stages = 221 KB of shared memory

// the two dependent CSR look-ups and the config of every model, done once so that the copy-issuing lane never waits on
// a dependent load (48 B per model: 1.4 % of the stream)
__global__ void __launch_bounds__(256) saturation_desc_kernel(SatIn in, SatDesc* desc) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= in.n_models) return;
  SatDesc d;
  d.v0 = in.model_variant_off[m]; d.v1 = in.model_variant_off[m + 1];
  d.r0 = in.variant_replica_off[d.v0]; d.r1 = in.variant_replica_off[d.v1];
  d.kvThr = in.cfg_kv_threshold[m]; d.qThr = in.cfg_queue_threshold[m];
  d.kvTrig = in.cfg_kv_trigger[m]; d.qTrig = in.cfg_queue_trigger[m];
  desc[m] = d;
}

// ---- PTX: mbarrier + 1-D bulk copy ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sat_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sat_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sat_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sat_mbar_expect_tx(unsigned bar_s, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sat_mbar_wait(unsigned bar_s, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar_s), "r"(parity) : "memory");
}
// global -> shared, `bytes` a positive multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void sat_bulk_g2s(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ bool sat_staged(int v0, int v1, int r0, int r1) { return r1 - r0 <= SAT_CAP && v1 - v0 <= SAT_VCAP; }
__device__ __forceinline__ void sat_cp4(unsigned dst_s, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void sat_cp8(unsigned dst_s, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void sat_cp16(unsigned dst_s, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
// Everything one model needs -> a stage (the whole warp calls this).  The two replica arrays — one contiguous range each,
// 1-2 KB — go through the TMA unit: lane 0 issues two cp.async.bulk copies that complete on the stage's mbarrier (range
// start rounded down / length up to 16 bytes; the over-read of < 16 B past an array's end stays inside the input arena,
// whose sub-arrays are 256-byte padded).  The per-variant values are one element per lane: every lane copies its own with
// cp.async (no registers, no scoreboard); lanes 0-2 copy the model's descriptor.  The caller commits the cp.async group.
__device__ __forceinline__ void sat_issue(const SatIn& in, const SatDesc* desc, long long m, int v0, int v1, int r0, int r1,
                                          unsigned st_s, unsigned bar_s) {
  // st_s / bar_s: shared-space addresses of the stage and of its barrier (computed once per kernel, not per copy)
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    const int ra = r0 & ~1;
    const unsigned b_rep = (unsigned)((r1 - ra + 1) & ~1) * 8u;
    sat_mbar_expect_tx(bar_s, 2 * b_rep);
    if (b_rep) {
      sat_bulk_g2s(st_s + (unsigned)offsetof(SatStage, kv), in.rep_kv + ra, b_rep, bar_s);
      sat_bulk_g2s(st_s + (unsigned)offsetof(SatStage, q), in.rep_queue + ra, b_rep, bar_s);
    }
  }
  const int v = v0 + lane;
  if (v < v1) {
    const unsigned l4 = st_s + 4u * lane;
    sat_cp4(l4 + (unsigned)offsetof(SatStage, lo), in.variant_replica_off + v);
    sat_cp4(l4 + (unsigned)offsetof(SatStage, cur), in.var_current + v);
    sat_cp4(l4 + (unsigned)offsetof(SatStage, des), in.var_desired + v);
    sat_cp4(l4 + (unsigned)offsetof(SatStage, pen), in.var_pending + v);
    sat_cp8(st_s + 8u * lane + (unsigned)offsetof(SatStage, cost), in.var_cost + v);
  }
  if (lane < 3) sat_cp16(st_s + (unsigned)offsetof(SatStage, desc) + 16u * lane, reinterpret_cast<const char*>(desc + m) + 16 * lane);
}

// order-preserving bit pattern of a float64 (-0 == +0; NaN sorts after +inf): costs are compared through it
__device__ __forceinline__ unsigned long long sat_sortable(double x) {
  if (x == 0.0) x = 0.0;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct SatTally { int n_up, n_down, n_trans; long long sum_targets; };   // (a warp sees < 2^31 models)

// One model, one warp.  kv / q: the model's replicas, element 0 = replica index `rbase` (a stage, or the global arrays
// with rbase = 0).  STAGED selects the plain per-lane loop (shared memory) or four independent loads in flight (global).
template <bool DETAIL, bool STAGED>
__device__ __forceinline__ void sat_model(const SatIn& in, const double* kvp, const long long* qp, const int rbase,
                                          const long long m, const int v0, const int v1, const SatOut& out,
                                          double* termsKv, double* termsQ, SatTally& tally) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool has_hs = in.var_has_state != nullptr;
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
    int cnt = 0, ns = 0, lo = 0, hi = 0;
    double sumKv = 0.0, sumQ = 0.0, maxKv = 0.0, avgKv = 0.0, avgQ = 0.0;
    long long maxQ = 0;
    // one coalesced round of loads per 32 variants
    const bool hs = act && (!has_hs || in.var_has_state[v]);
    if (act) { lo = in.variant_replica_off[v]; hi = in.variant_replica_off[v + 1]; }
    const int cur = hs ? in.var_current[v] : 0, des = hs ? in.var_desired[v] : 0;
    r_cur = cur; r_des = des; r_pen = hs ? in.var_pending[v] : 0; r_cost = act ? in.var_cost[v] : 0.0;
    cnt = hi - lo;
    if (STAGED) {
      for (int r = lo; r < hi; r++) {
        const double kv = kvp[r - rbase];
        const long long q = qp[r - rbase];
        const double qd = (double)q;
        const bool sat = kv >= kvThr || qd >= qThr;                       // analyzer.go:163-164
        if (DETAIL) { if (out.rep_saturated) out.rep_saturated[r] = sat ? 1 : 0; }
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
    } else {
      for (int base = lo; base < hi; base += 4) {
        double kvv[4]; long long qq[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const bool inb = base + j < hi;
          kvv[j] = inb ? __ldg(kvp + (base + j - rbase)) : 0.0;
          qq[j] = inb ? __ldg(qp + (base + j - rbase)) : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (base + j < hi) {
            const double kv = kvv[j];
            const long long q = qq[j];
            const double qd = (double)q;
            const bool sat = kv >= kvThr || qd >= qThr;
            if (DETAIL) { if (out.rep_saturated) out.rep_saturated[base + j] = sat ? 1 : 0; }
            if (!sat) {
              sumKv = d_add(sumKv, d_sub(kvThr, kv));
              sumQ = d_add(sumQ, d_sub(qThr, qd));
              ns++;
            }
            if (DETAIL) {
              if (kv > maxKv) maxKv = kv;
              if (q > maxQ) maxQ = q;
            }
          }
        }
      }
    }
    if (ns > 0) {                                                          // :190-193, one reciprocal for both quotients
      const double nd = (double)ns;
      if (ns < (1 << 24) && (in_window(sumKv) || sumKv == 0.0) && (in_window(sumQ) || sumQ == 0.0)) {
        const double rr = rcp_f32den((float)ns, nd);
        avgKv = div_f32den(sumKv, nd, rr); avgQ = div_f32den(sumQ, nd, rr);
      } else { avgKv = d_div(sumKv, nd); avgQ = d_div(sumQ, nd); }
    }
    if (DETAIL && act) {
      if (out.var_replica_count) out.var_replica_count[v] = cnt;
      if (out.var_non_saturated) out.var_non_saturated[v] = ns;
      if (out.var_max_kv) out.var_max_kv[v] = maxKv;
      if (out.var_max_queue) out.var_max_queue[v] = maxQ;
      if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = avgKv;
      if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = avgQ;
    }
    const bool analysed = act && cnt > 0;   // only variants with metrics enter VariantAnalyses
    // ordered accumulation over the chunk (analyzer.go:86-94).  A variant without metrics has
    // ns == 0 -> term +0.0, and x + 0.0 == x exactly, so all 32 slots are added unconditionally.
    const double termKv = analysed ? d_mul(avgKv, (double)ns) : 0.0, termQ = analysed ? d_mul(avgQ, (double)ns) : 0.0;
    __syncwarp();
    termsKv[lane] = termKv; termsQ[lane] = termQ;
    __syncwarp();
    {
      // lanes 0-15 run the KV chain, lanes 16-31 the queue chain (each exact and sequential; two terms per 16-byte load)
      const double2* col = reinterpret_cast<const double2*>(lane < 16 ? termsKv : termsQ);
      double acc = (lane < 16) ? totalSpareKv : totalSpareQueue;
#pragma unroll
      for (int l = 0; l < 16; l++) { const double2 t = col[l]; acc = d_add(d_add(acc, t.x), t.y); }
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
  // cheapest without pending, tie -> lower index (:378-395); else most expensive with base target > 1, tie -> higher
  // index (:407-425).  Costs are compared through their order-preserving bit pattern: a two-word warp arg-min / arg-max.
  int plus_v = -1, minus_v = -1;
  const bool stable = nAnalysed > 0 && !inTransition;
  if (stable && (up || downSafe)) {
    const bool want_min = up;
    int best_v = -1;
    unsigned long long best_k = 0;
    for (int c0 = v0; c0 < v1; c0 += 32) {
      const int v = c0 + lane;
      bool cand = false;
      double vcost = 0.0;
      if (v < v1) {
        int cnt, pen;
        if (single) { cnt = r_cnt; pen = r_pen; vcost = r_cost; }
        else {
          cnt = in.variant_replica_off[v + 1] - in.variant_replica_off[v];
          const bool hs2 = !has_hs || in.var_has_state[v];
          pen = hs2 ? in.var_pending[v] : 0;
          vcost = in.var_cost[v];
        }
        cand = cnt > 0 && (want_min ? (pen <= 0) : (cnt > 1));
      }
      const unsigned cm = __ballot_sync(full, cand);
      if (!cm) continue;
      unsigned long long k = sat_sortable(vcost);
      if (!want_min) k = ~k;                                   // arg-max as arg-min of the complement
      const unsigned khi = cand ? (unsigned)(k >> 32) : 0xffffffffu, klo = cand ? (unsigned)k : 0xffffffffu;
      const unsigned mh = __reduce_min_sync(full, khi);
      bool in_ = cand && khi == mh;
      const unsigned ml = __reduce_min_sync(full, in_ ? klo : 0xffffffffu);
      in_ = in_ && klo == ml;
      const unsigned wm = __ballot_sync(full, in_);
      // all-ones keys of non-candidates can only tie with a candidate whose key is all ones too; `in_` requires cand
      const int wl = want_min ? (__ffs(wm) - 1) : (31 - __clz(wm));
      const unsigned long long kk = ((unsigned long long)mh << 32) | ml;
      // across chunks of 32: strictly better replaces; on ties the lower index stays (min) / the higher replaces (max)
      if (best_v < 0 || kk < best_k || (!want_min && kk == best_k)) { best_v = c0 + wl; best_k = kk; }
    }
    if (want_min) plus_v = best_v; else minus_v = best_v;
  }
  if (lane == 0) {
    if (nAnalysed > 0 && inTransition) tally.n_trans++;
    if (plus_v >= 0) tally.n_up++;
    if (minus_v >= 0) tally.n_down++;
  }
  // ---- targets (analyzer.go:303-436) ------------------------------------------------------------
  for (int c0 = v0; c0 < v1; c0 += 32) {
    const int v = c0 + lane;
    if (v >= v1) continue;
    const int cnt = single ? r_cnt : in.variant_replica_off[v + 1] - in.variant_replica_off[v];
    const bool hs = !has_hs || in.var_has_state[v];
    int tgt;
    if (nAnalysed == 0) tgt = hs ? (single ? r_cur : in.var_current[v]) : -1;   // nil safety :303-309
    else if (cnt == 0) tgt = -1;                                      // not in VariantAnalyses
    else if (inTransition) {                                          // :350-359
      const int cur = single ? r_cur : (hs ? in.var_current[v] : 0), des = single ? r_des : (hs ? in.var_desired[v] : 0);
      tgt = (des != 0 && des != cur) ? des : cur;
    } else tgt = cnt + (v == plus_v ? 1 : 0) - (v == minus_v ? 1 : 0);   // :362, :399, :428
    if (out.var_target) out.var_target[v] = tgt;
    if (tgt >= 0) tally.sum_targets += tgt;
  }
}

// One model out of a stage (at most 32 variants, one per lane; everything in shared memory).  Same arithmetic and
// order as sat_model, without its chunk loops.
template <bool DETAIL>
__device__ __forceinline__ void sat_model_staged(const SatStage* st, const unsigned char* __restrict__ hs_col, const long long m,
                                                 const SatOut& out, double* termsKv, double* termsQ, SatTally& tally) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int v0 = st->desc.v0, v1 = st->desc.v1;
  const double kvThr = st->desc.kvThr, qThr = st->desc.qThr, kvTrig = st->desc.kvTrig, qTrig = st->desc.qTrig;
  const int rbase = st->desc.r0 & ~1;
  const int v = v0 + lane;
  const bool act = v < v1;
  int lo = st->desc.r1, cur = 0, des = 0, pen = 0;
  double cost = 0.0;
  bool hs = false;
  if (act) {
    lo = st->lo[lane];
    hs = !hs_col || hs_col[v];
    if (hs) { cur = st->cur[lane]; des = st->des[lane]; pen = st->pen[lane]; }
    cost = st->cost[lane];
  }
  // a variant's range ends where the next one starts; the model's last one ends at r1 (inactive lanes hold r1: empty)
  const int hi_n = __shfl_down_sync(full, lo, 1);
  const int hi = (lane == 31) ? st->desc.r1 : hi_n;
  const int cnt = hi - lo;
  int ns = 0;
  double sumKv = 0.0, sumQ = 0.0, maxKv = 0.0, avgKv = 0.0, avgQ = 0.0;
  long long maxQ = 0;
  for (int r = lo; r < hi; r++) {
    const double kv = st->kv[r - rbase];
    const long long q = st->q[r - rbase];
    const double qd = (double)q;
    const bool sat = kv >= kvThr || qd >= qThr;                       // analyzer.go:163-164
    if (DETAIL) { if (out.rep_saturated) out.rep_saturated[r] = sat ? 1 : 0; }
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
  if (ns > 0) {                                                        // :190-193, one reciprocal for both quotients
    const double nd = (double)ns;
    if ((in_window(sumKv) || sumKv == 0.0) && (in_window(sumQ) || sumQ == 0.0)) {
      const double rr = rcp_f32den((float)ns, nd);
      avgKv = div_f32den(sumKv, nd, rr); avgQ = div_f32den(sumQ, nd, rr);
    } else { avgKv = d_div(sumKv, nd); avgQ = d_div(sumQ, nd); }
  }
  if (DETAIL && act) {
    if (out.var_replica_count) out.var_replica_count[v] = cnt;
    if (out.var_non_saturated) out.var_non_saturated[v] = ns;
    if (out.var_max_kv) out.var_max_kv[v] = maxKv;
    if (out.var_max_queue) out.var_max_queue[v] = maxQ;
    if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = avgKv;
    if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = avgQ;
  }
  const bool analysed = cnt > 0;           // (inactive lanes have cnt == 0)
  // ordered accumulation (analyzer.go:86-94): a variant without metrics contributes an exact +0.0
  __syncwarp();
  termsKv[lane] = d_mul(avgKv, (double)ns); termsQ[lane] = d_mul(avgQ, (double)ns);
  __syncwarp();
  double totalSpareKv, totalSpareQueue;
  {
    const double2* col = reinterpret_cast<const double2*>(lane < 16 ? termsKv : termsQ);
    double acc = 0.0;
#pragma unroll
    for (int l = 0; l < 16; l++) { const double2 t = col[l]; acc = d_add(d_add(acc, t.x), t.y); }
    const double other = shfl_xor_d(full, acc, 16);
    totalSpareKv = (lane < 16) ? acc : other;
    totalSpareQueue = (lane < 16) ? other : acc;
  }
  const int nonSaturated = __reduce_add_sync(full, ns);
  const int totalReplicas = __reduce_add_sync(full, cnt);
  const unsigned anm = __ballot_sync(full, analysed);
  const bool inTransition = __any_sync(full, analysed && ((des != 0 && des != cur) || (cnt != cur)));   // :322-341

  // ---- model level (analyzer.go:96-121, 199-280) ---------------------------------------------
  double avgSpareKv = 0.0, avgSpareQueue = 0.0;
  bool up = false, downSafe = false, kvT = false, qT = false;
  if (totalReplicas > 0) {
    if (nonSaturated > 0) {
      const double nd = (double)nonSaturated;
      if ((in_window(totalSpareKv) || totalSpareKv == 0.0) && (in_window(totalSpareQueue) || totalSpareQueue == 0.0)) {
        const double rr = rcp_f32den((float)nonSaturated, nd);
        avgSpareKv = div_f32den(totalSpareKv, nd, rr); avgSpareQueue = div_f32den(totalSpareQueue, nd, rr);
      } else { avgSpareKv = d_div(totalSpareKv, nd); avgSpareQueue = d_div(totalSpareQueue, nd); }
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
  // ---- scaling candidate (analyzer.go:376-433): two-word warp arg-min on the order-preserving cost bits ------------
  int plus_l = -1, minus_l = -1;
  if (anm && !inTransition && (up || downSafe)) {
    const bool want_min = up;
    const bool cand = cnt > 0 && (want_min ? (pen <= 0) : (cnt > 1));
    const unsigned cm = __ballot_sync(full, cand);
    if (cm) {
      unsigned long long k = sat_sortable(cost);
      if (!want_min) k = ~k;
      const unsigned khi = cand ? (unsigned)(k >> 32) : 0xffffffffu, klo = cand ? (unsigned)k : 0xffffffffu;
      const unsigned mh = __reduce_min_sync(full, khi);
      bool in_ = cand && khi == mh;
      const unsigned ml = __reduce_min_sync(full, in_ ? klo : 0xffffffffu);
      in_ = in_ && klo == ml;
      const unsigned wm = __ballot_sync(full, in_);
      const int wl = want_min ? (__ffs(wm) - 1) : (31 - __clz(wm));
      if (want_min) plus_l = wl; else minus_l = wl;
    }
  }
  if (lane == 0) {
    if (anm && inTransition) tally.n_trans++;
    if (plus_l >= 0) tally.n_up++;
    if (minus_l >= 0) tally.n_down++;
  }
  // ---- targets (analyzer.go:303-436) ------------------------------------------------------------
  if (act) {
    int tgt;
    if (!anm) tgt = hs ? cur : -1;                                              // nil safety :303-309
    else if (cnt == 0) tgt = -1;                                                // not in VariantAnalyses
    else if (inTransition) tgt = (des != 0 && des != cur) ? des : cur;          // :350-359
    else tgt = cnt + (lane == plus_l ? 1 : 0) - (lane == minus_l ? 1 : 0);      // :362, :399, :428
    if (out.var_target) out.var_target[v] = tgt;
    if (tgt >= 0) tally.sum_targets += tgt;
  }
}

template <bool DETAIL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, SAT_WARPS_PER_SM / WARPS) saturation_kernel(SatIn in, SatOut out, const SatDesc* __restrict__ desc) {
  extern __shared__ __align__(16) unsigned char sat_smem[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SatWarpSmem* ws = reinterpret_cast<SatWarpSmem*>(sat_smem) + warp;
  SatTally tally = {0, 0, 0, 0};
  // model indices fit 32 bits (the CSR offsets are int32; the host rejects larger batches)
  const int gw = blockIdx.x * WARPS + warp, tw = gridDim.x * WARPS;
  const int M = (int)in.n_models;

  const unsigned ws_s = sat_smem_addr(ws);       // shared-space address of this warp's area
  if (lane == 0) {
    for (int s = 0; s < SAT_NS; s++) sat_mbar_init(&ws->bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  static_assert(SAT_NS == 2, "stage = trip parity");
  // Geometry (16 bytes: variants [x, y), replicas [z, w)) is only needed where copies are issued; the model analysed
  // reads its own from the stage.  `gn` = the model whose copies are issued at the end of this trip (SAT_NS models ahead),
  // loaded one trip earlier so that no descriptor load is consumed in the trip that issues it.  `staged` / `par`: one bit
  // per stage — the stage holds a staged model / phase parity of the next wait on its barrier.  All of it lives in
  // registers (no lambda, nothing by reference).
  const int4 none = make_int4(0, 0, 0, 0);
  unsigned staged = 0, par = 0;
  {
    const int4 g0 = gw < M ? __ldg(reinterpret_cast<const int4*>(desc + gw)) : none;
    const int4 g1 = (long long)gw + tw < M ? __ldg(reinterpret_cast<const int4*>(desc + gw + tw)) : none;
    // prologue: the first two models; one cp.async group per model, empty or not
    if (gw < M && sat_staged(g0.x, g0.y, g0.z, g0.w)) { sat_issue(in, desc, gw, g0.x, g0.y, g0.z, g0.w, ws_s, ws_s + (unsigned)offsetof(SatWarpSmem, bar)); staged |= 1u; }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if ((long long)gw + tw < M && sat_staged(g1.x, g1.y, g1.z, g1.w)) { sat_issue(in, desc, gw + tw, g1.x, g1.y, g1.z, g1.w, ws_s + (unsigned)sizeof(SatStage), ws_s + (unsigned)offsetof(SatWarpSmem, bar) + 8u); staged |= 2u; }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int4 gn = (long long)gw + 2LL * tw < M ? __ldg(reinterpret_cast<const int4*>(desc + gw + 2LL * tw)) : none;
  int S = 0;
  // (m + 3 tw can pass 2^31 only in the last trips: the look-ahead indices are compared in 64 bits, m itself stays int)
  for (int m = gw; m < M; S ^= 1) {
    SatStage* st = &ws->stage[S];
    const unsigned st_s = ws_s + (unsigned)S * (unsigned)sizeof(SatStage), bar_s = ws_s + (unsigned)offsetof(SatWarpSmem, bar) + 8u * S;
    const long long mf = (long long)m + 2LL * tw;                            // the model that takes this stage next
    const int4 g_after = mf + tw < M ? __ldg(reinterpret_cast<const int4*>(desc + mf + tw)) : none;
    asm volatile("cp.async.wait_group %0;" ::"n"(SAT_NS - 1) : "memory");   // this model's group is the oldest pending one
    if ((staged >> S) & 1u) {
      sat_mbar_wait(bar_s, (par >> S) & 1u);                                   // a model that is not staged never arms its barrier
      par ^= 1u << S;
      __syncwarp();                                                          // the other lanes' cp.async data
      sat_model_staged<DETAIL>(st, in.var_has_state, m, out, ws->termsKv, ws->termsQ, tally);
    } else {
      const int4 g = __ldg(reinterpret_cast<const int4*>(desc + m));
      sat_model<DETAIL, false>(in, in.rep_kv, in.rep_queue, 0, m, g.x, g.y, out, ws->termsKv, ws->termsQ, tally);
    }
    __syncwarp();                                                            // every lane is done with the stage before it is refilled
    staged &= ~(1u << S);
    if (mf < M && sat_staged(gn.x, gn.y, gn.z, gn.w)) { sat_issue(in, desc, mf, gn.x, gn.y, gn.z, gn.w, st_s, bar_s); staged |= 1u << S; }
    asm volatile("cp.async.commit_group;" ::: "memory");
    gn = g_after;
    if ((long long)m + tw >= M) break;
    m += tw;
  }
  if (out.partials) {
    long long sum_targets = tally.sum_targets;
    for (int o = 16; o; o >>= 1) sum_targets += __shfl_down_sync(full, sum_targets, o);
    if (lane == 0) {
      if (tally.n_up) atomicAdd((unsigned long long*)&out.partials[0], (unsigned long long)tally.n_up);
      if (tally.n_down) atomicAdd((unsigned long long*)&out.partials[1], (unsigned long long)tally.n_down);
      if (tally.n_trans) atomicAdd((unsigned long long*)&out.partials[2], (unsigned long long)tally.n_trans);
      if (sum_targets) atomicAdd((unsigned long long*)&out.partials[3], (unsigned long long)sum_targets);
    }
  }
}

// persistent launch: SAT_WARPS_PER_SM warps per SM in 256-thread blocks, each warp its own two-stage pipeline
// (one warp per block — block-uniform addresses — was tried and is slower: 1.18 ms vs 0.85 ms on configs[3])
static inline cudaError_t launch_saturation(bool detail, int sm_count, long long M, const SatIn& vin, const SatOut& w, const SatDesc* d_desc,
                                            cudaStream_t stream) {
  long long blocks = (long long)sm_count * (SAT_WARPS_PER_SM / SAT_WARPS);
  const long long need = (M + SAT_WARPS - 1) / SAT_WARPS;
  if (blocks > need) blocks = need;
  const size_t smem = sizeof(SatWarpSmem) * SAT_WARPS;
  auto k = detail ? saturation_kernel<true, SAT_WARPS> : saturation_kernel<false, SAT_WARPS>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<(unsigned)blocks, SAT_WARPS * 32, smem, stream>>>(vin, w, d_desc);
  return cudaGetLastError();
}

}  // namespace wva
