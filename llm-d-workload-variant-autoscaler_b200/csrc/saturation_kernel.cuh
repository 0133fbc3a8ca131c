// saturation_kernel.cuh — V1 saturation capacity model for a batch of models:
// saturation.Analyzer.AnalyzeModelSaturation / analyzeVariant / shouldScaleUp /
// isScaleDownSafe (internal/saturation/analyzer.go:31-280) and
// CalculateSaturationTargets (analyzer.go:290-439).
//
// HBM-bound by design: 16 B per replica (kv float64 + queue int64) + 32 B per variant + 40 B per model, each read once.
// One warp per GROUP of G = 2 consecutive models, every warp its own software pipeline over its groups — no block-level
// synchronisation anywhere (one persistent 768-thread block per SM, 24 warps x 8.3 KB of shared memory):
//   * consecutive models are contiguous in the CSR layout, so a group's replicas are ONE range of each replica array: lane
//     0 issues two `cp.async.bulk` (1-D TMA) copies global -> shared memory per group (2 x ~2.3 KB), completing on the
//     warp's own mbarrier; the per-variant columns (24 B per variant) follow as 16-byte `cp.async` chunks, one per lane and
//     column, and the models' 48-byte descriptors go to a slot selected by trip parity;
//   * a lane owns variant l of each model of the group — G independent dependency chains — and streams their replicas out
//     of shared memory in slice order (the per-variant float64 sums are order dependent) in a branch-free, warp-uniform
//     loop of four slots per round;
//   * as soon as that loop is over everything the trip still needs is in registers, and the warp issues the copies of its
//     NEXT group into the same stage: they land while the model-level part of this group runs (a single stage per warp —
//     twice the warps of a double-buffered design in the same shared memory);
//   * the variant -> model accumulation runs in ascending variant index as 2G chains (KV and queue of every model) in
//     disjoint lane groups of one fixed 32-step pass (lanes without metrics contribute an exact +0.0); the model-level
//     divisions are done once, every lane for its own model; the cheapest / most-expensive variant is a warp arg-min on
//     the order-preserving bit pattern of the cost.
// A group that does not fit a stage (more than CAP_R replicas, a model of more than 32 variants, a variant of more than 64
// replicas) takes the general path — the same arithmetic over global memory — in a second pass of the same kernel.
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
// A warp analyses a GROUP of G consecutive models per trip.  Consecutive models are contiguous in the CSR layout, so the
// group's replicas are ONE range of each replica array (two bulk copies per group instead of per model) and its variants
// one range of every per-variant column (16-byte cp.async chunks, a handful per lane per group).  The G models are then
// analysed side by side — lane l owns variant l of each of them: G independent dependency chains per lane — and their
// ordered model-level sums run as 2G chains in disjoint lane groups of one 32-step pass.
constexpr int SAT_VCAP = 32;         // variants of one model a stage holds (one per lane)
constexpr int SAT_MAXCNT = 64;       // replicas of one variant the staged path takes
#ifndef SAT_UNROLL
#define SAT_UNROLL 4                 // replica slots per round of the (warp-uniform) replica loop
#endif
template <int G> struct SatCfg;
// replicas a stage holds
template <> struct SatCfg<1> { static constexpr int CAP_R = 224; };
template <> struct SatCfg<2> { static constexpr int CAP_R = 352; };
template <> struct SatCfg<4> { static constexpr int CAP_R = 688; };

struct SatDesc {                     // per model, 48 bytes (written once per batch by saturation_desc_kernel)
  int v0, v1, r0, r1;                // variants [v0, v1), replicas [r0, r1)
  double kvThr, qThr, kvTrig, qTrig; // SaturationScalingConfig of the model
};
template <int G> struct alignas(16) SatStage {
  static constexpr int CAP_R = SatCfg<G>::CAP_R, CAP_V = 32 * G + 4;   // variant slots: range start rounded down to a multiple of 4, one extra offset
  double kv[CAP_R + 2];              // cp.async.bulk destinations: 16-byte aligned (range start rounded down to an even index)
  long long q[CAP_R + 2];
  double cost[CAP_V];                // per-variant columns, filled by 16-byte cp.async chunks
  int lo[CAP_V], cur[CAP_V], des[CAP_V], pen[CAP_V];
};
// ONE stage per warp: everything a trip needs from it is in registers once the replica loop is over, and the copies of the
// warp's next group are issued right there — they land while the model-level part of this group runs.
template <int G> struct alignas(16) SatWarpSmem {
  SatStage<G> stage;
  double terms[2 * G * 32];          // the variants' terms of the 2G ordered sums, one column each
  SatDesc desc[2][G];                // the models' descriptors, by trip parity: they are read until the end of the trip
  unsigned long long bar;
  unsigned long long pad_;
};
static_assert(sizeof(SatDesc) == 48, "layout");

// group geometry: {V0 (or a negative value: a model of the group has more than 32 variants), V1, R0, R1} per group of G models
template <int G> __host__ __device__ inline long long sat_groups(long long M) { return (M + G - 1) / G; }
// descriptor scratch: M descriptors, then (256-byte aligned) the group geometry of the largest group count (G = 1)
static inline size_t sat_desc_bytes(long long M) { return (((size_t)M * sizeof(SatDesc) + 255) & ~(size_t)255) + (size_t)M * 16 + 256; }
static inline int4* sat_geo_ptr(SatDesc* desc, long long M) {
  return reinterpret_cast<int4*>(reinterpret_cast<char*>(desc) + (((size_t)M * sizeof(SatDesc) + 255) & ~(size_t)255));
}

// the dependent CSR look-ups and the config of every model, done once so that the copy-issuing lane never waits on a
// dependent load (48 B per model + 16 B per group: 1.4 % of the stream)
template <int G>
__global__ void __launch_bounds__(256) saturation_desc_kernel(SatIn in, SatDesc* desc, int4* geo) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= in.n_models) return;
  SatDesc d;
  d.v0 = in.model_variant_off[m]; d.v1 = in.model_variant_off[m + 1];
  d.r0 = in.variant_replica_off[d.v0]; d.r1 = in.variant_replica_off[d.v1];
  d.kvThr = in.cfg_kv_threshold[m]; d.qThr = in.cfg_queue_threshold[m];
  d.kvTrig = in.cfg_kv_trigger[m]; d.qTrig = in.cfg_queue_trigger[m];
  desc[m] = d;
  if (m % G == 0) {
    const long long me = m + G < in.n_models ? m + G : in.n_models;
    bool ok = d.v1 - d.v0 <= SAT_VCAP;
    int vprev = d.v1;
    for (long long mm = m + 1; mm < me; mm++) { const int vn = in.model_variant_off[mm + 1]; ok = ok && vn - vprev <= SAT_VCAP; vprev = vn; }
    geo[m / G] = make_int4(ok ? d.v0 : -1 - d.v0, vprev, d.r0, in.variant_replica_off[vprev]);
  }
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
template <int G> __device__ __forceinline__ bool sat_staged(const int4 g) {   // g = the group's geometry (x < 0: not staged)
  return g.x >= 0 && g.w - (g.z & ~1) <= SatCfg<G>::CAP_R;
}
__device__ __forceinline__ void sat_cp16(unsigned dst_s, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
// Everything one group of models needs -> a stage (the whole warp calls this).  The two replica arrays — one contiguous
// range each, G x 1-2 KB — go through the TMA unit: lane 0 issues two cp.async.bulk copies that complete on the stage's
// mbarrier (range start rounded down / length up to 16 bytes; an over-read of < 16 B past an array's end stays inside the
// input arena, whose sub-arrays are 256-byte padded).  The per-variant columns are copied in 16-byte chunks from the
// range start rounded down to a multiple of 4 variants (the offsets column one entry further: a variant's range ends where
// the next one starts), one cp.async per lane and column; 3 x G lanes copy the models' descriptors.  No registers, no
// scoreboard.  The caller commits the cp.async group.
template <int G>
__device__ __forceinline__ void sat_issue(const SatIn& in, const SatDesc* desc, const int lane, const int m0, const int ng, const int4 g,
                                          unsigned st_s, unsigned desc_s, unsigned bar_s) {
  // st_s / desc_s / bar_s: shared-space addresses of the stage, of the descriptor slot and of the barrier
  using Stage = SatStage<G>;
  if (lane == 0) {
    const int ra = g.z & ~1;
    const unsigned b_rep = (unsigned)((g.w - ra + 1) & ~1) * 8u;
    sat_mbar_expect_tx(bar_s, 2 * b_rep);
    if (b_rep) {
      sat_bulk_g2s(st_s + (unsigned)offsetof(Stage, kv), in.rep_kv + ra, b_rep, bar_s);
      sat_bulk_g2s(st_s + (unsigned)offsetof(Stage, q), in.rep_queue + ra, b_rep, bar_s);
    }
  }
  const int va = g.x & ~3;
  const int n4 = (g.y - va + 4) >> 2;              // 16-byte chunks of an int column holding entries [va, V1]
  const int n2 = (g.y - va + 1) >> 1;              // 16-byte chunks of the float64 cost column
  const size_t e4 = (size_t)(unsigned)va * 4u + 16u * lane;   // byte offset of this lane's chunk in an int column
  const unsigned d = st_s + 16u * lane;
#pragma unroll
  for (int k = 0; k < (Stage::CAP_V / 4 + 31) / 32; k++) {
    if (lane + 32 * k < n4) {
      sat_cp16(d + 512u * k + (unsigned)offsetof(Stage, lo), reinterpret_cast<const char*>(in.variant_replica_off) + e4 + 512u * k);
      sat_cp16(d + 512u * k + (unsigned)offsetof(Stage, cur), reinterpret_cast<const char*>(in.var_current) + e4 + 512u * k);
      sat_cp16(d + 512u * k + (unsigned)offsetof(Stage, des), reinterpret_cast<const char*>(in.var_desired) + e4 + 512u * k);
      sat_cp16(d + 512u * k + (unsigned)offsetof(Stage, pen), reinterpret_cast<const char*>(in.var_pending) + e4 + 512u * k);
    }
  }
#pragma unroll
  for (int k = 0; k < (Stage::CAP_V / 2 + 31) / 32; k++)
    if (lane + 32 * k < n2)
      sat_cp16(d + 512u * k + (unsigned)offsetof(Stage, cost), reinterpret_cast<const char*>(in.var_cost) + 2 * (size_t)(unsigned)va * 4u + 16u * lane + 512u * k);
  if (lane < 3 * ng) sat_cp16(desc_s + 16u * lane, reinterpret_cast<const char*>(desc + m0) + 16 * lane);
}

// order-preserving bit pattern of a float64 (-0 == +0; NaN sorts after +inf): costs are compared through it
__device__ __forceinline__ unsigned long long sat_sortable(double x) {
  if (x == 0.0) x = 0.0;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// n_*: lane 0's copy counts (a warp sees < 2^31 models); sum_targets: per-lane partial sums (general path);
// sum_uniform: warp-uniform sums (staged path)
// the same key in two 32-bit words, branch-free (x + 0.0 turns -0 into +0; 4 instructions)
__device__ __forceinline__ void sat_sortable2(double x, unsigned& khi, unsigned& klo) {
  x = d_add(x, 0.0);
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int sgn = hi >> 31;                       // all ones for a negative value
  klo = (unsigned)(lo ^ sgn);
  khi = (unsigned)(hi ^ (sgn | (int)0x80000000));
}

struct SatTally { int n_up, n_down, n_trans; long long sum_targets, sum_uniform; };

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

// One group of up to G models out of a stage (each at most 32 variants, one per lane; everything in shared memory).
// Same arithmetic and order per model as sat_model; the G models only share instructions.  `issue_next()` is called
// exactly once if the function returns true — as soon as the stage is dead — and not at all if it returns false (a
// variant longer than SAT_MAXCNT: the caller takes the general path).
template <bool DETAIL, int G, class IssueNext>
__device__ __forceinline__ bool sat_group_staged(SatWarpSmem<G>* ws, const unsigned tb, const unsigned char* __restrict__ hs_col,
                                                 const int lane, const int m0, const int ng, const SatOut& out, SatTally& tally, IssueNext&& issue_next) {
  const unsigned full = 0xffffffffu;
  constexpr int LPC = 32 / (2 * G);          // lanes per chain of the ordered sums
  constexpr int LPM = 32 / G;                // lanes per model in the model-level part
  const int chain = lane / LPC, jj = lane / LPM;
  // (everything is addressed as ws-> + index: pointers into shared memory held in registers become generic pointers)
  const SatStage<G>* st = &ws->stage;
  const SatDesc* dsc = ws->desc[tb];
  double* terms = ws->terms;
  const int va = dsc[0].v0 & ~3;             // first variant / replica slot of the stage
  const int rbase = dsc[0].r0 & ~1;
  int cnt[G], cur[G], des[G], pen[G], vv[G], ns[G];
  double cost[G], kvThr[G], qThr[G], sumKv[G], sumQ[G], maxKv[G];
  long long maxQ[G];
  bool act[G], hs[G];
  int kvi[G];                                // first replica slot of the lane's variant
  int cmax = 0;
#pragma unroll
  for (int j = 0; j < G; j++) {
    // (descriptor slots past ng hold stale bytes: such a model is given an empty variant range at the stage's start)
    const int v0 = j < ng ? dsc[j].v0 : va, v1 = j < ng ? dsc[j].v1 : va;
    kvThr[j] = dsc[j].kvThr; qThr[j] = dsc[j].qThr;
    vv[j] = v0 + lane;
    act[j] = vv[j] < v1;
    const int i = vv[j] - va;                // < 32 G + 3 for every lane: inside the columns
    const int lo = st->lo[i];
    cnt[j] = act[j] ? st->lo[i + 1] - lo : 0;          // a variant's range ends where the next one starts
    hs[j] = act[j] && (!hs_col || hs_col[vv[j]]);
    cur[j] = hs[j] ? st->cur[i] : 0; des[j] = hs[j] ? st->des[i] : 0; pen[j] = hs[j] ? st->pen[i] : 0;
    cost[j] = st->cost[i];
    kvi[j] = act[j] ? lo - rbase : 0;                  // (inactive lanes read slot 0 of the stage and discard it)
    cmax = cnt[j] > cmax ? cnt[j] : cmax;
    ns[j] = 0; sumKv[j] = 0.0; sumQ[j] = 0.0; maxKv[j] = 0.0; maxQ[j] = 0;
  }
  // ---- analyzeVariant: the lane's variant of every model of the group, replicas in slice order ---------------------
  // Branch-free and warp-uniform: every lane runs to the longest variant of the warp, four slots per round; a lane past
  // the end of one of its variants still loads the slot and discards it.  The slot lies inside the stage: at most
  // SAT_MAXCNT + SAT_UNROLL slots past a range that ends inside kv[] / q[], and q[] is followed by more than that many bytes of
  // per-variant columns.  A group with a longer variant takes the general path.
  static_assert(sizeof(SatStage<G>) - offsetof(SatStage<G>, cost) >= (SAT_MAXCNT + SAT_UNROLL) * 8, "over-read stays inside the stage");
  const int cmaxw = __reduce_max_sync(full, cmax);
  if (cmaxw > SAT_MAXCNT) return false;
  for (int i0 = 0; i0 < cmaxw; i0 += SAT_UNROLL) {
#pragma unroll
    for (int u = 0; u < SAT_UNROLL; u++) {
#pragma unroll
      for (int j = 0; j < G; j++) {
        const int i = i0 + u;
        const bool ok = i < cnt[j];
        const double kv = st->kv[kvi[j] + i];
        const long long q = st->q[kvi[j] + i];
        const double qd = (double)q;
        const bool sat = kv >= kvThr[j] || qd >= qThr[j];                  // analyzer.go:163-164
        if (DETAIL) { if (ok && out.rep_saturated) out.rep_saturated[kvi[j] + rbase + i] = sat ? 1 : 0; }
        if (ok && !sat) {
          sumKv[j] = d_add(sumKv[j], d_sub(kvThr[j], kv));                 // :170-175
          sumQ[j] = d_add(sumQ[j], d_sub(qThr[j], qd));
          ns[j]++;
        }
        if (DETAIL) {
          if (ok && kv > maxKv[j]) maxKv[j] = kv;                          // :179-184
          if (ok && q > maxQ[j]) maxQ[j] = q;
        }
      }
    }
  }
  issue_next();                              // every value of the stage that is still needed lives in registers
  double termKv[G], termQ[G];
#pragma unroll
  for (int j = 0; j < G; j++) {
    double avgKv = 0.0, avgQ = 0.0;
    if (ns[j] > 0) {                                                       // :190-193, one reciprocal for both quotients
      const double nd = (double)ns[j];
      // (a sum over ns > 0 non-saturated replicas is a sum of strictly positive terms — kv < kvThr and queue < qThr,
      // and the difference of two distinct doubles is never zero — so it is positive or NaN, never 0: the window test
      // alone decides; ns <= SAT_MAXCNT < 2^24)
      if (in_window(sumKv[j]) && in_window(sumQ[j])) {
        const double rr = rcp_f32den((float)ns[j], nd);
        avgKv = div_f32den(sumKv[j], nd, rr); avgQ = div_f32den(sumQ[j], nd, rr);
      } else { avgKv = d_div(sumKv[j], nd); avgQ = d_div(sumQ[j], nd); }
    }
    if (DETAIL && act[j]) {
      const int v = vv[j];
      if (out.var_replica_count) out.var_replica_count[v] = cnt[j];
      if (out.var_non_saturated) out.var_non_saturated[v] = ns[j];
      if (out.var_max_kv) out.var_max_kv[v] = maxKv[j];
      if (out.var_max_queue) out.var_max_queue[v] = maxQ[j];
      if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = avgKv;
      if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = avgQ;
    }
    // a variant without metrics (and an inactive lane) contributes an exact +0.0 to the ordered sums (analyzer.go:86-94)
    termKv[j] = d_mul(avgKv, (double)ns[j]); termQ[j] = d_mul(avgQ, (double)ns[j]);
  }
  // ---- ordered accumulation: 2G chains (KV and queue of every model), each in its own group of 32 / 2G lanes --------
  __syncwarp();
#pragma unroll
  for (int j = 0; j < G; j++) { terms[(2 * j) * 32 + lane] = termKv[j]; terms[(2 * j + 1) * 32 + lane] = termQ[j]; }
  __syncwarp();
  double totalSpareKv, totalSpareQueue;
  {
    const double2* col = reinterpret_cast<const double2*>(terms + chain * 32);
    double acc = 0.0;
#pragma unroll
    for (int l = 0; l < 16; l++) { const double2 t = col[l]; acc = d_add(d_add(acc, t.x), t.y); }
    const double other = shfl_xor_d(full, acc, LPC);
    totalSpareKv = (chain & 1) ? other : acc;
    totalSpareQueue = (chain & 1) ? acc : other;
  }
  int nonSaturated = 0, totalReplicas = 0, tr[G];
  unsigned anm[G];
  bool inTr[G];
#pragma unroll
  for (int j = 0; j < G; j++) {
    const int nsj = __reduce_add_sync(full, ns[j]);
    tr[j] = __reduce_add_sync(full, cnt[j]);
    if (jj == j) { nonSaturated = nsj; totalReplicas = tr[j]; }
    anm[j] = __ballot_sync(full, cnt[j] > 0);                              // the variants that enter VariantAnalyses
    inTr[j] = __any_sync(full, cnt[j] > 0 && ((des[j] != 0 && des[j] != cur[j]) || (cnt[j] != cur[j])));   // :322-341
  }
  // ---- model level (analyzer.go:96-121, 199-280): every lane for its own model jj -----------------------------------
  const double m_kvThr = dsc[jj].kvThr, m_qThr = dsc[jj].qThr, kvTrig = dsc[jj].kvTrig, qTrig = dsc[jj].qTrig;
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
      const double avgKvLoad = d_sub(m_kvThr, avgSpareKv), avgQLoad = d_sub(m_qThr, avgSpareQueue);
      const double scale = div_small_int((double)nonSaturated, nonSaturated - 1);
      const double remKv = d_sub(m_kvThr, d_mul(avgKvLoad, scale)), remQ = d_sub(m_qThr, d_mul(avgQLoad, scale));
      downSafe = (remKv >= kvTrig) && (remQ >= qTrig);
    }
  }
  const unsigned mflags = (up ? SAT_FLAG_UP : 0) | (downSafe ? SAT_FLAG_DOWN : 0) | (kvT ? SAT_FLAG_KV : 0) | (qT ? SAT_FLAG_Q : 0);
  unsigned fl[G];
#pragma unroll
  for (int j = 0; j < G; j++) fl[j] = __shfl_sync(full, mflags, j * LPM) | (inTr[j] ? SAT_FLAG_TRANS : 0);
  if (lane == jj * LPM && jj < ng) {          // the first lane of every model's lane group writes the model's results
    const int m = m0 + jj;
    if (DETAIL) {
      if (out.mod_total_replicas) out.mod_total_replicas[m] = totalReplicas;
      if (out.mod_non_saturated) out.mod_non_saturated[m] = nonSaturated;
      if (out.mod_avg_spare_kv) out.mod_avg_spare_kv[m] = avgSpareKv;
      if (out.mod_avg_spare_queue) out.mod_avg_spare_queue[m] = avgSpareQueue;
    }
    unsigned f = fl[0];
#pragma unroll
    for (int j = 1; j < G; j++) f = (jj == j) ? fl[j] : f;
    if (out.mod_flags) out.mod_flags[m] = (unsigned char)f;
  }
#pragma unroll
  for (int j = 0; j < G; j++) {
    if (j >= ng) break;
    const bool upj = fl[j] & SAT_FLAG_UP, downj = fl[j] & SAT_FLAG_DOWN;
    int tgt;
    if (anm[j] && !inTr[j]) {
      // ---- scaling candidate (analyzer.go:376-433): two-word warp arg-min on the order-preserving cost bits --------
      // cheapest without pending, tie -> lower index (:378-395); else most expensive with more than one replica, tie ->
      // higher index (:407-425): the arg-max is the arg-min of the complemented key
      int adj = 0, moved = 0;                 // this lane's / the model's change of replicas
      if (upj || downj) {
        const bool cand = upj ? (cnt[j] > 0 && pen[j] <= 0) : (cnt[j] > 1);
        if (__any_sync(full, cand)) {
          unsigned khi, klo;
          sat_sortable2(cost[j], khi, klo);
          const unsigned flip = upj ? 0u : 0xffffffffu;
          khi = cand ? (khi ^ flip) : 0xffffffffu; klo = cand ? (klo ^ flip) : 0xffffffffu;
          const unsigned mh = __reduce_min_sync(full, khi);
          bool in_ = cand && khi == mh;
          // (all-ones keys of non-candidates can only tie with a candidate whose key is all ones too; `in_` requires cand)
          unsigned wm = __ballot_sync(full, in_);
          if (wm & (wm - 1)) {                // several candidates share the high word: the low word decides
            const unsigned ml = __reduce_min_sync(full, in_ ? klo : 0xffffffffu);
            wm = __ballot_sync(full, in_ && klo == ml);
          }
          const int wl = upj ? (__ffs(wm) - 1) : (31 - __clz(wm));
          moved = upj ? 1 : -1;
          adj = lane == wl ? moved : 0;
        }
      }
      tally.n_up += moved > 0; tally.n_down += moved < 0;
      tgt = cnt[j] > 0 ? cnt[j] + adj : -1;                                  // :362, :399, :428; -1: not in VariantAnalyses
      tally.sum_uniform += tr[j] + moved;
    } else {
      if (!anm[j]) tgt = hs[j] ? cur[j] : -1;                                // nil safety :303-309
      else {
        tally.n_trans++;
        tgt = cnt[j] == 0 ? -1 : ((des[j] != 0 && des[j] != cur[j]) ? des[j] : cur[j]);   // :350-359
      }
      tally.sum_uniform += __reduce_add_sync(full, act[j] && tgt > 0 ? tgt : 0);
    }
    if (act[j] && out.var_target) out.var_target[vv[j]] = tgt;
  }
  return true;
}

template <bool DETAIL, int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) saturation_kernel(SatIn in, SatOut out, const SatDesc* __restrict__ desc, int4* geo) {
  extern __shared__ __align__(16) unsigned char sat_smem[];
  using Warp = SatWarpSmem<G>;
  const unsigned full = 0xffffffffu;
  // lane, warp and the warp's shared-space base address are made opaque: the compiler otherwise re-derives them from the
  // special registers (S2R SR_TID / SR_CgaCtaId, ~40 instructions per model) at every use instead of keeping 3 registers
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("" : "+r"(lane));
  asm volatile("" : "+r"(warp));
  Warp* ws = reinterpret_cast<Warp*>(sat_smem) + warp;
  SatTally tally = {0, 0, 0, 0, 0};
  // model and group indices fit 32 bits (the CSR offsets are int32; the host rejects larger batches); look-ahead indices
  // are compared as unsigned (g + 2 tw < 2^32)
  const unsigned gw = blockIdx.x * WARPS + warp, tw = gridDim.x * WARPS;
  const int M = (int)in.n_models;
  const unsigned NG = (unsigned)sat_groups<G>(M);

  unsigned wbase_s = sat_smem_addr(ws);
  asm volatile("" : "+r"(wbase_s));
  const unsigned st_s = wbase_s + (unsigned)offsetof(Warp, stage), desc_s = wbase_s + (unsigned)offsetof(Warp, desc),
                 bar_s = wbase_s + (unsigned)offsetof(Warp, bar);
  if (lane == 0) {
    sat_mbar_init(&ws->bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // The geometry of a group (16 bytes: variants [x, y), replicas [z, w)) is only needed where its copies are issued; the
  // group analysed reads its models' own descriptors from shared memory.  `gn` = the warp's next group, whose copies are
  // issued in the middle of this trip; it is loaded right after the previous issue, half a trip before it is consumed.
  // `staged`: the stage holds (or is receiving) the group about to be analysed; `par`: phase parity of the next wait on
  // the barrier (a group that is not staged never arms it); `tb`: trip parity = descriptor slot of the group analysed.
  const int4 none = make_int4(-1, 0, 0, 0);
  constexpr unsigned DESC_SLOT = (unsigned)sizeof(SatDesc) * G;
  bool staged = false;
  unsigned par = 0, tb = 0;
  {
    const int4 g0 = gw < NG ? __ldg(geo + gw) : none;
    if (sat_staged<G>(g0)) { const int r = M - (int)gw * G; sat_issue<G>(in, desc, lane, (int)gw * G, r < G ? r : G, g0, st_s, desc_s, bar_s); staged = true; }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int4 gn = gw + tw < NG ? __ldg(geo + gw + tw) : none;
  for (unsigned g = gw; g < NG; g += tw, tb ^= 1u) {
    const unsigned g1 = g + tw;                                              // the warp's next group
    const int m0 = (int)g * G, rem = M - m0, ng = rem < G ? rem : G;
    const bool cur_staged = staged;
    bool issued = false, done = false;
    auto issue_next = [&]() {
      __syncwarp();                                                          // every lane is done with the stage
      staged = false;
      if (sat_staged<G>(gn)) {                                               // (`none` is not staged)
        const int m1 = (int)g1 * G, r1 = M - m1;
        sat_issue<G>(in, desc, lane, m1, r1 < G ? r1 : G, gn, st_s, desc_s + (tb ^ 1u) * DESC_SLOT, bar_s);
        staged = true;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      gn = g1 + tw < NG ? __ldg(geo + g1 + tw) : none;                       // the geometry the next trip issues
      issued = true;
    };
    if (cur_staged) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");                 // this lane's column chunks
      sat_mbar_wait(bar_s, par);                                             // the replica ranges
      par ^= 1u;
      __syncwarp();                                                          // the other lanes' cp.async data
      done = sat_group_staged<DETAIL, G>(ws, tb, in.var_has_state, lane, m0, ng, out, tally, issue_next);
    }
    if (!issued) issue_next();
    // a staged group that turned out to hold a variant of more than SAT_MAXCNT replicas is marked as not staged: the
    // second pass takes it (this warp again: its own store, read back with a coherent load)
    if (cur_staged && !done && lane == 0) reinterpret_cast<int*>(geo + g)[0] = -1;
  }
  // ---- second pass: the groups that were not staged take the general path (no limits: straight from global memory).
  // Rare, and kept out of the pipeline so that its registers and branches do not weigh on the loop above.  The warp looks
  // at 32 of its groups per round (one coherent load per lane) and walks the set bits.
  __syncwarp();
  for (unsigned gb = gw; gb < NG; gb += 32u * tw) {
    const unsigned long long gl = (unsigned long long)gb + (unsigned long long)lane * tw;
    const bool mine = gl < NG && !sat_staged<G>(__ldcg(geo + gl));
    unsigned todo = __ballot_sync(full, mine);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned g = gb + (unsigned)l * tw;
      const int m0 = (int)g * G, rem = M - m0, ng = rem < G ? rem : G;
      for (int j = 0; j < ng; j++) {
        const int4 d = __ldg(reinterpret_cast<const int4*>(desc + m0 + j));
        sat_model<DETAIL, false>(in, in.rep_kv, in.rep_queue, 0, m0 + j, d.x, d.y, out, ws->terms, ws->terms + 32, tally);
      }
    }
  }
  if (out.partials) {
    long long sum_targets = tally.sum_targets;
    for (int o = 16; o; o >>= 1) sum_targets += __shfl_down_sync(full, sum_targets, o);
    sum_targets += tally.sum_uniform;
    if (lane == 0) {
      if (tally.n_up) atomicAdd((unsigned long long*)&out.partials[0], (unsigned long long)tally.n_up);
      if (tally.n_down) atomicAdd((unsigned long long*)&out.partials[1], (unsigned long long)tally.n_down);
      if (tally.n_trans) atomicAdd((unsigned long long*)&out.partials[2], (unsigned long long)tally.n_trans);
      if (sum_targets) atomicAdd((unsigned long long*)&out.partials[3], (unsigned long long)sum_targets);
    }
  }
}

// persistent launch: one block per SM, every warp its own pipeline over groups of G consecutive models
// (G = 2: 24 warps x 8.2 KB per SM).  `desc` = scratch of sat_desc_bytes(M) bytes.
template <int G, int WARPS>
static inline cudaError_t launch_saturation_g(bool detail, int sm_count, long long M, const SatIn& vin, const SatOut& w, SatDesc* d_desc,
                                              cudaStream_t stream) {
  static_assert(sizeof(SatWarpSmem<G>) * WARPS <= 227 * 1024, "shared memory per SM");
  int4* geo = sat_geo_ptr(d_desc, M);
  saturation_desc_kernel<G><<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(vin, d_desc, geo);
  long long blocks = sm_count;
  const long long need = (sat_groups<G>(M) + WARPS - 1) / WARPS;
  if (blocks > need) blocks = need;
  const size_t smem = sizeof(SatWarpSmem<G>) * WARPS;
  auto k = detail ? saturation_kernel<true, G, WARPS> : saturation_kernel<false, G, WARPS>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<(unsigned)blocks, WARPS * 32, smem, stream>>>(vin, w, d_desc, geo);
  return cudaGetLastError();
}
template <int G, int WARPS> static inline cudaError_t sat_prepare_attributes_g() {
  const int smem = (int)(sizeof(SatWarpSmem<G>) * WARPS);
  cudaError_t e = cudaFuncSetAttribute(saturation_kernel<true, G, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(saturation_kernel<false, G, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}
// group size and warps per SM of the product path
constexpr int SAT_G = 2, SAT_WARPS = 24;
#ifdef WVA_SAT_VARIANTS
#define SAT_VARIANTS(X) X(1, 24) X(2, 16) X(2, 20) X(2, 24) X(4, 8) X(4, 12)
#else
#define SAT_VARIANTS(X) X(1, 24) X(2, 20) X(2, 24) X(4, 12)
#endif
// the kernels' dynamic shared-memory limits (callers that capture launch_saturation into a CUDA graph call this first)
static inline cudaError_t sat_prepare_attributes() {
  cudaError_t e = cudaSuccess;
#define X(g, w) if (e == cudaSuccess) e = sat_prepare_attributes_g<g, w>();
  SAT_VARIANTS(X)
#undef X
  return e;
}
// descriptors + analysis (two launches).  WVA_SAT_GROUP = 1 | 2 | 4 (and WVA_SAT_WARPS in a -DWVA_SAT_VARIANTS build)
// override the group size / the warps per SM: measurement and tests only.
static inline cudaError_t launch_saturation(bool detail, int sm_count, long long M, const SatIn& vin, const SatOut& w, SatDesc* d_desc,
                                            cudaStream_t stream) {
  const char* eg = getenv("WVA_SAT_GROUP");
  const char* ew = getenv("WVA_SAT_WARPS");
  const int G = eg ? atoi(eg) : SAT_G, W = ew ? atoi(ew) : 0;
#define X(g, w) if (G == g && W == w) return launch_saturation_g<g, w>(detail, sm_count, M, vin, w_, d_desc, stream);
  const SatOut& w_ = w;
  SAT_VARIANTS(X)
#undef X
  if (G == 1) return launch_saturation_g<1, 24>(detail, sm_count, M, vin, w, d_desc, stream);
  if (G == 4) return launch_saturation_g<4, 12>(detail, sm_count, M, vin, w, d_desc, stream);
  // (the kernel that also writes the analysis fields needs more registers: 20 warps per SM)
  if (detail) return launch_saturation_g<SAT_G, 20>(detail, sm_count, M, vin, w, d_desc, stream);
  return launch_saturation_g<SAT_G, SAT_WARPS>(detail, sm_count, M, vin, w, d_desc, stream);
}

}  // namespace wva
