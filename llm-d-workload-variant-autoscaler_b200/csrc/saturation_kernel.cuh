// saturation_kernel.cuh — V1 saturation capacity model for a batch of models:
// saturation.Analyzer.AnalyzeModelSaturation / analyzeVariant / shouldScaleUp /
// isScaleDownSafe (internal/saturation/analyzer.go:31-280) and
// CalculateSaturationTargets (analyzer.go:290-439).
//
// HBM-bound by design: 16 B per replica (kv float64 + queue int64) + 32 B per variant + 40 B per model, each read once.
// The stream is moved by the TMA unit, not by the warps: the models are cut into chunks of SAT_G = 8 consecutive
// models (one per warp of the CTA); the CSR layout makes everything a chunk needs CONTIGUOUS in each of the 13 input
// arrays, so one elected thread issues 13 `cp.async.bulk` (1-D TMA) copies global -> shared memory per chunk, completing
// on an mbarrier, into a ring of SAT_NS stages: two chunks (~2 x 25 KB) are always in flight per CTA, two CTAs per SM.
// The warps never touch global memory for input; they wait on the stage's mbarrier, then analyse their model out of
// shared memory: one lane per variant streams its replicas in slice order (the per-variant float64 sums are order
// dependent), the variant -> model accumulation runs in ascending variant index (broadcast reads, fixed 32-step
// unrolled chain; lanes without metrics contribute an exact +0.0), the cheapest / most-expensive variant search runs
// only for the models that scale.  A chunk whose ranges exceed a stage (a model with thousands of replicas) takes the
// same code over global memory instead.
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

// ---- chunk geometry ------------------------------------------------------------------------------------------------
constexpr int SAT_G = 8;             // models per chunk = warps per CTA
constexpr int SAT_NS = 3;            // stages per CTA
constexpr int SAT_CAP_REP = 1536;    // replicas a stage holds (8 models x 32 variants x 4.5 replicas = 1152 on average)
constexpr int SAT_CAP_VAR = 288;     // variants a stage holds

struct alignas(128) SatStage {       // every member starts on a 16-byte boundary (cp.async.bulk destination)
  double kv[SAT_CAP_REP + 2];
  long long q[SAT_CAP_REP + 2];
  double cost[SAT_CAP_VAR + 2];
  double cfg[4][SAT_G];
  int vro[SAT_CAP_VAR + 8];
  int cur[SAT_CAP_VAR + 4], des[SAT_CAP_VAR + 4], pen[SAT_CAP_VAR + 4];
  int mvo[SAT_G + 4];
  unsigned char hs[SAT_CAP_VAR + 32];
};
static_assert(sizeof(SatStage) % 128 == 0 && offsetof(SatStage, q) % 16 == 0 && offsetof(SatStage, cost) % 16 == 0 &&
              offsetof(SatStage, cfg) % 16 == 0 && offsetof(SatStage, vro) % 16 == 0 && offsetof(SatStage, cur) % 16 == 0 &&
              offsetof(SatStage, des) % 16 == 0 && offsetof(SatStage, pen) % 16 == 0 && offsetof(SatStage, mvo) % 16 == 0 &&
              offsetof(SatStage, hs) % 16 == 0, "SatStage members must be 16-byte aligned");
constexpr size_t SAT_SMEM_BYTES = sizeof(SatStage) * SAT_NS + 128;

struct SatChunk { int v_lo, v_hi, r_lo, r_hi; };   // variants [v_lo, v_hi) and replicas [r_lo, r_hi) of models [c*G, (c+1)*G)

// chunk descriptors: the two dependent CSR look-ups of every chunk, done once so that the copy-issuing thread of the
// main kernel never waits on global memory (16 B per 8 models: 0.06 % of the stream)
__global__ void __launch_bounds__(256) saturation_chunk_kernel(SatIn in, SatChunk* desc, long long n_chunks) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const long long m0 = c * SAT_G, m1 = min(in.n_models, m0 + SAT_G);
  SatChunk d;
  d.v_lo = in.model_variant_off[m0]; d.v_hi = in.model_variant_off[m1];
  d.r_lo = in.variant_replica_off[d.v_lo]; d.r_hi = in.variant_replica_off[d.v_hi];
  desc[c] = d;
}
__device__ __forceinline__ bool sat_chunk_fits(const SatChunk& d) {
  return d.r_hi - d.r_lo <= SAT_CAP_REP && d.v_hi - d.v_lo <= SAT_CAP_VAR;
}

// ---- PTX: mbarrier + 1-D bulk copy ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sat_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sat_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sat_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sat_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sat_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sat_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(sat_smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a positive multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void sat_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sat_smem_addr(dst)), "l"(src), "r"(bytes), "r"(sat_smem_addr(bar)) : "memory");
}

// where a model's inputs are read from: a stage in shared memory (indices rebased by the stage's aligned starts) or the
// global arrays themselves (all offsets 0)
struct SatSrc {
  const double* kv; const long long* q; const double* cost; const int *vro, *cur, *des, *pen; const unsigned char* hs;
  int r0, v0i, v0c, v0h;            // element index of kv[0] / vro[0], cur[0].. / cost[0] / hs[0]
};

// geometry of a staged chunk (shared by the issuing thread and the consumers)
struct SatGeom { int r_a, v_i, v_c, v_h; unsigned n_rep, n_vro, n_var4, n_cost, n_hs; };
__device__ __forceinline__ SatGeom sat_geom(const SatChunk& d) {
  SatGeom g;
  g.r_a = d.r_lo & ~1; g.v_i = d.v_lo & ~3; g.v_c = d.v_lo & ~1; g.v_h = d.v_lo & ~15;
  g.n_rep = (unsigned)((d.r_hi - g.r_a + 1) & ~1);
  g.n_vro = (unsigned)((d.v_hi + 1 - g.v_i + 3) & ~3);
  g.n_var4 = (unsigned)((d.v_hi - g.v_i + 3) & ~3);
  g.n_cost = (unsigned)((d.v_hi - g.v_c + 1) & ~1);
  g.n_hs = (unsigned)((d.v_hi - g.v_h + 15) & ~15);
  return g;
}

// one elected thread: all copies of a chunk onto the stage's mbarrier (over-reads of < 16 B past an array's end stay
// inside the input arena, whose sub-arrays are 256-byte padded — capi_aux.inl wva_saturation_upload)
__device__ __forceinline__ void sat_issue_chunk(const SatIn& in, const SatChunk& d, long long c, SatStage* st,
                                                unsigned long long* bar) {
  const SatGeom g = sat_geom(d);
  const long long m0 = c * SAT_G;
  const unsigned b_rep = g.n_rep * 8, b_vro = g.n_vro * 4, b_v4 = g.n_var4 * 4, b_cost = g.n_cost * 8, b_hs = in.var_has_state ? g.n_hs : 0;
  const unsigned total = 2 * b_rep + b_vro + 3 * b_v4 + b_cost + b_hs + (SAT_G + 4) * 4 + 4 * SAT_G * 8;
  sat_mbar_expect_tx(bar, total);
  if (b_rep) { sat_bulk_g2s(st->kv, in.rep_kv + g.r_a, b_rep, bar); sat_bulk_g2s(st->q, in.rep_queue + g.r_a, b_rep, bar); }
  sat_bulk_g2s(st->vro, in.variant_replica_off + g.v_i, b_vro, bar);
  if (b_v4) {
    sat_bulk_g2s(st->cur, in.var_current + g.v_i, b_v4, bar); sat_bulk_g2s(st->des, in.var_desired + g.v_i, b_v4, bar);
    sat_bulk_g2s(st->pen, in.var_pending + g.v_i, b_v4, bar);
  }
  if (b_cost) sat_bulk_g2s(st->cost, in.var_cost + g.v_c, b_cost, bar);
  if (b_hs) sat_bulk_g2s(st->hs, in.var_has_state + g.v_h, b_hs, bar);
  sat_bulk_g2s(st->mvo, in.model_variant_off + m0, (SAT_G + 4) * 4, bar);
  sat_bulk_g2s(st->cfg[0], in.cfg_kv_threshold + m0, SAT_G * 8, bar);
  sat_bulk_g2s(st->cfg[1], in.cfg_queue_threshold + m0, SAT_G * 8, bar);
  sat_bulk_g2s(st->cfg[2], in.cfg_kv_trigger + m0, SAT_G * 8, bar);
  sat_bulk_g2s(st->cfg[3], in.cfg_queue_trigger + m0, SAT_G * 8, bar);
}

struct SatTally { long long n_up, n_down, n_trans, sum_targets; };

// One model, one warp.  STAGED: inputs in shared memory (simple per-lane replica loop); otherwise global memory with
// four independent loads in flight per lane.
template <bool DETAIL, bool STAGED>
__device__ __forceinline__ void sat_model(const SatSrc& src, const bool has_hs, long long m, int v0, int v1, double kvThr,
                                          double qThr, double kvTrig, double qTrig, const SatOut& out, double2* my_terms,
                                          SatTally& tally) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
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
    const bool hs = act && (!has_hs || src.hs[v - src.v0h]);
    const int cur = hs ? src.cur[v - src.v0i] : 0, des = hs ? src.des[v - src.v0i] : 0;
    r_cur = cur; r_des = des; r_pen = hs ? src.pen[v - src.v0i] : 0; r_cost = act ? src.cost[v - src.v0c] : 0.0;
    if (act) {
      const int lo = src.vro[v - src.v0i], hi = src.vro[v + 1 - src.v0i];
      cnt = hi - lo;
      if (STAGED) {
        for (int r = lo; r < hi; r++) {
          const double kv = src.kv[r - src.r0];
          const long long q = src.q[r - src.r0];
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
            kvv[j] = inb ? __ldg(src.kv + base + j) : 0.0;      // unstaged: src is the global arrays, r0 == 0
            qq[j] = inb ? __ldg(src.q + base + j) : 0;
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
          cnt = src.vro[v + 1 - src.v0i] - src.vro[v - src.v0i];
          const bool hs2 = !has_hs || src.hs[v - src.v0h];
          pen = hs2 ? src.pen[v - src.v0i] : 0;
          vcost = src.cost[v - src.v0c];
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
    if (nAnalysed > 0 && inTransition) tally.n_trans++;
    if (plus_v >= 0) tally.n_up++;
    if (minus_v >= 0) tally.n_down++;
  }
  // ---- targets (analyzer.go:303-436) ------------------------------------------------------------
  for (int c0 = v0; c0 < v1; c0 += 32) {
    const int v = c0 + lane;
    if (v >= v1) continue;
    const int cnt = single ? r_cnt : src.vro[v + 1 - src.v0i] - src.vro[v - src.v0i];
    const bool hs = !has_hs || src.hs[v - src.v0h];
    int tgt;
    if (nAnalysed == 0) tgt = hs ? (single ? r_cur : src.cur[v - src.v0i]) : -1;   // nil safety :303-309
    else if (cnt == 0) tgt = -1;                                      // not in VariantAnalyses
    else if (inTransition) {                                          // :350-359
      const int cur = single ? r_cur : (hs ? src.cur[v - src.v0i] : 0), des = single ? r_des : (hs ? src.des[v - src.v0i] : 0);
      tgt = (des != 0 && des != cur) ? des : cur;
    } else tgt = cnt + (v == plus_v ? 1 : 0) - (v == minus_v ? 1 : 0);   // :362, :399, :428
    if (out.var_target) out.var_target[v] = tgt;
    if (tgt >= 0) tally.sum_targets += tgt;
  }
}

template <bool DETAIL>
__global__ void __launch_bounds__(SAT_G * 32, 2) saturation_kernel(SatIn in, SatOut out, const SatChunk* __restrict__ desc,
                                                                   long long n_chunks) {
  extern __shared__ __align__(128) unsigned char sat_smem[];
  SatStage* stages = reinterpret_cast<SatStage*>(sat_smem);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sat_smem + sizeof(SatStage) * SAT_NS);
  __shared__ double2 terms[SAT_G][32];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* my_terms = terms[warp];
  SatTally tally = {0, 0, 0, 0};
  const bool has_hs = in.var_has_state != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SAT_NS; s++) sat_mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // prologue: the first SAT_NS - 1 chunks of this CTA
  if (threadIdx.x == 0) {
    for (int k = 0; k < SAT_NS - 1; k++) {
      const long long c = (long long)blockIdx.x + (long long)k * gridDim.x;
      if (c < n_chunks) { const SatChunk d = desc[c]; if (sat_chunk_fits(d)) sat_issue_chunk(in, d, c, &stages[k], &bars[k]); }
    }
  }
  unsigned par_bits = 0;          // bit s = phase parity of the next wait on stage s (uniform over the CTA)
  long long it = 0;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
    // keep SAT_NS - 1 chunks in flight: the stage refilled here was consumed in the previous iteration (barrier below)
    if (threadIdx.x == 0) {
      const long long cn = c + (long long)(SAT_NS - 1) * gridDim.x;
      if (cn < n_chunks) {
        const int sn = (int)((it + SAT_NS - 1) % SAT_NS);
        const SatChunk dn = desc[cn];
        if (sat_chunk_fits(dn)) sat_issue_chunk(in, dn, cn, &stages[sn], &bars[sn]);
      }
    }
    const SatChunk d = desc[c];
    const long long m = c * SAT_G + warp;
    const int s = (int)(it % SAT_NS);
    if (sat_chunk_fits(d)) {
      // a chunk that is not staged never arms its stage's barrier: the parity is counted per staged use
      sat_mbar_wait(&bars[s], (par_bits >> s) & 1u);
      par_bits ^= 1u << s;
      if (m < in.n_models) {
        const SatStage* st = &stages[s];
        const SatGeom g = sat_geom(d);
        SatSrc src;
        src.kv = st->kv; src.q = st->q; src.cost = st->cost; src.vro = st->vro; src.cur = st->cur; src.des = st->des;
        src.pen = st->pen; src.hs = st->hs; src.r0 = g.r_a; src.v0i = g.v_i; src.v0c = g.v_c; src.v0h = g.v_h;
        sat_model<DETAIL, true>(src, has_hs, m, st->mvo[warp], st->mvo[warp + 1], st->cfg[0][warp], st->cfg[1][warp],
                                st->cfg[2][warp], st->cfg[3][warp], out, my_terms, tally);
      }
    } else if (m < in.n_models) {
      SatSrc src;
      src.kv = in.rep_kv; src.q = in.rep_queue; src.cost = in.var_cost; src.vro = in.variant_replica_off;
      src.cur = in.var_current; src.des = in.var_desired; src.pen = in.var_pending; src.hs = in.var_has_state;
      src.r0 = 0; src.v0i = 0; src.v0c = 0; src.v0h = 0;
      sat_model<DETAIL, false>(src, has_hs, m, in.model_variant_off[m], in.model_variant_off[m + 1], in.cfg_kv_threshold[m],
                               in.cfg_queue_threshold[m], in.cfg_kv_trigger[m], in.cfg_queue_trigger[m], out, my_terms, tally);
    }
    __syncthreads();              // every warp is done with stage s before it is refilled (next iteration, thread 0)
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

}  // namespace wva
