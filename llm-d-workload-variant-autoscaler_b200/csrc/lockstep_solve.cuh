// lockstep_solve.cuh — one MM1ModelStateDependent.Solve per lane
// (pkg/analyzer/mm1modelstatedependent.go:28-116), all 32 lanes of a warp advancing the
// state index together: pass 1 (sum of p~) for every lane, then pass 2 (normalise +
// accumulate).  The loops are unrolled in chunks of 8 states and carry NO per-state
// control flow:
//   - the early exit (E4) is decided once per chunk: a lane is `done` when its current
//     term is below 2^-54 of every accumulator (compared on the high words — conservative)
//     and the remaining terms are non-increasing; a done lane keeps executing the same
//     instructions, which by that very criterion are exact no-ops on its accumulators;
//   - the exponent window (E3) is tracked with one integer min and one max per state and
//     checked per chunk; a violation (float64 overflow / underflow regime, never seen on
//     sane inputs) marks the solve `bad` and the pair is redone by the literal slow path;
//   - the warp leaves a loop when every lane is done.
// Per state: pass 1 = 5 FP64-pipe ops, pass 2 = 13 FP64-pipe ops.
#pragma once
#include "wva_core.cuh"

namespace wva {

#if defined(__CUDACC__)

#define WVA_HI_LO 0x20B00000   // high word of 2^-500
#define WVA_HI_HI 0x5F300000   // high word of 2^500

// head-table accessors -----------------------------------------------------------------------------
struct WarpTable {   // one table per warp in shared memory: (mu_n, ~1/mu_n) as float64 pairs, broadcast reads
  const double2* t;
  __device__ __forceinline__ void load(int n, double& mu, double& r) const { double2 v = t[n]; mu = v.x; r = v.y; }
  __device__ __forceinline__ double mu_at(int n) const { return t[n].x; }
};
struct LaneTable {   // one float32 column per lane ([n][thread], bank = lane); 1/mu refined on the fly
  const float* t;
  int stride;
  __device__ __forceinline__ void load(int n, double& mu, double& r) const {
    float m32 = t[(size_t)n * stride]; mu = (double)m32; r = rcp_f32den(m32, mu);
  }
  __device__ __forceinline__ double mu_at(int n) const { return (double)t[(size_t)n * stride]; }
};

struct P1 { double p, sum; int mn, mx; };
struct P2 { double p, L, sumP, di, pi; int mn, mx; };

__device__ __forceinline__ void p1_step(P1& s, double lam, double mu, double r) {
  double x = d_mul(s.p, lam);
  int h = d_hi(x);
  s.mn = min(s.mn, h); s.mx = max(s.mx, h);
  s.p = div_f32den(x, mu, r);
  s.sum = d_add(s.sum, s.p);
}
__device__ __forceinline__ void p2_step(P2& s, double lam, double mu, double r, double sum, double rsum) {
  double x = d_mul(s.p, lam);
  int h = d_hi(x);
  s.mn = min(s.mn, h); s.mx = max(s.mx, h);
  s.p = div_f32den(x, mu, r);
  s.pi = div_markstein2(s.p, sum, rsum);
  s.di = d_add(s.di, 1.0);
  s.L = d_add(s.L, d_mul(s.di, s.pi));
  s.sumP = d_add(s.sumP, s.pi);
}

// `active` lanes solve at `lambda`; inactive lanes ride along (lambda 0).  On return `bad` is set
// for a lane whose solve left the exponent window (caller: redo the pair on the slow path).
// __noinline__: the sizer calls this from four places; one copy keeps the unrolled loops in the I-cache
template <class Tab>
__device__ __noinline__ void lockstep_solve(const PairModel& m, const Tab& tab, float lambda, bool active,
                                               SolveStats& st, int& states, bool& bad) {
  const unsigned full = 0xffffffffu;
  const int K = m.K, N = m.N, NH = N - 1;
  const double lam = active ? (double)lambda : 0.0;
  const double lamg = d_mul((double)lambda, 1.000001);
  const bool tail_ok = d_bits(lamg) <= d_bits(m.mu_last);
  const double mu_l = m.mu_last, r_l = m.r_last;
  bad = false;
  states = 0;
  // ------------------------------------------------------------------ pass 1
  P1 a; a.p = 1.0; a.sum = 1.0;
  bool done = !active;
  int n = 0;
  bool all_done = false;
  while (n < NH && !all_done) {                       // head: table entries n .. n+c-1
    const int c = min(8, NH - n);
    const bool eok = (n >= m.mono) && (d_bits(lamg) <= d_bits(tab.mu_at(n)));
    a.mn = 0x7fffffff; a.mx = 0;
    if (c == 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) { double mu, r; tab.load(n + j, mu, r); p1_step(a, lam, mu, r); }
    } else {
      for (int j = 0; j < c; j++) { double mu, r; tab.load(n + j, mu, r); p1_step(a, lam, mu, r); }
    }
    n += c;
    if (!done) {
      states += c;
      if (a.mn < WVA_HI_LO || a.mx >= WVA_HI_HI) { bad = true; done = true; }
      else if (eok && d_hi(a.p) < d_hi(d_mul(a.sum, 0x1p-54))) done = true;
    }
    all_done = !__any_sync(full, !done);
  }
  while (n < K && !all_done) {                        // tail: constant service rate
    const int c = min(8, K - n);
    a.mn = 0x7fffffff; a.mx = 0;
    if (c == 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) p1_step(a, lam, mu_l, r_l);
    } else {
      for (int j = 0; j < c; j++) p1_step(a, lam, mu_l, r_l);
    }
    n += c;
    if (!done) {
      states += c;
      if (a.mn < WVA_HI_LO || a.mx >= WVA_HI_HI) { bad = true; done = true; }
      else if (tail_ok && d_hi(a.p) < d_hi(d_mul(a.sum, 0x1p-54))) done = true;
    }
    all_done = !__any_sync(full, !done);
  }
  // ------------------------------------------------------------------ pass 2
  const double sum = a.sum;
  if (active && !bad && !in_window(sum)) bad = true;
  const double rsum = d_rcp(sum);
  const double cK = 0x1p-55 / (double)K;              // 2x margin covers the rounding of cK itself
  P2 b; b.p = 1.0; b.L = 0.0; b.di = 0.0; b.pi = 0.0;
  b.sumP = d_div(1.0, sum);                           // p[0] = 1/sum
  done = !active || bad;
  all_done = !__any_sync(full, !done);
  n = 0;
  while (n < NH && !all_done) {                       // head (i = n+1 <= N-1)
    const int c = min(8, NH - n);
    const bool eok = (n >= m.mono) && (d_bits(lamg) <= d_bits(tab.mu_at(n)));
    b.mn = 0x7fffffff; b.mx = 0;
    if (c == 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) { double mu, r; tab.load(n + j, mu, r); p2_step(b, lam, mu, r, sum, rsum); }
    } else {
      for (int j = 0; j < c; j++) { double mu, r; tab.load(n + j, mu, r); p2_step(b, lam, mu, r, sum, rsum); }
    }
    n += c;
    if (!done) {
      states += c;
      if (b.mn < WVA_HI_LO || b.mx >= WVA_HI_HI) { bad = true; done = true; }
      else if (eok) {
        int thr = min(d_hi(d_mul(b.L, cK)), d_hi(d_mul(b.sumP, 0x1p-54)));
        if (d_hi(b.pi) < thr) done = true;
      }
    }
    all_done = !__any_sync(full, !done);
  }
  double Lserv;
  if (all_done) {
    // every lane left before state N: the accumulators no longer change, so the value the
    // reference computes at i == N (mm1modelstatedependent.go:52-54) is the current one
    Lserv = d_add(b.L, d_mul(d_sub(1.0, b.sumP), (double)N));
  } else {
    // state i == N uses servRate[N-1]
    b.mn = 0x7fffffff; b.mx = 0;
    p2_step(b, lam, mu_l, r_l, sum, rsum);
    n = N;
    if (!done) { states += 1; if (b.mn < WVA_HI_LO || b.mx >= WVA_HI_HI) { bad = true; done = true; } }
    Lserv = d_add(b.L, d_mul(d_sub(1.0, b.sumP), (double)N));
    if (!done && tail_ok) {
      int thr = min(d_hi(d_mul(b.L, cK)), d_hi(d_mul(b.sumP, 0x1p-54)));
      if (d_hi(b.pi) < thr) done = true;
    }
    all_done = !__any_sync(full, !done);
  }
  bool reached_K = false;
  while (n < K && !all_done) {                        // tail (i = n+1 in N+1 .. K)
    const int c = min(8, K - n);
    b.mn = 0x7fffffff; b.mx = 0;
    if (c == 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) p2_step(b, lam, mu_l, r_l, sum, rsum);
    } else {
      for (int j = 0; j < c; j++) p2_step(b, lam, mu_l, r_l, sum, rsum);
    }
    n += c;
    if (!done) {
      states += c;
      if (b.mn < WVA_HI_LO || b.mx >= WVA_HI_HI) { bad = true; done = true; }
      else if (n == K) { reached_K = true; done = true; }
      else if (tail_ok) {
        int thr = min(d_hi(d_mul(b.L, cK)), d_hi(d_mul(b.sumP, 0x1p-54)));
        if (d_hi(b.pi) < thr) done = true;
      }
    }
    all_done = !__any_sync(full, !done);
  }
  const double pK = reached_K ? b.pi : 0.0;            // (E4): an early exit implies p[K] < 2^-53
  st.avgNumInServers = (float)Lserv;
  st.avgNumInSystem = (float)b.L;
  st.throughput = f_mul(lambda, f_sub(1.0f, (float)pK));
  st.avgRespTime = f_div(st.avgNumInSystem, st.throughput);
  st.avgServTime = f_div(st.avgNumInServers, st.throughput);
  float w = f_sub(st.avgRespTime, st.avgServTime);
  st.avgWaitTime = (w < 0.0f) ? 0.0f : w;
}

// evaluation values of a finished solve (EvalTTFT / EvalITL, queueanalyzer.go:283-308)
__device__ __forceinline__ void eval_values(const PairModel& m, const SolveStats& st, float* ttft, float* itl, float* pf) {
  *pf = prefill_time(m, st.avgNumInServers);
  *itl = f_div(f_sub(st.avgServTime, *pf), m.out_tok);
  *ttft = f_add(f_add(st.avgWaitTime, *pf), *itl);
}

#endif  // __CUDACC__
}  // namespace wva
