// grid_kernel.cuh — replica-grid evaluator: for every (server, accelerator, r) one
// QueueAnalyzer.Analyze(totalRate / r) (pkg/analyzer/queueanalyzer.go:127-167), the
// call CreateAllocation makes at pkg/core/allocation.go:140-148 with numReplicas = r.
//
// Mapping: one warp per (server, accelerator) pair ("one block per model sweeps its
// variant x replica grid" at warp granularity).  The pair's head table (mu_n, 1/mu_n)
// is built once by the warp into shared memory; the 32 lanes then take 32 consecutive
// replica levels per round and solve their chains in lock step (lockstep_solve.cuh):
// broadcast table reads, no per-state control flow, a lane whose terms became no-ops
// idles until the round's longest chain ends.  The per-pair frontier (smallest r meeting
// every SLO) is a warp-shuffle min.
#pragma once
#include "wva_core.cuh"
#include "lockstep_solve.cuh"

namespace wva {

struct GridOut {
  unsigned char* ok;
  float *ttft, *itl, *rho, *tput;
  int* frontier;
};

struct GridCounters {
  unsigned long long next_pair, solves, states, overflow;
  int limit_hit, pad_;
  unsigned long long slots0, slots_rest, rounds;   // 32 x longest chain of the round: first round of a pair / later rounds
};

// per-pair preparation shared by all lanes of the warp; false -> every level is "not ok"
__device__ __forceinline__ bool grid_setup(PairModel& m, const SysView& s, int srv, int acc, int n_limit,
                                           float* total_rate, float* slo_ttft, float* slo_itl, float* slo_tps,
                                           int* limit_hit) {
  float arrival = s.srv_arrival[srv];
  int in_tok = s.srv_in_tokens[srv], out_tok = s.srv_out_tokens[srv];
  int model = s.srv_model[srv];
  if (arrival < 0.0f || in_tok < 0 || out_tok < 0 || model < 0 || model >= s.n_models) return false;
  size_t pi = (size_t)model * s.n_acc + acc;
  if (!s.perf_present[pi] || !s.srv_target_present[srv] || arrival == 0.0f || out_tok == 0) return false;
  long long N;
  if (s.srv_max_batch[srv] > 0) N = s.srv_max_batch[srv];
  else { N = (long long)s.perf_max_batch[pi] * s.perf_at_tokens[pi] / out_tok; if (N < 1) N = 1; }
  if (N > n_limit) { *limit_hit = 1; return false; }
  model_init(m, s.perf_alpha[pi], s.perf_beta[pi], s.perf_gamma[pi], in_tok, out_tok, (int)N);
  *slo_ttft = s.srv_slo_ttft[srv]; *slo_itl = s.srv_slo_itl[srv]; *slo_tps = s.srv_slo_tps[srv];
  *total_rate = (*slo_tps == 0.0f) ? f_div(arrival, 60.0f) : f_div(*slo_tps, (float)out_tok);
  return true;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
grid_kernel(SysView s, int R, GridOut out, unsigned long long n_pairs, int nmax, GridCounters* ctr) {
  extern __shared__ double2 smem_grid[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* tab = smem_grid + (size_t)warp * nmax;
  float* tabf = (float*)(smem_grid + (size_t)WARPS * nmax) + (size_t)warp * nmax;
  unsigned long long my_solves = 0, my_states = 0, my_s0 = 0, my_sr = 0, my_rounds = 0;

  while (true) {
    unsigned long long pair = 0;
    if (lane == 0) pair = atomicAdd(&ctr->next_pair, 1ull);
    pair = __shfl_sync(full, pair, 0);
    if (pair >= n_pairs) break;
    const int srv = (int)(pair / (unsigned)s.n_acc), acc = (int)(pair % (unsigned)s.n_acc);
    const size_t obase = (size_t)pair * (size_t)R;
    PairModel m;
    float total_rate = 0, slo_ttft = 0, slo_itl = 0, slo_tps = 0;
    int lim = 0;
    bool valid = grid_setup(m, s, srv, acc, nmax, &total_rate, &slo_ttft, &slo_itl, &slo_tps, &lim);
    if (lim && lane == 0) ctr->limit_hit = 1;
    if (!valid) {
      for (int r = lane; r < R; r += 32) {
        if (out.ok) out.ok[obase + r] = 0;
        if (out.ttft) out.ttft[obase + r] = 0.0f;
        if (out.itl) out.itl[obase + r] = 0.0f;
        if (out.rho) out.rho[obase + r] = 0.0f;
        if (out.tput) out.tput[obase + r] = 0.0f;
      }
      if (out.frontier && lane == 0) out.frontier[pair] = 0;
      continue;
    }
    __syncwarp();
    for (int n = lane; n < m.N; n += 32) {
      float mu32 = serv_rate(m, n + 1);
      double mu = (double)mu32;
      tabf[n] = mu32;
      tab[n] = make_double2(mu, rcp_f32den(mu32, mu));
    }
    __syncwarp();
    model_finish(m, tabf, 1);
    const float lambda_tps = f_mul(m.lambda_max, f_sub(1.0f, WVA_STABILITY_SAFETY));
    int front = 0x7fffffff;
    bool first_round = true;

    for (int r0 = 0; r0 < R; r0 += 32) {
      const int r = r0 + lane + 1;
      const bool in_range = r <= R;
      const float rate = f_div(total_rate, (float)r);
      const bool admitted = in_range && analyze_admits(m, rate);      // queueanalyzer.go:128-136
      if (!__any_sync(full, admitted)) {
        if (in_range) {
          const size_t o = obase + (size_t)(r - 1);
          if (out.ok) out.ok[o] = 0;
          if (out.ttft) out.ttft[o] = 0.0f;
          if (out.itl) out.itl[o] = 0.0f;
          if (out.rho) out.rho[o] = 0.0f;
          if (out.tput) out.tput[o] = 0.0f;
        }
        continue;
      }
      const float lambda = f_div(rate, 1000.0f);
      SolveStats st;
      int sv = 0;
      bool bad = false, ovf = false;
      lockstep_solve(m, WarpTable{tab}, lambda, admitted, st, sv, bad);
      if (admitted) { my_solves++; my_states += (unsigned long long)sv; }
      {
        int mxs = admitted ? sv : 0;
        for (int o = 16; o; o >>= 1) mxs = max(mxs, __shfl_xor_sync(full, mxs, o));
        if (lane == 0) { if (first_round) my_s0 += 32ull * mxs; else my_sr += 32ull * mxs; my_rounds++; }
        first_round = false;
      }
      if (admitted && bad) {
        // outside the exponent window: redo this level alone through the per-lane state machine
        // (IEEE divisions); a true float64 overflow stays flagged (only the sizer has the rescale path)
        Chain c;
        chain_start(c, lambda);
        c.tail_ok = d_bits(c.lamg) <= d_bits(m.mu_last);
        while (!chain_step(c, m, st)) {}
        ovf = c.phase == CH_OVERFLOW;
        if (ovf) atomicAdd(&ctr->overflow, 1ull);
      }
      if (in_range) {
        const size_t o = obase + (size_t)(r - 1);
        if (!admitted || ovf) {
          if (out.ok) out.ok[o] = 0;
          if (out.ttft) out.ttft[o] = 0.0f;
          if (out.itl) out.itl[o] = 0.0f;
          if (out.rho) out.rho[o] = 0.0f;
          if (out.tput) out.tput[o] = 0.0f;
        } else {
          // Analyze: queueanalyzer.go:143-166
          float pf, dec, avg_ttft;
          eval_values(m, st, &avg_ttft, &dec, &pf);
          float rho = f_div(st.avgNumInServers, (float)m.N);
          rho = fminf(fmaxf(rho, 0.0f), 1.0f);
          if (out.ok) out.ok[o] = 1;
          if (out.ttft) out.ttft[o] = f_add(st.avgWaitTime, pf);   // allocation.go:148
          if (out.itl) out.itl[o] = dec;
          if (out.rho) out.rho[o] = rho;
          if (out.tput) out.tput[o] = f_mul(st.throughput, 1000.0f);
          bool meets = (slo_ttft <= 0.0f || avg_ttft <= slo_ttft) && (slo_itl <= 0.0f || dec <= slo_itl) &&
                       (slo_tps <= 0.0f || lambda <= lambda_tps);
          if (meets && r < front) front = r;
        }
      }
    }
    for (int o = 16; o; o >>= 1) front = min(front, __shfl_down_sync(full, front, o));
    if (out.frontier && lane == 0) out.frontier[pair] = (front == 0x7fffffff) ? 0 : front;
    __syncwarp();
  }
  for (int o = 16; o; o >>= 1) {
    my_solves += __shfl_down_sync(full, my_solves, o);
    my_states += __shfl_down_sync(full, my_states, o);
  }
  if (lane == 0) {
    atomicAdd(&ctr->solves, my_solves); atomicAdd(&ctr->states, my_states);
    atomicAdd(&ctr->slots0, my_s0); atomicAdd(&ctr->slots_rest, my_sr); atomicAdd(&ctr->rounds, my_rounds);
  }
}

}  // namespace wva
