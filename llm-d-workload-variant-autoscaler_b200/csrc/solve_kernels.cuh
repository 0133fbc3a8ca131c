// solve_kernels.cuh — the allocator side: Solver.SolveUnlimited (per-server argmin of
// value, pkg/solver/solver.go:63-79) as a warp-shuffle reduction, and
// System.AllocateByType (pkg/core/system.go:271-299) as a deterministic two-level
// reduction (int64 counts exact; float64 cost sums in a fixed order).
#pragma once
#include "wva_core.cuh"

namespace wva {

struct SolView {  // device image of wva_solution
  unsigned char* state;
  int *acc, *num_replicas, *batch_size;
  float *cost, *value, *itl, *ttft, *rho, *max_arrv_rate;
};

// One warp per server; lanes stride the accelerator axis.  Strict '<' against
// MaxFloat32 and ascending-index visiting order = the canonical tie-break
// (lowest accelerator index among equal minima), as in the oracle.
__global__ void __launch_bounds__(256) solve_unlimited_kernel(SysView s, CandView c, SolView o) {
  const unsigned full = 0xffffffffu;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= s.n_servers) return;
  size_t base = (size_t)warp * s.n_acc;
  float best = FLT_MAX;
  int best_a = -1;
  for (int a = lane; a < s.n_acc; a += 32) {
    if (c.state[base + a] == ALLOC_NONE) continue;
    float v = c.value[base + a];
    if (v < best) { best = v; best_a = a; }   // per-lane: ascending a, strict <
  }
  for (int off = 16; off; off >>= 1) {
    float ov = __shfl_down_sync(full, best, off);
    int oa = __shfl_down_sync(full, best_a, off);
    // take the other lane's candidate if strictly smaller, or equal with a lower index
    bool take = (oa >= 0) && (best_a < 0 || ov < best || (ov == best && oa < best_a));
    if (take) { best = ov; best_a = oa; }
  }
  if (lane == 0) {
    if (best_a < 0) {
      o.state[warp] = ALLOC_NONE; o.acc[warp] = -1; o.num_replicas[warp] = 0; o.batch_size[warp] = 0;
      o.cost[warp] = o.value[warp] = o.itl[warp] = o.ttft[warp] = o.rho[warp] = o.max_arrv_rate[warp] = 0.0f;
    } else {
      size_t i = base + best_a;
      unsigned char st = c.state[i];
      o.state[warp] = st;
      o.acc[warp] = (st == ALLOC_ACC) ? best_a : -1;
      o.num_replicas[warp] = c.num_replicas[i];
      o.batch_size[warp] = c.batch_size[i];
      o.cost[warp] = c.cost[i]; o.value[warp] = c.value[i]; o.itl[warp] = c.itl[i];
      o.ttft[warp] = c.ttft[i]; o.rho[warp] = c.rho[i]; o.max_arrv_rate[warp] = c.max_arrv_rate[i];
    }
  }
}

// AllocateByType, level 1: each block reduces a contiguous slice of servers into
// partial[block][type] (count, cost) with a fixed-order shared-memory tree.
#define WVA_MAX_TYPES 64
__global__ void __launch_bounds__(256) by_type_partial_kernel(SysView s, SolView o, long long* part_count,
                                                              double* part_cost) {
  __shared__ long long sc[256];
  __shared__ double sd[256];
  int srv = blockIdx.x * blockDim.x + threadIdx.x;
  int t = -1; long long cnt = 0; double cst = 0.0;
  if (srv < s.n_servers && o.state[srv] == ALLOC_ACC && s.srv_model[srv] >= 0) {
    int a = o.acc[srv];
    t = s.acc_type[a];
    cnt = (long long)o.num_replicas[srv] * num_instances(s, s.srv_model[srv], a) * s.acc_multiplicity[a];
    cst = (double)o.cost[srv];
  }
  for (int ty = 0; ty < s.n_types; ty++) {
    sc[threadIdx.x] = (t == ty) ? cnt : 0;
    sd[threadIdx.x] = (t == ty) ? cst : 0.0;
    __syncthreads();
    for (int w = 128; w; w >>= 1) {
      if (threadIdx.x < w) { sc[threadIdx.x] += sc[threadIdx.x + w]; sd[threadIdx.x] += sd[threadIdx.x + w]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) { part_count[(size_t)blockIdx.x * s.n_types + ty] = sc[0]; part_cost[(size_t)blockIdx.x * s.n_types + ty] = sd[0]; }
    __syncthreads();
  }
}
// level 2: one block per type sums the partials in a fixed order
__global__ void __launch_bounds__(256) by_type_final_kernel(int n_types, int n_parts, const long long* part_count,
                                                            const double* part_cost, long long* type_count,
                                                            double* type_cost) {
  __shared__ long long sc[256];
  __shared__ double sd[256];
  int ty = blockIdx.x;
  long long c = 0; double d = 0.0;
  for (int p = threadIdx.x; p < n_parts; p += 256) { c += part_count[(size_t)p * n_types + ty]; d += part_cost[(size_t)p * n_types + ty]; }
  sc[threadIdx.x] = c; sd[threadIdx.x] = d;
  __syncthreads();
  for (int w = 128; w; w >>= 1) {
    if (threadIdx.x < w) { sc[threadIdx.x] += sc[threadIdx.x + w]; sd[threadIdx.x] += sd[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { type_count[ty] = sc[0]; type_cost[ty] = sd[0]; }
}

}  // namespace wva
