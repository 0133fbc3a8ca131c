// limiter_kernel.cuh — GPU-count limiter: DefaultLimiter.Limit
// (internal/engines/pipeline/default_limiter.go:42-81), TypeInventory.CreateAllocator /
// typeAllocator.TryAllocate (type_inventory.go:222-243,347-373) and
// GreedyBySaturation.Allocate (greedy_saturation_algorithm.go:34-108).
//
// The reference's sequential "take min(requested, remaining)" per accelerator type
// is exactly a per-type exclusive prefix sum over the candidates in priority order:
//   remaining_i = max(0, avail_t - sum_{j<i, type t} requested_j).
// Device pipeline: used[t] by integer atomics -> order-preserving compaction of the
// scale-up candidates -> three stable LSD radix passes (cost, spare, type) carrying
// the decision index (ties end in ascending decision index = the oracle's canonical
// order) -> segmented exclusive scan by type -> scatter of the mutated fields.
// The radix sort / scan / select primitives are CUB (CUDA toolkit); the key
// construction, the allocation rule and the scatter are the kernels below.
#pragma once
#include "wva_core.cuh"
#include <cub/cub.cuh>

namespace wva {

// order-preserving map double -> uint64 (-0 == +0, NaN last)
__device__ __forceinline__ unsigned long long sortable_f64(double x) {
  if (x == 0.0) x = 0.0;
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) limiter_prepare_kernel(long long D, int T, const int* acc_type, const int* current,
                                                              const int* target, const int* gpr, long long* used,
                                                              unsigned char* is_cand, int* out_target, int* out_gpus,
                                                              unsigned char* out_limited) {
  long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  int t = acc_type[d];
  if (t >= 0 && t < T) {  // calculateUsedGPUs default_limiter.go:72-81 (raw GPUsPerReplica)
    long long u = (long long)current[d] * gpr[d];
    if (u) atomicAdd((unsigned long long*)&used[t], (unsigned long long)u);
  }
  is_cand[d] = target[d] > current[d] ? 1 : 0;   // filterScaleUpCandidates :50-58
  out_target[d] = target[d];
  out_gpus[d] = 0;
  out_limited[d] = 0;
}

// keys for the compacted candidate list
__global__ void __launch_bounds__(256) limiter_keys_kernel(int n, int T, const int* cand_idx, const int* acc_type,
                                                           const double* spare, const double* cost,
                                                           unsigned long long* k_cost, unsigned long long* k_spare) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = cand_idx[i];
  k_cost[i] = sortable_f64(cost[d]);
  k_spare[i] = sortable_f64(spare[d]);
}
__global__ void __launch_bounds__(256) limiter_gather_kernel(int n, int T, const int* order, const int* acc_type,
                                                             const int* current, const int* target, const int* gpr,
                                                             const double* spare, unsigned long long* k_spare_sorted,
                                                             int* k_type, long long* req) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = order[i];
  if (k_spare_sorted) k_spare_sorted[i] = sortable_f64(spare[d]);
  if (k_type) {
    int t = acc_type[d];
    k_type[i] = (t >= 0 && t < T) ? t : T;     // AcceleratorName == "" -> own segment, never allocates
  }
  if (req) {
    int g = gpr[d]; if (g <= 0) g = 1;          // greedy_saturation_algorithm.go:85-88
    req[i] = (long long)(target[d] - current[d]) * g;
  }
}

// allocateForDecision (greedy_saturation_algorithm.go:79-108) with TryAllocate folded in
__global__ void __launch_bounds__(256) limiter_apply_kernel(int n, int T, const int* order, const int* k_type,
                                                            const long long* req, const long long* prefix,
                                                            const long long* used, const int* type_limit,
                                                            const int* current, const int* target, const int* gpr,
                                                            int* out_target, int* out_gpus, unsigned char* out_limited) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = order[i];
  int t = k_type[i];
  int g = gpr[d]; if (g <= 0) g = 1;
  long long needed = (long long)target[d] - current[d];
  long long allocated = 0;
  if (t < T) {
    long long avail0 = (long long)type_limit[t] - used[t];      // CreateAllocator type_inventory.go:231-236
    if (avail0 < 0) avail0 = 0;
    long long remaining = avail0 - prefix[i];
    if (remaining < 0) remaining = 0;
    allocated = req[i] < remaining ? req[i] : remaining;         // TryAllocate :358-366
  }
  long long replicas = allocated / g;
  out_gpus[d] = (int)(replicas * g);
  out_target[d] = current[d] + (int)replicas;
  out_limited[d] = replicas < needed ? 1 : 0;
}

}  // namespace wva
