// greedy_solve.cuh — Solver.SolveGreedy (pkg/solver/greedy.go:35-341) on the device.
//
//   K1 greedy_prepare_kernel   (thread per server)  sort each server's candidates by value
//                              (slices.SortFunc greedy.go:61-63, stable = ascending accelerator
//                              index on ties), first delta (greedy.go:64-70)
//   sort                       entries by (priority asc, delta desc, value desc), greedy.go:76-87:
//                              three stable LSD radix passes (CUB) carrying the server index, so ties
//                              keep ascending server index — the oracle's canonical order
//   K2 greedy_allocate_kernel  allocate() + bestEffort() (greedy.go:107-316).  The reference keeps a
//                              sorted slice and re-inserts a bumped entry BEFORE equal elements
//                              (slices.BinarySearchFunc + slices.Insert, greedy.go:161-162).  That is
//                              the order (key asc; among equal keys re-inserted entries first, latest
//                              first; then the untouched entries in their sorted order), realised
//                              here as: the sorted array consumed from its head + a 32-ary heap of
//                              re-inserted entries keyed (key, insertion stamp desc); the next entry
//                              is the heap top when top <= head, else the head.
//
// K2 is inherently sequential (every fit test depends on all earlier takes of the type): one warp runs
// the sweep, the heap is 32-ary and its pops are warp-cooperative.  Chunked fast-forward over runs of
// fitting entries is next-round work.
#pragma once
#include "wva_core.cuh"
#include "solve_kernels.cuh"
#include <cub/cub.cuh>

namespace wva {

// Go cmp.Compare on float32 (NaN sorts first)
__device__ __forceinline__ int cmp_f32(float x, float y) {
  bool xn = x != x, yn = y != y;
  if (xn) return yn ? 0 : -1;
  if (yn) return 1;
  return x < y ? -1 : (x > y ? 1 : 0);
}

// order-preserving float32 -> uint32 with NaN first and -0 == +0; descending order = bitwise not
__device__ __forceinline__ unsigned sortable_f32(float x) {
  if (x != x) return 0u;
  if (x == 0.0f) x = 0.0f;
  unsigned b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct GreedyWs {   // device workspace views
  int* order;        // [S*A] accelerator index of the k-th cheapest candidate of a server
  int* ncand;        // [S]
  int* cur_idx;      // [S]
  unsigned* k_val;   // [S] sort keys (ping)
  unsigned* k_val2;  // [S] (pong)
  int* e_srv;        // [S] entries (ping)
  int* e_srv2;       // [S] (pong)
  unsigned char* flag;  // [S] server has candidates
  int* n_entries;    // [1]
  // heap of re-inserted entries
  int* h_srv; unsigned* h_tau; float* h_delta; float* h_value; int* h_prio;
  int* unalloc;      // [S]
  // allocateEqually tickets
  unsigned char* t_live; unsigned char* t_active; unsigned char* t_alloc; int* t_type; int* t_upr; int* t_nrep; int* t_final;
  long long* avail;  // [T]
};

__global__ void __launch_bounds__(128) greedy_prepare_kernel(SysView s, CandView c, GreedyWs w) {
  int srv = blockIdx.x * blockDim.x + threadIdx.x;
  if (srv >= s.n_servers) return;
  const int A = s.n_acc;
  int* ord = w.order + (size_t)srv * A;
  int n = 0;
  // insertion sort, stable: ascending accelerator index among equal values
  for (int a = 0; a < A; a++) {
    if (c.state[(size_t)srv * A + a] == ALLOC_NONE) continue;
    float v = c.value[(size_t)srv * A + a];
    int k = n;
    while (k > 0 && cmp_f32(c.value[(size_t)srv * A + ord[k - 1]], v) > 0) { ord[k] = ord[k - 1]; k--; }
    ord[k] = a;
    n++;
  }
  w.ncand[srv] = n;
  w.cur_idx[srv] = 0;
  w.flag[srv] = n > 0 ? 1 : 0;
}

// keys of the compacted entry list for one radix pass: which = 0 value (desc), 1 delta (desc), 2 priority (asc)
__global__ void __launch_bounds__(256) greedy_keys_kernel(SysView s, CandView c, GreedyWs w, const int* e_srv, int n,
                                                         int which, unsigned* keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int srv = e_srv[i];
  const int A = s.n_acc;
  const int* ord = w.order + (size_t)srv * A;
  float v0 = c.value[(size_t)srv * A + ord[0]];
  if (which == 0) keys[i] = ~sortable_f32(v0);
  else if (which == 1) {
    float d = (w.ncand[srv] > 1) ? f_sub(c.value[(size_t)srv * A + ord[1]], v0) : FLT_MAX;   // greedy.go:64-70
    keys[i] = ~sortable_f32(d);
  } else keys[i] = (unsigned)s.srv_priority[srv] ^ 0x80000000u;
}

struct GEntry { int prio; float delta, value; unsigned tau; int srv; };

// serverEntriesOrder (greedy.go:76-87)
__device__ __forceinline__ int g_order(const GEntry& a, const GEntry& b) {
  if (a.prio == b.prio) {
    if (a.delta == b.delta) return cmp_f32(b.value, a.value);
    return cmp_f32(b.delta, a.delta);
  }
  return a.prio < b.prio ? -1 : 1;
}
// heap order: key, then the latest insertion first
__device__ __forceinline__ bool g_before(const GEntry& a, const GEntry& b) {
  int o = g_order(a, b);
  return o < 0 || (o == 0 && a.tau > b.tau);
}

// 32-ary min-heap of re-inserted entries, operated by the whole warp: every lane runs the same control
// flow on the same values (the sweep itself is sequential), and a pop inspects the 32 children of a node
// with one coalesced load per field + a shuffle arg-min, so a heap of S entries is ~log32(S) <= 4 levels
// deep instead of the 17 dependent levels of a binary heap.
struct GHeap {
  GreedyWs w; int n;
  __device__ GEntry get(int i) const { GEntry e; e.prio = w.h_prio[i]; e.delta = w.h_delta[i]; e.value = w.h_value[i]; e.tau = w.h_tau[i]; e.srv = w.h_srv[i]; return e; }
  __device__ void put(int i, const GEntry& e) {
    if ((threadIdx.x & 31) == 0) { w.h_prio[i] = e.prio; w.h_delta[i] = e.delta; w.h_value[i] = e.value; w.h_tau[i] = e.tau; w.h_srv[i] = e.srv; }
  }
  __device__ void push(const GEntry& e) {
    int i = n++;
    while (i > 0) {
      int p = (i - 1) >> 5;
      GEntry pe = get(p);
      if (!g_before(e, pe)) break;
      put(i, pe);
      i = p;
    }
    put(i, e);
    __syncwarp();
  }
  __device__ GEntry pop() {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    GEntry top = get(0);
    GEntry last = get(--n);
    int i = 0;
    while (true) {
      const int c0 = 32 * i + 1;
      if (c0 >= n) break;
      const int ci = c0 + lane;
      const bool have = ci < n;
      GEntry me;
      if (have) me = get(ci);
      else { me.prio = 0x7fffffff; me.delta = 0.0f; me.value = 0.0f; me.tau = 0; me.srv = -1; }
      int who = lane;
      for (int o = 16; o; o >>= 1) {
        GEntry ot;
        ot.prio = __shfl_xor_sync(full, me.prio, o); ot.delta = __shfl_xor_sync(full, me.delta, o);
        ot.value = __shfl_xor_sync(full, me.value, o); ot.tau = __shfl_xor_sync(full, me.tau, o);
        ot.srv = __shfl_xor_sync(full, me.srv, o);
        const int ow = __shfl_xor_sync(full, who, o);
        // total order for the butterfly: heap order, then the lower lane (equal entries cannot occur: tau is unique)
        const bool take = (ot.srv >= 0) && (me.srv < 0 || g_before(ot, me) || (!g_before(me, ot) && ow < who));
        if (take) { me = ot; who = ow; }
      }
      if (me.srv < 0 || !g_before(me, last)) break;
      put(i, me);
      i = c0 + who;
    }
    if (n > 0) put(i, last);
    __syncwarp();
    return top;
  }
};

__device__ __forceinline__ long long g_upr(const SysView& s, int srv, int acc) {   // greedy.go:139
  return (long long)num_instances(s, s.srv_model[srv], acc) * s.acc_multiplicity[acc];
}

__device__ __forceinline__ void g_commit(const SysView& s, const CandView& c, const SolView& o, int srv, int acc,
                                         int replicas, float cost, float value) {
  size_t i = (size_t)srv * s.n_acc + acc;
  o.state[srv] = ALLOC_ACC; o.acc[srv] = acc; o.num_replicas[srv] = replicas; o.batch_size[srv] = c.batch_size[i];
  o.cost[srv] = cost; o.value[srv] = value; o.itl[srv] = c.itl[i]; o.ttft[srv] = c.ttft[i]; o.rho[srv] = c.rho[i];
  o.max_arrv_rate[srv] = c.max_arrv_rate[i];
}

// allocateMaximally (greedy.go:194-223)
__device__ void g_allocate_maximally(const SysView& s, const CandView& c, const SolView& o, const GreedyWs& w,
                                     const int* list, int n) {
  const int A = s.n_acc;
  for (int k = 0; k < n; k++) {
    int srv = list[k];
    if (s.srv_model[srv] < 0) continue;
    const int* ord = w.order + (size_t)srv * A;
    for (int j = 0; j < w.ncand[srv]; j++) {
      int acc = ord[j];
      size_t i = (size_t)srv * A + acc;
      if (c.state[i] != ALLOC_ACC) continue;
      long long upr = g_upr(s, srv, acc);
      if (upr <= 0) continue;
      int t = s.acc_type[acc];
      long long maxr = w.avail[t] / upr;
      int cur = c.num_replicas[i];
      if (maxr > cur) maxr = cur;
      if (maxr > 0) {
        float factor = f_div((float)maxr, (float)cur);
        g_commit(s, c, o, srv, acc, (int)maxr, f_mul(c.cost[i], factor), f_mul(c.value[i], factor));
        w.avail[t] -= maxr * upr;
        break;
      }
    }
  }
}

// allocateEqually (greedy.go:239-316) over list[0..n)
__device__ void g_allocate_equally(const SysView& s, const CandView& c, const SolView& o, const GreedyWs& w,
                                   const int* list, int n) {
  const int A = s.n_acc;
  int live = 0;
  for (int k = 0; k < n; k++) {
    int srv = list[k];
    bool ok = s.srv_model[srv] >= 0;
    w.t_live[srv] = ok ? 1 : 0; w.t_active[srv] = 0; w.t_alloc[srv] = 0; w.t_nrep[srv] = 0;
    if (ok) live++;
  }
  while (live > 0) {
    for (int k = 0; k < n; k++) {
      int srv = list[k];
      if (!w.t_live[srv]) continue;
      if (!w.t_active[srv]) {
        const int* ord = w.order + (size_t)srv * A;
        for (int j = 0; j < w.ncand[srv]; j++) {
          int acc = ord[j];
          if (c.state[(size_t)srv * A + acc] != ALLOC_ACC) continue;
          long long upr = g_upr(s, srv, acc);
          if (upr > 0 && w.avail[s.acc_type[acc]] >= upr) {
            w.t_active[srv] = 1; w.t_type[srv] = s.acc_type[acc]; w.t_upr[srv] = (int)upr; w.t_final[srv] = acc;
            break;
          }
        }
        if (!w.t_active[srv]) { w.t_live[srv] = 0; live--; continue; }
      }
      long long ra = w.avail[w.t_type[srv]] / w.t_upr[srv];
      int want = c.num_replicas[(size_t)srv * A + w.t_final[srv]];
      long long allocatable = ra < want ? ra : want;
      if (allocatable > 0) { w.t_nrep[srv]++; w.avail[w.t_type[srv]] -= w.t_upr[srv]; w.t_alloc[srv] = 1; }
      else { w.t_live[srv] = 0; live--; }
    }
  }
  for (int k = 0; k < n; k++) {
    int srv = list[k];
    if (!w.t_alloc[srv]) continue;
    int acc = w.t_final[srv];
    size_t i = (size_t)srv * A + acc;
    float factor = f_div((float)w.t_nrep[srv], (float)c.num_replicas[i]);
    g_commit(s, c, o, srv, acc, w.t_nrep[srv], f_mul(c.cost[i], factor), f_mul(c.value[i], factor));
  }
}

// bestEffort (greedy.go:169-192)
__device__ void g_best_effort(const SysView& s, const CandView& c, const SolView& o, const GreedyWs& w,
                              const int* list, int n, int policy) {
  if (policy == 1) g_allocate_maximally(s, c, o, w, list, n);
  else if (policy == 2) {
    int i = 0;
    while (i < n) {   // makePriorityGroups (greedy.go:321-341)
      int j = i + 1;
      while (j < n && s.srv_priority[list[j]] == s.srv_priority[list[i]]) j++;
      g_allocate_equally(s, c, o, w, list + i, j - i);
      i = j;
    }
  } else if (policy == 3) g_allocate_equally(s, c, o, w, list, n);
}

__global__ void greedy_allocate_kernel(SysView s, CandView c, SolView o, GreedyWs w, const int* e_srv, int delayed,
                                       int policy) {
  if (blockIdx.x != 0) return;   // one warp; every lane executes the sweep redundantly, the heap uses all 32
  const bool writer = (threadIdx.x & 31) == 0;
  const int A = s.n_acc;
  const int n0 = *w.n_entries;
  if (writer) for (int t = 0; t < s.n_types; t++) w.avail[t] = s.type_count[t];   // greedy.go:38-39
  __syncwarp();
  GHeap heap; heap.w = w; heap.n = 0;
  unsigned tau = 0;
  int head = 0, n_un = 0, group_un0 = 0;
  int group_prio = n0 > 0 ? s.srv_priority[e_srv[0]] : 0;
  while (true) {
    // non-delayed mode: allocate + bestEffort run per priority group (greedy.go:96-103); the array is
    // sorted by priority and a re-inserted entry keeps its priority, so a group ends when both the
    // heap and the group's stretch of the array are exhausted
    bool head_ok = head < n0 && (delayed || s.srv_priority[e_srv[head]] == group_prio);
    if (!head_ok && heap.n == 0) {
      if (!delayed) {
        if (writer) g_best_effort(s, c, o, w, w.unalloc + group_un0, n_un - group_un0, policy);   // sequential: lane 0
        __syncwarp();
        group_un0 = n_un;
        if (head < n0) { group_prio = s.srv_priority[e_srv[head]]; continue; }
      }
      break;
    }
    GEntry e;
    bool from_heap = false;
    if (head_ok) {
      int srv = e_srv[head];
      const int* ord = w.order + (size_t)srv * A;
      e.srv = srv; e.prio = s.srv_priority[srv]; e.tau = 0;
      e.value = c.value[(size_t)srv * A + ord[0]];
      e.delta = (w.ncand[srv] > 1) ? f_sub(c.value[(size_t)srv * A + ord[1]], e.value) : FLT_MAX;
      if (heap.n > 0) {
        GEntry top = heap.get(0);
        if (g_order(top, e) <= 0) from_heap = true;   // inserted BEFORE equal elements
      }
    } else from_heap = true;
    if (from_heap) e = heap.pop(); else head++;
    const int srv = e.srv;
    if (s.srv_model[srv] < 0) continue;                                // greedy.go:126-129
    const int* ord = w.order + (size_t)srv * A;
    int ci = w.cur_idx[srv];
    int acc = ord[ci];
    size_t i = (size_t)srv * A + acc;
    if (c.state[i] != ALLOC_ACC) continue;                             // accelerator "" -> nil (greedy.go:133-136)
    int t = s.acc_type[acc];
    long long count = (long long)c.num_replicas[i] * g_upr(s, srv, acc);
    if (w.avail[t] >= count) {                                         // greedy.go:143-145
      if (writer) { w.avail[t] -= count; g_commit(s, c, o, srv, acc, c.num_replicas[i], c.cost[i], c.value[i]); }
      __syncwarp();
    } else {
      ci++;
      if (writer) w.cur_idx[srv] = ci;
      __syncwarp();
      const int n = w.ncand[srv];
      if (policy == 0 && ci < n) {
        // Policy None: bestEffort() is a no-op, so an entry that can no longer be satisfied changes
        // nothing whenever its remaining candidates are tried.  `available` only decreases, hence a
        // candidate that does not fit NOW never will: if none of the remaining candidates fits now the
        // entry is dropped at once (the lanes test the candidates in parallel).  Exact for this policy;
        // for the others the order of the unallocated list matters and the literal sweep is kept.
        bool fits = false;
        for (int j0 = ci; j0 < n; j0 += 32) {
          const int j = j0 + (int)(threadIdx.x & 31);
          if (j < n) {
            const int a2 = ord[j];
            const size_t i2 = (size_t)srv * A + a2;
            if (c.state[i2] == ALLOC_ACC) {
              const long long cnt2 = (long long)c.num_replicas[i2] * g_upr(s, srv, a2);
              if (w.avail[s.acc_type[a2]] >= cnt2) fits = true;
            }
          }
        }
        if (!__any_sync(0xffffffffu, fits)) { if (writer) w.unalloc[n_un] = srv; n_un++; __syncwarp(); continue; }
      }
      if (ci + 1 < n) e.delta = f_sub(c.value[(size_t)srv * A + ord[ci + 1]], c.value[(size_t)srv * A + ord[ci]]);
      else if (ci == n) { if (writer) w.unalloc[n_un] = srv; n_un++; __syncwarp(); continue; }
      else e.delta = FLT_MAX;
      e.value = c.value[(size_t)srv * A + ord[ci]];
      e.tau = ++tau;
      heap.push(e);
    }
  }
  if (delayed && writer) g_best_effort(s, c, o, w, w.unalloc, n_un, policy);
}

__global__ void __launch_bounds__(256) greedy_clear_solution_kernel(SolView o, int S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  o.state[i] = ALLOC_NONE; o.acc[i] = -1; o.num_replicas[i] = 0; o.batch_size[i] = 0;
  o.cost[i] = o.value[i] = o.itl[i] = o.ttft[i] = o.rho[i] = o.max_arrv_rate[i] = 0.0f;
}

// host driver; ws/ws_cap: a growable device allocation owned by the ctx
static inline int32_t run_solve_greedy(const SysView& s, const CandView& c, const SolView& o, int delayed, int policy,
                                       void** ws, size_t* ws_cap, cudaStream_t stream, long long* launches) {
  const size_t S = (size_t)s.n_servers, A = (size_t)s.n_acc, T = (size_t)s.n_types;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o2 = off; off = (off + b + 255) & ~(size_t)255; return o2; };
  size_t o_order = take(S * A * 4), o_ncand = take(S * 4), o_cur = take(S * 4), o_k1 = take(S * 4), o_k2 = take(S * 4),
         o_e1 = take(S * 4), o_e2 = take(S * 4), o_flag = take(S), o_ne = take(64), o_hs = take(S * 4), o_ht = take(S * 4),
         o_hd = take(S * 4), o_hv = take(S * 4), o_hp = take(S * 4), o_un = take(S * 4), o_tl = take(S), o_ta = take(S),
         o_tc = take(S), o_tt = take(S * 4), o_tu = take(S * 4), o_tn = take(S * 4), o_tf = take(S * 4), o_av = take(T * 8 + 8);
  size_t tmp = 0, tb = 0;
  cub::CountingInputIterator<int> cnt(0);
  cub::DeviceSelect::Flagged(nullptr, tb, cnt, (unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)S, stream);
  tmp = tb;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, (int)S, 0, 32, stream);
  if (tb > tmp) tmp = tb;
  size_t o_tmp = take(tmp + 256);
  if (off + 256 > *ws_cap) {
    if (*ws) cudaFree(*ws);
    *ws = nullptr; *ws_cap = 0;
    if (cudaMalloc(ws, off + 256) != cudaSuccess) return WVA_ERR_NOMEM;
    *ws_cap = off + 256;
  }
  char* d = (char*)*ws;
  GreedyWs w;
  w.order = (int*)(d + o_order); w.ncand = (int*)(d + o_ncand); w.cur_idx = (int*)(d + o_cur);
  w.k_val = (unsigned*)(d + o_k1); w.k_val2 = (unsigned*)(d + o_k2); w.e_srv = (int*)(d + o_e1); w.e_srv2 = (int*)(d + o_e2);
  w.flag = (unsigned char*)(d + o_flag); w.n_entries = (int*)(d + o_ne);
  w.h_srv = (int*)(d + o_hs); w.h_tau = (unsigned*)(d + o_ht); w.h_delta = (float*)(d + o_hd); w.h_value = (float*)(d + o_hv);
  w.h_prio = (int*)(d + o_hp); w.unalloc = (int*)(d + o_un);
  w.t_live = (unsigned char*)(d + o_tl); w.t_active = (unsigned char*)(d + o_ta); w.t_alloc = (unsigned char*)(d + o_tc);
  w.t_type = (int*)(d + o_tt); w.t_upr = (int*)(d + o_tu); w.t_nrep = (int*)(d + o_tn); w.t_final = (int*)(d + o_tf);
  w.avail = (long long*)(d + o_av);
  void* d_tmp = d + o_tmp;
  const unsigned nb = (unsigned)((S + 255) / 256);
  greedy_clear_solution_kernel<<<nb, 256, 0, stream>>>(o, (int)S);                       // server.RemoveAllocation()
  greedy_prepare_kernel<<<(unsigned)((S + 127) / 128), 128, 0, stream>>>(s, c, w);
  size_t t2 = tmp;
  if (cub::DeviceSelect::Flagged(d_tmp, t2, cnt, w.flag, w.e_srv, w.n_entries, (int)S, stream) != cudaSuccess) return WVA_ERR_CUDA;
  int n = 0;
  if (cudaMemcpyAsync(&n, w.n_entries, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return WVA_ERR_CUDA;
  if (cudaStreamSynchronize(stream) != cudaSuccess) return WVA_ERR_CUDA;
  *launches += 3;
  int* cur = w.e_srv; int* alt = w.e_srv2;
  if (n > 0) {
    const unsigned cb = (unsigned)((n + 255) / 256);
    for (int which = 0; which < 3; which++) {   // LSD: value, then delta, then priority
      greedy_keys_kernel<<<cb, 256, 0, stream>>>(s, c, w, cur, n, which, w.k_val);
      t2 = tmp;
      if (cub::DeviceRadixSort::SortPairs(d_tmp, t2, w.k_val, w.k_val2, cur, alt, n, 0, 32, stream) != cudaSuccess) return WVA_ERR_CUDA;
      int* sw = cur; cur = alt; alt = sw;
      *launches += 2;
    }
  }
  greedy_allocate_kernel<<<1, 32, 0, stream>>>(s, c, o, w, cur, delayed, policy);
  *launches += 1;
  return cudaGetLastError() == cudaSuccess ? WVA_OK : WVA_ERR_CUDA;
}

}  // namespace wva
