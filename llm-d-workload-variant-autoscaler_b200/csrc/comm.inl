// comm.inl — multi-GPU plumbing of the C-ABI (included at the end of capi.cu): one NCCL communicator per group of
// contexts, the model-sharded exchange steps of wva_solve / wva_saturation_run, and wva_group (one host process
// driving several GPUs with one host thread per device).
//
// The path shards by server.  What crosses NVLink, all of it directly between the device arenas:
//   limited capacity   ONE grouped in-place ncclAllGather over the ten candidate arrays (37 B + 4 B per pair) — the
//                      greedy sweep needs every server (pkg/solver/greedy.go:35-105) — then the sweep on every rank;
//   unlimited          the rank solves its own block (SolveUnlimited is per server, solver.go:63-79); ONE grouped
//                      in-place ncclAllGather of the ten solution arrays (37 B per server) and ONE ncclAllReduce(sum)
//                      pair of the by-type partials {count int64[T], cost float64[T]} (system.go:271-299);
//   V1 saturation      ONE ncclAllReduce(sum) of the four int64 partials.
// Every exchange starts with a 2-word ncclAllReduce(max) of the ranks' status so that a rank whose sizing failed
// makes EVERY rank return an error instead of leaving the others inside a collective.
//
// NCCL is resolved with dlopen at first use (libnccl.so.2; when the host process already carries a copy — e.g. the one
// bundled with torch — the loader hands back that one): a single-GPU caller never needs the library.
#include <dlfcn.h>
#include <nccl.h>
#include <thread>

namespace {

struct NcclApi {
  bool tried = false, ok = false;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (api.tried) return api;
  api.tried = true;
  void* h = nullptr;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { api.err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return api; }
#define WVA_NCCL_SYM(field, sym)                                                          \
  api.field = (decltype(api.field))dlsym(h, #sym);                                        \
  if (!api.field) { api.err = "libnccl: missing symbol " #sym; return api; }
  WVA_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
  WVA_NCCL_SYM(CommInitRank, ncclCommInitRank)
  WVA_NCCL_SYM(CommDestroy, ncclCommDestroy)
  WVA_NCCL_SYM(AllGather, ncclAllGather)
  WVA_NCCL_SYM(AllReduce, ncclAllReduce)
  WVA_NCCL_SYM(GroupStart, ncclGroupStart)
  WVA_NCCL_SYM(GroupEnd, ncclGroupEnd)
  WVA_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef WVA_NCCL_SYM
  api.ok = true;
  return api;
}

#define NK(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r__ = (call);                                                                    \
    if (r__ != ncclSuccess) {                                                                     \
      ctx->last_error = std::string(#call) + ": " + nccl_api().GetErrorString(r__);               \
      return WVA_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

// status agreement: max over ranks of (my status) -> everyone learns whether any rank failed
int32_t comm_agree(wva_ctx* ctx, int my_status, int* agreed) {
  NcclApi& n = nccl_api();
  CK(ctx->comm_ws.reserve(256));
  int* d = (int*)ctx->comm_ws.p;
  CK(cudaMemcpyAsync(d, &my_status, 4, cudaMemcpyHostToDevice, ctx->stream));
  NK(n.AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
  CK(cudaMemcpyAsync(agreed, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WVA_OK;
}

}  // namespace
static int32_t comm_agree_status(wva_ctx* ctx, int my_status, int* agreed) { return comm_agree(ctx, my_status, agreed); }

static void comm_release(wva_ctx* ctx) {
  if (ctx->comm && nccl_api().ok) nccl_api().CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr; ctx->world = 1; ctx->rank = 0;
}

// wva_solve on a context with a communicator (see the header of this file)
static int32_t comm_exchange_and_solve(wva_ctx* ctx) {
  NcclApi& n = nccl_api();
  const ncclComm_t comm = (ncclComm_t)ctx->comm;
  int agreed = 0;
  int32_t rc = comm_agree(ctx, ctx->calculated ? ctx->shard_status : WVA_ERR_STATE, &agreed);
  if (rc != WVA_OK) return rc;
  if (agreed != WVA_OK) {
    if (ctx->shard_status == WVA_OK && ctx->calculated) ctx->last_error = "wva_solve: the sizing of another rank failed";
    return agreed;
  }
  const size_t rows = (size_t)ctx->shard_rows, A = (size_t)ctx->A, T = (size_t)ctx->T;
  const size_t cnt = rows * A;                 // pairs per rank block (the last block may be partly padding)
  if (!ctx->unlimited) {
    // greedy needs every server: all-gather the candidate arena in place, then the same sweep on every rank
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (cnt > 0) {
      const CandView& c = ctx->cand;
      NK(n.GroupStart());
      NK(n.AllGather(c.state + ctx->rank * cnt, c.state, cnt, ncclUint8, comm, ctx->stream));
      void* f[9] = {c.num_replicas, c.batch_size, c.cost, c.value, c.itl, c.ttft, c.rho, c.max_arrv_rate, c.n_solves};
      for (void* q : f) NK(n.AllGather((char*)q + (size_t)ctx->rank * cnt * 4, q, cnt, ncclInt32, comm, ctx->stream));   // 4-byte words
      NK(n.GroupEnd());
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = solve_view(ctx, ctx->sys, ctx->cand, ctx->sol);
    if (rc != WVA_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->timing.exchange_ms = elapsed(ctx, 0, 1);
    return WVA_OK;
  }
  // unlimited: per-server argmin on the rank's block, then all-gather of the solution + all-reduce of the partials
  rc = solve_view(ctx, shard_sys(ctx), shard_cand(ctx), shard_sol(ctx));
  if (rc != WVA_OK) return rc;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  const SolView& o = ctx->sol;
  NK(n.GroupStart());
  if (rows > 0) {
    NK(n.AllGather(o.state + (size_t)ctx->rank * rows, o.state, rows, ncclUint8, comm, ctx->stream));
    void* f[9] = {o.acc, o.num_replicas, o.batch_size, o.cost, o.value, o.itl, o.ttft, o.rho, o.max_arrv_rate};
    for (void* q : f) NK(n.AllGather((char*)q + (size_t)ctx->rank * rows * 4, q, rows, ncclInt32, comm, ctx->stream));
  }
  if (T > 0) {
    NK(n.AllReduce(ctx->d_type_count, ctx->d_type_count, T, ncclInt64, ncclSum, comm, ctx->stream));
    NK(n.AllReduce(ctx->d_type_cost, ctx->d_type_cost, T, ncclDouble, ncclSum, comm, ctx->stream));
  }
  NK(n.GroupEnd());
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.exchange_ms = elapsed(ctx, 0, 1);
  return WVA_OK;
}

// partials of the V1 saturation run summed over the ranks (device, in place on out[0..3] -> all[0..3])
static int32_t comm_reduce_sat_partials(wva_ctx* ctx, long long* d_partials, long long* d_all) {
  CK(cudaMemcpyAsync(d_all, d_partials, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->world > 1) {
    NcclApi& n = nccl_api();
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    NK(n.AllReduce(d_all, d_all, 4, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  }
  return WVA_OK;
}

extern "C" int32_t wva_comm_unique_id(uint8_t id[WVA_COMM_ID_BYTES]) {
  if (!id) return WVA_ERR_ARG;
  NcclApi& n = nccl_api();
  if (!n.ok) return WVA_ERR_NO_DEVICE;
  static_assert(sizeof(ncclUniqueId) == WVA_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  if (n.GetUniqueId(&u) != ncclSuccess) return WVA_ERR_CUDA;
  memcpy(id, &u, sizeof(u));
  return WVA_OK;
}

extern "C" int32_t wva_comm_init_rank(wva_ctx* ctx, int32_t world, int32_t rank, const uint8_t id[WVA_COMM_ID_BYTES]) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return WVA_ERR_ARG;
  if (ctx->comm) { ctx->last_error = "wva_comm_init_rank: the context already has a communicator"; return WVA_ERR_STATE; }
  NcclApi& n = nccl_api();
  if (!n.ok) { ctx->last_error = n.err; return WVA_ERR_NO_DEVICE; }
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  NK(n.CommInitRank(&comm, world, u, rank));
  ctx->comm = comm; ctx->world = world; ctx->rank = rank;
  ctx->loaded = false; ctx->calculated = false; ctx->solved = false;   // arenas are sized per communicator
  return WVA_OK;
}

extern "C" int32_t wva_comm_shard(const wva_ctx* ctx, int32_t* lo, int32_t* hi) {
  if (!ctx || !lo || !hi) return WVA_ERR_ARG;
  *lo = ctx->shard_lo; *hi = ctx->shard_hi;
  return WVA_OK;
}

// ------------------------------------------------------------------ one process, several GPUs
struct wva_group {
  std::vector<wva_ctx*> ctx;
  PinBuf stage;     // host staging for the block outputs of wva_group_saturation_v1
};

namespace {
template <class F>
int32_t on_every_device(wva_group* g, F f) {     // one host thread per device; the first non-zero status wins
  const int n = (int)g->ctx.size();
  std::vector<int32_t> rc(n, WVA_OK);
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++) th.emplace_back([&, i] { rc[i] = f(i, g->ctx[i]); });
  rc[0] = f(0, g->ctx[0]);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++) if (rc[i] != WVA_OK) return rc[i];
  return WVA_OK;
}
}  // namespace

extern "C" int32_t wva_group_create(const int32_t* devices, int32_t n, wva_group** out) {
  if (!devices || !out || n < 1) return WVA_ERR_ARG;
  *out = nullptr;
  wva_group* g = new (std::nothrow) wva_group();
  if (!g) return WVA_ERR_NOMEM;
  for (int i = 0; i < n; i++) {
    wva_ctx* c = nullptr;
    int32_t rc = wva_create(devices[i], &c);
    if (rc != WVA_OK) { wva_group_destroy(g); return rc; }
    g->ctx.push_back(c);
  }
  if (n > 1) {
    uint8_t id[WVA_COMM_ID_BYTES];
    int32_t rc = wva_comm_unique_id(id);
    if (rc == WVA_OK) rc = on_every_device(g, [&](int i, wva_ctx* c) { return wva_comm_init_rank(c, n, i, id); });
    if (rc != WVA_OK) { wva_group_destroy(g); return rc; }
  }
  *out = g;
  return WVA_OK;
}

extern "C" int32_t wva_group_destroy(wva_group* g) {
  if (!g) return WVA_ERR_ARG;
  for (wva_ctx* c : g->ctx) wva_destroy(c);
  g->stage.release();
  delete g;
  return WVA_OK;
}

extern "C" int32_t wva_group_size(const wva_group* g) { return g ? (int32_t)g->ctx.size() : 0; }
extern "C" wva_ctx* wva_group_ctx(wva_group* g, int32_t i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

// Manager.Optimize over the group: replicated load, sharded sizing, exchange, allocator; solution from device 0
extern "C" int32_t wva_group_optimize(wva_group* g, const wva_system* sys, wva_solution* out) {
  if (!g || !sys || !out) return WVA_ERR_ARG;
  int32_t rc = on_every_device(g, [&](int, wva_ctx* c) {
    int32_t r = wva_load_system(c, sys);
    if (r != WVA_OK) return r;
    const int32_t rcalc = wva_calculate(c);
    const int32_t rsolve = wva_solve(c);          // always entered: its first step agrees on the ranks' status
    return rcalc != WVA_OK ? rcalc : rsolve;
  });
  if (rc != WVA_OK) return rc;
  return wva_get_solution(g->ctx[0], out);
}

// V1 saturation over the group: device i analyses the i-th contiguous block of models; outputs land at the block's
// offsets of the caller's arrays; partials_all = the all-reduced partials
extern "C" int32_t wva_group_saturation_v1(wva_group* g, const wva_saturation_in* in, const wva_saturation_out* out) {
  if (!g || !in || !out) return WVA_ERR_ARG;
  const int n = (int)g->ctx.size();
  const long long M = in->n_models, V = in->n_variants, P = in->n_replicas;
  if (M < 0 || V < 0 || P < 0 || !in->model_variant_off || !in->variant_replica_off) return WVA_ERR_ARG;
  if (!valid_offsets(in->model_variant_off, (size_t)M, (size_t)V) || !valid_offsets(in->variant_replica_off, (size_t)V, (size_t)P))
    return WVA_ERR_ARG;
  const long long per = (M + n - 1) / n;
  const bool detail = out->var_replica_count || out->var_non_saturated || out->var_max_kv || out->var_max_queue ||
                      out->var_avg_spare_kv || out->var_avg_spare_queue || out->rep_saturated || out->mod_total_replicas ||
                      out->mod_non_saturated || out->mod_avg_spare_kv || out->mod_avg_spare_queue;
  std::vector<std::vector<int32_t>> mvo(n), vro(n);
  std::vector<int64_t> part(n * 4, 0), part_all(n * 4, 0);
  int32_t rc = on_every_device(g, [&](int i, wva_ctx* c) {
    const long long m0 = std::min(M, per * i), m1 = std::min(M, per * (i + 1));
    const long long v0 = in->model_variant_off[m0], v1 = in->model_variant_off[m1];
    const long long p0 = in->variant_replica_off[v0], p1 = in->variant_replica_off[v1];
    // rebased CSR offsets of the block
    mvo[i].resize(m1 - m0 + 1); vro[i].resize(v1 - v0 + 1);
    for (long long m = m0; m <= m1; m++) mvo[i][m - m0] = (int32_t)(in->model_variant_off[m] - v0);
    for (long long v = v0; v <= v1; v++) vro[i][v - v0] = (int32_t)(in->variant_replica_off[v] - p0);
    wva_saturation_in b = *in;
    b.n_models = m1 - m0; b.n_variants = v1 - v0; b.n_replicas = p1 - p0;
    b.model_variant_off = mvo[i].data(); b.variant_replica_off = vro[i].data();
    b.rep_kv = in->rep_kv + p0; b.rep_queue = in->rep_queue + p0;
    b.var_cost = in->var_cost + v0; b.var_current = in->var_current + v0; b.var_desired = in->var_desired + v0;
    b.var_pending = in->var_pending + v0; b.var_has_state = in->var_has_state ? in->var_has_state + v0 : nullptr;
    b.cfg_kv_threshold = in->cfg_kv_threshold + m0; b.cfg_queue_threshold = in->cfg_queue_threshold + m0;
    b.cfg_kv_trigger = in->cfg_kv_trigger + m0; b.cfg_queue_trigger = in->cfg_queue_trigger + m0;
    int32_t r = wva_saturation_upload(c, &b);
    const int32_t r2 = wva_saturation_run(c, detail ? 1 : 0);   // always entered (the all-reduce inside it)
    if (r == WVA_OK) r = r2;
    if (r != WVA_OK) return r;
    wva_saturation_out o = {};
#define OFFS(field, base) o.field = out->field ? out->field + (base) : nullptr;
    OFFS(var_target, v0) OFFS(var_replica_count, v0) OFFS(var_non_saturated, v0) OFFS(var_max_kv, v0) OFFS(var_max_queue, v0)
    OFFS(var_avg_spare_kv, v0) OFFS(var_avg_spare_queue, v0) OFFS(rep_saturated, p0) OFFS(mod_total_replicas, m0)
    OFFS(mod_non_saturated, m0) OFFS(mod_avg_spare_kv, m0) OFFS(mod_avg_spare_queue, m0) OFFS(mod_flags, m0)
#undef OFFS
    o.partials = &part[i * 4]; o.partials_all = &part_all[i * 4];
    return wva_saturation_fetch(c, &o);
  });
  if (rc != WVA_OK) return rc;
  if (out->partials) for (int k = 0; k < 4; k++) { int64_t s = 0; for (int i = 0; i < n; i++) s += part[i * 4 + k]; out->partials[k] = s; }
  if (out->partials_all) for (int k = 0; k < 4; k++) out->partials_all[k] = part_all[k];
  return WVA_OK;
}
