// capi_aux.inl — second half of the C-ABI implementation (included at the end of
// capi.cu): replica grid, M/M/1/K leg, V1 saturation, limiter,
// FP64 microbenchmarks.

// ------------------------------------------------------------------ replica grid
template <int WARPS>
static cudaError_t launch_grid(wva_ctx* ctx, int blocks, size_t smem, int R, const GridOut& o, unsigned long long n_pairs,
                               int nmax, GridCounters* ctr, const GridDefer& df) {
  auto k = grid_kernel<WARPS>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<blocks, WARPS * 32, smem, ctx->stream>>>(shard_sys(ctx), R, o, n_pairs, nmax, ctr, df);
  ctx->launches++;
  return cudaGetLastError();
}

// Runs the grid on the resident system and keeps the results in HBM.
// full != 0 materialises ok/ttft/itl/rho/tput [S*A*R]; the frontier [S*A] is always produced.
namespace {
// CSR offsets as every batched entry point takes them: off[0] == 0, non-decreasing, off[n] == total
inline bool valid_offsets(const int32_t* off, size_t n, size_t total) {
  if (!off || off[0] != 0 || (size_t)off[n] != total) return false;
  for (size_t i = 0; i < n; i++) if (off[i + 1] < off[i]) return false;
  return true;
}
}  // namespace

extern "C" int32_t wva_grid_run(wva_ctx* ctx, int32_t R, int32_t full) {
  if (!ctx || R < 1) return WVA_ERR_ARG;
  if (!ctx->loaded) { ctx->last_error = "wva_grid_run before wva_load_system"; return WVA_ERR_STATE; }
  CK(cudaSetDevice(ctx->device));
  GridState& g = ctx->grid;
  const SysView sys_v = shard_sys(ctx);        // with a communicator: the rank's block of servers
  const size_t P = (size_t)sys_v.n_servers * ctx->A, n = P * (size_t)R;
  Layout L;
  size_t o_ctr = L.take(256), o_nmax = L.take(256), o_front = L.take(P * 4);
  size_t o_ok = 0, o_ttft = 0, o_itl = 0, o_rho = 0, o_tput = 0;
  if (full) { o_ok = L.take(n); o_ttft = L.take(n * 4); o_itl = L.take(n * 4); o_rho = L.take(n * 4); o_tput = L.take(n * 4); }
  CK(g.buf.reserve(L.off + 256));
  char* d = (char*)g.buf.p;
  g.ctr = (GridCounters*)(d + o_ctr);
  int* d_nmax = (int*)(d + o_nmax);
  g.view.frontier = (int*)(d + o_front);
  g.view.ok = full ? (unsigned char*)(d + o_ok) : nullptr;
  g.view.ttft = full ? (float*)(d + o_ttft) : nullptr;
  g.view.itl = full ? (float*)(d + o_itl) : nullptr;
  g.view.rho = full ? (float*)(d + o_rho) : nullptr;
  g.view.tput = full ? (float*)(d + o_tput) : nullptr;
  g.R = R; g.full = full != 0; g.ran = false;
  CK(cudaMemsetAsync(d, 0, 512, ctx->stream));
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (P > 0) {
    int blocks = (int)((P + 255) / 256); if (blocks > 4096) blocks = 4096;
    max_batch_kernel<<<blocks, 256, 0, ctx->stream>>>(sys_v, (unsigned long long)P, d_nmax);
    ctx->launches++;
    int nmax = 0;
    CK(cudaMemcpyAsync(&nmax, d_nmax, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (nmax < 1) nmax = 1;
    // per-warp head table: (mu, 1/mu) float64 pairs + the float32 copy = 20 B per state
    const size_t per_warp = (size_t)nmax * 20;
    if (per_warp + 1024 > 200 * 1024) { ctx->last_error = "grid: max batch size above 10188"; return WVA_ERR_LIMIT; }
    // ---- deferral of the near-saturation levels (grid_kernel.cuh): large systems only
    GridDefer df = {};
    unsigned long long* d_next = nullptr;
    unsigned long long* items_alt = nullptr; unsigned char* cls_alt = nullptr; void* d_sort_tmp = nullptr; size_t sort_tmp = 0;
    const int row_stride = (nmax + 31) & ~31;
    {
      const size_t rows_bytes = P * (size_t)row_stride * 4;
      const bool want = ctx->grid_defer == 2 || (ctx->grid_defer == 0 && P >= 20000);
      if (want && R <= 65535 && P < 0x7fffffffull && rows_bytes <= ((size_t)16 << 30)) {
        const unsigned long long cap = (unsigned long long)std::min((size_t)P * (size_t)R, (size_t)P * 16 + 1024);
        cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned char*)nullptr, (unsigned char*)nullptr, (unsigned long long*)nullptr,
                                        (unsigned long long*)nullptr, (int)std::min<unsigned long long>(cap, 0x7fffffffull), 0, 8, ctx->stream);
        Layout D;
        const size_t o_cnt = D.take(256), o_rows = D.take(rows_bytes), o_side = D.take(P * sizeof(GridSide)), o_it = D.take(cap * 8),
                     o_it2 = D.take(cap * 8), o_cl = D.take(cap), o_cl2 = D.take(cap), o_tmp = D.take(sort_tmp + 256);
        if (cap < 0x7fffffffull && g.defer.reserve(D.off + 256) == cudaSuccess) {
          char* dd = (char*)g.defer.p;
          CK(cudaMemsetAsync(dd + o_cnt, 0, 256, ctx->stream));
          df.n_items = (unsigned long long*)(dd + o_cnt); d_next = df.n_items + 1;
          df.rows = (float*)(dd + o_rows); df.side = (GridSide*)(dd + o_side);
          df.items = (unsigned long long*)(dd + o_it); items_alt = (unsigned long long*)(dd + o_it2);
          df.cls = (unsigned char*)(dd + o_cl); cls_alt = (unsigned char*)(dd + o_cl2); d_sort_tmp = dd + o_tmp;
          df.cap = cap; df.row_stride = row_stride; df.thr = 0.6f;
        }
      }
    }
    cudaError_t e;
    if (per_warp * 8 <= 64 * 1024) {
      int per_sm = (int)((200 * 1024) / (per_warp * 8 + 1024)); if (per_sm > 6) per_sm = 6; if (per_sm < 1) per_sm = 1;
      e = launch_grid<8>(ctx, ctx->sm_count * per_sm, per_warp * 8, R, g.view, P, nmax, g.ctr, df);
    } else if (per_warp * 4 + 1024 <= 200 * 1024) {
      int per_sm = (int)((200 * 1024) / (per_warp * 4 + 1024)); if (per_sm > 8) per_sm = 8; if (per_sm < 1) per_sm = 1;
      e = launch_grid<4>(ctx, ctx->sm_count * per_sm, per_warp * 4, R, g.view, P, nmax, g.ctr, df);
    } else {
      e = launch_grid<1>(ctx, ctx->sm_count * 4, per_warp, R, g.view, P, nmax, g.ctr, df);
    }
    if (e != cudaSuccess) { ctx->last_error = std::string("grid launch: ") + cudaGetErrorString(e); return WVA_ERR_CUDA; }
    if (df.rows) {
      unsigned long long n_items = 0;
      CK(cudaMemcpyAsync(&n_items, df.n_items, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      if (n_items > df.cap) n_items = df.cap;          // reservations past the capacity were solved in place
      if (n_items > 0) {
        size_t tb = sort_tmp;
        CK(cub::DeviceRadixSort::SortPairsDescending(d_sort_tmp, tb, df.cls, cls_alt, df.items, items_alt, (int)n_items, 0, 8, ctx->stream));
        CK(cudaFuncSetAttribute(grid_deferred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 32 * 33 * 4));
        grid_deferred_kernel<<<ctx->sm_count * 2, 256, 8 * 2 * 32 * 33 * 4, ctx->stream>>>(R, g.view, df, items_alt, n_items, g.ctr, d_next);
        ctx->launches += 2;
      }
      grid_frontier_fix_kernel<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(g.view.frontier, (unsigned long long)P);
      ctx->launches++;
      CK(cudaGetLastError());
    }
  }
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  GridCounters hc;
  CK(cudaMemcpyAsync(&hc, g.ctr, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.grid_ms = elapsed(ctx, 2, 3);
  ctx->timing.chain_solves = (int64_t)hc.solves;
  ctx->timing.chain_states = (int64_t)hc.states;
  ctx->timing.overflow_pairs = (int64_t)hc.overflow;
  if (getenv("WVA_SIZER_DEBUG"))
    fprintf(stderr, "grid: live %llu, lock-step slots first round %llu, later rounds %llu, deferred pass %llu, rounds %llu\n",
            hc.states, hc.slots0, hc.slots_rest, hc.slots_def, hc.rounds);
  if (hc.limit_hit) { ctx->last_error = "grid: a pair needs a larger max batch size than the launch was built for"; return WVA_ERR_LIMIT; }
  g.ran = true;
  return WVA_OK;
}

extern "C" int32_t wva_grid_fetch(wva_ctx* ctx, uint8_t* ok, float* ttft, float* itl, float* rho, float* tput, int32_t* frontier) {
  if (!ctx) return WVA_ERR_ARG;
  GridState& g = ctx->grid;
  if (!g.ran) { ctx->last_error = "wva_grid_fetch before wva_grid_run"; return WVA_ERR_STATE; }
  if ((ok || ttft || itl || rho || tput) && !g.full) { ctx->last_error = "grid was run without full outputs"; return WVA_ERR_STATE; }
  CK(cudaSetDevice(ctx->device));
  const size_t P = (size_t)(ctx->shard_hi - ctx->shard_lo) * ctx->A, n = P * (size_t)g.R;
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  if (ok && n) CK(cudaMemcpyAsync(ok, g.view.ok, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (ttft && n) CK(cudaMemcpyAsync(ttft, g.view.ttft, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (itl && n) CK(cudaMemcpyAsync(itl, g.view.itl, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (rho && n) CK(cudaMemcpyAsync(rho, g.view.rho, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (tput && n) CK(cudaMemcpyAsync(tput, g.view.tput, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (frontier && P) CK(cudaMemcpyAsync(frontier, g.view.frontier, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.d2h_ms = elapsed(ctx, 6, 7);
  return WVA_OK;
}

extern "C" int32_t wva_analyze_grid(wva_ctx* ctx, int32_t R, uint8_t* ok, float* ttft, float* itl, float* rho, float* tput,
                         int32_t* frontier) {
  int32_t rc = wva_grid_run(ctx, R, (ok || ttft || itl || rho || tput) ? 1 : 0);
  if (rc != WVA_OK) return rc;
  return wva_grid_fetch(ctx, ok, ttft, itl, rho, tput, frontier);
}

// ------------------------------------------------------------------ M/M/1/K leg
extern "C" int32_t wva_mm1k_eval(wva_ctx* ctx, int64_t n, const float* lambda, const float* mu, const int32_t* K, uint8_t* valid,
                      float* avg_resp, float* avg_wait, float* avg_serv, float* avg_num, float* avg_queue,
                      float* throughput, float* rho) {
  if (!ctx || n < 0) return WVA_ERR_ARG;
  if (n == 0) return WVA_OK;
  if (!lambda || !mu || !K || !valid || !avg_resp || !avg_wait || !avg_serv || !avg_num || !avg_queue || !throughput || !rho)
    return WVA_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  Layout L;
  size_t o_l = L.take(n * 4), o_m = L.take(n * 4), o_k = L.take(n * 4);
  CK(ctx->io_in.reserve(L.off + 256));
  Layout O;
  size_t o_v = O.take(n), o_f[7];
  for (int i = 0; i < 7; i++) o_f[i] = O.take(n * 4);
  CK(ctx->io_out.reserve(O.off + 256));
  char* di = (char*)ctx->io_in.p; char* dout = (char*)ctx->io_out.p;
  CK(cudaMemcpyAsync(di + o_l, lambda, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_m, mu, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_k, K, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  mm1k_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      n, (const float*)(di + o_l), (const float*)(di + o_m), (const int*)(di + o_k), (unsigned char*)(dout + o_v),
      (float*)(dout + o_f[0]), (float*)(dout + o_f[1]), (float*)(dout + o_f[2]), (float*)(dout + o_f[3]),
      (float*)(dout + o_f[4]), (float*)(dout + o_f[5]), (float*)(dout + o_f[6]));
  ctx->launches++;
  CK(cudaGetLastError());
  float* dst[7] = {avg_resp, avg_wait, avg_serv, avg_num, avg_queue, throughput, rho};
  CK(cudaMemcpyAsync(valid, dout + o_v, n, cudaMemcpyDeviceToHost, ctx->stream));
  for (int i = 0; i < 7; i++) CK(cudaMemcpyAsync(dst[i], dout + o_f[i], n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WVA_OK;
}

// ------------------------------------------------------------------ V1 saturation
extern "C" int32_t wva_saturation_upload(wva_ctx* ctx, const wva_saturation_in* in) {
  if (!ctx || !in) return WVA_ERR_ARG;
  const long long M = in->n_models, V = in->n_variants, P = in->n_replicas;
  if (M > 0x7fffffffLL) { ctx->last_error = "more than 2^31 - 1 models in one batch"; return WVA_ERR_LIMIT; }
  if (M < 0 || V < 0 || P < 0 || P > 0x7fffffffLL || V > 0x7fffffffLL) return WVA_ERR_ARG;
  if (!in->model_variant_off || !in->variant_replica_off) return WVA_ERR_ARG;
  if (!valid_offsets(in->model_variant_off, (size_t)M, (size_t)V) || !valid_offsets(in->variant_replica_off, (size_t)V, (size_t)P)) {
    ctx->last_error = "model_variant_off / variant_replica_off are not CSR offsets of the given sizes";
    return WVA_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  SatState& st = ctx->sat;
  Layout L;
  size_t o_mvo = L.take((M + 1) * 4), o_vro = L.take((V + 1) * 4), o_kv = L.take(P * 8), o_q = L.take(P * 8),
         o_cost = L.take(V * 8), o_cur = L.take(V * 4), o_des = L.take(V * 4), o_pen = L.take(V * 4), o_hs = L.take(V),
         o_c0 = L.take(M * 8), o_c1 = L.take(M * 8), o_c2 = L.take(M * 8), o_c3 = L.take(M * 8);
  CK(st.in.reserve(L.off + 256));
  char* d = (char*)st.in.p;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  struct { size_t off; const void* src; size_t b; } cp[] = {
      {o_mvo, in->model_variant_off, (size_t)(M + 1) * 4}, {o_vro, in->variant_replica_off, (size_t)(V + 1) * 4},
      {o_kv, in->rep_kv, (size_t)P * 8}, {o_q, in->rep_queue, (size_t)P * 8}, {o_cost, in->var_cost, (size_t)V * 8},
      {o_cur, in->var_current, (size_t)V * 4}, {o_des, in->var_desired, (size_t)V * 4}, {o_pen, in->var_pending, (size_t)V * 4},
      {o_hs, in->var_has_state, in->var_has_state ? (size_t)V : 0},
      {o_c0, in->cfg_kv_threshold, (size_t)M * 8}, {o_c1, in->cfg_queue_threshold, (size_t)M * 8},
      {o_c2, in->cfg_kv_trigger, (size_t)M * 8}, {o_c3, in->cfg_queue_trigger, (size_t)M * 8}};
  for (auto& x : cp) {
    if (x.b == 0) continue;
    if (!x.src) return WVA_ERR_ARG;
    CK(cudaMemcpyAsync(d + x.off, x.src, x.b, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  SatIn& v = st.vin;
  v.n_models = M; v.n_variants = V; v.n_replicas = P;
  v.model_variant_off = (const int*)(d + o_mvo); v.variant_replica_off = (const int*)(d + o_vro);
  v.rep_kv = (const double*)(d + o_kv); v.rep_queue = (const long long*)(d + o_q);
  v.var_cost = (const double*)(d + o_cost); v.var_current = (const int*)(d + o_cur);
  v.var_desired = (const int*)(d + o_des); v.var_pending = (const int*)(d + o_pen);
  v.var_has_state = in->var_has_state ? (const unsigned char*)(d + o_hs) : nullptr;
  v.cfg_kv_threshold = (const double*)(d + o_c0); v.cfg_queue_threshold = (const double*)(d + o_c1);
  v.cfg_kv_trigger = (const double*)(d + o_c2); v.cfg_queue_trigger = (const double*)(d + o_c3);
  // output arena
  Layout O;
  size_t q_part = O.take(64), q_t = O.take(V * 4), q_rc = O.take(V * 4), q_ns = O.take(V * 4), q_mk = O.take(V * 8),
         q_mq = O.take(V * 8), q_ak = O.take(V * 8), q_aq = O.take(V * 8), q_rs = O.take(P), q_mt = O.take(M * 4),
         q_mn = O.take(M * 4), q_mak = O.take(M * 8), q_maq = O.take(M * 8), q_mf = O.take(M);
  CK(st.out.reserve(O.off + 256));
  st.out_bytes = O.off;
  char* o = (char*)st.out.p;
  SatOut& w = st.vout;
  w.partials = (long long*)(o + q_part); w.var_target = (int*)(o + q_t); w.var_replica_count = (int*)(o + q_rc);
  w.var_non_saturated = (int*)(o + q_ns); w.var_max_kv = (double*)(o + q_mk); w.var_max_queue = (long long*)(o + q_mq);
  w.var_avg_spare_kv = (double*)(o + q_ak); w.var_avg_spare_queue = (double*)(o + q_aq);
  w.rep_saturated = (unsigned char*)(o + q_rs); w.mod_total_replicas = (int*)(o + q_mt);
  w.mod_non_saturated = (int*)(o + q_mn); w.mod_avg_spare_kv = (double*)(o + q_mak);
  w.mod_avg_spare_queue = (double*)(o + q_maq); w.mod_flags = (unsigned char*)(o + q_mf);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.h2d_ms = elapsed(ctx, 0, 1);
  st.M = M; st.V = V; st.P = P; st.uploaded = true; st.ran = false;
  return WVA_OK;
}

// detail != 0 also materialises the per-variant / per-replica / per-model analysis fields;
// detail == 0 writes only targets, model flags and the shard partials.
extern "C" int32_t wva_saturation_run(wva_ctx* ctx, int32_t detail) {
  if (!ctx) return WVA_ERR_ARG;
  SatState& st = ctx->sat;
  CK(cudaSetDevice(ctx->device));
  if (ctx->world > 1) {   // every rank learns whether every rank has a batch: nobody is left alone in the all-reduce
    int agreed = 0;
    int32_t rc = comm_agree_status(ctx, st.uploaded ? WVA_OK : WVA_ERR_STATE, &agreed);
    if (rc != WVA_OK) return rc;
    if (agreed != WVA_OK) { ctx->last_error = "wva_saturation_run: a rank has no uploaded batch"; return agreed; }
  }
  if (!st.uploaded) { ctx->last_error = "wva_saturation_run before wva_saturation_upload"; return WVA_ERR_STATE; }
  SatOut w = st.vout;
  if (!detail) {
    w.var_replica_count = nullptr; w.var_non_saturated = nullptr; w.var_max_kv = nullptr; w.var_max_queue = nullptr;
    w.var_avg_spare_kv = nullptr; w.var_avg_spare_queue = nullptr; w.rep_saturated = nullptr;
    w.mod_total_replicas = nullptr; w.mod_non_saturated = nullptr; w.mod_avg_spare_kv = nullptr;
    w.mod_avg_spare_queue = nullptr;
  }
  CK(cudaMemsetAsync(w.partials, 0, 64, ctx->stream));
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (st.M > 0) {
    // per-model descriptors + group geometry (the dependent CSR look-ups), then the streaming kernel: one persistent CTA
    // per SM, every warp its own cp.async.bulk pipeline over groups of consecutive models (saturation_kernel.cuh)
    CK(st.desc.reserve(sat_desc_bytes(st.M)));
    SatDesc* d_desc = (SatDesc*)st.desc.p;
    CK(launch_saturation(detail, ctx->sm_count, st.M, st.vin, w, d_desc, ctx->stream));
    ctx->launches += 2;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  {
    int32_t rc = comm_reduce_sat_partials(ctx, w.partials, w.partials + 4);   // partials_all (one ncclAllReduce)
    if (rc != WVA_OK) return rc;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.saturation_ms = elapsed(ctx, 2, 3);
  ctx->timing.exchange_ms = ctx->world > 1 ? elapsed(ctx, 0, 1) : 0.0f;
  st.ran = true;
  return WVA_OK;
}

extern "C" int32_t wva_saturation_fetch(wva_ctx* ctx, const wva_saturation_out* out) {
  if (!ctx || !out) return WVA_ERR_ARG;
  SatState& st = ctx->sat;
  if (!st.ran) { ctx->last_error = "wva_saturation_fetch before wva_saturation_run"; return WVA_ERR_STATE; }
  CK(cudaSetDevice(ctx->device));
  const size_t M = st.M, V = st.V, P = st.P;
  const SatOut& w = st.vout;
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  struct { void* dst; const void* src; size_t b; } cp[] = {
      {out->var_target, w.var_target, V * 4}, {out->var_replica_count, w.var_replica_count, V * 4},
      {out->var_non_saturated, w.var_non_saturated, V * 4}, {out->var_max_kv, w.var_max_kv, V * 8},
      {out->var_max_queue, w.var_max_queue, V * 8}, {out->var_avg_spare_kv, w.var_avg_spare_kv, V * 8},
      {out->var_avg_spare_queue, w.var_avg_spare_queue, V * 8}, {out->rep_saturated, w.rep_saturated, P},
      {out->mod_total_replicas, w.mod_total_replicas, M * 4}, {out->mod_non_saturated, w.mod_non_saturated, M * 4},
      {out->mod_avg_spare_kv, w.mod_avg_spare_kv, M * 8}, {out->mod_avg_spare_queue, w.mod_avg_spare_queue, M * 8},
      {out->mod_flags, w.mod_flags, M}, {out->partials, w.partials, 32}, {out->partials_all, w.partials + 4, 32}};
  for (auto& x : cp)
    if (x.dst && x.b) CK(cudaMemcpyAsync(x.dst, x.src, x.b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.d2h_ms = elapsed(ctx, 6, 7);
  return WVA_OK;
}

extern "C" int32_t wva_saturation_v1(wva_ctx* ctx, const wva_saturation_in* in, const wva_saturation_out* out) {
  if (!out) return WVA_ERR_ARG;
  int32_t rc = wva_saturation_upload(ctx, in);
  if (rc != WVA_OK) return rc;
  bool detail = out->var_replica_count || out->var_non_saturated || out->var_max_kv || out->var_max_queue ||
                out->var_avg_spare_kv || out->var_avg_spare_queue || out->rep_saturated || out->mod_total_replicas ||
                out->mod_non_saturated || out->mod_avg_spare_kv || out->mod_avg_spare_queue;
  rc = wva_saturation_run(ctx, detail ? 1 : 0);
  if (rc != WVA_OK) return rc;
  return wva_saturation_fetch(ctx, out);
}

// ------------------------------------------------------------------ limiter
extern "C" int32_t wva_limit(wva_ctx* ctx, int64_t D, int32_t T, const int32_t* acc_type, const int32_t* current,
                  const int32_t* target, const int32_t* gpr, const double* spare, const double* cost,
                  const int32_t* type_limit, int32_t* out_target, int32_t* out_gpus, uint8_t* out_limited) {
  if (!ctx || D < 0 || T < 0 || D > 0x7fffffffLL) return WVA_ERR_ARG;
  if (D == 0) return WVA_OK;   // default_limiter.go:43-45
  if (!acc_type || !current || !target || !gpr || !spare || !cost || !out_target || !out_gpus || !out_limited ||
      (T > 0 && !type_limit))
    return WVA_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)D;
  Layout L;
  size_t o_at = L.take(n * 4), o_cur = L.take(n * 4), o_tgt = L.take(n * 4), o_gpr = L.take(n * 4), o_sp = L.take(n * 8),
         o_co = L.take(n * 8), o_lim = L.take((size_t)(T + 1) * 4);
  CK(ctx->io_in.reserve(L.off + 256));
  char* di = (char*)ctx->io_in.p;
  Layout O;
  size_t q_used = O.take((size_t)(T + 1) * 8), q_ncand = O.take(64), q_flag = O.take(n), q_ot = O.take(n * 4), q_og = O.take(n * 4),
         q_ol = O.take(n), q_idxA = O.take(n * 4), q_idxB = O.take(n * 4), q_kA = O.take(n * 8), q_kB = O.take(n * 8),
         q_tA = O.take(n * 4), q_tB = O.take(n * 4), q_req = O.take(n * 8), q_pre = O.take(n * 8);
  size_t tmp_bytes = 0, tb;
  {
    cub::CountingInputIterator<int> cnt(0);
    cub::DeviceSelect::Flagged(nullptr, tb, cnt, (unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, ctx->stream);
    tmp_bytes = tb;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                    (int*)nullptr, (int)n, 0, 64, ctx->stream);
    if (tb > tmp_bytes) tmp_bytes = tb;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, 0, 8,
                                    ctx->stream);
    if (tb > tmp_bytes) tmp_bytes = tb;
    cub::DeviceScan::ExclusiveSumByKey(nullptr, tb, (int*)nullptr, (long long*)nullptr, (long long*)nullptr, (int)n,
                                       cub::Equality(), ctx->stream);
    if (tb > tmp_bytes) tmp_bytes = tb;
  }
  size_t q_tmp = O.take(tmp_bytes + 256);
  CK(ctx->io_out.reserve(O.off + 256));
  char* dd = (char*)ctx->io_out.p;
  CK(cudaMemcpyAsync(di + o_at, acc_type, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_cur, current, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_tgt, target, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_gpr, gpr, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_sp, spare, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di + o_co, cost, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (T > 0) CK(cudaMemcpyAsync(di + o_lim, type_limit, (size_t)T * 4, cudaMemcpyHostToDevice, ctx->stream));
  const int* d_at = (const int*)(di + o_at); const int* d_cur = (const int*)(di + o_cur); const int* d_tgt = (const int*)(di + o_tgt);
  const int* d_gpr = (const int*)(di + o_gpr); const double* d_sp = (const double*)(di + o_sp); const double* d_co = (const double*)(di + o_co);
  const int* d_lim = (const int*)(di + o_lim);
  long long* d_used = (long long*)(dd + q_used); int* d_ncand = (int*)(dd + q_ncand);
  unsigned char* d_flag = (unsigned char*)(dd + q_flag);
  int* d_ot = (int*)(dd + q_ot); int* d_og = (int*)(dd + q_og); unsigned char* d_ol = (unsigned char*)(dd + q_ol);
  int* idxA = (int*)(dd + q_idxA); int* idxB = (int*)(dd + q_idxB);
  unsigned long long* kA = (unsigned long long*)(dd + q_kA); unsigned long long* kB = (unsigned long long*)(dd + q_kB);
  int* tA = (int*)(dd + q_tA); int* tB = (int*)(dd + q_tB);
  long long* d_req = (long long*)(dd + q_req); long long* d_pre = (long long*)(dd + q_pre);
  void* d_tmp = dd + q_tmp;
  CK(cudaMemsetAsync(d_used, 0, (size_t)(T + 1) * 8, ctx->stream));
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  const unsigned nb = (unsigned)((n + 255) / 256);
  limiter_prepare_kernel<<<nb, 256, 0, ctx->stream>>>((long long)n, T, d_at, d_cur, d_tgt, d_gpr, d_used, d_flag, d_ot, d_og, d_ol);
  ctx->launches++;
  {
    cub::CountingInputIterator<int> cnt(0);
    size_t tb2 = tmp_bytes;
    CK(cub::DeviceSelect::Flagged(d_tmp, tb2, cnt, d_flag, idxA, d_ncand, (int)n, ctx->stream));
    ctx->launches++;
  }
  int nc = 0;
  CK(cudaMemcpyAsync(&nc, d_ncand, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (nc > 0) {
    const unsigned cb = (unsigned)((nc + 255) / 256);
    size_t tb2;
    // pass 1: by cost (payload = ascending decision index -> stable tie-break)
    limiter_keys_kernel<<<cb, 256, 0, ctx->stream>>>(nc, T, idxA, d_at, d_sp, d_co, kA, kB);
    tb2 = tmp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, tb2, kA, kB, idxA, idxB, nc, 0, 64, ctx->stream));
    // pass 2: by spare capacity
    limiter_gather_kernel<<<cb, 256, 0, ctx->stream>>>(nc, T, idxB, d_at, d_cur, d_tgt, d_gpr, d_sp, kA, nullptr, nullptr);
    tb2 = tmp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, tb2, kA, kB, idxB, idxA, nc, 0, 64, ctx->stream));
    // pass 3: by accelerator type (segments)
    limiter_gather_kernel<<<cb, 256, 0, ctx->stream>>>(nc, T, idxA, d_at, d_cur, d_tgt, d_gpr, d_sp, nullptr, tA, nullptr);
    tb2 = tmp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, tb2, tA, tB, idxA, idxB, nc, 0, 8, ctx->stream));
    // requested GPUs in final order, segmented exclusive scan, apply
    limiter_gather_kernel<<<cb, 256, 0, ctx->stream>>>(nc, T, idxB, d_at, d_cur, d_tgt, d_gpr, d_sp, nullptr, nullptr, d_req);
    tb2 = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSumByKey(d_tmp, tb2, tB, d_req, d_pre, nc, cub::Equality(), ctx->stream));
    limiter_apply_kernel<<<cb, 256, 0, ctx->stream>>>(nc, T, idxB, tB, d_req, d_pre, d_used, d_lim, d_cur, d_tgt, d_gpr, d_ot, d_og, d_ol);
    ctx->launches += 9;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CK(cudaMemcpyAsync(out_target, d_ot, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_gpus, d_og, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_limited, d_ol, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.limit_ms = elapsed(ctx, 2, 3);
  return WVA_OK;
}

// ------------------------------------------------------------------ FP64 microbenchmarks
__global__ void __launch_bounds__(256) mb_dfma_kernel(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void __launch_bounds__(256) mb_ddiv_kernel(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  const double m = 1.000001;
  for (int i = 0; i < iters; i++) {
    a0 = __ddiv_rn(a0, m); a1 = __ddiv_rn(a1, m); a2 = __ddiv_rn(a2, m); a3 = __ddiv_rn(a3, m);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

// ------------------------------------------------------------------ pinned host buffers for the caller
extern "C" int32_t wva_host_alloc(size_t bytes, void** out) {
  if (!out) return WVA_ERR_ARG;
  *out = nullptr;
  if (bytes == 0) return WVA_OK;
  return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? WVA_OK : WVA_ERR_NOMEM;
}
extern "C" int32_t wva_host_free(void* p) {
  if (!p) return WVA_OK;
  return cudaFreeHost(p) == cudaSuccess ? WVA_OK : WVA_ERR_CUDA;
}

// peak double-precision FMA and IEEE-divide instruction throughput (lane-ops per second)
extern "C" int32_t wva_microbench_fp64(wva_ctx* ctx, double* dfma_per_s, double* ddiv_per_s) {
  if (!ctx || !dfma_per_s || !ddiv_per_s) return WVA_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256;
  CK(ctx->io_out.reserve((size_t)blocks * threads * 8));
  double* d = (double*)ctx->io_out.p;
  const int it_fma = 20000, it_div = 4000;
  float best_fma = 1e30f, best_div = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    mb_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(d, it_fma);
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    mb_ddiv_kernel<<<blocks, threads, 0, ctx->stream>>>(d, it_div);
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 2;
    float a = elapsed(ctx, 2, 3), b = elapsed(ctx, 3, 4);
    if (rep > 0) { if (a < best_fma) best_fma = a; if (b < best_div) best_div = b; }
  }
  *dfma_per_s = (double)blocks * threads * 8.0 * it_fma / (best_fma * 1e-3);
  *ddiv_per_s = (double)blocks * threads * 4.0 * it_div / (best_div * 1e-3);
  return WVA_OK;
}

// ------------------------------------------------------------------ V2 pipeline (pipeline_v2_kernel.cuh)
namespace {
// host arrays -> one device arena, results back: a tiny staging helper for the three batched V2 entry points
struct Stage {
  wva_ctx* ctx; Layout L; std::vector<std::pair<size_t, std::pair<const void*, size_t>>> in;
  size_t add(const void* host, size_t bytes) { size_t o = L.take(bytes ? bytes : 1); if (host && bytes) in.push_back({o, {host, bytes}}); return o; }
  int32_t upload() {
    if (ctx->io_in.reserve(L.off + 256) != cudaSuccess) return WVA_ERR_NOMEM;
    for (auto& e : in)
      if (cudaMemcpyAsync((char*)ctx->io_in.p + e.first, e.second.first, e.second.second, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return WVA_ERR_CUDA;
    return WVA_OK;
  }
  template <class T> const T* at(size_t off, const void* host) const { return host ? (const T*)((char*)ctx->io_in.p + off) : nullptr; }
};
inline unsigned warp_grid(wva_ctx* ctx, long long n_models) {
  long long blocks = (n_models + 7) / 8;                    // 8 warps (models) per 256-thread block
  const long long cap = (long long)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}
}  // namespace

extern "C" int32_t wva_saturation_v2(wva_ctx* ctx, const wva_saturation_v2_in* in, const wva_saturation_v2_out* out) {
  if (!ctx || !in || !out || in->n_models < 0 || in->n_variants < 0 || in->n_replicas < 0) return WVA_ERR_ARG;
  const size_t M = (size_t)in->n_models, V = (size_t)in->n_variants, P = (size_t)in->n_replicas;
  if (M == 0) return WVA_OK;
  if (in->n_variants > 0x7fffffffLL || in->n_replicas > 0x7fffffffLL) return WVA_ERR_LIMIT;
  if (!in->model_variant_off || !in->variant_replica_off || !in->cfg_kv_threshold || !in->cfg_scale_up_threshold ||
      !in->cfg_scale_down_boundary || (V && (!in->var_current || !in->var_pending || !in->var_fallback_capacity)) ||
      (P && (!in->rep_total_kv_tokens || !in->rep_tokens_in_use || !in->rep_queue_length || !in->rep_avg_input_tokens ||
             !in->rep_avg_output_tokens || !in->rep_prefix_hit_rate || !in->rep_k2)) ||
      ((in->sched_queue_size == nullptr) != (in->sched_queue_bytes == nullptr)))
    return WVA_ERR_ARG;
  if (!valid_offsets(in->model_variant_off, M, V) || !valid_offsets(in->variant_replica_off, V, P)) {
    ctx->last_error = "model_variant_off / variant_replica_off are not CSR offsets of the given sizes";
    return WVA_ERR_ARG;
  }
  if (in->rep_slice_order)
    for (size_t m = 0; m < M; m++) {
      const int32_t a = in->variant_replica_off[in->model_variant_off[m]], b = in->variant_replica_off[in->model_variant_off[m + 1]];
      for (int32_t k = a; k < b; k++)
        if (in->rep_slice_order[k] < a || in->rep_slice_order[k] >= b) { ctx->last_error = "rep_slice_order leaves its model"; return WVA_ERR_ARG; }
    }
  CK(cudaSetDevice(ctx->device));
  Stage s{ctx};
  const size_t o_mvo = s.add(in->model_variant_off, (M + 1) * 4), o_vro = s.add(in->variant_replica_off, (V + 1) * 4),
               o_tk = s.add(in->rep_total_kv_tokens, P * 8), o_tu = s.add(in->rep_tokens_in_use, P * 8),
               o_ql = s.add(in->rep_queue_length, P * 8), o_ai = s.add(in->rep_avg_input_tokens, P * 8),
               o_ao = s.add(in->rep_avg_output_tokens, P * 8), o_hr = s.add(in->rep_prefix_hit_rate, P * 8),
               o_k2 = s.add(in->rep_k2, P * 8), o_so = s.add(in->rep_slice_order, P * 4), o_vc = s.add(in->var_current, V * 4),
               o_vp = s.add(in->var_pending, V * 4), o_vf = s.add(in->var_fallback_capacity, V * 8),
               o_kt = s.add(in->cfg_kv_threshold, M * 8), o_su = s.add(in->cfg_scale_up_threshold, M * 8),
               o_sd = s.add(in->cfg_scale_down_boundary, M * 8), o_qs = s.add(in->sched_queue_size, M * 8),
               o_qb = s.add(in->sched_queue_bytes, M * 8);
  int32_t rc = s.upload();
  if (rc != WVA_OK) return rc;
  Layout O;
  const size_t q_k1 = O.take(P * 8), q_ef = O.take(P * 8), q_de = O.take(P * 8), q_sa = O.take(P), q_vr = O.take(V * 4),
               q_vcap = O.take(V * 8), q_vt = O.take(V * 8), q_vd = O.take(V * 8), q_vu = O.take(V * 8), q_ms = O.take(M * 8),
               q_md = O.take(M * 8), q_mu = O.take(M * 8), q_mr = O.take(M * 8), q_mp = O.take(M * 8);
  CK(ctx->io_out.reserve(O.off + 256));
  char* dd = (char*)ctx->io_out.p;
  SatV2In di;
  di.n_models = in->n_models; di.n_variants = in->n_variants; di.n_replicas = in->n_replicas;
  di.model_variant_off = s.at<int>(o_mvo, in->model_variant_off); di.variant_replica_off = s.at<int>(o_vro, in->variant_replica_off);
  di.rep_total_kv = s.at<long long>(o_tk, in->rep_total_kv_tokens); di.rep_tokens_in_use = s.at<long long>(o_tu, in->rep_tokens_in_use);
  di.rep_queue_len = s.at<long long>(o_ql, in->rep_queue_length); di.rep_k2 = s.at<long long>(o_k2, in->rep_k2);
  di.rep_avg_in = s.at<double>(o_ai, in->rep_avg_input_tokens); di.rep_avg_out = s.at<double>(o_ao, in->rep_avg_output_tokens);
  di.rep_hit = s.at<double>(o_hr, in->rep_prefix_hit_rate); di.rep_slice_order = s.at<int>(o_so, in->rep_slice_order);
  di.var_current = s.at<int>(o_vc, in->var_current); di.var_pending = s.at<int>(o_vp, in->var_pending);
  di.var_fallback = s.at<double>(o_vf, in->var_fallback_capacity);
  di.cfg_kv_threshold = s.at<double>(o_kt, in->cfg_kv_threshold); di.cfg_scale_up = s.at<double>(o_su, in->cfg_scale_up_threshold);
  di.cfg_scale_down = s.at<double>(o_sd, in->cfg_scale_down_boundary);
  di.sched_size = s.at<long long>(o_qs, in->sched_queue_size); di.sched_bytes = s.at<long long>(o_qb, in->sched_queue_bytes);
  SatV2Out dout;
  dout.rep_k1 = out->rep_k1 ? (long long*)(dd + q_k1) : nullptr; dout.rep_effective = out->rep_effective ? (long long*)(dd + q_ef) : nullptr;
  dout.rep_demand = out->rep_demand ? (long long*)(dd + q_de) : nullptr; dout.rep_saturated = out->rep_saturated ? (unsigned char*)(dd + q_sa) : nullptr;
  dout.var_ready = out->var_ready ? (int*)(dd + q_vr) : nullptr; dout.var_cap = out->var_per_replica_capacity ? (double*)(dd + q_vcap) : nullptr;
  dout.var_total_cap = out->var_total_capacity ? (double*)(dd + q_vt) : nullptr; dout.var_total_demand = out->var_total_demand ? (double*)(dd + q_vd) : nullptr;
  dout.var_util = out->var_utilization ? (double*)(dd + q_vu) : nullptr;
  dout.mod_supply = out->mod_total_supply ? (double*)(dd + q_ms) : nullptr; dout.mod_demand = out->mod_total_demand ? (double*)(dd + q_md) : nullptr;
  dout.mod_util = out->mod_utilization ? (double*)(dd + q_mu) : nullptr; dout.mod_required = out->mod_required_capacity ? (double*)(dd + q_mr) : nullptr;
  dout.mod_spare = out->mod_spare_capacity ? (double*)(dd + q_mp) : nullptr;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  saturation_v2_kernel<<<warp_grid(ctx, in->n_models), 256, 0, ctx->stream>>>(di, dout);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  auto back = [&](void* host, size_t off, size_t bytes) -> cudaError_t {
    return (host && bytes) ? cudaMemcpyAsync(host, dd + off, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
  };
  CK(back(out->rep_k1, q_k1, P * 8)); CK(back(out->rep_effective, q_ef, P * 8)); CK(back(out->rep_demand, q_de, P * 8));
  CK(back(out->rep_saturated, q_sa, P)); CK(back(out->var_ready, q_vr, V * 4)); CK(back(out->var_per_replica_capacity, q_vcap, V * 8));
  CK(back(out->var_total_capacity, q_vt, V * 8)); CK(back(out->var_total_demand, q_vd, V * 8)); CK(back(out->var_utilization, q_vu, V * 8));
  CK(back(out->mod_total_supply, q_ms, M * 8)); CK(back(out->mod_total_demand, q_md, M * 8)); CK(back(out->mod_utilization, q_mu, M * 8));
  CK(back(out->mod_required_capacity, q_mr, M * 8)); CK(back(out->mod_spare_capacity, q_mp, M * 8));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.saturation_ms = elapsed(ctx, 2, 3);
  return WVA_OK;
}

extern "C" int32_t wva_cost_aware_optimize(wva_ctx* ctx, int64_t n_models, int64_t n_variants, const int32_t* model_variant_off,
                                           const double* mod_required, const double* mod_spare, const uint8_t* mod_has_result,
                                           const int32_t* var_current, const double* var_cost, const double* var_cap,
                                           int32_t* var_target) {
  if (!ctx || n_models < 0 || n_variants < 0) return WVA_ERR_ARG;
  if (n_models == 0) return WVA_OK;
  if (n_variants > 0x7fffffffLL) return WVA_ERR_LIMIT;
  if (!model_variant_off || !mod_required || !mod_spare || (n_variants && (!var_current || !var_cost || !var_cap || !var_target)))
    return WVA_ERR_ARG;
  if (!valid_offsets(model_variant_off, (size_t)n_models, (size_t)n_variants)) { ctx->last_error = "model_variant_off is not a CSR offset array"; return WVA_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const size_t M = (size_t)n_models, V = (size_t)n_variants;
  Stage s{ctx};
  const size_t o_mvo = s.add(model_variant_off, (M + 1) * 4), o_rq = s.add(mod_required, M * 8), o_sp = s.add(mod_spare, M * 8),
               o_hr = s.add(mod_has_result, M), o_cu = s.add(var_current, V * 4), o_co = s.add(var_cost, V * 8), o_ca = s.add(var_cap, V * 8);
  int32_t rc = s.upload();
  if (rc != WVA_OK) return rc;
  CK(ctx->io_out.reserve(V * 4 + 256));
  int* d_t = (int*)ctx->io_out.p;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  cost_aware_kernel<<<warp_grid(ctx, n_models), 256, 0, ctx->stream>>>(n_models, s.at<int>(o_mvo, model_variant_off), s.at<double>(o_rq, mod_required),
      s.at<double>(o_sp, mod_spare), s.at<unsigned char>(o_hr, mod_has_result), s.at<int>(o_cu, var_current), s.at<double>(o_co, var_cost),
      s.at<double>(o_ca, var_cap), d_t);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (V) CK(cudaMemcpyAsync(var_target, d_t, V * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.limit_ms = elapsed(ctx, 2, 3);
  return WVA_OK;
}

extern "C" int32_t wva_enforce(wva_ctx* ctx, int64_t n_models, int64_t n_variants, const int32_t* model_variant_off,
                               const uint8_t* mod_s2z, const double* mod_request_count, const uint8_t* mod_request_error,
                               const double* var_cost, const uint8_t* var_has_cost, int32_t* var_target, uint8_t* mod_applied) {
  if (!ctx || n_models < 0 || n_variants < 0) return WVA_ERR_ARG;
  if (n_models == 0) return WVA_OK;
  if (n_variants > 0x7fffffffLL) return WVA_ERR_LIMIT;
  if (!model_variant_off || !mod_s2z || !mod_request_count || (n_variants && (!var_cost || !var_target))) return WVA_ERR_ARG;
  if (!valid_offsets(model_variant_off, (size_t)n_models, (size_t)n_variants)) { ctx->last_error = "model_variant_off is not a CSR offset array"; return WVA_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const size_t M = (size_t)n_models, V = (size_t)n_variants;
  Stage s{ctx};
  const size_t o_mvo = s.add(model_variant_off, (M + 1) * 4), o_z = s.add(mod_s2z, M), o_rc = s.add(mod_request_count, M * 8),
               o_re = s.add(mod_request_error, M), o_co = s.add(var_cost, V * 8), o_hc = s.add(var_has_cost, V), o_tg = s.add(var_target, V * 4);
  int32_t rc = s.upload();
  if (rc != WVA_OK) return rc;
  CK(ctx->io_out.reserve(M + 256));
  unsigned char* d_app = (unsigned char*)ctx->io_out.p;
  int* d_t = (int*)((char*)ctx->io_in.p + o_tg);
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  enforce_kernel<<<warp_grid(ctx, n_models), 256, 0, ctx->stream>>>(n_models, s.at<int>(o_mvo, model_variant_off), s.at<unsigned char>(o_z, mod_s2z),
      s.at<double>(o_rc, mod_request_count), s.at<unsigned char>(o_re, mod_request_error), s.at<double>(o_co, var_cost),
      s.at<unsigned char>(o_hc, var_has_cost), nullptr, d_t, d_app);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (V) CK(cudaMemcpyAsync(var_target, d_t, V * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mod_applied) CK(cudaMemcpyAsync(mod_applied, d_app, M, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.limit_ms = elapsed(ctx, 2, 3);
  return WVA_OK;
}

// analyzer -> cost-aware optimizer -> enforcer for every model of a cycle: one upload, three launches chained on the
// device (the optimizer reads the analyzer's capacities and signals, the enforcer the optimizer's targets), one download
extern "C" int32_t wva_pipeline_v2(wva_ctx* ctx, const wva_saturation_v2_in* in, const double* var_cost, const int32_t* var_name_rank,
                                   const uint8_t* mod_s2z, const double* mod_request_count, const uint8_t* mod_request_error,
                                   const wva_saturation_v2_out* out, int32_t* var_target, uint8_t* mod_applied) {
  if (!ctx || !in || in->n_models < 0 || in->n_variants < 0 || in->n_replicas < 0) return WVA_ERR_ARG;
  const size_t M = (size_t)in->n_models, V = (size_t)in->n_variants, P = (size_t)in->n_replicas;
  if (M == 0) return WVA_OK;
  if (in->n_variants > 0x7fffffffLL || in->n_replicas > 0x7fffffffLL) return WVA_ERR_LIMIT;
  if (!in->model_variant_off || !in->variant_replica_off || !in->cfg_kv_threshold || !in->cfg_scale_up_threshold ||
      !in->cfg_scale_down_boundary || !mod_s2z || !mod_request_count ||
      (V && (!in->var_current || !in->var_pending || !in->var_fallback_capacity || !var_cost || !var_target)) ||
      (P && (!in->rep_total_kv_tokens || !in->rep_tokens_in_use || !in->rep_queue_length || !in->rep_avg_input_tokens ||
             !in->rep_avg_output_tokens || !in->rep_prefix_hit_rate || !in->rep_k2)) ||
      ((in->sched_queue_size == nullptr) != (in->sched_queue_bytes == nullptr)))
    return WVA_ERR_ARG;
  if (!valid_offsets(in->model_variant_off, M, V) || !valid_offsets(in->variant_replica_off, V, P)) {
    ctx->last_error = "model_variant_off / variant_replica_off are not CSR offsets of the given sizes";
    return WVA_ERR_ARG;
  }
  if (in->rep_slice_order)
    for (size_t m = 0; m < M; m++) {
      const int32_t a = in->variant_replica_off[in->model_variant_off[m]], b = in->variant_replica_off[in->model_variant_off[m + 1]];
      for (int32_t k = a; k < b; k++)
        if (in->rep_slice_order[k] < a || in->rep_slice_order[k] >= b) { ctx->last_error = "rep_slice_order leaves its model"; return WVA_ERR_ARG; }
    }
  CK(cudaSetDevice(ctx->device));
  Stage s{ctx};
  const size_t o_mvo = s.add(in->model_variant_off, (M + 1) * 4), o_vro = s.add(in->variant_replica_off, (V + 1) * 4),
               o_tk = s.add(in->rep_total_kv_tokens, P * 8), o_tu = s.add(in->rep_tokens_in_use, P * 8),
               o_ql = s.add(in->rep_queue_length, P * 8), o_ai = s.add(in->rep_avg_input_tokens, P * 8),
               o_ao = s.add(in->rep_avg_output_tokens, P * 8), o_hr = s.add(in->rep_prefix_hit_rate, P * 8),
               o_k2 = s.add(in->rep_k2, P * 8), o_so = s.add(in->rep_slice_order, P * 4), o_vc = s.add(in->var_current, V * 4),
               o_vp = s.add(in->var_pending, V * 4), o_vf = s.add(in->var_fallback_capacity, V * 8),
               o_kt = s.add(in->cfg_kv_threshold, M * 8), o_su = s.add(in->cfg_scale_up_threshold, M * 8),
               o_sd = s.add(in->cfg_scale_down_boundary, M * 8), o_qs = s.add(in->sched_queue_size, M * 8),
               o_qb = s.add(in->sched_queue_bytes, M * 8), o_co = s.add(var_cost, V * 8), o_nr = s.add(var_name_rank, V * 4),
               o_z = s.add(mod_s2z, M), o_rc = s.add(mod_request_count, M * 8), o_re = s.add(mod_request_error, M);
  int32_t rc = s.upload();
  if (rc != WVA_OK) return rc;
  Layout O;
  const size_t q_k1 = O.take(P * 8), q_ef = O.take(P * 8), q_de = O.take(P * 8), q_sa = O.take(P), q_vr = O.take(V * 4),
               q_vcap = O.take(V * 8), q_vt = O.take(V * 8), q_vd = O.take(V * 8), q_vu = O.take(V * 8), q_ms = O.take(M * 8),
               q_md = O.take(M * 8), q_mu = O.take(M * 8), q_mr = O.take(M * 8), q_mp = O.take(M * 8), q_tg = O.take(V * 4),
               q_ap = O.take(M);
  CK(ctx->io_out.reserve(O.off + 256));
  char* dd = (char*)ctx->io_out.p;
  const wva_saturation_v2_out none = {};
  const wva_saturation_v2_out& w = out ? *out : none;
  SatV2In di;
  di.n_models = in->n_models; di.n_variants = in->n_variants; di.n_replicas = in->n_replicas;
  di.model_variant_off = s.at<int>(o_mvo, in->model_variant_off); di.variant_replica_off = s.at<int>(o_vro, in->variant_replica_off);
  di.rep_total_kv = s.at<long long>(o_tk, in->rep_total_kv_tokens); di.rep_tokens_in_use = s.at<long long>(o_tu, in->rep_tokens_in_use);
  di.rep_queue_len = s.at<long long>(o_ql, in->rep_queue_length); di.rep_k2 = s.at<long long>(o_k2, in->rep_k2);
  di.rep_avg_in = s.at<double>(o_ai, in->rep_avg_input_tokens); di.rep_avg_out = s.at<double>(o_ao, in->rep_avg_output_tokens);
  di.rep_hit = s.at<double>(o_hr, in->rep_prefix_hit_rate); di.rep_slice_order = s.at<int>(o_so, in->rep_slice_order);
  di.var_current = s.at<int>(o_vc, in->var_current); di.var_pending = s.at<int>(o_vp, in->var_pending);
  di.var_fallback = s.at<double>(o_vf, in->var_fallback_capacity);
  di.cfg_kv_threshold = s.at<double>(o_kt, in->cfg_kv_threshold); di.cfg_scale_up = s.at<double>(o_su, in->cfg_scale_up_threshold);
  di.cfg_scale_down = s.at<double>(o_sd, in->cfg_scale_down_boundary);
  di.sched_size = s.at<long long>(o_qs, in->sched_queue_size); di.sched_bytes = s.at<long long>(o_qb, in->sched_queue_bytes);
  SatV2Out dout;
  dout.rep_k1 = w.rep_k1 ? (long long*)(dd + q_k1) : nullptr; dout.rep_effective = w.rep_effective ? (long long*)(dd + q_ef) : nullptr;
  dout.rep_demand = w.rep_demand ? (long long*)(dd + q_de) : nullptr; dout.rep_saturated = w.rep_saturated ? (unsigned char*)(dd + q_sa) : nullptr;
  dout.var_ready = w.var_ready ? (int*)(dd + q_vr) : nullptr;
  dout.var_cap = (double*)(dd + q_vcap);                                       // always: the optimizer reads it
  dout.var_total_cap = w.var_total_capacity ? (double*)(dd + q_vt) : nullptr; dout.var_total_demand = w.var_total_demand ? (double*)(dd + q_vd) : nullptr;
  dout.var_util = w.var_utilization ? (double*)(dd + q_vu) : nullptr;
  dout.mod_supply = w.mod_total_supply ? (double*)(dd + q_ms) : nullptr; dout.mod_demand = w.mod_total_demand ? (double*)(dd + q_md) : nullptr;
  dout.mod_util = w.mod_utilization ? (double*)(dd + q_mu) : nullptr;
  dout.mod_required = (double*)(dd + q_mr); dout.mod_spare = (double*)(dd + q_mp);   // always
  int* d_t = (int*)(dd + q_tg);
  unsigned char* d_app = (unsigned char*)(dd + q_ap);
  const unsigned grid = warp_grid(ctx, in->n_models);
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  saturation_v2_kernel<<<grid, 256, 0, ctx->stream>>>(di, dout);
  cost_aware_kernel<<<grid, 256, 0, ctx->stream>>>(in->n_models, di.model_variant_off, dout.mod_required, dout.mod_spare, nullptr,
                                                   di.var_current, s.at<double>(o_co, var_cost), dout.var_cap, d_t);
  enforce_kernel<<<grid, 256, 0, ctx->stream>>>(in->n_models, di.model_variant_off, s.at<unsigned char>(o_z, mod_s2z),
                                                s.at<double>(o_rc, mod_request_count), s.at<unsigned char>(o_re, mod_request_error),
                                                s.at<double>(o_co, var_cost), nullptr, s.at<int>(o_nr, var_name_rank), d_t, d_app);
  ctx->launches += 3;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  auto back = [&](void* host, size_t off, size_t bytes) -> cudaError_t {
    return (host && bytes) ? cudaMemcpyAsync(host, dd + off, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
  };
  CK(back(w.rep_k1, q_k1, P * 8)); CK(back(w.rep_effective, q_ef, P * 8)); CK(back(w.rep_demand, q_de, P * 8));
  CK(back(w.rep_saturated, q_sa, P)); CK(back(w.var_ready, q_vr, V * 4)); CK(back(w.var_per_replica_capacity, q_vcap, V * 8));
  CK(back(w.var_total_capacity, q_vt, V * 8)); CK(back(w.var_total_demand, q_vd, V * 8)); CK(back(w.var_utilization, q_vu, V * 8));
  CK(back(w.mod_total_supply, q_ms, M * 8)); CK(back(w.mod_total_demand, q_md, M * 8)); CK(back(w.mod_utilization, q_mu, M * 8));
  CK(back(w.mod_required_capacity, q_mr, M * 8)); CK(back(w.mod_spare_capacity, q_mp, M * 8));
  CK(back(var_target, q_tg, V * 4)); CK(back(mod_applied, q_ap, M));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.saturation_ms = elapsed(ctx, 2, 3);
  return WVA_OK;
}
