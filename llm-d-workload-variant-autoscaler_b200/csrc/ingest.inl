// ingest.inl — Collector -> SoA ingest and the streaming reconcile of BASELINE configs[4]:
// metric batch -> decisions as ONE CUDA graph launch (included at the end of capi.cu).
//
// The reference assembles []ReplicaMetrics per model with string-keyed joins: six Prometheus vectors are folded into a
// map[podName]*podMetricData (internal/collector/replica_metrics.go:118-290), every pod is matched to its
// VariantAutoscaling through PodVAMapper.FindVAForPod (source/pod_va_mapper.go:32, replica_metrics.go:321) and the
// records are appended in map order (:296-396).  Here the string work is done ONCE per pod: a pod is registered into a
// slot of its variant (the registry = CSR model -> variant -> slot, slots of a variant in ascending pod name = the
// canonical order of the per-variant float64 sums), and every cycle the response parser writes each sample straight into
// page-locked columns indexed by slot (one hash look-up pod name -> slot per sample; wva_ingest_write applies the
// reference's per-vector rules).  wva_ingest_commit then replays a captured CUDA graph:
//     H2D of the column arena (one copy) -> count the pods that reported, per variant -> exclusive scan = CSR offsets
//     -> pack (kv, queue) of those pods in slot order (pods with neither metric are skipped, a missing metric reads 0:
//     replica_metrics.go:296-318; queue = Go's int(float64)) -> model descriptors -> V1 saturation analysis + targets
//     (saturation_kernel.cuh, the same kernel wva_saturation_run launches) -> D2H of the decision arena (one copy).
// No per-cycle allocation, no host synchronisation inside, one launch.
namespace {

__global__ void __launch_bounds__(256) ingest_count_kernel(long long V, const int* __restrict__ vso,
                                                           const unsigned char* __restrict__ has, int* __restrict__ cnt) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int c = 0;
  for (int k = vso[v]; k < vso[v + 1]; k++) c += has[k] ? 1 : 0;
  cnt[v] = c;
}

// Go's int(float64) on amd64 (CVTTSD2SI): truncation; NaN and out-of-range values give the "integer indefinite" 1 << 63
__device__ __forceinline__ long long go_int_from_f64(double x) {
  if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return (long long)0x8000000000000000ull;
  return (long long)x;
}

__global__ void __launch_bounds__(256) ingest_pack_kernel(long long V, const int* __restrict__ vso, const int* __restrict__ vro,
                                                          const unsigned char* __restrict__ has, const double* __restrict__ kv,
                                                          const double* __restrict__ queue, double* __restrict__ rep_kv,
                                                          long long* __restrict__ rep_queue, int* __restrict__ rep_slot) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int o = vro[v];
  for (int k = vso[v]; k < vso[v + 1]; k++) {
    const unsigned char h = has[k];
    if (!h) continue;                                            // replica_metrics.go:298-300
    rep_kv[o] = (h & 1) ? kv[k] : 0.0;                           // :305-311
    rep_queue[o] = (h & 2) ? go_int_from_f64(queue[k]) : 0;      // :312-318, :160
    if (rep_slot) rep_slot[o] = k;
    o++;
  }
}

}  // namespace

struct wva_ingest {
  wva_ctx* ctx = nullptr;
  long long M = 0, V = 0, S = 0;      // models, variants, slots (registered pods)
  PinBuf h_in, h_out;                 // column arena (host, page-locked) and decision arena
  DevBuf d_in, d_reg, d_work, d_out;
  size_t in_bytes = 0, out_bytes = 0;
  wva_ingest_columns cols = {};
  wva_ingest_results res = {};
  SatIn vin = {};
  SatOut vout = {};
  // device views
  const int* d_vso = nullptr; int* d_cnt = nullptr; int* d_vro = nullptr; int* d_rep_slot = nullptr;
  const unsigned char* d_has = nullptr; const double* d_kv = nullptr; const double* d_queue = nullptr;
  double* d_rep_kv = nullptr; long long* d_rep_queue = nullptr;
  SatDesc* d_desc = nullptr;
  void* d_scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long cycles = 0;
  IngestScratch scratch;              // host scratch of wva_ingest_write
};

static cudaError_t ingest_enqueue(wva_ingest* g, cudaStream_t s) {
  const long long V = g->V, M = g->M;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(g->d_in.p, g->h_in.p, g->in_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
  if (V > 0) {
    const unsigned vb = (unsigned)((V + 255) / 256);
    ingest_count_kernel<<<vb, 256, 0, s>>>(V, g->d_vso, g->d_has, g->d_cnt);
    size_t tb = g->scan_tmp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(g->d_scan_tmp, tb, g->d_cnt, g->d_vro, (int)(V + 1), s)) != cudaSuccess) return e;
    ingest_pack_kernel<<<vb, 256, 0, s>>>(V, g->d_vso, g->d_vro, g->d_has, g->d_kv, g->d_queue, g->d_rep_kv, g->d_rep_queue,
                                          g->d_rep_slot);
  }
  if ((e = cudaMemsetAsync(g->vout.partials, 0, 64, s)) != cudaSuccess) return e;
  if (M > 0) {
    if ((e = launch_saturation(true, g->ctx->sm_count, M, g->vin, g->vout, g->d_desc, s)) != cudaSuccess) return e;
  }
  if ((e = cudaMemcpyAsync(g->h_out.p, g->d_out.p, g->out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  return cudaGetLastError();
}

extern "C" int32_t wva_ingest_destroy(wva_ingest* g) {
  if (!g) return WVA_ERR_ARG;
  if (g->ctx) cudaSetDevice(g->ctx->device);
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  g->h_in.release(); g->h_out.release(); g->d_in.release(); g->d_reg.release(); g->d_work.release(); g->d_out.release();
  delete g;
  return WVA_OK;
}

extern "C" int32_t wva_ingest_create(wva_ctx* ctx, int64_t n_models, int64_t n_variants, int64_t n_slots,
                                     const int32_t* model_variant_off, const int32_t* variant_slot_off, wva_ingest** out,
                                     wva_ingest_columns* cols, wva_ingest_results* res) {
  if (!ctx || !out || !cols || !res || n_models < 0 || n_variants < 0 || n_slots < 0) return WVA_ERR_ARG;
  *out = nullptr;
  if (n_variants > 0x7ffffff0LL || n_slots > 0x7ffffff0LL) return WVA_ERR_LIMIT;
  if (!valid_offsets(model_variant_off, (size_t)n_models, (size_t)n_variants) ||
      !valid_offsets(variant_slot_off, (size_t)n_variants, (size_t)n_slots)) {
    ctx->last_error = "wva_ingest_create: model_variant_off / variant_slot_off are not CSR offsets of the given sizes";
    return WVA_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  wva_ingest* g = new (std::nothrow) wva_ingest();
  if (!g) return WVA_ERR_NOMEM;
  g->ctx = ctx; g->M = n_models; g->V = n_variants; g->S = n_slots;
  const size_t M = (size_t)n_models, V = (size_t)n_variants, S = (size_t)n_slots;
  auto fail = [&](int32_t rc) { wva_ingest_destroy(g); return rc; };
  // ---- column arena (host page-locked image == device image): per-slot columns, per-variant state, per-model config
  Layout L;
  const size_t o_kv = L.take(S * 8), o_q = L.take(S * 8), o_has = L.take(S), o_cost = L.take(V * 8), o_cur = L.take(V * 4),
               o_des = L.take(V * 4), o_pen = L.take(V * 4), o_c0 = L.take(M * 8), o_c1 = L.take(M * 8), o_c2 = L.take(M * 8),
               o_c3 = L.take(M * 8);
  g->in_bytes = L.off + 256;
  if (g->h_in.reserve(g->in_bytes) != cudaSuccess || g->d_in.reserve(g->in_bytes) != cudaSuccess) return fail(WVA_ERR_NOMEM);
  memset(g->h_in.p, 0, g->in_bytes);
  char* h = (char*)g->h_in.p; char* d = (char*)g->d_in.p;
  g->cols.n_slots = n_slots; g->cols.n_variants = n_variants; g->cols.n_models = n_models;
  g->cols.kv = (double*)(h + o_kv); g->cols.queue = (double*)(h + o_q); g->cols.has = (uint8_t*)(h + o_has);
  g->cols.var_cost = (double*)(h + o_cost); g->cols.var_current = (int32_t*)(h + o_cur); g->cols.var_desired = (int32_t*)(h + o_des);
  g->cols.var_pending = (int32_t*)(h + o_pen); g->cols.cfg_kv_threshold = (double*)(h + o_c0);
  g->cols.cfg_queue_threshold = (double*)(h + o_c1); g->cols.cfg_kv_trigger = (double*)(h + o_c2);
  g->cols.cfg_queue_trigger = (double*)(h + o_c3);
  g->d_kv = (const double*)(d + o_kv); g->d_queue = (const double*)(d + o_q); g->d_has = (const unsigned char*)(d + o_has);
  // ---- registry + work arrays
  Layout R;
  const size_t r_mvo = R.take((M + 1) * 4), r_vso = R.take((V + 1) * 4);
  if (g->d_reg.reserve(R.off + 256) != cudaSuccess) return fail(WVA_ERR_NOMEM);
  char* dr = (char*)g->d_reg.p;
  if (cudaMemcpyAsync(dr + r_mvo, model_variant_off, (M + 1) * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(dr + r_vso, variant_slot_off, (V + 1) * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return fail(WVA_ERR_CUDA);
  g->d_vso = (const int*)(dr + r_vso);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (int*)nullptr, (int*)nullptr, (int)(V + 1), ctx->stream);
  g->scan_tmp_bytes = tb;
  Layout W;
  const size_t w_cnt = W.take((V + 2) * 4), w_vro = W.take((V + 2) * 4), w_rk = W.take(S * 8 + 16), w_rq = W.take(S * 8 + 16),
               w_rs = W.take(S * 4), w_desc = W.take(sat_desc_bytes(M)), w_tmp = W.take(tb + 256);
  if (g->d_work.reserve(W.off + 256) != cudaSuccess) return fail(WVA_ERR_NOMEM);
  char* dw = (char*)g->d_work.p;
  if (cudaMemsetAsync(dw, 0, g->d_work.cap, ctx->stream) != cudaSuccess) return fail(WVA_ERR_CUDA);   // cnt[V] = 0: the scan's extra element
  g->d_cnt = (int*)(dw + w_cnt); g->d_vro = (int*)(dw + w_vro); g->d_rep_kv = (double*)(dw + w_rk);
  g->d_rep_queue = (long long*)(dw + w_rq); g->d_rep_slot = (int*)(dw + w_rs); g->d_desc = (SatDesc*)(dw + w_desc);
  g->d_scan_tmp = dw + w_tmp;
  // ---- decision arena
  Layout O;
  const size_t q_part = O.take(64), q_t = O.take(V * 4), q_rc = O.take(V * 4), q_ns = O.take(V * 4), q_ak = O.take(V * 8),
               q_aq = O.take(V * 8), q_mf = O.take(M), q_mt = O.take(M * 4);
  g->out_bytes = O.off + 256;
  if (g->h_out.reserve(g->out_bytes) != cudaSuccess || g->d_out.reserve(g->out_bytes) != cudaSuccess) return fail(WVA_ERR_NOMEM);
  memset(g->h_out.p, 0, g->out_bytes);
  char* ho = (char*)g->h_out.p; char* dob = (char*)g->d_out.p;
  g->res.partials = (int64_t*)(ho + q_part); g->res.var_target = (int32_t*)(ho + q_t); g->res.var_replica_count = (int32_t*)(ho + q_rc);
  g->res.var_non_saturated = (int32_t*)(ho + q_ns); g->res.var_avg_spare_kv = (double*)(ho + q_ak);
  g->res.var_avg_spare_queue = (double*)(ho + q_aq); g->res.mod_flags = (uint8_t*)(ho + q_mf);
  g->res.mod_total_replicas = (int32_t*)(ho + q_mt);
  SatIn& v = g->vin;
  v.n_models = n_models; v.n_variants = n_variants; v.n_replicas = n_slots;
  v.model_variant_off = (const int*)(dr + r_mvo); v.variant_replica_off = g->d_vro;
  v.rep_kv = g->d_rep_kv; v.rep_queue = g->d_rep_queue;
  v.var_cost = (const double*)(d + o_cost); v.var_current = (const int*)(d + o_cur); v.var_desired = (const int*)(d + o_des);
  v.var_pending = (const int*)(d + o_pen); v.var_has_state = nullptr;
  v.cfg_kv_threshold = (const double*)(d + o_c0); v.cfg_queue_threshold = (const double*)(d + o_c1);
  v.cfg_kv_trigger = (const double*)(d + o_c2); v.cfg_queue_trigger = (const double*)(d + o_c3);
  SatOut& w = g->vout;
  w = SatOut{};
  w.partials = (long long*)(dob + q_part); w.var_target = (int*)(dob + q_t); w.var_replica_count = (int*)(dob + q_rc);
  w.var_non_saturated = (int*)(dob + q_ns); w.var_avg_spare_kv = (double*)(dob + q_ak); w.var_avg_spare_queue = (double*)(dob + q_aq);
  w.mod_flags = (unsigned char*)(dob + q_mf); w.mod_total_replicas = (int*)(dob + q_mt);
  // (the kernel's shared-memory attribute is set before the capture as well: launch_saturation sets it again, harmlessly)
  if (sat_prepare_attributes() != cudaSuccess) return fail(WVA_ERR_CUDA);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(WVA_ERR_CUDA);
  // ---- capture the cycle once
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail(WVA_ERR_CUDA);
  const cudaError_t ce = ingest_enqueue(g, ctx->stream);
  const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g->graph);
  if (ce != cudaSuccess || ee != cudaSuccess || !g->graph) {
    ctx->last_error = std::string("wva_ingest_create: graph capture failed: ") + cudaGetErrorString(ce != cudaSuccess ? ce : ee);
    return fail(WVA_ERR_CUDA);
  }
  if (cudaGraphInstantiate(&g->exec, g->graph, 0) != cudaSuccess) return fail(WVA_ERR_CUDA);
  *cols = g->cols; *res = g->res; *out = g;
  return WVA_OK;
}

extern "C" int32_t wva_ingest_begin(wva_ingest* g) {
  if (!g) return WVA_ERR_ARG;
  if (g->S) memset(g->cols.has, 0, (size_t)g->S);
  return WVA_OK;
}

// one Prometheus-shaped vector keyed by slot, in result order (later samples of a pod overwrite earlier ones: the
// reference assigns into a map, replica_metrics.go:133-160); slot < 0 = sample without a pod label or of an unknown pod
extern "C" int32_t wva_ingest_write(wva_ingest* g, int32_t which, int64_t n, const int32_t* slot, const double* value) {
  if (!g || n < 0 || (n > 0 && (!slot || !value)) || (which != WVA_VEC_KV_CACHE_USAGE && which != WVA_VEC_QUEUE_LENGTH)) return WVA_ERR_ARG;
  double* col = which == WVA_VEC_KV_CACHE_USAGE ? g->cols.kv : g->cols.queue;
  const uint8_t bit = which == WVA_VEC_KV_CACHE_USAGE ? 1 : 2;
  // large vectors: two-pass radix partition over host threads (ingest_scatter.hpp), same result as the serial loop
  return ingest_scatter(col, g->cols.has, g->S, bit, n, slot, value, g->scratch) ? WVA_OK : WVA_ERR_ARG;
}

extern "C" int32_t wva_ingest_commit(wva_ingest* g) {
  if (!g) return WVA_ERR_ARG;
  wva_ctx* ctx = g->ctx;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  CK(cudaGraphLaunch(g->exec, ctx->stream));
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.saturation_ms = elapsed(ctx, 2, 3);       // the whole graph: H2D + pack + analysis + D2H
  ctx->launches += 6;
  g->cycles++;
  return WVA_OK;
}
