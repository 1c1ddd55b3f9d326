// gmql_ctx.cu — context, dataset upload, result accessors of the C ABI (include/gmql_cuda.h).
#include "index.cuh"
#include <algorithm>

extern "C" int gmql_abi_version(void) { return GMQL_ABI_VERSION; }

extern "C" int gmql_ctx_create(int device, void* stream, gmql_ctx** out) {
  if (!out) return GMQL_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) { cudaGetLastError(); return GMQL_ERR_NODEVICE; }
  if (device < 0 || device >= count) return GMQL_ERR_INVALID;
  gmql_ctx* c = new gmql_ctx();
  c->device = device;
  try {
    CUDA_TRY(cudaSetDevice(device));
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaEventCreate(&c->ev_k0));
    CUDA_TRY(cudaEventCreate(&c->ev_k1));
    CUDA_TRY(cudaEventCreate(&c->ev_h0));
    CUDA_TRY(cudaEventCreate(&c->ev_h1));
    // a private stream-ordered pool (the process's default pool is left alone: other allocators of the host process are not
    // starved by what this library keeps cached); released blocks stay in it until gmql_ctx_destroy
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CUDA_TRY(cudaMemPoolCreate(&c->pool, &props));
    uint64_t thr = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    c->arena_limit = total_b / 4;
  } catch (const gmql_error& ex) {
    delete c;
    return ex.code;
  }
  *out = c;
  return GMQL_OK;
}

extern "C" void gmql_ctx_destroy(gmql_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto* d : c->datasets) if (d) d->owner = nullptr;   // datasets / results must be freed before their context (header)
  arena_trim(c);
  cudaStreamSynchronize(c->stream);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  if (c->ev_k0) cudaEventDestroy(c->ev_k0);
  if (c->ev_k1) cudaEventDestroy(c->ev_k1);
  if (c->ev_h0) cudaEventDestroy(c->ev_h0);
  if (c->ev_h1) cudaEventDestroy(c->ev_h1);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  for (auto& pin : c->pins) if (pin.first) cudaFreeHost(pin.first);
  if (c->h_defer) cudaFreeHost(c->h_defer);
  delete c;
}

extern "C" const char* gmql_last_error(const gmql_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int gmql_ctx_set_option(gmql_ctx* c, const char* name, const char* value) {
  if (!c || !name) return GMQL_ERR_INVALID;
  static const char* known[] = {"cover.path", "cover.logw", "cover.waves", "cover.stitch", "cover.speculate", "cover.kernel", "cover.dbg", "cover.blocks", "map.index", "map.grouped", "map.lean",
                                "defer_check", "join.md", "join.md_memo", "cover.gmap4", "cover.bins", "cover.finish", nullptr};
  bool ok = false;
  for (int k = 0; known[k]; ++k) if (!strcmp(known[k], name)) ok = true;
  if (!ok) { c->err = std::string("unknown option: ") + name; return GMQL_ERR_INVALID; }
  if (value && *value) c->options[name] = value; else c->options.erase(name);
  return GMQL_OK;
}

// flags of operators that returned without their last synchronisation (option "defer_check"); call with the stream idle
int gmql_check_deferred(gmql_ctx* c) {
  int rc = GMQL_OK;
  for (auto& d : c->deferred)
    if (*d.h_flag) { c->err = std::string("deferred check of an earlier call failed: ") + d.what; rc = GMQL_ERR_INVALID; }
  c->deferred.clear();
  return rc;
}

extern "C" int gmql_ctx_sync(gmql_ctx* c) {
  if (!c) return GMQL_ERR_INVALID;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  c->pending = false;
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return GMQL_ERR_CUDA; }
  return gmql_check_deferred(c);
}

extern "C" int64_t gmql_ctx_launch_count(const gmql_ctx* c) { return c ? c->launches : 0; }

extern "C" int gmql_host_alloc(size_t bytes, void** out) {
  if (!out) return GMQL_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return GMQL_ERR_NOMEM; }
  return GMQL_OK;
}
extern "C" void gmql_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" double gmql_last_device_ms(const gmql_ctx* c) {
  if (!c) return -1.0;
  float ms = 0.f;
  if (cudaEventSynchronize(c->ev_k1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
  return (double)ms;
}
extern "C" double gmql_last_h2d_ms(const gmql_ctx* c) {
  if (!c) return -1.0;
  float ms = 0.f;
  if (cudaEventSynchronize(c->ev_h1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, c->ev_h0, c->ev_h1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
  return (double)ms;
}

// ------------------------------------------------------------------------------------------------
static const void* upload_array(gmql_ctx* ctx, const void* src, size_t bytes, int location, DevRegions* out) {
  if (!src) return nullptr;
  if (location == GMQL_LOC_DEVICE) return src;
  DevBuf<uint8_t> b(ctx, (int64_t)(bytes ? bytes : 1));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  out->h2d_bytes += (int64_t)bytes;
  const void* p = b.p;
  out->owned.push_back(std::move(b));
  return p;
}

template <typename T>
static __global__ void k_widen_chrom(const T* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

// stop = start + length (lengths sent as uint16 / uint32)
template <typename C, typename LEN>
static __global__ void k_stop_from_length(const C* __restrict__ start, const LEN* __restrict__ len, int64_t n, C* __restrict__ stop) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) stop[i] = start[i] + (C)len[i];
}

// sample_offsets (one run of rows per sample) -> per-row sample index; one block per sample, grid-stride
static __global__ void k_expand_sample_offsets(const int64_t* __restrict__ off, int n_samples, int32_t* __restrict__ sample) {
  for (int s = blockIdx.x; s < n_samples; s += gridDim.x) {
    const int64_t a = off[s], b = off[s + 1];
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) sample[i] = s;
  }
}

// the dataset whose resident arrays a device-located gmql_regions points at (views made by gmql_dataset_upload), or null
static gmql_dataset* find_dataset(gmql_ctx* ctx, const gmql_regions* r) {
  if (r->location != GMQL_LOC_DEVICE || !r->start) return nullptr;
  for (auto* d : ctx->datasets)
    if (d && d->dev.start == r->start && d->dev.n == r->n && d->dev.coord_bits == r->coord_bits) return d;
  return nullptr;
}

void gmql_upload_regions(gmql_ctx* ctx, const gmql_regions* r, const std::vector<int>& need_cols, DevRegions* out, bool keep_narrow) {
  GMQL_REQUIRE(r != nullptr, GMQL_ERR_INVALID, "null gmql_regions");
  GMQL_REQUIRE(r->n >= 0 && r->n < (int64_t)0x7fffffff, GMQL_ERR_UNSUPPORTED, "datasets of >= 2^31 regions are not supported in this version");
  GMQL_REQUIRE(r->coord_bits == 32 || r->coord_bits == 64, GMQL_ERR_INVALID, "coord_bits must be 32 or 64");
  GMQL_REQUIRE(r->n == 0 || (r->chrom && r->start && r->stop), GMQL_ERR_INVALID, "chrom/start/stop are required");
  GMQL_REQUIRE(r->n_chrom > 0 && r->n_samples > 0, GMQL_ERR_INVALID, "n_chrom and n_samples must be positive");
  out->n = r->n;
  out->coord_bits = r->coord_bits;
  out->n_chrom = r->n_chrom;
  out->n_samples = r->n_samples;
  out->n_cols = r->n_cols;
  out->span_hint = r->chrom_span_hint;
  out->sorted_hint = r->sorted_by_position != 0;
  gmql_dataset* ds = find_dataset(ctx, r);
  out->dataset = ds;
  size_t cb = (size_t)(r->coord_bits / 8);
  size_t n = (size_t)r->n;
  GMQL_REQUIRE(r->chrom_bits == 0 || r->chrom_bits == 8 || r->chrom_bits == 16 || r->chrom_bits == 32, GMQL_ERR_INVALID, "chrom_bits must be 0, 8, 16 or 32");
  out->wide_ready = true;
  if (r->chrom_bits == 8 || r->chrom_bits == 16) {
    GMQL_REQUIRE(r->n_chrom <= (r->chrom_bits == 8 ? 256 : 65536), GMQL_ERR_INVALID, "n_chrom does not fit chrom_bits");
    out->chrom_narrow = upload_array(ctx, r->chrom, n * (size_t)(r->chrom_bits / 8), r->location, out);
    out->chrom_bits = r->chrom_bits;
    out->chrom = nullptr;
    out->wide_ready = false;
  } else {
    out->chrom = (const int32_t*)upload_array(ctx, r->chrom, n * 4, r->location, out);
  }
  out->start = upload_array(ctx, r->start, n * cb, r->location, out);
  GMQL_REQUIRE(r->stop_is_length == 0 || r->stop_is_length == 16 || r->stop_is_length == 32, GMQL_ERR_INVALID, "stop_is_length must be 0, 16 or 32");
  if (r->stop_is_length) {
    out->len_narrow = upload_array(ctx, r->stop, n * (size_t)(r->stop_is_length / 8), r->location, out);
    out->len_bits = r->stop_is_length;
    out->stop = nullptr;
    out->wide_ready = false;
  } else {
    out->stop = upload_array(ctx, r->stop, n * cb, r->location, out);
  }
  out->strand = (const int8_t*)upload_array(ctx, r->strand, n, r->location, out);
  out->host_sample_col = r->sample != nullptr;
  if (r->sample_offsets) {
    // rows grouped by sample.  Without a sample column the offsets define it; with one they are a promise about its
    // layout, which the kernels that rely on it check against the column.
    const int64_t* off = r->sample_offsets;   // host array by contract
    if (ds && !ds->sample_offsets.empty() && off == ds->sample_offsets.data()) {
      out->sample_off = ds->dev.sample_off;    // validated and uploaded once, with the dataset
      out->max_sample_rows = ds->dev.max_sample_rows;
    } else {
      GMQL_REQUIRE(off[0] == 0 && off[r->n_samples] == r->n, GMQL_ERR_INVALID, "sample_offsets must start at 0 and end at n");
      int64_t widest = 0;
      for (int s = 0; s < r->n_samples; ++s) {
        GMQL_REQUIRE(off[s] <= off[s + 1], GMQL_ERR_INVALID, "sample_offsets must be non-decreasing");
        widest = std::max<int64_t>(widest, off[s + 1] - off[s]);
      }
      const size_t ob = 8 * ((size_t)r->n_samples + 1);
      DevBuf<uint8_t> doff(ctx, (int64_t)ob);
      int64_t* h_off = (int64_t*)ctx_pinned(ctx, ob);
      memcpy(h_off, off, ob);
      CUDA_TRY(cudaMemcpyAsync(doff.p, h_off, ob, cudaMemcpyHostToDevice, ctx->stream));
      if (r->location != GMQL_LOC_DEVICE) out->h2d_bytes += (int64_t)ob;
      out->sample_off = (const int64_t*)doff.p;
      out->max_sample_rows = widest;
      out->owned.push_back(std::move(doff));
    }
    for (const auto& tg : ctx->trusted_groupings) if (tg.first == (const void*)r->sample && tg.second == (const void*)r->sample_offsets && r->sample) out->sample_trusted = true;
    if (!r->sample) {
      out->sample_trusted = true;
      out->sample = nullptr;
      out->wide_ready = false;
    } else {
      out->sample = (const int32_t*)upload_array(ctx, r->sample, n * 4, r->location, out);
    }
  } else {
    out->sample = (const int32_t*)upload_array(ctx, r->sample, n * 4, r->location, out);
  }
  out->dup_of = (const int32_t*)upload_array(ctx, r->dup_of, n * 4, r->location, out);
  out->cols.assign((size_t)std::max(0, r->n_cols), nullptr);
  out->valid.assign((size_t)std::max(0, r->n_cols), nullptr);
  for (int c : need_cols) {
    GMQL_REQUIRE(c >= 0 && c < r->n_cols && r->cols && r->cols[c], GMQL_ERR_INVALID, "aggregate refers to a value column that is not provided");
    if (out->cols[c]) continue;
    out->cols[c] = (const double*)upload_array(ctx, r->cols[c], n * 8, r->location, out);
    if (r->valid && r->valid[c]) out->valid[c] = (const uint8_t*)upload_array(ctx, r->valid[c], n, r->location, out);
  }
  if (!keep_narrow) gmql_ensure_wide(ctx, out);
}

// int32 chromosome indices, stop coordinates and the per-row sample column of a dataset that arrived in the narrow encodings.
// For a resident dataset they are materialised once and kept with it; otherwise they live as long as `d`.
void gmql_ensure_wide(gmql_ctx* ctx, DevRegions* d) {
  if (d->wide_ready) return;
  const int64_t n = d->n;
  gmql_dataset* ds = d->dataset;
  if (ds && ds->wide_done) {
    if (!d->chrom) d->chrom = (const int32_t*)ds->wide_chrom.p;
    if (!d->stop) d->stop = ds->wide_stop.p;
    if (!d->sample && d->sample_off) d->sample = (const int32_t*)ds->wide_sample.p;
    d->wide_ready = true;
    return;
  }
  const size_t cb = (size_t)(d->coord_bits / 8);
  const int g = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256 * 4), (int64_t)ctx->num_sms * 16));
  DevBuf<uint8_t> wc, ws, wm;
  if (!d->chrom && d->chrom_narrow) {
    wc.alloc(ctx, (int64_t)(n * 4 ? n * 4 : 1));
    if (n) {
      if (d->chrom_bits == 8) LAUNCH(ctx, (k_widen_chrom<uint8_t>), g, 256, 0, (const uint8_t*)d->chrom_narrow, n, (int32_t*)wc.p);
      else LAUNCH(ctx, (k_widen_chrom<uint16_t>), g, 256, 0, (const uint16_t*)d->chrom_narrow, n, (int32_t*)wc.p);
    }
    d->chrom = (const int32_t*)wc.p;
  }
  if (!d->stop && d->len_narrow) {
    ws.alloc(ctx, (int64_t)((size_t)n * cb ? (size_t)n * cb : 1));
    if (n) {
      if (cb == 4 && d->len_bits == 16) LAUNCH(ctx, (k_stop_from_length<int32_t, uint16_t>), g, 256, 0, (const int32_t*)d->start, (const uint16_t*)d->len_narrow, n, (int32_t*)ws.p);
      else if (cb == 4) LAUNCH(ctx, (k_stop_from_length<int32_t, uint32_t>), g, 256, 0, (const int32_t*)d->start, (const uint32_t*)d->len_narrow, n, (int32_t*)ws.p);
      else if (d->len_bits == 16) LAUNCH(ctx, (k_stop_from_length<int64_t, uint16_t>), g, 256, 0, (const int64_t*)d->start, (const uint16_t*)d->len_narrow, n, (int64_t*)ws.p);
      else LAUNCH(ctx, (k_stop_from_length<int64_t, uint32_t>), g, 256, 0, (const int64_t*)d->start, (const uint32_t*)d->len_narrow, n, (int64_t*)ws.p);
    }
    d->stop = ws.p;
  }
  if (!d->sample && d->sample_off) {
    wm.alloc(ctx, (int64_t)(n * 4 ? n * 4 : 1));
    if (n) LAUNCH(ctx, k_expand_sample_offsets, (int)std::min<int64_t>(d->n_samples, (int64_t)ctx->num_sms * 16), 256, 0, d->sample_off, d->n_samples, (int32_t*)wm.p);
    d->sample = (const int32_t*)wm.p;
  }
  if (ds) {
    ds->wide_chrom = std::move(wc); ds->wide_stop = std::move(ws); ds->wide_sample = std::move(wm);
    ds->wide_done = true;
  } else {
    if (wc.p) d->owned.push_back(std::move(wc));
    if (ws.p) d->owned.push_back(std::move(ws));
    if (wm.p) d->owned.push_back(std::move(wm));
  }
  d->wide_ready = true;
}

// ------------------------------------------------------------------------------------------------
extern "C" int64_t gmql_result_rows(const gmql_result* r) { return r ? r->rows : -1; }

extern "C" int64_t gmql_result_column_len(const gmql_result* r, int32_t col) {
  if (!r) return -1;
  auto it = r->cols.find(col);
  return it == r->cols.end() ? -1 : it->second.len;
}
extern "C" int32_t gmql_result_column_elem_size(const gmql_result* r, int32_t col) {
  if (!r) return -1;
  auto it = r->cols.find(col);
  return it == r->cols.end() ? -1 : it->second.elem;
}
extern "C" const void* gmql_result_column_device(const gmql_result* r, int32_t col) {
  if (!r) return nullptr;
  auto it = r->cols.find(col);
  return it == r->cols.end() ? nullptr : it->second.p;
}
extern "C" int gmql_result_column_copy(gmql_ctx* ctx, const gmql_result* r, int32_t col, void* dst) {
  if (!ctx || !r) return GMQL_ERR_INVALID;
  auto it = r->cols.find(col);
  if (it == r->cols.end()) { ctx->err = "result has no such column"; return GMQL_ERR_INVALID; }
  size_t bytes = (size_t)it->second.len * (size_t)it->second.elem;
  if (bytes) {
    if (!dst) { ctx->err = "null destination"; return GMQL_ERR_INVALID; }
    cudaError_t e = cudaMemcpyAsync(dst, it->second.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->pending = false;
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return GMQL_ERR_CUDA; }
    if (!ctx->deferred.empty()) return gmql_check_deferred(ctx);
  }
  return GMQL_OK;
}

extern "C" void gmql_result_free(gmql_ctx* ctx, gmql_result* r) {
  if (!r) return;
  gmql_trace("result_free");
  gmql_ctx* owner = r->owner ? r->owner : ctx;
  for (auto& kv : r->cols)
    if (kv.second.p) { if (owner) ctx_free(owner, kv.second.p); else cudaFree(kv.second.p); }
  if (r->view_acc_double) { if (owner) ctx_free(owner, r->view_acc_double); else cudaFree(r->view_acc_double); }
  delete r;
}

// ------------------------------------------------------------------------------------------------
// Upload-time statistics of a dataset, in ONE read of its arrays in whatever encoding they arrived: largest stop per chromosome
// (the linear layout every operator derives), longest region, "a zero-length region exists", "a '*' strand exists", and row
// validity (chromosome / sample index in range, 0 <= start <= stop, strand code, no 32-bit overflow of start + length).
static __global__ void __launch_bounds__(256) k_dataset_stats(const void* __restrict__ chrom, int chrom_bits, const void* __restrict__ start, int bits, const void* __restrict__ stop_or_len,
                                                              int len_bits, const int8_t* __restrict__ strand, const int32_t* __restrict__ sample, int64_t n, int n_chrom, int n_samples,
                                                              unsigned long long* __restrict__ span, int* __restrict__ flags /* [0] '*' seen, [1] zero-length, [2] bad row, [3] longest (clamped) */) {
  extern __shared__ unsigned long long sspan[];
  const bool use_sh = n_chrom <= 4096;
  if (use_sh) {
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) sspan[c] = 0ull;
    __syncthreads();
  }
  int has_star = 0, zero = 0, bad = 0;
  long long maxlen = 0;
  int64_t i_begin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the narrow encodings without strand / sample columns (the production upload): four rows per thread from three vector loads,
  // and the per-chromosome maximum is only touched when a row can raise it (a plain read first: the maximum only grows, so a
  // stale value filters safely) — one shared-memory atomic per row on two dozen addresses kept this kernel at 0.9 TB/s
  if (use_sh && chrom_bits == 8 && bits == 32 && len_bits == 16 && !strand && !sample &&
      ((((uintptr_t)chrom) & 3u) | (((uintptr_t)start) & 15u) | (((uintptr_t)stop_or_len) & 7u)) == 0) {
    const int64_t n4 = n >> 2;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
      const int4 vs = ((const int4*)start)[g];
      const uint2 vl = ((const uint2*)stop_or_len)[g];
      const uint32_t vc = ((const uint32_t*)chrom)[g];
      const int ss[4] = {vs.x, vs.y, vs.z, vs.w};
      const uint32_t ll[4] = {vl.x & 0xffffu, vl.x >> 16, vl.y & 0xffffu, vl.y >> 16};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = (int)((vc >> (8 * q)) & 0xffu);
        const int64_t s0 = ss[q], e0 = s0 + (int64_t)ll[q];
        if (c >= n_chrom || s0 < 0 || e0 > 0x7fffffffLL) { bad = 1; continue; }
        if (ll[q] == 0u) zero = 1;
        if ((long long)ll[q] > maxlen) maxlen = ll[q];
        if ((unsigned long long)e0 > sspan[c]) atomicMax(&sspan[c], (unsigned long long)e0);
      }
    }
    has_star = 1;                                          // no strand column: every region is '*'
    i_begin += n4 << 2;                                    // the last n % 4 rows go through the loop below
  }
  for (int64_t i = i_begin; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = chrom_bits == 8 ? (int)((const uint8_t*)chrom)[i] : (chrom_bits == 16 ? (int)((const uint16_t*)chrom)[i] : ((const int32_t*)chrom)[i]);
    const int64_t s = ld_coord(start, bits, i);
    int64_t e;
    if (len_bits == 16) e = s + (int64_t)((const uint16_t*)stop_or_len)[i];
    else if (len_bits == 32) e = s + (int64_t)((const uint32_t*)stop_or_len)[i];
    else e = ld_coord(stop_or_len, bits, i);
    const int st = strand ? (int)strand[i] : 0;
    const int sm = sample ? sample[i] : 0;
    if (st == 0) has_star = 1;
    if (c < 0 || c >= n_chrom || s < 0 || s > e || st < 0 || st > 2 || sm < 0 || sm >= n_samples || (bits == 32 && e > 0x7fffffffLL)) { bad = 1; continue; }
    if (s == e) zero = 1;
    if (e - s > maxlen) maxlen = e - s;
    if (use_sh) { if ((unsigned long long)e > sspan[c]) atomicMax(&sspan[c], (unsigned long long)e); } else atomicMax(&span[c], (unsigned long long)e);
  }
  __shared__ int s_fl[4];
  if (threadIdx.x < 4) s_fl[threadIdx.x] = 0;
  __syncthreads();
  {
    int ml = maxlen > 0x7fffffffLL ? 0x7fffffff : (int)maxlen;
    ml = __reduce_max_sync(0xffffffffu, ml);
    const bool a0 = __any_sync(0xffffffffu, has_star), a1 = __any_sync(0xffffffffu, zero), a2 = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      if (a0) s_fl[0] = 1;
      if (a1) s_fl[1] = 1;
      if (a2) s_fl[2] = 1;
      if (ml > 0) atomicMax(&s_fl[3], ml);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                  // one store / RED per block and flag
    if (s_fl[0]) flags[0] = 1;
    if (s_fl[1]) flags[1] = 1;
    if (s_fl[2]) flags[2] = 1;
    if (s_fl[3] > 0) atomicMax(&flags[3], s_fl[3]);
  }
  if (use_sh)
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) if (sspan[c]) atomicMax(&span[c], sspan[c]);
}

extern "C" int gmql_dataset_upload(gmql_ctx* ctx, const gmql_regions* host, gmql_dataset** out, gmql_regions* view) {
  if (!ctx || !host || !out || !view) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_dataset* d = new gmql_dataset();
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    std::vector<int> all;
    for (int c = 0; c < host->n_cols; ++c) all.push_back(c);
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    gmql_upload_regions(ctx, host, all, &d->dev, /*keep_narrow=*/true);   // the narrow encodings stay resident as they arrived
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    d->owner = ctx;
    if (host->sample_ids) d->sample_ids.assign(host->sample_ids, host->sample_ids + host->n_samples);
    else d->sample_ids.assign((size_t)host->n_samples, 0);
    d->cols = d->dev.cols;
    d->valid = d->dev.valid;
    const DevRegions& dv = d->dev;
    memset(view, 0, sizeof *view);
    view->n = dv.n;
    view->location = GMQL_LOC_DEVICE;
    view->coord_bits = dv.coord_bits;
    view->chrom = dv.chrom_narrow ? (const int32_t*)dv.chrom_narrow : dv.chrom;
    view->chrom_bits = dv.chrom_narrow ? dv.chrom_bits : 0;
    view->start = dv.start;
    view->stop = dv.len_narrow ? dv.len_narrow : dv.stop;
    view->stop_is_length = dv.len_narrow ? dv.len_bits : 0;
    view->strand = dv.strand;
    view->sample = dv.sample;             // null when the upload had sample_offsets only
    if (host->sample_offsets) {
      d->sample_offsets.assign(host->sample_offsets, host->sample_offsets + host->n_samples + 1);
      view->sample_offsets = d->sample_offsets.data();   // the view keeps the promise: rows grouped by sample
      if (dv.sample && dv.sample_trusted) ctx->trusted_groupings.push_back(std::make_pair((const void*)dv.sample, (const void*)d->sample_offsets.data()));
    }
    view->n_chrom = dv.n_chrom;
    view->n_samples = dv.n_samples;
    view->sample_ids = d->sample_ids.data();
    view->n_cols = dv.n_cols;
    view->cols = d->cols.empty() ? nullptr : d->cols.data();
    view->valid = d->valid.empty() ? nullptr : d->valid.data();
    view->dup_of = dv.dup_of;
    view->chrom_span_hint = host->chrom_span_hint;
    view->sorted_by_position = host->sorted_by_position;
    if (dv.n > 0) {
      // dataset statistics at load time (what GMQL's own profiler keeps per dataset).  The view hands the spans on as
      // chrom_span_hint, so operators on the resident dataset know the genome's layout up front; the flags let them choose
      // their path without a pass over the data.
      DevBuf<unsigned long long> span(ctx, dv.n_chrom);
      DevBuf<int> fl(ctx, 4);
      span.zero(); fl.zero();
      const int nb = (int)std::min<int64_t>(ceil_div64(dv.n, 256 * 8), (int64_t)ctx->num_sms * 8);
      const size_t sh = dv.n_chrom <= 4096 ? sizeof(unsigned long long) * (size_t)dv.n_chrom : 0;
      LAUNCH(ctx, k_dataset_stats, nb, 256, sh, dv.chrom_narrow ? dv.chrom_narrow : (const void*)dv.chrom, dv.chrom_narrow ? dv.chrom_bits : 32, dv.start, dv.coord_bits,
             dv.len_narrow ? dv.len_narrow : dv.stop, dv.len_narrow ? dv.len_bits : 0, dv.strand, dv.sample, dv.n, dv.n_chrom, dv.n_samples, span.p, fl.p);
      unsigned long long* hs = (unsigned long long*)ctx_pinned(ctx, 8 * (size_t)dv.n_chrom);
      int* hf = (int*)ctx_pinned(ctx, 4 * sizeof(int));
      CUDA_TRY(cudaMemcpyAsync(hs, span.p, 8 * (size_t)dv.n_chrom, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(hf, fl.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(GMQL_SYNC(ctx));
      if (!hf[2]) {                                       // invalid rows: no statistics, no hint; the operators report them
        d->stats_ok = true;
        d->has_star = hf[0] != 0 || dv.strand == nullptr;
        d->has_zero = hf[1] != 0;
        d->maxlen = hf[3];
        bool hint_ok = host->chrom_span_hint != nullptr;
        if (hint_ok) for (int c = 0; c < dv.n_chrom; ++c) if (host->chrom_span_hint[c] < (int64_t)hs[c]) hint_ok = false;
        if (hint_ok) d->span_hint.assign(host->chrom_span_hint, host->chrom_span_hint + dv.n_chrom);   // the caller's bounds hold: keep its layout
        else if (!host->chrom_span_hint) d->span_hint.assign(hs, hs + dv.n_chrom);
        if (!d->span_hint.empty()) {
          view->chrom_span_hint = d->span_hint.data();
          gmql_cover_prepare_dataset(ctx, d, 2000);       // BinSize.Cover's default (core/BinSize.scala:14); other bin sizes recount
        }
      }
    }
    ctx->datasets.push_back(d);
    *out = d;
    return GMQL_OK;
  } catch (const gmql_error& ex) {
    ctx->err = ex.what();
    delete d;
    return ex.code;
  }
}

extern "C" void gmql_dataset_free(gmql_ctx* ctx, gmql_dataset* d) {
  (void)ctx;
  if (!d) return;
  if (d->owner) {
    auto& tg = d->owner->trusted_groupings;
    for (size_t k = 0; k < tg.size(); ++k)
      if (tg[k].first == (const void*)d->dev.sample && tg[k].second == (const void*)d->sample_offsets.data()) { tg.erase(tg.begin() + (long)k); break; }
    auto& ds = d->owner->datasets;
    for (size_t k = 0; k < ds.size(); ++k) if (ds[k] == d) { ds.erase(ds.begin() + (long)k); break; }
  }
  delete d;  // DevBuf destructors hand the buffers back to the owning context's arena
}

extern "C" int gmql_ctx_set_profiling(gmql_ctx* c, int enabled) {
  if (!c) return GMQL_ERR_INVALID;
  c->profiling = enabled != 0;
  return GMQL_OK;
}

extern "C" int gmql_ctx_profile_report(gmql_ctx* c, char* buf, size_t cap) {
  if (!c || !buf || cap == 0) return GMQL_ERR_INVALID;
  cudaStreamSynchronize(c->stream);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : c->prof) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    auto& e = agg[r.name];
    e.first++;
    e.second += ms;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c->prof.clear();
  std::vector<std::pair<std::string, std::pair<int, double>>> v(agg.begin(), agg.end());
  std::sort(v.begin(), v.end(), [](const auto& x, const auto& y) { return x.second.second > y.second.second; });
  std::string s;
  for (auto& e : v) {
    char line[512];
    snprintf(line, sizeof line, "%s %d %.6f\n", e.first.c_str(), e.second.first, e.second.second);
    s += line;
  }
  size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return GMQL_OK;
}
