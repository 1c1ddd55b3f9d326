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
    // keep freed blocks cached in the stream-ordered pool (operators allocate scratch every call)
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
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
  if (c->ev_k0) cudaEventDestroy(c->ev_k0);
  if (c->ev_k1) cudaEventDestroy(c->ev_k1);
  if (c->ev_h0) cudaEventDestroy(c->ev_h0);
  if (c->ev_h1) cudaEventDestroy(c->ev_h1);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  for (auto& pin : c->pins) if (pin.first) cudaFreeHost(pin.first);
  delete c;
}

extern "C" const char* gmql_last_error(const gmql_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int gmql_ctx_sync(gmql_ctx* c) {
  if (!c) return GMQL_ERR_INVALID;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return GMQL_ERR_CUDA; }
  return GMQL_OK;
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

void gmql_upload_regions(gmql_ctx* ctx, const gmql_regions* r, const std::vector<int>& need_cols, DevRegions* out) {
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
  size_t cb = (size_t)(r->coord_bits / 8);
  size_t n = (size_t)r->n;
  GMQL_REQUIRE(r->chrom_bits == 0 || r->chrom_bits == 8 || r->chrom_bits == 16 || r->chrom_bits == 32, GMQL_ERR_INVALID, "chrom_bits must be 0, 8, 16 or 32");
  if (r->chrom_bits == 8 || r->chrom_bits == 16) {
    GMQL_REQUIRE(r->n_chrom <= (r->chrom_bits == 8 ? 256 : 65536), GMQL_ERR_INVALID, "n_chrom does not fit chrom_bits");
    const size_t cw = (size_t)(r->chrom_bits / 8);
    const void* raw = upload_array(ctx, r->chrom, n * cw, r->location, out);
    DevBuf<uint8_t> wide(ctx, (int64_t)(n * 4 ? n * 4 : 1));
    if (n) {
      int g = (int)std::min<int64_t>(ceil_div64((int64_t)n, 256 * 4), (int64_t)ctx->num_sms * 16);
      if (r->chrom_bits == 8) LAUNCH(ctx, (k_widen_chrom<uint8_t>), g, 256, 0, (const uint8_t*)raw, (int64_t)n, (int32_t*)wide.p);
      else LAUNCH(ctx, (k_widen_chrom<uint16_t>), g, 256, 0, (const uint16_t*)raw, (int64_t)n, (int32_t*)wide.p);
    }
    out->chrom = (const int32_t*)wide.p;
    if (r->location != GMQL_LOC_DEVICE) out->owned.pop_back();   // the narrow copy is released in stream order
    out->owned.push_back(std::move(wide));
  } else {
    out->chrom = (const int32_t*)upload_array(ctx, r->chrom, n * 4, r->location, out);
  }
  out->start = upload_array(ctx, r->start, n * cb, r->location, out);
  GMQL_REQUIRE(r->stop_is_length == 0 || r->stop_is_length == 16 || r->stop_is_length == 32, GMQL_ERR_INVALID, "stop_is_length must be 0, 16 or 32");
  if (r->stop_is_length) {
    const size_t lw = (size_t)(r->stop_is_length / 8);
    const void* raw = upload_array(ctx, r->stop, n * lw, r->location, out);
    DevBuf<uint8_t> full(ctx, (int64_t)(n * cb ? n * cb : 1));
    if (n) {
      int g = (int)std::min<int64_t>(ceil_div64((int64_t)n, 256 * 4), (int64_t)ctx->num_sms * 16);
      if (cb == 4 && lw == 2) LAUNCH(ctx, (k_stop_from_length<int32_t, uint16_t>), g, 256, 0, (const int32_t*)out->start, (const uint16_t*)raw, (int64_t)n, (int32_t*)full.p);
      else if (cb == 4) LAUNCH(ctx, (k_stop_from_length<int32_t, uint32_t>), g, 256, 0, (const int32_t*)out->start, (const uint32_t*)raw, (int64_t)n, (int32_t*)full.p);
      else if (lw == 2) LAUNCH(ctx, (k_stop_from_length<int64_t, uint16_t>), g, 256, 0, (const int64_t*)out->start, (const uint16_t*)raw, (int64_t)n, (int64_t*)full.p);
      else LAUNCH(ctx, (k_stop_from_length<int64_t, uint32_t>), g, 256, 0, (const int64_t*)out->start, (const uint32_t*)raw, (int64_t)n, (int64_t*)full.p);
    }
    out->stop = full.p;
    if (r->location != GMQL_LOC_DEVICE) out->owned.pop_back();   // the length copy is released in stream order
    out->owned.push_back(std::move(full));
  } else {
    out->stop = upload_array(ctx, r->stop, n * cb, r->location, out);
  }
  out->strand = (const int8_t*)upload_array(ctx, r->strand, n, r->location, out);
  if (r->sample_offsets) {
    // rows grouped by sample.  Without a sample column the offsets define it; with one they are a promise about its
    // layout, which the kernels that rely on it check against the column.
    const int64_t* off = r->sample_offsets;   // host array by contract
    GMQL_REQUIRE(off[0] == 0 && off[r->n_samples] == r->n, GMQL_ERR_INVALID, "sample_offsets must start at 0 and end at n");
    int64_t widest = 0;
    for (int s = 0; s < r->n_samples; ++s) {
      GMQL_REQUIRE(off[s] <= off[s + 1], GMQL_ERR_INVALID, "sample_offsets must be non-decreasing");
      widest = std::max<int64_t>(widest, off[s + 1] - off[s]);
    }
    const size_t ob = 8 * ((size_t)r->n_samples + 1);
    DevBuf<uint8_t> doff(ctx, (int64_t)ob);
    CUDA_TRY(cudaMemcpyAsync(doff.p, off, ob, cudaMemcpyHostToDevice, ctx->stream));   // small pageable source: staged by the driver before the call returns
    if (r->location != GMQL_LOC_DEVICE) out->h2d_bytes += (int64_t)ob;
    out->sample_off = (const int64_t*)doff.p;
    out->max_sample_rows = widest;
    for (const auto& tg : ctx->trusted_groupings) if (tg.first == (const void*)r->sample && tg.second == (const void*)r->sample_offsets && r->sample) out->sample_trusted = true;
    if (!r->sample) {
      out->sample_trusted = true;
      DevBuf<uint8_t> smp(ctx, (int64_t)(n * 4 ? n * 4 : 1));
      if (n) LAUNCH(ctx, k_expand_sample_offsets, (int)std::min<int64_t>(r->n_samples, (int64_t)ctx->num_sms * 16), 256, 0, (const int64_t*)doff.p, r->n_samples, (int32_t*)smp.p);
      out->sample = (const int32_t*)smp.p;
      out->owned.push_back(std::move(smp));
    } else {
      out->sample = (const int32_t*)upload_array(ctx, r->sample, n * 4, r->location, out);
    }
    out->owned.push_back(std::move(doff));
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
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return GMQL_ERR_CUDA; }
  }
  return GMQL_OK;
}

extern "C" void gmql_result_free(gmql_ctx* ctx, gmql_result* r) {
  if (!r) return;
  gmql_trace("result_free");
  for (auto& kv : r->cols)
    if (kv.second.p) { if (ctx) cudaFreeAsync(kv.second.p, ctx->stream); else cudaFree(kv.second.p); }
  if (r->view_acc_double) { if (ctx) cudaFreeAsync(r->view_acc_double, ctx->stream); else cudaFree(r->view_acc_double); }
  delete r;
}

// ------------------------------------------------------------------------------------------------
struct gmql_dataset {
  DevRegions dev;
  std::vector<int64_t> sample_ids;
  std::vector<int64_t> sample_offsets;   // host copy when the upload declared the rows grouped by sample
  gmql_ctx* owner = nullptr;             // context holding this dataset's trusted-grouping entry
  std::vector<int64_t> span_hint;        // largest stop per chromosome, computed at upload when the caller gave no hint
  std::vector<const double*> cols;
  std::vector<const uint8_t*> valid;
};

extern "C" int gmql_dataset_upload(gmql_ctx* ctx, const gmql_regions* host, gmql_dataset** out, gmql_regions* view) {
  if (!ctx || !host || !out || !view) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_dataset* d = new gmql_dataset();
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<int> all;
    for (int c = 0; c < host->n_cols; ++c) all.push_back(c);
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    gmql_upload_regions(ctx, host, all, &d->dev);
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    if (host->sample_ids) d->sample_ids.assign(host->sample_ids, host->sample_ids + host->n_samples);
    else d->sample_ids.assign((size_t)host->n_samples, 0);
    d->cols = d->dev.cols;
    d->valid = d->dev.valid;
    memset(view, 0, sizeof *view);
    view->n = d->dev.n;
    view->location = GMQL_LOC_DEVICE;
    view->coord_bits = d->dev.coord_bits;
    view->chrom = d->dev.chrom;
    view->start = d->dev.start;
    view->stop = d->dev.stop;
    view->strand = d->dev.strand;
    view->sample = d->dev.sample;
    if (host->sample_offsets) {
      d->sample_offsets.assign(host->sample_offsets, host->sample_offsets + host->n_samples + 1);
      view->sample_offsets = d->sample_offsets.data();   // the view keeps the promise: rows grouped by sample
      if (d->dev.sample_trusted) { ctx->trusted_groupings.push_back(std::make_pair((const void*)d->dev.sample, (const void*)d->sample_offsets.data())); d->owner = ctx; }
    }
    view->n_chrom = d->dev.n_chrom;
    view->n_samples = d->dev.n_samples;
    view->sample_ids = d->sample_ids.data();
    view->n_cols = d->dev.n_cols;
    view->cols = d->cols.empty() ? nullptr : d->cols.data();
    view->valid = d->valid.empty() ? nullptr : d->valid.data();
    view->dup_of = d->dev.dup_of;
    view->chrom_span_hint = host->chrom_span_hint;
    if (!host->chrom_span_hint && d->dev.n > 0) {
      // dataset statistics at load time (what GMQL's own profiler keeps per dataset): the largest stop of every chromosome.
      // The view hands them on as chrom_span_hint, so operators on the resident dataset know the genome's layout up front.
      DevBuf<unsigned long long> span(ctx, d->dev.n_chrom);
      DevBuf<int> bad(ctx, 1);
      span.zero(); bad.zero();
      accumulate_spans(ctx, d->dev, span.p, bad.p);
      std::vector<unsigned long long> hs((size_t)d->dev.n_chrom);
      int h_bad = 0;
      CUDA_TRY(cudaMemcpyAsync(hs.data(), span.p, 8 * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(GMQL_SYNC(ctx));
      if (!h_bad) {                                     // invalid rows: no hint; the operators report them
        d->span_hint.assign(hs.begin(), hs.end());
        view->chrom_span_hint = d->span_hint.data();
      }
    }
    view->sorted_by_position = host->sorted_by_position;
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
  if (d && d->owner) {
    auto& tg = d->owner->trusted_groupings;
    for (size_t k = 0; k < tg.size(); ++k)
      if (tg[k].first == (const void*)d->dev.sample && tg[k].second == (const void*)d->sample_offsets.data()) { tg.erase(tg.begin() + (long)k); break; }
  }
  delete d;  // DevBuf destructors release on the owning context's stream
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
