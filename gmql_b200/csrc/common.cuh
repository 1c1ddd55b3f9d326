// common.cuh — context, device buffers, launch/error helpers shared by the operator translation units.
// Product code: no CPU fallback anywhere; every CUDA failure surfaces as a negative gmql_status.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>
#include <stdexcept>
#include "../../include/gmql_cuda.h"

#define GMQL_NUM_SMS_DEFAULT 148

struct gmql_error : public std::runtime_error {
  int code;
  gmql_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct gmql_dataset;
struct gmql_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = GMQL_NUM_SMS_DEFAULT;
  std::string err;
  int64_t launches = 0;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_h0 = nullptr, ev_h1 = nullptr;
  double last_device_ms = 0.0, last_h2d_ms = 0.0;
  // optional per-launch profiling
  bool profiling = false;
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  // pinned host scratch for small host->device tables (truly asynchronous copies); slots are reused from the start of
  // every API call, when the stream is idle
  std::vector<std::pair<void*, size_t>> pins;
  size_t pin_next = 0;
  bool pending = false;   // a kernel has been launched since the last stream synchronisation
  // (sample column, sample_offsets) pairs of the views handed out by gmql_dataset_upload: the column was expanded from those
  // very offsets on the device, so the layout promise needs no re-check when such a view comes back
  std::vector<std::pair<const void*, const void*>> trusted_groupings;
  // per-context options (gmql_ctx_set_option): test and tuning hooks live here, never in process-global state
  std::map<std::string, std::string> options;
  // Device memory: a private stream-ordered pool (nothing is left reserved in the process's default pool) under an
  // exact-size cache of released buffers.  Everything of a context runs on ONE stream, so a buffer released on the host
  // can be handed to the next request at once: whatever still uses it was queued earlier on that stream.  In steady state
  // (the same operators on datasets of the same shape) no allocation reaches the driver at all.
  cudaMemPool_t pool = nullptr;
  std::multimap<size_t, void*> arena_free;
  std::unordered_map<void*, size_t> arena_live;   // every buffer handed out and not yet returned -> its size
  size_t arena_cached = 0;           // bytes held in arena_free
  size_t arena_limit = 0;            // cached bytes above which the cache is emptied on the next miss
  int64_t arena_hits = 0, arena_misses = 0;
  // datasets made resident by gmql_dataset_upload on this context (their upload-time statistics license operator fast paths)
  std::vector<struct gmql_dataset*> datasets;
  // deferred validation (option "defer_check" = "1"): device flags of operators that returned without their last synchronisation
  struct Deferred { const int* h_flag; const char* what; };
  std::vector<Deferred> deferred;
  int* h_defer = nullptr;            // pinned ring of flag words
  int h_defer_next = 0;
  bool last_deferred = false;        // the operator call in progress returned without its last synchronisation
};

static inline const char* ctx_opt(const gmql_ctx* c, const char* name) {
  auto it = c->options.find(name);
  return it == c->options.end() ? nullptr : it->second.c_str();
}

// Start of an operator call.  Pinned scratch slots are recycled only when the stream is known to be idle (an operator that
// returned without its last synchronisation, option "defer_check", may still be reading them); the validity flags of such
// operators are looked at as soon as they have reached the host.
int gmql_check_deferred(gmql_ctx* c);   // GMQL_OK, or GMQL_ERR_INVALID with the message set (gmql_ctx.cu)
static inline void ctx_begin_call(gmql_ctx* c) {
  if (!c->pending) {
    c->pin_next = 0;
    if (!c->deferred.empty() && gmql_check_deferred(c) != GMQL_OK) throw gmql_error(GMQL_ERR_INVALID, c->err);
  }
}

// ---- device memory ------------------------------------------------------------------------------------------------
static inline void arena_trim(gmql_ctx* c) {
  for (auto& kv : c->arena_free) cudaFreeAsync(kv.second, c->stream);
  c->arena_free.clear();
  c->arena_cached = 0;
}
static inline void* ctx_malloc(gmql_ctx* c, size_t bytes) {
  if (bytes == 0) return nullptr;
  bytes = (bytes + 255) & ~(size_t)255;
  auto it = c->arena_free.find(bytes);
  if (it != c->arena_free.end()) {
    void* p = it->second;
    c->arena_free.erase(it);
    c->arena_cached -= bytes;
    c->arena_hits++;
    c->arena_live[p] = bytes;
    return p;
  }
  c->arena_misses++;
  if (c->arena_cached > c->arena_limit) arena_trim(c);
  void* q = nullptr;
  cudaError_t e = c->pool ? cudaMallocFromPoolAsync(&q, bytes, c->pool, c->stream) : cudaMallocAsync(&q, bytes, c->stream);
  if (e != cudaSuccess && !c->arena_free.empty()) {      // out of memory with buffers cached: give them back and retry
    cudaGetLastError();
    arena_trim(c);
    cudaStreamSynchronize(c->stream);
    e = c->pool ? cudaMallocFromPoolAsync(&q, bytes, c->pool, c->stream) : cudaMallocAsync(&q, bytes, c->stream);
  }
  if (e != cudaSuccess) {
    char b[256];
    snprintf(b, sizeof b, "device allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
    cudaGetLastError();
    throw gmql_error(GMQL_ERR_NOMEM, b);
  }
  c->arena_live[q] = bytes;
  return q;
}
static inline void ctx_free(gmql_ctx* c, void* p) {
  if (!p) return;
  auto it = c->arena_live.find(p);
  if (it == c->arena_live.end()) { cudaFreeAsync(p, c->stream); return; }   // not ours (never happens for DevBuf memory)
  const size_t bytes = it->second;
  c->arena_live.erase(it);
  c->arena_free.insert(std::make_pair(bytes, p));
  c->arena_cached += bytes;
}

// pinned host memory valid until the next API call on this context
static inline void* ctx_pinned(gmql_ctx* c, size_t bytes) {
  if (bytes < 4096) bytes = 4096;
  if (c->pin_next < c->pins.size()) {
    auto& slot = c->pins[c->pin_next];
    if (slot.second < bytes) {
      cudaFreeHost(slot.first);
      slot.first = nullptr; slot.second = 0;
      if (cudaHostAlloc(&slot.first, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw gmql_error(GMQL_ERR_NOMEM, "pinned host allocation failed"); }
      slot.second = bytes;
    }
    return c->pins[c->pin_next++].first;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw gmql_error(GMQL_ERR_NOMEM, "pinned host allocation failed"); }
  c->pins.push_back(std::make_pair(p, bytes));
  c->pin_next = c->pins.size();
  return p;
}

static inline void gmql_prof_begin(gmql_ctx* c, const char* name) {
  gmql_ctx::ProfRec r; r.name = name;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c->stream);
  c->prof.push_back(r);
}
static inline void gmql_prof_end(gmql_ctx* c) { cudaEventRecord(c->prof.back().b, c->stream); }

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw gmql_error(GMQL_ERR_CUDA, _b);                                                    \
    }                                                                                         \
  } while (0)

#define GMQL_REQUIRE(cond, code, msg)                                                         \
  do {                                                                                        \
    if (!(cond)) throw gmql_error((code), std::string(msg));                                  \
  } while (0)

// host-side trace (GMQL_TRACE=1): wall-clock microseconds of every launch / synchronisation, to stderr
#include <chrono>
#include <cstdlib>
inline bool gmql_trace_on() { static int on = -1; if (on < 0) { const char* e = getenv("GMQL_TRACE"); on = (e && *e == '1') ? 1 : 0; } return on == 1; }
inline void gmql_trace(const char* what) {
  if (!gmql_trace_on()) return;
  static auto t0 = std::chrono::steady_clock::now();
  fprintf(stderr, "[trace] %10.1f us  %s\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(), what);
}

// kernel launch on the context stream; counts launches for bench.py's gpu_launches
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                           \
  do {                                                                                        \
    gmql_trace(#kernel);                                                                      \
    if ((ctx)->profiling) gmql_prof_begin((ctx), #kernel);                                    \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                          \
    if ((ctx)->profiling) gmql_prof_end((ctx));                                               \
    (ctx)->launches++;                                                                        \
    (ctx)->pending = true;                                                                    \
    CUDA_TRY(cudaGetLastError());                                                             \
  } while (0)

#define GMQL_SYNC(ctx) ([&]() { gmql_trace("sync>"); cudaError_t _e = cudaStreamSynchronize((ctx)->stream); (ctx)->pending = false; gmql_trace("<sync"); return _e; }())
// end of an operator: everything the caller can observe is complete unless a kernel was launched after the last synchronisation
// (stream-ordered frees and event records may still be queued; consumers of the result use the same stream)
#define GMQL_SYNC_IF_PENDING(ctx) ((ctx)->pending ? GMQL_SYNC(ctx) : cudaSuccess)

// stream-ordered device buffer (cudaMallocAsync pool: frees do not synchronise)
template <typename T>
struct DevBuf {
  gmql_ctx* ctx = nullptr;
  T* p = nullptr;
  int64_t n = 0;
  DevBuf() {}
  DevBuf(gmql_ctx* c, int64_t count) { alloc(c, count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  void alloc(gmql_ctx* c, int64_t count) {
    release();
    ctx = c; n = count;
    if (count > 0) p = (T*)ctx_malloc(c, sizeof(T) * (size_t)count);
  }
  void zero() { if (p) CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, ctx->stream)); }
  void fill_byte(int v) { if (p) CUDA_TRY(cudaMemsetAsync(p, v, sizeof(T) * (size_t)n, ctx->stream)); }
  void release() {
    if (p) { ctx_free(ctx, p); p = nullptr; }
    n = 0;
  }
  T* detach() { T* q = p; p = nullptr; n = 0; return q; }
  ~DevBuf() { release(); }
};

// A dataset resident on the device (either adopted device pointers or uploaded copies).
struct DevRegions {
  int64_t n = 0;
  int coord_bits = 32;
  int n_chrom = 0, n_samples = 0, n_cols = 0;
  const int32_t* chrom = nullptr;
  const void* start = nullptr;
  const void* stop = nullptr;
  const int8_t* strand = nullptr;   // may be null
  const int32_t* sample = nullptr;  // may be null
  const int32_t* dup_of = nullptr;  // may be null
  std::vector<const double*> cols;
  std::vector<const uint8_t*> valid;
  std::vector<DevBuf<uint8_t>> owned;  // uploaded copies
  int64_t h2d_bytes = 0;
  const int64_t* span_hint = nullptr;  // host [n_chrom] upper bounds of stop, or null
  bool sorted_hint = false;            // rows ordered by (chrom, start) according to the caller (verified on the device)
  const int64_t* sample_off = nullptr; // device [n_samples + 1] when the rows are grouped by sample (sample_offsets given), else null
  int64_t max_sample_rows = 0;         // largest sample (rows), valid with sample_off
  bool sample_trusted = false;         // the sample column is known to match sample_off (expanded from it by this library)
  // The ABI's narrow encodings stay as they arrived: the operators' fast paths (tile-resident COVER, per-sample COUNT of MAP)
  // read them directly; gmql_ensure_wide() materialises the int32 chromosome / stop / per-row sample columns for everything
  // else (once per resident dataset, once per call otherwise).  While !wide_ready, chrom / stop / sample may be null.
  const void* chrom_narrow = nullptr;  // uint8_t[] / uint16_t[] when chrom_bits is 8 / 16
  int chrom_bits = 32;
  const void* len_narrow = nullptr;    // uint16_t[] / uint32_t[] region lengths when len_bits is 16 / 32
  int len_bits = 0;
  bool wide_ready = true;
  struct gmql_dataset* dataset = nullptr;   // the resident dataset these arrays belong to (gmql_dataset_upload), or null
  bool host_sample_col = false;        // the caller supplied a per-row sample column (not only sample_offsets)
};

// upload (or adopt) a gmql_regions; only the value columns listed in need_cols are moved.  keep_narrow: leave narrow encodings
// as they are (the caller has a fast path for them, or calls gmql_ensure_wide itself)
void gmql_upload_regions(gmql_ctx* ctx, const gmql_regions* r, const std::vector<int>& need_cols, DevRegions* out, bool keep_narrow = false);
void gmql_ensure_wide(gmql_ctx* ctx, DevRegions* d);

// A dataset made resident by gmql_dataset_upload: the raw arrays in the encodings they arrived in, what the upload learned
// about them (the statistics GMQL's own profiler keeps per dataset) and what operators derive once per dataset.
struct gmql_dataset {
  DevRegions dev;
  std::vector<int64_t> sample_ids;
  std::vector<int64_t> sample_offsets;   // host copy when the upload declared the rows grouped by sample
  gmql_ctx* owner = nullptr;
  std::vector<int64_t> span_hint;        // largest stop per chromosome (or the caller's own hint)
  std::vector<const double*> cols;
  std::vector<const uint8_t*> valid;
  // upload-time statistics over every row (stats_ok: all rows valid)
  bool stats_ok = false, has_zero = false, has_star = false;
  int64_t maxlen = 0;
  // What COVER's tile path derives once per dataset (gmql_cover.cu): the linear layout coff = prefix(span_hint + lay_pad) and,
  // for tiles of 65 536 positions, the first record of every tile (tile_off, the scanned tile histogram) and the start of every
  // (coarse bin, input-tile group) range of the coarse partition pass (coarse_off)
  DevBuf<uint64_t> lay_coff;
  DevBuf<unsigned long long> lay_span;
  uint64_t lay_pad = 0, lay_G = 0;
  DevBuf<uint32_t> tile_off, coarse_off;
  int64_t hist_nT = 0; int hist_groups = 0, hist_n_coarse = 0;
  // wide columns, materialised on first demand
  DevBuf<uint8_t> wide_chrom, wide_stop, wide_sample;
  bool wide_done = false;
};
// fills the dataset's tile histogram (defined in gmql_cover.cu next to the kernels that consume it)
void gmql_cover_prepare_dataset(gmql_ctx* ctx, gmql_dataset* d, int64_t bin_size);

// result object: named device columns
struct ResultCol {
  void* p = nullptr;
  int64_t len = 0;
  int32_t elem = 0;
};
struct gmql_result {
  gmql_ctx* owner = nullptr;       // the context whose arena holds the columns
  int64_t rows = 0;
  int kind = 0;  // 1 map, 2 cover, 3 join
  std::map<int, ResultCol> cols;
  // cover view support
  int32_t n_chrom = 0;
  std::vector<int64_t> group_ids;
  std::vector<const double*> view_cols;
  std::vector<const uint8_t*> view_valid;
  void* view_acc_double = nullptr;
  std::vector<int64_t> span_hint;    // cover results: upper bound of stop per chromosome (host)
  bool sorted = false;               // cover results: rows ordered by (chrom, start) (single class)
  bool unified = false;              // cover results: strands unified, every row '*'
};
template <typename T>
static inline void result_add(gmql_result* r, int id, DevBuf<T>& b) {
  ResultCol c; c.len = b.n; c.elem = (int32_t)sizeof(T); c.p = (void*)b.detach();
  r->cols[id] = c;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ int64_t ld_coord(const void* p, int bits, int64_t i) {
  return bits == 32 ? (int64_t)((const int32_t*)p)[i] : ((const int64_t*)p)[i];
}

__device__ __forceinline__ bool strand_ok(int a, int b) { return a == 0 || b == 0 || a == b; }

// timing brackets: device time from the first kernel to the last (CUDA events on the context stream)
struct DeviceTimer {
  gmql_ctx* ctx;
  explicit DeviceTimer(gmql_ctx* c) : ctx(c) { cudaEventRecord(c->ev_k0, c->stream); }
  void stop() { cudaEventRecord(ctx->ev_k1, ctx->stream); }
};
