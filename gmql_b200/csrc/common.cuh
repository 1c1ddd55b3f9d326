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
#include <stdexcept>
#include "../../include/gmql_cuda.h"

#define GMQL_NUM_SMS_DEFAULT 148

struct gmql_error : public std::runtime_error {
  int code;
  gmql_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

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
};

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
    if (count > 0) {
      void* q = nullptr;
      cudaError_t e = cudaMallocAsync(&q, sizeof(T) * (size_t)count, c->stream);
      if (e != cudaSuccess) {
        char b[256];
        snprintf(b, sizeof b, "device allocation of %lld bytes failed: %s", (long long)(sizeof(T) * (size_t)count), cudaGetErrorString(e));
        cudaGetLastError();
        throw gmql_error(GMQL_ERR_NOMEM, b);
      }
      p = (T*)q;
    }
  }
  void zero() { if (p) CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, ctx->stream)); }
  void fill_byte(int v) { if (p) CUDA_TRY(cudaMemsetAsync(p, v, sizeof(T) * (size_t)n, ctx->stream)); }
  void release() {
    if (p) { cudaFreeAsync(p, ctx->stream); p = nullptr; }
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
};

// upload (or adopt) a gmql_regions; only the value columns listed in need_cols are moved
void gmql_upload_regions(gmql_ctx* ctx, const gmql_regions* r, const std::vector<int>& need_cols, DevRegions* out);

// result object: named device columns
struct ResultCol {
  void* p = nullptr;
  int64_t len = 0;
  int32_t elem = 0;
};
struct gmql_result {
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
