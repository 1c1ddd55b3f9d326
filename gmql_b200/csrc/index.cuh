// index.cuh — the interval index every operator searches: regions of the *indexed* side sorted by a linearised
// start coordinate, with the running maximum of the linearised stops ("running max-stop array", north_star).
//
// Linearisation: lin(block, chrom, pos) = block * G + coff[chrom] + pos, where coff is the exclusive prefix sum of
// (span[chrom] + 2) and G its total; span[chrom] = max stop seen on that chromosome.  `block` separates
// (group, strand) classes for COVER and is 0 for MAP/JOIN.  Because linear stops of earlier chromosomes are
// always below the next chromosome's offset, the running max needs no segmentation.
//
// Query (streamed side): all indexed regions x with x.start < q.stop && x.stop > q.start are found by
//   hi = lower_bound(start, lin(q.stop));  for k = hi-1 down while pmax[k] > lin(q.start): test stop[k] > lin(q.start).
#pragma once
#include "prims.cuh"

struct LinMap {
  const uint64_t* coff;  // [n_chrom+1] device
  const int64_t* span;   // [n_chrom] device
  uint64_t G;            // genome total (host copy)
  int n_chrom;
};

static __global__ void k_chrom_span(const int32_t* __restrict__ chrom, const void* __restrict__ stop, int bits, int64_t n, int n_chrom, unsigned long long* __restrict__ span, int* __restrict__ bad) {
  extern __shared__ unsigned long long sspan[];
  bool use_sh = n_chrom <= 4096;
  if (use_sh) {
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) sspan[c] = 0ull;
    __syncthreads();
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = chrom[i];
    int64_t e = ld_coord(stop, bits, i);
    if (c < 0 || c >= n_chrom || e < 0) { *bad = 1; continue; }
    if (use_sh) atomicMax(&sspan[c], (unsigned long long)e);
    else atomicMax(&span[c], (unsigned long long)e);
  }
  if (use_sh) {
    __syncthreads();
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x)
      if (sspan[c]) atomicMax(&span[c], sspan[c]);
  }
}

static __global__ void k_chrom_offsets(const unsigned long long* __restrict__ span, int n_chrom, uint64_t* __restrict__ coff, unsigned long long pad) {
  // tiny: one thread (n_chrom is tens to a few thousand)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint64_t run = 0;
    for (int c = 0; c < n_chrom; ++c) { coff[c] = run; run += span[c] + pad; }
    coff[n_chrom] = run;
  }
}

// Accumulates max stop per chromosome of `r` into span (device, zero-initialised by the caller).
static void accumulate_spans(gmql_ctx* ctx, const DevRegions& r, unsigned long long* span, int* bad) {
  if (r.n == 0) return;
  int nb = (int)std::min<int64_t>(ceil_div64(r.n, 256 * 8), (int64_t)ctx->num_sms * 8);
  size_t sh = r.n_chrom <= 4096 ? sizeof(unsigned long long) * (size_t)r.n_chrom : 0;
  LAUNCH(ctx, k_chrom_span, nb, 256, sh, r.chrom, r.stop, r.coord_bits, r.n, r.n_chrom, span, bad);
}

struct SpanTable {
  DevBuf<unsigned long long> span;  // [n_chrom]
  DevBuf<uint64_t> coff;            // [n_chrom+1]
  DevBuf<int> bad;
  uint64_t G = 0;
  int n_chrom = 0;
  LinMap map() const { LinMap m; m.coff = coff.p; m.span = (const int64_t*)span.p; m.G = G; m.n_chrom = n_chrom; return m; }
};

// builds the chromosome offset table from one or two datasets; synchronises once to read G
static void build_span_table(gmql_ctx* ctx, const DevRegions* a, const DevRegions* b, SpanTable* t, unsigned long long pad = 2ull) {
  int n_chrom = a->n_chrom;
  t->n_chrom = n_chrom;
  t->span.alloc(ctx, n_chrom);
  t->span.zero();
  t->coff.alloc(ctx, n_chrom + 1);
  t->bad.alloc(ctx, 1);
  t->bad.zero();
  accumulate_spans(ctx, *a, t->span.p, t->bad.p);
  if (b) accumulate_spans(ctx, *b, t->span.p, t->bad.p);
  LAUNCH(ctx, k_chrom_offsets, 1, 32, 0, t->span.p, n_chrom, t->coff.p, pad);
  uint64_t G = 0;
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&G, t->coff.p + n_chrom, sizeof G, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&bad, t->bad.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(!bad, GMQL_ERR_INVALID, "region with chromosome index out of range or negative coordinate");
  t->G = G;
}

// span table from a caller-supplied upper bound per chromosome: no pass over the data, no host round trip.  Regions beyond
// their bound must be detected by the consumer (k_index_stops' `bad`).
static void span_table_from_hint(gmql_ctx* ctx, const int64_t* hint, int n_chrom, SpanTable* t, unsigned long long pad = 2ull) {
  t->n_chrom = n_chrom;
  t->span.alloc(ctx, n_chrom);
  t->coff.alloc(ctx, n_chrom + 1);
  t->bad.alloc(ctx, 1);
  t->bad.zero();
  unsigned long long* hs = (unsigned long long*)ctx_pinned(ctx, sizeof(unsigned long long) * (size_t)n_chrom);
  uint64_t* hc = (uint64_t*)ctx_pinned(ctx, sizeof(uint64_t) * ((size_t)n_chrom + 1));
  uint64_t run = 0;
  for (int c = 0; c < n_chrom; ++c) {
    GMQL_REQUIRE(hint[c] >= 0, GMQL_ERR_INVALID, "negative chrom_span_hint");
    hs[c] = (unsigned long long)hint[c]; hc[c] = run; run += hs[c] + pad;
  }
  hc[n_chrom] = run;
  CUDA_TRY(cudaMemcpyAsync(t->span.p, hs, sizeof(unsigned long long) * (size_t)n_chrom, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(t->coff.p, hc, sizeof(uint64_t) * ((size_t)n_chrom + 1), cudaMemcpyHostToDevice, ctx->stream));
  t->G = run;
}

template <typename K>
struct IntervalIndex {
  int64_t n = 0;
  DevBuf<K> start_a, start_b;        // ping-pong key buffers; `start` points at the sorted one
  DevBuf<uint32_t> row_a, row_b;
  K* start = nullptr;                // [n] sorted linear starts
  uint32_t* row = nullptr;           // [n] original row of each sorted entry
  DevBuf<K> stop;                    // [n] linear stop of each sorted entry
  DevBuf<K> pmax;                    // [n] running max of stop
};

// block_of(i) = class block of row i (0 when blocks == nullptr)
// payload = original row, or (stop - start) when `stop` is given (no later gather needed; rows are then unavailable)
template <typename K>
__global__ void k_index_keys(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ blocks,
                             int64_t n, LinMap lm, K* __restrict__ keys, uint32_t* __restrict__ rows, int* __restrict__ bad = nullptr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t blk = blocks ? blocks[i] : 0u;
    int64_t s = ld_coord(start, bits, i);
    int c = chrom[i];
    // span tables built from a caller's hint have seen no row yet: an out-of-range chromosome or a negative start is
    // reported (and the row parked at the origin), never dereferenced
    if (bad && (c < 0 || c >= lm.n_chrom || s < 0)) { *bad = 1; c = 0; s = 0; }
    uint64_t lin = blk * lm.G + lm.coff[c] + (uint64_t)s;
    keys[i] = (K)lin;
    rows[i] = stop ? (uint32_t)(ld_coord(stop, bits, i) - s) : (uint32_t)i;
  }
}

template <typename K>
__global__ void k_index_stops_from_len(const K* __restrict__ starts, const uint32_t* __restrict__ lens, int64_t n, K* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = starts[k] + (K)lens[k];
}

// bad (optional): set when a region lies outside its chromosome's span (span tables built from a caller's hint)
template <typename K>
__global__ void k_index_stops(const int32_t* __restrict__ chrom, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ blocks,
                              int64_t n, LinMap lm, const uint32_t* __restrict__ rows, K* __restrict__ out, int* __restrict__ bad, const K* __restrict__ check_sorted) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    if (check_sorted && k > 0 && check_sorted[k - 1] > check_sorted[k]) *bad = 1;   // caller promised sorted rows
    uint32_t i = rows[k];
    uint64_t blk = blocks ? blocks[i] : 0u;
    int c = chrom[i];
    int64_t e = ld_coord(stop, bits, i);
    if (bad && (c < 0 || c >= lm.n_chrom || e < 0 || e > lm.span[c])) { *bad = 1; c = 0; e = 0; }
    out[k] = (K)(blk * lm.G + lm.coff[c] + (uint64_t)e);
  }
}

static inline int grid_for(gmql_ctx* ctx, int64_t n, int threads, int per_thread = 4) {
  int64_t nb = ceil_div64(n, (int64_t)threads * per_thread);
  int64_t cap = (int64_t)ctx->num_sms * 16;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

template <typename K>
static void alloc_index(gmql_ctx* ctx, int64_t n, IntervalIndex<K>* ix) {
  ix->n = n;
  ix->start_a.alloc(ctx, n); ix->start_b.alloc(ctx, n);
  ix->row_a.alloc(ctx, n); ix->row_b.alloc(ctx, n);
  ix->stop.alloc(ctx, n); ix->pmax.alloc(ctx, n);
  ix->start = ix->start_a.p; ix->row = ix->row_a.p;
}

// start_a / row_a hold the unsorted keys and payloads (row, or length when payload_len): sort, derive stops, running max
template <typename K>
static void finish_index(gmql_ctx* ctx, const DevRegions& r, const uint32_t* blocks, uint64_t n_blocks, const LinMap& lm, IntervalIndex<K>* ix, bool payload_len, int* bad = nullptr) {
  int64_t n = ix->n;
  if (n == 0) return;
  int g = grid_for(ctx, n, 256);
  uint64_t max_lin = n_blocks * lm.G;
  // rows promised to be in (chrom, start) order need no sort (single block of the linear coordinate only); the promise is
  // checked by k_index_stops
  const bool presorted = r.sorted_hint && blocks == nullptr && !payload_len && bad != nullptr;
  if (!presorted) radix_sort<K>(ctx, ix->start_a.p, ix->start_b.p, ix->row_a.p, ix->row_b.p, n, bits_for(max_lin), &ix->start, &ix->row);
  if (payload_len) LAUNCH(ctx, (k_index_stops_from_len<K>), g, 256, 0, ix->start, ix->row, n, ix->stop.p);
  else LAUNCH(ctx, (k_index_stops<K>), g, 256, 0, r.chrom, r.stop, r.coord_bits, blocks, n, lm, ix->row, ix->stop.p, bad, presorted ? (const K*)ix->start : (const K*)nullptr);
  device_scan<K, K, OpMax<K>, true>(ctx, ix->stop.p, ix->pmax.p, n, (K*)nullptr);
}

template <typename K>
static void build_index(gmql_ctx* ctx, const DevRegions& r, const uint32_t* blocks, uint64_t n_blocks, const LinMap& lm, IntervalIndex<K>* ix, bool payload_len = false, int* bad = nullptr) {
  alloc_index<K>(ctx, r.n, ix);
  if (r.n == 0) return;
  // payload_len requires region lengths < 2^32 (checked by the caller through the span table)
  LAUNCH(ctx, (k_index_keys<K>), grid_for(ctx, r.n, 256), 256, 0, r.chrom, r.start, payload_len ? r.stop : (const void*)nullptr, r.coord_bits, blocks, r.n, lm, ix->start_a.p, ix->row_a.p, bad);
  finish_index<K>(ctx, r, blocks, n_blocks, lm, ix, payload_len, bad);
}
