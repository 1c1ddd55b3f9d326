// cover_tiled.cuh — tile-resident COVER / FLAT (both phases) for dense inputs with 32-bit linear coordinates.
//
// Instead of fully sorting 2n events and merging them, the linear genome is cut into tiles of W = 2^logw positions
// (W <= 65 536).  Regions (linear start, length) are only PARTITIONED by the tile of their start (radix passes on the
// top key bits; the order inside a tile is irrelevant).  One thread block then owns a tile:
//   * its events are the +1 of every region that starts in the tile and the -1 (at stop', the extractor's extended stop,
//     GenometricCover.scala:353-354) of every region of this tile or the previous one whose stop' falls inside; regions
//     are shorter than W, so no other tile can contribute, and the depth entering the tile is simply the number of
//     regions of the previous tile that are still open;
//   * the tile's distinct event positions are marked in a W-bit occupancy bitmap in shared memory; a popcount prefix
//     over the bitmap words RANKS every position, i.e. sorts the distinct positions without sorting anything;
//   * +1 / -1 are summed per rank with shared-memory atomics — coincident events are summed exactly as the reference's
//     reduceByKey does (:98-108) — and a block-wide prefix sum over the (sparse) ranks gives the depth profile;
//   * in-range stretches become PIECES (clipped at tile edges, like the reference's bins :110-118);
//   * GMAP4 (GMAP4.scala:20-100) runs region-driven inside the tile: every region of the tile (and the regions of the
//     previous tile that reach into it) finds the pieces it overlaps and accumulates count / startMin / endMax /
//     startMax / endMin with warp-aggregated shared-memory atomics.  A (run, region) pair is accounted exactly once, in
//     tile max(tile(run.start), tile(region.start)) — the reference's own first-bin rule (GMAP4.scala:44-45) at tile
//     granularity;
//   * a small stitch pass fuses pieces that touch at tile edges (the reference's boundary merge, :121-152).
// Work per tile is proportional to its events (plus W/32 bitmap words), not to W.  Four 256-thread blocks per SM (54 KB of shared memory each).
// Preconditions (checked by the caller, which otherwise uses the sort-and-merge path): uint32 keys, COVER or FLAT,
// no zero-length regions, max region length < W - 2 (value aggregates are a post-pass of the caller over the runs).  Capacity overflows (distinct positions, pieces or
// halo regions of one tile) raise `fail` and the caller falls back as well.
#pragma once
#include <cooperative_groups.h>
#include "index.cuh"

namespace tiled {

constexpr int MAX_LOGW = 16;
// Four 256-thread blocks per SM (64 registers per thread, 54 KB of shared memory each): a tile costs a fixed number of block-wide
// scans and barriers, so fewer threads per tile and more independent tiles in flight per SM beat two 512-thread blocks.
constexpr int THREADS = 256;
constexpr int CTAS_PER_SM = 4;
constexpr int NWARPS = THREADS / 32;
constexpr int MAXWORDS = (1 << MAX_LOGW) / 32;
constexpr int CAP_EV = 13312;           // distinct event positions per tile
constexpr int MAXP = 256;               // pieces per tile
constexpr int HCAP = 512;               // regions of the previous tile still open at the tile start
constexpr int MAX_TILE_REGIONS = 32000; // own + open regions of one tile: keeps every 16-bit net change exact
constexpr uint32_t EXT = 0x80000000u;   // length payload bit: stop' = stop + 1
// Partitioned records are 4 bytes: tile-local start (16 bits) | length (16 bits); regions are shorter than a tile, so both fit.
// The extractor's one-base extension (stop on a bin multiple, GenometricCover.scala:353-354) is a property of the chromosome-local
// stop: the tile kernel recomputes it from the tile's chromosome offset (TileChrom) instead of carrying a bit.
struct TileChrom {                       // chromosome layout around one tile (block-uniform)
  uint32_t lo, next, base;               // linear positions [lo, next) belong to one chromosome of one class, whose position 0 is `base`
};
__device__ __forceinline__ uint32_t lin_base_slow(const LinMap& lm, uint32_t x) {      // linear position of the chromosome start at or before x
  uint64_t cls = 0, rem = x;
  if (rem >= lm.G) { cls = rem / lm.G; rem -= cls * lm.G; }   // one class (the usual case): no division
  int lo = 0, hi = lm.n_chrom;           // last chromosome with coff <= rem
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lm.coff[mid] <= rem) lo = mid; else hi = mid; }
  return (uint32_t)(cls * lm.G + lm.coff[lo]);
}
__device__ __forceinline__ TileChrom tile_chrom(const LinMap& lm, uint32_t tile_lo) {
  uint64_t cls = 0, rem = tile_lo;
  if (rem >= lm.G) { cls = rem / lm.G; rem -= cls * lm.G; }
  int lo = 0, hi = lm.n_chrom;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lm.coff[mid] <= rem) lo = mid; else hi = mid; }
  TileChrom t;
  t.base = (uint32_t)(cls * lm.G + lm.coff[lo]);
  t.lo = t.base;
  const uint64_t nx = cls * lm.G + (lo + 1 < lm.n_chrom ? lm.coff[lo + 1] : lm.G);
  t.next = nx > 0xffffffffull ? 0xffffffffu : (uint32_t)nx;
  return t;
}
struct BinMod { uint64_t magic; int64_t B; };   // x % B == 0 without a division (Lemire, Kaser, Kurz 2019), exact for 32-bit x and B < 2^31
__device__ __forceinline__ BinMod make_binmod(int64_t B) { BinMod m; m.B = B; m.magic = B > 1 ? 0xffffffffffffffffull / (uint64_t)B + 1ull : 0ull; return m; }
__device__ __forceinline__ bool is_bin_multiple(const BinMod& m, uint32_t x) {
  return m.B < 0x7fffffffLL ? (m.magic * (uint64_t)x <= m.magic - 1ull) : ((uint64_t)x % (uint64_t)m.B == 0ull);
}
// x % B == 0 with one multiply-high (x < 2^32): q = floor(x * floor((2^32 - 1) / B) / 2^32) underestimates x / B by at most 2.
// Built on the host, passed by value.
struct BinMod32 { uint32_t B, inv; int big; };
static inline BinMod32 make_binmod32(int64_t B) {
  BinMod32 m;
  m.big = B >= 0x7fffffffLL ? 1 : 0;
  m.B = B > 0xffffffffLL ? 0u : (uint32_t)B;    // a bin wider than any 32-bit coordinate: only 0 is a multiple
  m.inv = m.big ? 0u : 0xffffffffu / m.B;
  return m;
}
__device__ __forceinline__ bool is_bin_multiple32(const BinMod32& m, uint32_t x) {
  if (m.big) return x == 0u || (m.B != 0u && x == m.B);
  const uint32_t r = x - __umulhi(x, m.inv) * m.B;
  return r == 0u || r == m.B || r == 2u * m.B;
}
// length payload of a record that starts at linear position rs: length | EXT when stop' = stop + 1
__device__ __forceinline__ uint32_t rec_len_payload(uint32_t rec, uint32_t rs, const TileChrom& tc, const LinMap& lm, const BinMod& bm, bool xb) {
  if (xb) return (rec & 0x7fffu) | ((rec & 0x8000u) << 16);   // the partition left the extension in bit 15 of the record
  const uint32_t len = rec & 0xffffu;
  const uint32_t base = (rs >= tc.lo && rs < tc.next) ? tc.base : lin_base_slow(lm, rs);
  return len | ((len != 0u && is_bin_multiple(bm, rs + len - base)) ? EXT : 0u);
}
constexpr size_t SMEM_BYTES = 4 * MAXWORDS + 2 * MAXWORDS + 2 * CAP_EV + 16 * HCAP + 32 * MAXP;

// net +1 / -1 per distinct position, two 16-bit signed counters per shared-memory word.  Adding +-1 to the low half may borrow
// from / carry into the high half; the word is always high * 65536 + low (mod 2^32), so decoding the low half as signed and
// removing it first recovers both exactly as long as every net change fits 16 bits (MAX_TILE_REGIONS).
__device__ __forceinline__ void delta_add(uint32_t* d, int r, int v) { atomicAdd(&d[r >> 1], (r & 1) ? ((uint32_t)v << 16) : (uint32_t)v); }
__device__ __forceinline__ int32_t delta_get(const uint32_t* d, int k) {
  const uint32_t w = d[k >> 1];
  const int32_t lo = (int32_t)(int16_t)(w & 0xffffu);
  return (k & 1) ? (int32_t)(int16_t)((w - (uint32_t)lo) >> 16) : lo;
}

struct Cfg {
  int n_groups, n_classes;
  const int32_t* mn;
  const int32_t* mx;
  int32_t mn0, mx0;    // thresholds of group 0 (fast path when n_groups == 1)
  uint64_t G;
  int xb;              // 1: partitioned records carry the one-base extension in bit 15 (lengths below 32 768)
  int dbg;             // profiling hook (option "cover.dbg"): 1 / 2 / 3 cut the word-count tile kernel short after a phase; results are then WRONG
};

struct Pieces {
  uint32_t* pstart; uint32_t* pend; int32_t* acc; int32_t* cnt;
  uint32_t* smin; uint32_t* emax; uint32_t* smax; uint32_t* emin; uint8_t* flags;   // bit0: continues from the previous tile, bit1: reaches the tile end
  uint32_t* blk_cnt;   // [gridDim.x] pieces written by every block: block b owns the slots [b * seg, (b + 1) * seg)
  int64_t seg;         // so that the pieces come out in genome order without sorting (block ranges and tiles ascend)
  unsigned long long* ticks;   // profiling hook (cover.dbg = 9): clock cycles per phase of the word-count kernel, summed over the tiles (thread 0 of every block)
  int* fail;   // 1: output capacity, 2: distinct positions, 3: pieces, 4: halo regions of one tile
};

__device__ __forceinline__ bool in_range(const Cfg& c, int32_t d, uint64_t pos) {
  if (c.n_groups == 1) return d >= c.mn0 && d <= c.mx0;
  int g = (int)((pos / c.G) / (uint64_t)c.n_classes);
  return d >= c.mn[g] && d <= c.mx[g];
}

// one packed 64-bit record per region: linear start key (high word) and length payload (low word; bit 31: the
// extractor's one-base extension, GenometricCover.scala:353-354).  Partitioning sorts on the high bits only.
static __global__ void k_tile_keys(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ cls,
                                   int64_t n, LinMap lm, int64_t B, uint64_t* __restrict__ packed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * lm.G + lm.coff[chrom[i]];
    bool ext = (e % B == 0 && s / B != e / B);
    packed[i] = ((base + (uint64_t)s) << 32) | (uint64_t)((uint32_t)(e - s) | (ext ? EXT : 0u));
  }
}

// ---- partition of the regions by tile: counting + two multisplit passes ----------------------------------------------
// The order inside a tile is irrelevant, so no stable ranking is needed: a block counts its records per bin with
// shared-memory atomics (the returned value is the record's rank inside the block), reserves the output range of every
// bin with ONE global atomicAdd on a cursor that starts at the bin's final offset, regroups the records in shared memory
// and writes each bin's run coalesced.  Pass 1 splits by coarse bin (256 tiles) straight from the input columns (the
// packed record never exists unpartitioned); pass 2 splits every coarse bin into its tiles.
constexpr int MSP_THREADS = 512;
constexpr int MSP_ITEMS = 8;
constexpr int MSP_TILE = MSP_THREADS * MSP_ITEMS;
constexpr int MSP_BINS = 1024;          // block-local bins (pass 2: tiles reachable from the block's first coarse bin)
constexpr int HIST_CAP = 54 * 1024;     // tiles whose counters fit one block's shared memory

__device__ __forceinline__ uint64_t make_record(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits,
                                                const uint32_t* __restrict__ cls, const uint64_t* coff, uint64_t G, int64_t B, int64_t i) {
  int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
  uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * G + coff[chrom[i]];
  // stop on a bin multiple and the region crosses a bin edge; with stop = q*B the second condition (start / B != q)
  // is simply start < stop
  bool ext;
  if (bits == 32 && B > 0 && B < 0x7fffffffLL) ext = ((uint32_t)e % (uint32_t)B == 0) && s < e;
  else ext = (e % B == 0) && s < e;
  return ((base + (uint64_t)s) << 32) | (uint64_t)((uint32_t)(e - s) | (ext ? EXT : 0u));
}

// regions per tile: block-private counters in shared memory, merged with one RED per non-empty counter.  Block b counts
// the 4096-region input tiles b, b + G, b + 2G, ... (G = gridDim.x) — the same tiles the coarse multisplit pass hands to
// its blocks b, b + G, ... — and also reports its own count per COARSE bin (256 tiles): the coarse pass then reserves
// output space on a cursor per (coarse bin, group b) instead of one per coarse bin, which every block of the chip would
// hit once per tile (tens of thousands of same-address atomics with a returned value, serialised in the L2).
static __global__ void __launch_bounds__(1024, 1) k_tile_hist(const int32_t* __restrict__ chrom, const void* __restrict__ start, int bits, const uint32_t* __restrict__ cls,
                                                               int64_t n, LinMap lm, int logw, int nT, uint32_t* __restrict__ hist, uint32_t* __restrict__ coarse_cnt, int n_coarse,
                                                               int vec_ok) {
  extern __shared__ uint32_t sh_hist[];                    // [nT] then chromosome offsets, then the coarse counters
  uint64_t* s_coff = (uint64_t*)(sh_hist + ((nT + 1) & ~1));
  const bool chrom_sh = lm.n_chrom <= 1024;
  uint32_t* s_coarse = (uint32_t*)(s_coff + (chrom_sh ? lm.n_chrom : 0));
  for (int k = threadIdx.x; k < nT; k += blockDim.x) sh_hist[k] = 0;
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) s_coarse[k] = 0;
  if (chrom_sh) for (int k = threadIdx.x; k < lm.n_chrom; k += blockDim.x) s_coff[k] = lm.coff[k];
  __syncthreads();
  const uint64_t* coff = chrom_sh ? s_coff : lm.coff;
  const int64_t n_tiles = (n + MSP_TILE - 1) / MSP_TILE;
  const bool vec = vec_ok && bits == 32;
  for (int64_t t0 = blockIdx.x; t0 < n_tiles; t0 += 2 * (int64_t)gridDim.x) {
    int c[8]; uint32_t st[8], cl[8];                       // two tiles in flight: four consecutive regions per thread and tile
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = t0 + h * (int64_t)gridDim.x;
      const int64_t i0 = t * MSP_TILE + (int64_t)threadIdx.x * 4;
      if (t < n_tiles && vec && i0 + 4 <= n) {
        const int4 vc = *(const int4*)(chrom + i0), vs = *(const int4*)((const int32_t*)start + i0);
        uint4 vk = make_uint4(0u, 0u, 0u, 0u);
        if (cls) vk = *(const uint4*)(cls + i0);
        c[4 * h] = vc.x; c[4 * h + 1] = vc.y; c[4 * h + 2] = vc.z; c[4 * h + 3] = vc.w;
        st[4 * h] = (uint32_t)vs.x; st[4 * h + 1] = (uint32_t)vs.y; st[4 * h + 2] = (uint32_t)vs.z; st[4 * h + 3] = (uint32_t)vs.w;
        cl[4 * h] = vk.x; cl[4 * h + 1] = vk.y; cl[4 * h + 2] = vk.z; cl[4 * h + 3] = vk.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ok = t < n_tiles && i0 + q < n;
          c[4 * h + q] = ok ? chrom[i0 + q] : -1;
          st[4 * h + q] = ok ? (uint32_t)ld_coord(start, bits, i0 + q) : 0u;   // the tile path runs on 32-bit linear keys only
          cl[4 * h + q] = ok && cls ? cls[i0 + q] : 0u;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c[j] < 0) continue;
      const uint32_t key = (uint32_t)((uint64_t)cl[j] * lm.G + coff[c[j]]) + st[j];
      atomicAdd(&sh_hist[key >> logw], 1u);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nT; k += blockDim.x) {
    const uint32_t c = sh_hist[k];
    if (c) { atomicAdd(&hist[k], c); atomicAdd(&s_coarse[k >> 8], c); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) coarse_cnt[(size_t)k * gridDim.x + blockIdx.x] = s_coarse[k];
}

// the same counts from the ABI's narrow encodings (uint8 chromosome, int32 start): what gmql_dataset_upload keeps resident
static __global__ void __launch_bounds__(1024, 1) k_tile_hist_narrow(const uint8_t* __restrict__ chrom8, const int32_t* __restrict__ start, int64_t n, LinMap lm, int logw, int nT,
                                                                      uint32_t* __restrict__ hist, uint32_t* __restrict__ coarse_cnt, int n_coarse, int vec_ok) {
  extern __shared__ uint32_t sh_hist[];                    // [nT], chromosome offsets (<= 256), coarse counters
  uint32_t* s_coff = sh_hist + nT;
  uint32_t* s_coarse = s_coff + 256;
  for (int k = threadIdx.x; k < nT; k += blockDim.x) sh_hist[k] = 0;
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) s_coarse[k] = 0;
  for (int k = threadIdx.x; k < 256; k += blockDim.x) s_coff[k] = k < lm.n_chrom ? (uint32_t)lm.coff[k] : 0u;
  __syncthreads();
  const int64_t n_tiles = (n + MSP_TILE - 1) / MSP_TILE;
  for (int64_t t0 = blockIdx.x; t0 < n_tiles; t0 += 2 * (int64_t)gridDim.x) {
    uint32_t c[8], st[8]; bool ok[8];                      // two input tiles in flight: four consecutive regions per thread and tile
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = t0 + h * (int64_t)gridDim.x;
      const int64_t i0 = t * MSP_TILE + (int64_t)threadIdx.x * 4;
      if (t < n_tiles && vec_ok && i0 + 4 <= n) {
        const uint32_t vc = *(const uint32_t*)(chrom8 + i0);
        const int4 vs = *(const int4*)(start + i0);
        c[4 * h] = vc & 0xffu; c[4 * h + 1] = (vc >> 8) & 0xffu; c[4 * h + 2] = (vc >> 16) & 0xffu; c[4 * h + 3] = vc >> 24;
        st[4 * h] = (uint32_t)vs.x; st[4 * h + 1] = (uint32_t)vs.y; st[4 * h + 2] = (uint32_t)vs.z; st[4 * h + 3] = (uint32_t)vs.w;
        ok[4 * h] = ok[4 * h + 1] = ok[4 * h + 2] = ok[4 * h + 3] = true;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool v = t < n_tiles && i0 + q < n;
          ok[4 * h + q] = v;
          c[4 * h + q] = v ? (uint32_t)chrom8[i0 + q] : 0u;
          st[4 * h + q] = v ? (uint32_t)start[i0 + q] : 0u;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (ok[j]) atomicAdd(&sh_hist[(s_coff[c[j]] + st[j]) >> logw], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nT; k += blockDim.x) {
    const uint32_t cnt = sh_hist[k];
    if (cnt) { atomicAdd(&hist[k], cnt); atomicAdd(&s_coarse[k >> 8], cnt); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) coarse_cnt[(size_t)k * gridDim.x + blockIdx.x] = s_coarse[k];
}

// ---- speculative launch (caller knows the chromosome spans up front: chrom_span_hint) ----------------------------------
// With the layout of the linear coordinate known before any data is read, gmql_cover does not wait for its prepass: ONE
// kernel reads the columns once and produces both what the prepass reports (validity / zero-length flags, longest region,
// observed spans — and "beyond its hint" as a validity error) and the tile histogram, and the rest of the tile pipeline is
// queued behind it at once.  The kernels that would misbehave on an ineligible input (zero-length regions, a region as
// long as a tile, invalid rows) look at the flags first and return; the host reads the flags with the pipeline's single
// synchronisation and, in that rare case, runs the operator again the ordinary way.
struct Guard { const int* flags; int max_len; };   // flags: [1] zero-length seen, [2] invalid row, [3] longest region; null = no guard
__device__ __forceinline__ bool guard_trips(const Guard& g) {
  return g.flags != nullptr && (g.flags[1] != 0 || g.flags[2] != 0 || g.flags[3] >= g.max_len);
}

static __global__ void __launch_bounds__(1024, 1) k_prepass_hist(const int32_t* __restrict__ chrom, const int32_t* __restrict__ start, const int32_t* __restrict__ stop,
                                                                  const int32_t* __restrict__ sample, int64_t n, int n_samples, LinMap lm, int logw, int nT,
                                                                  uint32_t* __restrict__ hist, uint32_t* __restrict__ coarse_cnt, int n_coarse, int vec_ok,
                                                                  unsigned long long* __restrict__ act_span, int* __restrict__ flags) {
  extern __shared__ uint32_t sh_hist[];                    // [nT], chromosome offsets, coarse counters, hint spans, observed spans
  uint64_t* s_coff = (uint64_t*)(sh_hist + ((nT + 1) & ~1));
  uint32_t* s_coarse = (uint32_t*)(s_coff + lm.n_chrom);
  uint32_t* s_hint = s_coarse + n_coarse;                  // [n_chrom] upper bound of stop (32-bit coordinates)
  uint32_t* s_seen = s_hint + lm.n_chrom;                  // [n_chrom] largest stop seen
  __shared__ int s_fl[4];
  for (int k = threadIdx.x; k < nT; k += blockDim.x) sh_hist[k] = 0;
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) s_coarse[k] = 0;
  for (int k = threadIdx.x; k < lm.n_chrom; k += blockDim.x) {
    s_coff[k] = lm.coff[k]; s_seen[k] = 0;
    const int64_t h = lm.span[k];
    s_hint[k] = h > 0x7fffffffLL ? 0x7fffffffu : (uint32_t)h;
  }
  if (threadIdx.x < 4) s_fl[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n_tiles = (n + MSP_TILE - 1) / MSP_TILE;
  int zero = 0, bad = 0, maxlen = 0;
  for (int64_t t0 = blockIdx.x; t0 < n_tiles; t0 += 2 * (int64_t)gridDim.x) {
    int c[8], st[8], en[8], sm[8];                         // two input tiles in flight, four consecutive regions per thread and tile
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = t0 + h * (int64_t)gridDim.x;
      const int64_t i0 = t * MSP_TILE + (int64_t)threadIdx.x * 4;
      if (t < n_tiles && vec_ok && i0 + 4 <= n) {
        const int4 vc = *(const int4*)(chrom + i0), vs = *(const int4*)(start + i0), ve = *(const int4*)(stop + i0);
        int4 vb = make_int4(0, 0, 0, 0);
        if (sample) vb = *(const int4*)(sample + i0);
        c[4 * h] = vc.x; c[4 * h + 1] = vc.y; c[4 * h + 2] = vc.z; c[4 * h + 3] = vc.w;
        st[4 * h] = vs.x; st[4 * h + 1] = vs.y; st[4 * h + 2] = vs.z; st[4 * h + 3] = vs.w;
        en[4 * h] = ve.x; en[4 * h + 1] = ve.y; en[4 * h + 2] = ve.z; en[4 * h + 3] = ve.w;
        sm[4 * h] = vb.x; sm[4 * h + 1] = vb.y; sm[4 * h + 2] = vb.z; sm[4 * h + 3] = vb.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ok = t < n_tiles && i0 + q < n;
          c[4 * h + q] = ok ? chrom[i0 + q] : -2;          // -2: padding lane
          st[4 * h + q] = ok ? start[i0 + q] : 0;
          en[4 * h + q] = ok ? stop[i0 + q] : 1;
          sm[4 * h + q] = ok && sample ? sample[i0 + q] : 0;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c[j] == -2) continue;
      if (st[j] == en[j]) zero = 1;
      if (c[j] < 0 || c[j] >= lm.n_chrom || st[j] < 0 || st[j] > en[j] || sm[j] < 0 || sm[j] >= n_samples || (uint32_t)en[j] > s_hint[c[j]]) { bad = 1; continue; }
      maxlen = max(maxlen, en[j] - st[j]);
      atomicMax(&s_seen[c[j]], (uint32_t)en[j]);
      const uint32_t key = (uint32_t)s_coff[c[j]] + (uint32_t)st[j];
      atomicAdd(&sh_hist[key >> logw], 1u);
    }
  }
  {
    maxlen = __reduce_max_sync(0xffffffffu, maxlen);
    const bool a1 = __any_sync(0xffffffffu, zero), a2 = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      if (a1) s_fl[1] = 1;
      if (a2) s_fl[2] = 1;
      if (maxlen > 0) atomicMax(&s_fl[3], maxlen);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                  // one store / RED per block and flag (see k_cover_prepass)
    flags[0] = 1;                                          // no strand column: every region counts as '*'
    if (s_fl[1]) flags[1] = 1;
    if (s_fl[2]) flags[2] = 1;
    if (s_fl[3] > 0) atomicMax(&flags[3], s_fl[3]);
  }
  for (int k = threadIdx.x; k < lm.n_chrom; k += blockDim.x) { const uint32_t v = s_seen[k]; if (v) atomicMax(&act_span[k], (unsigned long long)v); }
  for (int k = threadIdx.x; k < nT; k += blockDim.x) {
    const uint32_t cnt = sh_hist[k];
    if (cnt) { atomicAdd(&hist[k], cnt); atomicAdd(&s_coarse[k >> 8], cnt); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_coarse; k += blockDim.x) coarse_cnt[(size_t)k * gridDim.x + blockIdx.x] = s_coarse[k];
}

// cursors: every tile starts at its final offset (the coarse cursors are the exclusive scan of k_tile_hist's coarse counts)
static __global__ void k_init_cursors(const uint32_t* __restrict__ off, int nT, uint32_t* __restrict__ cur_tile) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nT; k += gridDim.x * blockDim.x) cur_tile[k] = off[k];
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) completing on a shared-memory mbarrier -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy reads of a staging buffer are ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Three variants of one kernel:
//   MSP_WIDE / MSP_NARROW  coarse pass: records are built from the input columns — int32 (or int64) chrom / start / stop, or the
//                          ABI's narrow encodings uint8 chrom, int32 start, uint16 length (7 bytes per region) — and split by
//                          coarse bin (tile >> 8); written as two arrays, linear start key (uint32) and length (uint16),
//                          6 bytes per region;
//   MSP_FINE               fine pass: reads those two arrays and splits every coarse bin into its tiles; written as ONE 4-byte
//                          record per region (tile-local start << 16 | length), which is all the tile kernel reads.
// bin_shift = 32 + logw (+ 8 for the coarse pass) applies to the 64-bit register form (key << 32 | length).
// The input of the NEXT tile of a block is fetched by TMA bulk copies into a staging buffer while the current tile is ranked,
// regrouped and written: every thread moves its items from the staging buffer into registers first, so one staging buffer
// and one mbarrier suffice, and the load latency no longer sits at the head of every tile (registers are at the
// 64-per-thread limit of two 512-thread blocks per SM, so a register prefetch is not an option).
enum { MSP_FINE = 0, MSP_WIDE = 1, MSP_NARROW = 2 };
struct MspIn {
  const uint32_t* key; const uint16_t* len;                                                         // MSP_FINE
  const int32_t* chrom; const void* start; const void* stop; int bits; const uint32_t* cls;       // MSP_WIDE
  const uint8_t* chrom8; const uint16_t* len16;                                                     // MSP_NARROW (start: int32 in `start`)
};
struct MspOut { uint32_t* key; uint16_t* len; uint32_t* rec; };
template <int MODE>
__host__ __device__ constexpr size_t msp_stage_bytes() { return (MODE == MSP_FINE ? 6 : (MODE == MSP_WIDE ? 12 : 7)) * (size_t)MSP_TILE; }
template <int MODE>
constexpr size_t msp_smem_bytes() { return 8 * (size_t)MSP_TILE + msp_stage_bytes<MODE>() + 3 * 4 * (size_t)MSP_BINS + (MODE != MSP_FINE ? 8 * 256 : 0); }

template <int MODE>
static __global__ void __launch_bounds__(MSP_THREADS, MODE == 0 ? 3 : 2) k_multisplit(MspIn in, LinMap lm, int64_t B, int64_t n, int bin_shift, int logw, uint32_t* __restrict__ cursor, MspOut out,
                                                                      int vec_ok, int groups, Guard guard, int xb, BinMod32 bm32) {
  if (guard_trips(guard)) return;
  constexpr bool FUSED = MODE != MSP_FINE;
  extern __shared__ __align__(128) unsigned char msp_smem[];
  uint64_t* sdata = (uint64_t*)msp_smem;                                    // [MSP_TILE] regroup buffer
  unsigned char* stage = msp_smem + 8 * (size_t)MSP_TILE;                   // TMA landing buffer
  uint32_t* cnt = (uint32_t*)(stage + msp_stage_bytes<MODE>());
  uint32_t* lexcl = cnt + MSP_BINS;
  uint32_t* gbase = lexcl + MSP_BINS;
  uint64_t* s_coff = (uint64_t*)(gbase + MSP_BINS);                         // [256], coarse passes only
  __shared__ uint32_t wtot[MSP_THREADS / 32];
  __shared__ uint32_t s_base_bin;
  __shared__ uint64_t s_bar;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool chrom_sh = FUSED && lm.n_chrom <= 256;
  if (chrom_sh) for (int k = tid; k < lm.n_chrom; k += MSP_THREADS) s_coff[k] = lm.coff[k];
  const uint64_t* coff = chrom_sh ? s_coff : lm.coff;
  const int64_t n_tiles = (n + MSP_TILE - 1) / MSP_TILE;
  // full tiles of suitably aligned input go through the staging buffer
  const bool tma_ok = MODE == MSP_FINE ? ((((uintptr_t)in.key | (uintptr_t)in.len) & 15u) == 0)
                    : (MODE == MSP_WIDE ? (vec_ok && in.bits == 32 && in.cls == nullptr) : (vec_ok != 0));
  auto issue = [&](int64_t t) {                            // one thread: arm the barrier, start the copies of tile t
    const int64_t base = t * MSP_TILE;
    mbar_expect_tx(&s_bar, (uint32_t)msp_stage_bytes<MODE>());
    if (MODE == MSP_WIDE) {
      tma_load_1d(stage, in.chrom + base, 4 * MSP_TILE, &s_bar);
      tma_load_1d(stage + 4 * MSP_TILE, (const int32_t*)in.start + base, 4 * MSP_TILE, &s_bar);
      tma_load_1d(stage + 8 * MSP_TILE, (const int32_t*)in.stop + base, 4 * MSP_TILE, &s_bar);
    } else if (MODE == MSP_NARROW) {
      tma_load_1d(stage, (const int32_t*)in.start + base, 4 * MSP_TILE, &s_bar);
      tma_load_1d(stage + 4 * MSP_TILE, in.len16 + base, 2 * MSP_TILE, &s_bar);
      tma_load_1d(stage + 6 * MSP_TILE, in.chrom8 + base, MSP_TILE, &s_bar);
    } else {
      tma_load_1d(stage, in.key + base, 4 * MSP_TILE, &s_bar);
      tma_load_1d(stage + 4 * MSP_TILE, in.len + base, 2 * MSP_TILE, &s_bar);
    }
  };
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    if (tma_ok && (int64_t)blockIdx.x < n_tiles && n - (int64_t)blockIdx.x * MSP_TILE >= MSP_TILE) issue(blockIdx.x);
  }
  uint32_t parity = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = t * MSP_TILE;
    const int count = (int)((n - base) < (int64_t)MSP_TILE ? (n - base) : (int64_t)MSP_TILE);
    const bool staged = tma_ok && count == MSP_TILE;
    // coarse pass: the cursor of (bin, group of this input tile), see k_tile_hist; fine pass: one cursor per tile
    const uint32_t cmul = FUSED ? (uint32_t)groups : 1u, cadd = FUSED ? (uint32_t)(t % groups) : 0u;
    for (int k = tid; k < MSP_BINS; k += MSP_THREADS) cnt[k] = 0;
    __syncthreads();                                       // also orders s_coff, the barrier's init and the previous iteration's reads
    // item j of a thread is record idx_of(j) of the tile: striped for the fine pass, four consecutive regions per
    // vector load when the records are built from the columns
    uint64_t rec[MSP_ITEMS];
    if (staged) { mbar_wait(&s_bar, parity); parity ^= 1u; }
    if (MODE == MSP_WIDE) {
#pragma unroll
      for (int j4 = 0; j4 < MSP_ITEMS / 4; ++j4) {
        const int idx0 = (j4 * MSP_THREADS + tid) * 4;
        const int64_t i0 = base + idx0;
        if (staged || (idx0 + 4 <= count && vec_ok && in.bits == 32)) {
          int4 vc, vs, ve;
          uint4 vk = make_uint4(0u, 0u, 0u, 0u);
          if (staged) {
            vc = ((const int4*)stage)[j4 * MSP_THREADS + tid];
            vs = ((const int4*)(stage + 4 * MSP_TILE))[j4 * MSP_THREADS + tid];
            ve = ((const int4*)(stage + 8 * MSP_TILE))[j4 * MSP_THREADS + tid];
          } else {
            vc = *(const int4*)(in.chrom + i0); vs = *(const int4*)((const int32_t*)in.start + i0); ve = *(const int4*)((const int32_t*)in.stop + i0);
            if (in.cls) vk = *(const uint4*)(in.cls + i0);
          }
          const int cc[4] = {vc.x, vc.y, vc.z, vc.w}, ss[4] = {vs.x, vs.y, vs.z, vs.w}, ee[4] = {ve.x, ve.y, ve.z, ve.w};
          const uint32_t kk[4] = {vk.x, vk.y, vk.z, vk.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t key = (uint32_t)((uint64_t)kk[q] * lm.G + coff[cc[q]]) + (uint32_t)ss[q];
            uint32_t lq = (uint32_t)(ee[q] - ss[q]) & 0xffffu;
            if (xb && lq != 0u && is_bin_multiple32(bm32, (uint32_t)ee[q])) lq |= 0x8000u;   // the extractor's one-base extension (GenometricCover.scala:353-354)
            rec[j4 * 4 + q] = ((uint64_t)key << 32) | (uint64_t)lq;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint64_t r = 0ull;
            if (idx0 + q < count) {
              r = make_record(in.chrom, in.start, in.stop, in.bits, in.cls, coff, lm.G, B, i0 + q);
              r = xb ? ((r & 0xffffffff00007fffull) | ((r >> 16) & 0x8000ull)) : (r & 0xffffffff0000ffffull);   // key << 32 | length (| extension in bit 15)
            }
            rec[j4 * 4 + q] = r;
          }
        }
      }
    } else if (MODE == MSP_NARROW) {
#pragma unroll
      for (int j4 = 0; j4 < MSP_ITEMS / 4; ++j4) {
        const int idx0 = (j4 * MSP_THREADS + tid) * 4;
        const int64_t i0 = base + idx0;
        int ss[4]; uint32_t ll[4], cc[4];
        if (staged || (idx0 + 4 <= count && vec_ok)) {
          int4 vs; uint2 vl; uint32_t vc;
          if (staged) {
            vs = ((const int4*)stage)[j4 * MSP_THREADS + tid];
            vl = ((const uint2*)(stage + 4 * MSP_TILE))[j4 * MSP_THREADS + tid];
            vc = ((const uint32_t*)(stage + 6 * MSP_TILE))[j4 * MSP_THREADS + tid];
          } else {
            vs = *(const int4*)((const int32_t*)in.start + i0); vl = *(const uint2*)(in.len16 + i0); vc = *(const uint32_t*)(in.chrom8 + i0);
          }
          ss[0] = vs.x; ss[1] = vs.y; ss[2] = vs.z; ss[3] = vs.w;
          ll[0] = vl.x & 0xffffu; ll[1] = vl.x >> 16; ll[2] = vl.y & 0xffffu; ll[3] = vl.y >> 16;
          cc[0] = vc & 0xffu; cc[1] = (vc >> 8) & 0xffu; cc[2] = (vc >> 16) & 0xffu; cc[3] = vc >> 24;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ok = idx0 + q < count;
            ss[q] = ok ? ((const int32_t*)in.start)[i0 + q] : 0;
            ll[q] = ok ? (uint32_t)in.len16[i0 + q] : 0u;
            cc[q] = ok ? (uint32_t)in.chrom8[i0 + q] : 0u;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t key = (uint32_t)coff[cc[q]] + (uint32_t)ss[q];
          uint32_t lq = ll[q];
          if (xb && lq != 0u && is_bin_multiple32(bm32, (uint32_t)ss[q] + lq)) lq |= 0x8000u;   // the extractor's one-base extension (GenometricCover.scala:353-354)
          rec[j4 * 4 + q] = idx0 + q < count ? (((uint64_t)key << 32) | (uint64_t)lq) : 0ull;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < MSP_ITEMS; ++j) {
        const int idx = j * MSP_THREADS + tid;
        uint32_t k32 = 0u, l16 = 0u;
        if (staged) { k32 = ((const uint32_t*)stage)[idx]; l16 = ((const uint16_t*)(stage + 4 * MSP_TILE))[idx]; }
        else if (idx < count) { k32 = in.key[base + idx]; l16 = in.len[base + idx]; }
        rec[j] = ((uint64_t)k32 << 32) | (uint64_t)l16;
      }
    }
    if (!FUSED && tid == 0) s_base_bin = (uint32_t)(rec[0] >> bin_shift) & ~255u;   // local bins are counted from the coarse bin of the tile's first record
    if (!FUSED || tma_ok) __syncthreads();                 // every thread holds its items: the staging buffer is free again
    if (tma_ok && tid == 0) {
      const int64_t tn = t + gridDim.x;
      if (tn < n_tiles && n - tn * MSP_TILE >= MSP_TILE) { fence_proxy_async(); issue(tn); }
    }
    const uint32_t base_bin = FUSED ? 0u : s_base_bin;
    const uint32_t wmask = (1u << logw) - 1u;
    auto emit = [&](uint32_t g, uint64_t r) {              // one record at its final place
      if (FUSED) { out.key[g] = (uint32_t)(r >> 32); out.len[g] = (uint16_t)r; }
      else out.rec[g] = ((((uint32_t)(r >> 32)) & wmask) << 16) | ((uint32_t)r & 0xffffu);
    };
    uint32_t rk[MSP_ITEMS];
#pragma unroll
    for (int j = 0; j < MSP_ITEMS; ++j) {
      const int idx = FUSED ? ((j >> 2) * MSP_THREADS + tid) * 4 + (j & 3) : j * MSP_THREADS + tid;
      rk[j] = 0xffffffffu;
      if (idx < count) {
        uint32_t lb = (uint32_t)(rec[j] >> bin_shift) - base_bin;
        if (lb < (uint32_t)MSP_BINS) rk[j] = atomicAdd(&cnt[lb], 1u);
        else {                                             // far bin (tiny coarse bins only): placed directly
          uint32_t g = atomicAdd(&cursor[(uint32_t)(rec[j] >> bin_shift) * cmul + cadd], 1u);
          emit(g, rec[j]);
        }
      }
    }
    __syncthreads();
    // reserve every bin's output range; exclusive offsets of the bins inside the block
    uint32_t g0 = 0, g1 = 0;
    {
      uint32_t c0 = cnt[2 * tid], c1 = cnt[2 * tid + 1];
      // the returned offsets are only needed by the final write: they stay in registers until after the regroup, so the
      // atomics' round trip to the L2 overlaps the scan, the barriers and the regroup
      if (c0) g0 = atomicAdd(&cursor[(base_bin + 2 * tid) * cmul + cadd], c0);
      if (c1) g1 = atomicAdd(&cursor[(base_bin + 2 * tid + 1) * cmul + cadd], c1);
      uint32_t v = c0 + c1, x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane == 31) wtot[w] = x;
      __syncthreads();
      uint32_t tt = lane < MSP_THREADS / 32 ? wtot[lane] : 0;
#pragma unroll
      for (int o = 1; o < MSP_THREADS / 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, tt, o); if (lane >= o) tt += y; }
      uint32_t below = __shfl_sync(0xffffffffu, tt, w > 0 ? w - 1 : 0);
      uint32_t ex = (w > 0 ? below : 0) + x - v;
      lexcl[2 * tid] = ex; lexcl[2 * tid + 1] = ex + c0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MSP_ITEMS; ++j) {
      if (rk[j] != 0xffffffffu) {
        uint32_t lb = (uint32_t)(rec[j] >> bin_shift) - base_bin;
        sdata[lexcl[lb] + rk[j]] = rec[j];
      }
    }
    gbase[2 * tid] = g0; gbase[2 * tid + 1] = g1;
    __syncthreads();
    const int local_total = (int)(lexcl[MSP_BINS - 1] + cnt[MSP_BINS - 1]);
    for (int i = tid; i < local_total; i += MSP_THREADS) {
      uint64_t r = sdata[i];
      uint32_t lb = (uint32_t)(r >> bin_shift) - base_bin;
      emit(gbase[lb] + ((uint32_t)i - lexcl[lb]), r);
    }
    __syncthreads();
  }
}

// the radix-sort route (more tiles than the counting kernel's shared memory holds): sorted 64-bit records -> 4-byte records
static __global__ void k_pack_rec32(const uint64_t* __restrict__ packed, int64_t n, int logw, uint32_t* __restrict__ rec, int xb) {
  const uint32_t wmask = (1u << logw) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t r = packed[i];
    rec[i] = ((((uint32_t)(r >> 32)) & wmask) << 16) | (xb ? (((uint32_t)r & 0x7fffu) | (((uint32_t)r >> 16) & 0x8000u)) : ((uint32_t)r & 0xffffu));
  }
}

// records partitioned by tile -> first index of every tile (off[nT] = n)
static __global__ void k_tile_offsets(const uint64_t* __restrict__ packed, int64_t n, int64_t nT, int logw, uint32_t* __restrict__ off) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = (int64_t)(packed[i] >> (32 + logw));
    int64_t tp = i > 0 ? (int64_t)(packed[i - 1] >> (32 + logw)) : -1;
    for (int64_t u = tp + 1; u <= t; ++u) off[u] = (uint32_t)i;
    if (i == n - 1) for (int64_t u = t + 1; u <= nT; ++u) off[u] = (uint32_t)n;
  }
}

// One barrier: the caller hands every call site of a tile its own scratch array, so a scratch array is only rewritten
// after the block has passed other barriers.
__device__ __forceinline__ int32_t block_excl_scan(int32_t v, int32_t* total, int32_t* scratch /*[NWARPS]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) scratch[w] = x;
  __syncthreads();
  int32_t t = lane < NWARPS ? scratch[lane] : 0;           // every warp scans the NWARPS warp totals itself
#pragma unroll
  for (int o = 1; o < NWARPS; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
  const int32_t tot = __shfl_sync(0xffffffffu, t, NWARPS - 1);
  const int32_t below = __shfl_sync(0xffffffffu, t, w > 0 ? w - 1 : 0);
  *total = tot;
  return (w > 0 ? below : 0) + x - v;
}

template <bool MULTI>
__device__ __forceinline__ bool in_rng(const Cfg& c, int32_t mn0, int32_t mx0, int32_t d, uint64_t pos) {
  if (!MULTI) return d >= mn0 && d <= mx0;
  int g = (int)((pos / c.G) / (uint64_t)c.n_classes);
  return d >= c.mn[g] && d <= c.mx[g];
}

// Positions of the tile's distinct events in rank order, read off the occupancy bitmap (no per-rank position array):
// seek(k) finds the word holding rank k through the popcount prefix, next() walks the set bits upward.
struct RankPos {
  const uint32_t* bitmap; int w; uint32_t word;
  __device__ __forceinline__ void seek(const uint32_t* bm, const uint16_t* wpre, int nW, int k) {
    bitmap = bm;
    int lo = 0, hi = nW;                                   // last word whose prefix is <= k: it holds rank k
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((int)wpre[mid] <= k) lo = mid; else hi = mid; }
    w = lo;
    word = bm[lo];
    const int r = k - (int)wpre[lo];                       // drop the r lower set bits of the word
    if (r > 0) word &= ~((2u << __fns(word, 0, r)) - 1u);
  }
  __device__ __forceinline__ uint32_t next() {             // tile-local position of the current rank; advances
    while (word == 0) { ++w; word = bitmap[w]; }
    const uint32_t pos = ((uint32_t)w << 5) + (uint32_t)__ffs(word) - 1u;
    word &= word - 1u;
    return pos;
  }
};

constexpr int RC = 8;   // regions of the tile each thread keeps in registers (the rest are re-read from L1/L2)

// GMAP4 contribution of one region per thread (block-uniform call): find the pieces it overlaps, accumulate with
// warp-aggregated shared-memory atomics
__device__ __forceinline__ void gmap4_round(bool valid, uint32_t rs, uint32_t re, bool own, bool cont0, int np, const uint32_t* pst, const uint32_t* pen,
                                            int32_t* pcnt, uint32_t* psmin, uint32_t* pemax, uint32_t* psmax, uint32_t* pemin) {
  const int lane = threadIdx.x & 31;
  int j = np;
  if (valid) {
    int lo = 0, hi = np;                 // first piece with end > rs
    while (lo < hi) { int mid = (lo + hi) >> 1; if (pen[mid] > rs) hi = mid; else lo = mid + 1; }
    j = lo;
  }
  if (np > 6) {
    // many pieces (sparse coverage, thresholds near the typical depth): the lanes of a warp rarely meet on a piece, and the
    // warp-aggregated loop below would run once per distinct piece — every lane accumulates on its own instead
    if (valid) {
      for (; j < np && pst[j] < re; ++j) {
        if (own || !(j == 0 && cont0)) {   // halo regions only feed runs that START in this tile
          atomicAdd(&pcnt[j], 1);
          atomicMin(&psmin[j], rs); atomicMax(&pemax[j], re);
          atomicMax(&psmax[j], rs); atomicMin(&pemin[j], re);
        }
      }
    }
    return;
  }
  while (true) {
    bool ov = valid && j < np && pst[j] < re;
    unsigned bal = __ballot_sync(0xffffffffu, ov);
    if (!bal) break;
    int J = __shfl_sync(0xffffffffu, j, __ffs(bal) - 1);
    bool mine = ov && j == J;
    bool add = mine && (own || !(J == 0 && cont0));   // halo regions only feed runs that START in this tile
    unsigned m = __ballot_sync(0xffffffffu, add);
    if (m) {
      uint32_t a_smin = __reduce_min_sync(0xffffffffu, add ? rs : 0xffffffffu);
      uint32_t a_emax = __reduce_max_sync(0xffffffffu, add ? re : 0u);
      uint32_t a_smax = __reduce_max_sync(0xffffffffu, add ? rs : 0u);
      uint32_t a_emin = __reduce_min_sync(0xffffffffu, add ? re : 0xffffffffu);
      if (lane == (__ffs(m) - 1)) {
        atomicAdd(&pcnt[J], __popc(m));
        atomicMin(&psmin[J], a_smin); atomicMax(&pemax[J], a_emax);
        atomicMax(&psmax[J], a_smax); atomicMin(&pemin[J], a_emin);
      }
    }
    if (mine) ++j;
  }
}

// Every block owns a CONTIGUOUS range of tiles holding an equal share of the regions.  Inside its range the regions
// still open at a tile's end are handed to the next tile through shared memory; only the first tile of a range scans
// its predecessor's regions in global memory.
template <bool MULTI>
static __global__ void __launch_bounds__(THREADS, CTAS_PER_SM) k_cover_tile(const uint32_t* __restrict__ P /* tile-local start << 16 | length */, const uint32_t* __restrict__ offS,
                                                                  int64_t t_lo, int64_t t_hi, int logw, Cfg cfg, LinMap lm, int64_t B, Pieces out, Guard guard) {
  if (guard_trips(guard)) return;
  const BinMod bm = make_binmod(B);
  const bool xb = cfg.xb != 0;
  const uint32_t lmask = xb ? 0x7fffu : 0xffffu;
  extern __shared__ uint32_t sm[];
  uint32_t* bitmap = sm;                                   // [MAXWORDS] occupancy of the tile's positions
  uint16_t* wpre = (uint16_t*)(bitmap + MAXWORDS);         // [MAXWORDS] set bits before each word
  uint32_t* delta = (uint32_t*)(wpre + MAXWORDS);          // [CAP_EV / 2] net +1/-1 per distinct position (by rank), 16 bits each
  uint32_t* hinS = delta + CAP_EV / 2;                     // [HCAP] regions still open at the tile start: start, length payload
  uint32_t* hinL = hinS + HCAP;
  uint32_t* houtS = hinL + HCAP;                           // [HCAP] regions still open at the tile end (next tile's hin)
  uint32_t* houtL = houtS + HCAP;
  uint32_t* pst = houtL + HCAP;                            // [MAXP] x 8: the tile's pieces and their GMAP4 accumulators
  uint32_t* pen = pst + MAXP;
  int32_t* pacc = (int32_t*)(pen + MAXP);
  int32_t* pcnt = pacc + MAXP;
  uint32_t* psmin = (uint32_t*)(pcnt + MAXP);
  uint32_t* pemax = psmin + MAXP;
  uint32_t* psmax = pemax + MAXP;
  uint32_t* pemin = psmax + MAXP;
  __shared__ int32_t scratch[3][NWARPS];                   // one per scan of a tile (see block_excl_scan)
  __shared__ int s_nh, s_nout;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t W = 1u << logw;
  const int nW = (int)(W >> 5) > 0 ? (int)(W >> 5) : 1;
  const int WPT = (nW + THREADS - 1) / THREADS;            // bitmap words per thread (consecutive)
  const int32_t mn0 = cfg.mn0, mx0 = cfg.mx0;

  // this block's tile range inside [t_lo, t_hi) (the whole genome, or the tiles one shard owns): first tile whose region
  // offset reaches the block's share of the regions that start there
  int64_t t_begin, t_end;
  {
    const uint32_t o_lo = offS[t_lo], o_hi = offS[t_hi];
    int64_t bounds[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t c = (int64_t)blockIdx.x + q;
      if (c <= 0) bounds[q] = t_lo;
      else if (c >= (int64_t)gridDim.x) bounds[q] = t_hi;
      else {
        const uint32_t target = o_lo + (uint32_t)(((int64_t)(o_hi - o_lo) * c) / (int64_t)gridDim.x);
        int64_t lo = t_lo, hi = t_hi;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (offS[mid] < target) lo = mid + 1; else hi = mid; }
        bounds[q] = lo;
      }
    }
    t_begin = bounds[0]; t_end = bounds[1];
  }
  int nh_carry = -1;                                       // -1: the open regions must be found by scanning the previous tile
  int64_t my_cnt = 0;                                      // pieces this block has published (block-uniform)

  // Invariants at the top of every tile iteration (established here, restored by the tile that breaks them, so that no
  // barrier is spent on clearing): the bitmap and the piece accumulators are clean, s_nout = 0 and s_nh = the number of
  // regions carried in hin (0 when the previous tile must be scanned).
  for (int k = tid; k < nW; k += THREADS) bitmap[k] = 0;
  for (int k = tid; k < MAXP; k += THREADS) { pacc[k] = 0; pcnt[k] = 0; psmin[k] = 0xffffffffu; pemax[k] = 0; psmax[k] = 0; pemin[k] = 0xffffffffu; pen[k] = 0; }
  if (tid == 0) { s_nh = 0; s_nout = 0; }
  __syncthreads();

  uint32_t s1_next = t_begin < t_end ? offS[t_begin] : 0u, s2_next = t_begin < t_end ? offS[t_begin + 1] : 0u;
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const uint32_t s0 = s1_next, s1 = s2_next;
    s1_next = s1;
    if (tile + 1 < t_end) s2_next = offS[tile + 2];        // the next tile's range, loaded a whole tile ahead
    const uint32_t nS = s1 - s0;
    const bool scan_prev = nh_carry < 0;
    const uint32_t h0 = (scan_prev && tile > 0) ? offS[tile - 1] : s0;
    const uint32_t tile_lo = (uint32_t)tile << logw;
    if (nS == 0 && (scan_prev ? h0 == s0 : nh_carry == 0)) { nh_carry = 0; continue; }   // depth 0 throughout, nothing left open
    const TileChrom tch = tile_chrom(lm, tile_lo);
    // the next tile's records: into L2 while this tile is processed
    if (tile + 1 < t_end) {
      const uintptr_t a0 = (uintptr_t)(P + s1) & ~(uintptr_t)127, a1 = (uintptr_t)(P + s2_next);
      for (uintptr_t a = a0 + ((uintptr_t)tid << 7); a < a1; a += (uintptr_t)THREADS << 7) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
    // ---- phase 1: mark event positions; regions of the previous tile that are still open; own regions into registers
    if (scan_prev) {
      for (uint32_t i0 = h0; i0 < s0; i0 += THREADS) {
        uint32_t i = i0 + tid;
        if (i < s0) {
          const uint32_t pr = P[i]; uint32_t rs = tile_lo - W + (pr >> 16), lp = rec_len_payload(pr, rs, tch, lm, bm, xb);
          uint32_t ee = rs + (lp & ~EXT) + (lp >> 31);
          if (ee >= tile_lo) {
            int slot = atomicAdd(&s_nh, 1);
            if (slot < HCAP) { hinS[slot] = rs; hinL[slot] = lp; }
            uint32_t q = ee - tile_lo;
            atomicOr(&bitmap[q >> 5], 1u << (q & 31));
          }
        }
      }
    } else {
      for (int k = tid; k < nh_carry; k += THREADS) {
        uint32_t lp = hinL[k];
        uint32_t q = hinS[k] + (lp & ~EXT) + (lp >> 31) - tile_lo;
        atomicOr(&bitmap[q >> 5], 1u << (q & 31));
      }
    }
    uint32_t rsr[RC], lpr[RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      uint32_t i = s0 + tid + r * THREADS;
      rsr[r] = 0; lpr[r] = 0;
      if (i < s1) { const uint32_t pr = P[i]; rsr[r] = tile_lo + (pr >> 16); lpr[r] = rec_len_payload(pr, rsr[r], tch, lm, bm, xb); }
    }
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      if (s0 + tid + r * THREADS < s1) {
        uint32_t q = rsr[r] - tile_lo;
        atomicOr(&bitmap[q >> 5], 1u << (q & 31));
        uint32_t qe = q + (lpr[r] & ~EXT) + (lpr[r] >> 31);
        if (qe < W) atomicOr(&bitmap[qe >> 5], 1u << (qe & 31));
        else { int slot = atomicAdd(&s_nout, 1); if (slot < HCAP) { houtS[slot] = rsr[r]; houtL[slot] = lpr[r]; } }
      }
    }
    for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
      const uint32_t pr = P[i]; uint32_t rs = tile_lo + (pr >> 16), lp = rec_len_payload(pr, rs, tch, lm, bm, xb);
      uint32_t q = rs - tile_lo;
      atomicOr(&bitmap[q >> 5], 1u << (q & 31));
      uint32_t qe = q + (lp & ~EXT) + (lp >> 31);
      if (qe < W) atomicOr(&bitmap[qe >> 5], 1u << (qe & 31));
      else { int slot = atomicAdd(&s_nout, 1); if (slot < HCAP) { houtS[slot] = rs; houtL[slot] = lp; } }
    }
    __syncthreads();
    const int nh = s_nh, nout = s_nout;
    const int32_t depth_in = nh;                           // regions still open when the tile begins
    // ---- phase 2: rank = popcount prefix over the bitmap
    int32_t total;
    {
      int32_t c = 0;
      const int w0 = tid * WPT;
      for (int q = 0; q < WPT; ++q) if (w0 + q < nW) c += __popc(bitmap[w0 + q]);
      int32_t pre = block_excl_scan(c, &total, scratch[0]);
      if (nh > HCAP || nout > HCAP || total > CAP_EV || nS > (uint32_t)MAX_TILE_REGIONS) {    // uniform across the block
        if (tid == 0) { *out.fail = (total > CAP_EV || nS > (uint32_t)MAX_TILE_REGIONS) ? 2 : 4; s_nh = 0; s_nout = 0; }
        nh_carry = -1;
        for (int k = tid; k < nW; k += THREADS) bitmap[k] = 0;   // every popcount above happened before the scan's barrier
        __syncthreads();
        continue;
      }
      if (tid == 0) { s_nh = nout; s_nout = 0; }           // every thread has read both (before the scan's barrier)
      for (int q = 0; q < WPT; ++q) if (w0 + q < nW) { wpre[w0 + q] = (uint16_t)pre; pre += __popc(bitmap[w0 + q]); }
    }
    for (int k = tid; k < ((total + 1) >> 1); k += THREADS) delta[k] = 0;
    __syncthreads();
    // ---- phase 3: sum +1 / -1 per distinct position
    for (int k = tid; k < nh; k += THREADS) {
      uint32_t lp = hinL[k];
      uint32_t q = hinS[k] + (lp & ~EXT) + (lp >> 31) - tile_lo;
      int r = wpre[q >> 5] + __popc(bitmap[q >> 5] & ((1u << (q & 31)) - 1u));
      delta_add(delta, r, -1);
    }
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      if (s0 + tid + r * THREADS < s1) {
        uint32_t q = rsr[r] - tile_lo;
        int rk = wpre[q >> 5] + __popc(bitmap[q >> 5] & ((1u << (q & 31)) - 1u));
        delta_add(delta, rk, 1);
        uint32_t qe = q + (lpr[r] & ~EXT) + (lpr[r] >> 31);
        if (qe < W) {
          int r2 = wpre[qe >> 5] + __popc(bitmap[qe >> 5] & ((1u << (qe & 31)) - 1u));
          delta_add(delta, r2, -1);
        }
      }
    }
    for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
      const uint32_t pr = P[i]; uint32_t rs = tile_lo + (pr >> 16), lp = rec_len_payload(pr, rs, tch, lm, bm, xb);
      uint32_t q = rs - tile_lo;
      int rk = wpre[q >> 5] + __popc(bitmap[q >> 5] & ((1u << (q & 31)) - 1u));
      delta_add(delta, rk, 1);
      uint32_t qe = q + (lp & ~EXT) + (lp >> 31);
      if (qe < W) {
        int r2 = wpre[qe >> 5] + __popc(bitmap[qe >> 5] & ((1u << (qe & 31)) - 1u));
        delta_add(delta, r2, -1);
      }
    }
    __syncthreads();
    // ---- phase 4: depth profile over the distinct positions, pieces
    const bool cont0 = in_rng<MULTI>(cfg, mn0, mx0, depth_in, tile_lo);
    const int IT = ((total + THREADS - 1) / THREADS) | 1;  // ranks per thread; odd stride: conflict-free shared-memory access
    const int k0 = tid * IT < total ? tid * IT : total;
    const int k1 = k0 + IT < total ? k0 + IT : total;
    // per-thread chunk: net change and the range of the running depth relative to the chunk start
    int32_t csum = 0, rmin = 0x7fffffff, rmax = -0x7fffffff;
    for (int k = k0; k < k1; ++k) { csum += delta_get(delta, k); rmin = min(rmin, csum); rmax = max(rmax, csum); }
    int32_t tile_net;
    const int32_t d_start = depth_in + block_excl_scan(csum, &tile_net, scratch[1]);
    // the state just before my first rank: in range iff the depth there is (the previous position belongs to the same
    // class unless the depth is 0, which is never in range)
    RankPos rp;                                            // positions are needed at transitions (and, with groups, everywhere)
    uint32_t pos_k0 = 0;
    if (MULTI && k0 < total) { rp.seek(bitmap, wpre, nW, k0); RankPos peek = rp; pos_k0 = peek.next(); }
    const bool in0 = k0 < total ? in_rng<MULTI>(cfg, mn0, mx0, d_start, (uint64_t)tile_lo + pos_k0) : false;
    const bool prev0 = k0 == 0 ? cont0 : in0;
    // A chunk whose whole depth range stays inside [min, max] (and was inside before), or stays outside, has no
    // transition: no walk needed.  With groups the thresholds depend on the position's class: always walk.
    const bool has = k0 < k1;
    bool quiet = !has;
    if (!MULTI && has) {
      const int32_t dlo = d_start + rmin, dhi = d_start + rmax;
      const bool all_in = dlo >= mn0 && dhi <= mx0, all_out = dhi < mn0 || dlo > mx0;
      quiet = (all_in && prev0) || (all_out && !prev0);
    }
    int nopen = 0;
    if (!quiet) {
      int32_t d = d_start;
      bool prev_in = prev0;
      RankPos it = rp;
      for (int k = k0; k < k1; ++k) {
        d += delta_get(delta, k);
        bool cur = in_rng<MULTI>(cfg, mn0, mx0, d, (uint64_t)tile_lo + (MULTI ? it.next() : 0u));
        nopen += (cur && !prev_in) ? 1 : 0;
        prev_in = cur;
      }
    }
    int32_t open_total = 0, open_pre = 0;
    if (__syncthreads_or(quiet ? 0 : 1)) open_pre = block_excl_scan(nopen, &open_total, scratch[2]);   // block-uniform
    const int np = (cont0 ? 1 : 0) + open_total;
    if (np > MAXP) {
      if (tid == 0) { *out.fail = 3; s_nh = 0; }
      nh_carry = -1;
      for (int k = tid; k < nW; k += THREADS) bitmap[k] = 0;    // the counting walk ended before the barrier above
      __syncthreads();
      continue;
    }
    uint32_t last_pos = 0;
    if (MULTI && total > 0) { RankPos lp; lp.seek(bitmap, wpre, nW, total - 1); last_pos = lp.next(); }
    // depth after the tile's last event, judged with that event's group (tile_lo + W - 1 may lie beyond the last class)
    const bool reach_end = np > 0 && in_rng<MULTI>(cfg, mn0, mx0, depth_in + tile_net, (uint64_t)tile_lo + last_pos);
    // The piece tables are clean (pen = 0, accumulators neutral).  Thread 0 writes the two entries no walking thread
    // writes: the start of a piece continuing from the previous tile and the end of a piece still open at the tile end.
    if (tid == 0) {
      if (cont0) { pst[0] = tile_lo; atomicMax(&pacc[0], depth_in); }
      if (reach_end) pen[np - 1] = tile_lo + W;
    }
    {
      const int cur0 = (cont0 ? 1 : 0) + open_pre - 1;     // piece open at my first rank (when prev0)
      if (!quiet) {
        int32_t d = d_start;
        bool prev_in = prev0;
        int cur = cur0;
        int32_t run_max = 0;
        if (!MULTI) rp.seek(bitmap, wpre, nW, k0);
        for (int k = k0; k < k1; ++k) {
          d += delta_get(delta, k);
          uint32_t pos = tile_lo + rp.next();
          bool cin = in_rng<MULTI>(cfg, mn0, mx0, d, pos);
          if (cin) {
            if (!prev_in) { ++cur; pst[cur] = pos; run_max = d; }
            else if (d > run_max) run_max = d;
          } else if (prev_in) {
            pen[cur] = pos;
            if (run_max > 0) atomicMax(&pacc[cur], run_max);
            run_max = 0;
          }
          prev_in = cin;
        }
        if (prev_in && run_max > 0 && cur >= 0) atomicMax(&pacc[cur], run_max);
      }
      // quiet chunks inside a run only raise the run's maximum depth: one atomic per warp when all lanes agree on the run
      const int32_t qmax = (quiet && has && prev0) ? d_start + rmax : 0;
      const unsigned qm = __ballot_sync(0xffffffffu, qmax > 0);
      if (qm) {
        const int c0 = __shfl_sync(0xffffffffu, cur0, __ffs(qm) - 1);
        if (__all_sync(0xffffffffu, qmax <= 0 || cur0 == c0)) {
          const int32_t wmax = __reduce_max_sync(0xffffffffu, qmax);
          if (lane == 0) atomicMax(&pacc[c0], wmax);
        } else if (qmax > 0) atomicMax(&pacc[cur0], qmax);
      }
    }
    __syncthreads();
    for (int k = tid; k < nW; k += THREADS) bitmap[k] = 0;    // no reader left; the next tile marks after another barrier
    // ---- phase 5: GMAP4, region-driven: the regions open at the tile start, then the tile's own
    if (np == 1) {
      // one piece (the usual case under dense coverage): no search, thread-local accumulation, one reduction per warp
      const uint32_t p0 = pst[0], p1 = pen[0];
      int32_t cnt = 0;
      uint32_t a_smin = 0xffffffffu, a_emax = 0, a_smax = 0, a_emin = 0xffffffffu;
      if (!cont0) {                                         // the open regions only feed a run that STARTS in this tile
        for (int k = tid; k < nh; k += THREADS) {
          uint32_t rs = hinS[k], re = rs + (hinL[k] & ~EXT);
          if (re > tile_lo && p1 > rs && p0 < re) { ++cnt; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
        }
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        if (s0 + tid + r * THREADS < s1) {
          uint32_t rs = rsr[r], re = rs + (lpr[r] & ~EXT);
          if (p1 > rs && p0 < re) { ++cnt; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
        }
      }
      for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
        const uint32_t pr = P[i]; uint32_t rs = tile_lo + (pr >> 16), re = rs + (pr & lmask);
        if (p1 > rs && p0 < re) { ++cnt; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      a_smin = __reduce_min_sync(0xffffffffu, a_smin); a_emax = __reduce_max_sync(0xffffffffu, a_emax);
      a_smax = __reduce_max_sync(0xffffffffu, a_smax); a_emin = __reduce_min_sync(0xffffffffu, a_emin);
      if (lane == 0 && cnt > 0) {
        atomicAdd(&pcnt[0], cnt);
        atomicMin(&psmin[0], a_smin); atomicMax(&pemax[0], a_emax);
        atomicMax(&psmax[0], a_smax); atomicMin(&pemin[0], a_emin);
      }
    } else if (np > 1) {
      for (int k0h = 0; k0h < nh; k0h += THREADS) {
        int k = k0h + tid;
        bool valid = k < nh;
        uint32_t rs = 0, re = 0;
        if (valid) { rs = hinS[k]; re = rs + (hinL[k] & ~EXT); valid = re > tile_lo; }
        gmap4_round(valid, rs, re, false, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        if ((uint32_t)(r * THREADS) < nS) {                 // block-uniform
          bool valid = s0 + tid + r * THREADS < s1;
          gmap4_round(valid, rsr[r], rsr[r] + (lpr[r] & ~EXT), true, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
        }
      }
      for (uint32_t i0 = s0 + RC * THREADS; i0 < s1; i0 += THREADS) {
        uint32_t i = i0 + tid;
        bool valid = i < s1;
        uint32_t rs = 0, re = 0;
        if (valid) { const uint32_t pr = P[i]; rs = tile_lo + (pr >> 16); re = rs + (pr & lmask); }
        gmap4_round(valid, rs, re, true, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
      }
    }
    if (np > 0) {
      __syncthreads();
      // ---- publish the tile's pieces in this block's segment
      if (my_cnt + np > out.seg) { if (tid == 0) *out.fail = 1; }
      else {
        const int64_t gb = (int64_t)blockIdx.x * out.seg + my_cnt;
        for (int k = tid; k < np; k += THREADS) {
          out.pstart[gb + k] = pst[k]; out.pend[gb + k] = pen[k]; out.acc[gb + k] = pacc[k]; out.cnt[gb + k] = pcnt[k];
          out.smin[gb + k] = psmin[k]; out.emax[gb + k] = pemax[k]; out.smax[gb + k] = psmax[k]; out.emin[gb + k] = pemin[k];
          out.flags[gb + k] = (uint8_t)(((k == 0 && cont0) ? 1 : 0) | ((k == np - 1 && reach_end) ? 2 : 0));
        }
        my_cnt += np;
      }
      for (int k = tid; k < np; k += THREADS) { pacc[k] = 0; pcnt[k] = 0; psmin[k] = 0xffffffffu; pemax[k] = 0; psmax[k] = 0; pemin[k] = 0xffffffffu; pen[k] = 0; }
    } else __syncthreads();
    // the regions still open at the tile end are the next tile's open regions
    { uint32_t* t = hinS; hinS = houtS; houtS = t; t = hinL; hinL = houtL; houtL = t; }
    nh_carry = nout;
  }
  if (tid == 0) out.blk_cnt[blockIdx.x] = (uint32_t)my_cnt;
}

#include "cover_tile_wc.cuh"

// per-block piece counts -> first global index of every block's pieces, and their total (one block)
static __global__ void __launch_bounds__(1024) k_seg_offsets(const uint32_t* __restrict__ blk_cnt, int nb, uint32_t* __restrict__ blk_off, unsigned long long* __restrict__ total) {
  __shared__ uint32_t wt[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < nb ? blk_cnt[i] : 0u, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wt[w] = x;
    __syncthreads();
    uint32_t b = carry;
    for (int q = 0; q < w; ++q) b += wt[q];
    if (i < nb) blk_off[i] = b + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = b + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) { blk_off[nb] = carry; *total = carry; }
}

// global piece index -> slot (pieces in genome order), and the run tables' initial values
struct Runs;

// ---- stitch: pieces sorted by start -> runs ------------------------------------------------------------------------
static __global__ void k_piece_heads(const uint32_t* __restrict__ order, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend, const uint8_t* __restrict__ flags, int64_t np, int32_t* __restrict__ head) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    bool merge = false;
    if (k > 0) {
      uint32_t a = order[k - 1], b = order[k];
      merge = (flags[b] & 1) && (flags[a] & 2) && pend[a] == pstart[b];
    }
    head[k] = merge ? 0 : 1;
  }
}

struct Runs {
  uint32_t* rstart; uint32_t* rend; int32_t* acc; int32_t* cnt; uint32_t* smin; uint32_t* emax; uint32_t* smax; uint32_t* emin;
  int32_t* edge;   // bit 0: the run continues a run of the tile before the first piece (another shard's), bit 1: it reaches beyond its last piece's tile
};

static __global__ void k_piece_order(const uint32_t* __restrict__ blk_off, int nb, int64_t seg, int64_t np, uint32_t* __restrict__ order, Runs r) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < np; g += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nb;                                   // last block whose first index is <= g
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((int64_t)blk_off[mid] <= g) lo = mid; else hi = mid; }
    order[g] = (uint32_t)((int64_t)lo * seg + (g - (int64_t)blk_off[lo]));
    r.acc[g] = 0; r.cnt[g] = 0; r.emax[g] = 0; r.smax[g] = 0; r.smin[g] = 0xffffffffu; r.emin[g] = 0xffffffffu; r.edge[g] = 0;
  }
}

static __global__ void k_piece_merge(const uint32_t* __restrict__ order, Pieces p, int64_t np, const int32_t* __restrict__ head, const int64_t* __restrict__ head_excl, Runs r) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t a = order[k];
    int64_t o = head[k] ? head_excl[k] : head_excl[k] - 1;
    if (head[k]) { r.rstart[o] = p.pstart[a]; if (p.flags[a] & 1) atomicOr(&r.edge[o], 1); }        // no piece before it here: the run comes from the previous shard
    if (k + 1 == np || head[k + 1]) { r.rend[o] = p.pend[a]; if (p.flags[a] & 2) atomicOr(&r.edge[o], 2); }   // ... and goes on in the next one
    atomicMax(&r.acc[o], p.acc[a]);
    if (p.cnt[a]) {
      atomicAdd(&r.cnt[o], p.cnt[a]);
      atomicMin(&r.smin[o], p.smin[a]); atomicMax(&r.emax[o], p.emax[a]);
      atomicMax(&r.smax[o], p.smax[a]); atomicMin(&r.emin[o], p.emin[a]);
    }
  }
}

// shard mode: per run, the raw material a cross-shard merge needs (null pointers otherwise).  Coordinates are chromosome-local;
// smin / emin = -1 and emax / smax = -1 when the run has no contributor on this shard.
struct ShardRaw { int32_t* edge; int32_t* cnt; int64_t* rstart; int64_t* rstop; int64_t* smin; int64_t* emax; int64_t* smax; int64_t* emin; };

// runs -> the arrays the common output stage (k_cover_compact) consumes
static __global__ void k_runs_finalize(Runs r, int64_t nr, const int64_t* __restrict__ n_runs, LinMap lm, int flat, uint32_t* __restrict__ ocls, int32_t* __restrict__ ochrom, int64_t* __restrict__ fstart, int64_t* __restrict__ fstop,
                                       int32_t* __restrict__ oacc, int32_t* __restrict__ keep, double* __restrict__ j1, double* __restrict__ j2, ShardRaw raw) {
  const int64_t live = *n_runs;                  // entries [live, nr) are unused capacity
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nr; k += (int64_t)gridDim.x * blockDim.x) {
    if (k >= live) {
      keep[k] = 0; ocls[k] = 0; ochrom[k] = 0; oacc[k] = 0; fstart[k] = 0; fstop[k] = 0; j1[k] = 0.0; j2[k] = 0.0;
      if (raw.edge) { raw.edge[k] = 0; raw.cnt[k] = 0; raw.rstart[k] = 0; raw.rstop[k] = 0; raw.smin[k] = -1; raw.emax[k] = -1; raw.smax[k] = -1; raw.emin[k] = -1; }
      continue;
    }
    uint64_t x = r.rstart[k];
    uint64_t cls = x / lm.G, rem = x - cls * lm.G;
    int lo = 0, hi = lm.n_chrom;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (lm.coff[mid] <= rem) lo = mid; else hi = mid; }
    uint64_t base = cls * lm.G + lm.coff[lo];
    int64_t cs = (int64_t)(x - base), ce = (int64_t)((uint64_t)r.rend[k] - base);
    int32_t cnt = r.cnt[k];
    ocls[k] = (uint32_t)cls; ochrom[k] = lo; oacc[k] = r.acc[k];
    keep[k] = cnt > 0 ? 1 : 0;                                     // GMAP4.scala:33-34,48
    if (raw.edge) {
      const int32_t eg = r.edge[k];
      if (eg) keep[k] = 1;                                         // its contributors may all live on the neighbouring shard
      raw.edge[k] = eg; raw.cnt[k] = cnt; raw.rstart[k] = cs; raw.rstop[k] = ce;
      raw.smin[k] = cnt > 0 ? (int64_t)((uint64_t)r.smin[k] - base) : -1; raw.emax[k] = cnt > 0 ? (int64_t)((uint64_t)r.emax[k] - base) : -1;
      raw.smax[k] = cnt > 0 ? (int64_t)((uint64_t)r.smax[k] - base) : -1; raw.emin[k] = cnt > 0 ? (int64_t)((uint64_t)r.emin[k] - base) : -1;
    }
    if (cnt > 0) {
      uint64_t smin = r.smin[k], emax = r.emax[k], smax = r.smax[k], emin = r.emin[k];
      int64_t os = flat ? (int64_t)(smin - base) : cs, oe = flat ? (int64_t)(emax - base) : ce;   // :83-84
      fstart[k] = os; fstop[k] = oe;
      long long den = (long long)(emax - smin);
      double num2 = cnt == 1 ? (double)(long long)(emin - smax) : (smax < emin ? (double)(long long)(emin - smax) : 0.0);   // :74
      j1[k] = den != 0 ? fabs(((double)oe - (double)os) / (double)den) : 0.0;                      // :91
      j2[k] = den != 0 ? fabs(num2 / (double)den) : 0.0;                                           // :93
    } else { fstart[k] = cs; fstop[k] = ce; j1[k] = 0.0; j2[k] = 0.0; }
  }
}


// ---- fused stitch: one thread-block cluster from the pieces to the result rows ---------------------------------------
// Everything that follows the tile kernel when there are at most STITCH_MAX pieces (COVER / FLAT without aggregates,
// not a shard) runs in ONE launch of one cluster of STITCH_CL blocks: block offsets, genome order, run heads, run
// indices, the merge of the pieces of a run, the GMAP4 epilogue of k_runs_finalize and the compaction of the kept runs
// into the result columns.  Every block owns a contiguous range of the pieces (then of the runs); the two prefix sums
// that cross block ranges (runs before my range, kept runs before my range) exchange one total per block through
// distributed shared memory, with the hardware cluster barrier in between.  Nothing returns to the host, so the
// operator needs a single synchronisation after the tile kernel; as separate kernels these steps were launch- and
// round-trip-bound.
constexpr int STITCH_T = 1024;
constexpr int STITCH_CL = 8;            // blocks of the cluster (portable maximum)
constexpr int STITCH_MAX = 1 << 18;     // pieces
constexpr int STITCH_MAXB = 4096;       // tile-kernel blocks whose offsets fit the shared table

struct StitchOut {
  int32_t* group; int8_t* strand; int32_t* chrom; int64_t* start; int64_t* stop; int32_t* acc; double* accd; double* j1; double* j2;
  unsigned long long* totals;           // [0] pieces, [1] runs, [2] kept runs, [3] 1 = too many pieces for this kernel (nothing else written)
};

__device__ __forceinline__ int32_t stitch_incl_scan(int32_t v, int32_t* total, int32_t* wt /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  __syncthreads();                                         // the previous call's readers are done with wt
  if (lane == 31) wt[w] = x;
  __syncthreads();
  int32_t t = wt[lane];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
  *total = __shfl_sync(0xffffffffu, t, 31);
  const int32_t below = __shfl_sync(0xffffffffu, t, w > 0 ? w - 1 : 0);
  return (w > 0 ? below : 0) + x;
}

static __global__ void __cluster_dims__(STITCH_CL, 1, 1) __launch_bounds__(STITCH_T, 1)
k_stitch_small(const uint32_t* __restrict__ blk_cnt, int nb, uint32_t* __restrict__ blk_off, Pieces p, uint32_t* __restrict__ order, int32_t* __restrict__ ridx, Runs r, LinMap lm,
               int flat, int n_classes, int unified, StitchOut out, Guard guard) {
  if (guard_trips(guard)) return;                          // uniform over the cluster
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();
  __shared__ uint32_t s_off[STITCH_MAXB + 1];
  __shared__ int32_t wt[32];
  __shared__ int32_t x_heads, x_kept;                      // this block's totals, read by the other blocks of the cluster
  __shared__ int32_t hp[STITCH_CL + 1], kp[STITCH_CL + 1]; // exclusive prefixes over the cluster's blocks
  const int tid = threadIdx.x, lane = tid & 31;
  // block offsets of the tile kernel's segments (every block of the cluster computes the small table itself)
  int32_t carry = 0;
  for (int base = 0; base < nb; base += STITCH_T) {
    const int i = base + tid;
    const int32_t v = i < nb ? (int32_t)blk_cnt[i] : 0;
    int32_t tot;
    const int32_t incl = stitch_incl_scan(v, &tot, wt);
    if (i < nb) s_off[i] = (uint32_t)(carry + incl - v);
    carry += tot;
  }
  const int32_t np = carry;
  if (tid == 0) s_off[nb] = (uint32_t)np;
  __syncthreads();
  if (cr == 0) {
    for (int i = tid; i <= nb; i += STITCH_T) blk_off[i] = s_off[i];
    if (tid == 0) { out.totals[0] = (unsigned long long)np; if (np > STITCH_MAX) out.totals[3] = 1ull; }
  }
  if (np > STITCH_MAX) return;                             // the whole cluster leaves together
  auto slot_of = [&](int32_t g) -> uint32_t {              // last block whose first index is <= g holds piece g
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int32_t)s_off[mid] <= g) lo = mid; else hi = mid; }
    return (uint32_t)((int64_t)lo * p.seg + (int64_t)(g - (int32_t)s_off[lo]));
  };
  // genome order, run heads, run index of every piece inside its block's range
  const int32_t L = max(1, (np + STITCH_CL - 1) / STITCH_CL);
  const int32_t g_lo = min(np, cr * L), g_hi = min(np, g_lo + L);
  carry = 0;
  for (int base = g_lo; base < g_hi; base += STITCH_T) {
    const int32_t g = base + tid;
    int32_t head = 0;
    if (g < g_hi) {
      const uint32_t a = slot_of(g);
      order[g] = a;
      head = 1;
      if (g > 0 && (p.flags[a] & 1)) {
        const uint32_t b = slot_of(g - 1);
        if ((p.flags[b] & 2) && p.pend[b] == p.pstart[a]) head = 0;
      }
    }
    int32_t tot;
    const int32_t incl = stitch_incl_scan(head, &tot, wt);
    if (g < g_hi) ridx[g] = carry + incl - 1;
    carry += tot;
  }
  if (tid == 0) x_heads = carry;
  cluster.sync();
  if (tid == 0) {
    int32_t run = 0;
    for (int c = 0; c < STITCH_CL; ++c) { hp[c] = run; run += *cluster.map_shared_rank(&x_heads, c); }
    hp[STITCH_CL] = run;
  }
  __syncthreads();
  const int32_t n_runs = hp[STITCH_CL];
  for (int k = cr * STITCH_T + tid; k < n_runs; k += STITCH_CL * STITCH_T) { r.acc[k] = 0; r.cnt[k] = 0; r.smin[k] = 0xffffffffu; r.emax[k] = 0; r.smax[k] = 0; r.emin[k] = 0xffffffffu; }
  cluster.sync();
  // merge the pieces of every run: segmented warp reduction, then one atomic per warp and run
  for (int base = g_lo; base < g_hi; base += STITCH_T) {
    const int32_t g = base + tid;
    const bool valid = g < g_hi;
    int32_t key = -1 - lane;
    int32_t v_acc = 0, v_cnt = 0;
    uint32_t v_smin = 0xffffffffu, v_emax = 0, v_smax = 0, v_emin = 0xffffffffu;
    if (valid) {
      const uint32_t a = order[g];
      key = hp[cr] + ridx[g];
      const int32_t kp_ = g > 0 ? hp[(g - 1) / L] + ridx[g - 1] : -1, kn_ = g + 1 < np ? hp[(g + 1) / L] + ridx[g + 1] : -1;
      if (kp_ != key) r.rstart[key] = p.pstart[a];
      if (kn_ != key) r.rend[key] = p.pend[a];
      v_acc = p.acc[a]; v_cnt = p.cnt[a];
      if (v_cnt) { v_smin = p.smin[a]; v_emax = p.emax[a]; v_smax = p.smax[a]; v_emin = p.emin[a]; }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {                     // equal keys are contiguous: reduce towards the first lane of a run
      const int32_t k2 = __shfl_down_sync(0xffffffffu, key, d);
      const int32_t a2 = __shfl_down_sync(0xffffffffu, v_acc, d), c2 = __shfl_down_sync(0xffffffffu, v_cnt, d);
      const uint32_t s2 = __shfl_down_sync(0xffffffffu, v_smin, d), e2 = __shfl_down_sync(0xffffffffu, v_emax, d);
      const uint32_t m2 = __shfl_down_sync(0xffffffffu, v_smax, d), n2 = __shfl_down_sync(0xffffffffu, v_emin, d);
      if (lane + d < 32 && k2 == key) { v_acc = max(v_acc, a2); v_cnt += c2; v_smin = min(v_smin, s2); v_emax = max(v_emax, e2); v_smax = max(v_smax, m2); v_emin = min(v_emin, n2); }
    }
    const int32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
    if (valid && (lane == 0 || kprev != key)) {
      atomicMax(&r.acc[key], v_acc);
      if (v_cnt) {
        atomicAdd(&r.cnt[key], v_cnt);
        atomicMin(&r.smin[key], v_smin); atomicMax(&r.emax[key], v_emax);
        atomicMax(&r.smax[key], v_smax); atomicMin(&r.emin[key], v_emin);
      }
    }
  }
  cluster.sync();
  // runs -> result rows (the epilogue of k_runs_finalize; runs without a contributor are dropped, GMAP4.scala:33-34,48)
  const int32_t R = max(1, (n_runs + STITCH_CL - 1) / STITCH_CL);
  const int32_t k_lo = min(n_runs, cr * R), k_hi = min(n_runs, k_lo + R);
  {
    int32_t mine = 0;
    for (int k = k_lo + tid; k < k_hi; k += STITCH_T) mine += r.cnt[k] > 0 ? 1 : 0;
    int32_t tot;
    stitch_incl_scan(mine, &tot, wt);
    if (tid == 0) x_kept = tot;
  }
  cluster.sync();
  if (tid == 0) {
    int32_t run = 0;
    for (int c = 0; c < STITCH_CL; ++c) { kp[c] = run; run += *cluster.map_shared_rank(&x_kept, c); }
    kp[STITCH_CL] = run;
  }
  __syncthreads();
  carry = kp[cr];
  for (int base = k_lo; base < k_hi; base += STITCH_T) {
    const int32_t k = base + tid;
    int32_t keep = 0, cnt = 0, chrom = 0;
    uint64_t cls = 0, basek = 0;
    int64_t cs = 0, ce = 0;
    if (k < k_hi) {
      const uint64_t x = r.rstart[k];
      cls = x / lm.G;
      const uint64_t rem = x - cls * lm.G;
      int lo = 0, hi = lm.n_chrom;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lm.coff[mid] <= rem) lo = mid; else hi = mid; }
      chrom = lo;
      basek = cls * lm.G + lm.coff[lo];
      cs = (int64_t)(x - basek); ce = (int64_t)((uint64_t)r.rend[k] - basek);
      cnt = r.cnt[k];
      keep = cnt > 0 ? 1 : 0;
    }
    int32_t tot;
    const int32_t incl = stitch_incl_scan(keep, &tot, wt);
    if (keep) {
      const int64_t o = (int64_t)(carry + incl - 1);
      const uint64_t smin = r.smin[k], emax = r.emax[k], smax = r.smax[k], emin = r.emin[k];
      const int64_t os = flat ? (int64_t)(smin - basek) : cs, oe = flat ? (int64_t)(emax - basek) : ce;   // :83-84
      const long long den = (long long)(emax - smin);
      const double num2 = cnt == 1 ? (double)(long long)(emin - smax) : (smax < emin ? (double)(long long)(emin - smax) : 0.0);   // :74
      const int32_t acc = r.acc[k];
      out.group[o] = (int32_t)((uint32_t)cls / (uint32_t)n_classes);
      out.strand[o] = unified ? (int8_t)0 : (int8_t)((uint32_t)cls % (uint32_t)n_classes + 1);
      out.chrom[o] = chrom;
      out.start[o] = os; out.stop[o] = oe;
      out.acc[o] = acc; out.accd[o] = (double)acc;
      out.j1[o] = den != 0 ? fabs(((double)oe - (double)os) / (double)den) : 0.0;                      // :91
      out.j2[o] = den != 0 ? fabs(num2 / (double)den) : 0.0;                                           // :93
    }
    carry += tot;
  }
  if (cr == 0 && tid == 0) { out.totals[1] = (unsigned long long)n_runs; out.totals[2] = (unsigned long long)kp[STITCH_CL]; }
  cluster.sync();                                          // no block leaves while its totals may still be read
}

}  // namespace tiled
