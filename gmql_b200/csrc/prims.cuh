// prims.cuh — hand-written device primitives: LSD radix sort (32/64-bit keys, optional 32-bit payload),
// device-wide scans, lower/upper bound.  HBM-bound integer work: coalesced warp-striped loads, shared-memory
// ranking, grids sized as multiples of the SM count (148 on B200).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// binary searches over sorted device arrays
// ------------------------------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ int64_t lower_bound_dev(const K* __restrict__ a, int64_t n, K x) {
  int64_t lo = 0, hi = n;  // first index with a[i] >= x
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
template <typename K>
__device__ __forceinline__ int64_t upper_bound_dev(const K* __restrict__ a, int64_t n, K x) {
  int64_t lo = 0, hi = n;  // first index with a[i] > x
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// device-wide scan (reduce -> scan partials -> apply); Op is an associative functor with identity()
// ------------------------------------------------------------------------------------------------
template <typename T> struct OpSum { __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } __device__ __forceinline__ static T identity() { return (T)0; } };
template <typename T> struct OpMax { __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; } __device__ __forceinline__ static T identity() { return (T)0; } };

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T, typename Op>
__device__ __forceinline__ T block_scan_excl(T v, T* total, Op op) {
  // exclusive scan of one value per thread across a 256-thread block
  __shared__ T warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x = op(y, x);
  }
  if (lane == 31) warp_tot[w] = x;
  __syncthreads();
  T base = Op::identity();
  for (int i = 0; i < w; ++i) base = op(base, warp_tot[i]);
  T tot = Op::identity();
  for (int i = 0; i < SCAN_THREADS / 32; ++i) tot = op(tot, warp_tot[i]);
  T incl = op(base, x);
  T prev = __shfl_up_sync(0xffffffffu, incl, 1);
  T excl = lane == 0 ? base : prev;
  __syncthreads();
  *total = tot;
  return excl;
}

template <typename Tin, typename T, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const Tin* __restrict__ in, int64_t n, T* __restrict__ partials, int64_t tiles_per_block) {
  Op op;
  T acc = Op::identity();
  int64_t t0 = (int64_t)blockIdx.x * tiles_per_block;
  for (int64_t t = t0; t < t0 + tiles_per_block; ++t) {
    int64_t base = t * SCAN_TILE;
    if (base >= n) break;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
      int64_t i = base + (int64_t)j * SCAN_THREADS + threadIdx.x;
      if (i < n) acc = op(acc, (T)in[i]);
    }
  }
  __shared__ T red[SCAN_THREADS];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = SCAN_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = op(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

template <typename T, typename Op>
__global__ void __launch_bounds__(1024) k_scan_partials(T* partials, int nb, T* total_out) {
  // single block: exclusive scan of nb partials (nb <= a few thousand)
  Op op;
  __shared__ T sh[1024];
  __shared__ T carry;
  if (threadIdx.x == 0) carry = Op::identity();
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    T v = i < nb ? partials[i] : Op::identity();
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      T y = threadIdx.x >= o ? sh[threadIdx.x - o] : Op::identity();
      __syncthreads();
      sh[threadIdx.x] = op(y, sh[threadIdx.x]);
      __syncthreads();
    }
    T incl = sh[threadIdx.x];
    T c = carry;
    T excl = threadIdx.x == 0 ? c : op(c, sh[threadIdx.x - 1]);
    if (i < nb) partials[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = op(c, incl);
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename Tin, typename T, typename Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const Tin* __restrict__ in, T* __restrict__ out, int64_t n, const T* __restrict__ partials, int64_t tiles_per_block) {
  Op op;
  T carry = partials[blockIdx.x];
  int64_t t0 = (int64_t)blockIdx.x * tiles_per_block;
  for (int64_t t = t0; t < t0 + tiles_per_block; ++t) {
    int64_t base = t * SCAN_TILE;
    if (base >= n) break;
    // blocked arrangement: thread owns SCAN_ITEMS consecutive elements
    T v[SCAN_ITEMS];
    int64_t i0 = base + (int64_t)threadIdx.x * SCAN_ITEMS;
    T tsum = Op::identity();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
      v[j] = (i0 + j < n) ? (T)in[i0 + j] : Op::identity();
      tsum = op(tsum, v[j]);
    }
    T tile_total;
    T excl = block_scan_excl<T, Op>(tsum, &tile_total, op);
    T run = op(carry, excl);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
      if (INCLUSIVE) { run = op(run, v[j]); if (i0 + j < n) out[i0 + j] = run; }
      else { if (i0 + j < n) out[i0 + j] = run; run = op(run, v[j]); }
    }
    carry = op(carry, tile_total);
  }
}

// small inputs: ONE block scans everything, 4 096 elements per iteration (one launch instead of three; for a few thousand
// elements the launches, not the work, are what costs)
constexpr int64_t SCAN_SMALL_MAX = 1 << 15;
template <typename Tin, typename T, typename Op, bool INCLUSIVE>
__global__ void __launch_bounds__(1024) k_scan_small(const Tin* in, T* out, int64_t n, T* total_out) {
  Op op;
  __shared__ T wt[32];
  __shared__ T s_carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_carry = Op::identity();
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t i0 = base + (int64_t)tid * 4;
    T v[4];
    T acc = Op::identity();
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = i0 + j < n ? (T)in[i0 + j] : Op::identity(); acc = op(acc, v[j]); }
    T x = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = op(y, x); }
    if (lane == 31) wt[w] = x;
    __syncthreads();
    T b = s_carry;
    for (int q = 0; q < w; ++q) b = op(b, wt[q]);
    T prev = __shfl_up_sync(0xffffffffu, x, 1);
    T run = lane == 0 ? b : op(b, prev);                    // exclusive prefix of this thread's four elements
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (INCLUSIVE) { run = op(run, v[j]); if (i0 + j < n) out[i0 + j] = run; }
      else { if (i0 + j < n) out[i0 + j] = run; run = op(run, v[j]); }
    }
    __syncthreads();
    if (tid == 1023) s_carry = run;
    __syncthreads();
  }
  if (total_out && tid == 0) *total_out = s_carry;
}

// out may alias in when Tin == T.  total (device pointer, optional) receives the reduction of all inputs.
template <typename Tin, typename T, typename Op, bool INCLUSIVE>
static void device_scan(gmql_ctx* ctx, const Tin* in, T* out, int64_t n, T* total_dev) {
  if (n <= 0) {
    if (total_dev) CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(T), ctx->stream));
    return;
  }
  if (n <= SCAN_SMALL_MAX) {
    LAUNCH(ctx, (k_scan_small<Tin, T, Op, INCLUSIVE>), 1, 1024, 0, in, out, n, total_dev);
    return;
  }
  int64_t tiles = ceil_div64(n, SCAN_TILE);
  int nb = (int)std::min<int64_t>(tiles, (int64_t)ctx->num_sms * 8);
  int64_t tpb = ceil_div64(tiles, nb);
  nb = (int)ceil_div64(tiles, tpb);
  DevBuf<T> partials(ctx, nb);
  LAUNCH(ctx, (k_scan_reduce<Tin, T, Op>), nb, SCAN_THREADS, 0, in, n, partials.p, tpb);
  LAUNCH(ctx, (k_scan_partials<T, Op>), 1, 1024, 0, partials.p, nb, total_dev);
  LAUNCH(ctx, (k_scan_apply<Tin, T, Op, INCLUSIVE>), nb, SCAN_THREADS, 0, in, out, n, partials.p, tpb);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable, "onesweep" organisation:
//   * one upfront kernel histograms every digit place in a single read of the keys;
//   * each pass is ONE kernel: a tile (256 threads x ITEMS keys) ranks its keys in shared memory, obtains its global
//     per-digit offsets by decoupled look-back over the tiles before it (tile ids are handed out by an atomic ticket,
//     so every predecessor is already running), reorders the tile in shared memory and writes digit runs with
//     coalesced stores.  Traffic per pass = read n + write n elements.
// ------------------------------------------------------------------------------------------------
constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
// resident blocks per SM the sort pass is compiled for: four with 32-bit keys (64 registers: the pass waits on its key loads,
// 0.44 -> 0.39 ms per pass of 5e7 keys + payload); with 64-bit keys the compiler's own choice is the fastest of 2 / 3 / 4 (measured)
#ifndef OS64_MINB
#define OS64_MINB 0
#endif
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#define OS_FLAG_AGG (1u << 30)
#define OS_FLAG_INCL (1u << 31)
#define OS_VALUE_MASK ((1u << 30) - 1u)
template <typename K> struct OsCfg {
  static constexpr int ITEMS = sizeof(K) == 4 ? 16 : 8;
  static constexpr int TILE = RS_THREADS * ITEMS;
};

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_os_hist(const K* __restrict__ keys, int64_t n, int passes, int begin_bit, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[8][RS_RADIX];
  for (int p = 0; p < passes; ++p) h[p][threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
    K k = keys[i];
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(uint32_t)(k >> (begin_bit + p * RS_BITS)) & (RS_RADIX - 1)], 1u);
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    uint32_t c = h[p][threadIdx.x];
    if (c) atomicAdd(&hist[p * RS_RADIX + threadIdx.x], c);
  }
}

// <<<passes, 256>>>: exclusive scan of each digit place's 256 counters, in place
static __global__ void __launch_bounds__(RS_RADIX) k_os_scan_hist(uint32_t* hist) {
  __shared__ uint32_t wt[RS_RADIX / 32];
  uint32_t* h = hist + blockIdx.x * RS_RADIX;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t v = h[threadIdx.x], x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wt[w] = x;
  __syncthreads();
  uint32_t base = 0;
  for (int q = 0; q < w; ++q) base += wt[q];
  h[threadIdx.x] = base + x - v;
}

template <typename K, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS, sizeof(K) == 4 ? 4 : OS64_MINB) k_os_pass(const K* __restrict__ kin, K* __restrict__ kout, const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                        int64_t n, int shift, const uint32_t* __restrict__ digit_start, volatile uint32_t* tile_state, uint32_t* ticket) {
  constexpr int ITEMS = OsCfg<K>::ITEMS;
  constexpr int TILE = OsCfg<K>::TILE;
  __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];
  __shared__ uint32_t tile_excl[RS_RADIX];
  __shared__ uint32_t glob_base[RS_RADIX];
  __shared__ uint32_t wtot[RS_WARPS];
  __shared__ K skeys[TILE];
  __shared__ uint32_t svals[HAS_VAL ? TILE : 1];
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
  for (int ww = 0; ww < RS_WARPS; ++ww) warp_cnt[ww][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t base = (int64_t)tile * TILE;
  const int64_t wbase = base + (int64_t)w * (32 * ITEMS) + lane;

  K key[ITEMS];
  uint32_t lrank[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    int64_t i = wbase + j * 32;
    if (i < n) key[j] = kin[i];
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    int64_t i = wbase + j * 32;
    bool ok = i < n;
    uint32_t d = ok ? ((uint32_t)(key[j] >> shift) & (RS_RADIX - 1)) : RS_RADIX;   // invalid lanes form their own group
    // lanes holding the same digit, by bit-wise ballots (match.any runs on the slow ADU pipe: ~60 cycles per warp)
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b <= RS_BITS; ++b) {
      uint32_t bit = (d >> b) & 1u;
      uint32_t bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
    uint32_t r = __popc(peers & lt_mask);
    uint32_t pre = ok ? warp_cnt[w][d] : 0;
    __syncwarp();
    if (ok && r == 0) warp_cnt[w][d] = pre + __popc(peers);
    __syncwarp();
    lrank[j] = pre + r;
  }
  __syncthreads();
  // thread d owns digit d: exclusive offsets of the warps inside the digit, the tile's count, publish, look back
  uint32_t tot = 0;
#pragma unroll
  for (int ww = 0; ww < RS_WARPS; ++ww) { uint32_t c = warp_cnt[ww][threadIdx.x]; warp_cnt[ww][threadIdx.x] = tot; tot += c; }
  tile_state[(int64_t)tile * RS_RADIX + threadIdx.x] = tot | (tile == 0 ? OS_FLAG_INCL : OS_FLAG_AGG);
  {
    uint32_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wtot[w] = x;
    __syncthreads();
    uint32_t b = 0;
    for (int q = 0; q < w; ++q) b += wtot[q];
    tile_excl[threadIdx.x] = b + x - tot;
  }
  __syncthreads();
  // reorder the tile in shared memory first (needs only tile-local offsets); the look-back below then finds more
  // predecessors already finished
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    int64_t i = wbase + j * 32;
    if (i < n) {
      uint32_t d = (uint32_t)(key[j] >> shift) & (RS_RADIX - 1);
      uint32_t pos = tile_excl[d] + warp_cnt[w][d] + lrank[j];
      skeys[pos] = key[j];
      if (HAS_VAL) svals[pos] = vin[i];
    }
  }
  // decoupled look-back, 8 predecessors per round trip
  uint32_t look = 0;
  if (tile > 0) {
    int64_t p = (int64_t)tile - 1;
    bool done = false;
    while (!done) {
      uint32_t s8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) s8[q] = (p - q >= 0) ? tile_state[(p - q) * RS_RADIX + threadIdx.x] : OS_FLAG_INCL;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (done) break;
        uint32_t sv = s8[q];
        if ((sv & (OS_FLAG_AGG | OS_FLAG_INCL)) == 0) break;     // not published yet: poll again from here
        look += sv & OS_VALUE_MASK;
        --p;
        if (sv & OS_FLAG_INCL) done = true;
      }
    }
    tile_state[(int64_t)tile * RS_RADIX + threadIdx.x] = ((look + tot) & OS_VALUE_MASK) | OS_FLAG_INCL;
  }
  glob_base[threadIdx.x] = digit_start[threadIdx.x] + look - tile_excl[threadIdx.x];
  __syncthreads();
  int64_t rem = n - base;
  int cnt = rem < TILE ? (int)rem : TILE;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    int i = threadIdx.x + j * RS_THREADS;
    if (i < cnt) {
      K k = skeys[i];
      uint32_t d = (uint32_t)(k >> shift) & (RS_RADIX - 1);
      uint32_t g = glob_base[d] + (uint32_t)i;
      kout[g] = k;
      if (HAS_VAL) vout[g] = svals[i];
    }
  }
}

// Sorts keys (and a 32-bit payload when vals != nullptr) on bits [0, key_bits).  Buffers ping-pong; on return
// *keys_out / *vals_out point at whichever buffer holds the sorted data.  n must be < 2^30.  With begin_bit > 0 only
// bits [begin_bit, key_bits) are ordered (a stable partition by the high bits).
template <typename K>
static void radix_sort(gmql_ctx* ctx, K* keys, K* keys_alt, uint32_t* vals, uint32_t* vals_alt, int64_t n, int key_bits, K** keys_out, uint32_t** vals_out, int begin_bit = 0) {
  K* kin = keys; K* kout = keys_alt;
  uint32_t* vin = vals; uint32_t* vout = vals_alt;
  if (n > 1 && key_bits > begin_bit) {
    GMQL_REQUIRE(n < (int64_t)(1 << 30), GMQL_ERR_UNSUPPORTED, "radix_sort: 2^30 or more elements in one sort");
    int passes = (key_bits - begin_bit + RS_BITS - 1) / RS_BITS;
    int64_t tiles = ceil_div64(n, OsCfg<K>::TILE);
    DevBuf<uint32_t> hist(ctx, (int64_t)8 * RS_RADIX);
    DevBuf<uint32_t> state(ctx, tiles * RS_RADIX + 32);
    hist.zero();
    int hb = (int)std::min<int64_t>(ceil_div64(n, RS_THREADS * 16), (int64_t)ctx->num_sms * 8);
    LAUNCH(ctx, (k_os_hist<K>), hb, RS_THREADS, 0, kin, n, passes, begin_bit, hist.p);
    LAUNCH(ctx, k_os_scan_hist, passes, RS_RADIX, 0, hist.p);
    for (int p = 0; p < passes; ++p) {
      int shift = begin_bit + p * RS_BITS;
      state.zero();
      uint32_t* ticket = state.p + tiles * RS_RADIX;
      if (vals) LAUNCH(ctx, (k_os_pass<K, true>), (unsigned)tiles, RS_THREADS, 0, kin, kout, vin, vout, n, shift, hist.p + p * RS_RADIX, state.p, ticket);
      else LAUNCH(ctx, (k_os_pass<K, false>), (unsigned)tiles, RS_THREADS, 0, kin, kout, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, shift, hist.p + p * RS_RADIX, state.p, ticket);
      std::swap(kin, kout);
      std::swap(vin, vout);
    }
  }
  *keys_out = kin;
  if (vals_out) *vals_out = vin;
}

static inline int bits_for(uint64_t max_value) {
  int b = 0;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}
