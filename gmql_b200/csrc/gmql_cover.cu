// gmql_cover.cu — COVER / FLAT / HISTOGRAM / SUMMIT on the device, both phases.
//
// Replaces GenometricCover.apply (GMQL-Spark/.../RegionsOperators/GenometricCover/GenometricCover.scala:30-158: assignGroups,
// extractor, per-bin coverHelper/histogramHelper/summitHelper, the boundary merge) and its second phase GMAP4.apply
// (RegionsOperators/GenometricMap/GMAP4.scala:20-100).
//
// Design (not a port of the per-bin HashMap/reduceByKey plan):
//  * every (group, unified strand) class gets its own block of a linearised genome coordinate, so ONE sort and ONE sweep
//    cover all groups, strands and chromosomes;
//  * COVER/FLAT/HISTOGRAM: the reference's per-2000bp-bin sweep + boundary merge equals a global sweep over
//    (start:+1, stop':-1), stop' = stop+1 when stop % bin == 0 and the region crosses a bin edge (extractor :353-354);
//    starts and stops' are radix-sorted separately (keys only) and merged on the fly (merge path): the depth after the
//    first m merged events is simply (#starts - #stops) consumed, so no prefix sum is needed;
//  * SUMMIT is bin-dependent in the reference (each bin restarts at depth 0, summitHelper :269-316), so the per-bin event
//    lists are emulated exactly: events are expanded per bin, sorted by (class, chrom, bin, offset), coalesced (zero-net
//    points kept) and each bin's state machine runs in one thread, followed by the cross-bin merge (:121-152);
//  * phase 2 (GMAP4): every phase-1 region queries the interval index of the input (sorted linear starts + running
//    max stop) with one warp, reducing count / startMin / endMax / startMax / endMin / aggregates over the overlapping
//    inputs; regions without contributors are dropped (:33-34,48).
#include "index.cuh"
#include "cover_tiled.cuh"
#include <cstdlib>
#include <algorithm>

namespace {

constexpr int MP_THREADS = 256;
constexpr int MP_ITEMS = 8;
constexpr int MP_TILE = MP_THREADS * MP_ITEMS;

struct CoverCfg {
  int n_groups;
  int n_classes;        // strand classes per group: 1 when unified, 2 ('+','-') otherwise
  int unified;
  int64_t B;            // BinSize.Cover
  const int32_t* sample_group;  // device [n_samples] or null
  const int32_t* mn;    // device [n_groups]
  const int32_t* mx;    // device [n_groups]
};

// ---- input scan: strand unification flag, samples that own a region, zero-length regions ----------------------------
__global__ void k_cover_inspect(const int8_t* __restrict__ strand, const int32_t* __restrict__ sample, const void* __restrict__ start, const void* __restrict__ stop, int bits,
                                int64_t n, int n_samples, int* __restrict__ flags /*[0]=has '*', [1]=zero-length, [2]=bad*/, int* __restrict__ sample_seen) {
  int has_star = 0, zero = 0, bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int st = strand ? (int)strand[i] : 0;
    if (st == 0) has_star = 1;
    if (st < 0 || st > 2) bad = 1;
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    if (s == e) zero = 1;
    if (s > e || s < 0) bad = 1;
    int sm = sample ? sample[i] : 0;
    if (sm < 0 || sm >= n_samples) bad = 1;
    else if (!sample_seen[sm]) sample_seen[sm] = 1;
  }
  if (__any_sync(0xffffffffu, has_star) && (threadIdx.x & 31) == 0) flags[0] = 1;
  if (__any_sync(0xffffffffu, zero) && (threadIdx.x & 31) == 0) flags[1] = 1;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) flags[2] = 1;
}

// gmql_cover's own prepass: per-chromosome max stop (span table) and the input flags in ONE read of the columns.
// Four consecutive regions per thread and iteration (128-bit column loads when the coordinates are 32-bit and the
// columns 16-byte aligned); the per-chromosome maxima live in shared memory (native 32-bit atomicMax for 32-bit coords).
__global__ void __launch_bounds__(256) k_cover_prepass(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits,
                                                       const int8_t* __restrict__ strand, const int32_t* __restrict__ sample, int64_t n, int n_chrom, int n_samples,
                                                       unsigned long long* __restrict__ span, int* __restrict__ flags, int vec_ok) {
  extern __shared__ unsigned long long sspan[];
  uint32_t* sspan32 = (uint32_t*)sspan;
  const bool use_sh = n_chrom <= 4096;
  const bool sh32 = use_sh && bits == 32;
  if (use_sh) {
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) { if (sh32) sspan32[c] = 0u; else sspan[c] = 0ull; }
    __syncthreads();
  }
  int has_star = 0, zero = 0, bad = 0;
  long long maxlen = 0;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x * 4) {
    int c[4], st[4], sm[4];
    int64_t s[4], e[4];
    if (i0 + 4 <= n && vec_ok && bits == 32) {
      int4 vc = *(const int4*)(chrom + i0), vs = *(const int4*)((const int32_t*)start + i0), ve = *(const int4*)((const int32_t*)stop + i0);
      c[0] = vc.x; c[1] = vc.y; c[2] = vc.z; c[3] = vc.w;
      s[0] = vs.x; s[1] = vs.y; s[2] = vs.z; s[3] = vs.w;
      e[0] = ve.x; e[1] = ve.y; e[2] = ve.z; e[3] = ve.w;
      if (sample) { int4 vb = *(const int4*)(sample + i0); sm[0] = vb.x; sm[1] = vb.y; sm[2] = vb.z; sm[3] = vb.w; }
      else { sm[0] = sm[1] = sm[2] = sm[3] = 0; }
      if (strand) { uint32_t w = *(const uint32_t*)(strand + i0); st[0] = (int)(int8_t)(w & 0xff); st[1] = (int)(int8_t)((w >> 8) & 0xff); st[2] = (int)(int8_t)((w >> 16) & 0xff); st[3] = (int)(int8_t)(w >> 24); }
      else { st[0] = st[1] = st[2] = st[3] = 0; }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t i = i0 + j;
        bool ok = i < n;
        c[j] = ok ? chrom[i] : -2;                         // -2: padding lane
        s[j] = ok ? ld_coord(start, bits, i) : 0;
        e[j] = ok ? ld_coord(stop, bits, i) : 1;
        st[j] = ok && strand ? (int)strand[i] : (ok ? 0 : 1);
        sm[j] = ok && sample ? sample[i] : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c[j] == -2) continue;
      if (st[j] == 0) has_star = 1;
      if (s[j] == e[j]) zero = 1;
      if (e[j] - s[j] > maxlen) maxlen = e[j] - s[j];
      if (c[j] < 0 || c[j] >= n_chrom || s[j] < 0 || s[j] > e[j] || st[j] < 0 || st[j] > 2 || sm[j] < 0 || sm[j] >= n_samples) { bad = 1; continue; }
      if (sh32) atomicMax(&sspan32[c[j]], (uint32_t)e[j]);
      else if (use_sh) atomicMax(&sspan[c[j]], (unsigned long long)e[j]);
      else atomicMax(&span[c[j]], (unsigned long long)e[j]);
    }
  }
  // flags and the longest region: collected per block in shared memory, then ONE store / RED per block and flag (no
  // value is read back: thousands of warps storing to, and reading from, the same 32-byte sector serialise in the L2)
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
  if (threadIdx.x == 0) {
    if (s_fl[0]) flags[0] = 1;
    if (s_fl[1]) flags[1] = 1;
    if (s_fl[2]) flags[2] = 1;
    if (s_fl[3] > 0) atomicMax(&flags[3], s_fl[3]);
  }
  if (use_sh) {
    for (int c = threadIdx.x; c < n_chrom; c += blockDim.x) {
      unsigned long long v = sh32 ? (unsigned long long)sspan32[c] : sspan[c];
      if (v) atomicMax(&span[c], v);
    }
  }
}

__global__ void k_iota_u32(uint32_t* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

__global__ void k_count_flags(const int* __restrict__ seen, int n, int* __restrict__ out) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += seen[i] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__device__ __forceinline__ uint32_t class_of(const CoverCfg& cfg, int sample, int strand) {
  int g = cfg.sample_group ? cfg.sample_group[sample] : 0;
  int sc = cfg.unified ? 0 : (strand - 1);
  return (uint32_t)(g * cfg.n_classes + sc);
}

__global__ void k_cover_classes(const int8_t* __restrict__ strand, const int32_t* __restrict__ sample, int64_t n, CoverCfg cfg, uint32_t* __restrict__ cls) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cls[i] = class_of(cfg, sample ? sample[i] : 0, strand ? (int)strand[i] : 0);
}

// ---- phase 1 (COVER / FLAT / HISTOGRAM): event keys -------------------------------------------------------------------
// start key / stop' key of every region in the linear space (extractor semantics, GenometricCover.scala:345-360)
template <typename K>
__global__ void k_cover_event_keys(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ cls,
                                   int64_t n, LinMap lm, int64_t B, K* __restrict__ skeys /*may be null*/, K* __restrict__ ekeys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * lm.G + lm.coff[chrom[i]];
    int64_t ks, ke;
    if (s == e) {           // HashMap(off->1, off->-1) keeps only the -1 (:351); its effect ends with the bin
      ks = (s / B + 1) * B;
      ke = s;
    } else {
      ks = s;
      ke = (e % B == 0 && s / B != e / B) ? e + 1 : e;   // stopOff == 0 is replaced by 1 (:354)
    }
    if (skeys) skeys[i] = (K)(base + (uint64_t)ks);
    ekeys[i] = (K)(base + (uint64_t)ke);
  }
}

// no zero-length regions: the index's start keys ARE the phase-1 start keys, so one kernel writes the start key, the
// length payload and the stop' key
template <typename K>
__global__ void k_cover_all_keys(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ cls,
                                 int64_t n, LinMap lm, int64_t B, K* __restrict__ skeys, uint32_t* __restrict__ lens, K* __restrict__ ekeys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * lm.G + lm.coff[chrom[i]];
    int64_t ke = (e % B == 0 && s / B != e / B) ? e + 1 : e;
    skeys[i] = (K)(base + (uint64_t)s);
    lens[i] = (uint32_t)(e - s);
    ekeys[i] = (K)(base + (uint64_t)ke);
  }
}

// ---- merge path ----------------------------------------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ int64_t merge_path(const K* __restrict__ A, int64_t nA, const K* __restrict__ Bv, int64_t nB, int64_t d) {
  // number of A elements among the first d of merge(A, B), ties: A first
  int64_t lo = d > nB ? d - nB : 0, hi = d < nA ? d : nA;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (A[mid] <= Bv[d - 1 - mid]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// merge-path split of every tile boundary, computed once (one thread per boundary) instead of by two threads of each
// tile while the other 254 wait at the barrier
template <typename K>
__global__ void k_merge_splits(const K* __restrict__ A, int64_t nA, const K* __restrict__ Bv, int64_t nB, int64_t n_tiles, int64_t* __restrict__ splits) {
  int64_t total = nA + nB;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t d = t * MP_TILE;
    if (d > total) d = total;
    splits[t] = merge_path<K>(A, nA, Bv, nB, d);
  }
}

// One pass over merge(starts, stops').  Emits one point (x, depth after x) per distinct position.
// WRITE == false: per-tile point counts; WRITE == true: points are written at tile_off[tile] + rank.
template <typename K, bool WRITE>
__global__ void __launch_bounds__(MP_THREADS) k_cover_points(const K* __restrict__ A, int64_t nA, const K* __restrict__ Bv, int64_t nB, int64_t n_tiles,
                                                             const int64_t* __restrict__ splits, int64_t* __restrict__ tile_cnt, const int64_t* __restrict__ tile_off,
                                                             K* __restrict__ px, int32_t* __restrict__ pd) {
  __shared__ K sA[MP_TILE + 1];
  __shared__ K sB[MP_TILE + 1];
  __shared__ int warp_tot[MP_THREADS / 32];
  const int64_t total = nA + nB;
  const K KMAX = ~(K)0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t d0 = tile * MP_TILE, d1 = d0 + MP_TILE;
    if (d1 > total) d1 = total;
    const int64_t i0 = splits[tile], i1 = splits[tile + 1];
    const int64_t j0 = d0 - i0, j1 = d1 - i1;
    const int na = (int)(i1 - i0), nb = (int)(j1 - j0);
    for (int t = threadIdx.x; t <= na; t += MP_THREADS) sA[t] = (i0 + t < nA) ? A[i0 + t] : KMAX;
    for (int t = threadIdx.x; t <= nb; t += MP_THREADS) sB[t] = (j0 + t < nB) ? Bv[j0 + t] : KMAX;
    __syncthreads();
    int ld = threadIdx.x * MP_ITEMS;
    int len = (int)(d1 - d0);
    if (ld > len) ld = len;
    int ai;
    {  // local merge path in shared memory (lookahead slots are never consumed: bounded by na / nb)
      int lo = ld > nb ? ld - nb : 0, hi = ld < na ? ld : na;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sA[mid] <= sB[ld - 1 - mid]) lo = mid + 1; else hi = mid;
      }
      ai = lo;
    }
    int bi = ld - ai;
    K xs[MP_ITEMS];
    int32_t ds[MP_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < MP_ITEMS; ++s) {
      if (ld + s < len) {
        bool takeA = (bi >= nb) || (ai < na && sA[ai] <= sB[bi]);
        K x = takeA ? sA[ai] : sB[bi];
        if (takeA) ++ai; else ++bi;
        K nxt = sA[ai] < sB[bi] ? sA[ai] : sB[bi];   // lookahead slots hold the next global elements (or KMAX)
        if (nxt != x) {
          xs[cnt] = x;
          ds[cnt] = (int32_t)((i0 + ai) - (j0 + bi));
          ++cnt;
        }
      }
    }
    // block exclusive scan of cnt
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int q = 0; q < MP_THREADS / 32; ++q) { if (q < w) base += warp_tot[q]; tot += warp_tot[q]; }
    if (!WRITE) {
      if (threadIdx.x == 0) tile_cnt[tile] = tot;
    } else {
      int64_t o = tile_off[tile] + base + incl - cnt;
      for (int s = 0; s < cnt; ++s) { px[o + s] = xs[s]; pd[o + s] = ds[s]; }
    }
    __syncthreads();
  }
}

// group of a linear coordinate (forward declaration for the fused sweep)
template <typename K>
__device__ __forceinline__ int group_of_lin(K x, uint64_t G, int n_classes);

// COVER / FLAT fused sweep: the same merge as k_cover_points, but run opens/closes are decided on the fly, so the
// point list is never materialised.  A run opens at the first distinct position whose depth enters [min, max] and
// closes at the first one that leaves it (coverHelper :181-208 + boundary merge :130-150 == maximal runs).
// WRITE == false: per-tile number of opens; WRITE == true: run k gets start (at its open), stop (at its close) and
// AccIndex = max depth (atomicMax, one per thread and run).
template <typename K, bool WRITE>
__global__ void __launch_bounds__(MP_THREADS) k_cover_runs_fused(const K* __restrict__ A, int64_t nA, const K* __restrict__ Bv, int64_t nB, int64_t n_tiles, uint64_t G, CoverCfg cfg,
                                                                 const int64_t* __restrict__ splits, int64_t* __restrict__ tile_cnt, const int64_t* __restrict__ tile_off,
                                                                 K* __restrict__ rs, K* __restrict__ re, int32_t* __restrict__ racc) {
  __shared__ K sAm[MP_TILE + 2];
  __shared__ K sBm[MP_TILE + 2];
  __shared__ int warp_tot[MP_THREADS / 32];
  K* sA = sAm + 1;   // sA[-1] = look-behind element
  K* sB = sBm + 1;
  const int64_t total = nA + nB;
  const K KMAX = ~(K)0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t d0 = tile * MP_TILE, d1 = d0 + MP_TILE;
    if (d1 > total) d1 = total;
    const int64_t i0 = splits[tile], i1 = splits[tile + 1];
    const int64_t j0 = d0 - i0, j1 = d1 - i1;
    const int na = (int)(i1 - i0), nb = (int)(j1 - j0);
    for (int t = threadIdx.x; t <= na + 1; t += MP_THREADS) { int64_t g = i0 - 1 + t; sAm[t] = (g >= 0 && g < nA) ? A[g] : KMAX; }
    for (int t = threadIdx.x; t <= nb + 1; t += MP_THREADS) { int64_t g = j0 - 1 + t; sBm[t] = (g >= 0 && g < nB) ? Bv[g] : KMAX; }
    __syncthreads();
    int len = (int)(d1 - d0);
    int ld = threadIdx.x * MP_ITEMS;
    if (ld > len) ld = len;
    int ai;
    {
      int lo = ld > nb ? ld - nb : 0, hi = ld < na ? ld : na;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sA[mid] <= sB[ld - 1 - mid]) lo = mid + 1; else hi = mid;
      }
      ai = lo;
    }
    int bi = ld - ai;
    // depth at the last position boundary at or before this thread's first element
    int32_t prev_depth = 0;
    bool active = ld < len;
    if (active) {
      K xn = sA[ai] < sB[bi] ? sA[ai] : sB[bi];
      if (ai >= na) xn = sB[bi]; else if (bi >= nb) xn = sA[ai];
      bool has_prev = (i0 + ai) > 0 || (j0 + bi) > 0;
      K xp = 0;
      if (has_prev) {
        bool pa = (i0 + ai) > 0, pb = (j0 + bi) > 0;
        K va = pa ? sA[ai - 1] : 0, vb = pb ? sB[bi - 1] : 0;
        xp = (pa && pb) ? (va > vb ? va : vb) : (pa ? va : vb);
      }
      if (!has_prev) prev_depth = 0;
      else if (xp != xn) prev_depth = (int32_t)((i0 + ai) - (j0 + bi));
      else {
        // the chunk starts inside a group of equal positions: depth before the group = #(A < x) - #(B < x)
        int la = 0, ha = ai;
        while (la < ha) { int mid = (la + ha) >> 1; if (sA[mid] < xn) la = mid + 1; else ha = mid; }
        int lb = 0, hb = bi;
        while (lb < hb) { int mid = (lb + hb) >> 1; if (sB[mid] < xn) lb = mid + 1; else hb = mid; }
        int64_t ga = (la > 0 || i0 == 0) ? i0 + la : lower_bound_dev<K>(A, i0, xn);
        int64_t gb = (lb > 0 || j0 == 0) ? j0 + lb : lower_bound_dev<K>(Bv, j0, xn);
        prev_depth = (int32_t)(ga - gb);
      }
    }
    K ev_x[MP_ITEMS];
    int8_t ev_kind[MP_ITEMS];      // 1 = open, 2 = close
    int32_t ev_max[MP_ITEMS];      // at a close / chunk end: max depth seen in the run part owned by this thread
    int nev = 0, nopen = 0;
    bool prev_in = false;
    int32_t run_max = 0;
    if (active) {
      K xfirst = (ai < na && (bi >= nb || sA[ai] <= sB[bi])) ? sA[ai] : sB[bi];
      int g = group_of_lin<K>(xfirst, G, cfg.n_classes);
      prev_in = prev_depth >= cfg.mn[g] && prev_depth <= cfg.mx[g];
    }
    bool in_at_start = prev_in;
#pragma unroll
    for (int s = 0; s < MP_ITEMS; ++s) {
      if (ld + s < len) {
        bool takeA = (bi >= nb) || (ai < na && sA[ai] <= sB[bi]);
        K x = takeA ? sA[ai] : sB[bi];
        if (takeA) ++ai; else ++bi;
        K nxt = sA[ai] < sB[bi] ? sA[ai] : sB[bi];
        if (nxt != x) {
          int32_t d = (int32_t)((i0 + ai) - (j0 + bi));
          int g = group_of_lin<K>(x, G, cfg.n_classes);
          bool cur_in = d >= cfg.mn[g] && d <= cfg.mx[g];
          if (cur_in) { if (!prev_in) { ev_x[nev] = x; ev_kind[nev] = 1; ev_max[nev] = 0; ++nev; ++nopen; run_max = d; } else if (d > run_max) run_max = d; }
          else if (prev_in) { ev_x[nev] = x; ev_kind[nev] = 2; ev_max[nev] = run_max; ++nev; run_max = 0; }
          prev_in = cur_in;
        }
      }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = nopen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int q = 0; q < MP_THREADS / 32; ++q) { if (q < w) base += warp_tot[q]; tot += warp_tot[q]; }
    if (!WRITE) {
      if (threadIdx.x == 0) tile_cnt[tile] = tot;
    } else {
      int64_t opens_before = tile_off[tile] + base + incl - nopen;   // opens strictly before this thread's chunk
      int64_t cur_run = opens_before - 1;                              // run open at chunk start (valid if in_at_start)
      (void)in_at_start;
      for (int s = 0; s < nev; ++s) {
        if (ev_kind[s] == 1) { ++cur_run; rs[cur_run] = ev_x[s]; }
        else { re[cur_run] = ev_x[s]; if (ev_max[s] > 0) atomicMax(&racc[cur_run], ev_max[s]); }
      }
      if (prev_in && run_max > 0) atomicMax(&racc[cur_run], run_max);   // run continues past this chunk
    }
    __syncthreads();
  }
}

// group of a linear coordinate
template <typename K>
__device__ __forceinline__ int group_of_lin(K x, uint64_t G, int n_classes) { return (int)(((uint64_t)x / G) / (uint64_t)n_classes); }

// ---- COVER / FLAT run detection over the point list --------------------------------------------------------------------
template <typename K>
__global__ void k_cover_open_flags(const K* __restrict__ px, const int32_t* __restrict__ pd, int64_t np, uint64_t G, CoverCfg cfg, int32_t* __restrict__ open) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x) {
    int g = group_of_lin<K>(px[m], G, cfg.n_classes);
    int d = pd[m];
    bool inr = d >= cfg.mn[g] && d <= cfg.mx[g];
    bool prev = false;
    if (m > 0) { int gp = group_of_lin<K>(px[m - 1], G, cfg.n_classes); int dp = pd[m - 1]; prev = dp >= cfg.mn[gp] && dp <= cfg.mx[gp]; }
    open[m] = (inr && !prev) ? 1 : 0;
  }
}

template <typename K>
__global__ void k_cover_runs(const K* __restrict__ px, const int32_t* __restrict__ pd, int64_t np, uint64_t G, CoverCfg cfg, const int64_t* __restrict__ open_excl,
                             K* __restrict__ rs, K* __restrict__ re, int32_t* __restrict__ racc) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x) {
    int g = group_of_lin<K>(px[m], G, cfg.n_classes);
    int d = pd[m];
    bool inr = d >= cfg.mn[g] && d <= cfg.mx[g];
    bool prev = false;
    if (m > 0) { int gp = group_of_lin<K>(px[m - 1], G, cfg.n_classes); int dp = pd[m - 1]; prev = dp >= cfg.mn[gp] && dp <= cfg.mx[gp]; }
    int64_t before = open_excl[m];
    if (inr) {
      int64_t run = prev ? before - 1 : before;
      if (!prev) rs[run] = px[m];
      atomicMax(&racc[run], d);            // AccIndex = max depth in the run (coverHelper countMax, :183-192)
    } else if (prev) {
      re[before - 1] = px[m];
    }
  }
}

// ---- HISTOGRAM segments ----------------------------------------------------------------------------------------------------
template <typename K>
__global__ void k_hist_flags(const K* __restrict__ px, const int32_t* __restrict__ pd, int64_t np, uint64_t G, CoverCfg cfg, int32_t* __restrict__ emit) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x) {
    int d = pd[m];
    int dp = m > 0 ? pd[m - 1] : 0;
    int g = group_of_lin<K>(px[m], G, cfg.n_classes);
    // a segment starts at every point that changes the depth to an accepted, non-zero value (histogramHelper :241-250)
    emit[m] = (d != dp && d != 0 && d >= cfg.mn[g] && d <= cfg.mx[g]) ? 1 : 0;
  }
}

template <typename K>
__global__ void k_hist_segments(const K* __restrict__ px, const int32_t* __restrict__ pd, int64_t np, const int32_t* __restrict__ emit, const int64_t* __restrict__ emit_excl,
                                K* __restrict__ rs, K* __restrict__ re, int32_t* __restrict__ racc) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x) {
    if (!emit[m]) continue;
    int d = pd[m];
    int64_t q = m + 1;
    while (q < np && pd[q] == d) ++q;      // zero-net points do not cut a segment (newCount != count, :247)
    int64_t o = emit_excl[m];
    rs[o] = px[m];
    re[o] = q < np ? px[q] : px[m];
    racc[o] = d;
  }
}

// linear -> (class, chrom, pos)
template <typename K>
__global__ void k_lin_to_regions(const K* __restrict__ rs, const K* __restrict__ re, int64_t n, LinMap lm, uint32_t* __restrict__ ocls, int32_t* __restrict__ ochrom,
                                 int64_t* __restrict__ ostart, int64_t* __restrict__ ostop) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)rs[k];
    uint64_t cls = x / lm.G;
    uint64_t r = x - cls * lm.G;
    int lo = 0, hi = lm.n_chrom;           // last chrom with coff <= r
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (lm.coff[mid] <= r) lo = mid; else hi = mid; }
    ocls[k] = (uint32_t)cls;
    ochrom[k] = lo;
    ostart[k] = (int64_t)(r - lm.coff[lo]);
    ostop[k] = (int64_t)((uint64_t)re[k] - cls * lm.G - lm.coff[lo]);
  }
}

// ---- SUMMIT: per-bin emulation ---------------------------------------------------------------------------------------------
// bin-space linearisation: lin2(class, chrom, bin, off) = class*G2 + coff2[chrom] + bin*(B+1) + off, off in [0, B]
__global__ void k_summit_offsets(const unsigned long long* __restrict__ span, int n_chrom, int64_t B, uint64_t* __restrict__ coff2) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint64_t run = 0;
    for (int c = 0; c < n_chrom; ++c) { coff2[c] = run; run += (span[c] / (uint64_t)B + 2ull) * (uint64_t)(B + 1); }
    coff2[n_chrom] = run;
  }
}

__global__ void k_summit_event_counts(const void* __restrict__ start, const void* __restrict__ stop, int bits, int64_t n, int64_t B, int64_t* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    int64_t sb = s / B, eb = e / B;
    cnt[i] = (s == e) ? 1 : 2 * (eb - sb + 1);
  }
}

// event key = lin2 << 1 | (delta > 0)
__global__ void k_summit_events(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ cls,
                                int64_t n, int64_t B, const uint64_t* __restrict__ coff2, uint64_t G2, const int64_t* __restrict__ off, uint64_t* __restrict__ keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
    int64_t sb = s / B, eb = e / B;
    uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * G2 + coff2[chrom[i]];
    uint64_t W = (uint64_t)(B + 1);
    int64_t o = off[i];
    if (s == e) {                                        // only the -1 survives (:351)
      keys[o] = ((base + (uint64_t)sb * W + (uint64_t)(s - sb * B)) << 1);
    } else if (sb == eb) {
      keys[o] = ((base + (uint64_t)sb * W + (uint64_t)(s - sb * B)) << 1) | 1ull;
      keys[o + 1] = ((base + (uint64_t)sb * W + (uint64_t)(e - sb * B)) << 1);
    } else {
      keys[o++] = ((base + (uint64_t)sb * W + (uint64_t)(s - sb * B)) << 1) | 1ull;   // map_start (:353)
      keys[o++] = ((base + (uint64_t)sb * W + (uint64_t)B) << 1);
      for (int64_t b = sb + 1; b < eb; ++b) {                                           // map_int (:355)
        keys[o++] = ((base + (uint64_t)b * W) << 1) | 1ull;
        keys[o++] = ((base + (uint64_t)b * W + (uint64_t)B) << 1);
      }
      int64_t eo = e - eb * B;
      keys[o++] = ((base + (uint64_t)eb * W) << 1) | 1ull;                              // map_stop (:354)
      keys[o++] = ((base + (uint64_t)eb * W + (uint64_t)(eo != 0 ? eo : 1)) << 1);
    }
  }
}

__global__ void k_summit_heads(const uint64_t* __restrict__ keys, int64_t ne, int32_t* __restrict__ head) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < ne; m += (int64_t)gridDim.x * blockDim.x)
    head[m] = (m == 0 || (keys[m] >> 1) != (keys[m - 1] >> 1)) ? 1 : 0;
}

// coalesce coincident events (reduceByKey :98-108): one point per distinct lin2, zero-net points kept
__global__ void k_summit_points(const uint64_t* __restrict__ keys, int64_t ne, const int32_t* __restrict__ head, const int64_t* __restrict__ head_excl,
                                uint64_t* __restrict__ pkey, int32_t* __restrict__ pdelta) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < ne; m += (int64_t)gridDim.x * blockDim.x) {
    if (!head[m]) continue;
    uint64_t k = keys[m] >> 1;
    int net = 0;
    for (int64_t q = m; q < ne && (keys[q] >> 1) == k; ++q) net += (keys[q] & 1ull) ? 1 : -1;
    int64_t o = head_excl[m];
    pkey[o] = k;
    pdelta[o] = net;
  }
}

// first point of every bin -> compact list (so that the state machines below run with every lane busy)
__global__ void k_summit_bin_heads(const uint64_t* __restrict__ pkey, int64_t np, int64_t B, int32_t* __restrict__ head) {
  const uint64_t W = (uint64_t)(B + 1);
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x)
    head[m] = (m == 0 || pkey[m - 1] / W != pkey[m] / W) ? 1 : 0;
}
__global__ void k_summit_bin_list(const int32_t* __restrict__ head, const int64_t* __restrict__ head_excl, int64_t np, uint32_t* __restrict__ bin_first) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < np; m += (int64_t)gridDim.x * blockDim.x)
    if (head[m]) bin_first[head_excl[m]] = (uint32_t)m;
}

// dense inputs: nearly every bin of the genome holds points, so the first point of EVERY bin (empty ones included) comes from
// one binary search per bin instead of head flags + scan + compaction over all points
__global__ void k_summit_bin_search(const uint64_t* __restrict__ pkey, int64_t np, uint64_t W, int64_t n_bins, uint32_t* __restrict__ bin_first) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x)
    bin_first[b] = (uint32_t)lower_bound_dev<uint64_t>(pkey, np, (uint64_t)b * W);
}

// one thread per bin: the reference's per-bin state machine over the bin's coalesced points — summitHelper (:269-316),
// coverHelper (:172-218, also FLAT) or histogramHelper (:229-258), literally.  A piece is stored at the index of the point that
// emits it (its stop is that point's offset, clipped at the bin size by k_summit_pieces).  SUMMIT always runs here (every bin
// restarts at depth 0); COVER / FLAT / HISTOGRAM only when minAcc < 1 makes depth 0 "in range": the reference's result then
// depends on which bins hold events at all (:57-79 with the bins of :98-118), which only the bin-wise machine reproduces.
enum { BINS_SUMMIT = 0, BINS_COVER = 1, BINS_HISTOGRAM = 2 };
// A thread per bin walks its own ~60 points: neighbouring lanes read addresses a whole bin apart, every 8-byte key costs a
// 32-byte sector and the working set of a full SM (1.4 MB) evicts the sectors before their reuse — ncu: 4.1 GB read for 1.1 GB of
// points, issue 4.6 % (profiles/r02z_ncu_c2_summary.txt).  So a block of BINS_T consecutive bins first brings ITS points — one
// contiguous range — into shared memory with coalesced loads (key as a 32-bit offset from the range's first key, net change), the
// threads run their state machines out of shared memory, leave the emit flags in the net-change slots, and the block writes
// them back coalesced.  A range beyond the stage (or wider than 2^32 keys: a chromosome / class boundary) is read in place.
constexpr int BINS_T = 128;
constexpr int BINS_CAP = 11264;                            // staged points per block (88 KB: two blocks per SM)
__global__ void __launch_bounds__(BINS_T) k_summit_bins(const uint64_t* __restrict__ pkey, const int32_t* __restrict__ pdelta, int64_t np, const uint32_t* __restrict__ bin_first, int64_t n_bins,
                              int64_t B, uint64_t G2, CoverCfg cfg, int kind, int cap, int32_t* __restrict__ emit, int64_t* __restrict__ pstart_off, int32_t* __restrict__ pval) {
  extern __shared__ uint32_t bins_sm[];
  uint32_t* s_k = bins_sm;                                 // [BINS_CAP] key - first key of the range
  int32_t* s_d = (int32_t*)(bins_sm + BINS_CAP);           // [BINS_CAP] net change, then the emit flag
  const uint64_t W = (uint64_t)(B + 1);
  const int tid = threadIdx.x;
  for (int64_t b0 = (int64_t)blockIdx.x * BINS_T; b0 < n_bins; b0 += (int64_t)gridDim.x * BINS_T) {
    const int64_t b1 = b0 + BINS_T < n_bins ? b0 + BINS_T : n_bins;
    const int64_t m0 = bin_first[b0], m1 = b1 < n_bins ? (int64_t)bin_first[b1] : np;
    if (m1 == m0) continue;                                // (a stretch of empty bins of the dense table; block-uniform)
    const uint64_t key0 = pkey[m0];
    const bool staged = m1 - m0 <= cap && pkey[m1 - 1] - key0 < 0xffffffffull;   // block-uniform
    if (staged) {
      for (int64_t i = tid; i < m1 - m0; i += BINS_T) { s_k[i] = (uint32_t)(pkey[m0 + i] - key0); s_d[i] = pdelta[m0 + i]; }
    }
    __syncthreads();
    const int64_t bi = b0 + tid;
    if (bi < b1) {
    const int64_t m = bin_first[bi];
    const int64_t m_end = bi + 1 < n_bins ? (int64_t)bin_first[bi + 1] : np;
    if (m < m_end) {
    auto key_at = [&](int64_t q) -> uint64_t { return staged ? key0 + (uint64_t)s_k[q - m0] : pkey[q]; };
    auto delta_at = [&](int64_t q) -> int { return staged ? s_d[q - m0] : pdelta[q]; };
    auto put = [&](int64_t q, bool out) { if (staged) s_d[q - m0] = out ? 1 : 0; else emit[q] = out ? 1 : 0; };
    const uint64_t bin = key_at(m) / W;
    const uint64_t bin_base = bin * W;
    int g = (int)((key_at(m) / G2) / (uint64_t)cfg.n_classes);
    int mn = cfg.mn[g], mx = cfg.mx[g];
    int64_t start = 0;
    int count = 0;
    if (kind == BINS_SUMMIT) {
      bool valid = false, reentered = false, growing = false;
      for (int64_t q = m; q < m_end; ++q) {
        int64_t off = (int64_t)(key_at(q) - bin_base);
        int nc = count + delta_at(q);
        bool nvalid = nc >= mn && nc <= mx;
        bool ngrow = nc > count || (growing && nc == count);
        bool nre = !valid && nvalid;
        int64_t nstart = (nvalid && nc != count) ? off : start;
        bool out = ((valid && growing && (!ngrow || !nvalid)) || (reentered && !ngrow)) && count != 0;
        put(q, out);
        if (out) { pstart_off[q] = start; pval[q] = count; }
        start = nstart; count = nc; valid = nvalid; reentered = nre; growing = ngrow;
      }
    } else if (kind == BINS_COVER) {
      int count_max = 0;
      bool recording = false;
      for (int64_t q = m; q < m_end; ++q) {
        const int64_t off = (int64_t)(key_at(q) - bin_base);
        const int nc = count + delta_at(q);
        const bool in_range = nc >= mn && nc <= mx;
        const bool nrec = in_range && q + 1 < m_end;       // `si.hasNext` (:180): the bin's last point never records
        const int ncmax = (!nrec && recording) ? 0 : ((nrec && nc > count_max) ? nc : count_max);
        const int64_t nstart = (in_range && !recording) ? off : start;
        const bool out = !nrec && recording;
        put(q, out);
        if (out) { pstart_off[q] = nstart; pval[q] = count_max; }
        count = nc; start = nstart; count_max = ncmax; recording = nrec;
      }
    } else {
      for (int64_t q = m; q < m_end; ++q) {
        const int64_t off = (int64_t)(key_at(q) - bin_base);
        const int nc = count + delta_at(q);
        const int64_t nstart = (nc >= mn && nc <= mx && nc != count) ? off : start;
        const bool out = count >= mn && count <= mx && nc != count && count != 0;
        put(q, out);
        if (out) { pstart_off[q] = start; pval[q] = count; }
        start = nstart; count = nc;
      }
    }
    }
    }
    __syncthreads();
    if (staged) for (int64_t i = tid; i < m1 - m0; i += BINS_T) emit[m0 + i] = s_d[i];
    __syncthreads();                                       // the stage is rewritten by the next range
  }
}

// compact emitted pieces into (class, chrom, start, stop, value) in genome order
__global__ void k_summit_pieces(const uint64_t* __restrict__ pkey, int64_t np, int64_t B, uint64_t G2, const uint64_t* __restrict__ coff2, int n_chrom,
                                const int32_t* __restrict__ emit, const int64_t* __restrict__ emit_excl, const int64_t* __restrict__ pstart_off, const int32_t* __restrict__ pval,
                                uint32_t* __restrict__ ocls, int32_t* __restrict__ ochrom, int64_t* __restrict__ ostart, int64_t* __restrict__ ostop, int32_t* __restrict__ oval) {
  const uint64_t W = (uint64_t)(B + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (int64_t)gridDim.x * blockDim.x) {
    if (!emit[q]) continue;
    uint64_t x = pkey[q];
    uint64_t cls = x / G2, r = x - cls * G2;
    int lo = 0, hi = n_chrom;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (coff2[mid] <= r) lo = mid; else hi = mid; }
    uint64_t inchr = r - coff2[lo];
    int64_t bin = (int64_t)(inchr / W), off = (int64_t)(inchr % W);
    int64_t o = emit_excl[q];
    ocls[o] = (uint32_t)cls;
    ochrom[o] = lo;
    ostart[o] = pstart_off[q] + bin * B;
    ostop[o] = (off > B ? B : off) + bin * B;
    oval[o] = pval[q];
  }
}

// cross-bin merge (GenometricCover.scala:121-152): pieces touching a bin edge chain when prev.stop == cur.start
__global__ void k_chain_heads(const uint32_t* __restrict__ cls, const int32_t* __restrict__ chrom, const int64_t* __restrict__ s, const int64_t* __restrict__ e, const int32_t* __restrict__ v, int eq_vals,
                              int64_t n, int64_t B, int32_t* __restrict__ head) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    bool merge = false;
    if (k > 0) {
      bool nv = (s[k] % B == 0) || (e[k] % B == 0);
      bool nvp = (s[k - 1] % B == 0) || (e[k - 1] % B == 0);
      merge = nv && nvp && cls[k] == cls[k - 1] && chrom[k] == chrom[k - 1] && e[k - 1] == s[k] && (!eq_vals || v[k - 1] == v[k]);   // HISTOGRAM: equal depths only (:139)
    }
    head[k] = merge ? 0 : 1;
  }
}
__global__ void k_chain_merge(const uint32_t* __restrict__ cls, const int32_t* __restrict__ chrom, const int64_t* __restrict__ s, const int64_t* __restrict__ e, const int32_t* __restrict__ v, int64_t n,
                              const int32_t* __restrict__ head, const int64_t* __restrict__ head_excl,
                              uint32_t* __restrict__ ocls, int32_t* __restrict__ ochrom, int64_t* __restrict__ os, int64_t* __restrict__ oe, int32_t* __restrict__ ov) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t o = head[k] ? head_excl[k] : head_excl[k] - 1;
    if (head[k]) { ocls[o] = cls[k]; ochrom[o] = chrom[k]; os[o] = s[k]; }
    if (k + 1 == n || head[k + 1]) oe[o] = e[k];
    atomicMax(&ov[o], v[k]);
  }
}

// ---- phase 2: GMAP4 ------------------------------------------------------------------------------------------------------------
struct CovAggPlan {
  int n_cols = 0;
  int col_id[8];
  int n_aggs = 0;
  int kind[16], slot[16];
  int need_bag = 0;      // some aggregate is BAG / BAGD: the contributors of every output region are listed (GMQL_COL_BAG_*)
};
struct CovCols { const double* v[8]; const uint8_t* ok[8]; };
struct CovOut { double* value[16]; uint8_t* valid[16]; };

__device__ __forceinline__ unsigned long long cov_key_min(double v) {
  if (v != v) return 0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long cov_key_max(double v) {
  if (v != v) return ~0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double cov_from_key(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// one warp per phase-1 region
template <typename K>
__global__ void __launch_bounds__(256) k_gmap4(const uint32_t* __restrict__ ccls, const int32_t* __restrict__ cchrom, const int64_t* __restrict__ cstart, const int64_t* __restrict__ cstop, int64_t nc,
                                               LinMap lm, const K* __restrict__ xs, const K* __restrict__ xe, const K* __restrict__ xpmax, const uint32_t* __restrict__ xrow, int64_t nx,
                                               CovAggPlan plan, CovCols cc, int flat,
                                               int32_t* __restrict__ keep, int64_t* __restrict__ fstart, int64_t* __restrict__ fstop, double* __restrict__ j1, double* __restrict__ j2,
                                               int32_t* __restrict__ ocount, CovOut out, const int* __restrict__ only_if /* null, or run only when *only_if != 0 */) {
  if (only_if && !*only_if) return;
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < nc; c += nwarps) {
    uint64_t base = (uint64_t)ccls[c] * lm.G + lm.coff[cchrom[c]];
    K cs = (K)(base + (uint64_t)cstart[c]), ce = (K)(base + (uint64_t)cstop[c]);
    int64_t hi = lower_bound_dev<K>(xs, nx, ce);            // inputs with start < c.stop
    int64_t lo = upper_bound_dev<K>(xpmax, hi, cs);          // first index whose running max stop exceeds c.start
    long long cnt = 0;
    unsigned long long smin = ~0ull, emax = 0ull, smax = 0ull, emin = ~0ull;
    int nn[8];
    double sum[8];
    unsigned long long vmin[8], vmax[8];
    for (int q = 0; q < plan.n_cols; ++q) { nn[q] = 0; sum[q] = 0.0; vmin[q] = ~0ull; vmax[q] = 0ull; }
    for (int64_t k = lo + lane; k < hi; k += 32) {
      K e = xe[k];
      if (e <= cs) continue;                                 // strict overlap (GMAP4.scala:41)
      K s = xs[k];
      ++cnt;
      if ((unsigned long long)s < smin) smin = s;
      if ((unsigned long long)e > emax) emax = e;
      if ((unsigned long long)s > smax) smax = s;
      if ((unsigned long long)e < emin) emin = e;
      if (plan.n_cols) {
        uint32_t row = xrow[k];
        for (int q = 0; q < plan.n_cols; ++q) {
          if (cc.ok[q] && !cc.ok[q][row]) continue;
          double v = cc.v[q][row];
          nn[q]++; sum[q] += v;
          unsigned long long a = cov_key_min(v), b = cov_key_max(v);
          if (a < vmin[q]) vmin[q] = a;
          if (b > vmax[q]) vmax[q] = b;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      unsigned long long t;
      t = __shfl_xor_sync(0xffffffffu, smin, o); if (t < smin) smin = t;
      t = __shfl_xor_sync(0xffffffffu, emax, o); if (t > emax) emax = t;
      t = __shfl_xor_sync(0xffffffffu, smax, o); if (t > smax) smax = t;
      t = __shfl_xor_sync(0xffffffffu, emin, o); if (t < emin) emin = t;
    }
    for (int q = 0; q < plan.n_cols; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        nn[q] += __shfl_xor_sync(0xffffffffu, nn[q], o);
        sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], o);
        unsigned long long t;
        t = __shfl_xor_sync(0xffffffffu, vmin[q], o); if (t < vmin[q]) vmin[q] = t;
        t = __shfl_xor_sync(0xffffffffu, vmax[q], o); if (t > vmax[q]) vmax[q] = t;
      }
    }
    if (lane == 0) {
      keep[c] = cnt > 0 ? 1 : 0;                             // dropped when nothing overlaps (:33-34,48)
      ocount[c] = (int32_t)cnt;
      if (cnt > 0) {
        int64_t start_min = (int64_t)(smin - base), end_max = (int64_t)(emax - base);
        int64_t os = flat ? start_min : cstart[c], oe = flat ? end_max : cstop[c];   // :83-84
        fstart[c] = os; fstop[c] = oe;
        long long den = (long long)(emax - smin);
        double num2;
        if (cnt == 1) num2 = (double)(long long)(emin - smax);                        // single contributor: raw (start, stop)
        else num2 = smax < emin ? (double)(long long)(emin - smax) : 0.0;            // (0,0) once the intersection is empty (:74)
        j1[c] = den != 0 ? fabs(((double)oe - (double)os) / (double)den) : 0.0;      // :91
        j2[c] = den != 0 ? fabs(num2 / (double)den) : 0.0;                           // :93
      } else { fstart[c] = cstart[c]; fstop[c] = cstop[c]; j1[c] = 0.0; j2[c] = 0.0; }
      for (int a = 0; a < plan.n_aggs; ++a) {
        int kind = plan.kind[a], q = plan.slot[a];
        double v = 0.0; uint8_t ok = 0;
        if (kind == GMQL_AGG_COUNT) { v = (double)cnt; ok = 1; }
        else if (q >= 0 && nn[q] > 0) {
          ok = 1;
          if (kind == GMQL_AGG_SUM) v = sum[q];
          else if (kind == GMQL_AGG_AVG) v = sum[q] / (double)nn[q];
          else if (kind == GMQL_AGG_MIN) v = cov_from_key(vmin[q]);
          else if (kind == GMQL_AGG_MAX) v = cov_from_key(vmax[q]);
          else if (kind == GMQL_AGG_MEDIAN) { double m = cov_from_key(vmin[q]); v = nn[q] >= 2 ? (m + m) / 2 : m; }
        }
        out.value[a][c] = v;
        out.valid[a][c] = ok;
      }
    }
  }
}

// ---- GMAP4, region-driven: for many small phase-1 regions (HISTOGRAM segments, SUMMIT plateaus) ------------------------
// The phase-1 regions of every flag are sorted by linear start and pairwise disjoint, so their stops are sorted too.  Instead
// of one warp searching the input index per phase-1 region, every INPUT region finds the phase-1 regions it overlaps (bucket
// table over their starts, then a backward walk that ends at the first non-overlap) and accumulates with REDs.
__device__ __forceinline__ void amin_k(uint32_t* p, uint32_t v) { atomicMin(p, v); }
__device__ __forceinline__ void amax_k(uint32_t* p, uint32_t v) { atomicMax(p, v); }
__device__ __forceinline__ void amin_k(uint64_t* p, uint64_t v) { atomicMin((unsigned long long*)p, (unsigned long long)v); }
__device__ __forceinline__ void amax_k(uint64_t* p, uint64_t v) { atomicMax((unsigned long long*)p, (unsigned long long)v); }

template <typename K>
struct ScatterAcc {
  int32_t* cnt; K* smin; K* emax; K* smax; K* emin;
  int32_t* nn[8]; double* sum[8]; unsigned long long* vmin[8]; unsigned long long* vmax[8];
};

template <typename K>
__global__ void k_p1_linear(const uint32_t* __restrict__ ccls, const int32_t* __restrict__ cchrom, const int64_t* __restrict__ cstart, const int64_t* __restrict__ cstop, int64_t nc,
                            LinMap lm, K* __restrict__ cs, K* __restrict__ ce, int* __restrict__ unsorted) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    uint64_t base = (uint64_t)ccls[c] * lm.G + lm.coff[cchrom[c]];
    K s = (K)(base + (uint64_t)cstart[c]), e = (K)(base + (uint64_t)cstop[c]);
    cs[c] = s; ce[c] = e;
    if (c > 0) {   // sorted and disjoint: the previous region ends at or before this start
      uint64_t pb = (uint64_t)ccls[c - 1] * lm.G + lm.coff[cchrom[c - 1]];
      if ((K)(pb + (uint64_t)cstop[c - 1]) > s) *unsorted = 1;
    }
  }
}

template <typename K>
__global__ void __launch_bounds__(256) k_gmap4_scatter(const K* __restrict__ xs, const K* __restrict__ xe, const uint32_t* __restrict__ xrow, int64_t nx,
                                                       const K* __restrict__ cs, const K* __restrict__ ce, int64_t nc, const uint32_t* __restrict__ tbl, int tbl_shift, int64_t n_buckets,
                                                       CovAggPlan plan, CovCols cc, ScatterAcc<K> acc, const int* __restrict__ unsorted) {
  if (*unsorted) return;                                   // the warp-per-region kernel takes over
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nx; k += (int64_t)gridDim.x * blockDim.x) {
    const K s = xs[k], e = xe[k];
    // last = number of phase-1 regions with start < e
    int64_t lo, hi;
    {
      uint64_t b = (uint64_t)e >> tbl_shift;
      if (b >= (uint64_t)n_buckets) { lo = hi = nc; }
      else { lo = tbl[b]; hi = tbl[b + 1]; }
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (cs[mid] < e) lo = mid + 1; else hi = mid; }
    }
    uint32_t row = 0;
    if (plan.n_cols) row = xrow[k];
    for (int64_t j = lo - 1; j >= 0 && ce[j] > s; --j) {     // strict overlap (GMAP4.scala:41); stops are sorted
      atomicAdd(&acc.cnt[j], 1);
      amin_k(&acc.smin[j], s); amax_k(&acc.emax[j], e);
      amax_k(&acc.smax[j], s); amin_k(&acc.emin[j], e);
      for (int q = 0; q < plan.n_cols; ++q) {
        if (cc.ok[q] && !cc.ok[q][row]) continue;
        double v = cc.v[q][row];
        atomicAdd(&acc.nn[q][j], 1);
        atomicAdd(&acc.sum[q][j], v);
        atomicMin(&acc.vmin[q][j], cov_key_min(v));
        atomicMax(&acc.vmax[q][j], cov_key_max(v));
      }
    }
  }
}

// ---- GMAP4, merge form: count and the four bounds without atomics (no value aggregates) ---------------------------------------
// The scatter above issues five REDs per (input region, phase-1 region) pair — 1.1e9 for C2's HISTOGRAM (96 M segments), 7 ms.
// Both lists are sorted, so a block takes a CHUNK of consecutive phase-1 regions and stages the slice of the input index they can
// see — from the first entry whose running max stop exceeds the chunk's first start to the last entry starting before the chunk's
// last stop — in shared memory (sorted starts, stops, running maxima); every thread then finds its own region's window inside
// the slice by two binary searches in shared memory (inputs with start < c.stop; first index whose running max exceeds c.start:
// exactly k_gmap4's bounds) and scans it: a handful of entries.  Plain stores, nothing to initialise; k_gmap4_finish reads the
// same accumulators.  A slice larger than the stage (long regions, deep pile-ups) is read from global memory by the same code.
constexpr int GM_T = 256, GM_PER = 8, GM_CHUNK = GM_T * GM_PER;
template <typename K>
__global__ void __launch_bounds__(GM_T) k_gmap4_merge(const K* __restrict__ xs, const K* __restrict__ xe, const K* __restrict__ xpmax, int64_t nx,
                                                      const K* __restrict__ cs, const K* __restrict__ ce, int64_t nc, ScatterAcc<K> acc, const int* __restrict__ unsorted, int* __restrict__ check) {
  // check (optional): the phase-1 lists come straight from the sweep, sorted and disjoint by construction; a violation is raised
  // here and reported as an internal error by the host, never computed around
  if (*unsorted) return;                                   // the warp-per-region kernel takes over
  constexpr int CAP = sizeof(K) == 4 ? 3072 : 1536;
  __shared__ K s_xs[CAP], s_xe[CAP], s_pm[CAP];
  __shared__ long long s_lo, s_hi;
  const int tid = threadIdx.x;
  for (int64_t c0 = (int64_t)blockIdx.x * GM_CHUNK; c0 < nc; c0 += (int64_t)gridDim.x * GM_CHUNK) {
    const int64_t c1 = c0 + GM_CHUNK < nc ? c0 + GM_CHUNK : nc;
    if (tid == 0) s_hi = lower_bound_dev<K>(xs, nx, ce[c1 - 1]);        // stops ascend: nothing of the chunk sees an entry from here on
    if (tid == 32) s_lo = upper_bound_dev<K>(xpmax, nx, cs[c0]);         // starts ascend: nothing of the chunk sees an entry before this one
    __syncthreads();
    const int64_t LO = s_lo, HI = s_hi > s_lo ? s_hi : s_lo;
    const int64_t m = HI - LO;
    const bool staged = m <= CAP;
    if (staged) for (int64_t i = tid; i < m; i += GM_T) { s_xs[i] = xs[LO + i]; s_xe[i] = xe[LO + i]; s_pm[i] = xpmax[LO + i]; }
    __syncthreads();
    const K* S = staged ? s_xs : xs + LO;
    const K* E = staged ? s_xe : xe + LO;
    const K* P = staged ? s_pm : xpmax + LO;
#pragma unroll 1
    for (int r = 0; r < GM_PER; ++r) {
      const int64_t c = c0 + (int64_t)r * GM_T + tid;
      if (c >= c1) break;
      const K a = cs[c], b = ce[c];
      if (check && c > 0 && ce[c - 1] > a) *check = 1;
      int64_t lo = 0, hi = m;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (S[mid] < b) lo = mid + 1; else hi = mid; }
      const int64_t kb = lo;                                 // inputs (of the slice) with start < c.stop
      lo = 0; hi = kb;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (P[mid] <= a) lo = mid + 1; else hi = mid; }
      int32_t cnt = 0;
      K smin = ~(K)0, emax = 0, smax = 0, emin = ~(K)0;
      for (int64_t k = lo; k < kb; ++k) {
        const K e = E[k];
        if (e <= a) continue;                                // strict overlap (GMAP4.scala:41)
        const K st = S[k];
        ++cnt;
        smin = st < smin ? st : smin; emax = e > emax ? e : emax;
        smax = st > smax ? st : smax; emin = e < emin ? e : emin;
      }
      acc.cnt[c] = cnt; acc.smin[c] = smin; acc.emax[c] = emax; acc.smax[c] = smax; acc.emin[c] = emin;
    }
    __syncthreads();                                         // the stage is rewritten by the next chunk
  }
}

template <typename K>
__global__ void k_p1_table(const K* __restrict__ cs, int64_t nc, int shift, int64_t n_buckets, uint32_t* __restrict__ tbl) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)b << shift;
    K kx = (x > (uint64_t)(~(K)0)) ? ~(K)0 : (K)x;
    tbl[b] = (uint32_t)lower_bound_dev<K>(cs, nc, kx);
  }
}

template <typename K>
__global__ void k_gmap4_finish(const uint32_t* __restrict__ ccls, const int32_t* __restrict__ cchrom, const int64_t* __restrict__ cstart, const int64_t* __restrict__ cstop, int64_t nc,
                               LinMap lm, CovAggPlan plan, ScatterAcc<K> acc, int flat, int32_t* __restrict__ keep, int64_t* __restrict__ fstart, int64_t* __restrict__ fstop,
                               double* __restrict__ j1, double* __restrict__ j2, int32_t* __restrict__ ocount, CovOut out, const int* __restrict__ unsorted) {
  if (*unsorted) return;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    uint64_t base = (uint64_t)ccls[c] * lm.G + lm.coff[cchrom[c]];
    long long cnt = acc.cnt[c];
    keep[c] = cnt > 0 ? 1 : 0;
    ocount[c] = (int32_t)cnt;
    if (cnt > 0) {
      unsigned long long smin = acc.smin[c], emax = acc.emax[c], smax = acc.smax[c], emin = acc.emin[c];
      int64_t start_min = (int64_t)(smin - base), end_max = (int64_t)(emax - base);
      int64_t os = flat ? start_min : cstart[c], oe = flat ? end_max : cstop[c];   // GMAP4.scala:83-84
      fstart[c] = os; fstop[c] = oe;
      long long den = (long long)(emax - smin);
      double num2;
      if (cnt == 1) num2 = (double)(long long)(emin - smax);
      else num2 = smax < emin ? (double)(long long)(emin - smax) : 0.0;            // :74
      j1[c] = den != 0 ? fabs(((double)oe - (double)os) / (double)den) : 0.0;      // :91
      j2[c] = den != 0 ? fabs(num2 / (double)den) : 0.0;                           // :93
    } else { fstart[c] = cstart[c]; fstop[c] = cstop[c]; j1[c] = 0.0; j2[c] = 0.0; }
    for (int a = 0; a < plan.n_aggs; ++a) {
      int kind = plan.kind[a], q = plan.slot[a];
      double v = 0.0; uint8_t ok = 0;
      if (kind == GMQL_AGG_COUNT) { v = (double)cnt; ok = 1; }
      else if (q >= 0 && acc.nn[q][c] > 0) {
        ok = 1;
        int nn = acc.nn[q][c];
        if (kind == GMQL_AGG_SUM) v = acc.sum[q][c];
        else if (kind == GMQL_AGG_AVG) v = acc.sum[q][c] / (double)nn;
        else if (kind == GMQL_AGG_MIN) v = cov_from_key(acc.vmin[q][c]);
        else if (kind == GMQL_AGG_MAX) v = cov_from_key(acc.vmax[q][c]);
        else if (kind == GMQL_AGG_MEDIAN) { double m = cov_from_key(acc.vmin[q][c]); v = nn >= 2 ? (m + m) / 2 : m; }
      }
      out.value[a][c] = v;
      out.valid[a][c] = ok;
    }
  }
}

struct FinalCols {
  int32_t* group; int32_t* chrom; int64_t* start; int64_t* stop; int8_t* strand; int32_t* acc; double* j1; double* j2;
  double* value[16]; uint8_t* valid[16];
};
// merge form, second half: keep = count > 0 (GMAP4.scala:33-34,48), and — once the kept rows are numbered — the GMAP4 formulas
// written straight into the compacted result columns: no intermediate start / stop / Jaccard arrays and no compaction pass over them
// (k_gmap4_finish + k_cover_compact moved 206 bytes per phase-1 region for HISTOGRAM's 96 M segments; this moves 109)
__global__ void k_keep_from_cnt(const int32_t* __restrict__ cnt, int64_t nc, int32_t* __restrict__ keep, const int* __restrict__ unsorted) {
  if (*unsorted) return;                                   // k_gmap4 decides
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) keep[c] = cnt[c] > 0 ? 1 : 0;
}
template <typename K>
__global__ void k_gmap4_finish_direct(const int32_t* __restrict__ keep, const int64_t* __restrict__ keep_excl, const uint32_t* __restrict__ ccls, const int32_t* __restrict__ cchrom,
                                      const int64_t* __restrict__ cstart, const int64_t* __restrict__ cstop, const int32_t* __restrict__ cacc, int64_t nc, LinMap lm, CovAggPlan plan,
                                      ScatterAcc<K> acc, int flat, int n_classes, int unified, FinalCols f, double* __restrict__ acc_d,
                                      const K* __restrict__ lin_s, const K* __restrict__ lin_e /* the phase-1 regions in linear form (then ccls .. cstop are null) */) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[c]) continue;
    const int64_t o = keep_excl[c];
    uint32_t cls;
    int chrom;
    int64_t c_start, c_stop;
    if (lin_s) {                                           // linear -> (class, chrom, pos), as k_lin_to_regions
      const uint64_t x = (uint64_t)lin_s[c];
      const uint64_t k = x / lm.G, r = x - k * lm.G;
      int lo = 0, hi = lm.n_chrom;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lm.coff[mid] <= r) lo = mid; else hi = mid; }
      cls = (uint32_t)k; chrom = lo;
      c_start = (int64_t)(r - lm.coff[lo]);
      c_stop = (int64_t)((uint64_t)lin_e[c] - k * lm.G - lm.coff[lo]);
    } else { cls = ccls[c]; chrom = cchrom[c]; c_start = cstart[c]; c_stop = cstop[c]; }
    const uint64_t base = (uint64_t)cls * lm.G + lm.coff[chrom];
    const long long cnt = acc.cnt[c];
    const unsigned long long smin = acc.smin[c], emax = acc.emax[c], smax = acc.smax[c], emin = acc.emin[c];
    const int64_t start_min = (int64_t)(smin - base), end_max = (int64_t)(emax - base);
    const int64_t os = flat ? start_min : c_start, oe = flat ? end_max : c_stop;       // GMAP4.scala:83-84
    const long long den = (long long)(emax - smin);
    double num2;
    if (cnt == 1) num2 = (double)(long long)(emin - smax);
    else num2 = smax < emin ? (double)(long long)(emin - smax) : 0.0;                  // :74
    f.group[o] = (int32_t)(cls / (uint32_t)n_classes);
    f.strand[o] = unified ? (int8_t)0 : (int8_t)(cls % (uint32_t)n_classes + 1);
    f.chrom[o] = chrom;
    f.start[o] = os;
    f.stop[o] = oe;
    f.acc[o] = cacc[c];
    acc_d[o] = (double)cacc[c];
    f.j1[o] = den != 0 ? fabs(((double)oe - (double)os) / (double)den) : 0.0;          // :91
    f.j2[o] = den != 0 ? fabs(num2 / (double)den) : 0.0;                               // :93
    for (int a = 0; a < plan.n_aggs; ++a) {                // no value columns in this form: COUNT, or a host-built BAG string
      const bool is_count = plan.kind[a] == GMQL_AGG_COUNT;
      f.value[a][o] = is_count ? (double)cnt : 0.0;
      f.valid[a][o] = is_count ? 1 : 0;
    }
  }
}

// ---- aggregates for the tile-resident path: a region-driven post-pass over the runs ------------------------------------------
// The tile kernel leaves count / startMin / endMax / startMax / endMin per run; value aggregates are added by one more read of the
// input columns in their ORIGINAL order (values are fetched by row, no gather): every region finds the runs it overlaps through a
// bucket table over the sorted, disjoint run starts and accumulates with REDs.
struct RunAgg { int32_t* nn[8]; double* sum[8]; unsigned long long* vmin[8]; unsigned long long* vmax[8]; };

__global__ void k_run_table(const uint32_t* __restrict__ rstart, const int64_t* __restrict__ n_runs, int shift, int64_t n_buckets, uint32_t* __restrict__ tbl) {
  const int64_t nr = *n_runs;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)b << shift;
    uint32_t kx = x > 0xffffffffull ? 0xffffffffu : (uint32_t)x;
    tbl[b] = (uint32_t)lower_bound_dev<uint32_t>(rstart, nr, kx);
  }
}

__global__ void __launch_bounds__(256) k_run_agg_scatter(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, const uint32_t* __restrict__ cls,
                                                         int64_t n, LinMap lm, const uint32_t* __restrict__ rstart, const uint32_t* __restrict__ rend, const int64_t* __restrict__ n_runs,
                                                         const uint32_t* __restrict__ tbl, int tbl_shift, int64_t n_buckets, CovAggPlan plan, CovCols cc, RunAgg acc) {
  const int64_t nr = *n_runs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t base = (cls ? (uint64_t)cls[i] : 0ull) * lm.G + lm.coff[chrom[i]];
    const uint32_t s = (uint32_t)(base + (uint64_t)ld_coord(start, bits, i)), e = (uint32_t)(base + (uint64_t)ld_coord(stop, bits, i));
    int64_t lo, hi;                                          // runs with start < e
    {
      uint64_t b = (uint64_t)e >> tbl_shift;
      if (b >= (uint64_t)n_buckets) { lo = hi = nr; } else { lo = tbl[b]; hi = tbl[b + 1]; }
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (rstart[mid] < e) lo = mid + 1; else hi = mid; }
    }
    for (int64_t j = lo - 1; j >= 0 && rend[j] > s; --j) {   // strict overlap (GMAP4.scala:41); run ends are sorted
      for (int q = 0; q < plan.n_cols; ++q) {
        if (cc.ok[q] && !cc.ok[q][i]) continue;
        double v = cc.v[q][i];
        atomicAdd(&acc.nn[q][j], 1);
        atomicAdd(&acc.sum[q][j], v);
        atomicMin(&acc.vmin[q][j], cov_key_min(v));
        atomicMax(&acc.vmax[q][j], cov_key_max(v));
      }
    }
  }
}

__global__ void k_run_agg_finish(int64_t nr, CovAggPlan plan, RunAgg acc, const int32_t* __restrict__ cnt, CovOut out) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nr; c += (int64_t)gridDim.x * blockDim.x) {
    for (int a = 0; a < plan.n_aggs; ++a) {
      int kind = plan.kind[a], q = plan.slot[a];
      double v = 0.0; uint8_t ok = 0;
      if (kind == GMQL_AGG_COUNT) { v = (double)cnt[c]; ok = 1; }
      else if (q >= 0 && acc.nn[q][c] > 0) {
        ok = 1;
        int nn = acc.nn[q][c];
        if (kind == GMQL_AGG_SUM) v = acc.sum[q][c];
        else if (kind == GMQL_AGG_AVG) v = acc.sum[q][c] / (double)nn;
        else if (kind == GMQL_AGG_MIN) v = cov_from_key(acc.vmin[q][c]);
        else if (kind == GMQL_AGG_MAX) v = cov_from_key(acc.vmax[q][c]);
        else if (kind == GMQL_AGG_MEDIAN) { double m = cov_from_key(acc.vmin[q][c]); v = nn >= 2 ? (m + m) / 2 : m; }
      }
      out.value[a][c] = v;
      out.valid[a][c] = ok;
    }
  }
}

__global__ void k_cover_compact(int64_t nc, const int32_t* __restrict__ keep, const int64_t* __restrict__ keep_excl, const uint32_t* __restrict__ ccls, const int32_t* __restrict__ cchrom,
                                const int64_t* __restrict__ fstart, const int64_t* __restrict__ fstop, const int32_t* __restrict__ cacc, const double* __restrict__ j1, const double* __restrict__ j2,
                                CovOut in, int n_aggs, int n_classes, int unified, FinalCols f, double* __restrict__ acc_d, tiled::ShardRaw raw_in, tiled::ShardRaw raw_out) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[c]) continue;
    int64_t o = keep_excl[c];
    uint32_t cls = ccls[c];
    f.group[o] = (int32_t)(cls / (uint32_t)n_classes);
    f.strand[o] = unified ? (int8_t)0 : (int8_t)(cls % (uint32_t)n_classes + 1);
    f.chrom[o] = cchrom[c];
    f.start[o] = fstart[c];
    f.stop[o] = fstop[c];
    f.acc[o] = cacc[c];
    acc_d[o] = (double)cacc[c];       // the view of the result as a region dataset exposes AccIndex as a value column
    f.j1[o] = j1[c];
    f.j2[o] = j2[c];
    if (raw_in.edge) {
      raw_out.edge[o] = raw_in.edge[c]; raw_out.cnt[o] = raw_in.cnt[c]; raw_out.rstart[o] = raw_in.rstart[c]; raw_out.rstop[o] = raw_in.rstop[c];
      raw_out.smin[o] = raw_in.smin[c]; raw_out.emax[o] = raw_in.emax[c]; raw_out.smax[o] = raw_in.smax[c]; raw_out.emin[o] = raw_in.emin[c];
    }
    for (int a = 0; a < n_aggs; ++a) { f.value[a][o] = in.value[a][c]; f.valid[a][o] = in.valid[a][c]; }
  }
}

static void plan_cov_aggs(const gmql_agg* aggs, int n_aggs, int n_cols_avail, CovAggPlan* plan, std::vector<int>* need_cols) {
  GMQL_REQUIRE(n_aggs >= 0 && n_aggs <= 16, GMQL_ERR_UNSUPPORTED, "at most 16 aggregates per call");
  GMQL_REQUIRE(n_aggs == 0 || aggs != nullptr, GMQL_ERR_INVALID, "null aggregate list");
  *plan = CovAggPlan();
  plan->n_aggs = n_aggs;
  for (int a = 0; a < n_aggs; ++a) {
    int kind = aggs[a].kind;
    plan->kind[a] = kind;
    plan->slot[a] = -1;
    if (kind == GMQL_AGG_COUNT) continue;
    if (kind == GMQL_AGG_BAG || kind == GMQL_AGG_BAGD) { plan->need_bag = 1; continue; }   // strings are built on the host from the contributor lists
    GMQL_REQUIRE(kind >= GMQL_AGG_SUM && kind <= GMQL_AGG_MEDIAN, GMQL_ERR_INVALID, "unknown aggregate kind");
    int col = aggs[a].col;
    GMQL_REQUIRE(col >= 0 && col < n_cols_avail, GMQL_ERR_INVALID, "aggregate column index out of range");
    int q = -1;
    for (int j = 0; j < plan->n_cols; ++j) if (plan->col_id[j] == col) q = j;
    if (q < 0) {
      GMQL_REQUIRE(plan->n_cols < 8, GMQL_ERR_UNSUPPORTED, "at most 8 distinct aggregated columns per call");
      q = plan->n_cols++;
      plan->col_id[q] = col;
      need_cols->push_back(col);
    }
    plan->slot[a] = q;
  }
}

struct Phase1Out {
  DevBuf<uint32_t> cls; DevBuf<int32_t> chrom; DevBuf<int64_t> start, stop; DevBuf<int32_t> acc;
  int64_t n = 0;
};

template <typename T>
static int64_t read_back(gmql_ctx* ctx, const T* dev) {
  T v;
  CUDA_TRY(cudaMemcpyAsync(&v, dev, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  return (int64_t)v;
}

template <typename K>
static void phase1_sweep(gmql_ctx* ctx, const DevRegions& in, const uint32_t* cls, const LinMap& lm, const CoverCfg& cfg, int flag, uint64_t n_classes_total,
                         const K* sorted_starts /*may be null*/, DevBuf<K>& ek_a, bool ek_filled, Phase1Out* p1, DevBuf<K>* lin_s = nullptr, DevBuf<K>* lin_e = nullptr) {
  // lin_s / lin_e given: the phase-1 regions are handed out as they are produced — sorted, disjoint linear (start, stop) pairs —
  // and p1 gets only their number and AccIndex; p1_regions() converts them to (class, chrom, start, stop) if a consumer needs that
  const int64_t n = in.n;
  int g = grid_for(ctx, n, 256);
  DevBuf<K> sk_a, sk_b, ek_b(ctx, n);
  K* starts = const_cast<K*>(sorted_starts);
  int key_bits = bits_for(n_classes_total * lm.G);
  if (!starts) { sk_a.alloc(ctx, n); sk_b.alloc(ctx, n); }
  if (!ek_filled) LAUNCH(ctx, (k_cover_event_keys<K>), g, 256, 0, in.chrom, in.start, in.stop, in.coord_bits, cls, n, lm, cfg.B, starts ? (K*)nullptr : sk_a.p, ek_a.p);
  if (!starts) radix_sort<K>(ctx, sk_a.p, sk_b.p, nullptr, nullptr, n, key_bits, &starts, nullptr);
  K* stops = nullptr;
  radix_sort<K>(ctx, ek_a.p, ek_b.p, nullptr, nullptr, n, key_bits, &stops, nullptr);

  int64_t total = 2 * n;
  int64_t n_tiles = ceil_div64(total, MP_TILE);
  DevBuf<int64_t> tile_cnt(ctx, n_tiles + 1), splits(ctx, n_tiles + 1);
  int mg = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->num_sms * 8);
  LAUNCH(ctx, (k_merge_splits<K>), grid_for(ctx, n_tiles + 1, 128, 1), 128, 0, starts, n, stops, n, n_tiles, splits.p);
  DevBuf<K> rs, re;
  int64_t nr = 0;
  if (flag != GMQL_HISTOGRAM) {
    // COVER / FLAT: fused merge + run detection (no point list)
    LAUNCH(ctx, (k_cover_runs_fused<K, false>), mg, MP_THREADS, 0, starts, n, stops, n, n_tiles, lm.G, cfg, splits.p, tile_cnt.p, (const int64_t*)nullptr, (K*)nullptr, (K*)nullptr, (int32_t*)nullptr);
    device_scan<int64_t, int64_t, OpSum<int64_t>, false>(ctx, tile_cnt.p, tile_cnt.p, n_tiles, tile_cnt.p + n_tiles);
    nr = read_back<int64_t>(ctx, tile_cnt.p + n_tiles);
    rs.alloc(ctx, nr); re.alloc(ctx, nr);
    p1->acc.alloc(ctx, nr);
    p1->acc.zero();
    if (nr) LAUNCH(ctx, (k_cover_runs_fused<K, true>), mg, MP_THREADS, 0, starts, n, stops, n, n_tiles, lm.G, cfg, splits.p, (int64_t*)nullptr, tile_cnt.p, rs.p, re.p, p1->acc.p);
    sk_a.release(); sk_b.release(); ek_a.release(); ek_b.release();
  } else {
    // HISTOGRAM: distinct positions with the depth after them, then one segment per depth change
    LAUNCH(ctx, (k_cover_points<K, false>), mg, MP_THREADS, 0, starts, n, stops, n, n_tiles, splits.p, tile_cnt.p, (const int64_t*)nullptr, (K*)nullptr, (int32_t*)nullptr);
    device_scan<int64_t, int64_t, OpSum<int64_t>, false>(ctx, tile_cnt.p, tile_cnt.p, n_tiles, tile_cnt.p + n_tiles);
    int64_t np = read_back<int64_t>(ctx, tile_cnt.p + n_tiles);
    DevBuf<K> px(ctx, np);
    DevBuf<int32_t> pd(ctx, np);
    LAUNCH(ctx, (k_cover_points<K, true>), mg, MP_THREADS, 0, starts, n, stops, n, n_tiles, splits.p, (int64_t*)nullptr, tile_cnt.p, px.p, pd.p);
    sk_a.release(); sk_b.release(); ek_a.release(); ek_b.release();
    DevBuf<int32_t> flags(ctx, np);
    DevBuf<int64_t> excl(ctx, np + 1);
    int pg = grid_for(ctx, np, 256);
    if (np) LAUNCH(ctx, (k_hist_flags<K>), pg, 256, 0, px.p, pd.p, np, lm.G, cfg, flags.p);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, flags.p, excl.p, np, excl.p + np);
    nr = read_back<int64_t>(ctx, excl.p + np);
    rs.alloc(ctx, nr); re.alloc(ctx, nr);
    p1->acc.alloc(ctx, nr);
    p1->acc.zero();
    if (nr) LAUNCH(ctx, (k_hist_segments<K>), pg, 256, 0, px.p, pd.p, np, flags.p, excl.p, rs.p, re.p, p1->acc.p);
  }
  p1->n = nr;
  if (lin_s && lin_e) { *lin_s = std::move(rs); *lin_e = std::move(re); return; }
  p1->cls.alloc(ctx, nr); p1->chrom.alloc(ctx, nr); p1->start.alloc(ctx, nr); p1->stop.alloc(ctx, nr);
  if (nr) LAUNCH(ctx, (k_lin_to_regions<K>), grid_for(ctx, nr, 256), 256, 0, rs.p, re.p, nr, lm, p1->cls.p, p1->chrom.p, p1->start.p, p1->stop.p);
}
template <typename K>
static void p1_regions(gmql_ctx* ctx, const DevBuf<K>& lin_s, const DevBuf<K>& lin_e, const LinMap& lm, Phase1Out* p1) {
  const int64_t nr = p1->n;
  p1->cls.alloc(ctx, nr); p1->chrom.alloc(ctx, nr); p1->start.alloc(ctx, nr); p1->stop.alloc(ctx, nr);
  if (nr) LAUNCH(ctx, (k_lin_to_regions<K>), grid_for(ctx, nr, 256), 256, 0, (const K*)lin_s.p, (const K*)lin_e.p, nr, lm, p1->cls.p, p1->chrom.p, p1->start.p, p1->stop.p);
}

static void phase1_summit(gmql_ctx* ctx, const DevRegions& in, const uint32_t* cls, const SpanTable& spans, const CoverCfg& cfg, uint64_t n_classes_total, Phase1Out* p1, int kind) {
  const int64_t n = in.n;
  const int64_t B = cfg.B;
  int g = grid_for(ctx, n, 256);
  DevBuf<uint64_t> coff2(ctx, spans.n_chrom + 1);
  LAUNCH(ctx, k_summit_offsets, 1, 32, 0, spans.span.p, spans.n_chrom, B, coff2.p);
  DevBuf<int64_t> ecnt(ctx, n + 1);
  LAUNCH(ctx, k_summit_event_counts, g, 256, 0, in.start, in.stop, in.coord_bits, n, B, ecnt.p);
  device_scan<int64_t, int64_t, OpSum<int64_t>, false>(ctx, ecnt.p, ecnt.p, n, ecnt.p + n);
  uint64_t G2 = 0;
  int64_t ne = 0;
  CUDA_TRY(cudaMemcpyAsync(&G2, coff2.p + spans.n_chrom, sizeof G2, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&ne, ecnt.p + n, sizeof ne, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(ne < (int64_t)0xffffffffLL, GMQL_ERR_UNSUPPORTED, "SUMMIT: more than 2^32 per-bin events");
  GMQL_REQUIRE(bits_for(n_classes_total * G2) < 63, GMQL_ERR_UNSUPPORTED, "SUMMIT: coordinate space too large");
  DevBuf<uint64_t> ka(ctx, ne), kb(ctx, ne);
  LAUNCH(ctx, k_summit_events, g, 256, 0, in.chrom, in.start, in.stop, in.coord_bits, cls, n, B, coff2.p, G2, ecnt.p, ka.p);
  uint64_t* keys = nullptr;
  // (the sign bit below lin2 is only counted, never ordered: the sort starts at bit 1 — four passes instead of five on hg38)
  radix_sort<uint64_t>(ctx, ka.p, kb.p, nullptr, nullptr, ne, bits_for(n_classes_total * G2) + 1, &keys, nullptr, 1);
  // coalesce
  DevBuf<int32_t> head(ctx, ne);
  DevBuf<int64_t> head_excl(ctx, ne + 1);
  int eg = grid_for(ctx, ne, 256);
  if (ne) LAUNCH(ctx, k_summit_heads, eg, 256, 0, keys, ne, head.p);
  device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, head.p, head_excl.p, ne, head_excl.p + ne);
  int64_t np = read_back<int64_t>(ctx, head_excl.p + ne);
  DevBuf<uint64_t> pkey(ctx, np);
  DevBuf<int32_t> pdelta(ctx, np);
  if (ne) LAUNCH(ctx, k_summit_points, eg, 256, 0, keys, ne, head.p, head_excl.p, pkey.p, pdelta.p);
  ka.release(); kb.release(); head.release(); head_excl.release();
  // per-bin state machines
  DevBuf<int32_t> emit(ctx, np), pval(ctx, np);
  DevBuf<int64_t> pso(ctx, np), emit_excl(ctx, np + 1);
  {
    GMQL_REQUIRE(np < (int64_t)0xffffffffLL, GMQL_ERR_UNSUPPORTED, "SUMMIT: too many event points");
    DevBuf<uint32_t> bin_first;
    int64_t n_bins = 0;
    const uint64_t all_bins = (n_classes_total * G2) / (uint64_t)(B + 1);      // G2 is a whole number of bins
    const char* bo0 = ctx_opt(ctx, "cover.bins");             // "inplace": test hook (also: the bin list by head flags + scan)
    if (np > 0 && all_bins <= (uint64_t)(2 * np + 1024) && !(bo0 && !strcmp(bo0, "inplace"))) {
      // dense: one search per bin of the genome (empty bins are empty ranges)
      n_bins = (int64_t)all_bins;
      bin_first.alloc(ctx, n_bins);
      LAUNCH(ctx, k_summit_bin_search, grid_for(ctx, n_bins, 256), 256, 0, pkey.p, np, (uint64_t)(B + 1), n_bins, bin_first.p);
    } else {
    DevBuf<int32_t> bhead(ctx, np);
    DevBuf<int64_t> bhead_excl(ctx, np + 1);
    if (np) LAUNCH(ctx, k_summit_bin_heads, grid_for(ctx, np, 256), 256, 0, pkey.p, np, B, bhead.p);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, bhead.p, bhead_excl.p, np, bhead_excl.p + np);
    n_bins = read_back<int64_t>(ctx, bhead_excl.p + np);
    bin_first.alloc(ctx, n_bins);
    if (np) LAUNCH(ctx, k_summit_bin_list, grid_for(ctx, np, 256), 256, 0, bhead.p, bhead_excl.p, np, bin_first.p);
    }
    if (n_bins) {
      const size_t bsh = 8 * (size_t)BINS_CAP;
      const char* bo = ctx_opt(ctx, "cover.bins");            // "inplace": test hook (no staging of the points)
      const int bins_cap = (bo && !strcmp(bo, "inplace")) ? 0 : BINS_CAP;
      CUDA_TRY(cudaFuncSetAttribute(k_summit_bins, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsh));
      LAUNCH(ctx, k_summit_bins, (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_bins, BINS_T), (int64_t)ctx->num_sms * 2)), BINS_T, bsh,
             pkey.p, pdelta.p, np, bin_first.p, n_bins, B, G2, cfg, kind, bins_cap, emit.p, pso.p, pval.p);
    }
  }
  device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, emit.p, emit_excl.p, np, emit_excl.p + np);
  int64_t npieces = read_back<int64_t>(ctx, emit_excl.p + np);
  DevBuf<uint32_t> qcls(ctx, npieces); DevBuf<int32_t> qchrom(ctx, npieces), qval(ctx, npieces); DevBuf<int64_t> qs(ctx, npieces), qe(ctx, npieces);
  if (npieces) LAUNCH(ctx, k_summit_pieces, grid_for(ctx, np, 256), 256, 0, pkey.p, np, B, G2, coff2.p, spans.n_chrom, emit.p, emit_excl.p, pso.p, pval.p,
                      qcls.p, qchrom.p, qs.p, qe.p, qval.p);
  // cross-bin merge
  DevBuf<int32_t> chead(ctx, npieces);
  DevBuf<int64_t> chead_excl(ctx, npieces + 1);
  int qg = grid_for(ctx, npieces, 256);
  if (npieces) LAUNCH(ctx, k_chain_heads, qg, 256, 0, qcls.p, qchrom.p, qs.p, qe.p, qval.p, kind == BINS_HISTOGRAM ? 1 : 0, npieces, B, chead.p);
  device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, chead.p, chead_excl.p, npieces, chead_excl.p + npieces);
  int64_t nr = read_back<int64_t>(ctx, chead_excl.p + npieces);
  p1->n = nr;
  p1->cls.alloc(ctx, nr); p1->chrom.alloc(ctx, nr); p1->start.alloc(ctx, nr); p1->stop.alloc(ctx, nr); p1->acc.alloc(ctx, nr);
  p1->acc.zero();
  if (nr) LAUNCH(ctx, k_chain_merge, qg, 256, 0, qcls.p, qchrom.p, qs.p, qe.p, qval.p, npieces, chead.p, chead_excl.p, p1->cls.p, p1->chrom.p, p1->start.p, p1->stop.p, p1->acc.p);
}

struct ShardSpec { int64_t first_tile, end_tile; };
struct TiledRuns { DevBuf<uint32_t> rstart, rend; DevBuf<int32_t> cnt; DevBuf<int64_t> head_excl; int64_t np = 0; };   // for the aggregate post-pass
struct ShardBufs { DevBuf<int32_t> edge, cnt; DevBuf<int64_t> rstart, rstop, smin, emax, smax, emin; };
// speculative launch of the tile pipeline (cover_tiled.cuh, "speculative launch"): the prepass is fused with the tile histogram
// and nothing waits for its flags; they come back with the pipeline's one synchronisation
struct SpecCtx {
  int* flags = nullptr;                    // device [4]: any '*', zero-length seen, invalid row, longest region (zeroed by the caller)
  unsigned long long* act_span = nullptr;  // device [n_chrom]: largest stop seen per chromosome (zeroed by the caller)
  int* h_flags = nullptr;                  // pinned host copies, valid once `synced`
  unsigned long long* h_span = nullptr;
  int n_chrom = 0;
  bool synced = false;                     // the flags have reached the host
  bool tripped = false;                    // the input was not eligible after all (or invalid): the caller starts over without speculation
  bool declined = false;                   // the tile path was not chosen at all: nothing has run
};

// result columns written by the fused stitch kernel (capacity STITCH_MAX; n_out rows are valid)
struct FusedOut {
  bool ready = false;
  int64_t n_out = 0;
  DevBuf<int32_t> group, chrom, acc; DevBuf<int64_t> start, stop; DevBuf<int8_t> strand; DevBuf<double> j1, j2, accd;
};

// host side of the tile-resident path; returns false (nothing consumed) when a capacity limit was hit and the caller must
// fall back to the sort-and-merge path
// what a resident dataset contributes to the tile path (cover_impl's fast route): its narrow arrays are read as they are and
// the tile / coarse offsets computed at upload are reused when they were made for this layout
struct TileSource { bool narrow = false; gmql_dataset* ds = nullptr; };

static bool dataset_offsets_match(const gmql_dataset* ds, const LinMap& lm, uint64_t pad, int logw, int64_t nT, int groups, int n_coarse) {
  return ds && ds->tile_off.p && ds->lay_pad == pad && ds->lay_G == lm.G && logw == tiled::MAX_LOGW && ds->hist_nT == nT && ds->hist_groups == groups && ds->hist_n_coarse == n_coarse;
}

// counts + scans of the tile partition for a dataset's narrow arrays (at upload, or when an operator uses another layout)
static void tile_offsets_narrow(gmql_ctx* ctx, const DevRegions& in, const LinMap& lm, int logw, int64_t nT, int G, int n_coarse, uint32_t* tile_off /*[nT + 1]*/, uint32_t* coarse_off /*[n_coarse * G + 1]*/) {
  const int64_t n = in.n;
  DevBuf<uint32_t> hist(ctx, nT + 1), coarse_cnt(ctx, (int64_t)n_coarse * G);
  hist.zero();
  auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
  const int vec_ok = al16(in.chrom_narrow) && al16(in.start);
  const size_t hsh = sizeof(uint32_t) * ((size_t)nT + 256 + (size_t)n_coarse);
  CUDA_TRY(cudaFuncSetAttribute(tiled::k_tile_hist_narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsh));
  LAUNCH(ctx, tiled::k_tile_hist_narrow, (unsigned)G, 1024, hsh, (const uint8_t*)in.chrom_narrow, (const int32_t*)in.start, n, lm, logw, (int)nT, hist.p, coarse_cnt.p, n_coarse, vec_ok);
  device_scan<uint32_t, uint32_t, OpSum<uint32_t>, false>(ctx, hist.p, tile_off, nT, tile_off + nT);
  device_scan<uint32_t, uint32_t, OpSum<uint32_t>, false>(ctx, coarse_cnt.p, coarse_off, (int64_t)n_coarse * G, coarse_off + (int64_t)n_coarse * G);
}

// host side of the tile-resident path; returns false (nothing consumed) when a capacity limit was hit and the caller must
// fall back to the sort-and-merge path
static bool run_cover_tiled(gmql_ctx* ctx, const DevRegions& in, const uint32_t* cls, const LinMap& lm, const CoverCfg& cfg, int flag, uint64_t n_classes_total, int logw,
                            const int32_t* h_mn, const int32_t* h_mx, Phase1Out* p1, DevBuf<int32_t>& keep, DevBuf<int64_t>& fstart, DevBuf<int64_t>& fstop,
                            DevBuf<double>& j1, DevBuf<double>& j2, int64_t* nc_out, const ShardSpec* shard, ShardBufs* sb, TiledRuns* keep_runs, FusedOut* fo, SpecCtx* spec,
                            const TileSource& src, bool use_wc, int64_t maxlen_known, int* fail_code) {
  *fail_code = 0;
  const int64_t n = in.n;
  // lengths below 32 768 (known before the launch): the partition decides the one-base extension and leaves it in bit 15 of the record
  const int xb = (maxlen_known >= 0 && maxlen_known < 32768 && cfg.B > 0) ? 1 : 0;
  const tiled::BinMod32 bm32 = tiled::make_binmod32(cfg.B);
  tiled::Guard guard;
  guard.flags = (spec && !spec->synced) ? spec->flags : nullptr;
  guard.max_len = (int)((1ll << logw) - 2);
  const int key_bits = bits_for(n_classes_total * lm.G);
  const int64_t nT = (int64_t)((n_classes_total * lm.G) >> logw) + 1;
  // the tiles this call owns: all of them, or one shard's range (the tile before the range only lends its open regions)
  const int64_t t_lo = shard ? shard->first_tile : 0, t_hi = shard ? std::min<int64_t>(shard->end_tile, nT) : nT;
  GMQL_REQUIRE(t_lo >= 0 && t_lo <= t_hi, GMQL_ERR_INVALID, "bad tile range");
  DevBuf<uint32_t> rec(ctx, n);                           // the partitioned 4-byte records
  DevBuf<uint32_t> offS_own;
  const uint32_t* offS = nullptr;
  if (nT <= tiled::HIST_CAP) {
    // counting + two multisplit passes (cover_tiled.cuh): the records never exist unpartitioned
    const int n_coarse = (int)((nT + 255) >> 8) + 1;
    // groups of input tiles (k_tile_hist): block b of the counting kernel and the blocks b, b + G, ... of the coarse pass
    const int64_t in_tiles = ceil_div64(n, tiled::MSP_TILE);
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(in_tiles, (int64_t)ctx->num_sms));
    DevBuf<uint32_t> key_a(ctx, n);
    DevBuf<uint16_t> len_a(ctx, n);
    DevBuf<uint32_t> cur_tile(ctx, nT + 1), cur_coarse(ctx, (int64_t)n_coarse * G + 1);
    auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
    if (src.narrow && dataset_offsets_match(src.ds, lm, src.ds ? src.ds->lay_pad : 0, logw, nT, G, n_coarse)) {
      // everything up to the partition was done when the dataset was uploaded: the cursors start from its offsets
      offS = src.ds->tile_off.p;
      CUDA_TRY(cudaMemcpyAsync(cur_tile.p, src.ds->tile_off.p, 4 * (size_t)(nT + 1), cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(cur_coarse.p, src.ds->coarse_off.p, 4 * ((size_t)n_coarse * G + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      offS_own.alloc(ctx, nT + 1);
      offS = offS_own.p;
      if (src.narrow) {
        tile_offsets_narrow(ctx, in, lm, logw, nT, G, n_coarse, offS_own.p, cur_coarse.p);
      } else {
        DevBuf<uint32_t> hist(ctx, nT + 1), coarse_cnt(ctx, (int64_t)n_coarse * G);
        hist.zero();
        const int vec_ok = al16(in.chrom) && al16(in.start) && al16(in.stop) && al16(cls) && cfg.B > 0 && cfg.B < 0x7fffffffLL;
        if (guard.flags) {
          // speculative: the prepass and the histogram in one read of the columns
          size_t hsh = sizeof(uint32_t) * (size_t)((nT + 1) & ~1) + 16 * (size_t)lm.n_chrom + sizeof(uint32_t) * (size_t)n_coarse;
          CUDA_TRY(cudaFuncSetAttribute(tiled::k_prepass_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsh));
          const int pv_ok = al16(in.chrom) && al16(in.start) && al16(in.stop) && al16(in.sample);
          LAUNCH(ctx, tiled::k_prepass_hist, (unsigned)G, 1024, hsh, in.chrom, (const int32_t*)in.start, (const int32_t*)in.stop, in.sample, n, in.n_samples, lm, logw, (int)nT,
                 hist.p, coarse_cnt.p, n_coarse, pv_ok, spec->act_span, spec->flags);
        } else {
          size_t hsh = sizeof(uint32_t) * (size_t)((nT + 1) & ~1) + (lm.n_chrom <= 1024 ? 8 * (size_t)lm.n_chrom : 0) + sizeof(uint32_t) * (size_t)n_coarse;
          CUDA_TRY(cudaFuncSetAttribute(tiled::k_tile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsh));
          LAUNCH(ctx, tiled::k_tile_hist, (unsigned)G, 1024, hsh, in.chrom, in.start, in.coord_bits, cls, n, lm, logw, (int)nT, hist.p, coarse_cnt.p, n_coarse, vec_ok);
        }
        device_scan<uint32_t, uint32_t, OpSum<uint32_t>, false>(ctx, hist.p, offS_own.p, nT, offS_own.p + nT);
        device_scan<uint32_t, uint32_t, OpSum<uint32_t>, false>(ctx, coarse_cnt.p, cur_coarse.p, (int64_t)n_coarse * G, cur_coarse.p + (int64_t)n_coarse * G);
      }
      LAUNCH(ctx, tiled::k_init_cursors, grid_for(ctx, nT, 256), 256, 0, offS, (int)nT, cur_tile.p);
    }
    // the coarse pass runs whole multiples of G blocks, so that block x only sees input tiles of group x % G
    const unsigned mg_c = (unsigned)(G * std::max<int64_t>(1, std::min<int64_t>(16, ceil_div64(in_tiles, G))));
    const unsigned mg = (unsigned)std::min<int64_t>(in_tiles, (int64_t)ctx->num_sms * 16);
    tiled::MspIn mi;
    memset(&mi, 0, sizeof mi);
    tiled::MspOut mo;
    mo.key = key_a.p; mo.len = len_a.p; mo.rec = nullptr;
    if (src.narrow) {
      mi.chrom8 = (const uint8_t*)in.chrom_narrow; mi.start = in.start; mi.len16 = (const uint16_t*)in.len_narrow; mi.bits = 32;
      const int vec_ok = al16(in.chrom_narrow) && al16(in.start) && al16(in.len_narrow) && cfg.B > 0 && cfg.B < 0x7fffffffLL;
      CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::msp_smem_bytes<tiled::MSP_NARROW>()));
      CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_NARROW>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   // three blocks per SM
      LAUNCH(ctx, tiled::k_multisplit<tiled::MSP_NARROW>, mg_c, tiled::MSP_THREADS, tiled::msp_smem_bytes<tiled::MSP_NARROW>(), mi, lm, cfg.B, n, 32 + logw + 8, logw, cur_coarse.p, mo, vec_ok, G, guard, xb, bm32);
    } else {
      mi.chrom = in.chrom; mi.start = in.start; mi.stop = in.stop; mi.bits = in.coord_bits; mi.cls = cls;
      const int vec_ok = al16(in.chrom) && al16(in.start) && al16(in.stop) && al16(cls) && cfg.B > 0 && cfg.B < 0x7fffffffLL;
      CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::msp_smem_bytes<tiled::MSP_WIDE>()));
      CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   // three blocks per SM
      LAUNCH(ctx, tiled::k_multisplit<tiled::MSP_WIDE>, mg_c, tiled::MSP_THREADS, tiled::msp_smem_bytes<tiled::MSP_WIDE>(), mi, lm, cfg.B, n, 32 + logw + 8, logw, cur_coarse.p, mo, vec_ok, G, guard, xb, bm32);
    }
    memset(&mi, 0, sizeof mi);
    mi.key = key_a.p; mi.len = len_a.p;
    mo.key = nullptr; mo.len = nullptr; mo.rec = rec.p;
    CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_FINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::msp_smem_bytes<tiled::MSP_FINE>()));
      CUDA_TRY(cudaFuncSetAttribute(tiled::k_multisplit<tiled::MSP_FINE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   // three blocks per SM
    LAUNCH(ctx, tiled::k_multisplit<tiled::MSP_FINE>, mg, tiled::MSP_THREADS, tiled::msp_smem_bytes<tiled::MSP_FINE>(), mi, lm, cfg.B, n, 32 + logw, logw, cur_tile.p, mo, 0, 1, guard, xb, bm32);
  } else {
    GMQL_REQUIRE(!src.narrow, GMQL_ERR_INVALID, "internal: the radix route of the tile path needs the wide columns");
    DevBuf<uint64_t> pk_a(ctx, n), pk_b(ctx, n);
    uint64_t* P64;
    offS_own.alloc(ctx, nT + 1);
    offS = offS_own.p;
    LAUNCH(ctx, tiled::k_tile_keys, grid_for(ctx, n, 256), 256, 0, in.chrom, in.start, in.stop, in.coord_bits, cls, n, lm, cfg.B, pk_a.p);
    radix_sort<uint64_t>(ctx, pk_a.p, pk_b.p, nullptr, nullptr, n, 32 + key_bits, &P64, nullptr, 32 + logw);   // partition by tile only
    LAUNCH(ctx, tiled::k_tile_offsets, grid_for(ctx, n, 256), 256, 0, P64, n, nT, logw, offS_own.p);
    LAUNCH(ctx, tiled::k_pack_rec32, grid_for(ctx, n, 256), 256, 0, (const uint64_t*)P64, n, logw, rec.p, xb);
  }
  const uint32_t* P = rec.p;

  // two resident blocks per SM; four ranges per resident block even out the tail.  Every block publishes its pieces in
  // its own segment of the piece tables, so they come out in genome order and nothing has to be sorted.
  // whole waves of the 2 x num_sms resident blocks, about 16 tiles per block (the first tile of every block range re-reads
  // its predecessor's regions), at most four waves
  const int64_t resident = (int64_t)ctx->num_sms * tiled::CTAS_PER_SM;
  int64_t waves = std::max<int64_t>(1, std::min<int64_t>(4, (t_hi - t_lo + resident * 8) / (resident * 16)));
  if (cfg.n_groups == 1 && use_wc) waves = 1;             // the word-count kernel's blocks cost more to set up; their ranges hold equal shares of the regions anyway
  if (const char* we = ctx_opt(ctx, "cover.waves")) waves = std::max(1, atoi(we));   // tuning hook
  unsigned tgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(t_hi - t_lo, 1), resident * waves));
  if (const char* be = ctx_opt(ctx, "cover.blocks")) tgrid = (unsigned)std::max(1, atoi(be));   // tuning hook
  // piece capacity of a block's segment.  Dense inputs (the tile path's home ground) make few pieces: about two per tile.  With few
  // regions per tile nearly every region can be a piece of its own (at most MAXP per tile); the blocks hold equal shares of the
  // regions, so the pieces spread about evenly — half as much again covers the difference
  const int64_t n_own_tiles = std::max<int64_t>(t_hi - t_lo, 1);
  const int64_t est_pieces = n >= 1024 * n_own_tiles ? n / 16 + 2 * nT : std::min<int64_t>(n + n_own_tiles, n_own_tiles * (int64_t)tiled::MAXP) * 3 / 2;
  const int64_t seg = std::max<int64_t>(4096, est_pieces / tgrid + 1);
  const int64_t cap = seg * tgrid;
  GMQL_REQUIRE(cap < (int64_t)0xffffffffLL, GMQL_ERR_UNSUPPORTED, "piece table too large");
  DevBuf<uint32_t> pstart(ctx, cap), pend(ctx, cap), smin(ctx, cap), emax(ctx, cap), smax(ctx, cap), emin(ctx, cap);
  DevBuf<int32_t> acc(ctx, cap), cnt(ctx, cap);
  DevBuf<uint8_t> pflags(ctx, cap);
  DevBuf<uint32_t> blk_cnt(ctx, tgrid), blk_off(ctx, (int64_t)tgrid + 1);
  DevBuf<unsigned long long> np_dev(ctx, 1);
  DevBuf<int> fail(ctx, 4);                              // [0] failure code; [1..3] profiling counters of option cover.dbg = 8
  fail.zero();
  tiled::Cfg tc;
  tc.n_groups = cfg.n_groups; tc.n_classes = cfg.n_classes; tc.mn = cfg.mn; tc.mx = cfg.mx; tc.mn0 = h_mn[0]; tc.mx0 = h_mx[0]; tc.G = lm.G; tc.xb = xb; tc.dbg = ctx_opt(ctx, "cover.dbg") ? atoi(ctx_opt(ctx, "cover.dbg")) : 0;
  tiled::Pieces pc;
  pc.pstart = pstart.p; pc.pend = pend.p; pc.acc = acc.p; pc.cnt = cnt.p; pc.smin = smin.p; pc.emax = emax.p; pc.smax = smax.p; pc.emin = emin.p;
  pc.flags = pflags.p; pc.blk_cnt = blk_cnt.p; pc.seg = seg; pc.fail = fail.p;
  DevBuf<unsigned long long> ticks(ctx, 16);
  ticks.zero();
  pc.ticks = ticks.p;
  if (cfg.n_groups == 1 && use_wc) {
    // the word-count kernel (cover_tile_wc.cuh); a tile with too many words to resolve sends the operator to the rank kernel
    const tiled::BinMod32 bm = bm32;
#define WC_LAUNCH(F, X)                                                                                                                   \
    do {                                                                                                                                  \
      CUDA_TRY(cudaFuncSetAttribute((tiled::k_cover_tile_wc<F, X>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::WC_SMEM_BYTES)); \
      LAUNCH(ctx, (tiled::k_cover_tile_wc<F, X>), tgrid, tiled::THREADS, tiled::WC_SMEM_BYTES, P, offS, t_lo, t_hi, logw, tc, lm, bm, pc, guard); \
    } while (0)
    if (logw == tiled::MAX_LOGW) { if (xb) WC_LAUNCH(true, true); else WC_LAUNCH(true, false); }
    else { if (xb) WC_LAUNCH(false, true); else WC_LAUNCH(false, false); }
#undef WC_LAUNCH
  } else if (cfg.n_groups > 1) {
    CUDA_TRY(cudaFuncSetAttribute(tiled::k_cover_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::SMEM_BYTES));
    LAUNCH(ctx, tiled::k_cover_tile<true>, tgrid, tiled::THREADS, tiled::SMEM_BYTES, P, offS, t_lo, t_hi, logw, tc, lm, cfg.B, pc, guard);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(tiled::k_cover_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled::SMEM_BYTES));
    LAUNCH(ctx, tiled::k_cover_tile<false>, tgrid, tiled::THREADS, tiled::SMEM_BYTES, P, offS, t_lo, t_hi, logw, tc, lm, cfg.B, pc, guard);
  }
  unsigned long long h_np = 0;
  int h_fail = 0;
  const char* stitch_env = ctx_opt(ctx, "cover.stitch");   // "split": test hook, always the separate stitch kernels
  const bool fused = fo && !sb && !keep_runs && (int64_t)tgrid <= tiled::STITCH_MAXB && !(stitch_env && !strcmp(stitch_env, "split"));
  if (fused) {
    // few pieces (the usual case): one kernel from the pieces to the result columns, one synchronisation
    const int64_t C = tiled::STITCH_MAX;
    DevBuf<unsigned long long> totals(ctx, 4);
    totals.zero();
    DevBuf<uint32_t> order_buf(ctx, C), rstart(ctx, C), rend(ctx, C), rsmin(ctx, C), remax(ctx, C), rsmax(ctx, C), remin(ctx, C);
    DevBuf<int32_t> ridx(ctx, C), racc(ctx, C), rcnt(ctx, C);
    tiled::Runs rr;
    rr.edge = nullptr;
    rr.rstart = rstart.p; rr.rend = rend.p; rr.acc = racc.p; rr.cnt = rcnt.p; rr.smin = rsmin.p; rr.emax = remax.p; rr.smax = rsmax.p; rr.emin = remin.p;
    fo->group.alloc(ctx, C); fo->chrom.alloc(ctx, C); fo->acc.alloc(ctx, C); fo->start.alloc(ctx, C); fo->stop.alloc(ctx, C);
    fo->strand.alloc(ctx, C); fo->j1.alloc(ctx, C); fo->j2.alloc(ctx, C); fo->accd.alloc(ctx, C);
    tiled::StitchOut so;
    so.group = fo->group.p; so.strand = fo->strand.p; so.chrom = fo->chrom.p; so.start = fo->start.p; so.stop = fo->stop.p;
    so.acc = fo->acc.p; so.accd = fo->accd.p; so.j1 = fo->j1.p; so.j2 = fo->j2.p; so.totals = totals.p;
    LAUNCH(ctx, tiled::k_stitch_small, tiled::STITCH_CL, tiled::STITCH_T, 0, (const uint32_t*)blk_cnt.p, (int)tgrid, blk_off.p, pc, order_buf.p, ridx.p, rr, lm, flag == GMQL_FLAT ? 1 : 0,
           cfg.n_classes, cfg.unified, so, guard);
    unsigned long long* h_tot = (unsigned long long*)ctx_pinned(ctx, 64);
    int* h_f = (int*)(h_tot + 4);
    CUDA_TRY(cudaMemcpyAsync(h_tot, totals.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(h_f, fail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (guard.flags) {
      CUDA_TRY(cudaMemcpyAsync(spec->h_flags, spec->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(spec->h_span, spec->act_span, 8 * (size_t)spec->n_chrom, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(GMQL_SYNC(ctx));
    if (guard.flags) {
      spec->synced = true;
      if (spec->h_flags[1] || spec->h_flags[2] || spec->h_flags[3] >= guard.max_len) { spec->tripped = true; return false; }
    }
    if (tc.dbg == 9) {
      unsigned long long ht[16];
      CUDA_TRY(cudaMemcpy(ht, ticks.p, sizeof ht, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[cover.dbg] cycles per phase (sum over tiles, thread 0):");
      for (int k = 0; k < 12; ++k) fprintf(stderr, " %d:%.1fM", k, (double)ht[k] * 1e-6);
      fprintf(stderr, "\n");
    }
    if (tc.dbg == 8) {
      int h4[4];
      CUDA_TRY(cudaMemcpy(h4, fail.p, sizeof h4, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[cover.dbg] tile kernel: %d tiles on the fast path, %d on the regular path, %d words resolved\n", h4[1], h4[2], h4[3]);
    }
    if (*h_f) { *fail_code = *h_f; return false; }
    if (!h_tot[3]) {
      fo->ready = true;
      fo->n_out = (int64_t)h_tot[2];
      *nc_out = (int64_t)h_tot[1];
      p1->n = (int64_t)h_tot[1];
      return true;
    }
    h_np = h_tot[0];                                      // too many pieces for one block: the separate kernels below (blk_off is in place)
  } else {
    LAUNCH(ctx, tiled::k_seg_offsets, 1, 1024, 0, blk_cnt.p, (int)tgrid, blk_off.p, np_dev.p);
    CUDA_TRY(cudaMemcpyAsync(&h_np, np_dev.p, sizeof h_np, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&h_fail, fail.p, sizeof h_fail, cudaMemcpyDeviceToHost, ctx->stream));
    if (guard.flags) {
      CUDA_TRY(cudaMemcpyAsync(spec->h_flags, spec->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(spec->h_span, spec->act_span, 8 * (size_t)spec->n_chrom, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(GMQL_SYNC(ctx));
    if (guard.flags) {
      spec->synced = true;
      if (spec->h_flags[1] || spec->h_flags[2] || spec->h_flags[3] >= guard.max_len) { spec->tripped = true; return false; }
    }
    if (h_fail) { *fail_code = h_fail; return false; }
  }
  const int64_t np = (int64_t)h_np;
  rec.release();

  // pieces in genome order -> runs: fuse the ones that touch at a tile edge.  The number of runs (<= np) stays on the
  // device: the run tables are sized for np and the tail entries are dropped by the keep flags.
  const int64_t nr = np;
  DevBuf<uint32_t> order_buf(ctx, np);
  DevBuf<uint32_t> rstart(ctx, nr), rend(ctx, nr), rsmin(ctx, nr), remax(ctx, nr), rsmax(ctx, nr), remin(ctx, nr);
  DevBuf<int32_t> racc(ctx, nr), rcnt(ctx, nr), redge(ctx, nr);
  tiled::Runs rr;
  rr.edge = redge.p;
  rr.rstart = rstart.p; rr.rend = rend.p; rr.acc = racc.p; rr.cnt = rcnt.p; rr.smin = rsmin.p; rr.emax = remax.p; rr.smax = rsmax.p; rr.emin = remin.p;
  if (np) LAUNCH(ctx, tiled::k_piece_order, grid_for(ctx, np, 256), 256, 0, blk_off.p, (int)tgrid, seg, np, order_buf.p, rr);
  const uint32_t* order = order_buf.p;
  DevBuf<int32_t> head(ctx, np);
  DevBuf<int64_t> head_excl(ctx, np + 1);
  if (np) LAUNCH(ctx, tiled::k_piece_heads, grid_for(ctx, np, 256), 256, 0, order, pstart.p, pend.p, pflags.p, np, head.p);
  device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, head.p, head_excl.p, np, head_excl.p + np);
  if (np) LAUNCH(ctx, tiled::k_piece_merge, grid_for(ctx, np, 256), 256, 0, order, pc, np, head.p, head_excl.p, rr);
  p1->n = nr;
  p1->cls.alloc(ctx, nr); p1->chrom.alloc(ctx, nr); p1->acc.alloc(ctx, nr);
  keep.alloc(ctx, nr); fstart.alloc(ctx, nr); fstop.alloc(ctx, nr); j1.alloc(ctx, nr); j2.alloc(ctx, nr);
  tiled::ShardRaw raw;
  memset(&raw, 0, sizeof raw);
  if (sb) {
    sb->edge.alloc(ctx, nr); sb->cnt.alloc(ctx, nr); sb->rstart.alloc(ctx, nr); sb->rstop.alloc(ctx, nr);
    sb->smin.alloc(ctx, nr); sb->emax.alloc(ctx, nr); sb->smax.alloc(ctx, nr); sb->emin.alloc(ctx, nr);
    raw.edge = sb->edge.p; raw.cnt = sb->cnt.p; raw.rstart = sb->rstart.p; raw.rstop = sb->rstop.p;
    raw.smin = sb->smin.p; raw.emax = sb->emax.p; raw.smax = sb->smax.p; raw.emin = sb->emin.p;
  }
  if (nr) LAUNCH(ctx, tiled::k_runs_finalize, grid_for(ctx, nr, 256), 256, 0, rr, nr, (const int64_t*)(head_excl.p + np), lm, flag == GMQL_FLAT ? 1 : 0, p1->cls.p, p1->chrom.p, fstart.p, fstop.p, p1->acc.p, keep.p, j1.p, j2.p, raw);
  *nc_out = nr;
  if (keep_runs) {
    keep_runs->rstart = std::move(rstart); keep_runs->rend = std::move(rend); keep_runs->cnt = std::move(rcnt);
    keep_runs->head_excl = std::move(head_excl); keep_runs->np = np;
  }
  return true;
}

template <typename K>
static void run_cover(gmql_ctx* ctx, DevRegions& in, const int32_t* h_sample_group, int n_groups, int flag, const int32_t* h_mn, const int32_t* h_mx, int64_t B,
                      const CovAggPlan& plan, const SpanTable& spans, bool unified, bool has_zero, int64_t maxlen, gmql_result* res, const ShardSpec* shard = nullptr,
                      SpecCtx* spec = nullptr, TileSource src = TileSource()) {
  const int64_t n = in.n;
  CoverCfg cfg;
  cfg.n_groups = n_groups;
  cfg.unified = unified ? 1 : 0;
  cfg.n_classes = unified ? 1 : 2;
  cfg.B = B;
  DevBuf<int32_t> d_sg, d_mn(ctx, n_groups), d_mx(ctx, n_groups);
  if (h_sample_group) {
    d_sg.alloc(ctx, in.n_samples);
    CUDA_TRY(cudaMemcpyAsync(d_sg.p, h_sample_group, 4 * (size_t)in.n_samples, cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(cudaMemcpyAsync(d_mn.p, h_mn, 4 * (size_t)n_groups, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(d_mx.p, h_mx, 4 * (size_t)n_groups, cudaMemcpyHostToDevice, ctx->stream));
  cfg.sample_group = d_sg.p;
  cfg.mn = d_mn.p;
  cfg.mx = d_mx.p;
  const uint64_t n_classes_total = (uint64_t)n_groups * (uint64_t)cfg.n_classes;
  LinMap lm = spans.map();
  // bin-wise phase 1 (k_summit_bins): SUMMIT always; the others when minAcc < 1 (or option "cover.path" = "bins", a cross-check)
  bool binwise = flag == GMQL_SUMMIT;
  for (int g = 0; g < n_groups; ++g) if (h_mn[g] < 1) binwise = true;
  if (const char* f = ctx_opt(ctx, "cover.path")) if (!strcmp(f, "bins")) binwise = true;
  GMQL_REQUIRE(!(binwise && shard), GMQL_ERR_UNSUPPORTED, "gmql_cover_shard: minAcc < 1 needs the bin-wise path, which is not sharded by coordinate range");
  if (spec && binwise) { spec->declined = true; return; }

  // class of every region (only materialised when there is more than one class)
  DevBuf<uint32_t> cls;
  if (n_classes_total > 1) {
    gmql_ensure_wide(ctx, &in);
    src.narrow = false;
    cls.alloc(ctx, n);
    if (n) LAUNCH(ctx, k_cover_classes, grid_for(ctx, n, 256), 256, 0, in.strand, in.sample, n, cfg, cls.p);
  }

  Phase1Out p1;
  // merge form of GMAP4 (no value aggregates): count and bounds per phase-1 region, finished straight into the result columns
  DevBuf<int32_t> m_cnt;
  DevBuf<K> m_smin, m_emax, m_smax, m_emin;
  DevBuf<int> m_unsorted;
  bool m_direct = false;
  DevBuf<K> lin_s, lin_e;                                  // phase-1 regions in linear form (the sweep's own output)
  bool lin_form = false;                                   // ... and nothing ever converts them: merge form + direct finish
  DevBuf<int> lin_check;
  DevBuf<int32_t> keep, ocount;
  DevBuf<int64_t> keep_excl, fstart, fstop;
  DevBuf<double> j1, j2;
  CovOut tmp;
  memset(&tmp, 0, sizeof tmp);
  std::vector<DevBuf<double>> tv((size_t)plan.n_aggs);
  std::vector<DevBuf<uint8_t>> tk((size_t)plan.n_aggs);
  int64_t nc = 0;

  // ---- tile-resident path (cover_tiled.cuh): dense inputs, 32-bit keys, COVER / FLAT without aggregates
  bool done = false;
  ShardBufs sbufs;
  TiledRuns truns;
  FusedOut fo;
  if (sizeof(K) == 4 && n > 0 && (flag == GMQL_COVER || flag == GMQL_FLAT) && !binwise && !has_zero && (plan.n_aggs == 0 || !shard) &&
      n_classes_total * lm.G + 2ull * (1ull << tiled::MAX_LOGW) < 0xffffffffull) {
    // widest tile whose expected number of events leaves headroom in the kernel's tables
    const double total = (double)(n_classes_total * lm.G);
    int logw = tiled::MAX_LOGW;
    while (logw > 8 && 2.0 * (double)n * (double)(1ull << logw) / total > tiled::CAP_EV / 2.5) --logw;
    const char* force = ctx_opt(ctx, "cover.path");   // "tiled" / "sort": test hook; default = density heuristic
    const char* flw = ctx_opt(ctx, "cover.logw");
    if (flw) logw = atoi(flw);
    if (shard) logw = tiled::MAX_LOGW;                  // shards are cut on 65 536-position tiles
    bool fits = logw >= 8 && logw <= tiled::MAX_LOGW && maxlen < (int64_t)(1ll << logw) - 2;
    if (src.narrow && (int64_t)((n_classes_total * lm.G) >> logw) + 1 > tiled::HIST_CAP) { gmql_ensure_wide(ctx, &in); src.narrow = false; }   // radix route reads the wide columns
    bool want = fits && (double)n * (double)(1ull << logw) / total >= 256.0;   // enough regions per tile for the tile pass to pay
    if (force && !strcmp(force, "tiled")) want = fits;
    if (force && !strcmp(force, "sort")) want = false;
    if (shard) want = fits;
    if (spec && (!want || (int64_t)((n_classes_total * lm.G) >> logw) + 1 > tiled::HIST_CAP || lm.n_chrom > 1024)) { spec->declined = true; return; }
    // A pile-up (many regions on a few kilobases, as real peak data has at promoters) can overflow one tile's tables:
    // retry with tiles four and sixteen times narrower before giving the input to the sort-and-merge path
    const char* kopt = ctx_opt(ctx, "cover.kernel");      // "rank" / "count": test hook; default = the word-count kernel where it applies
    bool use_wc = !(kopt && !strcmp(kopt, "rank"));
    for (int attempt = 0; want && !done && attempt < 3; ++attempt) {
      const int lw = logw - 2 * attempt;
      if (lw < 8 || maxlen >= (int64_t)(1ll << lw) - 2) break;
      int fail_code = 0;
      done = run_cover_tiled(ctx, in, cls.p, lm, cfg, flag, n_classes_total, lw, h_mn, h_mx, &p1, keep, fstart, fstop, j1, j2, &nc, shard, shard ? &sbufs : nullptr, plan.n_aggs ? &truns : nullptr,
                             (plan.n_aggs == 0 && !shard) ? &fo : nullptr, spec, src, use_wc, (spec && !spec->synced) ? (int64_t)-1 : maxlen, &fail_code);
      if (!done && fail_code == 5 && use_wc && !(spec && spec->tripped)) {      // only the word-count kernel's own table overflowed: same tiles, rank kernel
        use_wc = false;
        --attempt;
        if (spec) maxlen = spec->h_flags[3];
        continue;
      }
      if (spec) {
        if (spec->tripped) return;                     // not eligible after all: the caller starts over with a blocking prepass
        maxlen = spec->h_flags[3];                     // the flags are on the host now: later attempts (and the sort path) are ordinary
      }
      if (flw || shard) break;                         // a forced width is not second-guessed
    }
  } else if (spec) { spec->declined = true; return; }

  if (!done || plan.n_aggs) gmql_ensure_wide(ctx, &in);   // everything below reads the int32 / int64 columns
  if (done && plan.n_aggs) {
    // value aggregates of the tile-resident path: region-driven post-pass over the runs
    const int64_t nr = nc;
    CovCols cc;
    memset(&cc, 0, sizeof cc);
    for (int q = 0; q < plan.n_cols; ++q) { cc.v[q] = in.cols[plan.col_id[q]]; cc.ok[q] = in.valid[plan.col_id[q]]; }
    for (int a = 0; a < plan.n_aggs; ++a) { tv[a].alloc(ctx, nr); tk[a].alloc(ctx, nr); tmp.value[a] = tv[a].p; tmp.valid[a] = tk[a].p; }
    RunAgg ra;
    memset(&ra, 0, sizeof ra);
    std::vector<DevBuf<int32_t>> a_nn((size_t)plan.n_cols);
    std::vector<DevBuf<double>> a_sum((size_t)plan.n_cols);
    std::vector<DevBuf<unsigned long long>> a_vmin((size_t)plan.n_cols), a_vmax((size_t)plan.n_cols);
    for (int q = 0; q < plan.n_cols; ++q) {
      a_nn[q].alloc(ctx, nr); a_nn[q].zero(); a_sum[q].alloc(ctx, nr); a_sum[q].zero();
      a_vmin[q].alloc(ctx, nr); a_vmin[q].fill_byte(0xff); a_vmax[q].alloc(ctx, nr); a_vmax[q].zero();
      ra.nn[q] = a_nn[q].p; ra.sum[q] = a_sum[q].p; ra.vmin[q] = a_vmin[q].p; ra.vmax[q] = a_vmax[q].p;
    }
    if (nr && plan.n_cols) {
      const uint64_t max_key = n_classes_total * lm.G;
      int tshift = 0;
      { int64_t want = std::max<int64_t>(1024, std::min<int64_t>((int64_t)1 << 22, 2 * nr + 1)); while (tshift < 62 && (int64_t)(max_key >> tshift) + 2 > want) ++tshift; }
      const int64_t nb = (int64_t)(max_key >> tshift) + 1;
      DevBuf<uint32_t> rtbl(ctx, nb + 2);
      const int64_t* n_runs = truns.head_excl.p + truns.np;
      LAUNCH(ctx, k_run_table, grid_for(ctx, nb + 1, 256), 256, 0, (const uint32_t*)truns.rstart.p, n_runs, tshift, nb, rtbl.p);
      LAUNCH(ctx, k_run_agg_scatter, grid_for(ctx, n, 256), 256, 0, in.chrom, in.start, in.stop, in.coord_bits, (const uint32_t*)cls.p, n, lm, (const uint32_t*)truns.rstart.p, (const uint32_t*)truns.rend.p,
             n_runs, (const uint32_t*)rtbl.p, tshift, nb, plan, cc, ra);
    }
    if (nr) LAUNCH(ctx, k_run_agg_finish, grid_for(ctx, nr, 256), 256, 0, nr, plan, ra, (const int32_t*)truns.cnt.p, tmp);
  }
  GMQL_REQUIRE(done || !shard, GMQL_ERR_UNSUPPORTED, "gmql_cover_shard needs the tile-resident path: COVER or FLAT without aggregates, one (group, strand) class, a linear genome below 2^32, no zero-length regions, regions shorter than 65 534 bases, and no tile denser than the kernel's tables");
  if (!done) {
  // interval index over the ORIGINAL coordinates (phase 2 contributors, GMAP4 uses the un-extended inputs)
  IntervalIndex<K> ix;
  const bool payload_len = plan.n_cols == 0 && spans.G < 0xffffffffull;
  const bool fused_keys = payload_len && !has_zero && !binwise && n > 0;
  DevBuf<K> ek_a(ctx, !binwise ? n : 0);
  if (fused_keys) {
    alloc_index<K>(ctx, n, &ix);
    LAUNCH(ctx, (k_cover_all_keys<K>), grid_for(ctx, n, 256), 256, 0, in.chrom, in.start, in.stop, in.coord_bits, cls.p, n, lm, cfg.B, ix.start_a.p, ix.row_a.p, ek_a.p);
    finish_index<K>(ctx, in, cls.p, n_classes_total, lm, &ix, true);
  } else {
    build_index<K>(ctx, in, cls.p, n_classes_total, lm, &ix, payload_len);
  }

  if (n == 0) {
    p1.n = 0;
  } else if (binwise) {
    phase1_summit(ctx, in, cls.p, spans, cfg, n_classes_total, &p1, flag == GMQL_SUMMIT ? BINS_SUMMIT : (flag == GMQL_HISTOGRAM ? BINS_HISTOGRAM : BINS_COVER));
  } else {
    // without zero-length regions the phase-1 start keys are exactly the index's sorted starts: reuse them
    phase1_sweep<K>(ctx, in, cls.p, lm, cfg, flag, n_classes_total, has_zero ? (const K*)nullptr : ix.start, ek_a, fused_keys, &p1, &lin_s, &lin_e);
    // many small phase-1 regions without value aggregates stay linear all the way (merge form, direct finish); everything else
    // gets its (class, chrom, start, stop) columns now
    const char* g4o = ctx_opt(ctx, "cover.gmap4");
    const char* fino = ctx_opt(ctx, "cover.finish");
    lin_form = p1.n > 0 && p1.n >= n / 8 && p1.n < (int64_t)0xfffffff0LL && plan.n_cols == 0 && !shard && !(g4o && !strcmp(g4o, "scatter")) && !(fino && !strcmp(fino, "split"));
    if (!lin_form) p1_regions<K>(ctx, lin_s, lin_e, lm, &p1);
  }

  // phase 2
  nc = p1.n;
  keep.alloc(ctx, nc);
  if (!lin_form) {                                         // (the linear form writes the result columns directly)
    ocount.alloc(ctx, nc);
    fstart.alloc(ctx, nc); fstop.alloc(ctx, nc);
    j1.alloc(ctx, nc); j2.alloc(ctx, nc);
  }
  CovCols cc;
  memset(&cc, 0, sizeof cc);
  for (int q = 0; q < plan.n_cols; ++q) { cc.v[q] = in.cols[plan.col_id[q]]; cc.ok[q] = in.valid[plan.col_id[q]]; }
  if (!lin_form) for (int a = 0; a < plan.n_aggs; ++a) { tv[a].alloc(ctx, nc); tk[a].alloc(ctx, nc); tmp.value[a] = tv[a].p; tmp.valid[a] = tk[a].p; }
  if (nc) {
    int64_t warps_needed = nc;
    int blocks = (int)std::min<int64_t>(ceil_div64(warps_needed, 8), (int64_t)ctx->num_sms * 16);
    if (nc >= n / 8 && nc < (int64_t)0xfffffff0LL) {
      // many small phase-1 regions: region-driven accumulation; the warp-per-region kernel only runs if the phase-1 list
      // turns out not to be sorted and disjoint (decided on the device, no host round trip)
      DevBuf<K> cs_own, ce_own;
      m_unsorted.alloc(ctx, 1);
      DevBuf<int>& unsorted = m_unsorted;
      unsorted.zero();
      if (lin_form) { lin_check.alloc(ctx, 1); lin_check.zero(); }
      else {
        cs_own.alloc(ctx, nc); ce_own.alloc(ctx, nc);
        LAUNCH(ctx, (k_p1_linear<K>), grid_for(ctx, nc, 256), 256, 0, p1.cls.p, p1.chrom.p, p1.start.p, p1.stop.p, nc, lm, cs_own.p, ce_own.p, unsorted.p);
      }
      struct { K* p; } cs = {lin_form ? lin_s.p : cs_own.p}, ce = {lin_form ? lin_e.p : ce_own.p};
      const uint64_t max_key = n_classes_total * lm.G;
      int tshift = 0;
      { int64_t want = std::max<int64_t>(1024, std::min<int64_t>((int64_t)1 << 22, nc / 2 + 1)); while (tshift < 62 && (int64_t)(max_key >> tshift) + 2 > want) ++tshift; }
      const int64_t nb = (int64_t)(max_key >> tshift) + 1;
      DevBuf<uint32_t> tbl(ctx, nb + 2);
      LAUNCH(ctx, (k_p1_table<K>), grid_for(ctx, nb + 1, 256), 256, 0, cs.p, nc, tshift, nb, tbl.p);
      ScatterAcc<K> sa;
      memset(&sa, 0, sizeof sa);
      m_cnt.alloc(ctx, nc); m_smin.alloc(ctx, nc); m_emax.alloc(ctx, nc); m_smax.alloc(ctx, nc); m_emin.alloc(ctx, nc);
      DevBuf<int32_t>& a_cnt = m_cnt;
      DevBuf<K>&a_smin = m_smin, &a_emax = m_emax, &a_smax = m_smax, &a_emin = m_emin;
      // without value aggregates the merge kernel writes count and bounds of every phase-1 region with plain stores
      const char* g4 = ctx_opt(ctx, "cover.gmap4");           // "scatter": test hook (the RED form)
      const bool merge_form = plan.n_cols == 0 && !(g4 && !strcmp(g4, "scatter"));
      if (!merge_form) { a_cnt.zero(); a_emax.zero(); a_smax.zero(); a_smin.fill_byte(0xff); a_emin.fill_byte(0xff); }
      sa.cnt = a_cnt.p; sa.smin = a_smin.p; sa.emax = a_emax.p; sa.smax = a_smax.p; sa.emin = a_emin.p;
      std::vector<DevBuf<int32_t>> a_nn((size_t)plan.n_cols);
      std::vector<DevBuf<double>> a_sum((size_t)plan.n_cols);
      std::vector<DevBuf<unsigned long long>> a_vmin((size_t)plan.n_cols), a_vmax((size_t)plan.n_cols);
      for (int q = 0; q < plan.n_cols; ++q) {
        a_nn[q].alloc(ctx, nc); a_nn[q].zero(); a_sum[q].alloc(ctx, nc); a_sum[q].zero();
        a_vmin[q].alloc(ctx, nc); a_vmin[q].fill_byte(0xff); a_vmax[q].alloc(ctx, nc); a_vmax[q].zero();
        sa.nn[q] = a_nn[q].p; sa.sum[q] = a_sum[q].p; sa.vmin[q] = a_vmin[q].p; sa.vmax[q] = a_vmax[q].p;
      }
      if (merge_form)
        LAUNCH(ctx, (k_gmap4_merge<K>), (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(nc, GM_CHUNK), (int64_t)ctx->num_sms * 4)), GM_T, 0,
               (const K*)ix.start, (const K*)ix.stop.p, (const K*)ix.pmax.p, n, (const K*)cs.p, (const K*)ce.p, nc, sa, (const int*)unsorted.p, lin_check.p);
      else
      LAUNCH(ctx, (k_gmap4_scatter<K>), grid_for(ctx, n, 256), 256, 0, (const K*)ix.start, (const K*)ix.stop.p, (const uint32_t*)ix.row, n, (const K*)cs.p, (const K*)ce.p, nc,
             (const uint32_t*)tbl.p, tshift, nb, plan, cc, sa, (const int*)unsorted.p);
      const char* fin = ctx_opt(ctx, "cover.finish");         // "split": test hook (k_gmap4_finish + k_cover_compact)
      m_direct = merge_form && !shard && !(fin && !strcmp(fin, "split"));
      if (m_direct) LAUNCH(ctx, k_keep_from_cnt, grid_for(ctx, nc, 256), 256, 0, (const int32_t*)a_cnt.p, nc, keep.p, (const int*)unsorted.p);
      else
      LAUNCH(ctx, (k_gmap4_finish<K>), grid_for(ctx, nc, 256), 256, 0, p1.cls.p, p1.chrom.p, p1.start.p, p1.stop.p, nc, lm, plan, sa, flag == GMQL_FLAT ? 1 : 0,
             keep.p, fstart.p, fstop.p, j1.p, j2.p, ocount.p, tmp, (const int*)unsorted.p);
      if (!lin_form)
      LAUNCH(ctx, (k_gmap4<K>), blocks, 256, 0, p1.cls.p, p1.chrom.p, p1.start.p, p1.stop.p, nc, lm, ix.start, ix.stop.p, ix.pmax.p, ix.row, n,
             plan, cc, flag == GMQL_FLAT ? 1 : 0, keep.p, fstart.p, fstop.p, j1.p, j2.p, ocount.p, tmp, (const int*)unsorted.p);
    } else {
      LAUNCH(ctx, (k_gmap4<K>), blocks, 256, 0, p1.cls.p, p1.chrom.p, p1.start.p, p1.stop.p, nc, lm, ix.start, ix.stop.p, ix.pmax.p, ix.row, n,
             plan, cc, flag == GMQL_FLAT ? 1 : 0, keep.p, fstart.p, fstop.p, j1.p, j2.p, ocount.p, tmp, (const int*)nullptr);
    }
  }
  }  // !done
  int64_t n_out;
  DevBuf<int32_t> o_group, o_chrom, o_acc;
  DevBuf<int64_t> o_start, o_stop;
  DevBuf<int8_t> o_strand;
  DevBuf<double> o_j1, o_j2, o_accd;
  if (fo.ready) {
    // the fused stitch kernel has already written the result rows
    n_out = fo.n_out;
    o_group = std::move(fo.group); o_chrom = std::move(fo.chrom); o_acc = std::move(fo.acc); o_start = std::move(fo.start); o_stop = std::move(fo.stop);
    o_strand = std::move(fo.strand); o_j1 = std::move(fo.j1); o_j2 = std::move(fo.j2); o_accd = std::move(fo.accd);
    o_group.n = o_chrom.n = o_acc.n = o_start.n = o_stop.n = o_strand.n = o_j1.n = o_j2.n = n_out;
  } else {
    keep_excl.alloc(ctx, nc + 1);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, keep.p, keep_excl.p, nc, keep_excl.p + nc);
    if (m_direct) {                                         // (one synchronisation for all three words)
      int h_uns = 0, h_chk = 0;
      CUDA_TRY(cudaMemcpyAsync(&h_uns, m_unsorted.p, sizeof h_uns, cudaMemcpyDeviceToHost, ctx->stream));
      if (lin_check.p) CUDA_TRY(cudaMemcpyAsync(&h_chk, lin_check.p, sizeof h_chk, cudaMemcpyDeviceToHost, ctx->stream));
      n_out = read_back<int64_t>(ctx, keep_excl.p + nc);
      GMQL_REQUIRE(!h_chk, GMQL_ERR_INVALID, "internal: the sweep's phase-1 regions are not sorted and disjoint");
      if (h_uns) m_direct = false;                          // the generic kernel filled the intermediate arrays: compact them
    } else
    n_out = read_back<int64_t>(ctx, keep_excl.p + nc);
    o_group.alloc(ctx, n_out); o_chrom.alloc(ctx, n_out); o_acc.alloc(ctx, n_out); o_start.alloc(ctx, n_out); o_stop.alloc(ctx, n_out);
    o_strand.alloc(ctx, n_out); o_j1.alloc(ctx, n_out); o_j2.alloc(ctx, n_out); o_accd.alloc(ctx, n_out > 0 ? n_out : 1);
  }
  std::vector<DevBuf<double>> ov((size_t)plan.n_aggs);
  std::vector<DevBuf<uint8_t>> ok((size_t)plan.n_aggs);
  FinalCols f;
  memset(&f, 0, sizeof f);
  f.group = o_group.p; f.chrom = o_chrom.p; f.start = o_start.p; f.stop = o_stop.p; f.strand = o_strand.p; f.acc = o_acc.p; f.j1 = o_j1.p; f.j2 = o_j2.p;
  for (int a = 0; a < plan.n_aggs; ++a) { ov[a].alloc(ctx, n_out); ok[a].alloc(ctx, n_out); f.value[a] = ov[a].p; f.valid[a] = ok[a].p; }
  ShardBufs so;
  tiled::ShardRaw raw_in, raw_out;
  memset(&raw_in, 0, sizeof raw_in); memset(&raw_out, 0, sizeof raw_out);
  if (shard) {
    so.edge.alloc(ctx, n_out); so.cnt.alloc(ctx, n_out); so.rstart.alloc(ctx, n_out); so.rstop.alloc(ctx, n_out);
    so.smin.alloc(ctx, n_out); so.emax.alloc(ctx, n_out); so.smax.alloc(ctx, n_out); so.emin.alloc(ctx, n_out);
    raw_in.edge = sbufs.edge.p; raw_in.cnt = sbufs.cnt.p; raw_in.rstart = sbufs.rstart.p; raw_in.rstop = sbufs.rstop.p;
    raw_in.smin = sbufs.smin.p; raw_in.emax = sbufs.emax.p; raw_in.smax = sbufs.smax.p; raw_in.emin = sbufs.emin.p;
    raw_out.edge = so.edge.p; raw_out.cnt = so.cnt.p; raw_out.rstart = so.rstart.p; raw_out.rstop = so.rstop.p;
    raw_out.smin = so.smin.p; raw_out.emax = so.emax.p; raw_out.smax = so.smax.p; raw_out.emin = so.emin.p;
  }
  if (nc && n_out && !fo.ready && m_direct) {
    ScatterAcc<K> sa;
    memset(&sa, 0, sizeof sa);
    sa.cnt = m_cnt.p; sa.smin = m_smin.p; sa.emax = m_emax.p; sa.smax = m_smax.p; sa.emin = m_emin.p;
    LAUNCH(ctx, (k_gmap4_finish_direct<K>), grid_for(ctx, nc, 256), 256, 0, (const int32_t*)keep.p, (const int64_t*)keep_excl.p, (const uint32_t*)p1.cls.p, (const int32_t*)p1.chrom.p,
           (const int64_t*)p1.start.p, (const int64_t*)p1.stop.p, (const int32_t*)p1.acc.p, nc, lm, plan, sa, flag == GMQL_FLAT ? 1 : 0, cfg.n_classes, cfg.unified, f, o_accd.p,
           (const K*)(lin_form ? lin_s.p : nullptr), (const K*)(lin_form ? lin_e.p : nullptr));
  } else
  if (nc && n_out && !fo.ready) LAUNCH(ctx, k_cover_compact, grid_for(ctx, nc, 256), 256, 0, nc, keep.p, keep_excl.p, p1.cls.p, p1.chrom.p, fstart.p, fstop.p, p1.acc.p, j1.p, j2.p,
                          tmp, plan.n_aggs, cfg.n_classes, cfg.unified, f, o_accd.p, raw_in, raw_out);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  res->view_acc_double = (void*)o_accd.detach();
  res->view_cols.assign(1, (const double*)res->view_acc_double);
  res->view_valid.assign(1, nullptr);

  res->rows = n_out;
  res->kind = 2;
  res->sorted = n_classes_total == 1;     // one (group, strand) class: the rows are in (chrom, start) order
  res->unified = cfg.unified != 0;
  res->n_chrom = in.n_chrom;
  res->group_ids.resize((size_t)n_groups);
  for (int gi = 0; gi < n_groups; ++gi) res->group_ids[gi] = gi;
  result_add(res, GMQL_COL_COV_GROUP, o_group);
  result_add(res, GMQL_COL_COV_CHROM, o_chrom);
  result_add(res, GMQL_COL_COV_START, o_start);
  result_add(res, GMQL_COL_COV_STOP, o_stop);
  result_add(res, GMQL_COL_COV_STRAND, o_strand);
  result_add(res, GMQL_COL_COV_ACC, o_acc);
  result_add(res, GMQL_COL_COV_JACC_INTERSECT, o_j1);
  result_add(res, GMQL_COL_COV_JACC_RESULT, o_j2);
  if (shard) {
    result_add(res, GMQL_COL_COV_EDGE, so.edge); result_add(res, GMQL_COL_COV_RAW_COUNT, so.cnt);
    result_add(res, GMQL_COL_COV_RAW_START, so.rstart); result_add(res, GMQL_COL_COV_RAW_STOP, so.rstop);
    result_add(res, GMQL_COL_COV_RAW_SMIN, so.smin); result_add(res, GMQL_COL_COV_RAW_EMAX, so.emax);
    result_add(res, GMQL_COL_COV_RAW_SMAX, so.smax); result_add(res, GMQL_COL_COV_RAW_EMIN, so.emin);
  }
  for (int a = 0; a < plan.n_aggs; ++a) {
    result_add(res, GMQL_COL_AGG_VALUE + a, ov[a]);
    result_add(res, GMQL_COL_AGG_VALID + a, ok[a]);
  }
}

struct Inspect { int has_star = 0, has_zero = 0, samples_with_regions = 0; };

static void inspect_input(gmql_ctx* ctx, const DevRegions& in, Inspect* out) {
  DevBuf<int> flags(ctx, 4), seen(ctx, in.n_samples);
  flags.zero();
  seen.zero();
  if (in.n) {
    LAUNCH(ctx, k_cover_inspect, grid_for(ctx, in.n, 256), 256, 0, in.strand, in.sample, in.start, in.stop, in.coord_bits, in.n, in.n_samples, flags.p, seen.p);
    LAUNCH(ctx, k_count_flags, (int)std::min<int64_t>(ceil_div64(in.n_samples, 256), 64), 256, 0, seen.p, in.n_samples, flags.p + 3);
  }
  int h[4];
  CUDA_TRY(cudaMemcpyAsync(h, flags.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(!h[2], GMQL_ERR_INVALID, "region with start > stop, negative start, bad strand code or sample index out of range");
  out->has_star = h[0];
  out->has_zero = h[1];
  out->samples_with_regions = h[3];
}

}  // namespace

extern "C" int gmql_cover_count_samples(gmql_ctx* ctx, const gmql_regions* in, int32_t* out_count) {
  if (!ctx || !in || !out_count) return GMQL_ERR_INVALID;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    // only the sample column matters, but the shared upload keeps the call simple (tiny next to gmql_cover itself)
    DevRegions d;
    gmql_upload_regions(ctx, in, std::vector<int>(), &d);
    Inspect ins;
    inspect_input(ctx, d, &ins);
    *out_count = ins.samples_with_regions;
    return GMQL_OK;
  } catch (const gmql_error& ex) {
    ctx->err = ex.what();
    cudaStreamSynchronize(ctx->stream);
    return ex.code;
  }
}

// a SpanTable whose device arrays belong to somebody else (a resident dataset's layout): the DevBufs are pointed at them for the
// duration of the call and detached again, so nothing is freed twice
struct LinMapBorrow {
  SpanTable* t;
  LinMapBorrow(SpanTable* tab, uint64_t* coff, unsigned long long* span) : t(tab) { t->coff.p = coff; t->span.p = span; }
  ~LinMapBorrow() { t->coff.detach(); t->span.detach(); }
};

// The linear layout of a resident dataset for BinSize.Cover = bin_size and what the tile path's partition needs before it can
// start: tile_off (first record of every 65 536-position tile) and coarse_off (start of every (coarse bin, input-tile group) range).
// Called by gmql_dataset_upload, and again by gmql_cover when an operator uses another bin size.
void gmql_cover_prepare_dataset(gmql_ctx* ctx, gmql_dataset* ds, int64_t bin_size) {
  const DevRegions& dv = ds->dev;
  ds->tile_off.release(); ds->coarse_off.release(); ds->lay_coff.release(); ds->lay_span.release();
  ds->lay_pad = 0; ds->lay_G = 0; ds->hist_nT = 0;
  if (!ds->stats_ok || ds->span_hint.empty() || dv.n <= 0) return;
  const int nc = dv.n_chrom;
  const uint64_t pad = (uint64_t)bin_size + 2ull;
  unsigned long long* hs = (unsigned long long*)ctx_pinned(ctx, 8 * (size_t)nc);
  uint64_t* hc = (uint64_t*)ctx_pinned(ctx, 8 * ((size_t)nc + 1));
  uint64_t run = 0;
  for (int c = 0; c < nc; ++c) { hs[c] = (unsigned long long)ds->span_hint[c]; hc[c] = run; run += hs[c] + pad; }
  hc[nc] = run;
  ds->lay_coff.alloc(ctx, nc + 1); ds->lay_span.alloc(ctx, nc);
  CUDA_TRY(cudaMemcpyAsync(ds->lay_span.p, hs, 8 * (size_t)nc, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ds->lay_coff.p, hc, 8 * ((size_t)nc + 1), cudaMemcpyHostToDevice, ctx->stream));
  ds->lay_pad = pad; ds->lay_G = run;
  // the partition offsets only for what the tile path's fast route accepts
  const int logw = tiled::MAX_LOGW;
  const int64_t nT = (int64_t)(run >> logw) + 1;
  if (!(dv.coord_bits == 32 && dv.chrom_narrow && dv.chrom_bits == 8 && dv.len_narrow && dv.len_bits == 16 && ds->has_star && !ds->has_zero &&
        ds->maxlen < (1ll << tiled::MAX_LOGW) - 2 && run + 2ull * (1ull << logw) < 0xffffffffull && nT <= tiled::HIST_CAP)) return;
  const int n_coarse = (int)((nT + 255) >> 8) + 1;
  const int64_t in_tiles = ceil_div64(dv.n, tiled::MSP_TILE);
  const int G = (int)std::max<int64_t>(1, std::min<int64_t>(in_tiles, (int64_t)ctx->num_sms));
  ds->tile_off.alloc(ctx, nT + 1); ds->coarse_off.alloc(ctx, (int64_t)n_coarse * G + 1);
  LinMap lm; lm.coff = ds->lay_coff.p; lm.span = (const int64_t*)ds->lay_span.p; lm.G = run; lm.n_chrom = nc;
  tile_offsets_narrow(ctx, dv, lm, logw, nT, G, n_coarse, ds->tile_off.p, ds->coarse_off.p);
  ds->hist_nT = nT; ds->hist_groups = G; ds->hist_n_coarse = n_coarse;
}

// BAG / BAGD in COVER (DefaultRegionsToRegionFactory.scala:128-168 through GMAP4.scala:57-66,85): the values of the regions that
// contribute to an output region, as text — strings are host business, so the device lists the contributing input rows of every
// output region and the host folds them.  A contributor of an output region is an input region of the same group that overlaps it
// with a compatible strand (GMAP4.scala:38-45; the first-bin clause only makes every pair count once): exactly MAP of the input,
// every sample relabelled with its group, onto the COVER result (whose "sample" is the group).  One internal gmql_map call with a
// BAG aggregate produces the lists; output row r of COVER is row r of that MAP (rows come out by group, then in row order).
__global__ void k_group_of_row(const int32_t* __restrict__ sample, const int32_t* __restrict__ sample_group, int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = sample_group[sample ? sample[i] : 0];
}

static void attach_bag_lists(gmql_ctx* ctx, DevRegions& d, const int32_t* h_sample_group, int n_groups, gmql_result* res, gmql_result* runs) {
  // `runs`: the rows whose coordinates define the contributors — the result itself, except for FLAT, whose rows carry the extent
  // of their contributors instead of the run they were found with (GMAP4.scala:83-84): there the same call with flag COVER
  if (res->rows == 0) return;
  gmql_ensure_wide(ctx, &d);
  gmql_regions rview;
  int rc = gmql_result_as_regions(ctx, runs, &rview);
  GMQL_REQUIRE(rc == GMQL_OK, rc, ctx->err);
  DevBuf<int32_t> grp, d_sg;
  if (h_sample_group && n_groups > 1) {
    grp.alloc(ctx, d.n); d_sg.alloc(ctx, d.n_samples);
    int32_t* h = (int32_t*)ctx_pinned(ctx, 4 * (size_t)d.n_samples);
    memcpy(h, h_sample_group, 4 * (size_t)d.n_samples);
    CUDA_TRY(cudaMemcpyAsync(d_sg.p, h, 4 * (size_t)d.n_samples, cudaMemcpyHostToDevice, ctx->stream));
    if (d.n) LAUNCH(ctx, k_group_of_row, grid_for(ctx, d.n, 256), 256, 0, d.sample, (const int32_t*)d_sg.p, d.n, grp.p);
    CUDA_TRY(GMQL_SYNC(ctx));            // the pinned table is recycled by the nested call below
  }
  std::vector<int64_t> gids((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) gids[g] = g;
  gmql_regions ev;
  memset(&ev, 0, sizeof ev);
  ev.n = d.n; ev.location = GMQL_LOC_DEVICE; ev.coord_bits = d.coord_bits;
  ev.chrom = d.chrom; ev.start = d.start; ev.stop = d.stop;
  ev.strand = res->unified ? (const int8_t*)nullptr : d.strand;     // unified strands: every input region counts as '*' (assignGroups, GenometricCover.scala:328-341)
  ev.sample = grp.p;                                                 // null with one group: every row is sample 0
  ev.n_chrom = d.n_chrom; ev.n_samples = n_groups; ev.sample_ids = gids.data();
  std::vector<gmql_pair> pairs((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) { pairs[g].left_sample = g; pairs[g].right_sample = g; }
  gmql_agg bag; bag.kind = GMQL_AGG_BAG; bag.col = 0;
  gmql_result* mp = nullptr;
  rc = gmql_map(ctx, &rview, &ev, pairs.data(), n_groups, &bag, 1, &mp);
  GMQL_REQUIRE(rc == GMQL_OK, rc, ctx->err);
  GMQL_REQUIRE(gmql_result_rows(mp) == res->rows, GMQL_ERR_INVALID, "internal: contributor lists do not line up with the COVER rows");
  for (int id : {GMQL_COL_BAG_OFFSETS, GMQL_COL_BAG_ROWS}) {
    auto it = mp->cols.find(id);
    GMQL_REQUIRE(it != mp->cols.end(), GMQL_ERR_INVALID, "internal: gmql_map returned no contributor lists");
    res->cols[id] = it->second;
    mp->cols.erase(it);
  }
  gmql_result_free(ctx, mp);
}

static int cover_impl(gmql_ctx* ctx, const gmql_regions* in, const int32_t* sample_group, int32_t n_groups, int32_t flag,
                      const int32_t* min_acc, const int32_t* max_acc, int64_t bin_size, const gmql_agg* aggs, int32_t n_aggs,
                      const gmql_cover_shard_t* shard, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    gmql_trace("ENTER gmql_cover");
    GMQL_REQUIRE(in != nullptr, GMQL_ERR_INVALID, "null dataset");
    GMQL_REQUIRE(n_groups >= 1 && min_acc && max_acc, GMQL_ERR_INVALID, "n_groups >= 1 and per-group thresholds are required");
    const bool force_unified = (flag & GMQL_COVER_FORCE_UNIFIED) != 0;
    flag &= ~GMQL_COVER_FORCE_UNIFIED;
    GMQL_REQUIRE(flag >= GMQL_COVER && flag <= GMQL_HISTOGRAM, GMQL_ERR_INVALID, "unknown cover flag");
    GMQL_REQUIRE(bin_size > 0, GMQL_ERR_INVALID, "bin_size must be positive (BinSize.Cover = 2000)");
    bool low_min = false;                                  // minAcc < 1: depth 0 is "in range" between the events of a bin (bin-wise path only)
    for (int g = 0; g < n_groups; ++g) if (min_acc[g] < 1) low_min = true;
    if (const char* f = ctx_opt(ctx, "cover.path")) if (!strcmp(f, "bins")) low_min = true;
    if (sample_group)
      for (int s = 0; s < in->n_samples; ++s)
        GMQL_REQUIRE(sample_group[s] >= 0 && sample_group[s] < n_groups, GMQL_ERR_INVALID, "sample_group entry out of range");
    CovAggPlan plan;
    std::vector<int> need_cols;
    plan_cov_aggs(aggs, n_aggs, in->n_cols, &plan, &need_cols);
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions d;
    gmql_upload_regions(ctx, in, need_cols, &d, /*keep_narrow=*/true);
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    // Resident dataset in the narrow encodings (uint8 chromosome, int32 start, uint16 length; strands unified) whose upload-time
    // statistics say it is eligible for the tile path: no pass over the data is needed to decide anything, the linear layout and
    // the partition offsets come from the dataset, and the operator starts at the partition itself, reading 7 bytes per region.
    {
      gmql_dataset* ds = d.dataset;
      const char* spec_opt = ctx_opt(ctx, "cover.speculate");   // "0": test hook (the generic route)
      if (ds && ds->stats_ok && !ds->span_hint.empty() && in->chrom_span_hint == ds->span_hint.data() && !shard && d.n > 0 && d.coord_bits == 32 && d.chrom_narrow &&
          d.chrom_bits == 8 && d.len_narrow && d.len_bits == 16 && ds->has_star /* strands unify to '*': the column, if any, is not read */ && n_groups == 1 && plan.n_aggs == 0 && (flag == GMQL_COVER || flag == GMQL_FLAT) &&
          !ds->has_zero && ds->maxlen < (1ll << tiled::MAX_LOGW) - 2 && !low_min && !(spec_opt && spec_opt[0] == '0')) {
        if (ds->lay_pad != (uint64_t)bin_size + 2ull) gmql_cover_prepare_dataset(ctx, ds, bin_size);   // another BinSize.Cover than the upload assumed
        if (ds->lay_coff.p && ds->lay_G + 2ull * (1ull << tiled::MAX_LOGW) < 0xffffffffull) {
          SpanTable hs;                                       // borrowed views of the dataset's layout
          hs.n_chrom = d.n_chrom; hs.G = ds->lay_G;
          LinMapBorrow lb(&hs, ds->lay_coff.p, ds->lay_span.p);
          res = new gmql_result();
          res->owner = ctx;
          TileSource src;
          src.narrow = true; src.ds = ds;
          run_cover<uint32_t>(ctx, d, sample_group, n_groups, flag, min_acc, max_acc, bin_size, plan, hs, /*unified=*/true, /*has_zero=*/false, ds->maxlen, res, nullptr, nullptr, src);
          res->span_hint.resize((size_t)d.n_chrom);
          for (int c = 0; c < d.n_chrom; ++c) res->span_hint[c] = ds->span_hint[c] + 1;   // results may carry the one-base extension
          CUDA_TRY(GMQL_SYNC_IF_PENDING(ctx));
          *out = res;
          gmql_trace("EXIT");
          return GMQL_OK;
        }
      }
    }
    gmql_ensure_wide(ctx, &d);
    // Speculative launch: the caller knows an upper bound of every chromosome's span (chrom_span_hint), so the linear layout is
    // fixed before any data is read and the tile pipeline is queued at once behind a fused prepass + histogram kernel; the
    // prepass flags come back with the pipeline's one synchronisation.  One (group, strand) class, 32-bit coordinates, no
    // strand column, COVER / FLAT without aggregates; anything else, and any input that turns out not to be eligible, takes
    // the ordinary route below.
    const char* spec_env = ctx_opt(ctx, "cover.speculate");   // "0": test hook
    if (in->chrom_span_hint && !shard && d.n > 0 && d.coord_bits == 32 && d.strand == nullptr && n_groups == 1 && plan.n_aggs == 0 &&
        (flag == GMQL_COVER || flag == GMQL_FLAT) && !low_min && !(spec_env && spec_env[0] == '0')) {
      SpanTable hs;
      span_table_from_hint(ctx, in->chrom_span_hint, d.n_chrom, &hs, (unsigned long long)bin_size + 2ull);
      if (hs.G + 2ull * (1ull << tiled::MAX_LOGW) < 0xffffffffull) {
        DevBuf<int> sflags(ctx, 4);
        DevBuf<unsigned long long> act(ctx, d.n_chrom);
        sflags.zero(); act.zero();
        SpecCtx sc;
        sc.flags = sflags.p; sc.act_span = act.p; sc.n_chrom = d.n_chrom;
        sc.h_flags = (int*)ctx_pinned(ctx, 4 * sizeof(int));
        sc.h_span = (unsigned long long*)ctx_pinned(ctx, 8 * (size_t)d.n_chrom);
        res = new gmql_result();
    res->owner = ctx;
        run_cover<uint32_t>(ctx, d, sample_group, n_groups, flag, min_acc, max_acc, bin_size, plan, hs, /*unified=*/true, /*has_zero=*/false, /*maxlen=*/0, res, nullptr, &sc);
        if (!sc.declined && !sc.tripped) {
          res->span_hint.resize((size_t)d.n_chrom);
          for (int c = 0; c < d.n_chrom; ++c) res->span_hint[c] = (int64_t)sc.h_span[c] + 1;   // results may carry the one-base extension
          CUDA_TRY(GMQL_SYNC_IF_PENDING(ctx));
          *out = res;
          gmql_trace("EXIT");
          return GMQL_OK;
        }
        gmql_result_free(ctx, res);                       // nothing (declined) or a guarded, empty pipeline (tripped) has run
        res = nullptr;
      }
    }
    // one read of the columns: chromosome spans + input flags
    SpanTable spans;
    spans.n_chrom = d.n_chrom;
    spans.span.alloc(ctx, d.n_chrom); spans.span.zero();
    spans.coff.alloc(ctx, d.n_chrom + 1);
    DevBuf<int> flags(ctx, 4);
    flags.zero();
    if (d.n) {
      int nb = (int)std::min<int64_t>(ceil_div64(d.n, 256 * 16), (int64_t)ctx->num_sms * 8);
      size_t sh = d.n_chrom <= 4096 ? sizeof(unsigned long long) * (size_t)d.n_chrom : 0;
      auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
      int vec_ok = al16(d.chrom) && al16(d.start) && al16(d.stop) && al16(d.sample) && ((uintptr_t)d.strand & 3u) == 0;
      LAUNCH(ctx, k_cover_prepass, nb, 256, sh, d.chrom, d.start, d.stop, d.coord_bits, d.strand, d.sample, d.n, d.n_chrom, d.n_samples, spans.span.p, flags.p, vec_ok);
    }
    LAUNCH(ctx, k_chrom_offsets, 1, 32, 0, spans.span.p, d.n_chrom, spans.coff.p, (unsigned long long)bin_size + 2ull);
    int hf[4];
    uint64_t G = 0;
    CUDA_TRY(cudaMemcpyAsync(hf, flags.p, sizeof hf, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&G, spans.coff.p + d.n_chrom, sizeof G, cudaMemcpyDeviceToHost, ctx->stream));
    unsigned long long* h_span = (unsigned long long*)ctx_pinned(ctx, 8 * (size_t)d.n_chrom);
    CUDA_TRY(cudaMemcpyAsync(h_span, spans.span.p, 8 * (size_t)d.n_chrom, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
    GMQL_REQUIRE(!hf[2], GMQL_ERR_INVALID, "region with chromosome/sample index out of range, start > stop, negative start or bad strand code");
    spans.G = G;
    ShardSpec sspec;
    if (shard) {
      // every shard must lay the genome out identically: the caller's global spans replace the ones seen in this shard's rows
      GMQL_REQUIRE(shard->chrom_span != nullptr, GMQL_ERR_INVALID, "gmql_cover_shard: chrom_span is required");
      unsigned long long* gs = (unsigned long long*)ctx_pinned(ctx, 8 * (size_t)d.n_chrom);
      uint64_t run = 0;
      for (int c = 0; c < d.n_chrom; ++c) {
        GMQL_REQUIRE(shard->chrom_span[c] >= 0 && (unsigned long long)shard->chrom_span[c] >= h_span[c], GMQL_ERR_INVALID, "gmql_cover_shard: a region lies beyond its chromosome's chrom_span");
        gs[c] = (unsigned long long)shard->chrom_span[c];
        h_span[c] = gs[c];
        run += gs[c] + (unsigned long long)bin_size + 2ull;
      }
      CUDA_TRY(cudaMemcpyAsync(spans.span.p, gs, 8 * (size_t)d.n_chrom, cudaMemcpyHostToDevice, ctx->stream));
      LAUNCH(ctx, k_chrom_offsets, 1, 32, 0, spans.span.p, d.n_chrom, spans.coff.p, (unsigned long long)bin_size + 2ull);
      spans.G = run;
      sspec.first_tile = shard->first_tile; sspec.end_tile = shard->end_tile;
    }
    Inspect ins;
    ins.has_star = hf[0];
    ins.has_zero = hf[1];
    bool unified = force_unified || ins.has_star != 0 || d.strand == nullptr || d.n == 0;   // assignGroups (:328-341)
    uint64_t total = (uint64_t)n_groups * (unified ? 1ull : 2ull) * spans.G;
    res = new gmql_result();
    res->owner = ctx;
    res->span_hint.resize((size_t)d.n_chrom);
    for (int c = 0; c < d.n_chrom; ++c) res->span_hint[c] = (int64_t)h_span[c] + 1;   // results may carry the one-base extension
    if (shard) GMQL_REQUIRE(unified && n_groups == 1 && total < 0xffffffffull, GMQL_ERR_UNSUPPORTED, "gmql_cover_shard: one group, unified strands and a linear genome below 2^32 are required");
    if (total < 0xffffffffull) run_cover<uint32_t>(ctx, d, sample_group, n_groups, flag, min_acc, max_acc, bin_size, plan, spans, unified, ins.has_zero != 0, (int64_t)hf[3], res, shard ? &sspec : nullptr);
    else run_cover<uint64_t>(ctx, d, sample_group, n_groups, flag, min_acc, max_acc, bin_size, plan, spans, unified, ins.has_zero != 0, (int64_t)hf[3], res);
    CUDA_TRY(GMQL_SYNC_IF_PENDING(ctx));
    if (plan.need_bag) {
      GMQL_REQUIRE(!shard, GMQL_ERR_UNSUPPORTED, "gmql_cover_shard has no aggregates");
      gmql_result* runs = res;
      if (flag == GMQL_FLAT) {
        // the runs behind the FLAT rows, in the same order (both drop the runs without contributors)
        int rc2 = cover_impl(ctx, in, sample_group, n_groups, GMQL_COVER | (force_unified ? GMQL_COVER_FORCE_UNIFIED : 0), min_acc, max_acc, bin_size, nullptr, 0, nullptr, &runs);
        GMQL_REQUIRE(rc2 == GMQL_OK, rc2, ctx->err);
        if (runs->rows != res->rows) { gmql_result_free(ctx, runs); throw gmql_error(GMQL_ERR_INVALID, "internal: FLAT rows do not line up with their runs"); }
      }
      try { attach_bag_lists(ctx, d, sample_group, n_groups, res, runs); }
      catch (...) { if (runs != res) gmql_result_free(ctx, runs); throw; }
      if (runs != res) gmql_result_free(ctx, runs);
    }
    *out = res;
    gmql_trace("EXIT");
    return GMQL_OK;
  } catch (const gmql_error& ex) {
    ctx->err = ex.what();
    if (res) gmql_result_free(ctx, res);
    cudaStreamSynchronize(ctx->stream);
    return ex.code;
  } catch (const std::exception& ex) {
    ctx->err = ex.what();
    if (res) gmql_result_free(ctx, res);
    return GMQL_ERR_INVALID;
  }
}

__global__ void k_i32_to_f64(const int32_t* __restrict__ in, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

extern "C" int gmql_result_as_regions(gmql_ctx* ctx, gmql_result* r, gmql_regions* view) {
  if (!ctx || !r || !view) return GMQL_ERR_INVALID;
  if (r->kind != 2) { ctx->err = "gmql_result_as_regions: only COVER results can be viewed as regions"; return GMQL_ERR_INVALID; }
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    int64_t n = r->rows;
    if (!r->view_acc_double) {
      void* p = nullptr;
      CUDA_TRY(cudaMallocAsync(&p, sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->stream));
      r->view_acc_double = p;
      if (n) LAUNCH(ctx, k_i32_to_f64, grid_for(ctx, n, 256), 256, 0, (const int32_t*)r->cols[GMQL_COL_COV_ACC].p, n, (double*)p);
      CUDA_TRY(GMQL_SYNC(ctx));
      r->view_cols.assign(1, (const double*)p);
      r->view_valid.assign(1, nullptr);
    }
    memset(view, 0, sizeof *view);
    view->n = n;
    view->location = GMQL_LOC_DEVICE;
    view->coord_bits = 64;
    view->chrom = (const int32_t*)r->cols[GMQL_COL_COV_CHROM].p;
    view->start = r->cols[GMQL_COL_COV_START].p;
    view->stop = r->cols[GMQL_COL_COV_STOP].p;
    view->strand = r->unified ? (const int8_t*)nullptr : (const int8_t*)r->cols[GMQL_COL_COV_STRAND].p;   // every row '*': same as no strand column
    view->sample = (const int32_t*)r->cols[GMQL_COL_COV_GROUP].p;
    view->n_chrom = r->n_chrom;
    view->n_samples = (int32_t)r->group_ids.size();
    view->sample_ids = r->group_ids.data();
    view->n_cols = 1;
    view->cols = r->view_cols.data();
    view->valid = r->view_valid.data();
    view->dup_of = nullptr;
    view->chrom_span_hint = r->span_hint.empty() ? nullptr : r->span_hint.data();
    view->sorted_by_position = r->sorted ? 1 : 0;
    return GMQL_OK;
  } catch (const gmql_error& ex) {
    ctx->err = ex.what();
    return ex.code;
  }
}

extern "C" int gmql_cover(gmql_ctx* ctx, const gmql_regions* in, const int32_t* sample_group, int32_t n_groups, int32_t flag,
                          const int32_t* min_acc, const int32_t* max_acc, int64_t bin_size, const gmql_agg* aggs, int32_t n_aggs, gmql_result** out) {
  return cover_impl(ctx, in, sample_group, n_groups, flag, min_acc, max_acc, bin_size, aggs, n_aggs, nullptr, out);
}

extern "C" int gmql_cover_shard(gmql_ctx* ctx, const gmql_regions* in, int32_t flag, int32_t min_acc, int32_t max_acc, int64_t bin_size,
                                const gmql_cover_shard_t* shard, gmql_result** out) {
  if (!shard) return GMQL_ERR_INVALID;
  return cover_impl(ctx, in, nullptr, 1, flag, &min_acc, &max_acc, bin_size, nullptr, 0, shard, out);
}
