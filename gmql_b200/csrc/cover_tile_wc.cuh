// cover_tile_wc.cuh — the word-count tile kernel of the tile-resident COVER / FLAT path (included by cover_tiled.cuh, namespace tiled).
//
// Same contract as k_cover_tile (the bitmap-rank kernel): a block owns a contiguous range of tiles, reads the 4-byte records of
// every tile, and publishes the tile's PIECES (in-range stretches clipped at tile edges) with their GMAP4 accumulators; the stitch
// kernels do the rest.  What differs is how the depth profile is obtained.  The rank kernel sorts the tile's distinct event
// positions (occupancy bitmap + popcount prefix), sums +1 / -1 per rank and scans the ranks: ~250 instructions and ~2 shared-memory
// wavefronts per region, most of them random (ncu, round 1: 59 % of the wavefronts were bank conflicts).  Here only COUNTS per
// 32-position word are kept:
//   * phase A  every region adds 1 to the start count of the word of its start and 1 to the stop count of the word of its stop'
//              (two shared-memory atomics per region, the only random read-modify-writes of the kernel);
//   * phase B  a prefix sum over the words gives the EXACT depth at every word boundary.  Inside a word the depth stays within
//              [d0 - nStops, d0 + nStarts]: a word whose bounds lie inside [min, max] holds no transition (IN), one whose bounds lie
//              on one side outside holds none either (OUT).  Only the remaining words, and the IN words whose upper bound exceeds the
//              largest boundary depth of their stretch of consecutive IN words (AccIndex = the run's maximum depth must be exact),
//              are resolved position by position — a few words per tile under dense coverage, a few per run boundary otherwise
//              (tools/tile_count_model.py is the host model this was validated with);
//   * phase C  the regions are visited once more: those with an event in a word to be resolved add it to that word's 32 per-position
//              net counters (two shared-memory LOADS per region; atomics only for the few hits);
//   * phase D  one warp per resolved word: lane p owns position p, a warp prefix sum gives the depth after every position, ballots
//              give the run opens / closes; piece indices come from a prefix sum over the resolved words' open counts.
// Coincident events are summed per position before the depth is judged, exactly as the reference's reduceByKey does
// (GenometricCover.scala:98-108).  GMAP4 (phase E) is region-driven as in the rank kernel.
// Capacity: WC_NCAP resolved words per tile; a tile beyond that raises fail = 5 and the caller reruns the operator with the rank
// kernel on the same partitioned records.
#pragma once

constexpr int WC_NCAP = 384;                    // words of one tile resolved position by position
constexpr int WC_SEGCAP = MAXWORDS / 2 + 2;     // stretches of consecutive IN words of one tile
constexpr int WC_MASKW = MAXWORDS / 32;         // words of the "to be resolved" bit mask
constexpr size_t WC_SMEM_BYTES = 4 * (size_t)MAXWORDS + 64 * (size_t)WC_NCAP + 2 * 4 * (size_t)HCAP + 32 * (size_t)MAXP + 4 * (size_t)WC_SEGCAP +
                                 2 * (size_t)WC_SEGCAP + 2 * 3 * (size_t)(WC_NCAP + 2) + 4 * (size_t)WC_NCAP + 2 * 4 * (size_t)WC_MASKW + 16;
static_assert(WC_SMEM_BYTES <= 56 * 1024, "four blocks per SM");

// the extractor's one-base extension of a region that starts at linear position rs (GenometricCover.scala:353-354)
__device__ __forceinline__ uint32_t wc_ext(uint32_t rs, uint32_t len, const TileChrom& tc, const LinMap& lm, const BinMod32& bm) {
  const uint32_t base = (rs >= tc.lo && rs < tc.next) ? tc.base : lin_base_slow(lm, rs);
  return (len != 0u && is_bin_multiple32(bm, rs + len - base)) ? 1u : 0u;
}

// FULL: tiles of 65 536 positions (the production width): 2 048 words, eight per thread, every word valid — the per-word loops then
// carry no bounds predicates.
// XB: the partition already decided the one-base extension and left it in bit 15 of the record (lengths below 32 768, the host
// knows the longest region): the tile kernel then never looks at chromosome offsets or the bin size.
template <bool FULL, bool XB>
static __global__ void __launch_bounds__(THREADS, CTAS_PER_SM) k_cover_tile_wc(const uint32_t* __restrict__ P /* tile-local start << 16 | length */, const uint32_t* __restrict__ offS,
                                                                               int64_t t_lo, int64_t t_hi, int logw, Cfg cfg, LinMap lm, BinMod32 bm, Pieces out, Guard guard) {
  if (guard_trips(guard)) return;
#ifdef GMQL_TILE_TICKS   // profiling build (make NVFLAGS+=-DGMQL_TILE_TICKS; option cover.dbg = 9): cycles per phase, thread 0 of every block
#define TICK(p) do { if (cfg.dbg == 9 && tid == 0) { const long long now_ = clock64(); atomicAdd((unsigned long long*)out.ticks + (p), (unsigned long long)(now_ - tick_last)); tick_last = now_; } } while (0)
  long long tick_last = 0;
#else
#define TICK(p) do { } while (0)
#endif
#define RLEN(pr) (XB ? ((pr) & 0x7fffu) : ((pr) & 0xffffu))
#define REXT(pr, rs, len) (XB ? (((pr) >> 15) & 1u) : wc_ext((rs), (len), tch, lm, bm))
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* cnt = sm;                                      // [MAXWORDS] starts | stops' << 16 per word
  uint32_t* npos = cnt + MAXWORDS;                         // [WC_NCAP * 16] net change per position of the resolved words (16 bits each)
  uint32_t* hin = npos + WC_NCAP * 16;                     // [HCAP] records of the regions still open at the tile start (they start one tile earlier)
  uint32_t* hout = hin + HCAP;                             // [HCAP] records of this tile's regions still open at its end
  uint32_t* pst = hout + HCAP;                             // [MAXP] x 8: the tile's pieces and their GMAP4 accumulators
  uint32_t* pen = pst + MAXP;
  int32_t* pacc = (int32_t*)(pen + MAXP);
  int32_t* pcnt = pacc + MAXP;
  uint32_t* psmin = (uint32_t*)(pcnt + MAXP);
  uint32_t* pemax = psmin + MAXP;
  uint32_t* psmax = pemax + MAXP;
  uint32_t* pemin = psmax + MAXP;
  int32_t* segmax = (int32_t*)(pemin + MAXP);              // [WC_SEGCAP] largest boundary depth of every stretch of IN words
  int32_t* slotd0 = segmax + WC_SEGCAP;                    // [WC_NCAP] depth before the resolved word
  uint32_t* nmask = (uint32_t*)(slotd0 + WC_NCAP);         // [WC_MASKW] bit w: word w is resolved position by position
  uint32_t* sbase = nmask + WC_MASKW;                      // [WC_MASKW] resolved words before the mask word
  uint16_t* segslot = (uint16_t*)(sbase + WC_MASKW);       // [WC_SEGCAP] resolved words before the stretch
  uint16_t* slotw = segslot + WC_SEGCAP;                   // [WC_NCAP + 2] word of every resolved slot (ascending)
  uint16_t* sopen = slotw + (WC_NCAP + 2);                 // [WC_NCAP + 2] run opens inside the slot
  uint16_t* sopre = sopen + (WC_NCAP + 2);                 // [WC_NCAP + 2] opens in the slots before
  __shared__ int32_t scratch[4][NWARPS];                   // one per scan of a tile (see block_excl_scan)
  __shared__ int32_t red[NWARPS][4];                       // per-warp {all words IN, largest boundary depth, largest upper bound}
  __shared__ uint32_t red2[NWARPS][5];                     // per-warp GMAP4 bounds and resolved maximum of the one-run tiles
  __shared__ int s_nh, s_nout;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t W = FULL ? 65536u : (1u << logw);
  const int nW = FULL ? MAXWORDS : (int)(W >> 5);
  const int WPT = FULL ? 8 : (nW + THREADS - 1) / THREADS; // words per thread (consecutive), <= 8
  const int w0 = tid * WPT;
  const uint32_t vmask = FULL ? 0xffu : (w0 < nW ? ((1u << min(WPT, nW - w0)) - 1u) : 0u);   // this thread's valid words
  const int32_t mn0 = cfg.mn0, mx0 = cfg.mx0;

  // this block's tile range inside [t_lo, t_hi): first tile whose region offset reaches the block's share of the regions
  int64_t t_begin, t_end;
  {
    const uint32_t o_lo = offS[t_lo], o_hi = offS[t_hi];
    int64_t bounds[2];
#pragma unroll 1
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
  // A run that crosses tile edges only needs its OVERALL maximum to be exact (the stitch kernel takes the maximum over its pieces).
  // acc_carry = what the earlier tiles of this block already proved about the run still open at the previous tile's end (0: none):
  // words whose depth bound stays below it need no resolution even when they hold this tile's own maximum.
  int32_t acc_carry = 0;

  // Invariants at the top of every tile iteration (restored by the tile that breaks them): cnt, npos, nmask, segmax and the piece
  // accumulators are clean, s_nout = 0 and s_nh = the number of regions carried in hin (0 when the previous tile must be scanned).
  for (int k = tid; k < MAXWORDS; k += THREADS) cnt[k] = 0;
  for (int k = tid; k < WC_NCAP * 16; k += THREADS) npos[k] = 0;
  for (int k = tid; k < WC_SEGCAP; k += THREADS) segmax[k] = 0;
  for (int k = tid; k < WC_MASKW; k += THREADS) nmask[k] = 0;
  for (int k = tid; k < MAXP; k += THREADS) { pacc[k] = 0; pcnt[k] = 0; psmin[k] = 0xffffffffu; pemax[k] = 0; psmax[k] = 0; pemin[k] = 0xffffffffu; pen[k] = 0; }
  if (tid == 0) { s_nh = 0; s_nout = 0; }
  __syncthreads();

  TileChrom tch;
  tch.lo = 1u; tch.next = 0u; tch.base = 0u;               // empty range: recomputed at the first tile and whenever a tile leaves it
  uint32_t s1_next = t_begin < t_end ? offS[t_begin] : 0u, s2_next = t_begin < t_end ? offS[t_begin + 1] : 0u;
#pragma unroll 1
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const uint32_t s0 = s1_next, s1 = s2_next;
    s1_next = s1;
    if (tile + 1 < t_end) s2_next = offS[tile + 2];        // the next tile's range, loaded a whole tile ahead
    const uint32_t nS_tile = s1 - s0;
#ifdef GMQL_TILE_TICKS
    if (cfg.dbg == 9 && tid == 0) tick_last = clock64();
#endif
    const bool scan_prev = nh_carry < 0;
    const uint32_t h0 = (scan_prev && tile > 0) ? offS[tile - 1] : s0;
    const uint32_t tile_lo = (uint32_t)tile << logw;
    if (nS_tile == 0 && (scan_prev ? h0 == s0 : nh_carry == 0)) { nh_carry = 0; acc_carry = 0; continue; }   // depth 0 throughout, nothing left open
    if (!XB && (tile_lo < tch.lo || tile_lo >= tch.next)) tch = tile_chrom(lm, tile_lo);                // a handful of times per chromosome
    if (tile + 1 < t_end) {                                // the next tile's records: into L2 while this tile is processed
      const uintptr_t a0 = (uintptr_t)(P + s1) & ~(uintptr_t)127, a1 = (uintptr_t)(P + s2_next);
      for (uintptr_t a = a0 + ((uintptr_t)tid << 7); a < a1; a += (uintptr_t)THREADS << 7) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
    // ---- phase A: counts of starts and stops' per word; the tile's own regions stay in registers
    if (scan_prev) {
#pragma unroll 1
      for (uint32_t i0 = h0; i0 < s0; i0 += THREADS) {
        const uint32_t i = i0 + tid;
        if (i < s0) {
          const uint32_t pr = P[i], rs = tile_lo - W + (pr >> 16), len = RLEN(pr);
          const uint32_t ee = rs + len + REXT(pr, rs, len);
          if (ee >= tile_lo) {
            const int slot = atomicAdd(&s_nh, 1);
            if (slot < HCAP) hin[slot] = pr;
            atomicAdd(&cnt[(ee - tile_lo) >> 5], 0x10000u);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int k = tid; k < nh_carry; k += THREADS) {
        const uint32_t pr = hin[k], rs = tile_lo - W + (pr >> 16), len = RLEN(pr);
        const uint32_t q = rs + len + REXT(pr, rs, len) - tile_lo;
        atomicAdd(&cnt[q >> 5], 0x10000u);
      }
    }
    uint32_t rcr[RC];                                      // records of the tile this thread keeps; bit r of rce: region r is extended
    uint32_t rce = 0;
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      const uint32_t i = s0 + tid + r * THREADS;
      rcr[r] = 0;
      if (i < s1) rcr[r] = P[i];
    }
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      if (s0 + tid + r * THREADS < s1) {
        const uint32_t q = rcr[r] >> 16, len = RLEN(rcr[r]);
        const uint32_t ex = REXT(rcr[r], tile_lo + q, len);
        rce |= ex << r;
        atomicAdd(&cnt[q >> 5], 1u);
        const uint32_t qe = q + len + ex;
        if (qe < W) atomicAdd(&cnt[qe >> 5], 0x10000u);
        else { const int slot = atomicAdd(&s_nout, 1); if (slot < HCAP) hout[slot] = rcr[r]; }
      }
    }
#pragma unroll 1
    for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
      const uint32_t pr = P[i], q = pr >> 16, len = RLEN(pr);
      const uint32_t qe = q + len + REXT(pr, tile_lo + q, len);
      atomicAdd(&cnt[q >> 5], 1u);
      if (qe < W) atomicAdd(&cnt[qe >> 5], 0x10000u);
      else { const int slot = atomicAdd(&s_nout, 1); if (slot < HCAP) hout[slot] = pr; }
    }
    __syncthreads();
    TICK(0);
    const int nh = s_nh, nout = s_nout;
    const int32_t depth_in = nh;                           // regions still open when the tile begins
    // ---- phase B: exact depth at every word boundary, classes of the words
    uint32_t cw[8];                                        // this thread's words ...
    uint32_t cprev = 0;                                    // ... and the word before them (its class decides whether a stretch starts here)
    if (FULL) {
      const uint4 a = ((const uint4*)cnt)[2 * tid], b = ((const uint4*)cnt)[2 * tid + 1];
      cw[0] = a.x; cw[1] = a.y; cw[2] = a.z; cw[3] = a.w; cw[4] = b.x; cw[5] = b.y; cw[6] = b.z; cw[7] = b.w;
      cprev = __shfl_up_sync(0xffffffffu, cw[7], 1);
      if (lane == 0) cprev = tid > 0 ? cnt[w0 - 1] : 0u;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) cw[j] = ((vmask >> j) & 1u) ? cnt[w0 + j] : 0u;
      if (w0 > 0 && w0 < nW) cprev = cnt[w0 - 1];
    }
    int32_t lsum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) lsum += (int32_t)(cw[j] & 0xffffu) - (int32_t)(cw[j] >> 16);
    int32_t tile_net;
    const int32_t d0 = depth_in + block_excl_scan(lsum, &tile_net, scratch[0]);
    // the counts are in registers: the table is clean again (every read of it happened before the scan's barrier)
    if (FULL) { ((uint4*)cnt)[2 * tid] = make_uint4(0u, 0u, 0u, 0u); ((uint4*)cnt)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u); }
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) if ((vmask >> j) & 1u) cnt[w0 + j] = 0u;
    }
    // failures (uniform across the block): this tile is skipped and left clean
    if (nh > HCAP || nout > HCAP || nS_tile > (uint32_t)MAX_TILE_REGIONS) {
      if (tid == 0) { *out.fail = nS_tile > (uint32_t)MAX_TILE_REGIONS ? 2 : 4; s_nh = 0; s_nout = 0; }
      nh_carry = -1; acc_carry = 0;
      __syncthreads();
      continue;
    }
    if (tid == 0) { s_nh = nout; s_nout = 0; }             // every thread has read both (before the scan's barrier)
    // classes: bit j of m_in / m_thr; a word is IN when its depth bounds lie inside [min, max], THR when they straddle a threshold.
    // bmax: largest boundary depth around this thread's IN words; hmax: largest upper bound d + starts of its IN words
    uint32_t m_in = 0, m_thr = 0;
    int32_t bmax = -1, hmax = -1;
    {
      int32_t d = d0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int32_t ns = (int32_t)(cw[j] & 0xffffu), ne = (int32_t)(cw[j] >> 16);
        const int32_t lo = d - ne, hi = d + ns, de = d + ns - ne;
        const bool ok = FULL || ((vmask >> j) & 1u);
        const bool in = ok && lo >= mn0 && hi <= mx0, outside = hi < mn0 || lo > mx0;
        if (in) { m_in |= 1u << j; bmax = max(bmax, max(d, de)); hmax = max(hmax, hi); }
        else if (ok && !outside) m_thr |= 1u << j;
        d = de;
      }
    }
    bool prev_in;
    {
      const int32_t pns = (int32_t)(cprev & 0xffffu), pne = (int32_t)(cprev >> 16);
      const int32_t pd0 = d0 - (pns - pne);
      prev_in = w0 > 0 && w0 < nW && (pd0 - pne >= mn0) && (pd0 + pns <= mx0);
    }
    // stretches of consecutive IN words: index, largest boundary depth
    const uint32_t m_start = m_in & ~((m_in << 1) | (prev_in ? 1u : 0u));
    const bool all_in = m_in == vmask;                     // (threads without words: vmask == 0, m_in == 0)
    // block-wide in ONE barrier: is every word of the tile IN, the largest boundary depth, the largest upper bound of a word
    {
      const bool w_all = __all_sync(0xffffffffu, all_in);
      const int32_t w_b = __reduce_max_sync(0xffffffffu, bmax), w_h = __reduce_max_sync(0xffffffffu, hmax);
      if (lane == 0) { red[warp][0] = w_all ? 1 : 0; red[warp][1] = w_b; red[warp][2] = w_h; }
    }
    __syncthreads();
    TICK(1);
    const bool blk_all = __all_sync(0xffffffffu, lane < NWARPS ? red[lane][0] != 0 : true);
    const int32_t blk_b = __reduce_max_sync(0xffffffffu, lane < NWARPS ? red[lane][1] : -1);
    const int32_t blk_h = __reduce_max_sync(0xffffffffu, lane < NWARPS ? red[lane][2] : -1);
    // ---- one-run tiles (three in four under dense coverage): every word is IN, so the whole tile is ONE piece [tile_lo, tile_lo + W)
    // of a run that began earlier; every region that starts here overlaps it (the open regions were accounted by the run's earlier
    // tiles, GMAP4.scala:44-45).  Only the run's maximum depth may still hide inside a word: M is what the word boundaries and the
    // run's earlier tiles prove; a word whose upper bound stays at or below M needs no look.  Half of these tiles resolve nothing
    // (four barriers, no second pass over the events); the others resolve a few words for their maximum only — no stretches, no
    // piece tables, no opens / closes.  (cover.dbg = 4, a test hook, sends them through the general path below instead.)
    const bool lite = blk_all && cfg.dbg != 4;
    const int32_t M = max(blk_b, acc_carry);
    int32_t seg_total = 0, seg_id = 0;                     // seg_id: stretch open before this thread's first word (-1: none)
    uint32_t m_noisy = 0;
    bool resolve = true;                                   // block-uniform
    if (lite) {
      resolve = blk_h > M;
      if (resolve) {
        int32_t d = d0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int32_t ns = (int32_t)(cw[j] & 0xffffu), ne = (int32_t)(cw[j] >> 16);
          if ((FULL || ((vmask >> j) & 1u)) && d + ns > M) m_noisy |= 1u << j;
          d += ns - ne;
        }
      }
    } else {
      if (!blk_all) {
        seg_id = block_excl_scan(__popc(m_start), &seg_total, scratch[1]) - 1;
        if (seg_total > WC_SEGCAP) seg_total = WC_SEGCAP;  // cannot happen: a stretch start needs a non-IN word before it
      } else { seg_total = 1; seg_id = tid == 0 ? -1 : 0; }  // every word of the tile is IN: one stretch, started by the tile's first word
      if (all_in) {
        if (vmask) atomicMax(&segmax[seg_id + (int32_t)(m_start & 1u)], bmax);     // one stretch covers all my words
      } else if (m_in) {
        int32_t s = seg_id, m = -1, d = d0;                // d: depth before word j
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const int32_t de = d + (int32_t)(cw[j] & 0xffffu) - (int32_t)(cw[j] >> 16);
          if ((m_in >> j) & 1u) {
            if ((m_start >> j) & 1u) { if (m >= 0) atomicMax(&segmax[s], m); ++s; m = -1; }
            m = max(m, max(d, de));
          } else if (m >= 0) { atomicMax(&segmax[s], m); m = -1; }
          d = de;
        }
        if (m >= 0) atomicMax(&segmax[s], m);
      }
      // the run entering the tile: its first stretch (index 0, when the tile's first word is IN) inherits what earlier tiles proved
      if (tid == 0 && acc_carry > 0 && (m_in & 1u)) atomicMax(&segmax[0], acc_carry);
      __syncthreads();
      TICK(2);
      // words to resolve: THR, and IN words whose upper bound exceeds what the stretch's boundaries already prove
      m_noisy = m_thr;
      if (m_in && !(all_in && hmax <= segmax[seg_id + (int32_t)(m_start & 1u)])) {
        int32_t s = seg_id, d = d0;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          if ((m_in >> j) & 1u) {
            if ((m_start >> j) & 1u) ++s;
            if (d + (int32_t)(cw[j] & 0xffffu) > segmax[s]) m_noisy |= 1u << j;
          }
          d += (int32_t)(cw[j] & 0xffffu) - (int32_t)(cw[j] >> 16);
        }
      }
    }
    int32_t nn_total = 0, slot0 = 0;
    if (resolve && (lite || __syncthreads_or(m_noisy ? 1 : 0))) slot0 = block_excl_scan(__popc(m_noisy), &nn_total, scratch[2]);
    TICK(3);
    if (nn_total > WC_NCAP) {                              // too many words to resolve: the rank kernel takes the operator
      if (tid == 0) { *out.fail = 5; s_nh = 0; }
      nh_carry = -1; acc_carry = 0;
      for (int k = tid; k < seg_total; k += THREADS) segmax[k] = 0;
      __syncthreads();
      continue;
    }
    if (cfg.dbg == 8 && tid == 0) { atomicAdd(out.fail + (lite ? 1 : 2), 1); atomicAdd(out.fail + 3, nn_total); }
    if (nn_total > 0) {
      if (m_noisy) atomicOr(&nmask[w0 >> 5], m_noisy << (w0 & 31));
      if (vmask && (w0 & 31) == 0) sbase[w0 >> 5] = (uint32_t)slot0;
      if (m_noisy | (lite ? 0u : m_start)) {
        int32_t s = seg_id, sl = slot0, d = d0;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          if (!lite && ((m_start >> j) & 1u)) { ++s; segslot[s] = (uint16_t)sl; }
          if ((m_noisy >> j) & 1u) { if (!lite) slotw[sl] = (uint16_t)(w0 + j); slotd0[sl] = d; ++sl; }
          d += (int32_t)(cw[j] & 0xffffu) - (int32_t)(cw[j] >> 16);
        }
      }
      __syncthreads();
      TICK(4);
      // ---- phase C: per-position net change of the words to resolve
      // (which of the 64 mask words hold anything at all: two registers, so that most events never touch shared memory)
      const unsigned nm_lo = __ballot_sync(0xffffffffu, nmask[lane] != 0u), nm_hi = __ballot_sync(0xffffffffu, nmask[32 + lane] != 0u);
      auto hit = [&](uint32_t q, int v) {                  // event at tile-local position q: counted when its word is resolved
        const uint32_t wd = q >> 5;
        if (!(((wd & 1024u) ? nm_hi : nm_lo) >> ((wd >> 5) & 31u) & 1u)) return;
        const uint32_t mk = nmask[wd >> 5];
        if ((mk >> (wd & 31u)) & 1u) {
          const uint32_t sl = sbase[wd >> 5] + (uint32_t)__popc(mk & ((1u << (wd & 31u)) - 1u));
          delta_add(npos, (int)(sl * 32u + (q & 31u)), v);
        }
      };
#pragma unroll 1
      for (int k = tid; k < nh; k += THREADS) {
        const uint32_t pr = hin[k], rs = tile_lo - W + (pr >> 16), len = RLEN(pr);
        hit(rs + len + REXT(pr, rs, len) - tile_lo, -1);
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        if (s0 + tid + r * THREADS < s1) {
          const uint32_t q = rcr[r] >> 16, qe = q + RLEN(rcr[r]) + ((XB ? (rcr[r] >> 15) : (rce >> r)) & 1u);
          hit(q, 1);
          if (qe < W) hit(qe, -1);
        }
      }
#pragma unroll 1
      for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
        const uint32_t pr = P[i], q = pr >> 16, len = RLEN(pr);
        hit(q, 1);
        const uint32_t qe = q + len + REXT(pr, tile_lo + q, len);
        if (qe < W) hit(qe, -1);
      }
      __syncthreads();
      TICK(5);
      if (tid < WC_MASKW) nmask[tid] = 0;                  // clean for the next tile (its next writer is several barriers away)
    }
    if (lite) {
      // the exact maximum inside every resolved word (one warp per word, lane = position)
      int32_t wmax = 0;
#pragma unroll 1
      for (int sl = warp; sl < nn_total; sl += NWARPS) {
        int32_t x = delta_get(npos, sl * 32 + lane);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < 16) npos[sl * 16 + lane] = 0;           // clean for the next tile (every lane's value went through the shuffles above)
        wmax = max(wmax, __reduce_max_sync(0xffffffffu, slotd0[sl] + x));
      }
      uint32_t a_smin = 0xffffffffu, a_emax = 0u, a_smax = 0u, a_emin = 0xffffffffu;
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        if (s0 + tid + r * THREADS < s1) {
          const uint32_t rs = tile_lo + (rcr[r] >> 16), re = rs + RLEN(rcr[r]);
          a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re);
        }
      }
#pragma unroll 1
      for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
        const uint32_t pr = P[i], rs = tile_lo + (pr >> 16), re = rs + RLEN(pr);
        a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re);
      }
      a_smin = __reduce_min_sync(0xffffffffu, a_smin); a_emax = __reduce_max_sync(0xffffffffu, a_emax);
      a_smax = __reduce_max_sync(0xffffffffu, a_smax); a_emin = __reduce_min_sync(0xffffffffu, a_emin);
      if (lane == 0) { red2[warp][0] = a_smin; red2[warp][1] = a_emax; red2[warp][2] = a_smax; red2[warp][3] = a_emin; red2[warp][4] = (uint32_t)wmax; }
      __syncthreads();
      const int32_t acc = max(M, (int32_t)__reduce_max_sync(0xffffffffu, lane < NWARPS ? red2[lane][4] : 0u));
      const bool fits = my_cnt + 1 <= out.seg;
      if (warp == 0) {
        a_smin = __reduce_min_sync(0xffffffffu, lane < NWARPS ? red2[lane][0] : 0xffffffffu);
        a_emax = __reduce_max_sync(0xffffffffu, lane < NWARPS ? red2[lane][1] : 0u);
        a_smax = __reduce_max_sync(0xffffffffu, lane < NWARPS ? red2[lane][2] : 0u);
        a_emin = __reduce_min_sync(0xffffffffu, lane < NWARPS ? red2[lane][3] : 0xffffffffu);
        if (lane == 0) {
          if (!fits) *out.fail = 1;
          else {
            const int64_t gb = (int64_t)blockIdx.x * out.seg + my_cnt;
            out.pstart[gb] = tile_lo; out.pend[gb] = tile_lo + W; out.acc[gb] = acc; out.cnt[gb] = (int32_t)nS_tile;
            out.smin[gb] = a_smin; out.emax[gb] = a_emax; out.smax[gb] = a_smax; out.emin[gb] = a_emin;
            out.flags[gb] = (uint8_t)3;
          }
        }
      }
      TICK(11);
      if (fits) my_cnt += 1;
      acc_carry = acc;
      { uint32_t* t = hin; hin = hout; hout = t; }
      nh_carry = nout;
      continue;
    }
    const bool cont0 = depth_in >= mn0 && depth_in <= mx0;
    int32_t open_total = 0;
    if (nn_total > 0) {
      // ---- phase D1: run opens inside every resolved word (one warp per word, lane = position)
#pragma unroll 1
      for (int sl = warp; sl < nn_total; sl += NWARPS) {
        int32_t x = delta_get(npos, sl * 32 + lane);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        const int32_t dd = slotd0[sl];
        const int32_t d = dd + x;
        const unsigned inm = __ballot_sync(0xffffffffu, d >= mn0 && d <= mx0);
        const unsigned prevm = (inm << 1) | ((dd >= mn0 && dd <= mx0) ? 1u : 0u);
        if (lane == 0) sopen[sl] = (uint16_t)__popc(inm & ~prevm);
      }
      __syncthreads();
      TICK(6);
      if (nn_total <= 32) {
        // few resolved words (the usual case): every warp sums the open counts itself, no block-wide scan
        const int32_t v = lane < nn_total ? (int32_t)sopen[lane] : 0;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        open_total = __shfl_sync(0xffffffffu, x, 31);
        if (warp == 0) { sopre[lane] = (uint16_t)(x - v); if (lane == 31) sopre[32] = (uint16_t)x; }
      } else {
        // two slots per thread: WC_NCAP <= 2 * THREADS
        const int a = 2 * tid, b = 2 * tid + 1;
        const int32_t va = a < nn_total ? (int32_t)sopen[a] : 0, vb = b < nn_total ? (int32_t)sopen[b] : 0;
        const int32_t pre = block_excl_scan(va + vb, &open_total, scratch[3]);
        if (a <= nn_total) sopre[a] = (uint16_t)pre;
        if (b <= nn_total) sopre[b] = (uint16_t)(pre + va);
      }
    }
    const int np = (cont0 ? 1 : 0) + open_total;
    if (np > MAXP) {
      if (tid == 0) { *out.fail = 3; s_nh = 0; }
      nh_carry = -1; acc_carry = 0;
      for (int k = tid; k < seg_total; k += THREADS) segmax[k] = 0;
      for (int k = tid; k < nn_total * 16; k += THREADS) npos[k] = 0;
      __syncthreads();
      continue;
    }
    if (nn_total > 0) __syncthreads();                     // sopre is complete
    TICK(7);
    // ---- phase D2: pieces.  Resolved words: exact opens, closes and maxima; stretches of IN words: their boundary maximum
    const int32_t d_end = depth_in + tile_net;
    const bool reach_end = np > 0 && d_end >= mn0 && d_end <= mx0;
    if (tid == 0) {
      if (cont0) { pst[0] = tile_lo; atomicMax(&pacc[0], depth_in); }
      if (reach_end) pen[np - 1] = tile_lo + W;
    }
#pragma unroll 1
    for (int sl = warp; sl < nn_total; sl += NWARPS) {
      int32_t x = delta_get(npos, sl * 32 + lane);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (lane < 16) npos[sl * 16 + lane] = 0;             // clean for the next tile (every lane's value went through the shuffles above)
      const int32_t dd = slotd0[sl];
      const int32_t d = dd + x;
      const bool in = d >= mn0 && d <= mx0;
      const unsigned inm = __ballot_sync(0xffffffffu, in);
      const unsigned prevm = (inm << 1) | ((dd >= mn0 && dd <= mx0) ? 1u : 0u);
      const unsigned openm = inm & ~prevm, closem = ~inm & prevm;
      const uint32_t pos = tile_lo + ((uint32_t)slotw[sl] << 5) + (uint32_t)lane;
      const int base = (cont0 ? 1 : 0) + (int)sopre[sl] - 1;                     // piece open when the word begins (if any)
      const int mine = base + __popc(openm & (0xffffffffu >> (31 - lane)));      // piece this position belongs to when in range
      if ((openm >> lane) & 1u) pst[mine] = pos;
      if ((closem >> lane) & 1u) pen[base + __popc(openm & ((1u << lane) - 1u))] = pos;
      unsigned todo = inm;                                 // maximum depth of every in-range stretch of the word
      while (todo) {
        const int first = __ffs(todo) - 1;
        const int pc = __shfl_sync(0xffffffffu, mine, first);
        const bool sel = in && mine == pc;
        const int32_t mx = __reduce_max_sync(0xffffffffu, sel ? d : 0);
        if (lane == first) atomicMax(&pacc[pc], mx);
        todo &= ~__ballot_sync(0xffffffffu, sel);
      }
    }
    if (m_start) {
      int32_t s = seg_id;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        if ((m_start >> j) & 1u) {
          ++s;
          const int pc = (cont0 ? 1 : 0) + (nn_total > 0 ? (int)sopre[segslot[s]] : 0) - 1;   // the run this stretch lies in
          if (pc >= 0) atomicMax(&pacc[pc], segmax[s]);
        }
      }
    }
    __syncthreads();
    TICK(8);
    for (int k = tid; k < seg_total; k += THREADS) segmax[k] = 0;                 // no reader left
    acc_carry = reach_end ? pacc[np - 1] : 0;              // block-uniform: read after the barrier, before the tables are cleaned
    // ---- phase E: GMAP4, region-driven: the regions open at the tile start, then the tile's own
    if (np == 1 || np == 2) {
      // one or two pieces (the usual cases under dense coverage): no search, thread-local accumulation, one reduction per warp and piece
#pragma unroll 1
      for (int pi = 0; pi < np; ++pi) {
        const uint32_t p0 = pst[pi], p1 = pen[pi];
        const bool from_prev = pi == 0 && cont0;           // the open regions only feed a run that STARTS in this tile
        int32_t c = 0;
        uint32_t a_smin = 0xffffffffu, a_emax = 0, a_smax = 0, a_emin = 0xffffffffu;
        if (!from_prev) {
#pragma unroll 1
          for (int k = tid; k < nh; k += THREADS) {
            const uint32_t rs = tile_lo - W + (hin[k] >> 16), re = rs + RLEN(hin[k]);
            if (re > tile_lo && p1 > rs && p0 < re) { ++c; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
          }
        }
#pragma unroll
        for (int r = 0; r < RC; ++r) {
          if (s0 + tid + r * THREADS < s1) {
            const uint32_t rs = tile_lo + (rcr[r] >> 16), re = rs + RLEN(rcr[r]);
            if (p1 > rs && p0 < re) { ++c; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
          }
        }
#pragma unroll 1
        for (uint32_t i = s0 + RC * THREADS + tid; i < s1; i += THREADS) {
          const uint32_t pr = P[i], rs = tile_lo + (pr >> 16), re = rs + RLEN(pr);
          if (p1 > rs && p0 < re) { ++c; a_smin = min(a_smin, rs); a_emax = max(a_emax, re); a_smax = max(a_smax, rs); a_emin = min(a_emin, re); }
        }
        c = __reduce_add_sync(0xffffffffu, c);
        a_smin = __reduce_min_sync(0xffffffffu, a_smin); a_emax = __reduce_max_sync(0xffffffffu, a_emax);
        a_smax = __reduce_max_sync(0xffffffffu, a_smax); a_emin = __reduce_min_sync(0xffffffffu, a_emin);
        if (lane == 0 && c > 0) {
          atomicAdd(&pcnt[pi], c);
          atomicMin(&psmin[pi], a_smin); atomicMax(&pemax[pi], a_emax);
          atomicMax(&psmax[pi], a_smax); atomicMin(&pemin[pi], a_emin);
        }
      }
    } else if (np > 2) {
#pragma unroll 1
      for (int k0h = 0; k0h < nh; k0h += THREADS) {
        const int k = k0h + tid;
        bool valid = k < nh;
        uint32_t rs = 0, re = 0;
        if (valid) { rs = tile_lo - W + (hin[k] >> 16); re = rs + RLEN(hin[k]); valid = re > tile_lo; }
        gmap4_round(valid, rs, re, false, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
      }
#pragma unroll 1
      for (int r = 0; r < RC; ++r) {
        if ((uint32_t)(r * THREADS) < nS_tile) {            // block-uniform
          const bool valid = s0 + tid + r * THREADS < s1;
          const uint32_t pr = valid ? P[s0 + tid + r * THREADS] : 0u;
          const uint32_t rs = tile_lo + (pr >> 16);
          gmap4_round(valid, rs, rs + RLEN(pr), true, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
        }
      }
#pragma unroll 1
      for (uint32_t i0 = s0 + RC * THREADS; i0 < s1; i0 += THREADS) {
        const uint32_t i = i0 + tid;
        const bool valid = i < s1;
        uint32_t rs = 0, re = 0;
        if (valid) { const uint32_t pr = P[i]; rs = tile_lo + (pr >> 16); re = rs + RLEN(pr); }
        gmap4_round(valid, rs, re, true, cont0, np, pst, pen, pcnt, psmin, pemax, psmax, pemin);
      }
    }
    TICK(9);
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
    } else __syncthreads();                                // (the clean-up above and the reads of hin are over before the next tile writes)
    TICK(10);
    // the regions still open at the tile end are the next tile's open regions
    { uint32_t* t = hin; hin = hout; hout = t; }
    nh_carry = nout;
  }
  if (tid == 0) out.blk_cnt[blockIdx.x] = (uint32_t)my_cnt;
#undef TICK
#undef RLEN
#undef REXT
}
