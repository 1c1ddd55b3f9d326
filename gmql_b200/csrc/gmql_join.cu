// gmql_join.cu — genometric JOIN on the device.
//
// Replaces GenometricJoin.apply (GMQL-Spark/.../RegionsOperators/GenometricJoin.scala:21-212) for one JoinQuadruple:
// atom split at MinDistance (:65-72), maxDistance rule (:75-79), the bin-quantised candidate window
// (computeBinStartRef/StopRef :284-309, binRightDs2 :334-342, first-bin de-duplication :120-122), first-round
// predicates (checkRegionCondition :246-282), MD(k) with ties per (left record, right sample) (:130-151),
// second-round predicates (:160-188), INTERSECTION overlap filter (:190-198), region builders (:345-372) and
// DISTINCT (:205-218).
//
// Design (not a port of the bin-replicated RDD join): the right side becomes an interval index (sorted linear
// starts + running max stop); every left region turns its bin range [binStart, binEnd] into one coordinate window
//     r.start < (binEnd+1)*B   and   r.stop >= binStart*B
// (exactly the set of right regions sharing a bin with it), and a warp scans that window of the index.
// Without MD the index pools all right samples and the sample pairing is a filter; with MD(k) the index is keyed by
// right sample so that each lane owns one (left region, right sample) group and selects its k-th distance with ties.
// All variable-size output uses count -> scan -> write.
#include "index.cuh"

namespace {

struct JoinCfg {
  int f_has_max, f_has_min, f_stream;   // first round; stream: 0 none, 1 upstream('+'), 2 downstream('-')
  int64_t f_max, f_min;
  int has_md;
  int64_t k;
  int s_has_max, s_has_min, s_stream;   // second round
  int64_t s_max, s_min;
  int64_t maxD, B;
  int builder;
  int has_eq;
};

__device__ __forceinline__ int64_t join_distance(int64_t as, int64_t ae, int64_t bs, int64_t be) {
  // distanceCalculator (:375-386)
  int64_t d1 = as - be, d2 = bs - ae;
  int64_t a1 = d1 < 0 ? -d1 : d1, a2 = d2 < 0 ? -d2 : d2;
  int64_t m = a1 < a2 ? a1 : a2;
  return (ae < bs || be < as) ? m : -m;
}

__device__ __forceinline__ bool stream_ok(int stream, int lstrand, int64_t ls, int64_t le, int64_t rs, int64_t re) {
  // :255-279 / :171-183 — uses the LEFT strand only; '*' behaves as '+'
  if (stream == 1) return (lstrand != 2) ? (re <= ls) : (rs >= le);
  if (stream == 2) return (lstrand != 2) ? (rs >= le) : (re <= ls);
  return true;
}

__device__ __forceinline__ void left_bins(const JoinCfg& c, int lstrand, int64_t ls, int64_t le, int64_t* b0, int64_t* b1) {
  // computeBinStartRef (:284-296), computeBinStopRef (:298-309)
  int64_t v0 = (c.f_stream == 0 || c.f_stream == lstrand || (lstrand == 0 && c.f_stream == 1)) ? ls - c.maxD : le;
  if (v0 < 0) v0 = 0;
  int64_t v1 = (c.f_stream == 0 || c.f_stream != lstrand || (lstrand == 0 && c.f_stream == 2)) ? le + c.maxD : ls;
  *b0 = v0 / c.B;
  *b1 = v1 / c.B;
}

__device__ __forceinline__ bool first_round_ok(const JoinCfg& c, int lstrand, int64_t ls, int64_t le, int rstrand, int64_t rs, int64_t re, int64_t b0, int64_t* dist) {
  if (re < b0 * c.B) return false;                       // right bins must reach the left's first bin: re / B < b0 (re, b0 >= 0)
  if (!strand_ok(lstrand, rstrand)) return false;        // :123
  int64_t d = join_distance(ls, le, rs, re);
  *dist = d;
  if (c.f_has_max && !(c.f_max > d)) return false;       // :250
  if (c.f_has_min && !(c.f_min < d)) return false;       // :251
  return stream_ok(c.f_stream, lstrand, ls, le, rs, re);
}

__device__ __forceinline__ bool second_round_ok(const JoinCfg& c, int lstrand, int64_t ls, int64_t le, int64_t rs, int64_t re, int64_t d) {
  if (c.s_has_max && !(c.s_max > d)) return false;       // :169
  if (c.s_has_min && !(c.s_min < d)) return false;       // :170
  if (!stream_ok(c.s_stream, lstrand, ls, le, rs, re)) return false;
  if (c.builder == GMQL_B_INTERSECTION && !(ls < re && rs < le)) return false;   // :190-196
  return true;
}

struct LeftSide {
  const int32_t* chrom; const void* start; const void* stop; int bits; const int8_t* strand; const int32_t* sample; const int64_t* eq; int64_t n;
  const uint32_t* perm;   // optional: left rows in genome order (slot w handles row perm[w]); counts / offsets are per slot
};
template <typename K>
struct RightIndex {
  const K* start; const K* stop; const K* pmax; const uint32_t* row; int64_t n;
  const uint32_t* tbl; int tbl_shift; int64_t n_buckets;   // tbl[b] = lower_bound(start, b << tbl_shift), b = 0..n_buckets
  const int8_t* strand;    // per sorted entry
  const int32_t* sample;   // per sorted entry
  const int64_t* eq;       // per ORIGINAL row (or null)
  const int4* packed;      // per sorted entry {local start, local stop, original row, strand} when the right coordinates are 32-bit, else null
};

// lower_bound over the sorted starts through the bucket table: two table reads and a search inside one bucket
// (a few entries) instead of ~log2(n) dependent loads across the whole index
template <typename K>
__device__ __forceinline__ int64_t start_lower_bound(const RightIndex<K>& R, uint64_t x) {
  uint64_t b = x >> R.tbl_shift;
  if (b >= (uint64_t)R.n_buckets) return R.n;            // beyond the last bucket boundary: every start is smaller
  int64_t lo = R.tbl[b], hi = R.tbl[b + 1];
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (R.start[mid] < (K)x) lo = mid + 1; else hi = mid; }
  return lo;
}

// window of the index a left region must scan: [lo, hi)
template <typename K>
__device__ __forceinline__ void window(const RightIndex<K>& R, const LinMap& lm, uint64_t block_base, int c, int64_t b0, int64_t b1, int64_t B, int64_t* lo, int64_t* hi) {
  int64_t span = lm.span[c];
  int64_t wend = (b1 + 1) * B;                            // r.start / B <= binEnd
  if (wend > span + 1) wend = span + 1;
  int64_t wbeg = b0 * B;                                  // r.stop / B >= binStart
  if (wbeg > span + 1) wbeg = span + 1;
  uint64_t base = block_base + lm.coff[c];
  const int64_t h = start_lower_bound<K>(R, base + (uint64_t)wend);
  // first entry whose running max stop reaches wbeg: entries starting at or after wbeg qualify (stop >= start), so the
  // answer is at or before lower_bound(start, wbeg); gallop back over the non-decreasing pmax from there
  const K y = (K)(base + (uint64_t)wbeg);
  int64_t r = start_lower_bound<K>(R, base + (uint64_t)wbeg);
  if (r > h) r = h;
  int64_t step = 1, l = r;
  while (l > 0 && R.pmax[l - 1] >= y) { r = l; l = l > step ? l - step : 0; step <<= 1; }
  // invariant: pmax[r] >= y (or r == h) and (l == 0 or pmax[l-1] < y): the first qualifying entry lies in [l, r]
  while (l < r) { int64_t mid = (l + r) >> 1; if (R.pmax[mid] < y) l = mid + 1; else r = mid; }
  *hi = h;
  *lo = l;
}

template <typename K>
__global__ void k_start_table(const K* __restrict__ start, int64_t n, int shift, int64_t n_buckets, uint32_t* __restrict__ tbl) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)b << shift;
    K kx = (x > (uint64_t)(~(K)0)) ? ~(K)0 : (K)x;
    tbl[b] = (uint32_t)lower_bound_dev<K>(start, n, kx);
  }
}

// ---- no MD: warp per left region over the pooled right index -------------------------------------------------------------
template <typename K, bool WRITE>
__global__ void __launch_bounds__(256) k_join_scan(LeftSide L, RightIndex<K> R, LinMap lm, JoinCfg cfg, const int32_t* __restrict__ pair_idx, int n_rs,
                                                   int64_t* __restrict__ cnt, const int64_t* __restrict__ off, int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int32_t* __restrict__ out_p) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t slot = warp; slot < L.n; slot += nwarps) {
    const int64_t i = L.perm ? (int64_t)L.perm[slot] : slot;
    int c = L.chrom[i];
    int64_t ls = ld_coord(L.start, L.bits, i), le = ld_coord(L.stop, L.bits, i);
    int lst = L.strand ? (int)L.strand[i] : 0;
    int la = L.sample ? L.sample[i] : 0;
    int64_t b0, b1;
    left_bins(cfg, lst, ls, le, &b0, &b1);
    int64_t total = 0;
    int64_t wbase = WRITE ? off[slot] : 0;
    if (b0 <= b1 && c >= 0 && c < lm.n_chrom) {
      int64_t lo, hi;
      window<K>(R, lm, 0ull, c, b0, b1, cfg.B, &lo, &hi);
      uint64_t cbase = lm.coff[c];
      int64_t leq = cfg.has_eq ? L.eq[i] : 0;
      for (int64_t k0 = lo; k0 < hi; k0 += 32) {
        int64_t k = k0 + lane;
        bool ok = false;
        int p = -1;
        uint32_t rrow = 0;
        if (k < hi && R.packed) {                          // 32-bit right coordinates: one 16-byte record per candidate
          const int4 en = R.packed[k];
          int64_t d;
          if (first_round_ok(cfg, lst, ls, le, en.w & 3, (int64_t)en.x, (int64_t)en.y, b0, &d)) {
            p = pair_idx[(int64_t)la * n_rs + (en.w >> 2)];                 // :124 (the bin join has no sample in its key)
            rrow = (uint32_t)en.z;
            ok = p >= 0 && second_round_ok(cfg, lst, ls, le, (int64_t)en.x, (int64_t)en.y, d);
            if (ok && cfg.has_eq) ok = R.eq[rrow] == leq;                   // :102-114
          }
        } else if (k < hi) {
          int64_t rs = (int64_t)((uint64_t)R.start[k] - cbase), re = (int64_t)((uint64_t)R.stop[k] - cbase);
          int64_t d;
          if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d)) {
            p = pair_idx[(int64_t)la * n_rs + R.sample[k]];                 // :124 (the bin join has no sample in its key)
            rrow = R.row[k];
            ok = p >= 0 && second_round_ok(cfg, lst, ls, le, rs, re, d);
            if (ok && cfg.has_eq) ok = R.eq[rrow] == leq;                   // :102-114
          }
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (WRITE && ok) {
          int64_t o = wbase + total + __popc(m & ((1u << lane) - 1u));
          out_l[o] = (int32_t)i; out_r[o] = (int32_t)rrow; out_p[o] = p;
        }
        total += __popc(m);
      }
    }
    if (!WRITE && lane == 0) cnt[slot] = total;
  }
}

constexpr int MD_WCAP = 8;   // candidates of one (left region, right sample) group kept in registers

// ---- MD(k): warp per left region, one lane per allowed right sample, index keyed by sample -----------------------------
template <typename K, bool WRITE>
__global__ void __launch_bounds__(256) k_join_md(LeftSide L, RightIndex<K> R, LinMap lm, JoinCfg cfg, const int64_t* __restrict__ allow_off, const int32_t* __restrict__ allow_rs,
                                                 const int32_t* __restrict__ allow_pair, const int32_t* __restrict__ mult,
                                                 int64_t* __restrict__ cnt, const int64_t* __restrict__ off, int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int32_t* __restrict__ out_p,
                                                 uint32_t* __restrict__ memo_lo, uint16_t* __restrict__ memo_ok, int64_t memo_stride) {
  // memo (optional): the count pass records, per (slot, allowed sample) group handled in registers, the window start and the
  // mask of emitted candidates; the write pass then only reads the rows back instead of repeating search and selection
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t slot = warp; slot < L.n; slot += nwarps) {
    const int64_t i = L.perm ? (int64_t)L.perm[slot] : slot;
    int c = L.chrom[i];
    int64_t ls = ld_coord(L.start, L.bits, i), le = ld_coord(L.stop, L.bits, i);
    int lst = L.strand ? (int)L.strand[i] : 0;
    int la = L.sample ? L.sample[i] : 0;
    int64_t b0, b1;
    left_bins(cfg, lst, ls, le, &b0, &b1);
    int64_t total = 0;
    int64_t wbase = WRITE ? off[slot] : 0;
    if (b0 <= b1 && c >= 0 && c < lm.n_chrom) {
      int64_t a0 = allow_off[la], a1 = allow_off[la + 1];
      int64_t m_dup = mult ? mult[i] : 1;
      int64_t leq = cfg.has_eq ? L.eq[i] : 0;
      uint64_t cbase = lm.coff[c];
      for (int64_t q0 = a0; q0 < a1; q0 += 32) {
        int64_t q = q0 + lane;
        int64_t mine = 0;
        int64_t lo = 0, hi = 0, thr = 0;
        int p = -1;
        uint64_t bbase = 0;
        int64_t dreg[MD_WCAP];
        int rreg[MD_WCAP];
        unsigned ok2 = 0;
        int nreg = 0;
        bool in_regs = false;
        bool use_memo = false;
        if (WRITE && memo_ok && q < a1) {
          const uint16_t m = memo_ok[slot * memo_stride + (q - a0)];
          if (!(m & 0x8000u)) {
            use_memo = true; in_regs = true; nreg = MD_WCAP;
            ok2 = m & 0xffu; mine = __popc(ok2);
            lo = memo_lo[slot * memo_stride + (q - a0)];
            p = allow_pair[q];
          }
        }
        if (q < a1 && !use_memo) {
          int b = allow_rs[q];
          p = allow_pair[q];
          bbase = (uint64_t)b * lm.G;
          window<K>(R, lm, bbase, c, b0, b1, cfg.B, &lo, &hi);
          if (R.packed && hi - lo <= MD_WCAP) {
            // the usual case: a handful of candidates.  Load each ONCE (one 16-byte record), keep distance and flags in
            // registers, select the k-th distance there
            nreg = (int)(hi - lo);
            int len = 0;
#pragma unroll
            for (int j = 0; j < MD_WCAP; ++j) {
              dreg[j] = 0x7fffffffffffffffLL;
              if (j < nreg) {
                const int4 en = R.packed[lo + j];
                int64_t d;
                if (first_round_ok(cfg, lst, ls, le, en.w, (int64_t)en.x, (int64_t)en.y, b0, &d) && (!cfg.has_eq || R.eq[(uint32_t)en.z] == leq)) {
                  dreg[j] = d; ++len;
                  if (second_round_ok(cfg, lst, ls, le, (int64_t)en.x, (int64_t)en.y, d)) ok2 |= 1u << j;
                }
                rreg[j] = en.z;
              }
            }
            if (len > 0) {
              int64_t want = m_dup * len < cfg.k ? m_dup * len : cfg.k;   // count = min(len, k), identical left records pool their lists (:134)
              int64_t idx = (want - 1) / m_dup;                           // position in the de-duplicated sorted list
              // thr = the distance at sorted position idx = the smallest d_j with #{i : d_i <= d_j} > idx
              thr = 0x7fffffffffffffffLL;
              if (idx == 0) {                                             // MD(1): the minimum
#pragma unroll
                for (int j = 0; j < MD_WCAP; ++j) thr = dreg[j] < thr ? dreg[j] : thr;
              } else
#pragma unroll
              for (int j = 0; j < MD_WCAP; ++j) {
                if (dreg[j] == 0x7fffffffffffffffLL) continue;
                int le_cnt = 0;
#pragma unroll
                for (int i2 = 0; i2 < MD_WCAP; ++i2) le_cnt += (dreg[i2] <= dreg[j]) ? 1 : 0;
                if (le_cnt > idx && dreg[j] < thr) thr = dreg[j];
              }
#pragma unroll
              for (int j = 0; j < MD_WCAP; ++j) if (dreg[j] <= thr && ((ok2 >> j) & 1u)) ++mine; else ok2 &= ~(1u << j);
            } else ok2 = 0;
            in_regs = true;
            if (!WRITE && memo_ok) { memo_ok[slot * memo_stride + (q - a0)] = (uint16_t)ok2; memo_lo[slot * memo_stride + (q - a0)] = (uint32_t)lo; }
          } else {
          if (!WRITE && memo_ok) memo_ok[slot * memo_stride + (q - a0)] = 0x8000u;   // handled by the streaming path in both passes
          // first-round survivors of this (left region, right sample) group and their k-th smallest distance (:136-146)
          int64_t len = 0;
          for (int64_t k = lo; k < hi; ++k) {
            int64_t rs = (int64_t)((uint64_t)R.start[k] - bbase - cbase), re = (int64_t)((uint64_t)R.stop[k] - bbase - cbase);
            int64_t d;
            if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq)) ++len;
          }
          if (len > 0) {
            // identical left records pool their candidate lists (MapKey, :134): the list is repeated m_dup times
            int64_t want = m_dup * len < cfg.k ? m_dup * len : cfg.k;   // count = min(len, k)
            int64_t idx = (want - 1) / m_dup;                           // position in the de-duplicated sorted list
            // selection without storage: walk the distinct distances in increasing order
            int64_t seen = 0;
            int64_t prev = 0;
            bool first = true;
            while (true) {
              int64_t cur = 0, ncur = 0;
              bool any = false;
              for (int64_t k = lo; k < hi; ++k) {
                int64_t rs = (int64_t)((uint64_t)R.start[k] - bbase - cbase), re = (int64_t)((uint64_t)R.stop[k] - bbase - cbase);
                int64_t d;
                if (!(first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq))) continue;
                if (!first && d <= prev) continue;
                if (!any || d < cur) { cur = d; ncur = 1; any = true; }
                else if (d == cur) ++ncur;
              }
              seen += ncur;
              prev = cur;
              first = false;
              if (seen > idx) break;
            }
            thr = prev;
            for (int64_t k = lo; k < hi; ++k) {
              int64_t rs = (int64_t)((uint64_t)R.start[k] - bbase - cbase), re = (int64_t)((uint64_t)R.stop[k] - bbase - cbase);
              int64_t d;
              if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq) &&
                  d <= thr && second_round_ok(cfg, lst, ls, le, rs, re, d)) ++mine;
            }
          }
        }
          }
        // exclusive scan of `mine` across the warp
        int64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        int64_t chunk_total = __shfl_sync(0xffffffffu, incl, 31);
        if (WRITE && mine > 0 && in_regs) {
          int64_t o = wbase + total + incl - mine;
#pragma unroll
          for (int j = 0; j < MD_WCAP; ++j) {
            if (j < nreg && ((ok2 >> j) & 1u)) { out_l[o] = (int32_t)i; out_r[o] = use_memo ? R.packed[lo + j].z : rreg[j]; out_p[o] = p; ++o; }
          }
        } else if (WRITE && mine > 0) {
          int64_t o = wbase + total + incl - mine;
          for (int64_t k = lo; k < hi; ++k) {
            int64_t rs = (int64_t)((uint64_t)R.start[k] - bbase - cbase), re = (int64_t)((uint64_t)R.stop[k] - bbase - cbase);
            int64_t d;
            if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq) &&
                d <= thr && second_round_ok(cfg, lst, ls, le, rs, re, d)) {
              out_l[o] = (int32_t)i; out_r[o] = (int32_t)R.row[k]; out_p[o] = p;
              ++o;
            }
          }
        }
        total += chunk_total;
      }
    }
    if (!WRITE && lane == 0) cnt[slot] = total;
  }
}

// ---- MD(k), team form: a warp owns 32 CONSECUTIVE left regions of the genome order and visits the right samples one by one ----
// k_join_md gives every (left region, right sample) group to one lane: each lane searches a different sample's index block and
// loads its own handful of candidates — ~15 dependent, uncoalesced sector reads per group, and the kernel is bound by the L2's
// sector rate (C4: 2e8 groups, 16 + 10 ms).  Neighbouring left regions of the genome order look at (nearly) the same candidates
// of a sample.  Here a round takes 32 right samples: lane j searches sample j ONCE for the whole team (the union of the team's
// bin ranges) and copies its few candidates (16-byte packed records) into the sample's slot in shared memory; then the team
// walks the 32 slots and every lane selects from a slot with its own window, first-round predicates and k-th distance
// (:117-151) in 32-bit arithmetic.  Same results group by group; the rows of a left region come out by right sample, then in
// index order.  Samples with more candidates near the team than a slot holds are staged one at a time (up to MDT_CAP), and
// beyond that every lane streams its own window as k_join_md does.
constexpr int MDT_SLOT = 16;  // candidates per (team, right sample) slot (a 16-bit memo mask per lane)
constexpr int MDT_CAP = 32;   // candidates of one (team, right sample) staged one sample at a time (a 32-bit survivor mask)
constexpr int MDT_WB = 8;     // entries the write pass keeps in flight
constexpr uint32_t MDT_NONE = 0xffffffffu, MDT_REDO = 0xfffffffeu;   // memo: no candidate near the team / not memoised

// the distance thresholds of both rounds as closed intervals lo <= d <= hi on 32-bit distances (d lies in [-(2^31-1), 2^31-1]);
// never_f / never_s: a round no distance can pass
struct JoinCfg32 { int f_lo, f_hi, s_lo, s_hi, never_f, never_s, f_stream, s_stream, builder, has_eq; };
static JoinCfg32 narrow_cfg(const JoinCfg& c) {
  auto fit = [](int has_max, int64_t mx, int has_min, int64_t mn, int* lo, int* hi, int* never) {
    *lo = (int)-0x7fffffff; *hi = 0x7fffffff; *never = 0;
    if (has_max) {                                         // max > d  <=>  d <= max - 1
      if (mx <= -0x7fffffffLL) *never = 1; else if (mx <= 0x7fffffffLL) *hi = (int)(mx - 1);
    }
    if (has_min) {                                         // min < d  <=>  d >= min + 1
      if (mn >= 0x7fffffffLL) *never = 1; else if (mn >= -0x7fffffffLL) *lo = (int)(mn + 1);
    }
  };
  JoinCfg32 r;
  fit(c.f_has_max, c.f_max, c.f_has_min, c.f_min, &r.f_lo, &r.f_hi, &r.never_f);
  fit(c.s_has_max, c.s_max, c.s_has_min, c.s_min, &r.s_lo, &r.s_hi, &r.never_s);
  r.f_stream = c.f_stream; r.s_stream = c.s_stream; r.builder = c.builder; r.has_eq = c.has_eq;
  return r;
}
__device__ __forceinline__ int join_distance32(int as, int ae, int bs, int be) {   // distanceCalculator (:375-386)
  const int d1 = as - be, d2 = bs - ae;
  const int a1 = d1 < 0 ? -d1 : d1, a2 = d2 < 0 ? -d2 : d2;
  const int m = a1 < a2 ? a1 : a2;
  return (ae < bs || be < as) ? m : -m;
}

// what a lane keeps about its left region for the selection: window, and the stream conditions (:255-279, :171-183; they look at
// the LEFT strand only) of both rounds as bounds on the candidate's stop / start
struct LaneQ { int ls, le, lst, wbeg, wlast, f_re_max, f_rs_min, s_re_max, s_rs_min; };
__device__ __forceinline__ void stream_bounds(int stream, int lst, int ls, int le, int* re_max, int* rs_min) {
  const bool need_re = (stream == 1 && lst != 2) || (stream == 2 && lst == 2);   // re <= ls
  const bool need_rs = (stream == 1 && lst == 2) || (stream == 2 && lst != 2);   // rs >= le
  *re_max = need_re ? ls : 0x7fffffff;
  *rs_min = need_rs ? le : (int)0x80000000;
}

// one lane's (left region, right sample) group over `n` <= 32 staged candidates (cand[t * stride], sorted by start):
// count = min(len, k) with ties (:136-146), second round (:160-198); returns the mask of the candidates that become rows.
// The scan is branch-free: the lanes of a team disagree about nearly every candidate.
__device__ __forceinline__ uint32_t md_select32(const JoinCfg32& c, const int4* cand, int stride, int n, const LaneQ& q, int64_t k, int64_t m_dup,
                                                const int64_t* __restrict__ r_eq, int64_t leq) {
  uint32_t fr = 0, em = 0;                                 // first-round survivors; rows to emit
  int len = 0, dmin = 0x7fffffff;
  const bool inter = c.builder == GMQL_B_INTERSECTION;
  for (int t = 0; t < n; ++t) {
    const int4 en = cand[t * stride];
    const int d = join_distance32(q.ls, q.le, en.x, en.y);
    // window: r.start / B <= binEnd, r.stop / B >= binStart (:120-122); strands (:123); distance and stream (:250-279)
    bool ok = (en.x <= q.wlast) & (en.y >= q.wbeg) & strand_ok(q.lst, en.w) & (d <= c.f_hi) & (d >= c.f_lo) & (en.y <= q.f_re_max) & (en.x >= q.f_rs_min);
    if (c.has_eq) { if (ok) ok = r_eq[(uint32_t)en.z] == leq; }   // :102-114
    const bool s2 = ok & (d <= c.s_hi) & (d >= c.s_lo) & (en.y <= q.s_re_max) & (en.x >= q.s_rs_min) & (!inter | ((q.ls < en.y) & (en.x < q.le)));
    fr |= (uint32_t)ok << t; len += ok;
    const bool lt = ok & (d < dmin);
    dmin = lt ? d : dmin; em = lt ? 0u : em;
    em |= (uint32_t)(s2 & (d == dmin)) << t;
  }
  if (len == 0) return 0u;
  const int64_t want = m_dup * len < k ? m_dup * len : k;   // count = min(len, k); identical left records pool their lists (:134)
  const int64_t idx = m_dup == 1 ? want - 1 : (want - 1) / m_dup;   // position in the de-duplicated sorted list (no 64-bit division in the usual case)
  if (idx > 0) {                                            // beyond the minimum: walk the distinct distances in increasing order
    int64_t seen = 0;
    int prev = 0;
    bool first = true;
    while (true) {
      int cur = 0, ncur = 0;
      bool any = false;
      for (uint32_t m = fr; m; m &= m - 1) {
        const int4 en = cand[(__ffs((int)m) - 1) * stride];
        const int d = join_distance32(q.ls, q.le, en.x, en.y);
        if (!first && d <= prev) continue;
        if (!any || d < cur) { cur = d; ncur = 1; any = true; }
        else if (d == cur) ++ncur;
      }
      seen += ncur; prev = cur; first = false;
      if (seen > idx) break;
    }
    em = 0;
    for (uint32_t m = fr; m; m &= m - 1) {
      const int t = __ffs((int)m) - 1;
      const int4 en = cand[t * stride];
      const int d = join_distance32(q.ls, q.le, en.x, en.y);
      const bool s2 = (d <= prev) & (d <= c.s_hi) & (d >= c.s_lo) & (en.y <= q.s_re_max) & (en.x >= q.s_rs_min) & (!inter | ((q.ls < en.y) & (en.x < q.le)));
      em |= (uint32_t)s2 << t;
    }
  }
  return em;
}

// one (left region, right sample) group on the unpacked index, no storage: count = min(len, k) with ties (:136-146), second round
template <typename K, bool WRITE>
__device__ int64_t md_lane_stream(const RightIndex<K>& R, const JoinCfg& cfg, int lst, int64_t ls, int64_t le, int64_t b0, int64_t leq, int64_t m_dup,
                                  uint64_t base, int64_t lo, int64_t hi, int32_t i, int p, int64_t o, int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int32_t* __restrict__ out_p) {
  int64_t len = 0;
  for (int64_t k = lo; k < hi; ++k) {
    int64_t rs = (int64_t)((uint64_t)R.start[k] - base), re = (int64_t)((uint64_t)R.stop[k] - base), d;
    if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq)) ++len;
  }
  if (len == 0) return 0;
  const int64_t want = m_dup * len < cfg.k ? m_dup * len : cfg.k;
  const int64_t idx = (want - 1) / m_dup;
  int64_t seen = 0, prev = 0;
  bool first = true;
  while (true) {                                           // walk the distinct distances in increasing order
    int64_t cur = 0, ncur = 0;
    bool any = false;
    for (int64_t k = lo; k < hi; ++k) {
      int64_t rs = (int64_t)((uint64_t)R.start[k] - base), re = (int64_t)((uint64_t)R.stop[k] - base), d;
      if (!(first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq))) continue;
      if (!first && d <= prev) continue;
      if (!any || d < cur) { cur = d; ncur = 1; any = true; }
      else if (d == cur) ++ncur;
    }
    seen += ncur; prev = cur; first = false;
    if (seen > idx) break;
  }
  int64_t mine = 0;
  for (int64_t k = lo; k < hi; ++k) {
    int64_t rs = (int64_t)((uint64_t)R.start[k] - base), re = (int64_t)((uint64_t)R.stop[k] - base), d;
    if (first_round_ok(cfg, lst, ls, le, (int)R.strand[k], rs, re, b0, &d) && (!cfg.has_eq || R.eq[R.row[k]] == leq) && d <= prev &&
        second_round_ok(cfg, lst, ls, le, rs, re, d)) {
      if (WRITE) { out_l[o + mine] = i; out_r[o + mine] = (int32_t)R.row[k]; out_p[o + mine] = p; }
      ++mine;
    }
  }
  return mine;
}

// teams = runs of up to 32 consecutive left regions of the genome order that lie on ONE chromosome (a team across a chromosome
// boundary would visit every right sample once per chromosome and could not be memoised): per chromosome the first sorted left
// slot and the first team, then per team its first slot and size.  At most ceil(n / 32) + n_chrom teams; the rest are empty.
template <typename K>
__global__ void k_team_chroms(const K* __restrict__ lstart, int64_t n, LinMap lm, int64_t* __restrict__ lb, int64_t* __restrict__ toff) {
  for (int c = threadIdx.x; c <= lm.n_chrom; c += blockDim.x)
    lb[c] = c == lm.n_chrom ? n : (lstart ? lower_bound_dev<K>(lstart, n, (K)lm.coff[c]) : 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int c = 0; c < lm.n_chrom; ++c) { toff[c] = t; t += (lb[c + 1] - lb[c] + 31) >> 5; }
    toff[lm.n_chrom] = t;
  }
}
__global__ void k_team_table(const int64_t* __restrict__ lb, const int64_t* __restrict__ toff, int n_chrom, int64_t n_teams, int2* __restrict__ teams) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_teams; t += (int64_t)gridDim.x * blockDim.x) {
    int2 r = make_int2(0, 0);
    if (t < toff[n_chrom]) {
      int lo = 0, hi = n_chrom;                            // the last chromosome whose first team is <= t
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (toff[mid] <= t) lo = mid; else hi = mid; }
      const int64_t first = lb[lo] + ((t - toff[lo]) << 5), left = lb[lo + 1] - first;
      r = make_int2((int)first, (int)(left < 32 ? left : 32));
    }
    teams[t] = r;
  }
}

template <typename K, bool WRITE>
__global__ void __launch_bounds__(256, 2) k_join_md_team(LeftSide L, RightIndex<K> R, LinMap lm, JoinCfg cfg, JoinCfg32 c32, const int32_t* __restrict__ pair_t /* [right sample][left sample] */, int n_ls, int n_rs,
                                                         const int32_t* __restrict__ mult, int64_t* __restrict__ cnt, const int64_t* __restrict__ off,
                                                         int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int32_t* __restrict__ out_p,
                                                         uint32_t* __restrict__ memo_lo, uint16_t* __restrict__ memo_em, int32_t* __restrict__ cnt_ts,
                                                         const int2* __restrict__ teams, int64_t n_teams) {
  // memo (count pass, optional): per (team, right sample) the first index of the slot's candidates (or MDT_NONE / MDT_REDO), the
  // number of rows of the whole team with that sample (cnt_ts, instead of the per-region cnt) and per lane the 16-bit mask of the
  // candidates that become rows — k_join_md_team_write then needs neither searches nor selection and writes contiguous runs
  extern __shared__ int4 s_team[];                         // per warp: 32 slots [entry][sample], then the one-sample stage
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int4* slots = s_team + (threadIdx.x >> 5) * (32 * MDT_SLOT + MDT_CAP);
  int4* big = slots + 32 * MDT_SLOT;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t team = warp; team < n_teams; team += nwarps) {
    const int2 tm = teams[team];
    const int64_t slot = (int64_t)tm.x + lane;
    bool valid = lane < tm.y;
    const bool mine_slot = valid;
    int64_t i = 0, b0 = 1, b1 = 0, m_dup = 1, leq = 0;
    int c = -1, lst = 0, la = 0, ls = 0, le = 0;
    if (valid) {
      i = L.perm ? (int64_t)L.perm[slot] : slot;
      c = L.chrom[i];
      ls = ((const int32_t*)L.start)[i]; le = ((const int32_t*)L.stop)[i];
      lst = L.strand ? (int)L.strand[i] : 0;
      la = L.sample ? L.sample[i] : 0;
      left_bins(cfg, lst, ls, le, &b0, &b1);
      m_dup = mult ? mult[i] : 1;
      leq = cfg.has_eq ? L.eq[i] : 0;
      valid = b0 <= b1 && c >= 0 && c < lm.n_chrom;
    }
    // the lane's own window in 32 bits: r.stop >= binStart * B (at most the left stop), r.start <= (binEnd + 1) * B - 1
    const int wbeg = (int)(b0 * cfg.B);
    const int64_t wl = (b1 + 1) * cfg.B - 1;
    const int wlast = wl > 0x7fffffffLL ? 0x7fffffff : (int)wl;
    LaneQ q;
    q.ls = ls; q.le = le; q.lst = lst; q.wbeg = wbeg; q.wlast = wlast;
    stream_bounds(c32.f_stream, lst, ls, le, &q.f_re_max, &q.f_rs_min);
    stream_bounds(c32.s_stream, lst, ls, le, &q.s_re_max, &q.s_rs_min);
    if (c32.never_f) q.wlast = (int)0x80000000;            // no candidate start lies at or below it
    if (c32.never_s) q.s_re_max = (int)0x80000000;         // nor a stop
    // kept in registers: recomputing these selects for every candidate costs more than the registers
    asm volatile("" : "+r"(q.wbeg), "+r"(q.wlast), "+r"(q.f_re_max), "+r"(q.f_rs_min), "+r"(q.s_re_max), "+r"(q.s_rs_min));
    const int64_t wbase = (WRITE && mine_slot) ? off[slot] : 0;
    int64_t total = 0;
    unsigned todo = __ballot_sync(FULL, valid);
    // a team that spans a chromosome boundary visits the samples once per chromosome: nothing to memoise, the write pass redoes it
    const bool memo_on = !WRITE && memo_lo != nullptr;
    if (todo == 0u) {                                      // an empty team, or no left region with a window: no rows
      if (memo_on) for (int b = lane; b < n_rs; b += 32) { memo_lo[team * n_rs + b] = MDT_NONE; cnt_ts[team * n_rs + b] = 0; }
      if (!WRITE && !memo_on && mine_slot) cnt[slot] = 0;
      continue;
    }
    const int c_first = __shfl_sync(FULL, c, todo ? __ffs(todo) - 1 : 0);   // (not inside a short-circuit: every lane takes part)
    const unsigned same = __ballot_sync(FULL, valid && c == c_first);
    const bool memo = memo_on && todo == same;
    if (memo_on && !memo) for (int b = lane; b < n_rs; b += 32) { memo_lo[team * n_rs + b] = MDT_REDO; cnt_ts[team * n_rs + b] = 0; }
    while (todo) {                                         // one pass per chromosome present in the team (nearly always one)
      const int cs = __shfl_sync(FULL, c, __ffs(todo) - 1);
      const bool act = valid && c == cs;
      todo &= ~__ballot_sync(FULL, act);
      int64_t ub0 = act ? b0 : 0x7fffffffffffffffLL, ub1 = act ? b1 : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int64_t x0 = __shfl_xor_sync(FULL, ub0, o), x1 = __shfl_xor_sync(FULL, ub1, o);
        ub0 = x0 < ub0 ? x0 : ub0; ub1 = x1 > ub1 ? x1 : ub1;
      }
      const uint64_t cbase = lm.coff[cs];
      for (int s0 = 0; s0 < n_rs; s0 += 32) {
        // lane j: sample s0 + j's candidates near the team -> its slot
        int64_t LO = 0, HI = 0;
        if (s0 + lane < n_rs) window<K>(R, lm, (uint64_t)(s0 + lane) * lm.G, cs, ub0, ub1, cfg.B, &LO, &HI);
        const int64_t nn = HI - LO;
        const int nmine = nn > MDT_CAP ? MDT_CAP + 1 : (int)nn;
        __syncwarp();                                      // the previous round's readers are done with the slots
        if (nmine <= MDT_SLOT) {
#pragma unroll
          for (int t = 0; t < MDT_SLOT; ++t) if (t < nmine) slots[t * 32 + lane] = R.packed[LO + t];   // [entry][sample]: conflict-free both ways
        }
        __syncwarp();
        if (memo && s0 + lane < n_rs) memo_lo[team * n_rs + s0 + lane] = nmine == 0 ? MDT_NONE : (nmine <= MDT_SLOT ? (uint32_t)LO : MDT_REDO);
        const int ns = n_rs - s0 < 32 ? n_rs - s0 : 32;
        int p_next = act ? pair_t[(int64_t)s0 * n_ls + la] : -1;
        int round_cnt = 0;                                 // lane j: rows of the whole team with sample s0 + j (memo mode)
        for (int jj = 0; jj < ns; ++jj) {
          const int n = __shfl_sync(FULL, nmine, jj);
          const int p = p_next;
          if (jj + 1 < ns) p_next = act ? pair_t[(int64_t)(s0 + jj + 1) * n_ls + la] : -1;
          if (n == 0) continue;                            // no candidate of this sample near the team
          const int b = s0 + jj;
          int mine = 0;
          if (n <= MDT_SLOT) {
            const uint32_t em = p >= 0 ? md_select32(c32, slots + jj, 32, n, q, cfg.k, m_dup, R.eq, leq) : 0u;
            mine = __popc(em);
            if (memo) memo_em[((team * n_rs + b) << 5) + lane] = (uint16_t)em;
            if (WRITE) { int64_t o = wbase + total; for (uint32_t m = em; m; m &= m - 1) { out_l[o] = (int32_t)i; out_r[o] = slots[(__ffs((int)m) - 1) * 32 + jj].z; out_p[o] = p; ++o; } }
          } else if (n <= MDT_CAP) {                       // more than a slot holds: stage this sample alone
            const int64_t lo = __shfl_sync(FULL, LO, jj);
            __syncwarp();
            for (int t = lane; t < n; t += 32) big[t] = R.packed[lo + t];
            __syncwarp();
            const uint32_t em = p >= 0 ? md_select32(c32, big, 1, n, q, cfg.k, m_dup, R.eq, leq) : 0u;
            mine = __popc(em);
            if (WRITE) { int64_t o = wbase + total; for (uint32_t m = em; m; m &= m - 1) { out_l[o] = (int32_t)i; out_r[o] = big[__ffs((int)m) - 1].z; out_p[o] = p; ++o; } }
          } else if (p >= 0) {                             // a dense right sample: every lane on its own, as k_join_md's streaming path
            int64_t l1, h1;
            window<K>(R, lm, (uint64_t)b * lm.G, cs, b0, b1, cfg.B, &l1, &h1);
            mine = (int)md_lane_stream<K, WRITE>(R, cfg, lst, ls, le, b0, leq, m_dup, (uint64_t)b * lm.G + cbase, l1, h1, (int32_t)i, p, wbase + total, out_l, out_r, out_p);
          }
          total += mine;
          if (memo_on) { const int tot = __reduce_add_sync(FULL, mine); if (lane == jj) round_cnt = tot; }
        }
        if (memo_on && s0 + lane < n_rs) {
          if (memo) cnt_ts[team * n_rs + s0 + lane] = round_cnt; else cnt_ts[team * n_rs + s0 + lane] += round_cnt;
        }
      }
    }
    if (!WRITE && !memo_on && mine_slot) cnt[slot] = total;
  }
}

// write pass of the team form from the count pass's memo.  The order of the output rows is free, so the rows of a (team, right
// sample) entry go out as ONE contiguous run at the entry's offset (exclusive scan of cnt_ts): lanes in order, a lane's rows in
// index order.  Entries that were not memoised (a sample with more candidates than a slot, a team across a chromosome boundary)
// are redone lane by lane on the index — the same exact selection, so the same rows as were counted.
template <typename K>
__global__ void __launch_bounds__(256, 2) k_join_md_team_write(LeftSide L, RightIndex<K> R, LinMap lm, JoinCfg cfg, const int32_t* __restrict__ pair_t, int n_ls, int n_rs,
                                                            const int32_t* __restrict__ mult, const int64_t* __restrict__ off_ts, const uint32_t* __restrict__ memo_lo, const uint16_t* __restrict__ memo_em,
                                                            int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int32_t* __restrict__ out_p,
                                                            const int2* __restrict__ teams, int64_t n_teams) {
  __shared__ __align__(16) uint16_t s_em_all[8][32 * 32];  // per warp: the masks of a round of 32 entries
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint16_t* s_em = s_em_all[threadIdx.x >> 5];
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t team = warp; team < n_teams; team += nwarps) {
    const int2 tm = teams[team];
    if (tm.y == 0) continue;
    const int64_t slot = (int64_t)tm.x + lane;
    const bool valid = lane < tm.y;
    const int64_t i = valid ? (L.perm ? (int64_t)L.perm[slot] : slot) : 0;
    const int la = (valid && L.sample) ? L.sample[i] : 0;
    for (int s0 = 0; s0 < n_rs; s0 += 32) {
      const int64_t e0 = team * n_rs + s0;
      const bool has = s0 + lane < n_rs;
      const uint32_t lo_j = has ? memo_lo[e0 + lane] : MDT_NONE;
      const int64_t base_j = has ? off_ts[e0 + lane] : 0;
      const int ns = n_rs - s0 < 32 ? n_rs - s0 : 32;
      // memoised entries.  The round's masks are one contiguous block (64 bytes per entry): four independent 16-byte loads per lane
      // bring them into shared memory; then eight entries at a time — nothing in an entry depends on another, so their pair and
      // first-row loads are issued back to back and the latencies overlap (a warp meets ~400 dependent global loads per team otherwise)
      {
        const uint4* src = (const uint4*)(memo_em + (e0 << 5));
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k * 32 + lane < ns * 4) v[k] = src[k * 32 + lane];
        __syncwarp();                                      // the previous round's readers are done
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k * 32 + lane < ns * 4) ((uint4*)s_em)[k * 32 + lane] = v[k];
        __syncwarp();
      }
      for (int j0 = 0; j0 < ns; j0 += MDT_WB) {
        uint32_t lo8[MDT_WB], em8[MDT_WB];
#pragma unroll
        for (int u = 0; u < MDT_WB; ++u) {
          lo8[u] = __shfl_sync(FULL, lo_j, (j0 + u) & 31);
          em8[u] = (j0 + u < ns && lo8[u] < MDT_REDO) ? (uint32_t)s_em[((j0 + u) << 5) + lane] : 0u;   // (never written for other entries)
        }
        int64_t o8[MDT_WB];
        int p8[MDT_WB], r8[MDT_WB];
#pragma unroll
        for (int u = 0; u < MDT_WB; ++u) {
          const int c = __popc(em8[u]);
          int incl = c;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
          o8[u] = __shfl_sync(FULL, base_j, (j0 + u) & 31) + incl - c;
          p8[u] = 0; r8[u] = 0;
          if (em8[u]) { p8[u] = pair_t[(int64_t)(s0 + j0 + u) * n_ls + la]; r8[u] = R.packed[lo8[u] + (__ffs((int)em8[u]) - 1)].z; }
        }
#pragma unroll
        for (int u = 0; u < MDT_WB; ++u) {
          uint32_t em = em8[u];
          if (em) {
            int64_t o = o8[u];
            out_l[o] = (int32_t)i; out_r[o] = r8[u]; out_p[o] = p8[u];
            for (em &= em - 1; em; em &= em - 1) { ++o; out_l[o] = (int32_t)i; out_r[o] = R.packed[lo8[u] + (__ffs((int)em) - 1)].z; out_p[o] = p8[u]; }
          }
        }
      }
      // entries that were not memoised: count, place, write — lane by lane on the index
      unsigned redo = __ballot_sync(FULL, lo_j == MDT_REDO);
      for (; redo; redo &= redo - 1) {
        const int jj = __ffs((int)redo) - 1, b = s0 + jj;
        const int64_t base = __shfl_sync(FULL, base_j, jj);
        int p = -1, c = -1, lst = 0;
        int64_t ls = 0, le = 0, b0 = 1, b1 = 0, l1 = 0, h1 = 0, leq = 0, m_dup = 1;
        if (valid) {
          p = pair_t[(int64_t)b * n_ls + la];
          c = L.chrom[i];
          ls = ld_coord(L.start, L.bits, i); le = ld_coord(L.stop, L.bits, i);
          lst = L.strand ? (int)L.strand[i] : 0;
          left_bins(cfg, lst, ls, le, &b0, &b1);
          leq = cfg.has_eq ? L.eq[i] : 0;
          m_dup = mult ? mult[i] : 1;
        }
        const bool go = valid && p >= 0 && c >= 0 && c < lm.n_chrom && b0 <= b1;
        int mine = 0;
        uint64_t gbase = 0;
        if (go) {
          gbase = (uint64_t)b * lm.G + lm.coff[c];
          window<K>(R, lm, (uint64_t)b * lm.G, c, b0, b1, cfg.B, &l1, &h1);
          mine = (int)md_lane_stream<K, false>(R, cfg, lst, ls, le, b0, leq, m_dup, gbase, l1, h1, (int32_t)i, p, 0, out_l, out_r, out_p);
        }
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
        if (mine) md_lane_stream<K, true>(R, cfg, lst, ls, le, b0, leq, m_dup, gbase, l1, h1, (int32_t)i, p, base + incl - mine, out_l, out_r, out_p);
      }
    }
  }
}

// per sorted index entry: strand and sample of the right row
__global__ void k_join_payload(const uint32_t* __restrict__ rows, int64_t n, const int8_t* __restrict__ strand, const int32_t* __restrict__ sample, int8_t* __restrict__ s_strand, int32_t* __restrict__ s_sample) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r = rows[k];
    s_strand[k] = strand ? strand[r] : (int8_t)0;
    s_sample[k] = sample ? sample[r] : 0;
  }
}

// one 16-byte record per sorted entry: everything the MD round needs about a candidate in one load
// `sample` given (the pooled index of joins without MD): w = strand | sample << 2, so a candidate is ONE 16-byte load
__global__ void k_join_packed(const uint32_t* __restrict__ rows, int64_t n, const int32_t* __restrict__ start, const int32_t* __restrict__ stop, const int8_t* __restrict__ strand,
                              const int32_t* __restrict__ sample, int4* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r = rows[k];
    out[k] = make_int4(start[r], stop[r], (int)r, (strand ? (int)strand[r] : 0) | (sample ? sample[r] << 2 : 0));
  }
}

__global__ void k_sample_classes(const int32_t* __restrict__ sample, int64_t n, int n_samples, uint32_t* __restrict__ cls, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = sample ? sample[i] : 0;
    if (s < 0 || s >= n_samples) { *bad = 1; s = 0; }
    cls[i] = (uint32_t)s;
  }
}

__global__ void k_check_samples(const int32_t* __restrict__ sample, int64_t n, int n_samples, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = sample ? sample[i] : 0;
    if (s < 0 || s >= n_samples) *bad = 1;
  }
}

__global__ void k_dup_counts(const int32_t* __restrict__ dup_of, int64_t n, int32_t* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(&cnt[dup_of[i]], 1);
}
__global__ void k_dup_mult(const int32_t* __restrict__ dup_of, int64_t n, const int32_t* __restrict__ cnt, int32_t* __restrict__ mult) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mult[i] = cnt[dup_of[i]];
}

// built region (joinRegions :345-372)
__global__ void k_join_build(const int32_t* __restrict__ ol, const int32_t* __restrict__ orr, int64_t n, LeftSide L, const int32_t* __restrict__ r_chrom, const void* __restrict__ r_start, const void* __restrict__ r_stop, int r_bits,
                             const int8_t* __restrict__ r_strand, int builder, int32_t* __restrict__ chrom, int64_t* __restrict__ start, int64_t* __restrict__ stop, int8_t* __restrict__ strand) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    int32_t i = ol[o], j = orr[o];
    int64_t ls = 0, le = 0, rs = 0, re = 0;
    int lst = 0, rst = 0, lc = 0, rc = 0;
    if (i >= 0) { ls = ld_coord(L.start, L.bits, i); le = ld_coord(L.stop, L.bits, i); lst = L.strand ? (int)L.strand[i] : 0; lc = L.chrom[i]; }
    if (j >= 0) { rs = ld_coord(r_start, r_bits, j); re = ld_coord(r_stop, r_bits, j); rst = r_strand ? (int)r_strand[j] : 0; rc = r_chrom[j]; }
    int64_t s, e; int st, c;
    switch (builder) {
      case GMQL_B_RIGHT: case GMQL_B_RIGHT_DISTINCT: s = rs; e = re; st = rst; c = rc; break;
      case GMQL_B_CONTIG: s = ls < rs ? ls : rs; e = le > re ? le : re; st = lst; c = lc; break;
      case GMQL_B_INTERSECTION: s = ls > rs ? ls : rs; e = le < re ? le : re; st = (lst == rst) ? lst : 0; c = lc; break;
      default: s = ls; e = le; st = lst; c = lc; break;   // LEFT, LEFT_DISTINCT, BOTH
    }
    chrom[o] = c; start[o] = s; stop[o] = e; strand[o] = (int8_t)st;
  }
}

// DISTINCT: key = pair << 32 | canonical row
__global__ void k_distinct_keys(const int32_t* __restrict__ rows, const int32_t* __restrict__ pair, int64_t n, const int32_t* __restrict__ dup_of, uint64_t* __restrict__ keys) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    int32_t r = rows[o];
    if (dup_of) r = dup_of[r];
    keys[o] = ((uint64_t)(uint32_t)pair[o] << 32) | (uint64_t)(uint32_t)r;
  }
}
__global__ void k_unique_flags(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) flag[o] = (o == 0 || keys[o] != keys[o - 1]) ? 1 : 0;
}
__global__ void k_unique_write(const uint64_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ flag, const int64_t* __restrict__ excl, int left_side,
                               int32_t* __restrict__ ol, int32_t* __restrict__ orr, int32_t* __restrict__ op) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    if (!flag[o]) continue;
    int64_t w = excl[o];
    int32_t row = (int32_t)(uint32_t)(keys[o] & 0xffffffffull), p = (int32_t)(uint32_t)(keys[o] >> 32);
    ol[w] = left_side ? row : -1;
    orr[w] = left_side ? -1 : row;
    op[w] = p;
  }
}

static void parse_atoms(const gmql_atom* atoms, int n_atoms, int64_t max_bin_distance, JoinCfg* c) {
  GMQL_REQUIRE(n_atoms >= 1 && n_atoms <= 8 && atoms, GMQL_ERR_INVALID, "a join condition needs 1..8 atoms");
  memset(c, 0, sizeof *c);
  int i = 0;
  // createExecutionParameters (:388-404): later atoms of one kind overwrite earlier ones
  for (; i < n_atoms && atoms[i].kind != GMQL_J_MINDIST; ++i) {
    switch (atoms[i].kind) {
      case GMQL_J_DISTLESS: c->f_has_max = 1; c->f_max = atoms[i].arg; break;
      case GMQL_J_DISTGREATER: c->f_has_min = 1; c->f_min = atoms[i].arg; break;
      case GMQL_J_UPSTREAM: c->f_stream = 1; break;
      case GMQL_J_DOWNSTREAM: c->f_stream = 2; break;
      default: throw gmql_error(GMQL_ERR_INVALID, "unknown join atom");
    }
  }
  if (i < n_atoms) {
    c->has_md = 1;
    c->k = atoms[i].arg;
    GMQL_REQUIRE(c->k >= 1, GMQL_ERR_INVALID, "MD(k) needs k >= 1");
    ++i;
    for (; i < n_atoms; ++i) {
      switch (atoms[i].kind) {
        case GMQL_J_DISTLESS: c->s_has_max = 1; c->s_max = atoms[i].arg; break;
        case GMQL_J_DISTGREATER: c->s_has_min = 1; c->s_min = atoms[i].arg; break;
        case GMQL_J_UPSTREAM: c->s_stream = 1; break;
        case GMQL_J_DOWNSTREAM: c->s_stream = 2; break;
        case GMQL_J_MINDIST: throw gmql_error(GMQL_ERR_INVALID, "a second MD atom is a MatchError in the reference (createExecutionParameters)");
        default: throw gmql_error(GMQL_ERR_INVALID, "unknown join atom");
      }
    }
  }
  // maxDistance (:75-79)
  if (c->f_has_max) c->maxD = c->f_max > 0 ? c->f_max : 0;
  else if (c->s_has_max) c->maxD = c->s_max > max_bin_distance ? c->s_max : max_bin_distance;
  else if (c->f_has_min) c->maxD = c->f_min + max_bin_distance;
  else c->maxD = max_bin_distance;
}

template <typename K>
static void run_join(gmql_ctx* ctx, const DevRegions& left, const DevRegions& right, const gmql_pair* pairs, int64_t n_pairs, JoinCfg cfg, bool rows_only,
                     const int64_t* d_left_eq, const int64_t* d_right_eq, const SpanTable& spans, gmql_result* res) {
  const int64_t nl = left.n, nr = right.n;
  const int n_ls = left.n_samples, n_rs = right.n_samples;
  LinMap lm = spans.map();
  DevBuf<int> bad(ctx, 1);
  bad.zero();

  // pair tables (host-built: n_pairs is #samples^2 at most)
  std::vector<int32_t> h_pair((size_t)n_ls * (size_t)n_rs, -1);
  std::vector<int64_t> h_aoff((size_t)n_ls + 1, 0);
  for (int64_t p = 0; p < n_pairs; ++p) {
    int a = pairs[p].left_sample, b = pairs[p].right_sample;
    GMQL_REQUIRE(a >= 0 && a < n_ls && b >= 0 && b < n_rs, GMQL_ERR_INVALID, "pair refers to a sample index out of range");
    GMQL_REQUIRE(h_pair[(size_t)a * n_rs + b] < 0, GMQL_ERR_INVALID, "the same (left sample, right sample) pair is listed twice");
    h_pair[(size_t)a * n_rs + b] = (int32_t)p;
    h_aoff[a + 1]++;
  }
  for (int a = 0; a < n_ls; ++a) h_aoff[a + 1] += h_aoff[a];
  std::vector<int32_t> h_ars((size_t)n_pairs), h_apair((size_t)n_pairs);
  {
    std::vector<int64_t> fill(h_aoff.begin(), h_aoff.end() - 1);
    for (int a = 0; a < n_ls; ++a)
      for (int b = 0; b < n_rs; ++b) {
        int32_t p = h_pair[(size_t)a * n_rs + b];
        if (p >= 0) { h_ars[fill[a]] = b; h_apair[fill[a]] = p; fill[a]++; }
      }
  }
  DevBuf<int32_t> d_pair(ctx, (int64_t)n_ls * n_rs), d_ars(ctx, n_pairs), d_apair(ctx, n_pairs);
  DevBuf<int64_t> d_aoff(ctx, n_ls + 1);
  CUDA_TRY(cudaMemcpyAsync(d_pair.p, h_pair.data(), 4 * h_pair.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(d_aoff.p, h_aoff.data(), 8 * h_aoff.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (n_pairs) {
    CUDA_TRY(cudaMemcpyAsync(d_ars.p, h_ars.data(), 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_apair.p, h_apair.data(), 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(GMQL_SYNC(ctx));   // host vectors go out of scope below
  cudaEventRecord(ctx->ev_h1, ctx->stream);
  cudaEventRecord(ctx->ev_k0, ctx->stream);

  if (nl) LAUNCH(ctx, k_check_samples, grid_for(ctx, nl, 256), 256, 0, left.sample, nl, n_ls, bad.p);

  // right index: pooled (no MD) or keyed by sample (MD)
  DevBuf<uint32_t> cls;
  uint64_t n_blocks = 1;
  if (cfg.has_md) {
    cls.alloc(ctx, nr);
    if (nr) LAUNCH(ctx, k_sample_classes, grid_for(ctx, nr, 256), 256, 0, right.sample, nr, n_rs, cls.p, bad.p);
    n_blocks = (uint64_t)n_rs;
  } else if (nr) {
    LAUNCH(ctx, k_check_samples, grid_for(ctx, nr, 256), 256, 0, right.sample, nr, n_rs, bad.p);
  }
  IntervalIndex<K> ix;
  build_index<K>(ctx, right, cfg.has_md ? cls.p : nullptr, n_blocks, lm, &ix);
  DevBuf<int8_t> s_strand(ctx, nr);
  DevBuf<int32_t> s_sample(ctx, nr);
  if (nr) LAUNCH(ctx, k_join_payload, grid_for(ctx, nr, 256), 256, 0, ix.row, nr, right.strand, right.sample, s_strand.p, s_sample.p);

  int h_bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "sample index out of range in a region dataset");

  // MD: visit the left regions in genome order, so that successive warps walk every right sample's index block forward and the
  // few cache lines a (left region, right sample) group touches are still in L2 (the order of the output rows is free)
  IntervalIndex<K> lix;
  if (cfg.has_md && nl > 1) build_index<K>(ctx, left, nullptr, 1, lm, &lix);
  LeftSide L;
  L.perm = (cfg.has_md && nl > 1) ? lix.row : nullptr;
  L.chrom = left.chrom; L.start = left.start; L.stop = left.stop; L.bits = left.coord_bits; L.strand = left.strand; L.sample = left.sample; L.eq = d_left_eq; L.n = nl;
  // bucket table over the linear start key (about two entries per bucket, at most 2^22 buckets)
  const uint64_t max_key = n_blocks * lm.G;
  int tbl_shift = 0;
  {
    int64_t want = std::max<int64_t>(1024, std::min<int64_t>((int64_t)1 << 22, nr / 2 + 1));
    while (tbl_shift < 62 && (int64_t)(max_key >> tbl_shift) + 2 > want) ++tbl_shift;
  }
  const int64_t n_buckets = (int64_t)(max_key >> tbl_shift) + 1;
  DevBuf<uint32_t> tbl(ctx, n_buckets + 2);
  LAUNCH(ctx, (k_start_table<K>), grid_for(ctx, n_buckets + 1, 256), 256, 0, ix.start, nr, tbl_shift, n_buckets, tbl.p);
  DevBuf<int4> packed;
  if (right.coord_bits == 32 && nr && (cfg.has_md || (right.sample != nullptr && n_rs < (1 << 29)))) {
    packed.alloc(ctx, nr);
    LAUNCH(ctx, k_join_packed, grid_for(ctx, nr, 256), 256, 0, ix.row, nr, (const int32_t*)right.start, (const int32_t*)right.stop, right.strand,
           cfg.has_md ? (const int32_t*)nullptr : right.sample, packed.p);
  }
  RightIndex<K> R;
  R.packed = packed.p;
  R.tbl = tbl.p; R.tbl_shift = tbl_shift; R.n_buckets = n_buckets;
  R.start = ix.start; R.stop = ix.stop.p; R.pmax = ix.pmax.p; R.row = ix.row; R.n = nr; R.strand = s_strand.p; R.sample = s_sample.p; R.eq = d_right_eq;

  // multiplicity of identical left records (MapKey pooling in the MD round)
  DevBuf<int32_t> mult;
  if (cfg.has_md && left.dup_of && nl) {
    DevBuf<int32_t> dc(ctx, nl);
    dc.zero();
    mult.alloc(ctx, nl);
    LAUNCH(ctx, k_dup_counts, grid_for(ctx, nl, 256), 256, 0, left.dup_of, nl, dc.p);
    LAUNCH(ctx, k_dup_mult, grid_for(ctx, nl, 256), 256, 0, left.dup_of, nl, dc.p, mult.p);
  }

  // MD: the team kernel when the right coordinates are packed, the left regions are visited in genome order and the pair table is
  // dense enough that a team meets most samples anyway (it walks all right samples); option "join.md" = "team" | "lane" forces one
  const bool team_ok = cfg.has_md && R.packed && left.coord_bits == 32;
  bool md_team = team_ok && L.perm && n_pairs * 4 >= (int64_t)n_ls * n_rs;
  if (const char* o = ctx_opt(ctx, "join.md")) {
    if (!strcmp(o, "team")) md_team = team_ok;
    else if (!strcmp(o, "lane")) md_team = false;
  }
  DevBuf<int32_t> d_pair_t;
  if (md_team) {
    std::vector<int32_t> h_pair_t((size_t)n_ls * (size_t)n_rs);
    for (int a = 0; a < n_ls; ++a)
      for (int b = 0; b < n_rs; ++b) h_pair_t[(size_t)b * n_ls + a] = h_pair[(size_t)a * n_rs + b];
    d_pair_t.alloc(ctx, (int64_t)n_ls * n_rs);
    CUDA_TRY(cudaMemcpyAsync(d_pair_t.p, h_pair_t.data(), 4 * h_pair_t.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
  }
  // memo of the count pass for the write pass (6 bytes per (left region, allowed right sample) group), when it fits
  DevBuf<uint32_t> memo_lo;
  DevBuf<uint16_t> memo_ok;
  int64_t memo_stride = 0;
  if (cfg.has_md && R.packed && nl && !md_team) {
    for (int a = 0; a < n_ls; ++a) memo_stride = std::max<int64_t>(memo_stride, h_aoff[a + 1] - h_aoff[a]);
    if (memo_stride > 0 && nl * memo_stride <= (int64_t)700000000) { memo_lo.alloc(ctx, nl * memo_stride); memo_ok.alloc(ctx, nl * memo_stride); }
  }
  const int64_t n_teams = md_team ? ceil_div64(nl, 32) + lm.n_chrom : 0;
  const int team_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_teams, 8), (int64_t)ctx->num_sms * 8));
  DevBuf<int2> teams;
  if (md_team && nl) {
    DevBuf<int64_t> lb(ctx, lm.n_chrom + 1), toff(ctx, lm.n_chrom + 1);
    teams.alloc(ctx, n_teams);
    LAUNCH(ctx, (k_team_chroms<K>), 1, 256, 0, (const K*)(L.perm ? lix.start : nullptr), nl, lm, lb.p, toff.p);
    LAUNCH(ctx, k_team_table, grid_for(ctx, n_teams, 256), 256, 0, lb.p, toff.p, lm.n_chrom, n_teams, teams.p);
  }
  // the team form's memo for its write pass: 80 bytes per (team, right sample) with the offsets, when it fits (option "join.md_memo" = "0": never)
  const size_t team_smem = 8 * (32 * MDT_SLOT + MDT_CAP) * sizeof(int4);
  if (md_team) {
    CUDA_TRY(cudaFuncSetAttribute(k_join_md_team<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_join_md_team<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_smem));
  }
  DevBuf<uint32_t> team_lo;
  DevBuf<uint16_t> team_em;
  DevBuf<int32_t> team_cnt;
  {
    const int64_t n_e = n_teams * n_rs;
    const char* o = ctx_opt(ctx, "join.md_memo");
    if (md_team && n_e > 0 && n_e * 80 <= ((int64_t)1 << 30) && !(o && !strcmp(o, "0"))) { team_lo.alloc(ctx, n_e); team_em.alloc(ctx, n_e * 32); team_cnt.alloc(ctx, n_e); }
  }

  DevBuf<int64_t> cnt(ctx, nl + 1);
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(nl, 8), (int64_t)ctx->num_sms * 16));
  if (nl) {
    if (md_team) LAUNCH(ctx, (k_join_md_team<K, false>), team_blocks, 256, team_smem, L, R, lm, cfg, narrow_cfg(cfg), d_pair_t.p, n_ls, n_rs, mult.p, cnt.p, (const int64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                        team_lo.p, team_em.p, team_cnt.p, teams.p, n_teams);
    else if (cfg.has_md) LAUNCH(ctx, (k_join_md<K, false>), blocks, 256, 0, L, R, lm, cfg, d_aoff.p, d_ars.p, d_apair.p, mult.p, cnt.p, (const int64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                           memo_lo.p, memo_ok.p, memo_stride);
    else LAUNCH(ctx, (k_join_scan<K, false>), blocks, 256, 0, L, R, lm, cfg, d_pair.p, n_rs, cnt.p, (const int64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr);
  }
  // offsets: per left region (slot), or per (team, right sample) entry when the team kernel memoised its count pass
  DevBuf<int64_t> team_off;
  const int64_t n_ts = team_lo.p ? n_teams * n_rs : 0;
  const int64_t* p_total = cnt.p + nl;
  if (n_ts) {
    team_off.alloc(ctx, n_ts + 1);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, team_cnt.p, team_off.p, n_ts, team_off.p + n_ts);
    p_total = team_off.p + n_ts;
  } else device_scan<int64_t, int64_t, OpSum<int64_t>, false>(ctx, cnt.p, cnt.p, nl, cnt.p + nl);
  int64_t n_out = 0;
  CUDA_TRY(cudaMemcpyAsync(&n_out, p_total, sizeof n_out, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(n_out < (int64_t)0x7fffffff * 8, GMQL_ERR_UNSUPPORTED, "join output too large for this version");
  DevBuf<int32_t> ol(ctx, n_out), orr(ctx, n_out), op(ctx, n_out);
  if (nl && n_out) {
    if (md_team && team_lo.p) LAUNCH(ctx, (k_join_md_team_write<K>), team_blocks, 256, 0, L, R, lm, cfg, d_pair_t.p, n_ls, n_rs, mult.p, team_off.p, team_lo.p, team_em.p, ol.p, orr.p, op.p, teams.p, n_teams);
    else if (md_team) LAUNCH(ctx, (k_join_md_team<K, true>), team_blocks, 256, team_smem, L, R, lm, cfg, narrow_cfg(cfg), d_pair_t.p, n_ls, n_rs, mult.p, (int64_t*)nullptr, cnt.p, ol.p, orr.p, op.p,
                             (uint32_t*)nullptr, (uint16_t*)nullptr, (int32_t*)nullptr, teams.p, n_teams);
    else if (cfg.has_md) LAUNCH(ctx, (k_join_md<K, true>), blocks, 256, 0, L, R, lm, cfg, d_aoff.p, d_ars.p, d_apair.p, mult.p, (int64_t*)nullptr, cnt.p, ol.p, orr.p, op.p,
                           memo_lo.p, memo_ok.p, memo_stride);
    else LAUNCH(ctx, (k_join_scan<K, true>), blocks, 256, 0, L, R, lm, cfg, d_pair.p, n_rs, (int64_t*)nullptr, cnt.p, ol.p, orr.p, op.p);
  }

  // DISTINCT builders (:205-218): sort-unique on (pair, canonical row)
  if (cfg.builder == GMQL_B_LEFT_DISTINCT || cfg.builder == GMQL_B_RIGHT_DISTINCT) {
    bool ls = cfg.builder == GMQL_B_LEFT_DISTINCT;
    DevBuf<uint64_t> ka(ctx, n_out), kb(ctx, n_out);
    int g = grid_for(ctx, n_out, 256);
    if (n_out) LAUNCH(ctx, k_distinct_keys, g, 256, 0, ls ? ol.p : orr.p, op.p, n_out, ls ? left.dup_of : right.dup_of, ka.p);
    uint64_t* keys = ka.p;
    GMQL_REQUIRE(n_out < (int64_t)0xffffffffLL, GMQL_ERR_UNSUPPORTED, "DISTINCT over more than 2^32 candidate pairs");
    radix_sort<uint64_t>(ctx, ka.p, kb.p, nullptr, nullptr, n_out, 32 + bits_for((uint64_t)std::max<int64_t>(n_pairs, 1)), &keys, nullptr);
    DevBuf<int32_t> flag(ctx, n_out);
    DevBuf<int64_t> excl(ctx, n_out + 1);
    if (n_out) LAUNCH(ctx, k_unique_flags, g, 256, 0, keys, n_out, flag.p);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, flag.p, excl.p, n_out, excl.p + n_out);
    int64_t n_u = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_u, excl.p + n_out, sizeof n_u, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
    DevBuf<int32_t> ul(ctx, n_u), ur(ctx, n_u), up(ctx, n_u);
    if (n_out) LAUNCH(ctx, k_unique_write, g, 256, 0, keys, n_out, flag.p, excl.p, ls ? 1 : 0, ul.p, ur.p, up.p);
    ol = std::move(ul); orr = std::move(ur); op = std::move(up);
    n_out = n_u;
  }

  res->rows = n_out;
  res->kind = 3;
  if (!rows_only) {
    DevBuf<int32_t> bc(ctx, n_out);
    DevBuf<int64_t> bs(ctx, n_out), be(ctx, n_out);
    DevBuf<int8_t> bst(ctx, n_out);
    if (n_out) LAUNCH(ctx, k_join_build, grid_for(ctx, n_out, 256), 256, 0, ol.p, orr.p, n_out, L, right.chrom, right.start, right.stop, right.coord_bits, right.strand, cfg.builder,
                      bc.p, bs.p, be.p, bst.p);
    result_add(res, GMQL_COL_JOIN_CHROM, bc);
    result_add(res, GMQL_COL_JOIN_START, bs);
    result_add(res, GMQL_COL_JOIN_STOP, be);
    result_add(res, GMQL_COL_JOIN_STRAND, bst);
  }
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  result_add(res, GMQL_COL_JOIN_LEFT_ROW, ol);
  result_add(res, GMQL_COL_JOIN_RIGHT_ROW, orr);
  result_add(res, GMQL_COL_JOIN_PAIR, op);
}

}  // namespace

extern "C" int gmql_join(gmql_ctx* ctx, const gmql_regions* left, const gmql_regions* right, const gmql_pair* pairs, int64_t n_pairs,
                         const gmql_atom* atoms, int32_t n_atoms, int32_t builder, int64_t bin_size, int64_t max_bin_distance,
                         const int64_t* left_eq, const int64_t* right_eq, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(left && right, GMQL_ERR_INVALID, "null dataset");
    GMQL_REQUIRE(n_pairs >= 0 && (n_pairs == 0 || pairs), GMQL_ERR_INVALID, "bad pair list");
    GMQL_REQUIRE(left->n_chrom == right->n_chrom, GMQL_ERR_INVALID, "left and right must share one chromosome dictionary (n_chrom differs)");
    GMQL_REQUIRE(bin_size > 0 && max_bin_distance >= 0, GMQL_ERR_INVALID, "bin_size must be positive (BinSize.Join = 5000)");
    bool rows_only = (builder & GMQL_JOIN_ROWS_ONLY) != 0;
    builder &= ~GMQL_JOIN_ROWS_ONLY;
    GMQL_REQUIRE(builder >= GMQL_B_LEFT && builder <= GMQL_B_BOTH, GMQL_ERR_INVALID, "unknown region builder");
    GMQL_REQUIRE((left_eq == nullptr) == (right_eq == nullptr), GMQL_ERR_INVALID, "left_eq and right_eq must be given together");
    JoinCfg cfg;
    parse_atoms(atoms, n_atoms, max_bin_distance, &cfg);
    cfg.B = bin_size;
    cfg.builder = builder;
    cfg.has_eq = left_eq ? 1 : 0;
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions dl, dr;
    gmql_upload_regions(ctx, left, std::vector<int>(), &dl);
    gmql_upload_regions(ctx, right, std::vector<int>(), &dr);
    DevBuf<int64_t> d_leq, d_req;
    const int64_t* p_leq = nullptr; const int64_t* p_req = nullptr;
    if (left_eq) {
      if (left->location == GMQL_LOC_DEVICE) { p_leq = left_eq; }
      else { d_leq.alloc(ctx, dl.n); if (dl.n) CUDA_TRY(cudaMemcpyAsync(d_leq.p, left_eq, 8 * (size_t)dl.n, cudaMemcpyHostToDevice, ctx->stream)); p_leq = d_leq.p; }
      if (right->location == GMQL_LOC_DEVICE) { p_req = right_eq; }
      else { d_req.alloc(ctx, dr.n); if (dr.n) CUDA_TRY(cudaMemcpyAsync(d_req.p, right_eq, 8 * (size_t)dr.n, cudaMemcpyHostToDevice, ctx->stream)); p_req = d_req.p; }
    }
    SpanTable spans;
    build_span_table(ctx, &dr, nullptr, &spans);
    res = new gmql_result();
    res->owner = ctx;
    run_join<uint64_t>(ctx, dl, dr, pairs, n_pairs, cfg, rows_only, p_leq, p_req, spans, res);
    CUDA_TRY(GMQL_SYNC(ctx));
    *out = res;
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
