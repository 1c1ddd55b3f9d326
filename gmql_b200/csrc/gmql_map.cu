// gmql_map.cu — genometric MAP on the device.
//
// Replaces GenometricMap71.apply/execute (GMQL-Spark/.../RegionsOperators/GenometricMap/GenometricMap71.scala:28-150)
// and the aggregate folds of DefaultRegionsToRegionFactory (GMQL-Server/.../DefaultRegionsToRegionFactory.scala:58-126).
//
// Design (not a port of the bin/cogroup/reduceByKey plan): the *small* side (reference regions, all ref samples
// pooled) becomes an interval index (sorted linear starts + running max stop); the *large* side (experiment regions)
// is streamed once from HBM in its original order, each region finds the reference regions it overlaps and
// accumulates straight into the output row of (pair, ref region).  The output table — one row per (pair, ref region),
// zero rows included (GenometricMap71.scala:115-116) — is written by a fill kernel and a finalise kernel.
#include <cooperative_groups.h>
#include "index.cuh"

namespace {

struct MapAggPlan {
  // per distinct input column
  int n_cols = 0;
  int col_id[8];
  int need_sum[8], need_min[8], need_max[8];
  // per aggregate
  int n_aggs = 0;
  int kind[16], slot[16];  // slot = index into the distinct-column tables
  int need_bag = 0;        // some aggregate is BAG / BAGD: emit the matched experiment rows of every output row
};

struct MapAcc {
  int* count;                    // [n_rows]
  int* nn[8];                    // non-null contributors per distinct column
  double* sum[8];
  unsigned long long* mn[8];
  unsigned long long* mx[8];
};

struct ExpCols {
  const double* v[8];
  const uint8_t* ok[8];
};

__device__ __forceinline__ unsigned long long dbl_key_min(double v) {
  // order-preserving map double -> uint64 (-0.0 < +0.0 as in java.lang.Math.min); NaN maps to 0 so that it wins a min
  if (v != v) return 0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long dbl_key_max(double v) {
  if (v != v) return ~0ull;  // NaN wins a max (Math.max)
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dbl_from_key(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// ---- reference bookkeeping ---------------------------------------------------------------------
__global__ void k_sample_keys(const int32_t* __restrict__ sample, int64_t n, int n_samples, uint32_t* __restrict__ keys, uint32_t* __restrict__ rows, unsigned int* __restrict__ counts, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = sample ? sample[i] : 0;
    if (s < 0 || s >= n_samples) { *bad = 1; s = 0; }
    keys[i] = (uint32_t)s;
    rows[i] = (uint32_t)i;
    // warp-aggregated histogram: neighbouring rows usually belong to one sample
    unsigned act = __activemask();
    unsigned peers = __match_any_sync(act, s);
    if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&counts[s], (unsigned)__popc(peers));
  }
}

__global__ void k_iota_rows(uint32_t* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

// rank of each ref row inside its sample (ascending original row index), for canonical rows of dup groups
__global__ void k_ref_rank(const uint32_t* __restrict__ rows_by_sample, const uint32_t* __restrict__ keys_sorted, const int64_t* __restrict__ sample_off, int64_t n, uint32_t* __restrict__ rank_of_row) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t row = rows_by_sample[k];
    rank_of_row[row] = (uint32_t)(k - sample_off[keys_sorted[k]]);
  }
}

__global__ void k_pair_rows(const int32_t* __restrict__ pl, const int32_t* __restrict__ pr, int64_t n_pairs, int n_ls, int n_rs, const unsigned int* __restrict__ sample_cnt, int64_t* __restrict__ rows_of_pair, int* __restrict__ bad) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    int a = pl[p], b = pr[p];
    if (a < 0 || a >= n_ls || b < 0 || b >= n_rs) { *bad = 1; rows_of_pair[p] = 0; continue; }
    rows_of_pair[p] = (int64_t)sample_cnt[a];
  }
}

__global__ void k_pair_base(const int32_t* __restrict__ pl, const int32_t* __restrict__ pr, int64_t n_pairs, int n_rs, const int64_t* __restrict__ row_off, long long* __restrict__ pair_base, int* __restrict__ bad) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = (int64_t)pl[p] * n_rs + pr[p];
    long long old = atomicExch((unsigned long long*)&pair_base[cell], (unsigned long long)row_off[p]);
    if (old != -1ll) *bad = 2;  // the same (ref, exp) pair listed twice
  }
}

// One packed record per sorted index entry: everything the streaming kernel needs about a candidate in ONE 16-byte
// (32-bit keys) or 32-byte (64-bit keys) load.  pmax_prev = running max of the stops of the entries BEFORE this one, so
// the backward walk knows whether to continue without touching the next record.
template <typename K>
struct alignas(16) RefEntry {
  K stop;
  K pmax_prev;
  uint32_t rank;   // rank of the canonical (dup_of) row inside its ref sample = output row offset
  uint32_t meta;   // ref sample << 2 | strand
};

template <typename K>
__global__ void k_index_entries(const uint32_t* __restrict__ rows, int64_t n, const K* __restrict__ stop, const K* __restrict__ pmax, const int32_t* __restrict__ sample,
                                const int8_t* __restrict__ strand, const int32_t* __restrict__ dup_of, const uint32_t* __restrict__ rank_of_row, int n_samples, int* __restrict__ bad,
                                RefEntry<K>* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t row = rows[k];
    uint32_t canon = dup_of ? (uint32_t)dup_of[row] : row;
    RefEntry<K> e;
    e.stop = stop[k];
    e.pmax_prev = k > 0 ? pmax[k - 1] : (K)0;
    e.rank = rank_of_row ? rank_of_row[canon] : canon;     // single ref sample: the rank inside the sample is the row itself
    int sm = sample ? sample[row] : 0;
    if (sm < 0 || sm >= n_samples) { *bad = 1; sm = 0; }
    e.meta = ((uint32_t)sm << 2) | (uint32_t)(strand ? strand[row] : (int8_t)0);
    out[k] = e;
  }
}

// bucket table over the linear coordinate: tbl[b] = lower_bound(r_start, b << shift).  A query then binary-searches only
// inside [tbl[b], tbl[b+1]] (usually 0-2 entries) instead of doing ~log2(n_ref) dependent L1/L2 loads.
template <typename K>
__global__ void k_bucket_table(const K* __restrict__ r_start, int64_t n_ref, int shift, int64_t n_buckets, uint32_t* __restrict__ tbl) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)b << shift;
    K kx = (x > (uint64_t)(~(K)0)) ? ~(K)0 : (K)x;
    tbl[b] = (uint32_t)lower_bound_dev<K>(r_start, n_ref, kx);
  }
}

// ---- small pre-sorted reference: the whole index in one launch ------------------------------------------------------
// MAP onto a COVER result or an annotation: one reference sample of at most MIX_MAX rows, promised to be in (chrom, start)
// order.  Linear keys (with the order and span checks of k_index_stops), stops, the running maximum of the stops, the
// packed entries and the bucket table come from ONE cluster of MIX_CL blocks instead of six launch-bound kernels.  Every
// block owns a contiguous range of the rows; the running maximum crosses block ranges through one value per block in
// distributed shared memory.
constexpr int MIX_T = 1024;
constexpr int MIX_CL = 8;
constexpr int64_t MIX_MAX = 1 << 17;

template <typename K>
__device__ __forceinline__ K mix_incl_max(K v, K* total, K* wt /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  K x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { K y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o && y > x) x = y; }
  __syncthreads();
  if (lane == 31) wt[w] = x;
  __syncthreads();
  K t = wt[lane];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { K y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o && y > t) t = y; }
  *total = __shfl_sync(0xffffffffu, t, 31);
  const K below = __shfl_sync(0xffffffffu, t, w > 0 ? w - 1 : 0);
  return (w > 0 && below > x) ? below : x;
}

template <typename K>
__global__ void __cluster_dims__(MIX_CL, 1, 1) __launch_bounds__(MIX_T, 1)
k_map_index_small(const int32_t* __restrict__ chrom, const void* __restrict__ start, const void* __restrict__ stop, int bits, int n, LinMap lm,
                  const int32_t* __restrict__ sample, const int8_t* __restrict__ strand, const int32_t* __restrict__ dup_of, int n_samples,
                  int tbl_shift, int64_t n_buckets, K* __restrict__ keys, uint32_t* __restrict__ rows, uint32_t* __restrict__ rows_by_sample,
                  K* __restrict__ stops, K* __restrict__ pmax, uint32_t* __restrict__ tbl, RefEntry<K>* __restrict__ entries, int* __restrict__ bad) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();
  __shared__ K wt[32];
  __shared__ K x_max;
  const int tid = threadIdx.x;
  const int L = (n + MIX_CL - 1) / MIX_CL;
  const int lo = min(n, cr * L), hi = min(n, lo + L);
  K carry = 0;
  for (int base = lo; base < hi; base += MIX_T) {
    const int i = base + tid;
    K st = 0;
    if (i < hi) {
      int c = chrom[i];
      int64_t s = ld_coord(start, bits, i), e = ld_coord(stop, bits, i);
      if (c < 0 || c >= lm.n_chrom || s < 0 || e < 0 || e > lm.span[c]) { *bad = 1; c = 0; s = 0; e = 0; }
      const K key = (K)(lm.coff[c] + (uint64_t)s);
      if (i > 0) {                                           // the caller promised (chrom, start) order
        const int cp = chrom[i - 1];
        const int64_t sp = ld_coord(start, bits, i - 1);
        if (cp >= 0 && cp < lm.n_chrom && sp >= 0 && (K)(lm.coff[cp] + (uint64_t)sp) > key) *bad = 1;
      }
      st = (K)(lm.coff[c] + (uint64_t)e);
      keys[i] = key; rows[i] = (uint32_t)i; rows_by_sample[i] = (uint32_t)i; stops[i] = st;
    }
    K tot;
    const K incl = mix_incl_max<K>(st, &tot, wt);
    if (i < hi) pmax[i] = incl > carry ? incl : carry;       // running maximum inside this block's range
    if (tot > carry) carry = tot;
  }
  if (tid == 0) x_max = carry;
  cluster.sync();                                            // keys and block maxima are visible cluster-wide
  K before = 0;
  for (int c = 0; c < cr; ++c) { const K v = *cluster.map_shared_rank(&x_max, c); if (v > before) before = v; }
  for (int base = lo; base < hi; base += MIX_T) {
    const int i = base + tid;
    K pm = 0, pp = 0;
    if (i < hi) { pm = pmax[i]; pp = i > lo ? pmax[i - 1] : (K)0; }
    __syncthreads();                                         // all local values of the round are read before any is replaced
    if (i < hi) {
      if (before > pm) pm = before;
      if (before > pp) pp = before;
      pmax[i] = pm;
      const uint32_t canon = dup_of ? (uint32_t)dup_of[i] : (uint32_t)i;
      RefEntry<K> en;
      en.stop = stops[i];
      en.pmax_prev = i > 0 ? pp : (K)0;
      en.rank = canon;
      int sm = sample ? sample[i] : 0;
      if (sm < 0 || sm >= n_samples) { *bad = 1; sm = 0; }
      en.meta = ((uint32_t)sm << 2) | (uint32_t)(strand ? strand[i] : (int8_t)0);
      entries[i] = en;
    }
  }
  for (int64_t b = (int64_t)cr * MIX_T + tid; b <= n_buckets; b += (int64_t)MIX_CL * MIX_T) {
    const uint64_t x = (uint64_t)b << tbl_shift;
    const K kx = (x > (uint64_t)(~(K)0)) ? ~(K)0 : (K)x;
    tbl[b] = (uint32_t)lower_bound_dev<K>(keys, n, kx);
  }
  cluster.sync();                                            // no block leaves while its maximum may still be read
}

// ---- the streaming kernel ------------------------------------------------------------------------
// Persistent blocks (one per SM).  The bucket table (and the per-chromosome offsets) are staged in shared memory, so a
// query costs one shared-memory lookup, usually no search at all (most buckets hold no reference start), one packed
// RefEntry load per candidate and the accumulating REDs.  Every thread handles four consecutive experiment regions per
// iteration (128-bit column loads; four independent candidate loads in flight).
constexpr int MS_THREADS = 1024;
constexpr int MS_PER = 4;
constexpr int MS_MAX_TBL = 46 * 1024;      // bucket-table entries kept in shared memory (184 KB)
constexpr int MS_MAX_CHROM = 1024;         // chromosomes whose offset/span are kept in shared memory

template <typename K>
struct MapStreamArgs {
  const int32_t* e_chrom; const void* e_start; const void* e_stop; int e_bits;
  const int8_t* e_strand; const int32_t* e_sample; int64_t n_exp; int n_exp_samples;
  LinMap lm; const K* r_start; const RefEntry<K>* entries; int64_t n_ref;
  const long long* pair_base; int* bad;
  const uint32_t* tbl; int tbl_shift; int n_tbl;   // n_tbl = entries of tbl (buckets + 2)
  int vec_ok;                                        // column pointers are 16-byte aligned
  uint8_t* hit; int exact;                           // DIFFERENCE: per reference row "some paired experiment region overlaps it"
  const int64_t* bag_off; int32_t* bag_cursor; int32_t* bag_rows;   // BAG / BAGD: matched experiment rows of every output row
};

template <typename K>
__device__ __forceinline__ void map_hit(const MapStreamArgs<K>& a, const MapAggPlan& plan, const ExpCols& ec, const MapAcc& acc, const RefEntry<K>& e, int b, int st, int64_t i) {
  if (!strand_ok((int)(e.meta & 3u), st)) return;
  long long base = a.pair_base[(int64_t)(e.meta >> 2) * a.n_exp_samples + b];
  if (base < 0) return;
  int64_t row = base + e.rank;
  atomicAdd(&acc.count[row], 1);
  for (int q = 0; q < plan.n_cols; ++q) {
    if (ec.ok[q] && !ec.ok[q][i]) continue;
    double v = ec.v[q][i];
    atomicAdd(&acc.nn[q][row], 1);
    if (plan.need_sum[q]) atomicAdd(&acc.sum[q][row], v);
    if (plan.need_min[q]) atomicMin(&acc.mn[q][row], dbl_key_min(v));
    if (plan.need_max[q]) atomicMax(&acc.mx[q][row], dbl_key_max(v));
  }
}

// MODE 0: accumulate.  MODE 1 (DIFFERENCE): mark the reference row instead (GenometricDifference.scala:58-69; `exact`: identical
// coordinates).  MODE 2 (BAG / BAGD): append the experiment row to the output row's list (a second pass, after the counts
// have been scanned into list offsets); the strings are built on the host from these lists.
template <typename K, int MODE>
__global__ void __launch_bounds__(MS_THREADS, 1) k_map_stream(MapStreamArgs<K> a, MapAggPlan plan, ExpCols ec, MapAcc acc) {
  extern __shared__ uint32_t s_tbl[];                   // [n_tbl] bucket table, then chromosome offsets / spans
  const bool chrom_sh = a.lm.n_chrom <= MS_MAX_CHROM;
  uint64_t* s_coff = (uint64_t*)(s_tbl + ((a.n_tbl + 3) & ~3));
  int64_t* s_span = (int64_t*)(s_coff + (chrom_sh ? a.lm.n_chrom : 0));
  for (int k = threadIdx.x; k < a.n_tbl; k += MS_THREADS) s_tbl[k] = a.tbl[k];
  if (chrom_sh) for (int k = threadIdx.x; k < a.lm.n_chrom; k += MS_THREADS) { s_coff[k] = a.lm.coff[k]; s_span[k] = a.lm.span[k]; }
  __syncthreads();
  const uint64_t* coff = chrom_sh ? s_coff : a.lm.coff;
  const int64_t* span = chrom_sh ? s_span : a.lm.span;

  for (int64_t i0 = ((int64_t)blockIdx.x * MS_THREADS + threadIdx.x) * MS_PER; i0 < a.n_exp; i0 += (int64_t)gridDim.x * MS_THREADS * MS_PER) {
    int c[MS_PER], smp[MS_PER], st[MS_PER];
    int64_t s[MS_PER], e[MS_PER];
    const bool full = i0 + MS_PER <= a.n_exp;
    if (full && a.vec_ok && a.e_bits == 32) {
      int4 vc = *(const int4*)(a.e_chrom + i0), vs = *(const int4*)((const int32_t*)a.e_start + i0), ve = *(const int4*)((const int32_t*)a.e_stop + i0);
      c[0] = vc.x; c[1] = vc.y; c[2] = vc.z; c[3] = vc.w;
      s[0] = vs.x; s[1] = vs.y; s[2] = vs.z; s[3] = vs.w;
      e[0] = ve.x; e[1] = ve.y; e[2] = ve.z; e[3] = ve.w;
      if (a.e_sample) { int4 vb = *(const int4*)(a.e_sample + i0); smp[0] = vb.x; smp[1] = vb.y; smp[2] = vb.z; smp[3] = vb.w; }
      else { smp[0] = smp[1] = smp[2] = smp[3] = 0; }
      if (a.e_strand) { uint32_t w = *(const uint32_t*)(a.e_strand + i0); st[0] = (int)(int8_t)(w & 0xff); st[1] = (int)(int8_t)((w >> 8) & 0xff); st[2] = (int)(int8_t)((w >> 16) & 0xff); st[3] = (int)(int8_t)(w >> 24); }
      else { st[0] = st[1] = st[2] = st[3] = 0; }
    } else {
#pragma unroll
      for (int j = 0; j < MS_PER; ++j) {
        int64_t i = i0 + j;
        bool ok = i < a.n_exp;
        c[j] = ok ? a.e_chrom[i] : -2;                  // -2: padding lane, skipped silently
        s[j] = ok ? ld_coord(a.e_start, a.e_bits, i) : 0;
        e[j] = ok ? ld_coord(a.e_stop, a.e_bits, i) : 0;
        smp[j] = ok && a.e_sample ? a.e_sample[i] : 0;
        st[j] = ok && a.e_strand ? (int)a.e_strand[i] : 0;
      }
    }
    // ---- candidate range of every region: hi = number of reference entries with start < region.stop
    K ls[MS_PER], lee[MS_PER];
    int hi[MS_PER];
#pragma unroll
    for (int j = 0; j < MS_PER; ++j) {
      hi[j] = 0; ls[j] = 0; lee[j] = 0;
      if (c[j] == -2) continue;
      if (c[j] < 0 || c[j] >= a.lm.n_chrom) { *a.bad = 1; continue; }
      int64_t sp = span[c[j]];
      if (s[j] >= sp || e[j] <= 0) continue;            // no reference region can satisfy r.start < e.stop && e.start < r.stop
      if (s[j] < 0) { *a.bad = 1; continue; }
      const bool clamped = e[j] > sp + 1;
      if (clamped) e[j] = sp + 1;
      if (smp[j] < 0 || smp[j] >= a.n_exp_samples) { *a.bad = 1; continue; }
      uint64_t co = coff[c[j]];
      ls[j] = (K)(co + (uint64_t)s[j]);
      K le = (K)(co + (uint64_t)e[j]);
      lee[j] = clamped ? (K)0 : le;                     // a clamped stop equals no reference stop
      uint64_t b = (uint64_t)le >> a.tbl_shift;
      int lo2 = (int)s_tbl[b], hi2 = (int)s_tbl[b + 1];
      while (lo2 < hi2) { int mid = (lo2 + hi2) >> 1; if (a.r_start[mid] < le) lo2 = mid + 1; else hi2 = mid; }
      hi[j] = lo2;
    }
    // ---- first candidate of each region: independent loads, all in flight together
    RefEntry<K> ent[MS_PER];
#pragma unroll
    for (int j = 0; j < MS_PER; ++j) if (hi[j] > 0) ent[j] = a.entries[hi[j] - 1];
#pragma unroll
    for (int j = 0; j < MS_PER; ++j) {
      if (hi[j] <= 0) continue;
      int k = hi[j] - 1;
      RefEntry<K> en = ent[j];
      while (true) {
        if (en.stop > ls[j]) {
          if (MODE == 0) map_hit<K>(a, plan, ec, acc, en, smp[j], st[j], i0 + j);
          else if (MODE == 2) {
            if (strand_ok((int)(en.meta & 3u), st[j])) {
              long long base = a.pair_base[(int64_t)(en.meta >> 2) * a.n_exp_samples + smp[j]];
              if (base >= 0) {
                int64_t row = base + en.rank;
                int pos = atomicAdd(&a.bag_cursor[row], 1);
                a.bag_rows[a.bag_off[row] + pos] = (int32_t)(i0 + j);
              }
            }
          }
          else if (strand_ok((int)(en.meta & 3u), st[j]) && a.pair_base[(int64_t)(en.meta >> 2) * a.n_exp_samples + smp[j]] >= 0 &&
                   (!a.exact || (en.stop == lee[j] && a.r_start[k] == ls[j])))
            a.hit[en.rank] = 1;
        }
        if (k == 0 || en.pmax_prev <= ls[j]) break;
        --k;
        en = a.entries[k];
      }
    }
  }
}

// ---- COUNT onto a small reference, experiment rows grouped by sample --------------------------------------------------
// The shape of `M = MAP() C E` with C a COVER result (~10^4 regions) and E a few thousand sample files of ~10^5 regions:
// the streaming kernel above is then bound by its scattered accesses — one 16-byte reference entry load and one RED into a
// table of n_ref x n_samples counters per experiment region (L2 sector throughput, not DRAM).  Here a block takes one
// experiment sample at a time and keeps EVERYTHING it touches at random in shared memory: the reference as 16-byte records
// {start, stop, running max of the earlier stops, rank << 2 | strand}, the bucket table over the linear coordinate, and the
// sample's n_ref counters (16 bits each: a count cannot exceed the sample's rows, < 65 536).  Global memory sees only the
// coalesced stream of the sample's regions and the coalesced write of its finished row block (no REDs, no zero-fill).
// Conditions (else k_map_stream): one reference sample, 32-bit keys, no value aggregates / BAG, rows grouped by sample
// (sample_offsets), every sample below 65 536 rows, and the reference small enough for shared memory.
constexpr int64_t MG_SMEM = 224 * 1024;

struct MapGroupedArgs {
  const int32_t* e_chrom; const void* e_start; const void* e_stop; int e_bits;
  const uint8_t* e_chrom8; const uint16_t* e_len16;   // NARROW: the ABI's narrow encodings, read as they are (7 bytes per row)
  const int8_t* e_strand; const int32_t* e_sample; const int64_t* sample_off; int n_exp_samples;
  LinMap lm; const uint32_t* r_start; const RefEntry<uint32_t>* entries; int n_ref;
  const long long* pair_base; int* bad;
  const uint32_t* tbl; int tbl_shift; int n_tbl;
  int* count;
  int compact;   // every entry's rank is its index and the reference has no strand column: no rank|strand array in shared memory
};


// B32: 32-bit experiment coordinates (the usual case): the whole row pipeline runs on 32-bit registers, the chromosome's
// {offset, span} pair is one 8-byte shared-memory load, and the rows of the NEXT iteration are already in flight while
// the current ones are matched.
template <bool B32, bool NARROW>
__global__ void __launch_bounds__(MS_THREADS, 1) k_map_count_grouped(MapGroupedArgs a) {
  extern __shared__ __align__(16) uint32_t g_dyn[];
  // the reference as a structure of arrays: the in-bucket scan walks consecutive 4-byte starts
  uint32_t* s_start = g_dyn;                                      // [n_ref]
  uint32_t* s_stop = s_start + a.n_ref;                           // [n_ref]
  uint32_t* s_pmax = s_stop + a.n_ref;                            // [n_ref] running max of the EARLIER stops
  uint32_t* s_rs = s_pmax + a.n_ref;                              // [n_ref] rank << 2 | strand, absent when a.compact
  uint32_t* s_cnt = s_rs + (a.compact ? 0 : a.n_ref);             // [(n_ref + 1) / 2] two 16-bit counters per word
  const int n_cw = (a.n_ref + 1) >> 1;
  uint16_t* s_tbl = (uint16_t*)(s_cnt + n_cw);                    // [n_tbl] 16-bit entries (n_ref < 65 536): twice the buckets per byte
  uint2* s_cs = (uint2*)(g_dyn + (((a.compact ? 3 : 4) * (size_t)a.n_ref + n_cw + (((size_t)a.n_tbl + 1) >> 1) + 1) & ~(size_t)1));   // [n_chrom] {linear offset, span}
  for (int k = threadIdx.x; k < a.n_ref; k += MS_THREADS) {
    const RefEntry<uint32_t> e = a.entries[k];
    s_start[k] = a.r_start[k]; s_stop[k] = e.stop; s_pmax[k] = e.pmax_prev;
    if (!a.compact) s_rs[k] = (e.rank << 2) | (e.meta & 3u);
  }
  for (int k = threadIdx.x; k < a.n_tbl; k += MS_THREADS) s_tbl[k] = (uint16_t)a.tbl[k];
  for (int k = threadIdx.x; k < a.lm.n_chrom; k += MS_THREADS) s_cs[k] = make_uint2((uint32_t)a.lm.coff[k], (uint32_t)a.lm.span[k]);
  const int32_t* e_start32 = (const int32_t*)a.e_start; const int32_t* e_stop32 = (const int32_t*)a.e_stop;

  // one experiment row against the staged reference; coordinates already validated against 32-bit overflow by the caller
  auto match = [&](int c, int64_t s, int64_t e, int st) {
    if (c < 0 || c >= a.lm.n_chrom) { *a.bad = 1; return; }
    const uint2 cs = s_cs[c];
    const int64_t sp = (int64_t)cs.y;
    if (s >= sp || e <= 0) return;                                // no reference region can satisfy r.start < e.stop && e.start < r.stop
    if (s < 0) { *a.bad = 1; return; }
    if (e > sp + 1) e = sp + 1;
    const uint32_t ls = cs.x + (uint32_t)s, le = cs.x + (uint32_t)e;
    const uint32_t b = le >> a.tbl_shift;
    int k = (int)s_tbl[b];
    const int hi2 = (int)s_tbl[b + 1];
    if (hi2 - k > 4) { int lo2 = k, h2 = hi2; while (lo2 < h2) { const int mid = (lo2 + h2) >> 1; if (s_start[mid] < le) lo2 = mid + 1; else h2 = mid; } k = lo2; }
    else while (k < hi2 && s_start[k] < le) ++k;                  // buckets usually hold no or one reference start
    --k;                                                          // last reference region with start < row.stop
    while (k >= 0) {
      if (s_stop[k] > ls) {
        const uint32_t rs = a.compact ? ((uint32_t)k << 2) : s_rs[k];
        if (strand_ok((int)(rs & 3u), st)) {
          const uint32_t rank = rs >> 2;
          atomicAdd(&s_cnt[rank >> 1], 1u << ((rank & 1u) << 4));
        }
      }
      if (s_pmax[k] <= ls) break;
      --k;
    }
  };

  for (int smp = blockIdx.x; smp < a.n_exp_samples; smp += gridDim.x) {
    const long long base = a.pair_base[smp];                      // the only reference sample is sample 0
    if (base < 0) continue;                                       // this experiment sample is in no pair (block-uniform)
    for (int k = threadIdx.x; k < n_cw; k += MS_THREADS) s_cnt[k] = 0;
    __syncthreads();                                              // also orders the staging above
    const int64_t lo = a.sample_off[smp], hi = a.sample_off[smp + 1];
    if (B32) {
      int c[MS_PER], s[MS_PER], e[MS_PER], sm[MS_PER];            // the rows of the iteration being matched
      int nc[MS_PER], ns[MS_PER], ne[MS_PER], nsm[MS_PER];        // ... and of the next one, in flight
      auto fetch = [&](int64_t i0, int* fc, int* fs, int* fe, int* fsm) {
#pragma unroll
        for (int j = 0; j < MS_PER; ++j) {                        // coalesced 32-bit loads; nothing here waits for them
          const int64_t i = i0 + (int64_t)j * MS_THREADS;
          const bool ok = i < hi;
          fc[j] = ok ? (NARROW ? (int)a.e_chrom8[i] : a.e_chrom[i]) : -2;   // -2: padding lane
          fs[j] = ok ? e_start32[i] : 0;
          fe[j] = ok ? (NARROW ? fs[j] + (int)a.e_len16[i] : e_stop32[i]) : 0;
          fsm[j] = ok && a.e_sample ? a.e_sample[i] : smp;
        }
      };
      int64_t i0 = lo + threadIdx.x;
      if (i0 < hi) fetch(i0, nc, ns, ne, nsm);
      for (; i0 < hi; i0 += (int64_t)MS_THREADS * MS_PER) {
#pragma unroll
        for (int j = 0; j < MS_PER; ++j) { c[j] = nc[j]; s[j] = ns[j]; e[j] = ne[j]; sm[j] = nsm[j]; }
        const int64_t i1 = i0 + (int64_t)MS_THREADS * MS_PER;
        if (i1 < hi) fetch(i1, nc, ns, ne, nsm);
#pragma unroll
        for (int j = 0; j < MS_PER; ++j) {
          if (c[j] == -2) continue;
          if (sm[j] != smp) { *a.bad = 3; continue; }             // the rows are not grouped as sample_offsets promised
          const int st = a.e_strand ? (int)a.e_strand[i0 + (int64_t)j * MS_THREADS] : 0;
          match(c[j], (int64_t)s[j], (int64_t)e[j], st);
        }
      }
    } else {
      for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += (int64_t)MS_THREADS * MS_PER) {
#pragma unroll
        for (int j = 0; j < MS_PER; ++j) {
          const int64_t i = i0 + (int64_t)j * MS_THREADS;
          if (i >= hi) continue;
          if (a.e_sample && a.e_sample[i] != smp) { *a.bad = 3; continue; }
          match(a.e_chrom[i], ((const int64_t*)a.e_start)[i], ((const int64_t*)a.e_stop)[i], a.e_strand ? (int)a.e_strand[i] : 0);
        }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.n_ref; k += MS_THREADS) a.count[base + k] = (int)((s_cnt[k >> 1] >> ((k & 1) << 4)) & 0xffffu);
    __syncthreads();
  }
}

// The production shape of `M = MAP() C E` on a resident dataset: experiment rows in the ABI's narrow encodings (uint8 chromosome,
// int32 start, uint16 length: 7 bytes per row, the sample of a row known from sample_offsets), reference = a COVER result or a
// sorted annotation without strand column or duplicates (every rank is its index).  Same staging as k_map_count_grouped, but the
// row pipeline is pure 32-bit, a thread reads FOUR consecutive rows with three vector loads (16 + 8 + 4 bytes, the next group in
// flight while this one is matched), stop and running maximum of a reference region are one 8-byte shared-memory load, and no
// strand / sample / padding-lane bookkeeping is left: ~4.5 -> ~2 warp instructions per row (ncu, profiles/).
__global__ void __launch_bounds__(MS_THREADS, 1) k_map_count_lean(MapGroupedArgs a, int64_t n_exp_rows) {
  extern __shared__ __align__(16) uint32_t g_dyn[];
  const int n_ref = a.n_ref, n_ref2 = (n_ref + 2) & ~1;
  uint2* s_sp = (uint2*)g_dyn;                                    // [n_ref] {stop, running max of the EARLIER stops}
  uint32_t* s_start = g_dyn + 2 * (size_t)n_ref;                  // [n_ref] + a sentinel that ends every forward scan
  uint32_t* s_cnt = s_start + n_ref2;                             // [(n_ref + 1) / 2] two 16-bit counters per word
  const int n_cw = (n_ref + 1) >> 1;
  uint16_t* s_tbl = (uint16_t*)(s_cnt + n_cw);                    // [n_tbl]
  uint2* s_cs = (uint2*)(g_dyn + ((3 * (size_t)n_ref + 2 + n_cw + (((size_t)a.n_tbl + 1) >> 1) + 1) & ~(size_t)1));   // [n_chrom] {linear offset, span}
  for (int k = threadIdx.x; k < n_ref; k += MS_THREADS) {
    const RefEntry<uint32_t> e = a.entries[k];
    s_start[k] = a.r_start[k];
    s_sp[k] = make_uint2(e.stop, e.pmax_prev);
  }
  if (threadIdx.x == 0) s_start[n_ref] = 0xffffffffu;
  for (int k = threadIdx.x; k < a.n_tbl; k += MS_THREADS) s_tbl[k] = (uint16_t)a.tbl[k];
  for (int k = threadIdx.x; k < a.lm.n_chrom; k += MS_THREADS) s_cs[k] = make_uint2((uint32_t)a.lm.coff[k], (uint32_t)a.lm.span[k]);
  const int4* st4 = (const int4*)a.e_start;
  const uint2* ln4 = (const uint2*)a.e_len16;
  const uint32_t* ch4 = (const uint32_t*)a.e_chrom8;
  const int32_t* st1 = (const int32_t*)a.e_start;
  const uint32_t n_chrom = (uint32_t)a.lm.n_chrom;
  const int shift = a.tbl_shift;

  auto match = [&](uint32_t c, int32_t s, uint32_t len) {
    if (c >= n_chrom) { *a.bad = 1; return; }
    const uint2 cs = s_cs[c];
    if ((uint32_t)s >= cs.y) { if (s < 0) *a.bad = 1; return; }   // beyond every reference region of the chromosome (or invalid)
    uint32_t e = (uint32_t)s + len;
    if (e > cs.y + 1u) e = cs.y + 1u;
    const uint32_t ls = cs.x + (uint32_t)s, le = cs.x + e;
    const uint32_t b = le >> shift;
    int k = (int)s_tbl[b];                                        // first reference start at or after the bucket's first position
    // forward to the first start >= row.stop: buckets usually hold no or one reference start, and the sentinel ends the scan
    // without a look at the next bucket's entry; a crowded bucket is finished by bisection
    if (s_start[k] < le) {
      ++k;
      if (s_start[k] < le) {
        ++k;
        if (s_start[k] < le) {
          int lo2 = k + 1, h2 = (int)s_tbl[b + 1];
          while (lo2 < h2) { const int mid = (lo2 + h2) >> 1; if (s_start[mid] < le) lo2 = mid + 1; else h2 = mid; }
          k = lo2;
        }
      }
    }
    --k;                                                          // last reference region with start < row.stop
    while (k >= 0) {
      const uint2 sp = s_sp[k];
      if (sp.x > ls) atomicAdd(&s_cnt[k >> 1], 1u << ((k & 1) << 4));
      if (sp.y <= ls) break;
      --k;
    }
  };

  for (int smp = blockIdx.x; smp < a.n_exp_samples; smp += gridDim.x) {
    const long long base = a.pair_base[smp];                      // the only reference sample is sample 0
    if (base < 0) continue;                                       // this experiment sample is in no pair (block-uniform)
    for (int k = threadIdx.x; k < n_cw; k += MS_THREADS) s_cnt[k] = 0;
    __syncthreads();                                              // also orders the staging above
    const int64_t lo = a.sample_off[smp], hi = a.sample_off[smp + 1];
    // groups of four consecutive rows, aligned on the arrays: group g holds rows 4g .. 4g + 3
    const int64_t g_end = (hi + 3) >> 2;
    int64_t g = (lo >> 2) + threadIdx.x;
    int4 ns = make_int4(0, 0, 0, 0); uint2 nl = make_uint2(0u, 0u); uint32_t nc = 0u;
    auto fetch = [&](int64_t gg) {
      if (4 * gg + 4 <= n_exp_rows) { ns = st4[gg]; nl = ln4[gg]; nc = ch4[gg]; }
      else {                                                      // the last, partial group of the whole dataset
        int sv[4] = {0, 0, 0, 0}; uint32_t lv[4] = {0u, 0u, 0u, 0u}, cv[4] = {0u, 0u, 0u, 0u};
        for (int q = 0; q < 4; ++q) if (4 * gg + q < n_exp_rows) { sv[q] = st1[4 * gg + q]; lv[q] = a.e_len16[4 * gg + q]; cv[q] = a.e_chrom8[4 * gg + q]; }
        ns = make_int4(sv[0], sv[1], sv[2], sv[3]);
        nl = make_uint2(lv[0] | (lv[1] << 16), lv[2] | (lv[3] << 16));
        nc = cv[0] | (cv[1] << 8) | (cv[2] << 16) | (cv[3] << 24);
      }
    };
    if (g < g_end) fetch(g);
    for (; g < g_end; g += MS_THREADS) {
      const int4 vs = ns; const uint2 vl = nl; const uint32_t vc = nc;
      if (g + MS_THREADS < g_end) fetch(g + MS_THREADS);
      const int64_t r0 = 4 * g;
      if (r0 >= lo && r0 + 4 <= hi) {                             // (all but the first and last group of a sample)
        match(vc & 0xffu, vs.x, vl.x & 0xffffu);
        match((vc >> 8) & 0xffu, vs.y, vl.x >> 16);
        match((vc >> 16) & 0xffu, vs.z, vl.y & 0xffffu);
        match(vc >> 24, vs.w, vl.y >> 16);
      } else {
        if (r0 >= lo && r0 < hi) match(vc & 0xffu, vs.x, vl.x & 0xffffu);
        if (r0 + 1 >= lo && r0 + 1 < hi) match((vc >> 8) & 0xffu, vs.y, vl.x >> 16);
        if (r0 + 2 >= lo && r0 + 2 < hi) match((vc >> 16) & 0xffu, vs.z, vl.y & 0xffffu);
        if (r0 + 3 >= lo && r0 + 3 < hi) match(vc >> 24, vs.w, vl.y >> 16);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_ref; k += MS_THREADS) a.count[base + k] = (int)((s_cnt[k >> 1] >> ((k & 1) << 4)) & 0xffffu);
    __syncthreads();
  }
}

// rows of non-canonical duplicates are not emitted by the reference (MapKey collapse): flag them with count = -1
__global__ void k_mark_dups(const int32_t* __restrict__ pl, const int64_t* __restrict__ row_off, int64_t n_pairs, const int64_t* __restrict__ sample_off,
                            const uint32_t* __restrict__ rows_by_sample, const int32_t* __restrict__ dup_of, int* __restrict__ count) {
  // one block per pair (grid-stride), threads over the ref rows of the pair's sample
  for (int64_t p = blockIdx.x; p < n_pairs; p += gridDim.x) {
    int a = pl[p];
    int64_t s0 = sample_off[a], s1 = sample_off[a + 1];
    for (int64_t k = s0 + threadIdx.x; k < s1; k += blockDim.x) {
      uint32_t row = rows_by_sample[k];
      if ((uint32_t)dup_of[row] != row) count[row_off[p] + (k - s0)] = -1;
    }
  }
}

struct MapOut {
  double* value[16];
  uint8_t* valid[16];
};

__global__ void __launch_bounds__(256) k_map_finalize(int64_t n_rows, MapAggPlan plan, MapAcc acc, MapOut out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    int cnt = acc.count[r];
    for (int a = 0; a < plan.n_aggs; ++a) {
      int kind = plan.kind[a], q = plan.slot[a];
      double v = 0.0;
      uint8_t ok = 0;
      if (kind == GMQL_AGG_COUNT) { v = (double)(cnt < 0 ? 0 : cnt); ok = 1; }  // funOut = GDouble(count), DefaultRegionsToRegionFactory.scala:62
      else if (kind == GMQL_AGG_BAG || kind == GMQL_AGG_BAGD) { v = 0.0; ok = 0; }  // host-built string (GMQL_COL_BAG_*)
      else {
        int nn = acc.nn[q][r];
        if (nn > 0) {
          ok = 1;
          if (kind == GMQL_AGG_SUM) v = acc.sum[q][r];                       // :69-82
          else if (kind == GMQL_AGG_AVG) v = acc.sum[q][r] / (double)nn;     // :112-126
          else if (kind == GMQL_AGG_MIN) v = dbl_from_key(acc.mn[q][r]);     // :84-96
          else if (kind == GMQL_AGG_MAX) v = dbl_from_key(acc.mx[q][r]);     // :98-110
          else if (kind == GMQL_AGG_MEDIAN) {                                // :33-56 pairwise fold => min; (m+m)/2 once two values met
            double m = dbl_from_key(acc.mn[q][r]);
            v = nn >= 2 ? (m + m) / 2 : m;
          }
        }
      }
      out.value[a][r] = v;
      out.valid[a][r] = ok;
    }
  }
}

static void plan_aggs(const gmql_agg* aggs, int n_aggs, int n_cols_avail, MapAggPlan* plan, std::vector<int>* need_cols) {
  GMQL_REQUIRE(n_aggs >= 0 && n_aggs <= 16, GMQL_ERR_UNSUPPORTED, "at most 16 aggregates per call");
  GMQL_REQUIRE(n_aggs == 0 || aggs != nullptr, GMQL_ERR_INVALID, "null aggregate list");
  *plan = MapAggPlan();
  plan->n_aggs = n_aggs;
  for (int a = 0; a < n_aggs; ++a) {
    int kind = aggs[a].kind;
    plan->kind[a] = kind;
    plan->slot[a] = -1;
    if (kind == GMQL_AGG_COUNT) continue;
    if (kind == GMQL_AGG_BAG || kind == GMQL_AGG_BAGD) { plan->need_bag = 1; continue; }   // strings are built on the host from the match lists
    GMQL_REQUIRE(kind >= GMQL_AGG_SUM && kind <= GMQL_AGG_MEDIAN, GMQL_ERR_INVALID,
                 "unknown aggregate kind (user closures have no function_identifier and cannot run on the device)");
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
    if (kind == GMQL_AGG_SUM || kind == GMQL_AGG_AVG) plan->need_sum[q] = 1;
    if (kind == GMQL_AGG_MIN || kind == GMQL_AGG_MEDIAN) plan->need_min[q] = 1;
    if (kind == GMQL_AGG_MAX) plan->need_max[q] = 1;
  }
}

static int64_t read_back_i64(gmql_ctx* ctx, const int64_t* dev) {
  int64_t v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, dev, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  return v;
}

template <typename K>
static void run_map(gmql_ctx* ctx, const DevRegions& ref, DevRegions& exp, const gmql_pair* pairs, int64_t n_pairs,
                    const MapAggPlan& plan, const SpanTable& spans, gmql_result* res) {
  const int64_t n_ref = ref.n, n_exp = exp.n;
  const int n_ls = ref.n_samples, n_rs = exp.n_samples;

  // pairs -> device (through pinned scratch: truly asynchronous); the single-ref-sample path needs them only for dup marking
  int32_t* h_pl = (int32_t*)ctx_pinned(ctx, 4 * (size_t)n_pairs);
  int32_t* h_pr = (int32_t*)ctx_pinned(ctx, 4 * (size_t)n_pairs);
  for (int64_t p = 0; p < n_pairs; ++p) { h_pl[p] = pairs[p].left_sample; h_pr[p] = pairs[p].right_sample; }
  DevBuf<int32_t> d_pl(ctx, n_pairs), d_pr(ctx, n_pairs);
  if (n_pairs && (n_ls != 1 || ref.dup_of)) {
    CUDA_TRY(cudaMemcpyAsync(d_pl.p, h_pl, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_pr.p, h_pr, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
  }
  cudaEventRecord(ctx->ev_h1, ctx->stream);

  DevBuf<int> bad(ctx, 1);
  bad.zero();

  // reference rows grouped by sample (stable: ascending row index inside a sample); rows per pair, row offsets, dense
  // (ref sample x exp sample) -> first row table
  DevBuf<uint32_t> sk_a, sk_b, sr_a, sr_b, rank_of_row;
  DevBuf<unsigned int> sample_cnt;
  DevBuf<int64_t> sample_off(ctx, (int64_t)n_ls + 1);
  DevBuf<int64_t> row_off(ctx, n_pairs + 1);
  DevBuf<long long> pair_base(ctx, (int64_t)n_ls * n_rs);
  uint32_t* keys_sorted = nullptr; uint32_t* rows_by_sample = nullptr;
  int64_t n_rows = 0;
  int h_bad = 0;
  const bool single = n_ls == 1;
  // small pre-sorted single-sample reference: the whole index comes from one cluster launch (k_map_index_small)
  const char* mix_env = ctx_opt(ctx, "map.index");           // "split": test hook, always the separate index kernels
  const bool small_index = single && ref.sorted_hint && n_ref > 0 && n_ref <= MIX_MAX && !(mix_env && !strcmp(mix_env, "split"));
  if (single) {
    // one reference sample (MAP onto a COVER result, an annotation, ...): every table is known on the host; nothing to
    // sort, no host round trip.  Sample indices of the rows are validated by k_index_entries.
    sr_a.alloc(ctx, n_ref);
    rows_by_sample = sr_a.p;
    if (n_ref && !small_index) LAUNCH(ctx, k_iota_rows, grid_for(ctx, n_ref, 256), 256, 0, sr_a.p, n_ref);
    int64_t* h_so = (int64_t*)ctx_pinned(ctx, 16);
    h_so[0] = 0; h_so[1] = n_ref;
    CUDA_TRY(cudaMemcpyAsync(sample_off.p, h_so, 16, cudaMemcpyHostToDevice, ctx->stream));
    int64_t* h_ro = (int64_t*)ctx_pinned(ctx, 8 * ((size_t)n_pairs + 1));
    long long* h_pb = (long long*)ctx_pinned(ctx, 8 * (size_t)n_rs);
    for (int b = 0; b < n_rs; ++b) h_pb[b] = -1ll;
    for (int64_t p = 0; p < n_pairs; ++p) {
      GMQL_REQUIRE(h_pl[p] == 0 && h_pr[p] >= 0 && h_pr[p] < n_rs, GMQL_ERR_INVALID, "sample index out of range (regions or pair list)");
      GMQL_REQUIRE(h_pb[h_pr[p]] < 0, GMQL_ERR_INVALID, "the same (ref sample, exp sample) pair is listed twice");
      h_ro[p] = p * n_ref;
      h_pb[h_pr[p]] = (long long)(p * n_ref);
    }
    h_ro[n_pairs] = n_pairs * n_ref;
    CUDA_TRY(cudaMemcpyAsync(row_off.p, h_ro, 8 * ((size_t)n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(pair_base.p, h_pb, 8 * (size_t)n_rs, cudaMemcpyHostToDevice, ctx->stream));
    n_rows = n_pairs * n_ref;
  } else {
    sk_a.alloc(ctx, n_ref); sk_b.alloc(ctx, n_ref); sr_a.alloc(ctx, n_ref); sr_b.alloc(ctx, n_ref);
    sample_cnt.alloc(ctx, n_ls);
    sample_cnt.zero();
    if (n_ref) LAUNCH(ctx, k_sample_keys, grid_for(ctx, n_ref, 256), 256, 0, ref.sample, n_ref, n_ls, sk_a.p, sr_a.p, sample_cnt.p, bad.p);
    keys_sorted = sk_a.p; rows_by_sample = sr_a.p;
    radix_sort<uint32_t>(ctx, sk_a.p, sk_b.p, sr_a.p, sr_b.p, n_ref, bits_for((uint64_t)(n_ls - 1)), &keys_sorted, &rows_by_sample);
    device_scan<unsigned int, int64_t, OpSum<int64_t>, false>(ctx, sample_cnt.p, sample_off.p, n_ls, sample_off.p + n_ls);
    rank_of_row.alloc(ctx, n_ref);
    if (n_ref) LAUNCH(ctx, k_ref_rank, grid_for(ctx, n_ref, 256), 256, 0, rows_by_sample, keys_sorted, sample_off.p, n_ref, rank_of_row.p);
    pair_base.fill_byte(0xff);
    if (n_pairs) {
      LAUNCH(ctx, k_pair_rows, grid_for(ctx, n_pairs, 256), 256, 0, d_pl.p, d_pr.p, n_pairs, n_ls, n_rs, sample_cnt.p, row_off.p, bad.p);
      device_scan<int64_t, int64_t, OpSum<int64_t>, false>(ctx, row_off.p, row_off.p, n_pairs, row_off.p + n_pairs);
    } else {
      row_off.zero();
    }
    CUDA_TRY(cudaMemcpyAsync(&n_rows, row_off.p + n_pairs, sizeof n_rows, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
    GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "sample index out of range (regions or pair list)");
    if (n_pairs) LAUNCH(ctx, k_pair_base, grid_for(ctx, n_pairs, 256), 256, 0, d_pl.p, d_pr.p, n_pairs, n_rs, row_off.p, pair_base.p, bad.p);
  }

  // interval index over the reference
  LinMap lm = spans.map();
  IntervalIndex<K> ix;
  if (small_index) alloc_index<K>(ctx, n_ref, &ix);
  else build_index<K>(ctx, ref, nullptr, 1, lm, &ix, false, (ref.span_hint || ref.sorted_hint) ? bad.p : (int*)nullptr);
  // bucket table, resident in the streaming kernel's shared memory: as fine as fits (most buckets then hold no start
  // and a query needs no search at all), at least 2^10 buckets
  int tbl_shift = 0;
  // COUNT onto a small single-sample reference with the experiment rows grouped by sample: k_map_count_grouped
  const char* mg_env = ctx_opt(ctx, "map.grouped");         // "0" / "1": test hook
  const bool mg_compact = small_index && !ref.dup_of && !ref.strand;   // rank == index, no strand: 12 bytes per reference region
  const int64_t mg_fixed = (mg_compact ? 12 : 16) * n_ref + 4 * ((n_ref + 1) / 2) + 8 * (int64_t)lm.n_chrom + 16;
  const int64_t mg_tbl = (MG_SMEM - mg_fixed) / 2 - 16;   // 16-bit table entries
  // (16-bit counters: a count cannot exceed the sample's rows — unless duplicate reference records pool their counts on one row,
  // GenometricMap71.scala:186-203, so a reference with dup_of takes the streaming kernel)
  const bool grouped = single && sizeof(K) == 4 && plan.n_cols == 0 && !plan.need_bag && exp.sample_off != nullptr && exp.max_sample_rows <= 65535 && !ref.dup_of &&
                       n_ref > 0 && n_ref < 65536 && n_exp > 0 && n_pairs > 0 && mg_tbl >= 2048 && !(mg_env && mg_env[0] == '0') &&
                       (n_rs >= 4 * ctx->num_sms || (mg_env && mg_env[0] == '1'));   // one block per sample at a time: needs many samples to fill the GPU evenly
  {
    int64_t want = std::max<int64_t>(1024, std::min<int64_t>(grouped ? mg_tbl : (int64_t)MS_MAX_TBL - 2, (grouped ? 64 : 16) * std::max<int64_t>(n_ref, 1)));
    while (tbl_shift < 62 && (int64_t)(lm.G >> tbl_shift) + 2 > want) ++tbl_shift;
  }
  int64_t n_buckets = (int64_t)(lm.G >> tbl_shift) + 1;
  DevBuf<uint32_t> tbl(ctx, n_buckets + 2);
  DevBuf<RefEntry<K>> entries(ctx, n_ref);
  if (small_index) {
    LAUNCH(ctx, (k_map_index_small<K>), MIX_CL, MIX_T, 0, ref.chrom, ref.start, ref.stop, ref.coord_bits, (int)n_ref, lm, ref.sample, ref.strand, ref.dup_of, n_ls,
           tbl_shift, n_buckets, ix.start_a.p, ix.row_a.p, sr_a.p, ix.stop.p, ix.pmax.p, tbl.p, entries.p, bad.p);
  } else {
    LAUNCH(ctx, (k_bucket_table<K>), grid_for(ctx, n_buckets + 1, 256), 256, 0, ix.start, n_ref, tbl_shift, n_buckets, tbl.p);
    if (n_ref) LAUNCH(ctx, (k_index_entries<K>), grid_for(ctx, n_ref, 256), 256, 0, ix.row, n_ref, ix.stop.p, ix.pmax.p, ref.sample, ref.strand, ref.dup_of, single ? (const uint32_t*)nullptr : rank_of_row.p, n_ls, bad.p, entries.p);
  }

  // accumulators
  MapAcc acc;
  memset(&acc, 0, sizeof acc);
  DevBuf<int> count(ctx, n_rows);
  if (!grouped) count.zero();                            // the grouped kernel writes every row
  acc.count = count.p;
  DevBuf<int> nn[8];
  DevBuf<double> sum[8];
  DevBuf<unsigned long long> mn[8], mx[8];
  ExpCols ec;
  memset(&ec, 0, sizeof ec);
  for (int q = 0; q < plan.n_cols; ++q) {
    nn[q].alloc(ctx, n_rows); nn[q].zero(); acc.nn[q] = nn[q].p;
    if (plan.need_sum[q]) { sum[q].alloc(ctx, n_rows); sum[q].zero(); acc.sum[q] = sum[q].p; }
    if (plan.need_min[q]) { mn[q].alloc(ctx, n_rows); mn[q].fill_byte(0xff); acc.mn[q] = mn[q].p; }
    if (plan.need_max[q]) { mx[q].alloc(ctx, n_rows); mx[q].zero(); acc.mx[q] = mx[q].p; }
    ec.v[q] = exp.cols[plan.col_id[q]];
    ec.ok[q] = exp.valid[plan.col_id[q]];
  }

  // the per-sample COUNT kernel reads the narrow encodings as they are; everything else wants the int32 / int64 columns
  const bool exp_narrow = grouped && !exp.wide_ready && exp.coord_bits == 32 && exp.chrom_narrow && exp.chrom_bits == 8 && exp.len_narrow && exp.len_bits == 16 &&
                          exp.sample_trusted;
  if (!exp_narrow) gmql_ensure_wide(ctx, &exp);
  MapStreamArgs<K> ma_saved;
  memset(&ma_saved, 0, sizeof ma_saved);
  size_t sh_saved = 0;
  int g_saved = 1;
  if (grouped) {
    MapGroupedArgs ga;
    ga.e_chrom = exp.chrom; ga.e_start = exp.start; ga.e_stop = exp.stop; ga.e_bits = exp.coord_bits;
    ga.e_chrom8 = (const uint8_t*)exp.chrom_narrow; ga.e_len16 = (const uint16_t*)exp.len_narrow;
    ga.e_strand = exp.strand; ga.e_sample = exp.sample_trusted ? (const int32_t*)nullptr : exp.sample; ga.sample_off = exp.sample_off; ga.n_exp_samples = n_rs;
    ga.lm = lm; ga.r_start = (const uint32_t*)ix.start; ga.entries = (const RefEntry<uint32_t>*)entries.p; ga.n_ref = (int)n_ref;
    ga.pair_base = pair_base.p; ga.bad = bad.p;
    ga.tbl = tbl.p; ga.tbl_shift = tbl_shift; ga.n_tbl = (int)(n_buckets + 2);
    ga.count = count.p;
    ga.compact = mg_compact ? 1 : 0;
    const size_t sh = (size_t)mg_fixed + 2 * (size_t)((ga.n_tbl + 7) & ~7);
    const int mg_grid = (int)std::min<int64_t>(n_rs, (int64_t)ctx->num_sms);
    const char* lean_opt = ctx_opt(ctx, "map.lean");      // "0": test hook (the general per-sample kernel)
    const bool lean_ok = exp_narrow && mg_compact && (((uintptr_t)exp.start & 15u) | ((uintptr_t)exp.len_narrow & 7u) | ((uintptr_t)exp.chrom_narrow & 3u)) == 0 &&
                         !(lean_opt && lean_opt[0] == '0');
    if (lean_ok) {
      CUDA_TRY(cudaFuncSetAttribute(k_map_count_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      LAUNCH(ctx, k_map_count_lean, mg_grid, MS_THREADS, sh, ga, n_exp);
    } else if (exp_narrow) {
      CUDA_TRY(cudaFuncSetAttribute((k_map_count_grouped<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      LAUNCH(ctx, (k_map_count_grouped<true, true>), mg_grid, MS_THREADS, sh, ga);
    } else if (exp.coord_bits == 32) {
      CUDA_TRY(cudaFuncSetAttribute((k_map_count_grouped<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      LAUNCH(ctx, (k_map_count_grouped<true, false>), mg_grid, MS_THREADS, sh, ga);
    } else {
      CUDA_TRY(cudaFuncSetAttribute((k_map_count_grouped<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      LAUNCH(ctx, (k_map_count_grouped<false, false>), mg_grid, MS_THREADS, sh, ga);
    }
  } else if (n_exp && n_ref && n_rows) {
    MapStreamArgs<K> ma;
    ma.e_chrom = exp.chrom; ma.e_start = exp.start; ma.e_stop = exp.stop; ma.e_bits = exp.coord_bits;
    ma.e_strand = exp.strand; ma.e_sample = exp.sample; ma.n_exp = n_exp; ma.n_exp_samples = n_rs;
    ma.lm = lm; ma.r_start = ix.start; ma.entries = entries.p; ma.n_ref = n_ref;
    ma.pair_base = pair_base.p; ma.bad = bad.p;
    ma.tbl = tbl.p; ma.tbl_shift = tbl_shift; ma.n_tbl = (int)(n_buckets + 2);
    auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
    ma.vec_ok = al16(exp.chrom) && al16(exp.start) && al16(exp.stop) && al16(exp.sample) && ((uintptr_t)exp.strand & 3u) == 0;
    size_t sh = sizeof(uint32_t) * (size_t)((ma.n_tbl + 3) & ~3) + (lm.n_chrom <= MS_MAX_CHROM ? 16 * (size_t)lm.n_chrom : 0);
    CUDA_TRY(cudaFuncSetAttribute((k_map_stream<K, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    int g = (int)std::min<int64_t>(ceil_div64(n_exp, (int64_t)MS_THREADS * MS_PER), (int64_t)ctx->num_sms);
    ma.hit = nullptr; ma.exact = 0;
    ma.bag_off = nullptr; ma.bag_cursor = nullptr; ma.bag_rows = nullptr;
    LAUNCH(ctx, (k_map_stream<K, 0>), g, MS_THREADS, sh, ma, plan, ec, acc);
    ma_saved = ma; sh_saved = sh; g_saved = g;
  }
  // BAG / BAGD: count -> scan -> write of the matched experiment rows
  DevBuf<int64_t> bag_off;
  DevBuf<int32_t> bag_rows;
  if (plan.need_bag) {
    bag_off.alloc(ctx, n_rows + 1);
    device_scan<int, int64_t, OpSum<int64_t>, false>(ctx, count.p, bag_off.p, n_rows, bag_off.p + n_rows);
    int64_t n_match = n_rows ? read_back_i64(ctx, bag_off.p + n_rows) : 0;
    GMQL_REQUIRE(n_match < (int64_t)0x7fffffff * 4, GMQL_ERR_UNSUPPORTED, "BAG: match lists too large for this version");
    bag_rows.alloc(ctx, n_match);
    if (n_exp && n_ref && n_rows && n_match) {
      DevBuf<int32_t> cursor(ctx, n_rows);
      cursor.zero();
      ma_saved.bag_off = bag_off.p; ma_saved.bag_cursor = cursor.p; ma_saved.bag_rows = bag_rows.p;
      CUDA_TRY(cudaFuncSetAttribute((k_map_stream<K, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_saved));
      LAUNCH(ctx, (k_map_stream<K, 2>), g_saved, MS_THREADS, sh_saved, ma_saved, plan, ec, acc);
    }
  }
  if (ref.dup_of && n_pairs && n_rows)
    LAUNCH(ctx, k_mark_dups, (int)std::min<int64_t>(n_pairs, (int64_t)ctx->num_sms * 8), 128, 0, d_pl.p, row_off.p, n_pairs, sample_off.p, rows_by_sample, ref.dup_of, count.p);

  // outputs
  MapOut mo;
  memset(&mo, 0, sizeof mo);
  std::vector<DevBuf<double>> ov((size_t)plan.n_aggs);
  std::vector<DevBuf<uint8_t>> ok((size_t)plan.n_aggs);
  for (int a = 0; a < plan.n_aggs; ++a) {
    ov[a].alloc(ctx, n_rows); ok[a].alloc(ctx, n_rows);
    mo.value[a] = ov[a].p; mo.valid[a] = ok[a].p;
  }
  if (plan.n_aggs && n_rows) LAUNCH(ctx, k_map_finalize, grid_for(ctx, n_rows, 256), 256, 0, n_rows, plan, acc, mo);
  cudaEventRecord(ctx->ev_k1, ctx->stream);

  // REF_ROW: copy out of the ping-pong buffer that holds it (before the operator's one synchronisation)
  DevBuf<int32_t> ref_row(ctx, n_ref);
  if (n_ref) CUDA_TRY(cudaMemcpyAsync(ref_row.p, rows_by_sample, 4 * (size_t)n_ref, cudaMemcpyDeviceToDevice, ctx->stream));
  const char* defer = ctx_opt(ctx, "defer_check");
  if (defer && defer[0] == '1' && single) {
    // the caller asked not to wait: the flag travels to a pinned word and is looked at by the next call that finds the stream
    // idle (gmql_result_column_copy, gmql_ctx_sync, the start of the next operator); row counts and offsets are host-known
    if (!ctx->h_defer) { if (cudaHostAlloc((void**)&ctx->h_defer, 256 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw gmql_error(GMQL_ERR_NOMEM, "pinned host allocation failed"); } }
    if (ctx->deferred.size() >= 200) { CUDA_TRY(GMQL_SYNC(ctx)); GMQL_REQUIRE(gmql_check_deferred(ctx) == GMQL_OK, GMQL_ERR_INVALID, ctx->err); }
    int* slot = ctx->h_defer + (ctx->h_defer_next++ & 255);
    *slot = 0;
    CUDA_TRY(cudaMemcpyAsync(slot, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    gmql_ctx::Deferred df; df.h_flag = slot; df.what = "gmql_map: invalid region / sample index, reference region beyond its chrom_span_hint, reference rows out of the promised order, or a sample column that contradicts sample_offsets";
    ctx->deferred.push_back(df);
    ctx->last_deferred = true;
    h_bad = 0;
  } else {
    CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
  }
  GMQL_REQUIRE(h_bad != 2, GMQL_ERR_INVALID, "the same (ref sample, exp sample) pair is listed twice");
  GMQL_REQUIRE(h_bad != 3, GMQL_ERR_INVALID, "the sample column does not match sample_offsets (rows must be grouped by sample in offset order)");
  GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "region with chromosome/sample index out of range, negative start, reference region beyond its chrom_span_hint, or reference rows not in the promised (chrom, start) order");

  res->rows = n_rows;
  res->kind = 1;
  result_add(res, GMQL_COL_MAP_ROW_OFFSETS, row_off);
  result_add(res, GMQL_COL_MAP_REF_ROW, ref_row);
  result_add(res, GMQL_COL_MAP_REF_SAMPLE_OFFSETS, sample_off);
  result_add(res, GMQL_COL_MAP_COUNT, count);
  if (plan.need_bag) { result_add(res, GMQL_COL_BAG_OFFSETS, bag_off); result_add(res, GMQL_COL_BAG_ROWS, bag_rows); }
  for (int a = 0; a < plan.n_aggs; ++a) {
    result_add(res, GMQL_COL_AGG_VALUE + a, ov[a]);
    result_add(res, GMQL_COL_AGG_VALID + a, ok[a]);
  }
}

// ---- DIFFERENCE ------------------------------------------------------------------------------------
__global__ void k_diff_keep(const int32_t* __restrict__ sample, const int32_t* __restrict__ dup_of, const uint8_t* __restrict__ has_group, const uint8_t* __restrict__ hit,
                            int64_t n, int n_samples, int32_t* __restrict__ keep, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int sm = sample ? sample[i] : 0;
    if (sm < 0 || sm >= n_samples) { *bad = 1; keep[i] = 0; continue; }
    // identical records of one sample collapse into one output record (the reduce key is the record's content,
    // GenometricDifference.scala:54,79-81): only the canonical row can be emitted, and it carries every duplicate's hits
    bool canon = !dup_of || dup_of[i] == (int32_t)i;
    keep[i] = (canon && has_group[sm] && !hit[i]) ? 1 : 0;
  }
}

__global__ void k_diff_compact(const int32_t* __restrict__ keep, const int64_t* __restrict__ excl, int64_t n, int32_t* __restrict__ rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (keep[i]) rows[excl[i]] = (int32_t)i;
}

template <typename K>
static void run_difference(gmql_ctx* ctx, const DevRegions& ref, const DevRegions& exp, const gmql_pair* pairs, int64_t n_pairs, int exact,
                           const SpanTable& spans, gmql_result* res) {
  const int64_t n_ref = ref.n, n_exp = exp.n;
  const int n_ls = ref.n_samples, n_rs = exp.n_samples;
  DevBuf<int> bad(ctx, 1);
  bad.zero();
  // pair table and "sample has a group" flags, built on the host
  long long* h_pb = (long long*)ctx_pinned(ctx, 8 * (size_t)n_ls * (size_t)n_rs);
  uint8_t* h_hg = (uint8_t*)ctx_pinned(ctx, (size_t)n_ls);
  for (int64_t q = 0; q < (int64_t)n_ls * n_rs; ++q) h_pb[q] = -1ll;
  memset(h_hg, 0, (size_t)n_ls);
  for (int64_t p = 0; p < n_pairs; ++p) {
    int a = pairs[p].left_sample, b = pairs[p].right_sample;
    GMQL_REQUIRE(a >= 0 && a < n_ls && b >= 0 && b < n_rs, GMQL_ERR_INVALID, "sample index out of range in the pair list");
    h_pb[(int64_t)a * n_rs + b] = 0;
    h_hg[a] = 1;
  }
  DevBuf<long long> pair_base(ctx, (int64_t)n_ls * n_rs);
  DevBuf<uint8_t> has_group(ctx, n_ls), hit(ctx, n_ref);
  CUDA_TRY(cudaMemcpyAsync(pair_base.p, h_pb, 8 * (size_t)n_ls * (size_t)n_rs, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(has_group.p, h_hg, (size_t)n_ls, cudaMemcpyHostToDevice, ctx->stream));
  hit.zero();
  cudaEventRecord(ctx->ev_h1, ctx->stream);

  LinMap lm = spans.map();
  IntervalIndex<K> ix;
  build_index<K>(ctx, ref, nullptr, 1, lm, &ix, false, (ref.span_hint || ref.sorted_hint) ? bad.p : (int*)nullptr);
  int tbl_shift = 0;
  {
    int64_t want = std::max<int64_t>(1024, std::min<int64_t>((int64_t)MS_MAX_TBL - 2, 16 * std::max<int64_t>(n_ref, 1)));
    while (tbl_shift < 62 && (int64_t)(lm.G >> tbl_shift) + 2 > want) ++tbl_shift;
  }
  int64_t n_buckets = (int64_t)(lm.G >> tbl_shift) + 1;
  DevBuf<uint32_t> tbl(ctx, n_buckets + 2);
  LAUNCH(ctx, (k_bucket_table<K>), grid_for(ctx, n_buckets + 1, 256), 256, 0, ix.start, n_ref, tbl_shift, n_buckets, tbl.p);
  DevBuf<RefEntry<K>> entries(ctx, n_ref);
  if (n_ref) LAUNCH(ctx, (k_index_entries<K>), grid_for(ctx, n_ref, 256), 256, 0, ix.row, n_ref, ix.stop.p, ix.pmax.p, ref.sample, ref.strand, ref.dup_of, (const uint32_t*)nullptr, n_ls, bad.p, entries.p);
  if (n_exp && n_ref && n_pairs) {
    MapStreamArgs<K> ma;
    ma.e_chrom = exp.chrom; ma.e_start = exp.start; ma.e_stop = exp.stop; ma.e_bits = exp.coord_bits;
    ma.e_strand = exp.strand; ma.e_sample = exp.sample; ma.n_exp = n_exp; ma.n_exp_samples = n_rs;
    ma.lm = lm; ma.r_start = ix.start; ma.entries = entries.p; ma.n_ref = n_ref;
    ma.pair_base = pair_base.p; ma.bad = bad.p;
    ma.tbl = tbl.p; ma.tbl_shift = tbl_shift; ma.n_tbl = (int)(n_buckets + 2);
    auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
    ma.vec_ok = al16(exp.chrom) && al16(exp.start) && al16(exp.stop) && al16(exp.sample) && ((uintptr_t)exp.strand & 3u) == 0;
    ma.hit = hit.p; ma.exact = exact;
    ma.bag_off = nullptr; ma.bag_cursor = nullptr; ma.bag_rows = nullptr;
    size_t sh = sizeof(uint32_t) * (size_t)((ma.n_tbl + 3) & ~3) + (lm.n_chrom <= MS_MAX_CHROM ? 16 * (size_t)lm.n_chrom : 0);
    CUDA_TRY(cudaFuncSetAttribute((k_map_stream<K, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    int g = (int)std::min<int64_t>(ceil_div64(n_exp, (int64_t)MS_THREADS * MS_PER), (int64_t)ctx->num_sms);
    MapAggPlan plan = MapAggPlan();
    ExpCols ec; memset(&ec, 0, sizeof ec);
    MapAcc acc; memset(&acc, 0, sizeof acc);
    LAUNCH(ctx, (k_map_stream<K, 1>), g, MS_THREADS, sh, ma, plan, ec, acc);
  }
  DevBuf<int32_t> keep(ctx, n_ref);
  DevBuf<int64_t> excl(ctx, n_ref + 1);
  if (n_ref) LAUNCH(ctx, k_diff_keep, grid_for(ctx, n_ref, 256), 256, 0, ref.sample, ref.dup_of, has_group.p, hit.p, n_ref, n_ls, keep.p, bad.p);
  device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, keep.p, excl.p, n_ref, excl.p + n_ref);
  int64_t n_out = 0;
  int h_bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&n_out, excl.p + n_ref, sizeof n_out, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(GMQL_SYNC(ctx));
  GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "region with chromosome/sample index out of range, negative start, reference region beyond its chrom_span_hint, or reference rows not in the promised (chrom, start) order");
  DevBuf<int32_t> rows(ctx, n_out);
  if (n_ref && n_out) LAUNCH(ctx, k_diff_compact, grid_for(ctx, n_ref, 256), 256, 0, keep.p, excl.p, n_ref, rows.p);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  res->rows = n_out;
  res->kind = 4;
  result_add(res, GMQL_COL_DIFF_REF_ROW, rows);
}

}  // namespace

extern "C" int gmql_map(gmql_ctx* ctx, const gmql_regions* ref, const gmql_regions* exp, const gmql_pair* pairs, int64_t n_pairs,
                        const gmql_agg* aggs, int32_t n_aggs, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    gmql_trace("ENTER gmql_map");
    GMQL_REQUIRE(ref && exp, GMQL_ERR_INVALID, "null dataset");
    GMQL_REQUIRE(n_pairs >= 0 && (n_pairs == 0 || pairs), GMQL_ERR_INVALID, "bad pair list");
    GMQL_REQUIRE(ref->n_chrom == exp->n_chrom, GMQL_ERR_INVALID, "ref and exp must share one chromosome dictionary (n_chrom differs)");
    MapAggPlan plan;
    std::vector<int> need_cols;
    plan_aggs(aggs, n_aggs, exp->n_cols, &plan, &need_cols);
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions dref, dexp;
    gmql_upload_regions(ctx, ref, std::vector<int>(), &dref);
    gmql_upload_regions(ctx, exp, need_cols, &dexp, /*keep_narrow=*/true);   // run_map widens unless its per-sample COUNT kernel takes the rows as they are
    SpanTable spans;
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    if (dref.span_hint) span_table_from_hint(ctx, dref.span_hint, dref.n_chrom, &spans);
    else build_span_table(ctx, &dref, nullptr, &spans);
    res = new gmql_result();
    res->owner = ctx;
    ctx->last_deferred = false;
    if (spans.G < 0xffffffffull) run_map<uint32_t>(ctx, dref, dexp, pairs, n_pairs, plan, spans, res);
    else run_map<uint64_t>(ctx, dref, dexp, pairs, n_pairs, plan, spans, res);
    if (!ctx->last_deferred) CUDA_TRY(GMQL_SYNC_IF_PENDING(ctx));
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

extern "C" int gmql_difference(gmql_ctx* ctx, const gmql_regions* ref, const gmql_regions* exp, const gmql_pair* pairs, int64_t n_pairs,
                               int32_t exact, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(ref && exp, GMQL_ERR_INVALID, "null dataset");
    GMQL_REQUIRE(n_pairs >= 0 && (n_pairs == 0 || pairs), GMQL_ERR_INVALID, "bad pair list");
    GMQL_REQUIRE(ref->n_chrom == exp->n_chrom, GMQL_ERR_INVALID, "ref and exp must share one chromosome dictionary (n_chrom differs)");
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions dref, dexp;
    gmql_upload_regions(ctx, ref, std::vector<int>(), &dref);
    gmql_upload_regions(ctx, exp, std::vector<int>(), &dexp);
    SpanTable spans;
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    if (dref.span_hint) span_table_from_hint(ctx, dref.span_hint, dref.n_chrom, &spans);
    else build_span_table(ctx, &dref, nullptr, &spans);
    res = new gmql_result();
    res->owner = ctx;
    if (spans.G < 0xffffffffull) run_difference<uint32_t>(ctx, dref, dexp, pairs, n_pairs, exact != 0, spans, res);
    else run_difference<uint64_t>(ctx, dref, dexp, pairs, n_pairs, exact != 0, spans, res);
    CUDA_TRY(GMQL_SYNC(ctx));
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
