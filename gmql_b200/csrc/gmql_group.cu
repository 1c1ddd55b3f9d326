// gmql_group.cu — GROUP by coordinates on the device (SURVEY §8f-2's parenthetical).
//
// Replaces GroupRD.apply (GMQL-Spark/.../RegionsOperators/GroupRD.scala:22-64): the records of one sample that agree on
// (chrom, start, stop, strand) and on the text forms of the grouping fields (:34-42; the host knows strings, so it passes one
// equivalence code per row, as gmql_join's join-on-attributes codes) collapse into ONE record; every aggregate is applied to the
// group's whole value list (:48-55) — `fun` is n-ary here, not folded pairwise as in MAP / COVER, so MEDIAN is the true (lower)
// median of DefaultRegionsToRegionFactory.scala:33-56: sorted non-null values, element (m - 1) / 2.
//
// Design: nothing is hashed.  Rows are sorted by (sample, chromosome, start) with the interval index's radix sort (stable: equal
// keys keep their input order); rows of one sample starting at one position form a short run, inside which every row looks back
// for the first row with the same stop, strand and field code — its group's head.  Heads are numbered by a scan; counts, sums and
// min / max go to per-group accumulators with atomics (groups are tiny: they are duplicates), the median is selected by walking
// the group's distinct values in increasing order (no storage), and the members of every group are listed (input order) for the
// host's BAG / BAGD strings.  Cost: the sort (n log-free passes) + O(run length) per row.
#include "index.cuh"

namespace {

__global__ void k_group_classes(const int32_t* __restrict__ sample, int64_t n, int n_samples, uint32_t* __restrict__ cls, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = sample ? sample[i] : 0;
    if (s < 0 || s >= n_samples) { *bad = 1; s = 0; }
    cls[i] = (uint32_t)s;
  }
}

// per sorted entry j: head[j] = 1 when no earlier entry of its run equals it; rep[j] = sorted index of its head; rank[j] = members
// of its group before it
__global__ void k_group_rep(const uint64_t* __restrict__ key, const uint64_t* __restrict__ lstop, const uint32_t* __restrict__ row, int64_t n,
                            const int8_t* __restrict__ strand, const int64_t* __restrict__ code, int32_t* __restrict__ head, uint32_t* __restrict__ rep, uint32_t* __restrict__ rank) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = key[j], e = lstop[j];
    const uint32_t r = row[j];
    const int st = strand ? (int)strand[r] : 0;
    const int64_t cd = code ? code[r] : 0;
    int64_t first = j;
    uint32_t before = 0;
    for (int64_t i = j - 1; i >= 0 && key[i] == k; --i) {
      const uint32_t ri = row[i];
      if (lstop[i] == e && (strand ? (int)strand[ri] : 0) == st && (code ? code[ri] : 0) == cd) { first = i; ++before; }
    }
    head[j] = first == j ? 1 : 0;
    rep[j] = (uint32_t)first;
    rank[j] = before;
  }
}

struct GroupAcc {
  int32_t* count;                 // [n_groups] members
  int32_t* rep_row;               // [n_groups] input row of the head (the smallest input row of the group)
  int* nn[8]; double* sum[8]; unsigned long long* mn[8]; unsigned long long* mx[8];
};
struct GroupCols { const double* v[8]; const uint8_t* ok[8]; int n; };

// total order of java.lang.Double.compare (what `sorted` on List[Double] uses): -0.0 < 0.0, NaN above everything
__device__ __forceinline__ unsigned long long dbl_order_key(double v) {
  unsigned long long b = (v != v) ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// Math.min / Math.max keys (NaN wins both; DefaultRegionsToRegionFactory.scala:84-110 reduce with them)
__device__ __forceinline__ unsigned long long dbl_key_min(double v) {
  if (v != v) return 0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long dbl_key_max(double v) {
  if (v != v) return ~0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dbl_from_key(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void k_group_init(GroupAcc acc, int n_cols, int64_t ng) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
    acc.count[g] = 0;
    for (int q = 0; q < n_cols; ++q) { acc.nn[q][g] = 0; acc.sum[q][g] = 0.0; acc.mn[q][g] = ~0ull; acc.mx[q][g] = 0ull; }
  }
}

__global__ void k_group_accumulate(const uint32_t* __restrict__ row, const uint32_t* __restrict__ rep, const int64_t* __restrict__ head_excl, int64_t n, GroupCols cols, GroupAcc acc,
                                   uint32_t* __restrict__ group_of) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = head_excl[rep[j]];
    const uint32_t r = row[j];
    group_of[j] = (uint32_t)g;
    if (rep[j] == (uint32_t)j) acc.rep_row[g] = (int32_t)r;
    atomicAdd(&acc.count[g], 1);
    for (int q = 0; q < cols.n; ++q) {
      if (cols.ok[q] && !cols.ok[q][r]) continue;
      const double v = cols.v[q][r];
      atomicAdd(&acc.nn[q][g], 1);
      atomicAdd(&acc.sum[q][g], v);
      atomicMin(&acc.mn[q][g], dbl_key_min(v));
      atomicMax(&acc.mx[q][g], dbl_key_max(v));
    }
  }
}

// members of every group in input order (the sort is stable): for the host's BAG / BAGD strings
__global__ void k_group_members(const uint32_t* __restrict__ row, const uint32_t* __restrict__ group_of, const uint32_t* __restrict__ rank, const int64_t* __restrict__ off, int64_t n,
                                int32_t* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    out[off[group_of[j]] + rank[j]] = (int32_t)row[j];
}

// true (lower) median of a group's non-null values: the element at sorted position (m - 1) / 2, found by walking the distinct
// values in increasing Double.compare order over the group's members (they sit in the head's run, at or after the head)
__global__ void k_group_median(const uint64_t* __restrict__ key, const uint32_t* __restrict__ row, const uint32_t* __restrict__ rep, const int32_t* __restrict__ head,
                               const int64_t* __restrict__ head_excl, int64_t n, const double* __restrict__ v, const uint8_t* __restrict__ ok, const int* __restrict__ nn,
                               double* __restrict__ out, uint8_t* __restrict__ out_ok) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    if (!head[j]) continue;
    const int64_t g = head_excl[j];
    const int m = nn[g];
    if (m == 0) { out[g] = 0.0; out_ok[g] = 0; continue; }
    const int target = (m - 1) / 2;
    const uint64_t k = key[j];
    int seen = 0;
    unsigned long long prev = 0ull;
    bool first = true;
    while (true) {
      unsigned long long cur = 0ull;
      int ncur = 0;
      for (int64_t i = j; i < n && key[i] == k; ++i) {
        if (rep[i] != (uint32_t)j) continue;
        const uint32_t r = row[i];
        if (ok && !ok[r]) continue;
        const unsigned long long x = dbl_order_key(v[r]);
        if (!first && x <= prev) continue;
        if (ncur == 0 || x < cur) { cur = x; ncur = 1; }
        else if (x == cur) ++ncur;
      }
      seen += ncur; prev = cur; first = false;
      if (seen > target) break;
    }
    const double x = dbl_from_key(prev);
    out[g] = (m & 1) ? x : (x + x) / 2;                     // `(values(left) + values(left)) / 2` (:47): x unless x + x overflows
    out_ok[g] = 1;
  }
}

__global__ void k_group_finish(int kind, int64_t ng, const int32_t* __restrict__ count, const int* __restrict__ nn, const double* __restrict__ sum,
                               const unsigned long long* __restrict__ mn, const unsigned long long* __restrict__ mx, double* __restrict__ out, uint8_t* __restrict__ out_ok) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    uint8_t okv = 0;
    if (kind == GMQL_AGG_COUNT) { v = (double)count[g]; okv = 1; }                                   // funOut: GDouble(list length) (:58-67)
    else if (nn[g] > 0) {
      okv = 1;
      if (kind == GMQL_AGG_SUM) v = sum[g];
      else if (kind == GMQL_AGG_AVG) v = sum[g] / (double)nn[g];                                      // :112-126
      else if (kind == GMQL_AGG_MIN) v = dbl_from_key(mn[g]);
      else if (kind == GMQL_AGG_MAX) v = dbl_from_key(mx[g]);
    }
    out[g] = v; out_ok[g] = okv;
  }
}

}  // namespace

extern "C" int gmql_group(gmql_ctx* ctx, const gmql_regions* in, const int64_t* field_code, const gmql_agg* aggs, int32_t n_aggs, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(in != nullptr, GMQL_ERR_INVALID, "null dataset");
    GMQL_REQUIRE(n_aggs >= 0 && n_aggs <= 16 && (n_aggs == 0 || aggs), GMQL_ERR_INVALID, "0..16 aggregates");
    // distinct aggregated columns
    int slot[16], col_id[8], n_cols = 0;
    bool need_bag = false;
    std::vector<int> need_cols;
    for (int a = 0; a < n_aggs; ++a) {
      slot[a] = -1;
      const int kind = aggs[a].kind;
      if (kind == GMQL_AGG_COUNT) continue;
      if (kind == GMQL_AGG_BAG || kind == GMQL_AGG_BAGD) { need_bag = true; continue; }
      GMQL_REQUIRE(kind >= GMQL_AGG_SUM && kind <= GMQL_AGG_MEDIAN, GMQL_ERR_INVALID, "unknown aggregate kind");
      GMQL_REQUIRE(aggs[a].col >= 0 && aggs[a].col < in->n_cols, GMQL_ERR_INVALID, "aggregate column index out of range");
      int q = -1;
      for (int j = 0; j < n_cols; ++j) if (col_id[j] == aggs[a].col) q = j;
      if (q < 0) {
        GMQL_REQUIRE(n_cols < 8, GMQL_ERR_UNSUPPORTED, "at most 8 distinct aggregated columns per call");
        q = n_cols++; col_id[q] = aggs[a].col; need_cols.push_back(aggs[a].col);
      }
      slot[a] = q;
    }
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions d;
    gmql_upload_regions(ctx, in, need_cols, &d);
    const int64_t n = d.n;
    DevBuf<int64_t> d_code;
    const int64_t* p_code = nullptr;
    if (field_code && n) {
      if (in->location == GMQL_LOC_DEVICE) p_code = field_code;
      else { d_code.alloc(ctx, n); CUDA_TRY(cudaMemcpyAsync(d_code.p, field_code, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream)); p_code = d_code.p; }
    }
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    res = new gmql_result();
    res->owner = ctx;
    res->kind = 6;

    int64_t ng = 0;
    DevBuf<int32_t> g_count, g_row, bag_rows;
    DevBuf<int64_t> bag_off;
    std::vector<DevBuf<double>> val((size_t)n_aggs);
    std::vector<DevBuf<uint8_t>> valid((size_t)n_aggs);
    if (n > 0) {
      SpanTable spans;
      build_span_table(ctx, &d, nullptr, &spans);
      LinMap lm = spans.map();
      DevBuf<int> bad(ctx, 1);
      bad.zero();
      DevBuf<uint32_t> cls(ctx, n);
      LAUNCH(ctx, k_group_classes, grid_for(ctx, n, 256), 256, 0, d.sample, n, d.n_samples, cls.p, bad.p);
      IntervalIndex<uint64_t> ix;
      build_index<uint64_t>(ctx, d, cls.p, (uint64_t)std::max(1, d.n_samples), lm, &ix);
      DevBuf<int32_t> head(ctx, n);
      DevBuf<uint32_t> rep(ctx, n), rank(ctx, n), group_of(ctx, n);
      DevBuf<int64_t> head_excl(ctx, n + 1);
      const int g = grid_for(ctx, n, 256);
      LAUNCH(ctx, k_group_rep, g, 256, 0, ix.start, ix.stop.p, ix.row, n, d.strand, p_code, head.p, rep.p, rank.p);
      device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, head.p, head_excl.p, n, head_excl.p + n);
      int h_bad = 0;
      CUDA_TRY(cudaMemcpyAsync(&ng, head_excl.p + n, sizeof ng, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(GMQL_SYNC(ctx));
      GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "sample index out of range");
      g_count.alloc(ctx, ng); g_row.alloc(ctx, ng);
      std::vector<DevBuf<int>> nn((size_t)n_cols);
      std::vector<DevBuf<double>> sum((size_t)n_cols);
      std::vector<DevBuf<unsigned long long>> mn((size_t)n_cols), mx((size_t)n_cols);
      GroupAcc acc;
      memset(&acc, 0, sizeof acc);
      acc.count = g_count.p; acc.rep_row = g_row.p;
      GroupCols gc;
      memset(&gc, 0, sizeof gc);
      gc.n = n_cols;
      for (int q = 0; q < n_cols; ++q) {
        nn[q].alloc(ctx, ng); sum[q].alloc(ctx, ng); mn[q].alloc(ctx, ng); mx[q].alloc(ctx, ng);
        acc.nn[q] = nn[q].p; acc.sum[q] = sum[q].p; acc.mn[q] = mn[q].p; acc.mx[q] = mx[q].p;
        gc.v[q] = d.cols[col_id[q]]; gc.ok[q] = d.valid[col_id[q]];
      }
      const int gg = grid_for(ctx, ng, 256);
      LAUNCH(ctx, k_group_init, gg, 256, 0, acc, n_cols, ng);
      LAUNCH(ctx, k_group_accumulate, g, 256, 0, ix.row, rep.p, head_excl.p, n, gc, acc, group_of.p);
      for (int a = 0; a < n_aggs; ++a) {
        const int kind = aggs[a].kind;
        if (kind == GMQL_AGG_BAG || kind == GMQL_AGG_BAGD) continue;
        val[a].alloc(ctx, ng); valid[a].alloc(ctx, ng);
        const int q = slot[a];
        if (kind == GMQL_AGG_MEDIAN)
          LAUNCH(ctx, k_group_median, g, 256, 0, ix.start, ix.row, rep.p, head.p, head_excl.p, n, gc.v[q], gc.ok[q], nn[q].p, val[a].p, valid[a].p);
        else
          LAUNCH(ctx, k_group_finish, gg, 256, 0, kind, ng, g_count.p, q >= 0 ? nn[q].p : (const int*)nullptr, q >= 0 ? sum[q].p : (const double*)nullptr,
                 q >= 0 ? mn[q].p : (const unsigned long long*)nullptr, q >= 0 ? mx[q].p : (const unsigned long long*)nullptr, val[a].p, valid[a].p);
      }
      if (need_bag) {
        bag_off.alloc(ctx, ng + 1);
        bag_rows.alloc(ctx, n);
        device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, g_count.p, bag_off.p, ng, bag_off.p + ng);
        LAUNCH(ctx, k_group_members, g, 256, 0, ix.row, group_of.p, rank.p, bag_off.p, n, bag_rows.p);
      }
      CUDA_TRY(GMQL_SYNC(ctx));   // the accumulators above go out of scope
    } else if (need_bag) {
      bag_off.alloc(ctx, 1);
      bag_off.zero();
    }
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    res->rows = ng;
    result_add(res, GMQL_COL_GROUP_ROW, g_row);
    result_add(res, GMQL_COL_GROUP_COUNT, g_count);
    for (int a = 0; a < n_aggs; ++a) {
      if (aggs[a].kind == GMQL_AGG_BAG || aggs[a].kind == GMQL_AGG_BAGD) continue;
      if (!val[a].p && ng == 0) { val[a].alloc(ctx, 0); valid[a].alloc(ctx, 0); }
      result_add(res, GMQL_COL_AGG_VALUE + a, val[a]);
      result_add(res, GMQL_COL_AGG_VALID + a, valid[a]);
    }
    if (need_bag) { result_add(res, GMQL_COL_BAG_OFFSETS, bag_off); result_add(res, GMQL_COL_BAG_ROWS, bag_rows); }
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
