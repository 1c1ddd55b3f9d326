// gmql_text.cu — columns -> text: the per-region work of MATERIALIZE's writer on the device.
//
// Replaces the line construction of StoreTABRD (GMQL-Spark/.../RegionsOperators/SelectRegions/StoreTABRD.scala:62-80):
//     (recordKey.productIterator.drop(1) ++ values).mkString("\t")
// = chrom \t start \t stop \t strand \t value ... with GRecordKey.strand a Char ('*', '+', '-'), start + 1 for one-based schemas
// (:62-69), GDouble.toString = DecimalFormat("#.########", ENGLISH) with RoundingMode.FLOOR (GMQL-Core/.../DataTypes.scala:139-150)
// and GNull.toString = "null" (:185).  One line per region, '\n'-terminated, in input row order (StoreTABRD partitions by sample id:
// with rows grouped by sample a sample file is one contiguous byte range).
//
// GDouble.toString on the device.  DecimalFormat rounds the SHORTEST decimal that identifies the double (the digits of
// Double.toString), not the binary value: 0.3 prints "0.3" although the double lies below 0.3.  For |v| < 2^26 the spacing of doubles
// is below 1e-8, so at most one multiple t of 1e-8 lies in the double's rounding interval, and then the shortest decimal IS t (any
// decimal with more than eight fraction digits is longer); otherwise no multiple does and flooring the shortest decimal equals
// flooring the exact value.  So: m = floor(|v| * 1e8) in exact 128-bit integer arithmetic, and the neighbouring multiple is taken
// instead when it parses back to v (distance <= half an ulp, ties to the even mantissa) — no digit-generation tables at all.
// Integer-valued doubles below 2^53 print their integer.  Everything else (non-integers from 2^26 on, magnitudes from 2^53 on, NaN,
// infinities) is left to the host's own DecimalFormat: the field is written as the one-byte placeholder 0x1A and the row's status
// bit 0 is set — the same contract as gmql_parse_text's undecidable literals.
#include "index.cuh"

namespace {

struct TextFmt {
  const int32_t* chrom; const void* start; const void* stop; int bits; const int8_t* strand;
  int n_values; const double* v[16]; const uint8_t* ok[16];
  const char* names; const int32_t* name_off; int n_chrom; int one_based;
};

template <bool WRITE>
struct Writer {
  char* p; int n;
  __device__ __forceinline__ void put(char c) { if (WRITE) p[n] = c; ++n; }
  __device__ __forceinline__ void put_u64(unsigned long long x) {
    int nd = 1;
    for (unsigned long long t = x; t >= 10ull; t /= 10ull) ++nd;
    if (WRITE) { for (int k = nd - 1; k >= 0; --k) { p[n + k] = (char)('0' + (int)(x % 10ull)); x /= 10ull; } }
    n += nd;
  }
  __device__ __forceinline__ void put_i64(long long x) {
    if (x < 0) { put('-'); put_u64(0ull - (unsigned long long)x); } else put_u64((unsigned long long)x);
  }
};

// GDouble.toString; returns false when the value is left to the host (placeholder written)
template <bool WRITE>
__device__ bool put_gdouble(Writer<WRITE>& w, double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const bool neg = (bits >> 63) != 0;
  const int bexp = (int)((bits >> 52) & 0x7ffull);
  unsigned long long mant = bits & 0xfffffffffffffull;
  if (bexp == 0x7ff) { w.put((char)0x1a); return false; }                     // NaN, infinities: the host's DecimalFormat symbols
  int e;                                                                       // |v| = mant * 2^e
  if (bexp == 0) e = -1074; else { mant |= 1ull << 52; e = bexp - 1075; }
  if (mant == 0ull) { if (neg) w.put('-'); w.put('0'); return true; }           // DecimalFormat prints "-0" for -0.0
  if (e >= 0 || (e > -53 && (mant & ((1ull << -e) - 1ull)) == 0ull)) {
    // integer-valued
    if (e > 0) { w.put((char)0x1a); return false; }                             // >= 2^53: the shortest digits are not the integer's
    if (neg) w.put('-');
    w.put_u64(e == 0 ? mant : (mant >> -e));
    return true;
  }
  if (bexp - 1023 >= 26) { w.put((char)0x1a); return false; }                   // non-integer with an ulp of 1e-8 or more
  // P = mant * 10^8 (< 2^80) as hi:lo; |v| * 10^8 = P / 2^s, s = -e in [27, 1074]
  const unsigned long long lo = mant * 100000000ull, hi = __umul64hi(mant, 100000000ull);
  const int s = -e;
  unsigned long long m, rem_hi, rem_lo;                                          // quotient, remainder (below 2^s)
  if (s >= 128) { m = 0ull; rem_hi = hi; rem_lo = lo; }
  else if (s >= 64) { m = s == 64 ? hi : (hi >> (s - 64)); rem_hi = s == 64 ? 0ull : (hi & ((1ull << (s - 64)) - 1ull)); rem_lo = lo; }
  else { m = (hi << (64 - s)) | (lo >> s); rem_hi = 0ull; rem_lo = lo & ((1ull << s) - 1ull); }
  const bool exact = (rem_hi | rem_lo) == 0ull;
  // is a neighbouring multiple of 1e-8 the double's shortest decimal?  Its distance from |v| in units of 2^e * 1e-8 is rem (the
  // multiple below) or 2^s - rem (the one above); half an ulp is 5e7 of those units; a tie parses to the even mantissa.
  const unsigned long long half = 50000000ull;
  const bool even = (mant & 1ull) == 0ull;
  const bool below_in = rem_hi == 0ull && (rem_lo < half || (rem_lo == half && even));
  bool above_in = false;
  if (s < 128) {
    // gap = 2^s - rem as hi:lo
    unsigned long long g_lo = 0ull - rem_lo, g_hi = (s >= 64 ? (1ull << (s - 64)) : 0ull) - rem_hi - (rem_lo != 0ull ? 1ull : 0ull);
    if (s < 64) { g_lo = (1ull << s) - rem_lo; g_hi = 0ull; }
    above_in = g_hi == 0ull && (g_lo < half || (g_lo == half && even));
  }
  // FLOOR: towards minus infinity
  unsigned long long q;
  if (!neg) q = (!exact && above_in) ? m + 1ull : m;
  else q = (exact || below_in) ? m : m + 1ull;
  const unsigned long long ip = q / 100000000ull;
  unsigned int fp = (unsigned int)(q % 100000000ull);
  if (neg) w.put('-');                                                           // q > 0 for every negative value here ("-0" only for -0.0, above)
  w.put_u64(ip);
  if (fp != 0u) {
    w.put('.');
    int nd = 8;
    while (fp % 10u == 0u) { fp /= 10u; --nd; }
    if (WRITE) { for (int k = nd - 1; k >= 0; --k) { w.p[w.n + k] = (char)('0' + (int)(fp % 10u)); fp /= 10u; } }
    w.n += nd;
  }
  return true;
}

template <bool WRITE>
__device__ int put_line(const TextFmt& f, int64_t i, char* dst, bool* host_field) {
  Writer<WRITE> w;
  w.p = dst; w.n = 0;
  const int c = f.chrom[i];
  if (c >= 0 && c < f.n_chrom) { for (int k = f.name_off[c]; k < f.name_off[c + 1]; ++k) w.put(f.names[k]); }
  else w.put('?');
  w.put('\t');
  w.put_i64((long long)ld_coord(f.start, f.bits, i) + (f.one_based ? 1 : 0));
  w.put('\t');
  w.put_i64((long long)ld_coord(f.stop, f.bits, i));
  w.put('\t');
  const int st = f.strand ? (int)f.strand[i] : 0;
  w.put(st == GMQL_STRAND_PLUS ? '+' : (st == GMQL_STRAND_MINUS ? '-' : '*'));
  bool host = false;
  for (int k = 0; k < f.n_values; ++k) {
    w.put('\t');
    if (f.ok[k] && !f.ok[k][i]) { w.put('n'); w.put('u'); w.put('l'); w.put('l'); }
    else if (!put_gdouble<WRITE>(w, f.v[k][i])) host = true;
  }
  w.put('\n');
  *host_field = host;
  return w.n;
}

__global__ void k_fmt_len(TextFmt f, int64_t n, int32_t* __restrict__ len, uint8_t* __restrict__ status) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bool host;
    len[i] = put_line<false>(f, i, nullptr, &host);
    status[i] = host ? 1 : 0;
  }
}

// A block formats FMT_ROWS consecutive rows into shared memory — placed so that shared and global addresses agree modulo 16 — and
// copies the finished byte range out with 16-byte stores; a range that does not fit is written straight to global memory.
constexpr int FMT_ROWS = 256;
constexpr int FMT_SMEM = 40 * 1024;

__global__ void __launch_bounds__(FMT_ROWS) k_fmt_write(TextFmt f, int64_t n, const int64_t* __restrict__ off, char* __restrict__ out) {
  __shared__ __align__(16) char buf[FMT_SMEM];
  const int64_t n_blocks = (n + FMT_ROWS - 1) / FMT_ROWS;
  for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const int64_t r0 = b * FMT_ROWS, r1 = min(n, r0 + (int64_t)FMT_ROWS);
    const int64_t o0 = off[r0], o1 = off[r1];
    const int64_t i = r0 + threadIdx.x;
    const int shift = (int)((uintptr_t)(out + o0) & 15u);
    bool host;
    if (o1 - o0 + shift <= (int64_t)FMT_SMEM) {
      if (i < r1) put_line<true>(f, i, buf + shift + (int)(off[i] - o0), &host);
      __syncthreads();
      const int total = (int)(o1 - o0);
      char* dst = out + o0 - shift;                                             // 16-byte aligned
      const int first = shift, last = shift + total;                            // bytes [first, last) of buf are valid
      const int v0 = (first + 15) & ~15, v1 = last & ~15;                       // whole 16-byte words in between
      for (int k = first + threadIdx.x; k < min(v0, last); k += FMT_ROWS) dst[k] = buf[k];
      for (int k = v0 + 16 * threadIdx.x; k < v1; k += 16 * FMT_ROWS) *(uint4*)(dst + k) = *(const uint4*)(buf + k);
      for (int k = max(v1, v0) + threadIdx.x; k < last; k += FMT_ROWS) dst[k] = buf[k];
      __syncthreads();
    } else if (i < r1) {
      put_line<true>(f, i, out + off[i], &host);
    }
  }
}

}  // namespace

extern "C" int gmql_format_text(gmql_ctx* ctx, const gmql_regions* in, const gmql_text_out_schema* schema, const char* const* chrom_names, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(in && schema && chrom_names, GMQL_ERR_INVALID, "null argument");
    GMQL_REQUIRE(schema->n_values >= 0 && schema->n_values <= 16 && (schema->n_values == 0 || schema->value_cols), GMQL_ERR_INVALID, "0..16 value columns");
    std::vector<int> need;
    for (int k = 0; k < schema->n_values; ++k) {
      GMQL_REQUIRE(schema->value_cols[k] >= 0 && schema->value_cols[k] < in->n_cols, GMQL_ERR_INVALID, "value column index out of range");
      bool seen = false;
      for (int c : need) seen = seen || c == schema->value_cols[k];
      if (!seen) need.push_back(schema->value_cols[k]);
    }
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions d;
    gmql_upload_regions(ctx, in, need, &d);
    // the chromosome dictionary as one blob + offsets
    std::string blob;
    int32_t* h_off = (int32_t*)ctx_pinned(ctx, 4 * ((size_t)d.n_chrom + 1));
    for (int c = 0; c < d.n_chrom; ++c) {
      GMQL_REQUIRE(chrom_names[c] != nullptr, GMQL_ERR_INVALID, "null chromosome name");
      h_off[c] = (int32_t)blob.size();
      blob += chrom_names[c];
    }
    h_off[d.n_chrom] = (int32_t)blob.size();
    char* h_blob = (char*)ctx_pinned(ctx, blob.size() + 1);
    memcpy(h_blob, blob.data(), blob.size());
    DevBuf<char> d_names(ctx, (int64_t)blob.size() + 1);
    DevBuf<int32_t> d_off(ctx, d.n_chrom + 1);
    CUDA_TRY(cudaMemcpyAsync(d_off.p, h_off, 4 * ((size_t)d.n_chrom + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (!blob.empty()) CUDA_TRY(cudaMemcpyAsync(d_names.p, h_blob, blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    TextFmt f;
    memset(&f, 0, sizeof f);
    f.chrom = d.chrom; f.start = d.start; f.stop = d.stop; f.bits = d.coord_bits; f.strand = d.strand;
    f.n_values = schema->n_values;
    for (int k = 0; k < schema->n_values; ++k) { f.v[k] = d.cols[schema->value_cols[k]]; f.ok[k] = d.valid[schema->value_cols[k]]; }
    f.names = d_names.p; f.name_off = d_off.p; f.n_chrom = d.n_chrom; f.one_based = schema->one_based ? 1 : 0;
    const int64_t n = d.n;
    res = new gmql_result();
    res->owner = ctx;
    res->kind = 7;
    DevBuf<int32_t> len(ctx, n);
    DevBuf<int64_t> off(ctx, n + 1);
    DevBuf<uint8_t> status(ctx, n);
    int64_t total = 0;
    if (n) LAUNCH(ctx, k_fmt_len, grid_for(ctx, n, 256), 256, 0, f, n, len.p, status.p);
    device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, len.p, off.p, n, off.p + n);
    CUDA_TRY(cudaMemcpyAsync(&total, off.p + n, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
    DevBuf<uint8_t> bytes(ctx, total);
    if (n) {
      const int64_t nb = ceil_div64(n, FMT_ROWS);
      LAUNCH(ctx, k_fmt_write, (unsigned)std::min<int64_t>(nb, (int64_t)ctx->num_sms * 8), FMT_ROWS, 0, f, n, (const int64_t*)off.p, (char*)bytes.p);
    }
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    res->rows = n;
    result_add(res, GMQL_COL_TXT_BYTES, bytes);
    result_add(res, GMQL_COL_TXT_LINE_OFFSET, off);
    result_add(res, GMQL_COL_TXT_STATUS, status);
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
