// gmql_parse.cu — tab-separated region text (BED / narrowPeak / custom schemas) -> columnar SoA on the device.
//
// Replaces the per-line host work of BedParser.region_parser (GMQL-Spark/.../loaders/BedParser.scala:112-172: split on tabs,
// trim, strand "+"/"-" else '*', optional start - 1, numeric values through String.toDouble with "null"/"." -> GNull :30-48,
// '#' lines skipped :173-203).  count -> scan -> write: newlines are counted and scanned into line offsets, comment / empty
// lines are dropped by a second scan, and one thread parses each kept line.
//
// Doubles are converted with the Eisel-Lemire algorithm (64x128-bit product with a truncated power of five, pow5_table.h),
// which is correctly rounded — as Java's Double.parseDouble is — whenever it answers; the few inputs it cannot decide (more
// than 19 significant digits with an ambiguous tail) and every syntax beyond plain decimals (hex floats, NaN, Infinity,
// type suffixes) are flagged per row for the host's own parser instead of being guessed.
#include "index.cuh"
#include "pow5_table.h"

namespace {

constexpr int TXT_MAX_VALUES = 16;

struct TextPlan {
  int chr_pos, start_pos, stop_pos, strand_pos, start_minus_one;
  int n_values;
  int value_pos[TXT_MAX_VALUES];
  int max_pos;
};

struct ChromDict {
  const unsigned long long* hash;   // [cap] open addressing, 0 = empty
  const int32_t* index;             // [cap]
  const char* names;                // concatenated
  const int32_t* name_off;          // [n_chrom + 1]
  int cap_mask;
};

struct TextOut {
  int64_t* line_off; uint8_t* status; int32_t* chrom; int64_t* start; int64_t* stop; int8_t* strand;
  double* value[TXT_MAX_VALUES]; uint8_t* valid[TXT_MAX_VALUES];
};

__device__ __forceinline__ unsigned long long fnv1a(const char* p, int n) {
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
  return h ? h : 1ull;
}

// ---- newline bookkeeping ----------------------------------------------------------------------------------------------
__global__ void k_txt_count_nl(const char* __restrict__ t, int64_t n, int64_t chunk, int32_t* __restrict__ cnt) {
  // one thread per chunk of `chunk` bytes
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c * chunk < n; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t a = c * chunk, b = a + chunk < n ? a + chunk : n;
    int k = 0;
    for (int64_t i = a; i < b; ++i) k += t[i] == '\n';
    cnt[c] = k;
  }
}
// line k+1 starts after the k-th newline; line 0 starts at 0
__global__ void k_txt_line_starts(const char* __restrict__ t, int64_t n, int64_t chunk, const int64_t* __restrict__ cnt_excl, int64_t* __restrict__ starts) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c * chunk < n; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t a = c * chunk, b = a + chunk < n ? a + chunk : n;
    int64_t o = cnt_excl[c] + 1;
    if (c == 0) starts[0] = 0;
    for (int64_t i = a; i < b; ++i) if (t[i] == '\n') starts[o++] = i + 1;
  }
}
// ---- field parsers ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void trim(const char* t, int64_t& a, int64_t& b) {      // String.trim: drop chars <= ' '
  while (a < b && (unsigned char)t[a] <= ' ') ++a;
  while (b > a && (unsigned char)t[b - 1] <= ' ') --b;
}

__device__ bool parse_long(const char* t, int64_t a, int64_t b, int64_t* out) {
  trim(t, a, b);
  if (a >= b) return false;
  bool neg = false;
  if (t[a] == '+' || t[a] == '-') { neg = t[a] == '-'; ++a; }
  if (a >= b || b - a > 18) return false;
  long long v = 0;
  for (int64_t i = a; i < b; ++i) {
    unsigned d = (unsigned)(t[i] - '0');
    if (d > 9u) return false;
    v = v * 10 + (long long)d;
  }
  *out = neg ? -v : v;
  return true;
}

struct AdjMant { unsigned long long mant; int power2; };

__device__ AdjMant el_compute(long long q, unsigned long long w, const unsigned long long* __restrict__ pow5) {
  AdjMant r;
  if (w == 0 || q < GMQL_POW5_SMALLEST_Q) { r.mant = 0; r.power2 = 0; return r; }
  if (q > GMQL_POW5_LARGEST_Q) { r.mant = 0; r.power2 = 0x7FF; return r; }
  const int lz = __clzll((long long)w);
  w <<= lz;
  const int idx = 2 * (int)(q - GMQL_POW5_SMALLEST_Q);
  const unsigned long long p_hi = pow5[idx], p_lo = pow5[idx + 1];
  unsigned long long fhi = __umul64hi(w, p_hi), flo = w * p_hi;
  const unsigned long long precision_mask = 0xFFFFFFFFFFFFFFFFull >> 55;
  if ((fhi & precision_mask) == precision_mask) {
    const unsigned long long shi = __umul64hi(w, p_lo);
    flo += shi;
    if (shi > flo) ++fhi;
  }
  const int upperbit = (int)(fhi >> 63);
  const int shift = upperbit + 64 - 52 - 3;
  unsigned long long mant = fhi >> shift;
  int power2 = (int)((((152170ll + 65536ll) * q) >> 16) + 63) + upperbit - lz + 1023;
  if (power2 <= 0) {                                       // subnormal
    if (-power2 + 1 >= 64) { r.mant = 0; r.power2 = 0; return r; }
    mant >>= -power2 + 1;
    mant += mant & 1ull;
    mant >>= 1;
    r.mant = mant; r.power2 = mant < (1ull << 52) ? 0 : 1;
    return r;
  }
  if (flo <= 1ull && q >= -4 && q <= 23 && (mant & 3ull) == 1ull) {   // exactly halfway: round to even
    if ((mant << shift) == fhi) mant &= ~1ull;
  }
  mant += mant & 1ull;
  mant >>= 1;
  if (mant >= (2ull << 52)) { mant = 1ull << 52; ++power2; }
  mant &= ~(1ull << 52);
  if (power2 >= 0x7FF) { r.mant = 0; r.power2 = 0x7FF; return r; }
  r.mant = mant; r.power2 = power2;
  return r;
}

// 0: parsed; 1: GNull ("null" / "."); 2: leave to the host parser
__device__ int parse_value(const char* t, int64_t a, int64_t b, const unsigned long long* __restrict__ pow5, double* out) {
  // BedParser.parseRegion (:30-48) tests the UNtrimmed field against "null" / "." (case-insensitive), then String.toDouble trims
  if (b - a == 1 && t[a] == '.') return 1;
  if (b - a == 4 && (t[a] | 32) == 'n' && (t[a + 1] | 32) == 'u' && (t[a + 2] | 32) == 'l' && (t[a + 3] | 32) == 'l') return 1;
  trim(t, a, b);
  if (a >= b) return 2;
  bool neg = false;
  int64_t i = a;
  if (t[i] == '+' || t[i] == '-') { neg = t[i] == '-'; ++i; }
  unsigned long long w = 0;
  int nd = 0;
  long long q = 0;
  bool seen = false, trunc = false;
  for (; i < b; ++i) {
    unsigned d = (unsigned)(t[i] - '0');
    if (d > 9u) break;
    seen = true;
    if (nd < 19) { w = w * 10ull + d; nd += w != 0ull; } else { ++q; trunc = trunc || d != 0u; }
  }
  if (i < b && t[i] == '.') {
    ++i;
    for (; i < b; ++i) {
      unsigned d = (unsigned)(t[i] - '0');
      if (d > 9u) break;
      seen = true;
      if (nd < 19) { w = w * 10ull + d; nd += w != 0ull; --q; } else trunc = trunc || d != 0u;
    }
  }
  if (!seen) return 2;
  if (i < b && (t[i] == 'e' || t[i] == 'E')) {
    ++i;
    bool eneg = false;
    if (i < b && (t[i] == '+' || t[i] == '-')) { eneg = t[i] == '-'; ++i; }
    if (i >= b || (unsigned)(t[i] - '0') > 9u) return 2;
    long long e = 0;
    for (; i < b; ++i) {
      unsigned d = (unsigned)(t[i] - '0');
      if (d > 9u) break;
      if (e < 100000) e = e * 10 + (long long)d;
    }
    q += eneg ? -e : e;
  }
  if (i != b) return 2;                                    // suffixes, hex floats, NaN, Infinity: the host decides
  AdjMant am = el_compute(q, w, pow5);
  if (trunc) {
    AdjMant am2 = el_compute(q, w + 1ull, pow5);
    if (am.mant != am2.mant || am.power2 != am2.power2) return 2;
  }
  unsigned long long bits = ((unsigned long long)neg << 63) | ((unsigned long long)am.power2 << 52) | am.mant;
  *out = __longlong_as_double((long long)bits);
  return 0;
}

__device__ int lookup_chrom(const ChromDict& d, const char* t, int64_t a, int64_t b) {
  trim(t, a, b);
  const int n = (int)(b - a);
  unsigned long long h = fnv1a(t + a, n);
  for (int s = (int)(h & (unsigned long long)d.cap_mask);; s = (s + 1) & d.cap_mask) {
    unsigned long long e = d.hash[s];
    if (e == 0ull) return -1;
    if (e == h) {
      int idx = d.index[s];
      int o = d.name_off[idx], len = d.name_off[idx + 1] - o;
      bool same = len == n;
      for (int i = 0; same && i < n; ++i) same = d.names[o + i] == t[a + i];
      if (same) return idx;
    }
  }
}

struct ParsedLine {
  int64_t start, stop; int chrom; int8_t strand; uint8_t status;
  double value[TXT_MAX_VALUES]; uint8_t valid[TXT_MAX_VALUES];
};

// one line [a, b) (no newline / carriage return) -> fields.  status: bit 0 a numeric field is left to the host, bit 1
// malformed (the reference throws: missing column, bad integer), bit 2 chromosome not in the dictionary
__device__ void parse_line(const char* __restrict__ t, int64_t a, int64_t b, const TextPlan& plan, const ChromDict& dict, const unsigned long long* __restrict__ pow5, ParsedLine* pl) {
  int64_t f_chr[2] = {-1, -1}, f_start[2] = {-1, -1}, f_stop[2] = {-1, -1}, f_strand[2] = {-1, -1};
  int64_t f_val[TXT_MAX_VALUES][2];
  for (int k = 0; k < plan.n_values; ++k) { f_val[k][0] = -1; f_val[k][1] = -1; }
  int col = 0;
  int64_t tok = a;
  for (int64_t i = a; i <= b && col <= plan.max_pos; ++i) {      // split(delimiter, -1): trailing empty fields count
    if (i == b || t[i] == '\t') {
      if (col == plan.chr_pos) { f_chr[0] = tok; f_chr[1] = i; }
      if (col == plan.start_pos) { f_start[0] = tok; f_start[1] = i; }
      if (col == plan.stop_pos) { f_stop[0] = tok; f_stop[1] = i; }
      if (col == plan.strand_pos) { f_strand[0] = tok; f_strand[1] = i; }
      for (int k = 0; k < plan.n_values; ++k) if (col == plan.value_pos[k]) { f_val[k][0] = tok; f_val[k][1] = i; }
      ++col;
      tok = i + 1;
    }
  }
  uint8_t status = 0;
  int64_t s = 0, e = 0;
  int chrom = -1;
  int8_t strand = 0;
  if (f_chr[0] < 0 || f_start[0] < 0 || f_stop[0] < 0 || (plan.strand_pos >= 0 && f_strand[0] < 0)) status |= 2;   // missing columns
  else {
    if (!parse_long(t, f_start[0], f_start[1], &s) || !parse_long(t, f_stop[0], f_stop[1], &e)) status |= 2;
    if (plan.start_minus_one) s -= 1;                                                            // 1-based schemas (:164)
    chrom = lookup_chrom(dict, t, f_chr[0], f_chr[1]);
    if (chrom < 0) status |= 4;
    if (plan.strand_pos >= 0) {
      int64_t x = f_strand[0], y = f_strand[1];
      trim(t, x, y);
      strand = (y - x == 1 && t[x] == '+') ? 1 : (y - x == 1 && t[x] == '-') ? 2 : 0;            // anything else -> '*' (:166-170)
    }
  }
  for (int k = 0; k < plan.n_values; ++k) {
    double v = 0.0;
    uint8_t ok = 0;
    if (f_val[k][0] < 0) status |= 2;
    else {
      int r = parse_value(t, f_val[k][0], f_val[k][1], pow5, &v);
      if (r == 0) ok = 1;
      else if (r == 2) status |= 1;
    }
    pl->value[k] = v;
    pl->valid[k] = ok;
  }
  pl->start = s; pl->stop = e; pl->chrom = chrom; pl->strand = strand; pl->status = status;
}

__device__ __forceinline__ void line_span(const char* __restrict__ t, int64_t n, const int64_t* __restrict__ starts, int64_t n_nl, int64_t l, int64_t* a, int64_t* b) {
  *a = starts[l];
  *b = l < n_nl ? starts[l + 1] - 1 : n;                   // without the newline
  if (*b > *a && t[*b - 1] == '\r') --*b;
}

// A line that starts with '#' is dropped only when it does not parse (region_parser catches the exception and returns
// None for '#' lines, :173-203); every other line is a row — malformed ones included, so that the caller can raise as the
// reference does.
__global__ void __launch_bounds__(128) k_txt_keep(const char* __restrict__ t, int64_t n, const int64_t* __restrict__ starts, int64_t n_nl, int64_t n_lines, TextPlan plan, ChromDict dict,
                                                  const unsigned long long* __restrict__ pow5, int32_t* __restrict__ keep) {
  for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_lines; l += (int64_t)gridDim.x * blockDim.x) {
    int64_t a, b;
    line_span(t, n, starts, n_nl, l, &a, &b);
    int k = 1;
    if (b > a && t[a] == '#') {
      ParsedLine pl;
      parse_line(t, a, b, plan, dict, pow5, &pl);
      k = (pl.status & 2) ? 0 : 1;
    }
    keep[l] = k;
  }
}

__global__ void __launch_bounds__(128) k_txt_parse(const char* __restrict__ t, int64_t n, const int64_t* __restrict__ starts, int64_t n_nl, int64_t n_lines, const int32_t* __restrict__ keep,
                                                   const int64_t* __restrict__ keep_excl, TextPlan plan, ChromDict dict, const unsigned long long* __restrict__ pow5, TextOut out) {
  for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_lines; l += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[l]) continue;
    const int64_t o = keep_excl[l];
    int64_t a, b;
    line_span(t, n, starts, n_nl, l, &a, &b);
    ParsedLine pl;
    parse_line(t, a, b, plan, dict, pow5, &pl);
    for (int k = 0; k < plan.n_values; ++k) { out.value[k][o] = pl.value[k]; out.valid[k][o] = pl.valid[k]; }
    out.line_off[o] = a; out.status[o] = pl.status; out.chrom[o] = pl.chrom; out.start[o] = pl.start; out.stop[o] = pl.stop; out.strand[o] = pl.strand;
  }
}

}  // namespace

extern "C" int gmql_parse_text(gmql_ctx* ctx, const char* text, int64_t n_bytes, int32_t location, const gmql_text_schema* schema,
                               const char* const* chrom_names, int32_t n_chrom, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(schema && (n_bytes == 0 || text) && n_bytes >= 0, GMQL_ERR_INVALID, "null text or schema");
    GMQL_REQUIRE(schema->n_values >= 0 && schema->n_values <= TXT_MAX_VALUES, GMQL_ERR_UNSUPPORTED, "at most 16 numeric value columns per parse");
    GMQL_REQUIRE(schema->chr_pos >= 0 && schema->start_pos >= 0 && schema->stop_pos >= 0, GMQL_ERR_INVALID, "chr/start/stop positions are required");
    GMQL_REQUIRE(n_chrom >= 0 && (n_chrom == 0 || chrom_names), GMQL_ERR_INVALID, "bad chromosome dictionary");
    TextPlan plan;
    plan.chr_pos = schema->chr_pos; plan.start_pos = schema->start_pos; plan.stop_pos = schema->stop_pos;
    plan.strand_pos = schema->strand_pos; plan.start_minus_one = schema->start_minus_one; plan.n_values = schema->n_values;
    plan.max_pos = std::max(std::max(plan.chr_pos, plan.start_pos), std::max(plan.stop_pos, plan.strand_pos));
    for (int k = 0; k < plan.n_values; ++k) {
      GMQL_REQUIRE(schema->value_pos && schema->value_pos[k] >= 0, GMQL_ERR_INVALID, "bad value column position");
      plan.value_pos[k] = schema->value_pos[k];
      plan.max_pos = std::max(plan.max_pos, plan.value_pos[k]);
    }
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    // text -> device
    DevBuf<char> dtext;
    const char* t = text;
    if (location != GMQL_LOC_DEVICE && n_bytes) {
      dtext.alloc(ctx, n_bytes);
      CUDA_TRY(cudaMemcpyAsync(dtext.p, text, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->stream));
      t = dtext.p;
    }
    // chromosome dictionary: open-addressing table of FNV-1a hashes + the names for the exact comparison
    int cap = 16;
    while (cap < 4 * std::max(n_chrom, 1)) cap <<= 1;
    std::vector<unsigned long long> h_hash((size_t)cap, 0ull);
    std::vector<int32_t> h_index((size_t)cap, -1), h_off((size_t)n_chrom + 1, 0);
    std::string blob;
    for (int c = 0; c < n_chrom; ++c) {
      GMQL_REQUIRE(chrom_names[c] != nullptr, GMQL_ERR_INVALID, "null chromosome name");
      std::string nm(chrom_names[c]);
      unsigned long long h = 1469598103934665603ull;
      for (unsigned char ch : nm) { h ^= ch; h *= 1099511628211ull; }
      if (!h) h = 1ull;
      int s = (int)(h & (unsigned long long)(cap - 1));
      while (h_hash[s] != 0ull) s = (s + 1) & (cap - 1);
      h_hash[s] = h; h_index[s] = c;
      h_off[c] = (int32_t)blob.size();
      blob += nm;
    }
    h_off[n_chrom] = (int32_t)blob.size();
    DevBuf<unsigned long long> d_hash(ctx, cap), d_pow5(ctx, 2 * (GMQL_POW5_LARGEST_Q - GMQL_POW5_SMALLEST_Q + 1));
    DevBuf<int32_t> d_index(ctx, cap), d_off(ctx, n_chrom + 1);
    DevBuf<char> d_names(ctx, (int64_t)blob.size() + 1);
    CUDA_TRY(cudaMemcpyAsync(d_hash.p, h_hash.data(), 8 * (size_t)cap, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_index.p, h_index.data(), 4 * (size_t)cap, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_off.p, h_off.data(), 4 * ((size_t)n_chrom + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (!blob.empty()) CUDA_TRY(cudaMemcpyAsync(d_names.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_pow5.p, gmql_pow5_128_host, sizeof gmql_pow5_128_host, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));   // the host staging vectors above go out of scope at the end of this call; keep it simple
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    cudaEventRecord(ctx->ev_k0, ctx->stream);

    res = new gmql_result();
    res->owner = ctx;
    res->kind = 6;
    int64_t n_rows = 0;
    if (n_bytes > 0) {
      // lines: count newlines per chunk -> scan -> starts
      const int64_t chunk = 256;
      const int64_t n_chunks = ceil_div64(n_bytes, chunk);
      DevBuf<int32_t> cnt(ctx, n_chunks);
      DevBuf<int64_t> cnt_excl(ctx, n_chunks + 1);
      LAUNCH(ctx, k_txt_count_nl, grid_for(ctx, n_chunks, 128, 1), 128, 0, t, n_bytes, chunk, cnt.p);
      device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, cnt.p, cnt_excl.p, n_chunks, cnt_excl.p + n_chunks);
      int64_t n_nl = 0;
      char last = 0;
      CUDA_TRY(cudaMemcpyAsync(&n_nl, cnt_excl.p + n_chunks, sizeof n_nl, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(&last, t + n_bytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(GMQL_SYNC(ctx));
      const int64_t n_lines = n_nl + (last == '\n' ? 0 : 1);     // a trailing newline does not open another line
      DevBuf<int64_t> starts(ctx, n_nl + 1);
      LAUNCH(ctx, k_txt_line_starts, grid_for(ctx, n_chunks, 128, 1), 128, 0, t, n_bytes, chunk, cnt_excl.p, starts.p);
      ChromDict cd;
      cd.hash = d_hash.p; cd.index = d_index.p; cd.names = d_names.p; cd.name_off = d_off.p; cd.cap_mask = cap - 1;
      DevBuf<int32_t> keep(ctx, n_lines);
      DevBuf<int64_t> keep_excl(ctx, n_lines + 1);
      if (n_lines) LAUNCH(ctx, k_txt_keep, grid_for(ctx, n_lines, 128, 1), 128, 0, t, n_bytes, starts.p, n_nl, n_lines, plan, cd, d_pow5.p, keep.p);
      device_scan<int32_t, int64_t, OpSum<int64_t>, false>(ctx, keep.p, keep_excl.p, n_lines, keep_excl.p + n_lines);
      CUDA_TRY(cudaMemcpyAsync(&n_rows, keep_excl.p + n_lines, sizeof n_rows, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(GMQL_SYNC(ctx));
      DevBuf<int64_t> o_off(ctx, n_rows), o_start(ctx, n_rows), o_stop(ctx, n_rows);
      DevBuf<uint8_t> o_status(ctx, n_rows);
      DevBuf<int32_t> o_chrom(ctx, n_rows);
      DevBuf<int8_t> o_strand(ctx, n_rows);
      std::vector<DevBuf<double>> ov((size_t)plan.n_values);
      std::vector<DevBuf<uint8_t>> ok((size_t)plan.n_values);
      TextOut to;
      memset(&to, 0, sizeof to);
      to.line_off = o_off.p; to.status = o_status.p; to.chrom = o_chrom.p; to.start = o_start.p; to.stop = o_stop.p; to.strand = o_strand.p;
      for (int k = 0; k < plan.n_values; ++k) { ov[k].alloc(ctx, n_rows); ok[k].alloc(ctx, n_rows); to.value[k] = ov[k].p; to.valid[k] = ok[k].p; }
      if (n_rows) LAUNCH(ctx, k_txt_parse, grid_for(ctx, n_lines, 128, 1), 128, 0, t, n_bytes, starts.p, n_nl, n_lines, keep.p, keep_excl.p, plan, cd, d_pow5.p, to);
      cudaEventRecord(ctx->ev_k1, ctx->stream);
      result_add(res, GMQL_COL_TXT_LINE_OFFSET, o_off);
      result_add(res, GMQL_COL_TXT_STATUS, o_status);
      result_add(res, GMQL_COL_TXT_CHROM, o_chrom);
      result_add(res, GMQL_COL_TXT_START, o_start);
      result_add(res, GMQL_COL_TXT_STOP, o_stop);
      result_add(res, GMQL_COL_TXT_STRAND, o_strand);
      for (int k = 0; k < plan.n_values; ++k) { result_add(res, GMQL_COL_AGG_VALUE + k, ov[k]); result_add(res, GMQL_COL_AGG_VALID + k, ok[k]); }
    } else {
      cudaEventRecord(ctx->ev_k1, ctx->stream);
    }
    res->rows = n_rows;
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
