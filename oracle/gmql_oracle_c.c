/*
 * CPU ORACLE (C) — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A C/OpenMP port of the reference's bin-based algorithms for MAP and the COVER family, used (a) as the checker at
 * sizes the pure-Python literal oracle (oracle/gmql_oracle.py) cannot finish, and (b) as bench.py's `cpu_baseline`
 * ("port") and `--impl reference` arm.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors for this path and cannot be built here (no JVM/Scala/Spark);
 * this port is pinned against the literal Python restatement on randomised inputs (tests/test_oracle_c.py), which in
 * turn follows the Scala line by line.
 *
 * What is kept from the reference's algorithm (paths relative to /root/reference, RO = GMQL-Spark/src/main/scala/it/
 * polimi/genomics/spark/implementation/RegionsOperators):
 *   MAP    RO/GenometricMap/GenometricMap71.scala:49-150 — experiment and reference records are replicated into
 *          fixed-width bins (:153-181), co-grouped on (exp sample, chrom, bin) (:82-83), every reference record of a
 *          bin scans every experiment record of the bin (:90-107) with the first-bin rule (:95,106), partial results
 *          are reduced per (pair, ref region) (:123).  Spark's shuffles become sorts; the reduce is in place.
 *   COVER  RO/GenometricCover/GenometricCover.scala:30-158 — per-bin +1/-1 point maps (extractor :345-360) merged per
 *          (chrom, bin, group, strand) (:98-108), sorted per bin (:111), coverHelper / histogramHelper / summitHelper
 *          state machines per bin, split + boundary merge (:121-152), then
 *          RO/GenometricMap/GMAP4.scala:20-100 — binned join of the cover regions against the inputs, reduce per region.
 * Parallelism: OpenMP over independent keys — experiment samples for MAP; for COVER chunks of 1024 consecutive 2000-bp bins of a
 * (group, strand, chrom) segment in phase 1 (Spark keys that plan by (chrom, bin, group, strand), :51,98-108) and phase-1
 * regions in GMAP4 — the analogue of Spark's partitions in local[N].
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int strand_ok(int a, int b) { return a == 0 || b == 0 || a == b; }

static double java_min(double a, double b) {
  if (a != a) return a;
  if (a == 0.0 && b == 0.0 && signbit(b)) return b;
  return a <= b ? a : b;
}
static double java_max(double a, double b) {
  if (a != a) return a;
  if (a == 0.0 && b == 0.0 && signbit(a)) return b;
  return a >= b ? a : b;
}

/* bench.py: use every host core whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers) */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------------ MAP */
typedef struct { int32_t chrom; int64_t bin; int32_t idx; } BinRec;

static int cmp_binrec(const void* a, const void* b) {
  const BinRec* x = (const BinRec*)a; const BinRec* y = (const BinRec*)b;
  if (x->chrom != y->chrom) return x->chrom < y->chrom ? -1 : 1;
  if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/*
 * Output layout = the CUDA result's: rows grouped by pair in list order; inside pair (a, b) sample a's regions in
 * ascending row index.  out_count[n_rows]; for the single aggregated column (val may be NULL): out_nn, out_sum,
 * out_min, out_max [n_rows] (min/max untouched where nn == 0).  Duplicate reference records are NOT collapsed here
 * (the caller's generators never produce them; the Python literal oracle covers that rule).
 * Returns n_rows, or -1 on allocation failure.
 */
int64_t oracle_map(int64_t n_ref, const int32_t* r_chrom, const int64_t* r_start, const int64_t* r_stop, const int8_t* r_strand, const int32_t* r_sample, int32_t n_ref_samples,
                   int64_t n_exp, const int32_t* e_chrom, const int64_t* e_start, const int64_t* e_stop, const int8_t* e_strand, const int32_t* e_sample, int32_t n_exp_samples,
                   const double* e_val, const uint8_t* e_valid,
                   int64_t n_pairs, const int32_t* pair_l, const int32_t* pair_r, int64_t bin,
                   int32_t* out_count, int32_t* out_nn, double* out_sum, double* out_min, double* out_max) {
  if (bin == 0) bin = INT64_MAX;                                         /* GenometricMap71.scala:36-40 */
  /* ref rows grouped by sample, ascending row index */
  int64_t* s_off = (int64_t*)calloc((size_t)n_ref_samples + 1, sizeof(int64_t));
  int32_t* s_rows = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_ref ? n_ref : 1));
  int32_t* rank = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_ref ? n_ref : 1));
  int64_t* row_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_pairs + 1));
  if (!s_off || !s_rows || !rank || !row_off) return -1;
  for (int64_t i = 0; i < n_ref; ++i) s_off[r_sample[i] + 1]++;
  for (int32_t s = 0; s < n_ref_samples; ++s) s_off[s + 1] += s_off[s];
  {
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_ref_samples + 1));
    memcpy(fill, s_off, sizeof(int64_t) * (size_t)n_ref_samples);
    for (int64_t i = 0; i < n_ref; ++i) { int64_t p = fill[r_sample[i]]++; s_rows[p] = (int32_t)i; rank[i] = (int32_t)(p - s_off[r_sample[i]]); }
    free(fill);
  }
  row_off[0] = 0;
  for (int64_t p = 0; p < n_pairs; ++p) row_off[p + 1] = row_off[p] + (s_off[pair_l[p] + 1] - s_off[pair_l[p]]);
  int64_t n_rows = row_off[n_pairs];
  for (int64_t r = 0; r < n_rows; ++r) { out_count[r] = 0; if (out_nn) { out_nn[r] = 0; out_sum[r] = 0.0; } }
  /* exp rows grouped by sample; pairs grouped by exp sample */
  int64_t* e_off = (int64_t*)calloc((size_t)n_exp_samples + 1, sizeof(int64_t));
  int32_t* e_rows = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_exp ? n_exp : 1));
  int64_t* p_off = (int64_t*)calloc((size_t)n_exp_samples + 1, sizeof(int64_t));
  int64_t* p_list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_pairs ? n_pairs : 1));
  if (!e_off || !e_rows || !p_off || !p_list) return -1;
  for (int64_t i = 0; i < n_exp; ++i) e_off[e_sample[i] + 1]++;
  for (int32_t s = 0; s < n_exp_samples; ++s) e_off[s + 1] += e_off[s];
  {
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_exp_samples + 1));
    memcpy(fill, e_off, sizeof(int64_t) * (size_t)n_exp_samples);
    for (int64_t i = 0; i < n_exp; ++i) e_rows[fill[e_sample[i]]++] = (int32_t)i;
    for (int64_t p = 0; p < n_pairs; ++p) p_off[pair_r[p] + 1]++;
    for (int32_t s = 0; s < n_exp_samples; ++s) p_off[s + 1] += p_off[s];
    memcpy(fill, p_off, sizeof(int64_t) * (size_t)n_exp_samples);
    for (int64_t p = 0; p < n_pairs; ++p) p_list[fill[pair_r[p]]++] = p;
    free(fill);
  }
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t b = 0; b < n_exp_samples; ++b) {
    if (p_off[b + 1] == p_off[b]) continue;
    /* exp.binDS (:153-165): one record per bin start/bin .. stop/bin */
    int64_t ne = 0;
    for (int64_t q = e_off[b]; q < e_off[b + 1]; ++q) { int32_t i = e_rows[q]; ne += e_stop[i] / bin - e_start[i] / bin + 1; }
    BinRec* eb = (BinRec*)malloc(sizeof(BinRec) * (size_t)(ne ? ne : 1));
    if (!eb) { failed = 1; continue; }
    int64_t w = 0;
    for (int64_t q = e_off[b]; q < e_off[b + 1]; ++q) {
      int32_t i = e_rows[q];
      for (int64_t k = e_start[i] / bin; k <= e_stop[i] / bin; ++k) { eb[w].chrom = e_chrom[i]; eb[w].bin = k; eb[w].idx = i; ++w; }
    }
    qsort(eb, (size_t)ne, sizeof(BinRec), cmp_binrec);
    for (int64_t pp = p_off[b]; pp < p_off[b + 1]; ++pp) {
      int64_t p = p_list[pp];
      int32_t a = pair_l[p];
      /* ref.binDS (:167-181): the reference sample is replicated for this experiment sample, one record per bin */
      int64_t nr = 0;
      for (int64_t q = s_off[a]; q < s_off[a + 1]; ++q) { int32_t i = s_rows[q]; nr += r_stop[i] / bin - r_start[i] / bin + 1; }
      BinRec* rb = (BinRec*)malloc(sizeof(BinRec) * (size_t)(nr ? nr : 1));
      if (!rb) { failed = 1; continue; }
      int64_t v = 0;
      for (int64_t q = s_off[a]; q < s_off[a + 1]; ++q) {
        int32_t i = s_rows[q];
        for (int64_t k = r_start[i] / bin; k <= r_stop[i] / bin; ++k) { rb[v].chrom = r_chrom[i]; rb[v].bin = k; rb[v].idx = i; ++v; }
      }
      qsort(rb, (size_t)nr, sizeof(BinRec), cmp_binrec);
      /* cogroup (:82-83) + hot loop (:90-120) */
      int64_t ei = 0;
      for (int64_t ri = 0; ri < nr;) {
        int64_t rj = ri;
        while (rj < nr && rb[rj].chrom == rb[ri].chrom && rb[rj].bin == rb[ri].bin) ++rj;
        while (ei < ne && (eb[ei].chrom < rb[ri].chrom || (eb[ei].chrom == rb[ri].chrom && eb[ei].bin < rb[ri].bin))) ++ei;
        int64_t ej = ei;
        while (ej < ne && eb[ej].chrom == rb[ri].chrom && eb[ej].bin == rb[ri].bin) ++ej;
        int64_t key_bin = rb[ri].bin;
        for (int64_t x = ri; x < rj; ++x) {
          int32_t r = rb[x].idx;
          int ref_in_start_bin = (r_start[r] / bin) == key_bin;
          int64_t row = row_off[p] + rank[r];
          for (int64_t y = ei; y < ej; ++y) {
            int32_t e = eb[y].idx;
            if (r_start[r] < e_stop[e] && e_start[e] < r_stop[r] && strand_ok(r_strand ? r_strand[r] : 0, e_strand ? e_strand[e] : 0) &&
                (ref_in_start_bin || (e_start[e] / bin) == key_bin)) {
              out_count[row]++;                                         /* reduceFunc (:66-79), in place */
              if (out_nn && e_val && (!e_valid || e_valid[e])) {
                double val = e_val[e];
                if (out_nn[row] == 0) { out_sum[row] = val; out_min[row] = val; out_max[row] = val; }
                else { out_sum[row] += val; out_min[row] = java_min(out_min[row], val); out_max[row] = java_max(out_max[row], val); }
                out_nn[row]++;
              }
            }
          }
        }
        ri = rj;
      }
      free(rb);
    }
    free(eb);
  }
  free(s_off); free(s_rows); free(rank); free(row_off); free(e_off); free(e_rows); free(p_off); free(p_list);
  return failed ? -1 : n_rows;
}

/* ---------------------------------------------------------------------------------------------------------- COVER */
typedef struct { int64_t bin; int32_t off; int32_t delta; } Pt;
static int cmp_pt(const void* a, const void* b) {
  const Pt* x = (const Pt*)a; const Pt* y = (const Pt*)b;
  if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
  return x->off < y->off ? -1 : (x->off > y->off);
}
typedef struct { int64_t start, stop; int32_t acc; } Piece;
static int cmp_piece(const void* a, const void* b) {
  const Piece* x = (const Piece*)a; const Piece* y = (const Piece*)b;
  if (x->start != y->start) return x->start < y->start ? -1 : 1;
  return x->stop < y->stop ? -1 : (x->stop > y->stop);
}
typedef struct { Piece* p; int64_t n, cap; } PieceVec;
static int pv_push(PieceVec* v, int64_t s, int64_t e, int32_t acc) {
  if (v->n == v->cap) {
    int64_t nc = v->cap ? v->cap * 2 : 1024;
    Piece* q = (Piece*)realloc(v->p, sizeof(Piece) * (size_t)nc);
    if (!q) return -1;
    v->p = q; v->cap = nc;
  }
  v->p[v->n].start = s; v->p[v->n].stop = e; v->p[v->n].acc = acc; v->n++;
  return 0;
}

typedef struct {
  int32_t group, chrom; int8_t strand; int64_t start, stop; int32_t acc; double j1, j2;
  int32_t count, nn; double sum, vmin, vmax;
} CoverRow;

typedef struct { CoverRow* p; int64_t n, cap; } RowVec;

/* ---- phase 1 of one chunk of consecutive bins [b_lo, b_hi) of one (group, strand, chrom) segment ------------------------------
 * Spark keys the COVER plan by (chrom, bin, group, strand) (GenometricCover.scala:51,98-108): every bin is an independent unit
 * of work until the boundary merge.  The port keeps that granularity for its parallelism: a segment is cut into chunks of
 * CHUNK_BINS bins, a region is handed to every chunk it touches, and a chunk generates, sorts and sweeps only the events of its
 * own bins.  It also builds the chunk's share of GMAP4's binned input (GMAP4.scala:102-139). */
#define CHUNK_BINS 1024

typedef struct {
  int64_t b_lo, b_hi;                /* bins of this chunk */
  PieceVec pieces;                   /* phase-1 pieces, in bin order */
  BinRec* eb; int64_t nb;            /* GMAP4: (bin, region) records of the chunk's bins, sorted by bin */
  int rc;
} Chunk;

static void cover_chunk(int flag, int64_t B, int32_t mn, int32_t mx, const int32_t* idx, int64_t m, const int64_t* start, const int64_t* stop, Chunk* ck) {
  const int64_t b_lo = ck->b_lo, b_hi = ck->b_hi;
  /* extractor (:345-360), restricted to the bins of this chunk */
  int64_t ne = 0, nb = 0;
  for (int64_t q = 0; q < m; ++q) {
    int64_t s = start[idx[q]], e = stop[idx[q]];
    int64_t sb = s / B, eb = e / B;
    int64_t lo = sb > b_lo ? sb : b_lo, hi = eb < b_hi - 1 ? eb : b_hi - 1;
    if (hi >= lo) { ne += 2 * (hi - lo + 1); nb += hi - lo + 1; }
  }
  Pt* ev = (Pt*)malloc(sizeof(Pt) * (size_t)(ne ? ne : 1));
  ck->eb = (BinRec*)malloc(sizeof(BinRec) * (size_t)(nb ? nb : 1));
  if (!ev || !ck->eb) { free(ev); ck->rc = -1; return; }
  int64_t w = 0, v = 0;
  for (int64_t q = 0; q < m; ++q) {
    int64_t s = start[idx[q]], e = stop[idx[q]];
    int64_t sb = s / B, eb = e / B;
    int64_t lo = sb > b_lo ? sb : b_lo, hi = eb < b_hi - 1 ? eb : b_hi - 1;
    for (int64_t k = lo; k <= hi; ++k) { ck->eb[v].chrom = 0; ck->eb[v].bin = k; ck->eb[v].idx = idx[q]; ++v; }
    if (sb == eb) {
      if (sb < b_lo || sb >= b_hi) continue;
      if (s == e) { ev[w].bin = sb; ev[w].off = (int32_t)(s - sb * B); ev[w].delta = -1; ++w; }   /* duplicate HashMap key keeps the -1 (:351) */
      else {
        ev[w].bin = sb; ev[w].off = (int32_t)(s - sb * B); ev[w].delta = 1; ++w;
        ev[w].bin = sb; ev[w].off = (int32_t)(e - sb * B); ev[w].delta = -1; ++w;
      }
    } else {
      if (sb >= b_lo && sb < b_hi) {                                                              /* map_start (:353) */
        ev[w].bin = sb; ev[w].off = (int32_t)(s - sb * B); ev[w].delta = 1; ++w;
        ev[w].bin = sb; ev[w].off = (int32_t)B; ev[w].delta = -1; ++w;
      }
      if (eb >= b_lo && eb < b_hi) {                                                              /* map_stop (:354) */
        int64_t eo = e - eb * B;
        ev[w].bin = eb; ev[w].off = (int32_t)(eo != 0 ? eo : 1); ev[w].delta = -1; ++w;
        ev[w].bin = eb; ev[w].off = 0; ev[w].delta = 1; ++w;
      }
      int64_t klo = sb + 1 > b_lo ? sb + 1 : b_lo, khi = eb - 1 < b_hi - 1 ? eb - 1 : b_hi - 1;
      for (int64_t k = klo; k <= khi; ++k) {                                                      /* map_int (:355) */
        ev[w].bin = k; ev[w].off = 0; ev[w].delta = 1; ++w;
        ev[w].bin = k; ev[w].off = (int32_t)B; ev[w].delta = -1; ++w;
      }
    }
  }
  ne = w; ck->nb = v;
  qsort(ck->eb, (size_t)ck->nb, sizeof(BinRec), cmp_binrec);
  qsort(ev, (size_t)ne, sizeof(Pt), cmp_pt);
  /* reduceByKey: coincident points are summed, zero-net points stay (:98-108) */
  int64_t np = 0;
  for (int64_t q = 0; q < ne;) {
    int64_t r = q; int32_t d = 0;
    while (r < ne && ev[r].bin == ev[q].bin && ev[r].off == ev[q].off) { d += ev[r].delta; ++r; }
    ev[np].bin = ev[q].bin; ev[np].off = ev[q].off; ev[np].delta = d; ++np;
    q = r;
  }
  PieceVec* pv = &ck->pieces;
  int rc = 0;
  for (int64_t q = 0; q < np && rc == 0;) {
    int64_t r = q;
    while (r < np && ev[r].bin == ev[q].bin) ++r;
    int64_t bin = ev[q].bin;
    if (flag == 0 || flag == 1) {                       /* coverHelper (:172-218) */
      int64_t st = 0; int32_t count = 0, cmax = 0; int rec = 0;
      for (int64_t t = q; t < r; ++t) {
        int has_next = t + 1 < r;
        int32_t nc = count + ev[t].delta;
        int inr = nc >= mn && nc <= mx;
        int nrec = inr && has_next;
        int32_t ncmax = (!nrec && rec) ? 0 : ((nrec && nc > cmax) ? nc : cmax);
        int64_t nst = (inr && !rec) ? ev[t].off : st;
        if (!nrec && rec) rc |= pv_push(pv, nst + bin * B, (ev[t].off > B ? B : ev[t].off) + bin * B, cmax);
        count = nc; st = nst; cmax = ncmax; rec = nrec;
      }
    } else if (flag == 3) {                             /* histogramHelper (:229-258) */
      int32_t count = 0; int64_t st = 0;
      for (int64_t t = q; t < r; ++t) {
        int32_t nc = count + ev[t].delta;
        int64_t nst = (nc >= mn && nc <= mx && nc != count) ? ev[t].off : st;
        if (count >= mn && count <= mx && nc != count && count != 0) rc |= pv_push(pv, st + bin * B, ev[t].off + bin * B, count);
        st = nst; count = nc;
      }
    } else {                                            /* summitHelper (:269-316) */
      int64_t st = 0; int32_t count = 0; int valid_ = 0, re = 0, grow = 0;
      for (int64_t t = q; t < r; ++t) {
        int32_t nc = count + ev[t].delta;
        int nvalid = nc >= mn && nc <= mx;
        int ngrow = nc > count || (grow && nc == count);
        int nre = !valid_ && nvalid;
        int64_t nst = (nvalid && nc != count) ? ev[t].off : st;
        if (((valid_ && grow && (!ngrow || !nvalid)) || (re && !ngrow)) && count != 0)
          rc |= pv_push(pv, st + bin * B, (ev[t].off > B ? B : ev[t].off) + bin * B, count);
        st = nst; count = nc; valid_ = nvalid; re = nre; grow = ngrow;
      }
    }
    q = r;
  }
  free(ev);
  ck->rc = rc;
}

/* split + boundary merge of one segment's pieces (:121-152); pieces arrive in bin order */
static int merge_segment(int flag, int64_t B, const Chunk* chunks, int64_t n_chunks, PieceVec* fin) {
  PieceVec nv = {0, 0, 0};
  int rc = 0;
  for (int64_t c = 0; c < n_chunks; ++c)
    for (int64_t q = 0; q < chunks[c].pieces.n; ++q) {
      const Piece* p = &chunks[c].pieces.p[q];
      if (p->start % B != 0 && p->stop % B != 0) rc |= pv_push(fin, p->start, p->stop, p->acc);
      else rc |= pv_push(&nv, p->start, p->stop, p->acc);
    }
  if (nv.n) {
    qsort(nv.p, (size_t)nv.n, sizeof(Piece), cmp_piece);
    Piece old = nv.p[0];
    for (int64_t q = 1; q < nv.n; ++q) {
      Piece cur = nv.p[q];
      if (old.stop == cur.start && (flag != 3 || old.acc == cur.acc)) { old.stop = cur.stop; if (cur.acc > old.acc) old.acc = cur.acc; }
      else { rc |= pv_push(fin, old.start, old.stop, old.acc); old = cur; }
    }
    rc |= pv_push(fin, old.start, old.stop, old.acc);
  }
  free(nv.p);
  return rc;
}

/* GMAP4 (GMAP4.scala:20-100) for one phase-1 region against the chunks' binned inputs; returns 1 when a row was produced */
static int gmap4_region(int flag, int64_t B, int32_t group, int32_t chrom, int8_t strand, const Piece* pc, const Chunk* chunks, int64_t n_chunks, int64_t first_bin,
                        const int64_t* start, const int64_t* stop, const double* val, const uint8_t* valid, CoverRow* o) {
  int64_t cs = pc->start, ce = pc->stop;
  int64_t cnt = 0, smin = 0, emax = 0, smax = 0, emin = 0;
  int32_t nn = 0; double sum = 0, vmin = 0, vmax = 0;
  for (int64_t k = cs / B; k <= ce / B; ++k) {
    int64_t ci = (k - first_bin) / CHUNK_BINS;
    if (ci < 0 || ci >= n_chunks) continue;
    const Chunk* ck = &chunks[ci];
    int64_t lo = 0, hi = ck->nb;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (ck->eb[mid].bin < k) lo = mid + 1; else hi = mid; }
    for (int64_t y = lo; y < ck->nb && ck->eb[y].bin == k; ++y) {
      int32_t e = ck->eb[y].idx;
      if (cs < stop[e] && start[e] < ce && ((cs / B) == k || (start[e] / B) == k)) {   /* :39-46 */
        if (cnt == 0) { smin = start[e]; emax = stop[e]; smax = start[e]; emin = stop[e]; }
        else {                                                                         /* reduce (:68-74) */
          if (start[e] < smin) smin = start[e];
          if (stop[e] > emax) emax = stop[e];
          int64_t a = smax > start[e] ? smax : start[e], b2 = emin < stop[e] ? emin : stop[e];
          if (a < b2) { smax = a; emin = b2; } else { smax = 0; emin = 0; }
        }
        ++cnt;
        if (val && (!valid || valid[e])) {
          double x = val[e];
          if (nn == 0) { sum = x; vmin = x; vmax = x; } else { sum += x; vmin = java_min(vmin, x); vmax = java_max(vmax, x); }
          ++nn;
        }
      }
    }
  }
  if (cnt == 0) return 0;                                                              /* :33-34,48 */
  double os = flag == 1 ? (double)smin : (double)cs, oe = flag == 1 ? (double)emax : (double)ce;   /* :83-84 */
  o->group = group; o->chrom = chrom; o->strand = strand; o->start = (int64_t)os; o->stop = (int64_t)oe; o->acc = pc->acc;
  int64_t den = emax - smin;
  o->j1 = den != 0 ? fabs((oe - os) / (double)den) : 0.0;
  o->j2 = den != 0 ? fabs(((double)emin - (double)smax) / (double)den) : 0.0;
  o->count = (int32_t)cnt; o->nn = nn; o->sum = sum; o->vmin = vmin; o->vmax = vmax;
  return 1;
}

static CoverRow* g_rows = 0;
static int64_t g_nrows = 0;

/* Runs COVER(flag) and keeps the rows; fetch them with oracle_cover_fetch.  flag: 0 COVER, 1 FLAT, 2 SUMMIT, 3 HISTOGRAM.
 * sample_group may be NULL (one group).  mn/mx per group.  val/valid: one optional aggregated column.  Returns #rows or -1. */
int64_t oracle_cover(int64_t n, const int32_t* chrom, const int64_t* start, const int64_t* stop, const int8_t* strand, const int32_t* sample,
                     const int32_t* sample_group, int32_t n_groups, int32_t flag, const int32_t* mn, const int32_t* mx, int64_t B,
                     const double* val, const uint8_t* valid) {
  free(g_rows); g_rows = 0; g_nrows = 0;
  int unknown = strand == 0;                                                             /* assignGroups (:328) */
  for (int64_t i = 0; i < n && !unknown; ++i) if (strand[i] == 0) unknown = 1;
  /* segments (group, strand, chrom), in that order; regions of a segment in ascending row order (a counting sort) */
  int32_t n_chrom = 0;
  for (int64_t i = 0; i < n; ++i) if (chrom[i] + 1 > n_chrom) n_chrom = chrom[i] + 1;
  if (n_groups < 1) n_groups = 1;
  const int64_t n_seg_keys = (int64_t)n_groups * 3 * (n_chrom > 0 ? n_chrom : 1);
  int64_t* seg_cnt = (int64_t*)calloc((size_t)n_seg_keys + 1, sizeof(int64_t));
  int64_t* seg_maxbin = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_seg_keys ? n_seg_keys : 1));
  int64_t* seg_minbin = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_seg_keys ? n_seg_keys : 1));
  if (!seg_cnt || !seg_maxbin || !seg_minbin) return -1;
  for (int64_t k = 0; k < n_seg_keys; ++k) { seg_maxbin[k] = -1; seg_minbin[k] = INT64_MAX; }
#define SEGKEY(i) (((int64_t)(sample_group ? sample_group[sample ? sample[i] : 0] : 0) * 3 + (unknown ? 0 : strand[i])) * n_chrom + chrom[i])
  for (int64_t i = 0; i < n; ++i) {
    int64_t k = SEGKEY(i);
    int64_t sb = start[i] / B, eb = stop[i] / B;
    if (eb > seg_maxbin[k]) seg_maxbin[k] = eb;
    if (sb < seg_minbin[k]) seg_minbin[k] = sb;
  }
  /* chunks of every non-empty segment */
  int64_t* seg_chunk0 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_seg_keys + 1));
  if (!seg_chunk0) return -1;
  int64_t n_chunks = 0;
  for (int64_t k = 0; k < n_seg_keys; ++k) {
    seg_chunk0[k] = n_chunks;
    if (seg_maxbin[k] >= 0) n_chunks += (seg_maxbin[k] - seg_minbin[k]) / CHUNK_BINS + 1;
  }
  seg_chunk0[n_seg_keys] = n_chunks;
  Chunk* chunks = (Chunk*)calloc((size_t)(n_chunks ? n_chunks : 1), sizeof(Chunk));
  int64_t* ck_off = (int64_t*)calloc((size_t)n_chunks + 2, sizeof(int64_t));
  if (!chunks || !ck_off) return -1;
  /* a region goes to every chunk its bins touch (counting sort by chunk, ascending row order inside a chunk) */
  for (int64_t i = 0; i < n; ++i) {
    int64_t k = SEGKEY(i);
    int64_t c0 = seg_chunk0[k] + (start[i] / B - seg_minbin[k]) / CHUNK_BINS, c1 = seg_chunk0[k] + (stop[i] / B - seg_minbin[k]) / CHUNK_BINS;
    for (int64_t c = c0; c <= c1; ++c) ck_off[c + 1]++;
  }
  for (int64_t c = 0; c < n_chunks; ++c) ck_off[c + 1] += ck_off[c];
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ck_off[n_chunks] ? ck_off[n_chunks] : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_chunks + 1));
  if (!idx || !fill) return -1;
  memcpy(fill, ck_off, sizeof(int64_t) * (size_t)(n_chunks + 1));
  for (int64_t i = 0; i < n; ++i) {
    int64_t k = SEGKEY(i);
    int64_t c0 = seg_chunk0[k] + (start[i] / B - seg_minbin[k]) / CHUNK_BINS, c1 = seg_chunk0[k] + (stop[i] / B - seg_minbin[k]) / CHUNK_BINS;
    for (int64_t c = c0; c <= c1; ++c) idx[fill[c]++] = (int32_t)i;
  }
  free(fill);
  for (int64_t k = 0; k < n_seg_keys; ++k)
    for (int64_t c = seg_chunk0[k]; c < seg_chunk0[k + 1]; ++c) {
      chunks[c].b_lo = seg_minbin[k] + (c - seg_chunk0[k]) * CHUNK_BINS;
      chunks[c].b_hi = chunks[c].b_lo + CHUNK_BINS;
    }
  int failed = 0;
  /* phase 1: every chunk on its own (the analogue of Spark's per-(chrom, bin, group, strand) partitions in local[N]) */
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < n_chunks; ++c) {
    int64_t k = 0;
    { int64_t lo = 0, hi = n_seg_keys; while (hi - lo > 1) { int64_t mid = (lo + hi) / 2; if (seg_chunk0[mid] <= c) lo = mid; else hi = mid; } k = lo; while (seg_chunk0[k + 1] <= c) ++k; }
    int32_t g = (int32_t)(k / ((int64_t)3 * n_chrom));
    cover_chunk(flag, B, mn[g], mx[g], idx + ck_off[c], ck_off[c + 1] - ck_off[c], start, stop, &chunks[c]);
    if (chunks[c].rc) failed = 1;
  }
  /* boundary merge per segment, GMAP4 per phase-1 region */
  RowVec* per = (RowVec*)calloc((size_t)(n_seg_keys ? n_seg_keys : 1), sizeof(RowVec));
  if (!per) return -1;
  for (int64_t k = 0; k < n_seg_keys && !failed; ++k) {
    int64_t nck = seg_chunk0[k + 1] - seg_chunk0[k];
    if (nck == 0) continue;
    PieceVec fin = {0, 0, 0};
    if (merge_segment(flag, B, chunks + seg_chunk0[k], nck, &fin)) { failed = 1; free(fin.p); break; }
    CoverRow* rows = (CoverRow*)malloc(sizeof(CoverRow) * (size_t)(fin.n ? fin.n : 1));
    uint8_t* keep = (uint8_t*)calloc((size_t)(fin.n ? fin.n : 1), 1);
    if (!rows || !keep) { failed = 1; free(fin.p); free(rows); free(keep); break; }
    int32_t g = (int32_t)(k / ((int64_t)3 * n_chrom));
    int8_t st = (int8_t)((k / n_chrom) % 3);
    int32_t ch = (int32_t)(k % n_chrom);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t c = 0; c < fin.n; ++c)
      keep[c] = (uint8_t)gmap4_region(flag, B, g, ch, st, &fin.p[c], chunks + seg_chunk0[k], nck, seg_minbin[k], start, stop, val, valid, &rows[c]);
    int64_t w = 0;
    for (int64_t c = 0; c < fin.n; ++c) if (keep[c]) rows[w++] = rows[c];
    per[k].p = rows; per[k].n = w; per[k].cap = fin.n;
    free(keep); free(fin.p);
  }
  int64_t tot = 0;
  for (int64_t k = 0; k < n_seg_keys; ++k) tot += per[k].n;
  if (!failed) {
    g_rows = (CoverRow*)malloc(sizeof(CoverRow) * (size_t)(tot ? tot : 1));
    if (!g_rows) failed = 1;
  }
  if (!failed) {
    int64_t w = 0;
    for (int64_t k = 0; k < n_seg_keys; ++k) { if (per[k].n) memcpy(g_rows + w, per[k].p, sizeof(CoverRow) * (size_t)per[k].n); w += per[k].n; }
    g_nrows = tot;
  }
  for (int64_t k = 0; k < n_seg_keys; ++k) free(per[k].p);
  for (int64_t c = 0; c < n_chunks; ++c) { free(chunks[c].pieces.p); free(chunks[c].eb); }
  free(per); free(chunks); free(ck_off); free(idx); free(seg_cnt); free(seg_maxbin); free(seg_minbin); free(seg_chunk0);
#undef SEGKEY
  return failed ? -1 : tot;
}

void oracle_cover_fetch(int32_t* group, int32_t* chrom, int64_t* start, int64_t* stop, int8_t* strand, int32_t* acc, double* j1, double* j2,
                        int32_t* count, int32_t* nn, double* sum, double* vmin, double* vmax) {
  for (int64_t i = 0; i < g_nrows; ++i) {
    const CoverRow* r = &g_rows[i];
    group[i] = r->group; chrom[i] = r->chrom; start[i] = r->start; stop[i] = r->stop; strand[i] = r->strand; acc[i] = r->acc; j1[i] = r->j1; j2[i] = r->j2;
    if (count) count[i] = r->count;
    if (nn) { nn[i] = r->nn; sum[i] = r->sum; vmin[i] = r->vmin; vmax[i] = r->vmax; }
  }
}
