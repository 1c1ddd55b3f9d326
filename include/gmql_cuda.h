/*
 * gmql_cuda.h — C ABI of libgmqlcuda.so, the B200 (sm_100a) back end for GMQL's
 * region-operator hot path: genometric MAP, COVER/FLAT/SUMMIT/HISTOGRAM and JOIN.
 *
 * The reference (DEIB-GECO/GMQL) has no FFI: its plug-in point is the Scala class
 * `Implementation` (GMQL-Server/.../GMQLServer/Implementation.scala:22-71) and the
 * operator dispatch `GMQLSparkExecutor.implement_rd`
 * (GMQL-Spark/.../implementation/GMQLSparkExecutor.scala:251-285).  The entry points
 * below are what a JNI shim under a `GMQLCudaExecutor extends GMQLSparkExecutor`
 * binds for the three IR nodes it overrides (INTEGRATION.md shows that binding):
 *
 *   gmql_map    replaces GenometricMap71.apply   (RegionsOperators/GenometricMap/GenometricMap71.scala:28)
 *   gmql_cover  replaces GenometricCover.apply   (RegionsOperators/GenometricCover/GenometricCover.scala:30)
 *                        + GMAP4.apply           (RegionsOperators/GenometricMap/GMAP4.scala:20)
 *   gmql_join   replaces GenometricJoin.apply    (RegionsOperators/GenometricJoin.scala:21)
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - every call returns 0 on success or a negative gmql_status; the message is
 *    available from gmql_last_error(ctx) (reference convention: operators throw,
 *    GMQLSparkExecutor.scala:190-207 rethrows — the JNI shim turns <0 into an exception);
 *  - a context is used by one thread at a time; distinct contexts are independent
 *    (no global mutable state), matching GMQLExecute's 5-job pool (GMQLExecute.scala:29-30);
 *  - inputs are owned by the caller and must stay alive for the duration of the call;
 *    results are owned by the library until gmql_result_free; results and datasets live in their context's memory pool and
 *    must be freed (with that context) before gmql_ctx_destroy;
 *  - device memory comes from a pool private to the context (cudaMemPoolCreate): the process's default pool is untouched;
 *  - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef GMQL_CUDA_H
#define GMQL_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMQL_ABI_VERSION 1

typedef enum {
  GMQL_OK = 0,
  GMQL_ERR_INVALID = -1,      /* bad argument / inconsistent table */
  GMQL_ERR_UNSUPPORTED = -2,  /* semantically valid in GMQL but not accelerated (no silent fallback) */
  GMQL_ERR_CUDA = -3,         /* CUDA runtime error, message has the cudaError string */
  GMQL_ERR_NOMEM = -4,
  GMQL_ERR_NODEVICE = -5
} gmql_status;

typedef struct gmql_ctx gmql_ctx;       /* opaque: device, stream, memory pool, last error */
typedef struct gmql_result gmql_result; /* opaque: device-resident result columns */

/* strand encoding (GRecordKey.strand, core/GRecordKey.scala:8; BedParser.scala:166-170) */
enum { GMQL_STRAND_BOTH = 0 /* '*' */, GMQL_STRAND_PLUS = 1 /* '+' */, GMQL_STRAND_MINUS = 2 /* '-' */ };

enum { GMQL_LOC_HOST = 0, GMQL_LOC_DEVICE = 1 };

/*
 * A region dataset as structure-of-arrays: the columnar image of RDD[GRECORD]
 * (GRECORD = (GRecordKey(id, chrom, start, stop, strand), Array[GValue]),
 * core/DataTypes.scala:62).  String columns stay on the host and are re-attached by
 * row index; numeric GValues are doubles (BedParser.scala:35-42) with a validity byte
 * (0 = GNull).
 */
typedef struct {
  int64_t n;                    /* number of regions (< 2^31 in this version) */
  int32_t location;             /* GMQL_LOC_HOST (pinned or pageable) or GMQL_LOC_DEVICE: applies to every array below */
  int32_t coord_bits;           /* 32: start/stop are int32_t[]; 64: int64_t[] */
  const int32_t* chrom;         /* [n] chromosome index 0..n_chrom-1 into the caller's name dictionary */
  const void* start;            /* [n] 0-based, half-open [start, stop) */
  const void* stop;             /* [n] */
  const int8_t* strand;         /* [n] GMQL_STRAND_*; NULL = every region is '*' */
  const int32_t* sample;        /* [n] index 0..n_samples-1 into sample_ids; NULL = all sample 0 */
  int32_t n_chrom;
  int32_t n_samples;
  const int64_t* sample_ids;    /* [n_samples] HOST array: GMQL sample ids (md5-derived, Loaders.scala:203-208) */
  int32_t n_cols;               /* numeric value columns available to aggregates */
  const double* const* cols;    /* HOST array of n_cols pointers (each [n], in `location`) */
  const uint8_t* const* valid;  /* HOST array of n_cols pointers or NULL; valid[c]==NULL = no GNull in column c */
  const int32_t* dup_of;        /* [n] or NULL.  dup_of[i] = lowest row index whose full record (sample, coords,
                                   strand, ALL values incl. host-side strings) equals row i (== i if unique).
                                   Lets the device reproduce the MapKey collapse (GenometricMap71.scala:93,123,186-203;
                                   GenometricJoin.scala:134-136) whose value equality only the host can decide. */
  /* optional narrow encodings (fewer bytes over PCIe); zero / NULL = not used */
  int32_t chrom_bits;           /* 0 or 32: chrom is int32_t[]; 8: uint8_t[]; 16: uint16_t[] (n_chrom must fit) */
  const int64_t* sample_offsets;/* HOST array [n_samples+1] or NULL: rows [off[s], off[s+1]) belong to sample s (regions
                                   arrive one sample file after the other, Loaders.scala:203-208); off[0] = 0,
                                   off[n_samples] = n, non-decreasing.  With sample == NULL it defines the sample of
                                   every row; next to a sample column it is a promise about that column's layout
                                   (checked on the device where it is relied upon: gmql_map's per-sample COUNT kernel).
                                   Views returned by gmql_dataset_upload carry it when the upload had it. */
  const int64_t* chrom_span_hint;/* HOST array [n_chrom] or NULL: an upper bound of `stop` per chromosome (e.g. the
                                   assembly's chromosome sizes).  Lets gmql_map / gmql_difference lay out the linear coordinate
                                   without a pass over the data and a host round trip; a region beyond its bound is an
                                   error (GMQL_ERR_INVALID), never a wrong result.  gmql_cover uses it to queue its tile
                                   pipeline without waiting for its own prepass (a region beyond its bound there only
                                   makes the call take the ordinary route).  Views made by gmql_result_as_regions carry
                                   it, and so do views made by gmql_dataset_upload (the largest stop seen per
                                   chromosome, computed once at upload, when the caller gave none). */
  int32_t stop_is_length;       /* 0: `stop` holds stop coordinates.  16 / 32: `stop` holds region LENGTHS (stop - start) as
                                   uint16_t[] / uint32_t[] — narrowPeak-like data (lengths below 65 536) then crosses PCIe in
                                   2 bytes instead of 4 or 8 per region; the stop column is rebuilt on the device. */
  int32_t sorted_by_position;   /* 1: the rows are ordered by (chrom index, start) — e.g. a COVER result or a sorted annotation.
                                   gmql_map / gmql_difference then index the reference side without sorting it; the order is
                                   verified on the device and a violation is an error (GMQL_ERR_INVALID). */
} gmql_regions;

/* aggregates: DefaultRegionsToRegionFactory (GMQL-Server/.../DefaultRegionsToRegionFactory.scala:13-31),
 * dispatched on RegionsToRegion.function_identifier + index (RegionsToRegion.scala:9-17). */
typedef enum {
  GMQL_AGG_COUNT = 0, GMQL_AGG_SUM = 1, GMQL_AGG_MIN = 2, GMQL_AGG_MAX = 3,
  GMQL_AGG_AVG = 4, GMQL_AGG_MEDIAN = 5 /* pairwise-fold artefact == MIN, see oracle */,
  GMQL_AGG_BAG = 6, GMQL_AGG_BAGD = 7   /* device emits matched row lists; strings are built on the host */
} gmql_agg_kind;
typedef struct { int32_t kind; int32_t col; } gmql_agg;

/* (left/ref sample, right/exp sample) pairs from implement_mjd (MetaJoinMJD2.scala:26-111);
 * the output sample id md5(left_id, right_id) is computed by the host side (GenometricMap71.scala:174). */
typedef struct { int32_t left_sample; int32_t right_sample; } gmql_pair;

/* COVER variants (CoverParameters/CoverFlag; GenometricCover.scala:112-117) */
typedef enum { GMQL_COVER = 0, GMQL_FLAT = 1, GMQL_SUMMIT = 2, GMQL_HISTOGRAM = 3 } gmql_cover_flag;
/* OR-ed into `flag`: treat the dataset as containing a '*' region.  Strand unification is decided on the WHOLE dataset
 * (assignGroups, GenometricCover.scala:328); a caller that shards the input (by chromosome, across GPUs) passes the
 * global answer to every shard. */
#define GMQL_COVER_FORCE_UNIFIED 0x100

/* JOIN atoms after the Compiler's rewriting (GmqlParsers.scala:462-487):
 * DLE(n)->DISTLESS(n+1), DGE(n)->DISTGREATER(n-1); predicates are strict (GenometricJoin.scala:250-251). */
typedef enum { GMQL_J_DISTLESS = 0, GMQL_J_DISTGREATER = 1, GMQL_J_MINDIST = 2, GMQL_J_UPSTREAM = 3, GMQL_J_DOWNSTREAM = 4 } gmql_atom_kind;
typedef struct { int32_t kind; int64_t arg; } gmql_atom;

/* RegionBuilder (JoinParametersRD/RegionBuilder.scala:7-12; GenometricJoin.scala:345-372) */
typedef enum { GMQL_B_LEFT = 0, GMQL_B_LEFT_DISTINCT = 1, GMQL_B_RIGHT = 2, GMQL_B_RIGHT_DISTINCT = 3,
               GMQL_B_CONTIG = 4, GMQL_B_INTERSECTION = 5, GMQL_B_BOTH = 6 } gmql_builder;
/* OR-ed into `builder`: emit only LEFT_ROW / RIGHT_ROW / PAIR (8+4 B per output record); the built coordinates are
 * derivable from the rows and can be materialised by the caller */
#define GMQL_JOIN_ROWS_ONLY 0x100

/* ---- context ------------------------------------------------------------------------- */
int  gmql_abi_version(void);
/* device: CUDA ordinal; stream: a cudaStream_t to launch on (e.g. the caller's current stream) or NULL to create one */
int  gmql_ctx_create(int device, void* stream, gmql_ctx** out);
void gmql_ctx_destroy(gmql_ctx* ctx);
const char* gmql_last_error(const gmql_ctx* ctx);
int  gmql_ctx_sync(gmql_ctx* ctx);                      /* cudaStreamSynchronize on the context stream; reports deferred checks */
/* Per-context options (strings; value NULL or "" clears).  They never change results, only which device path computes them
 * (test hooks, tuning) or when an error is reported:
 *   "cover.path"      "tiled" | "sort" | "bins"   force the tile-resident, the sort-and-merge or the bin-wise path of COVER / FLAT / HISTOGRAM
 *   "cover.logw"      "8".."16"          tile width of the tile-resident path
 *   "cover.kernel"    "rank" | "count"   tile kernel: bitmap-rank (round 1) or word-count (default where it applies)
 *   "cover.waves", "cover.stitch" ("split"), "cover.speculate" ("0"), "map.index" ("split"), "map.grouped" ("0" | "1"), "map.lean" ("0")
 *   "cover.gmap4"     "scatter"          GMAP4 over many small phase-1 regions with REDs instead of the merge form
 *   "cover.finish"    "split"            merge-form GMAP4 through the intermediate arrays and the compaction pass instead of straight
 *                                        into the result columns
 *   "cover.bins"      "inplace"          the per-bin state machines read their points in place instead of from a shared-memory stage,
 *                                        and the bin list comes from head flags + scan instead of one search per bin
 *   "join.md"         "team" | "lane"    MD(k) kernel: a warp per 32 neighbouring left regions walking the right samples, or a lane per
 *                                        (left region, right sample) group (default: team when the pair table is dense)
 *   "join.md_memo"    "0"                the team kernel's write pass repeats the selection instead of reading the count pass's memo
 *   "defer_check"     "1"                gmql_map (one reference sample) returns without waiting for the device: an invalid
 *                                        input (bad index, broken layout promise) is then reported by the next call on this
 *                                        context that finds the stream idle — gmql_result_column_copy, gmql_ctx_sync or the next
 *                                        operator — as GMQL_ERR_INVALID.  For pipelines that want one host synchronisation per
 *                                        COVER -> MAP step.  Default: every operator reports its own errors.
 * Options are per context: nothing in this library reads process-global state (environment variables). */
int  gmql_ctx_set_option(gmql_ctx* ctx, const char* name, const char* value);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t gmql_ctx_launch_count(const gmql_ctx* ctx);
/* pinned host memory for the SoA columns (JNI: wrap with NewDirectByteBuffer) */
int  gmql_host_alloc(size_t bytes, void** out);
void gmql_host_free(void* p);

/* Device-resident copy of a host dataset: the analogue of the per-IR-node memo (IROperator.intermediateResult,
 * GMQLSparkExecutor.scala:252-253,282) so that a dataset consumed by several operators (`C = COVER(..) E; M = MAP() C E`)
 * crosses PCIe once.  `view` is filled with device pointers (location = GMQL_LOC_DEVICE) that stay valid until
 * gmql_dataset_free; the small host arrays of the view (sample_ids, cols, valid, sample_offsets, chrom_span_hint) are owned by
 * the dataset too.  All n_cols value columns are uploaded.  The arrays stay resident in the encodings they arrived in: a
 * dataset sent with chrom_bits = 8, stop_is_length = 16 and sample_offsets costs 7 bytes per region of device memory, and the
 * view says so (chrom_bits / stop_is_length / sample == NULL); operators without a path for the narrow encodings widen them once
 * per dataset.  The upload also gathers the dataset's statistics in one read (largest stop per chromosome — handed on as the
 * view's chrom_span_hint — longest region, zero-length / '*' flags, row validity) and, for datasets the tile-resident COVER
 * path accepts, the partition offsets of that path: the per-node memo of GMQLSparkExecutor.scala:252-253 extended to what every
 * operator on the dataset would otherwise recompute. */
typedef struct gmql_dataset gmql_dataset;
int  gmql_dataset_upload(gmql_ctx* ctx, const gmql_regions* host, gmql_dataset** out, gmql_regions* view);
void gmql_dataset_free(gmql_ctx* ctx, gmql_dataset* d);

/* Per-kernel device timing (CUDA events around every launch of this context).  Off by default; meant for one extra
 * profiled step, not for the timed region.  gmql_ctx_profile_report writes lines "kernel_name launches total_ms"
 * sorted by total time into buf (NUL-terminated, truncated to cap) and clears the log. */
int  gmql_ctx_set_profiling(gmql_ctx* ctx, int enabled);
int  gmql_ctx_profile_report(gmql_ctx* ctx, char* buf, size_t cap);

/* ---- MAP -------------------------------------------------------------------------------
 * For every pair (a, b) and every ref region r of sample a: S = { e in exp sample b : e.chrom == r.chrom,
 * r.start < e.stop, e.start < r.stop, strands compatible }; emits count = |S| (also when 0) and the
 * aggregates over S (GenometricMap71.scala:81-147; bin-invariant).
 *
 * Result columns (gmql_result_column): rows are grouped by pair in pair-list order; inside pair (a, b)
 * the rows follow sample a's regions in ascending original row index.
 *   GMQL_COL_MAP_ROW_OFFSETS  int64 [n_pairs+1]   first row of each pair
 *   GMQL_COL_MAP_REF_ROW      int32 [n_ref]       ref row indices grouped by ref sample (ascending inside a sample)
 *   GMQL_COL_MAP_REF_SAMPLE_OFFSETS int64 [n_ref_samples+1]  offsets into REF_ROW
 *   GMQL_COL_MAP_COUNT        int32 [n_rows]      |S|; -1 = row collapsed into its dup_of row (not emitted by GMQL)
 *   GMQL_COL_AGG_VALUE + k    double [n_rows]     aggregate k (BAG/BAGD: unused)
 *   GMQL_COL_AGG_VALID + k    uint8 [n_rows]      0 = GNull
 *   GMQL_COL_BAG_OFFSETS      int64 [n_rows+1], GMQL_COL_BAG_ROWS int32 [...]  matched exp rows (only if a BAG/BAGD is requested)
 */
int gmql_map(gmql_ctx* ctx, const gmql_regions* ref, const gmql_regions* exp,
             const gmql_pair* pairs, int64_t n_pairs,
             const gmql_agg* aggs, int32_t n_aggs, gmql_result** out);

/* ---- COVER family ----------------------------------------------------------------------
 * sample_group: HOST [in->n_samples] group index 0..n_groups-1 per sample (implement_mgd,
 * MetaGroupMGD.scala:20-79) or NULL = one group (id 0L).  min_acc/max_acc: HOST [n_groups] thresholds
 * already resolved from N/ANY/ALL (GenometricCover.scala:57-79); use gmql_cover_count_samples for the
 * ungrouped ALL case.  bin_size = BinSize.Cover (2000): observable in results (stop%bin==0 extension,
 * SUMMIT bin cuts; GenometricCover.scala:345-360, 269-316).  A threshold min_acc < 1 (N(0), ALL - k) makes depth 0
 * "in range": coverHelper (:172-218) then records from the first to the last event of every bin that holds events, so
 * the result depends on the bins; such calls (and SUMMIT always) run the reference's per-bin state machines literally.
 *
 * Result columns: one row per output region, sorted by (group, strand, chrom, start):
 *   GMQL_COL_COV_GROUP int32, GMQL_COL_COV_CHROM int32, GMQL_COL_COV_START int64, GMQL_COL_COV_STOP int64,
 *   GMQL_COL_COV_STRAND int8, GMQL_COL_COV_ACC int32 (AccIndex), GMQL_COL_COV_JACC_INTERSECT double,
 *   GMQL_COL_COV_JACC_RESULT double, GMQL_COL_AGG_VALUE/VALID + k.
 * BAG / BAGD (DefaultRegionsToRegionFactory.scala:128-168, folded by GMAP4.scala:57-66): strings are host business, so the
 * device lists the CONTRIBUTORS of every output row — the input regions of the row's group that overlap the run with a
 * compatible strand (GMAP4.scala:38-45) — as GMQL_COL_BAG_OFFSETS int64 [rows + 1] / GMQL_COL_BAG_ROWS int32 (row indices of
 * `in`, unordered inside a list); the host folds the values' text forms (gmql_b200/executor.py).  For FLAT the
 * contributors are those of the run behind the row, not of the row's own (wider) coordinates.
 */
int gmql_cover(gmql_ctx* ctx, const gmql_regions* in, const int32_t* sample_group, int32_t n_groups,
               int32_t flag, const int32_t* min_acc, const int32_t* max_acc, int64_t bin_size,
               const gmql_agg* aggs, int32_t n_aggs, gmql_result** out);
/* ---- COVER / FLAT on one coordinate-range shard (multi-GPU: north_star "by genomic coordinate range with halo handling") ----
 * The linear coordinate lin(chrom, pos) = coff[chrom] + pos, coff = exclusive prefix sum of (chrom_span[c] + bin_size + 2), is cut
 * into tiles of 65 536 positions; a shard owns the tiles [first_tile, end_tile).  `in` must hold every region whose start lies in
 * the tiles [first_tile - 1, end_tile) (the tile before the range only lends the regions still open at the cut: the halo; other
 * rows are ignored).  chrom_span (HOST [n_chrom], an upper bound of stop per chromosome) must be identical on every shard.
 * Requirements: COVER or FLAT, one group, unified strands, no aggregates, no zero-length regions, regions shorter than 65 534
 * bases, the linear genome below 2^32 (otherwise GMQL_ERR_UNSUPPORTED: shard by chromosome instead).
 *
 * Result: the columns of gmql_cover for the runs inside the shard, plus what a cross-shard merge needs (all chromosome-local):
 *   GMQL_COL_COV_EDGE int32 — bit 0: the run continues one that ends at the shard's first position, bit 1: it reaches the shard's
 *     end (such rows are kept even without a contributor here); rows with EDGE == 0 are final,
 *   GMQL_COL_COV_RAW_START / _RAW_STOP int64 (the run clipped to the shard), GMQL_COL_COV_RAW_COUNT int32 (contributors accounted
 *     on this shard: every (run, region) pair is accounted on exactly one shard, the one of max(run start, region start), the
 *     reference's first-bin rule GMAP4.scala:44-45 at shard granularity), GMQL_COL_COV_RAW_SMIN / _EMAX / _SMAX / _EMIN int64
 *     (startMin, endMax, startMax, endMin over them, -1 when none).
 * Runs that touch at a cut (left.RAW_STOP == right.RAW_START, EDGE bits 2 and 1) are one run of the reference (its boundary
 * merge, GenometricCover.scala:121-152): AccIndex = max, count = sum, min / max of the four bounds, then the GMAP4 formulas
 * (gmql_b200/dist.py::merge_cover_shards is the host side; NCCL only gathers the few edge rows). */
typedef struct {
  const int64_t* chrom_span;
  int64_t first_tile, end_tile;
} gmql_cover_shard_t;
int gmql_cover_shard(gmql_ctx* ctx, const gmql_regions* in, int32_t flag, int32_t min_acc, int32_t max_acc, int64_t bin_size,
                     const gmql_cover_shard_t* shard, gmql_result** out);

/* number of samples that own at least one region (GenometricCover.scala:64) */
int gmql_cover_count_samples(gmql_ctx* ctx, const gmql_regions* in, int32_t* out_count);

/* ---- JOIN ------------------------------------------------------------------------------
 * atoms: one JoinQuadruple's atomic conditions in script order (<= 4).  bin_size = BinSize.Join (5000) and
 * max_bin_distance (100000) define the candidate window exactly as GenometricJoin.scala:75-79,284-342.
 * left_eq/right_eq: optional per-row int64 codes; when both non-NULL a pair also needs equal codes
 * (joinOnAttributes, GenometricJoin.scala:102-114; the host dictionary-encodes the attribute tuple).
 *
 * Result columns: one row per output record (unordered, as the RDD is):
 *   GMQL_COL_JOIN_LEFT_ROW int32, GMQL_COL_JOIN_RIGHT_ROW int32 (-1 for LEFT_DISTINCT),
 *   GMQL_COL_JOIN_PAIR int32 (index into pairs), GMQL_COL_JOIN_START int64, GMQL_COL_JOIN_STOP int64,
 *   GMQL_COL_JOIN_STRAND int8, GMQL_COL_JOIN_CHROM int32 (built region, GenometricJoin.scala:345-372).
 */
int gmql_join(gmql_ctx* ctx, const gmql_regions* left, const gmql_regions* right,
              const gmql_pair* pairs, int64_t n_pairs,
              const gmql_atom* atoms, int32_t n_atoms, int32_t builder,
              int64_t bin_size, int64_t max_bin_distance,
              const int64_t* left_eq, const int64_t* right_eq,
              gmql_result** out);

/* ---- DIFFERENCE ------------------------------------------------------------------------
 * Replaces GenometricDifference.apply (RegionsOperators/GenometricDifference.scala:23-100; dispatch
 * GMQLSparkExecutor.scala:278, IRDifferenceRD(meta_join, left, right, exact)).  A reference region survives when
 * no region of ANY experiment sample paired with its sample overlaps it (r.start < e.stop, e.start < r.stop, strands
 * compatible; exact != 0: with identical coordinates as well, :67).  Regions of reference samples that appear in no
 * pair produce nothing (:104-114); identical records of one reference sample collapse into one output record (the
 * reduce key is the record's content, :54) — pass dup_of as for gmql_map.  The fixed 50 000-base binning (:19) is not
 * observable in the result.
 *
 * Result: GMQL_COL_DIFF_REF_ROW int32 [rows] — surviving rows of `ref`, ascending.  Output sample id =
 * md5(ref sample id) (:53), derived on the host.
 */
int gmql_difference(gmql_ctx* ctx, const gmql_regions* ref, const gmql_regions* exp,
                    const gmql_pair* pairs, int64_t n_pairs, int32_t exact, gmql_result** out);

/* ---- PROFILE ---------------------------------------------------------------------------
 * The reference profiles every materialised result with two foldByKey passes (Profiler.profile,
 * GMQL-Profiler/.../Profilers/Profiler.scala:124-153; call site GMQLSparkExecutor.scala:158-173).  gmql_profile
 * computes the per-sample ProfilerValue (:65) in one read of the coordinate columns: one row per sample index,
 * columns GMQL_COL_PROF_* (all int64; the two sums wrap like Java longs).  Samples without regions keep the fold's
 * zero value (leftMost = minLength = Long.MaxValue, rightMost = maxLength = Long.MinValue, count 0, :69); the
 * per-dataset fold and calculateResult (:84-92) are a few operations on these rows and stay on the host.
 */
int gmql_profile(gmql_ctx* ctx, const gmql_regions* in, gmql_result** out);

/* ---- GROUP by coordinates ----------------------------------------------------------------
 * GroupRD.apply (RegionsOperators/GroupRD.scala:22-64): the records of one sample that agree on (chrom, start, stop, strand)
 * and on the text forms of the grouping fields collapse into one record.  field_code: [in->n] (same location as `in`) one code
 * per row, equal exactly when the rows agree on every grouping field's text (the host knows strings; :34-42), or NULL without
 * grouping fields.  Every aggregate sees the group's whole value list (:48-55, `fun` is n-ary here): COUNT = members, SUM / AVG /
 * MIN / MAX over the non-null values (GNull when there is none), GMQL_AGG_MEDIAN = the TRUE lower median (sorted non-null values,
 * element (m - 1) / 2, DefaultRegionsToRegionFactory.scala:33-56 — unlike MAP / COVER, whose pairwise fold degrades it to MIN),
 * BAG / BAGD = member lists for the host.
 * Result: one row per group, ordered by (sample, chrom, start) and, inside that, by first appearance:
 *   GMQL_COL_GROUP_ROW int32 (the group's first input row: key and grouping-field values come from it),
 *   GMQL_COL_GROUP_COUNT int32 (members), GMQL_COL_AGG_VALUE / _VALID + k, and with BAG / BAGD GMQL_COL_BAG_OFFSETS int64
 *   [rows + 1] / GMQL_COL_BAG_ROWS int32 (the members of every group, input order).
 * MERGE (RegionsOperators/MergeRD.scala:20-45) relabels sample ids with group ids and touches no region: it stays on the host.
 */
int gmql_group(gmql_ctx* ctx, const gmql_regions* in, const int64_t* field_code, const gmql_agg* aggs, int32_t n_aggs, gmql_result** out);

/* ---- TEXT -> columns --------------------------------------------------------------------
 * Replaces the per-line work of BedParser.region_parser (GMQL-Spark/.../loaders/BedParser.scala:112-172; the concrete
 * parsers are column layouts: NarrowPeakParser :295-303, RnaSeqParser :346-348, BedScoreParser :525-527, custom schemas
 * :357-507): tab split, trim, strand "+" / "-" else '*', optional start - 1 for 1-based schemas (:164), numeric values
 * through String.toDouble with "null" / "." -> GNull (:30-48), '#' and empty lines skipped (:173-203).
 * `text` is one sample file (or several concatenated): n_bytes of tab-separated lines in `location`.  chrom_names is the
 * caller's chromosome dictionary (HOST strings); a name outside it yields chrom = -1 and status bit 2.
 *
 * Result: one row per kept line, in file order:
 *   GMQL_COL_TXT_LINE_OFFSET int64 (byte offset of the line: string columns stay in the text),
 *   GMQL_COL_TXT_STATUS uint8: 0 = parsed; bit 0 = a numeric field is left to the host's own parser (more than 19 significant
 *     digits with an undecidable tail, or a syntax beyond plain decimals: hex floats, NaN, Infinity, type suffixes) — its
 *     value is 0 / invalid here; bit 1 = malformed line (missing columns, bad integer: the reference throws, :173-203);
 *     bit 2 = chromosome name not in the dictionary,
 *   GMQL_COL_TXT_CHROM int32, GMQL_COL_TXT_START int64, GMQL_COL_TXT_STOP int64, GMQL_COL_TXT_STRAND int8,
 *   GMQL_COL_AGG_VALUE + k double, GMQL_COL_AGG_VALID + k uint8 for the k-th requested value column.
 * Doubles are correctly rounded (Eisel-Lemire), i.e. bit-identical to Double.parseDouble, whenever status bit 0 is clear.
 * The device columns can be handed to the operators as a GMQL_LOC_DEVICE gmql_regions (coord_bits = 64).
 */
typedef struct {
  int32_t chr_pos, start_pos, stop_pos;   /* 0-based column of each field */
  int32_t strand_pos;                     /* < 0: the format has no strand column (every region '*') */
  int32_t n_values;                       /* numeric value columns to extract (<= 16) */
  const int32_t* value_pos;               /* HOST [n_values] */
  int32_t start_minus_one;                /* 1-based schemas */
} gmql_text_schema;
int gmql_parse_text(gmql_ctx* ctx, const char* text, int64_t n_bytes, int32_t location, const gmql_text_schema* schema,
                    const char* const* chrom_names, int32_t n_chrom, gmql_result** out);

/* ---- columns -> TEXT --------------------------------------------------------------------
 * Replaces the per-region line construction of MATERIALIZE's writer, StoreTABRD
 * (GMQL-Spark/.../RegionsOperators/SelectRegions/StoreTABRD.scala:62-80): `(recordKey.productIterator.drop(1) ++ values)
 * .mkString("\t")`, i.e. chrom \t start \t stop \t strand \t value ... \n with the strand as '*' / '+' / '-', start + 1 for one-based
 * schemas (:62-69), GNull as "null" (DataTypes.scala:185) and GDouble.toString = DecimalFormat("#.########", ENGLISH) with
 * RoundingMode.FLOOR (GMQL-Core/.../DataTypes.scala:139-150): at most eight fraction digits, floored towards minus infinity on the
 * double's shortest decimal form, no trailing zeros, "-0" for -0.0.  One line per region of `in`, in row order (StoreTABRD writes one
 * file per sample id: with rows grouped by sample a file is one contiguous byte range of the result).  String-valued columns stay on
 * the host, as everywhere in this ABI; value_cols lists the numeric columns of `in` to print, in output order.
 *
 * Result: GMQL_COL_TXT_BYTES uint8 [total] (the text), GMQL_COL_TXT_LINE_OFFSET int64 [n + 1] (byte offset of every line; the last
 * entry is the total), GMQL_COL_TXT_STATUS uint8 [n]: bit 0 = some numeric field of the line is left to the host's own DecimalFormat
 * and stands in the text as the single byte 0x1A — non-integers from 2^26 on, magnitudes from 2^53 on, NaN and infinities (below
 * 2^26 the digits are derived exactly, see gmql_text.cu).  Reading the three columns back is the D2H of the MATERIALIZE step.
 */
typedef struct {
  int32_t one_based;          /* GMQLSchemaCoordinateSystem.OneBased: print start + 1 */
  int32_t n_values;           /* numeric value columns to print (<= 16) */
  const int32_t* value_cols;  /* HOST [n_values] indices into in->cols */
} gmql_text_out_schema;
int gmql_format_text(gmql_ctx* ctx, const gmql_regions* in, const gmql_text_out_schema* schema, const char* const* chrom_names, gmql_result** out);

/* ---- results --------------------------------------------------------------------------- */
enum {
  GMQL_COL_MAP_ROW_OFFSETS = 1, GMQL_COL_MAP_REF_ROW = 2, GMQL_COL_MAP_REF_SAMPLE_OFFSETS = 3, GMQL_COL_MAP_COUNT = 4,
  GMQL_COL_BAG_OFFSETS = 5, GMQL_COL_BAG_ROWS = 6,
  GMQL_COL_COV_GROUP = 16, GMQL_COL_COV_CHROM = 17, GMQL_COL_COV_START = 18, GMQL_COL_COV_STOP = 19,
  GMQL_COL_COV_STRAND = 20, GMQL_COL_COV_ACC = 21, GMQL_COL_COV_JACC_INTERSECT = 22, GMQL_COL_COV_JACC_RESULT = 23,
  GMQL_COL_COV_EDGE = 24, GMQL_COL_COV_RAW_COUNT = 25, GMQL_COL_COV_RAW_START = 26, GMQL_COL_COV_RAW_STOP = 27,
  GMQL_COL_COV_RAW_SMIN = 28, GMQL_COL_COV_RAW_EMAX = 29, GMQL_COL_COV_RAW_SMAX = 30, GMQL_COL_COV_RAW_EMIN = 31,
  GMQL_COL_JOIN_LEFT_ROW = 32, GMQL_COL_JOIN_RIGHT_ROW = 33, GMQL_COL_JOIN_PAIR = 34, GMQL_COL_JOIN_START = 35,
  GMQL_COL_JOIN_STOP = 36, GMQL_COL_JOIN_STRAND = 37, GMQL_COL_JOIN_CHROM = 38,
  GMQL_COL_DIFF_REF_ROW = 48,
  GMQL_COL_PROF_LEFTMOST = 64, GMQL_COL_PROF_RIGHTMOST = 65, GMQL_COL_PROF_MIN_LENGTH = 66, GMQL_COL_PROF_MAX_LENGTH = 67,
  GMQL_COL_PROF_SUM_LENGTH = 68, GMQL_COL_PROF_SUM_LENGTH_SQ = 69, GMQL_COL_PROF_COUNT = 70,
  GMQL_COL_TXT_LINE_OFFSET = 80, GMQL_COL_TXT_STATUS = 81, GMQL_COL_TXT_CHROM = 82, GMQL_COL_TXT_START = 83, GMQL_COL_TXT_STOP = 84,
  GMQL_COL_TXT_STRAND = 85, GMQL_COL_TXT_BYTES = 86,
  GMQL_COL_GROUP_ROW = 96, GMQL_COL_GROUP_COUNT = 97,
  GMQL_COL_AGG_VALUE = 256, /* + aggregate index */
  GMQL_COL_AGG_VALID = 512  /* + aggregate index */
};
int64_t gmql_result_rows(const gmql_result* r);                 /* output rows (regions / pairs) */
/* element count and element size of a column; <0 if the result has no such column */
int64_t gmql_result_column_len(const gmql_result* r, int32_t col);
int32_t gmql_result_column_elem_size(const gmql_result* r, int32_t col);
/* device pointer of a column (valid until gmql_result_free), NULL if absent */
const void* gmql_result_column_device(const gmql_result* r, int32_t col);
/* copy a column to host memory (dst must hold len*elem_size bytes); synchronises the context stream */
int  gmql_result_column_copy(gmql_ctx* ctx, const gmql_result* r, int32_t col, void* dst);
/* COVER results only: a device-located gmql_regions view (sample = group index, one value column = AccIndex)
 * so that `C = COVER(..) E; M = MAP() C E` chains without a host round trip (the analogue of the reference's
 * per-node memo, GMQLSparkExecutor.scala:252-253).  The small host arrays the view points at (sample_ids, cols)
 * are owned by the result and live until gmql_result_free.  The view carries chrom_span_hint and, for a single
 * (group, strand) class, sorted_by_position; when the strands were unified (every row '*') it has no strand column
 * (strand = NULL means exactly that). */
int  gmql_result_as_regions(gmql_ctx* ctx, gmql_result* r, gmql_regions* view);
void gmql_result_free(gmql_ctx* ctx, gmql_result* r);

/* device time (ms, CUDA events on the context stream) of the last operator call: first kernel -> last kernel,
 * host<->device copies excluded (BASELINE.md §2 t_device), and the H2D part separately */
double gmql_last_device_ms(const gmql_ctx* ctx);
double gmql_last_h2d_ms(const gmql_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GMQL_CUDA_H */
