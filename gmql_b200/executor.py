"""GMQLCudaExecutor — host-side mirror of the three operator entry points the CUDA back end replaces.

Reference interface being mirrored (same argument meaning, same error behaviour: every failure raises):
  GenometricMap71.apply(executor, grouping, aggregator, reference, experiments, BINNING_PARAMETER, REF_PARALLILISM, sc)
      RegionsOperators/GenometricMap/GenometricMap71.scala:28
  GenometricCover.apply(executor, coverFlag, min, max, aggregators, grouping, inputDataset, binSize, sc)
      RegionsOperators/GenometricCover/GenometricCover.scala:30
  GenometricJoin.apply(executor, metajoinCondition, distanceJoinCondition, regionBuilder, leftDataset, rightDataset,
                       joinOnAttributes, BINNING_PARAMETER, MAXIMUM_DISTANCE, sc)
      RegionsOperators/GenometricJoin.scala:21-30
with the executor constants of GMQLSparkExecutor(binSize = BinSize(), maxBinDistance = 100000, ...)
(GMQLSparkExecutor.scala:43-47).  The metadata-driven inputs (`grouping` -> sample pairs, `groups` -> sample->group)
arrive already evaluated, as they do from implement_mjd / implement_mgd.

All arithmetic happens in libgmqlcuda.so; this module only packs columns, resolves thresholds/ids and re-attaches
host-side (string) values.  No CPU fallback exists.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .ids import pair_id
from .params import Agg, BinSize, CoverFlag, RegionBuilder, atom_tuple
from .regions import RegionDataset

INT_MAX = 2 ** 31 - 1


class _Result:
    """Owns a gmql_result handle and copies columns out on demand."""

    def __init__(self, ex, handle):
        self.ex, self.h = ex, handle

    def rows(self):
        return int(self.ex.lib.gmql_result_rows(self.h))

    def has(self, col):
        return self.ex.lib.gmql_result_column_len(self.h, col) >= 0

    def column(self, col, dtype):
        n = self.ex.lib.gmql_result_column_len(self.h, col)
        if n < 0:
            raise KeyError(col)
        out = np.empty(int(n), dtype=dtype)
        assert out.itemsize == self.ex.lib.gmql_result_column_elem_size(self.h, col)
        self.ex._check(self.ex.lib.gmql_result_column_copy(self.ex.ctx, self.h, col, out.ctypes.data))
        return out

    def device_ptr(self, col):
        return self.ex.lib.gmql_result_column_device(self.h, col)

    def free(self):
        if self.h:
            self.ex.lib.gmql_result_free(self.ex.ctx, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GMQLCudaExecutor:
    def __init__(self, binSize=BinSize(), maxBinDistance=100000, device=0, stream=None, narrow=False):
        self.binSize = binSize
        self.narrow = narrow     # compact ABI encodings (8/16-bit chrom, sample_offsets) where they apply
        self.maxBinDistance = maxBinDistance
        self.lib = L.load_library()
        ctx = C.c_void_p()
        rc = self.lib.gmql_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(ctx))
        if rc != 0:
            raise L.GmqlCudaError(rc, "cannot create a CUDA context (no CUDA device? this back end has no CPU fallback)")
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.gmql_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise L.GmqlCudaError(rc, self.lib.gmql_last_error(self.ctx).decode("utf-8", "replace"))

    def set_option(self, name, value):
        """gmql_ctx_set_option: per-context test / tuning hooks and "defer_check" (include/gmql_cuda.h)."""
        self._check(self.lib.gmql_ctx_set_option(self.ctx, name.encode(), None if value is None else str(value).encode()))

    def last_device_ms(self):
        return float(self.lib.gmql_last_device_ms(self.ctx))

    def launch_count(self):
        return int(self.lib.gmql_ctx_launch_count(self.ctx))

    # ---- helpers ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _c_aggs(aggs, ds):
        colmap = ds.numeric_col_map()
        arr = (L.Agg * max(1, len(aggs)))()
        for i, a in enumerate(aggs):
            arr[i].kind = a.kind
            if a.function_identifier.upper() in ("COUNT", "BAG", "BAGD"):
                arr[i].col = 0      # BAG / BAGD: the device emits match lists, the strings are built here (any column type)
            else:
                if a.index not in colmap:
                    raise L.GmqlCudaError(-2, "aggregate %s on a non-numeric column (index %d) cannot run on the device" % (a.function_identifier, a.index))
                arr[i].col = colmap[a.index]
        return arr

    @staticmethod
    def _c_pairs(pairs, left, right):
        lidx = {int(s): i for i, s in enumerate(left.sample_ids)}
        ridx = {int(s): i for i, s in enumerate(right.sample_ids)}
        use = [(l, r) for (l, r) in pairs if l in lidx and r in ridx]
        arr = (L.Pair * max(1, len(use)))()
        for i, (l, r) in enumerate(use):
            arr[i].left_sample = lidx[l]
            arr[i].right_sample = ridx[r]
        return arr, use

    # ---- MAP --------------------------------------------------------------------------------------------------------------
    def map_raw(self, pairs, aggregator, reference, experiments):
        """Runs gmql_map; returns (result handle wrapper, used pair list).  pairs: [(ref_sample_id, exp_sample_id)]."""
        names = sorted(set(reference.chrom_names) | set(experiments.chrom_names))
        ref = reference.with_chrom_names(names)
        exp = experiments.with_chrom_names(names)
        c_ref, k1 = ref.as_c(dup_of=ref.dup_of(), narrow=self.narrow)
        c_exp, k2 = exp.as_c(narrow=self.narrow)
        c_pairs, used = self._c_pairs(pairs, ref, exp)
        c_aggs = self._c_aggs(aggregator, exp)
        out = C.c_void_p()
        self._check(self.lib.gmql_map(self.ctx, C.byref(c_ref), C.byref(c_exp), c_pairs, len(used), c_aggs, len(aggregator), C.byref(out)))
        del k1, k2
        return _Result(self, out), used, ref

    def GenometricMap71(self, grouping, aggregator, reference, experiments):
        """MAP.  grouping: list of (ref_sample_id, exp_sample_id) pairs (implement_mjd flattened), or None for every
        ref sample x every exp sample (MetaJoinMJD2.scala:105-109).  Returns the output as records
        (new_id, chrom, start, stop, strand, refValues ++ [count] ++ aggregates), one per (pair, distinct ref region)."""
        if grouping is None:
            grouping = [(int(a), int(b)) for a in reference.sample_ids for b in experiments.sample_ids]
        res, used, ref = self.map_raw(grouping, aggregator, reference, experiments)
        row_off = res.column(L.COL_MAP_ROW_OFFSETS, np.int64)
        ref_rows = res.column(L.COL_MAP_REF_ROW, np.int32)
        s_off = res.column(L.COL_MAP_REF_SAMPLE_OFFSETS, np.int64)
        count = res.column(L.COL_MAP_COUNT, np.int32)
        vals = [res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(len(aggregator))]
        oks = [res.column(L.COL_AGG_VALID + k, np.uint8) for k in range(len(aggregator))]
        bag_slots = [q for q, a in enumerate(aggregator) if a.function_identifier.upper() in ("BAG", "BAGD")]
        if bag_slots:
            from .javafmt import bag_fold
            bag_off = res.column(L.COL_BAG_OFFSETS, np.int64)
            bag_rows = res.column(L.COL_BAG_ROWS, np.int32)
            exp_records = experiments.with_chrom_names(sorted(set(reference.chrom_names) | set(experiments.chrom_names))).to_records()
        res.free()
        ref_records = ref.to_records()
        lidx = {int(s): i for i, s in enumerate(ref.sample_ids)}
        out = []
        for p, (l, r) in enumerate(used):
            new_id = pair_id(l, r)
            a = lidx[l]
            rows = ref_rows[s_off[a]:s_off[a + 1]]
            base = int(row_off[p])
            for k, row in enumerate(rows):
                c = int(count[base + k])
                if c < 0:
                    continue  # collapsed duplicate (MapKey)
                rec = ref_records[row]
                aggs = [(float(vals[q][base + k]) if oks[q][base + k] else None) for q in range(len(aggregator))]
                for q in bag_slots:      # matched experiment rows in input order, folded pairwise as the operators do
                    rows_q = np.sort(bag_rows[bag_off[base + k]:bag_off[base + k + 1]])
                    aggs[q] = bag_fold([exp_records[r][5][aggregator[q].index] for r in rows_q], aggregator[q].function_identifier.upper() == "BAGD") if c > 0 else None
                aggs = tuple(aggs)
                out.append((new_id, rec[1], rec[2], rec[3], rec[4], tuple(rec[5]) + (float(c),) + aggs))
        return out

    # ---- DIFFERENCE ---------------------------------------------------------------------------------------------------------
    def GenometricDifference(self, grouping, leftDataset, rightDataset, exact=False):
        """DIFFERENCE (GenometricDifference.apply, RegionsOperators/GenometricDifference.scala:23).  grouping: list of
        (left_sample_id, right_sample_id) pairs (implement_mjd flattened) or None for all pairs.  Returns the surviving
        left records as (md5(left sample id), chrom, start, stop, strand, values)."""
        from .ids import long_id
        if grouping is None:
            grouping = [(int(a), int(b)) for a in leftDataset.sample_ids for b in rightDataset.sample_ids]
        names = sorted(set(leftDataset.chrom_names) | set(rightDataset.chrom_names))
        ref = leftDataset.with_chrom_names(names)
        exp = rightDataset.with_chrom_names(names)
        c_ref, k1 = ref.as_c(dup_of=ref.dup_of(), narrow=self.narrow)
        c_exp, k2 = exp.as_c(narrow=self.narrow)
        c_pairs, used = self._c_pairs(grouping, ref, exp)
        out = C.c_void_p()
        self._check(self.lib.gmql_difference(self.ctx, C.byref(c_ref), C.byref(c_exp), c_pairs, len(used), 1 if exact else 0, C.byref(out)))
        del k1, k2
        res = _Result(self, out)
        rows = res.column(L.COL_DIFF_REF_ROW, np.int32)
        res.free()
        recs = ref.to_records()
        return [(long_id(recs[i][0]), recs[i][1], recs[i][2], recs[i][3], recs[i][4], tuple(recs[i][5])) for i in rows]

    # ---- GROUP by coordinates, MERGE ------------------------------------------------------------------------------------------
    def GroupRD(self, groupingParameters, aggregates, regionDataset):
        """GROUP (GroupRD.apply, RegionsOperators/GroupRD.scala:22-64).  groupingParameters: list of value positions (FIELD(pos))
        or None; aggregates: list of Agg or None.  Records of one sample with equal (chrom, start, stop, strand) and equal text
        forms of the grouping fields collapse into one: (id, chrom, start, stop, strand, groupingFields ++ aggregates); every
        aggregate sees the group's whole value list (MEDIAN is the true lower median here)."""
        from .javafmt import bag_text, gvalue_text
        fields = list(groupingParameters or [])
        aggregates = list(aggregates or [])
        recs = regionDataset.to_records()
        code = None
        if fields and recs:
            seen = {}
            code = np.array([seen.setdefault(tuple(gvalue_text(r[5][p]) for p in fields), len(seen)) for r in recs], dtype=np.int64)
        c_ds, keep = regionDataset.as_c(narrow=self.narrow)
        c_aggs = self._c_aggs(aggregates, regionDataset)
        out = C.c_void_p()
        self._check(self.lib.gmql_group(self.ctx, C.byref(c_ds), code.ctypes.data if code is not None else None, c_aggs, len(aggregates), C.byref(out)))
        del keep
        res = _Result(self, out)
        rows = res.column(L.COL_GROUP_ROW, np.int32)
        bag_slots = [q for q, a in enumerate(aggregates) if a.function_identifier.upper() in ("BAG", "BAGD")]
        vals = [None if q in bag_slots else res.column(L.COL_AGG_VALUE + q, np.float64) for q in range(len(aggregates))]
        oks = [None if q in bag_slots else res.column(L.COL_AGG_VALID + q, np.uint8) for q in range(len(aggregates))]
        if bag_slots:
            bag_off = res.column(L.COL_BAG_OFFSETS, np.int64)
            bag_rows = res.column(L.COL_BAG_ROWS, np.int32)
        res.free()
        outp = []
        for g, row in enumerate(rows):
            rec = recs[row]
            aggs = []
            for q, a in enumerate(aggregates):
                if q in bag_slots:
                    members = bag_rows[bag_off[g]:bag_off[g + 1]]
                    aggs.append(bag_text([recs[m][5][a.index] for m in members], a.function_identifier.upper() == "BAGD"))
                else:
                    aggs.append(float(vals[q][g]) if oks[q][g] else None)
            outp.append((rec[0], rec[1], rec[2], rec[3], rec[4], tuple(rec[5][p] for p in fields) + tuple(aggs)))
        return outp

    @staticmethod
    def MergeRD(groups, regionDataset):
        """MERGE (MergeRD.apply, RegionsOperators/MergeRD.scala:20-45): every record keeps its coordinates and values and takes its
        sample's group id (samples without a group are dropped by the join, :41-44), or 1L without grouping (:33-36).  Nothing
        is computed per region, so this stays on the host: a relabelling of the sample ids."""
        out = []
        for (sid, chrom, start, stop, strand, vals) in regionDataset.to_records():
            if groups is None:
                out.append((1, chrom, start, stop, strand, tuple(vals)))
            elif sid in groups:
                out.append((groups[sid], chrom, start, stop, strand, tuple(vals)))
        return out

    # ---- TEXT -> columns ---------------------------------------------------------------------------------------------------
    def parse_text(self, text, parser="NarrowPeakParser", chrom_names=None, start_minus_one=False):
        """Parses one sample file's tab-separated text on the device (gmql_parse_text; BedParser.region_parser,
        loaders/BedParser.scala:112-172).  parser: a key of regions.PARSERS or a tuple (chr, start, stop, strand|None,
        [(pos, "DOUBLE"|"STRING"), ...]).  Returns a dict of numpy columns: line_offset, status, chrom, start, stop, strand,
        values (list of (float64[n], uint8[n]) for the DOUBLE columns, in schema order) and chrom_names.  String columns
        stay in the text (line_offset locates them)."""
        from .regions import PARSERS
        chr_pos, start_pos, stop_pos, strand_pos, other = PARSERS[parser] if isinstance(parser, str) else parser
        data = text if isinstance(text, (bytes, bytearray)) else text.encode("utf-8")
        if chrom_names is None:      # the caller normally knows the assembly; otherwise one cheap host scan of the first column
            chrom_names = sorted({ln.split(b"\t", 1)[0].strip().decode() for ln in data.split(b"\n") if ln and not ln.startswith(b"#")})
        vpos = np.array([p for (p, typ) in other if typ != "STRING"], dtype=np.int32)
        sch = L.TextSchema()
        sch.chr_pos, sch.start_pos, sch.stop_pos = chr_pos, start_pos, stop_pos
        sch.strand_pos = -1 if strand_pos is None else strand_pos
        sch.n_values = len(vpos)
        sch.value_pos = vpos.ctypes.data if len(vpos) else None
        sch.start_minus_one = 1 if start_minus_one else 0
        names = (C.c_char_p * max(1, len(chrom_names)))(*[n.encode() for n in chrom_names])
        out = C.c_void_p()
        self._check(self.lib.gmql_parse_text(self.ctx, bytes(data), len(data), 0, C.byref(sch), names, len(chrom_names), C.byref(out)))
        res = _Result(self, out)
        cols = dict(line_offset=res.column(L.COL_TXT_LINE_OFFSET, np.int64), status=res.column(L.COL_TXT_STATUS, np.uint8),
                    chrom=res.column(L.COL_TXT_CHROM, np.int32), start=res.column(L.COL_TXT_START, np.int64),
                    stop=res.column(L.COL_TXT_STOP, np.int64), strand=res.column(L.COL_TXT_STRAND, np.int8),
                    values=[(res.column(L.COL_AGG_VALUE + k, np.float64), res.column(L.COL_AGG_VALID + k, np.uint8)) for k in range(len(vpos))],
                    chrom_names=list(chrom_names))
        res.free()
        return cols

    # ---- columns -> TEXT ---------------------------------------------------------------------------------------------------
    def format_text(self, ds, value_cols=None, one_based=False, host_format=None):
        """The region lines StoreTABRD writes for `ds` (SelectRegions/StoreTABRD.scala:62-80), built on the device
        (gmql_format_text): returns (bytes, line_offsets int64[n + 1], status uint8[n]).  value_cols: indices of the numeric
        columns to print (default: all).  Fields the device leaves to the host's DecimalFormat (status bit 0; placeholder byte
        0x1A) are filled in with host_format(value) when given (the JVM side passes GDouble.toString; tests pass the oracle's)."""
        c_ds, keep = ds.as_c(narrow=False)
        nums = [c for c in ds.columns if c[0] == "num"]
        vc = np.arange(len(nums), dtype=np.int32) if value_cols is None else np.asarray(value_cols, dtype=np.int32)
        sch = L.TextOutSchema()
        sch.one_based, sch.n_values = (1 if one_based else 0), len(vc)
        sch.value_cols = vc.ctypes.data if len(vc) else None
        names = (C.c_char_p * max(1, len(ds.chrom_names)))(*([n.encode() for n in ds.chrom_names] or [b""]))
        out = C.c_void_p()
        self._check(self.lib.gmql_format_text(self.ctx, C.byref(c_ds), C.byref(sch), names, C.byref(out)))
        res = _Result(self, out)
        text, off, status = res.column(L.COL_TXT_BYTES, np.uint8).tobytes(), res.column(L.COL_TXT_LINE_OFFSET, np.int64), res.column(L.COL_TXT_STATUS, np.uint8)
        res.free()
        if host_format is not None and status.any():
            lines = []
            pos = 0
            for i in np.flatnonzero(status & 1):
                lines.append(text[pos:off[i]])
                fields = text[off[i]:off[i + 1] - 1].split(b"\t")
                for k, col in enumerate(vc):
                    if fields[4 + k] == b"\x1a":
                        fields[4 + k] = host_format(float(nums[col][1][i])).encode()
                lines.append(b"\t".join(fields) + b"\n")
                pos = off[i + 1]
            lines.append(text[pos:])
            text = b"".join(lines)
            ln = np.frombuffer(text, dtype=np.uint8) == 10
            off = np.concatenate([[0], np.flatnonzero(ln) + 1]).astype(np.int64)
        return text, off, status

    # ---- PROFILE ----------------------------------------------------------------------------------------------------------
    def profile(self, ds):
        """Per-sample statistics of Profiler.profile (GMQL-Profiler/.../Profilers/Profiler.scala:124-153): returns
        {sample_id: {leftMost, rightMost, minLength, maxLength, count, avgLength, varianceLength}} (calculateResult :84-92)."""
        c_ds, keep = ds.as_c(narrow=self.narrow)
        out = C.c_void_p()
        self._check(self.lib.gmql_profile(self.ctx, C.byref(c_ds), C.byref(out)))
        del keep
        res = _Result(self, out)
        cols = {k: res.column(c, np.int64) for k, c in (("l", L.COL_PROF_LEFTMOST), ("r", L.COL_PROF_RIGHTMOST), ("mn", L.COL_PROF_MIN_LENGTH),
                                                        ("mx", L.COL_PROF_MAX_LENGTH), ("sl", L.COL_PROF_SUM_LENGTH), ("sq", L.COL_PROF_SUM_LENGTH_SQ),
                                                        ("c", L.COL_PROF_COUNT))}
        res.free()
        prof = {}
        for i, sid in enumerate(ds.sample_ids):
            c = int(cols["c"][i])
            if c > 0:
                mean = float(int(cols["sl"][i])) / c
                var = float(int(cols["sq"][i])) / c - mean * mean
                prof[int(sid)] = dict(leftMost=int(cols["l"][i]), rightMost=int(cols["r"][i]), minLength=int(cols["mn"][i]), maxLength=int(cols["mx"][i]),
                                      count=c, avgLength=mean, varianceLength=var)
            else:
                prof[int(sid)] = dict(leftMost=0, rightMost=0, minLength=0, maxLength=0, count=0, avgLength=0.0, varianceLength=0.0)
        return prof

    # ---- COVER ------------------------------------------------------------------------------------------------------------
    def cover_raw(self, coverFlag, min, max, aggregators, groups, ds, force_unified=False):
        c_ds, keep = ds.as_c(narrow=self.narrow)
        n_s = int(ds.sample_ids.shape[0])
        if groups is not None:
            gids = sorted(set(groups.values()))
            gpos = {g: i for i, g in enumerate(gids)}
            try:
                sample_group = np.array([gpos[groups[int(s)]] for s in ds.sample_ids], dtype=np.int32)
            except KeyError as e:  # plain Map.apply in assignGroups throws (GenometricCover.scala:341)
                raise L.GmqlCudaError(-1, "sample %s has no group" % e)
            sizes = {}
            for _s, g in groups.items():
                sizes[g] = sizes.get(g, 0) + 1
        else:
            gids = [0]
            sample_group = None
            sizes = None
        # thresholds (GenometricCover.scala:57-79)
        if min.kind == "ALL" or max.kind == "ALL":
            if sizes is not None:
                all_value = {g: sizes[g] for g in gids}
            else:
                cnt = C.c_int32()
                self._check(self.lib.gmql_cover_count_samples(self.ctx, C.byref(c_ds), C.byref(cnt)))
                all_value = {0: int(cnt.value)}
        else:
            all_value = {g: 0 for g in gids}
        mn = np.array([(min.fun(all_value[g]) if min.kind == "ALL" else (1 if min.kind == "ANY" else min.n)) for g in gids], dtype=np.int32)
        mx = np.array([(max.fun(all_value[g]) if max.kind == "ALL" else (INT_MAX if max.kind == "ANY" else max.n)) for g in gids], dtype=np.int32)
        c_aggs = self._c_aggs(aggregators, ds)
        out = C.c_void_p()
        self._check(self.lib.gmql_cover(self.ctx, C.byref(c_ds), sample_group.ctypes.data if sample_group is not None else None,
                                        len(gids), CoverFlag.CODE[coverFlag] | (0x100 if force_unified else 0), mn.ctypes.data, mx.ctypes.data,
                                        self.binSize.Cover, c_aggs, len(aggregators), C.byref(out)))
        del keep
        return _Result(self, out), gids

    def cover_shard_raw(self, coverFlag, min_acc, max_acc, ds, chrom_span, first_tile, end_tile):
        """gmql_cover_shard on the rows of `ds` (which must hold the regions of the tiles [first_tile - 1, end_tile), see
        dist.shard_by_tiles): COVER / FLAT of one coordinate-range shard.  Returns a dict of numpy columns (the gmql_cover
        columns plus edge flags and the raw accumulators dist.merge_cover_shards needs)."""
        c_ds, keep = ds.as_c(narrow=self.narrow)
        span = np.ascontiguousarray(chrom_span, dtype=np.int64)
        sh = L.CoverShard()
        sh.chrom_span, sh.first_tile, sh.end_tile = span.ctypes.data, int(first_tile), int(end_tile)
        out = C.c_void_p()
        self._check(self.lib.gmql_cover_shard(self.ctx, C.byref(c_ds), CoverFlag.CODE[coverFlag], int(min_acc), int(max_acc), self.binSize.Cover, C.byref(sh), C.byref(out)))
        del keep
        res = _Result(self, out)
        cols = {k: res.column(c, dt) for k, c, dt in (
            ("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64), ("acc", L.COL_COV_ACC, np.int32),
            ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64), ("edge", L.COL_COV_EDGE, np.int32), ("cnt", L.COL_COV_RAW_COUNT, np.int32),
            ("rstart", L.COL_COV_RAW_START, np.int64), ("rstop", L.COL_COV_RAW_STOP, np.int64), ("smin", L.COL_COV_RAW_SMIN, np.int64),
            ("emax", L.COL_COV_RAW_EMAX, np.int64), ("smax", L.COL_COV_RAW_SMAX, np.int64), ("emin", L.COL_COV_RAW_EMIN, np.int64))}
        res.free()
        return cols

    def GenometricCover(self, coverFlag, min, max, aggregators, groups, ds, force_unified=False):
        """COVER / FLAT / SUMMIT / HISTOGRAM.  groups: dict sample_id -> group_id (implement_mgd) or None.  Returns records
        (group_id, chrom, start, stop, strand, (AccIndex, JaccardIntersect, JaccardResult, aggregates...)).
        force_unified: the caller sharded the dataset and some OTHER shard holds a '*' region (strand unification is global)."""
        res, gids = self.cover_raw(coverFlag, min, max, aggregators, groups, ds, force_unified)
        g = res.column(L.COL_COV_GROUP, np.int32)
        chrom = res.column(L.COL_COV_CHROM, np.int32)
        start = res.column(L.COL_COV_START, np.int64)
        stop = res.column(L.COL_COV_STOP, np.int64)
        strand = res.column(L.COL_COV_STRAND, np.int8)
        acc = res.column(L.COL_COV_ACC, np.int32)
        j1 = res.column(L.COL_COV_J1, np.float64)
        j2 = res.column(L.COL_COV_J2, np.float64)
        vals = [res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(len(aggregators))]
        oks = [res.column(L.COL_AGG_VALID + k, np.uint8) for k in range(len(aggregators))]
        # BAG / BAGD: the device lists the contributing input rows of every output region (GMQL_COL_BAG_*); the strings are folded
        # here, pairwise in input row order, as in GenometricMap71 (the reference's own order is whatever Spark's shuffle produced)
        bag_slots = [q for q, a in enumerate(aggregators) if a.function_identifier.upper() in ("BAG", "BAGD")]
        if bag_slots and len(g):
            from .javafmt import bag_fold
            bag_off = res.column(L.COL_BAG_OFFSETS, np.int64)
            bag_rows = res.column(L.COL_BAG_ROWS, np.int32)
            recs = ds.to_records()
        res.free()
        sc = "*+-"
        out = []
        for i in range(len(g)):
            aggs = [(float(vals[q][i]) if oks[q][i] else None) for q in range(len(aggregators))]
            for q in bag_slots:
                rows_q = np.sort(bag_rows[bag_off[i]:bag_off[i + 1]])
                aggs[q] = bag_fold([recs[r][5][aggregators[q].index] for r in rows_q], aggregators[q].function_identifier.upper() == "BAGD")
            out.append((gids[g[i]], ds.chrom_names[chrom[i]], int(start[i]), int(stop[i]), sc[strand[i]],
                        (float(acc[i]), float(j1[i]), float(j2[i])) + tuple(aggs)))
        return out

    # ---- JOIN -------------------------------------------------------------------------------------------------------------
    def join_raw(self, pairs, atoms, regionBuilder, left, right, left_eq=None, right_eq=None, dedup=True):
        names = sorted(set(left.chrom_names) | set(right.chrom_names))
        l = left.with_chrom_names(names)
        r = right.with_chrom_names(names)
        c_l, k1 = l.as_c(dup_of=l.dup_of() if dedup else None, narrow=self.narrow)
        c_r, k2 = r.as_c(dup_of=r.dup_of() if dedup else None, narrow=self.narrow)
        c_pairs, used = self._c_pairs(pairs, l, r)
        c_atoms = (L.Atom * max(1, len(atoms)))()
        for i, a in enumerate(atoms):
            code, arg, _ = atom_tuple(a)
            c_atoms[i].kind, c_atoms[i].arg = code, arg
        le = np.ascontiguousarray(left_eq, dtype=np.int64) if left_eq is not None else None
        re_ = np.ascontiguousarray(right_eq, dtype=np.int64) if right_eq is not None else None
        out = C.c_void_p()
        self._check(self.lib.gmql_join(self.ctx, C.byref(c_l), C.byref(c_r), c_pairs, len(used), c_atoms, len(atoms),
                                       RegionBuilder.CODE[regionBuilder], self.binSize.Join, self.maxBinDistance,
                                       le.ctypes.data if le is not None else None, re_.ctypes.data if re_ is not None else None,
                                       C.byref(out)))
        del k1, k2
        return _Result(self, out), used, l, r

    def GenometricJoin(self, metajoin_pairs, distanceJoinCondition, regionBuilder, leftDataset, rightDataset, joinOnAttributes=None):
        """JOIN for one JoinQuadruple.  metajoin_pairs: list of (left_sample_id, right_sample_id) or None = all pairs.
        distanceJoinCondition: list of AtomicCondition in script order.  joinOnAttributes: optional list of
        (left_value_index, right_value_index).  Returns built records as joinRegions does (GenometricJoin.scala:345-372)."""
        if metajoin_pairs is None:
            metajoin_pairs = [(int(a), int(b)) for a in leftDataset.sample_ids for b in rightDataset.sample_ids]
        left_eq = right_eq = None
        lrec = leftDataset.to_records()
        rrec = rightDataset.to_records()
        if joinOnAttributes:
            codes = {}
            left_eq = np.array([codes.setdefault(tuple(rec[5][i] for (i, _j) in joinOnAttributes), len(codes)) for rec in lrec], dtype=np.int64)
            right_eq = np.array([codes.get(tuple(rec[5][j] for (_i, j) in joinOnAttributes), -1) for rec in rrec], dtype=np.int64)
        res, used, l, r = self.join_raw(metajoin_pairs, distanceJoinCondition, regionBuilder, leftDataset, rightDataset, left_eq, right_eq)
        lrow = res.column(L.COL_JOIN_LEFT_ROW, np.int32)
        rrow = res.column(L.COL_JOIN_RIGHT_ROW, np.int32)
        pidx = res.column(L.COL_JOIN_PAIR, np.int32)
        start = res.column(L.COL_JOIN_START, np.int64)
        stop = res.column(L.COL_JOIN_STOP, np.int64)
        strand = res.column(L.COL_JOIN_STRAND, np.int8)
        chrom = res.column(L.COL_JOIN_CHROM, np.int32)
        res.free()
        ids = [pair_id(a, b) for (a, b) in used]
        sc = "*+-"
        out = []
        for i in range(len(lrow)):
            lv = tuple(lrec[lrow[i]][5]) if lrow[i] >= 0 else ()
            rv = tuple(rrec[rrow[i]][5]) if rrow[i] >= 0 else ()
            if regionBuilder == RegionBuilder.LEFT_DISTINCT:
                vals = lv
            elif regionBuilder == RegionBuilder.RIGHT_DISTINCT:
                vals = rv
            elif regionBuilder == RegionBuilder.BOTH:
                rr = rrec[rrow[i]]
                vals = lv + (rr[1], float(rr[2]), float(rr[3]), rr[4]) + rv
            else:
                vals = lv + rv
            out.append((ids[pidx[i]], l.chrom_names[chrom[i]], int(start[i]), int(stop[i]), sc[strand[i]], vals))
        return out
