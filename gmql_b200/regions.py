"""RegionDataset — the columnar (structure-of-arrays) image of RDD[GRECORD] that crosses the C ABI.

GRECORD = (GRecordKey(id, chrom, start, stop, strand), Array[GValue])  (core/DataTypes.scala:62, core/GRecordKey.scala:8).
Numeric GValues are float64 with a validity byte (0 = GNull); string columns stay on the host as object arrays and are
re-attached to results by row index.
"""
import ctypes as C

import numpy as np

from . import _lib as L

STRAND_CODE = {"*": 0, "+": 1, "-": 2}
STRAND_CHAR = np.array(["*", "+", "-"])

# the reference's stock parsers (loaders/BedParser.scala:295-303, :346-348, :525-527): (chr, start, stop, strand, [(pos, type)])
PARSERS = {
    "NarrowPeakParser": (0, 1, 2, 5, [(3, "STRING"), (4, "DOUBLE"), (6, "DOUBLE"), (7, "DOUBLE"), (8, "DOUBLE"), (9, "DOUBLE")]),
    "RnaSeqParser": (0, 1, 2, 3, [(4, "STRING"), (5, "DOUBLE")]),
    "BedScoreParser": (0, 1, 2, None, [(3, "DOUBLE")]),
}


class RegionDataset:
    def __init__(self, chrom_names, chrom, start, stop, strand, sample, sample_ids, columns=None):
        self.chrom_names = list(chrom_names)
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int64)
        self.stop = np.ascontiguousarray(stop, dtype=np.int64)
        self.strand = np.ascontiguousarray(strand, dtype=np.int8)
        self.sample = np.ascontiguousarray(sample, dtype=np.int32)
        self.sample_ids = np.ascontiguousarray(sample_ids, dtype=np.int64)
        # columns: list of ("num", float64[n], uint8[n] | None) or ("str", object[n])
        self.columns = list(columns or [])

    @property
    def n(self):
        return int(self.chrom.shape[0])

    # ---- construction ---------------------------------------------------------------------------------------------
    @classmethod
    def from_records(cls, records, sample_ids=None, n_values=None):
        """records: iterable of (sample_id, chrom, start, stop, strand, values) — the oracle's record model."""
        records = list(records)
        if sample_ids is None:
            sample_ids = sorted({r[0] for r in records})
        sidx = {s: i for i, s in enumerate(sample_ids)}
        names = sorted({r[1] for r in records})
        cidx = {c: i for i, c in enumerate(names)}
        n = len(records)
        if n_values is None:
            n_values = len(records[0][5]) if n else 0
        chrom = np.fromiter((cidx[r[1]] for r in records), dtype=np.int32, count=n)
        start = np.fromiter((r[2] for r in records), dtype=np.int64, count=n)
        stop = np.fromiter((r[3] for r in records), dtype=np.int64, count=n)
        strand = np.fromiter((STRAND_CODE[r[4]] for r in records), dtype=np.int8, count=n)
        sample = np.fromiter((sidx[r[0]] for r in records), dtype=np.int32, count=n)
        cols = []
        for k in range(n_values):
            vals = [r[5][k] for r in records]
            if any(isinstance(v, str) for v in vals):
                cols.append(("str", np.array(vals, dtype=object)))
            else:
                valid = np.fromiter((0 if v is None else 1 for v in vals), dtype=np.uint8, count=n)
                num = np.fromiter((0.0 if v is None else v for v in vals), dtype=np.float64, count=n)
                cols.append(("num", num, valid if not valid.all() else None))
        return cls(names, chrom, start, stop, strand, sample, sample_ids, cols)

    @classmethod
    def from_text(cls, files, parser="NarrowPeakParser"):
        """files: list of (file_name, text).  Sample id = md5(file name) as loaders/Loaders.scala:203-208; lines follow
        BedParser.region_parser (BedParser.scala:112-172): '#' lines skipped, strand anything but +/- -> '*',
        'null'/'.' numeric -> GNull."""
        from .ids import sample_id_from_filename
        chr_pos, start_pos, stop_pos, strand_pos, other = PARSERS[parser]
        recs = []
        ids = []
        for name, text in files:
            sid = sample_id_from_filename(name)
            ids.append(sid)
            for line in text.splitlines():
                if not line or line.startswith("#"):
                    continue
                s = line.split("\t")
                vals = []
                for pos, typ in other:
                    x = s[pos]
                    if typ == "STRING":
                        vals.append(x.strip())
                    else:
                        vals.append(None if x.lower() in ("null", ".") else float(x))
                strand = "*"
                if strand_pos is not None:
                    t = s[strand_pos].strip()
                    strand = t if t in ("+", "-") else "*"
                recs.append((sid, s[chr_pos].strip(), int(s[start_pos].strip()), int(s[stop_pos].strip()), strand, tuple(vals)))
        return cls.from_records(recs, sample_ids=ids, n_values=len(other))

    def to_records(self):
        out = []
        cols = []
        for c in self.columns:
            if c[0] == "str":
                cols.append(list(c[1]))
            else:
                num, valid = c[1], c[2]
                if valid is None:
                    cols.append([float(v) for v in num])
                else:
                    cols.append([float(v) if ok else None for v, ok in zip(num, valid)])
        for i in range(self.n):
            out.append((int(self.sample_ids[self.sample[i]]), self.chrom_names[self.chrom[i]], int(self.start[i]),
                        int(self.stop[i]), str(STRAND_CHAR[self.strand[i]]), tuple(c[i] for c in cols)))
        return out

    # ---- helpers for the boundary -------------------------------------------------------------------------------------
    def with_chrom_names(self, names):
        """Re-index chromosomes into a shared dictionary (both operands of MAP/JOIN must use one)."""
        if list(names) == self.chrom_names:
            return self
        pos = {c: i for i, c in enumerate(names)}
        remap = np.array([pos[c] for c in self.chrom_names], dtype=np.int32) if self.chrom_names else np.zeros(0, np.int32)
        chrom = remap[self.chrom] if self.n else self.chrom
        return RegionDataset(names, chrom, self.start, self.stop, self.strand, self.sample, self.sample_ids, self.columns)

    def dup_of(self):
        """dup_of[i] = lowest row whose full record equals row i (MapKey equality, GenometricMap71.scala:186-203);
        None when every record is unique."""
        seen = {}
        out = np.arange(self.n, dtype=np.int32)
        any_dup = False
        cols = []
        for c in self.columns:
            if c[0] == "str":
                cols.append(c[1])
            else:
                v = c[1] if c[2] is None else np.where(c[2] != 0, c[1], np.nan)
                cols.append(v)
        for i in range(self.n):
            key = (int(self.sample[i]), int(self.chrom[i]), int(self.start[i]), int(self.stop[i]), int(self.strand[i]),
                   tuple((None if (isinstance(c[i], float) and c[i] != c[i]) else c[i]) for c in cols))
            j = seen.setdefault(key, i)
            if j != i:
                out[i] = j
                any_dup = True
        return out if any_dup else None

    def numeric_col_map(self):
        """value-column index -> position among the numeric columns passed over the ABI (strings stay on the host)."""
        m = {}
        k = 0
        for idx, c in enumerate(self.columns):
            if c[0] == "num":
                m[idx] = k
                k += 1
        return m

    def as_c(self, dup_of=None, coord32=None, narrow=False, sorted_by_position=False, chrom_span_hint=None):
        """Builds the gmql_regions struct over this dataset's numpy buffers.  Returns (struct, keepalive).
        narrow=True uses the ABI's compact encodings where they apply: 8/16-bit chromosome indices and, when the rows are
        ordered by sample, sample_offsets instead of a per-row sample column."""
        keep = []
        n = self.n
        if coord32 is None:
            coord32 = n == 0 or (int(self.stop.max(initial=0)) < 2 ** 31 and int(self.start.max(initial=0)) < 2 ** 31
                                 and int(self.start.min(initial=0)) >= -(2 ** 31))
        if coord32:
            start = self.start.astype(np.int32)
            stop = self.stop.astype(np.int32)
        else:
            start, stop = self.start, self.stop
        keep += [start, stop, self.chrom, self.strand, self.sample, self.sample_ids]
        r = L.Regions()
        r.n = n
        r.location = 0
        r.coord_bits = 32 if coord32 else 64
        r.chrom = self.chrom.ctypes.data
        r.start = start.ctypes.data
        r.stop = stop.ctypes.data
        r.strand = self.strand.ctypes.data
        r.sample = self.sample.ctypes.data
        r.n_chrom = max(1, len(self.chrom_names))
        r.n_samples = max(1, int(self.sample_ids.shape[0]))
        r.sample_ids = self.sample_ids.ctypes.data
        if narrow and n:
            if r.n_chrom <= 65536:
                cn = self.chrom.astype(np.uint8 if r.n_chrom <= 256 else np.uint16)
                keep.append(cn)
                r.chrom = cn.ctypes.data
                r.chrom_bits = 8 if r.n_chrom <= 256 else 16
            ln = self.stop - self.start
            if int(ln.min(initial=0)) >= 0 and int(ln.max(initial=0)) < 65536:
                l16 = ln.astype(np.uint16)
                keep.append(l16)
                r.stop = l16.ctypes.data
                r.stop_is_length = 16
            if np.all(np.diff(self.sample) >= 0):
                off = np.searchsorted(self.sample, np.arange(r.n_samples + 1), side="left").astype(np.int64)
                keep.append(off)
                r.sample = None
                r.sample_offsets = off.ctypes.data
        nums = [c for c in self.columns if c[0] == "num"]
        r.n_cols = len(nums)
        if nums:
            cols = (C.c_void_p * len(nums))()
            valid = (C.c_void_p * len(nums))()
            for k, c in enumerate(nums):
                a = np.ascontiguousarray(c[1], dtype=np.float64)
                keep.append(a)
                cols[k] = a.ctypes.data
                if c[2] is not None:
                    v = np.ascontiguousarray(c[2], dtype=np.uint8)
                    keep.append(v)
                    valid[k] = v.ctypes.data
                else:
                    valid[k] = None
            keep += [cols, valid]
            r.cols = C.cast(cols, C.POINTER(C.c_void_p))
            r.valid = C.cast(valid, C.POINTER(C.c_void_p))
        if dup_of is not None:
            d = np.ascontiguousarray(dup_of, dtype=np.int32)
            keep.append(d)
            r.dup_of = d.ctypes.data
        if sorted_by_position:
            r.sorted_by_position = 1
        if chrom_span_hint is not None:
            h = np.ascontiguousarray(chrom_span_hint, dtype=np.int64)
            assert h.shape[0] == r.n_chrom
            keep.append(h)
            r.chrom_span_hint = h.ctypes.data
        return r, keep
