"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A literal, line-by-line Python restatement of GMQL-Spark's region-operator hot path
(MAP, COVER/FLAT/SUMMIT/HISTOGRAM incl. the GMAP4 second phase, genometric JOIN) and of
the aggregate factory, following the reference's *bin-based* algorithm verbatim, plus
bin-free "direct" forms used to cross-check the reading.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; `gmql_b200` never does.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or expected
outputs for this path (no assertions anywhere in the tree; `conf/*.xml`
hold queries and generator parameters only) and cannot be compiled or run here (no JVM,
Scala, Spark or Maven in the image).  What pins this oracle instead:
  * it is a literal transcription (each function cites the file:line it follows; paths are
    relative to /root/reference, RO = GMQL-Spark/src/main/scala/it/polimi/genomics/spark/
    implementation/RegionsOperators, SRV = GMQL-Server/src/main/scala/it/polimi/genomics/GMQLServer);
  * the literal bin-based forms are checked against the independent direct forms on
    randomised inputs with bin-aligned coordinates (tests/test_oracle_selfcheck.py);
  * Guava's md5 `putLong/putString/asLong` semantics (a transitive, un-vendored dependency:
    com.google.guava via spark-core_2.11:2.2.0, no explicit pin in any pom) are restated from
    its published algorithm: little-endian primitive sinks, asLong = first 8 digest bytes LE.

Record model (core/DataTypes.scala:62, core/GRecordKey.scala:8):
    region = (sample_id:int, chrom:str, start:int, stop:int, strand:str in '+-*', values:tuple)
    a value (GValue) is float (GDouble), str (GString) or None (GNull).
"""
from __future__ import annotations

import hashlib
import math
import struct
from collections import defaultdict
from decimal import Decimal, ROUND_FLOOR

# --------------------------------------------------------------------------------------
# ids (Guava Hashing.md5 semantics)
# --------------------------------------------------------------------------------------

def md5_pair_id(left_id: int, right_id: int) -> int:
    """Hashing.md5().newHasher().putLong(l).putLong(r).hash().asLong
    (RO/GenometricMap/GenometricMap71.scala:174, RO/GenometricJoin.scala:46)."""
    d = hashlib.md5(struct.pack("<qq", left_id, right_id)).digest()
    return int.from_bytes(d[:8], "little", signed=True)


def md5_string_id(s: str) -> int:
    """Hashing.md5().newHasher().putString(s, UTF_8).hash().asLong
    (loaders/Loaders.scala:203-208 sample id from file name; RO/GenometricMap/GMAP4.scala:32)."""
    d = hashlib.md5(s.encode("utf-8")).digest()
    return int.from_bytes(d[:8], "little", signed=True)


def sample_id_from_filename(name: str) -> int:
    """loaders/Loaders.scala:203-208: strip a trailing '.meta', drop '/', md5 -> long."""
    ext = name[name.rfind(".") + 1:]
    base = name if ext != "meta" else name[: name.rfind(".")]
    return md5_string_id(base.replace("/", ""))


# --------------------------------------------------------------------------------------
# GValue helpers (core/DataTypes.scala:76-194)
# --------------------------------------------------------------------------------------

_WIDE_CONTEXT = __import__("decimal").Context(prec=400)


def gdouble_to_string(v: float) -> str:
    """GDouble.toString: DecimalFormat("#.########", ENGLISH), RoundingMode.FLOOR
    (core/DataTypes.scala:139-148)."""
    if v != v:
        return "NaN"  # DecimalFormat prints U+FFFD for NaN; never produced on this path
    if math.isinf(v):
        return "∞" if v > 0 else "-∞"
    # (a context wide enough for the 309 integer digits of the largest double)
    q = Decimal(repr(v)).quantize(Decimal("0.00000001"), rounding=ROUND_FLOOR, context=_WIDE_CONTEXT)
    s = format(q, "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    if s in ("-0", ""):
        s = "-0" if s == "-0" else "0"
    return s


def gvalue_to_string(v) -> str:
    if v is None:
        return "null"  # GNull.toString, DataTypes.scala:185
    if isinstance(v, float):
        return gdouble_to_string(v)
    return str(v)


def store_tab_line(rec, one_based=False) -> str:
    """One region line of StoreTABRD (SelectRegions/StoreTABRD.scala:62-80): `(recordKey.productIterator.drop(1) ++ values)
    .mkString("\\t")` — chrom, start (+ 1 for one-based schemas, :62-69), stop, strand, then every GValue's toString."""
    _sid, chrom, start, stop, strand, vals = rec
    return "\t".join([chrom, str(start + (1 if one_based else 0)), str(stop), strand] + [gvalue_to_string(v) for v in vals])


def _gvalue_cmp_key(v):
    """GValue ordering: GDouble > GString > GNull (DataTypes.scala:77-103)."""
    if v is None:
        return (0, 0.0, "")
    if isinstance(v, str):
        return (1, 0.0, v)
    return (2, v, "")


def _java_min(a: float, b: float) -> float:
    """java.lang.Math.min(double,double): NaN propagates, -0.0 < +0.0."""
    if a != a:
        return a
    if a == 0.0 and b == 0.0 and math.copysign(1.0, b) < 0:
        return b
    return a if a <= b else b


def _java_max(a: float, b: float) -> float:
    if a != a:
        return a
    if a == 0.0 and b == 0.0 and math.copysign(1.0, a) < 0:
        return b
    return a if a >= b else b


def _java_double_to_string(v: float) -> str:
    """java.lang.Double.toString (BAG / BAGD, DefaultRegionsToRegionFactory.scala:139): decimal notation for
    1e-3 <= |v| < 1e7, else d.dddE[-]n, shortest identifying digits (extreme subnormals, where the JDK prints more digits
    than necessary, excepted)."""
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    d = Decimal(repr(abs(v)))
    sign, digs, exp = d.as_tuple()
    digs = list(digs)
    while len(digs) > 1 and digs[-1] == 0:
        digs.pop()
        exp += 1
    ds = "".join(map(str, digs))
    point = len(ds) + exp                      # position of the decimal point relative to the digit string
    neg = "-" if v < 0 else ""
    if 1e-3 <= abs(v) < 1e7:
        if point <= 0:
            return neg + "0." + "0" * (-point) + ds
        if point >= len(ds):
            return neg + ds + "0" * (point - len(ds)) + ".0"
        return neg + ds[:point] + "." + ds[point:]
    return "%s%s.%sE%d" % (neg, ds[0], ds[1:] or "0", point - 1)


# --------------------------------------------------------------------------------------
# aggregates (SRV/DefaultRegionsToRegionFactory.scala)
# --------------------------------------------------------------------------------------

class Agg:
    """RegionsToRegion (core/DataStructures/RegionAggregate/RegionsToRegion.scala:9-20):
    `fun` folds a List[GValue] (applied pairwise by the operators), `fun_out` finalises with
    (count, nonNullCount)."""

    def __init__(self, kind: str, index: int = 0):
        self.kind = kind.upper()
        self.index = 0 if self.kind == "COUNT" else index  # COUNT.index = 0 (:60)

    def _doubles(self, line):
        return [v for v in line if isinstance(v, float)]

    def fun(self, line):
        k = self.kind
        if k == "COUNT":  # :63-66
            return float(len(line))
        if k in ("SUM", "AVG"):  # :74-81, :118-125
            ss = self._doubles(line)
            if ss:
                acc = ss[0]
                for x in ss[1:]:
                    acc = acc + x
                return acc
            return None
        if k == "MIN":  # :89-95
            ss = self._doubles(line)
            if ss:
                acc = ss[0]
                for x in ss[1:]:
                    acc = _java_min(acc, x)
                return acc
            return None
        if k == "MAX":  # :103-109
            ss = self._doubles(line)
            if ss:
                acc = ss[0]
                for x in ss[1:]:
                    acc = _java_max(acc, x)
                return acc
            return None
        if k == "MEDIAN":  # :38-54 (note values(left)+values(left))
            values = sorted(self._doubles(line))
            if values:
                if len(values) % 2 == 0:
                    left = len(values) // 2 - 1
                    return (values[left] + values[left]) / 2
                return values[len(values) // 2]
            return None
        if k in ("BAG", "BAGD"):  # :133-146, :154-167
            if not line:
                return "."
            items = list(line)
            if k == "BAGD":
                seen = []
                for it in items:
                    if it not in seen:
                        seen.append(it)
                items = seen
            items = sorted(items, key=_gvalue_cmp_key)
            out = []
            for g in items:
                if g is None:
                    out.append(".")
                elif isinstance(g, float):
                    out.append(_java_double_to_string(g))
                else:
                    out.append(g)
            return ",".join(out)
        raise ValueError("No map function with the given name (%s) found." % k)  # :16,:29

    def fun_out(self, v, counts):
        count, nonnull = counts
        k = self.kind
        if k == "COUNT":  # :62
            return float(count)
        if k == "AVG":  # :116-117
            if nonnull > 0 and v is not None:
                return v / nonnull
            return None
        if k == "MEDIAN":  # :37
            return v
        return v if count > 0 else None  # SUM :73, MIN :88, MAX :102, BAG :132, BAGD :153


def _strand_ok(l: str, r: str) -> bool:
    return l == "*" or r == "*" or l == r


# --------------------------------------------------------------------------------------
# MAP — literal (RO/GenometricMap/GenometricMap71.scala)
# --------------------------------------------------------------------------------------

def _reduce_func(aggs, left, right):
    """reduceFunc, GenometricMap71.scala:66-79."""
    lv, lc, lcs = left
    rv, rc, rcs = right
    if lv and rv:
        values = tuple(a.fun([lv[i], rv[i]]) for i, a in enumerate(aggs))
    elif rv:
        values = rv
    else:
        values = lv
    n = max(len(lcs), len(rcs))
    lcs2 = tuple(lcs) + (0,) * (n - len(lcs))
    rcs2 = tuple(rcs) + (0,) * (n - len(rcs))
    return (values, lc + rc, tuple(x + y for x, y in zip(lcs2, rcs2)))


def map_reference(ref, exp, pairs, aggs, bin_size=50000):
    """GenometricMap71.apply/execute (:28-150).  `pairs`: dict ref_sample_id -> list of exp ids
    (implement_mjd flattened then groupByKey, :50-52).  Returns a list of
    (new_id, chrom, start, stop, strand, values) with values = refValues ++ [count] ++ aggs."""
    if bin_size == 0:
        bin_size = 2 ** 63 - 1  # :36-40
    # exp.binDS (:153-165)
    exp_binned = defaultdict(list)
    for (sid, chrom, start, stop, strand, vals) in exp:
        new_val = tuple(vals[a.index] for a in aggs)
        for b in range(start // bin_size, stop // bin_size + 1):
            exp_binned[(sid, chrom, b)].append((start, stop, strand, new_val))
    # ref.binDS (:167-181)
    ref_binned = defaultdict(list)
    for (sid, chrom, start, stop, strand, vals) in ref:
        for exp_id in pairs.get(sid, ()):
            new_id = md5_pair_id(sid, exp_id)
            for b in range(start // bin_size, stop // bin_size + 1):
                ref_binned[(exp_id, chrom, b)].append((new_id, start, stop, strand, tuple(vals)))
    # cogroup + hot loop (:81-121); cogroup keys with an empty ref side emit nothing
    joined = []
    for key, refs in ref_binned.items():
        exps = exp_binned.get(key, ())
        for (new_id, rstart, rstop, rstrand, rvals) in refs:
            map_key = (new_id, key[1], rstart, rstop, rstrand, rvals)
            ref_in_start_bin = (rstart // bin_size) == key[2]
            filt = [e for e in exps
                    if (rstart < e[1] and e[0] < rstop)
                    and _strand_ok(rstrand, e[2])
                    and (ref_in_start_bin or (e[0] // bin_size) == key[2])]
            if filt:
                items = [(e[3], 1, tuple(0 if v is None else 1 for v in e[3])) for e in filt]
                acc = items[0]
                for it in items[1:]:
                    acc = _reduce_func(aggs, acc, it)
                joined.append((map_key, acc))
            elif ref_in_start_bin:
                joined.append((map_key, ((), 0, ())))
    # reduceByKey (:123)
    reduced = {}
    for k, v in joined:
        reduced[k] = _reduce_func(aggs, reduced[k], v) if k in reduced else v
    # output (:125-147)
    out = []
    for (new_id, chrom, rstart, rstop, rstrand, rvals), (values, count, counts) in reduced.items():
        new_val = tuple(
            a.fun_out(values[i] if values else None, (count, counts[i] if counts else 0))
            for i, a in enumerate(aggs))
        out.append((new_id, chrom, rstart, rstop, rstrand, tuple(rvals) + (float(count),) + new_val))
    return out


def map_direct(ref, exp, pairs, aggs):
    """Bin-free form of MAP: direct overlap predicate, MapKey collapse, closed-form
    aggregates folded in exp input order."""
    exp_by = defaultdict(list)
    for e in exp:
        exp_by[(e[0], e[1])].append(e)
    acc = {}
    for (sid, chrom, start, stop, strand, vals) in ref:
        for exp_id in pairs.get(sid, ()):
            key = (md5_pair_id(sid, exp_id), chrom, start, stop, strand, tuple(vals))
            S = [e for e in exp_by.get((exp_id, chrom), ())
                 if start < e[3] and e[2] < stop and _strand_ok(strand, e[4])]
            acc.setdefault(key, []).extend(S)
    out = []
    for key, S in acc.items():
        new_val = []
        for a in aggs:
            col = [e[5][a.index] for e in S]
            nonnull = sum(1 for v in col if v is not None)
            if not col:
                v = None
            else:
                v = col[0]
                for x in col[1:]:
                    v = a.fun([v, x])
            new_val.append(a.fun_out(v, (len(S), nonnull)))
        out.append(key[:5] + (tuple(key[5]) + (float(len(S)),) + tuple(new_val),))
    return out


# --------------------------------------------------------------------------------------
# COVER family — literal (RO/GenometricCover/GenometricCover.scala, RO/GenometricMap/GMAP4.scala)
# --------------------------------------------------------------------------------------

INT_MAX = 2 ** 31 - 1


def group_reference(ds, grouping_fields, aggs):
    """GroupRD.apply (RegionsOperators/GroupRD.scala:22-64), literally: every record's values become one-element lists (:31),
    records are grouped by the md5 of the UNSEPARATED string id + chrom + start + stop + strand (+ "\u00a7" + text of every
    grouping field, :32-42; grouped here by the string itself), a group's lists are concatenated (:44), every aggregate is
    applied ONCE to the whole list of its column with (length, non-null count) (:48-52), and the output values are the first
    record's grouping fields followed by the aggregates (:56-61).  The fold order is Spark's; here: input order."""
    fields = list(grouping_fields or [])
    groups = {}
    for rec in ds:
        sid, chrom, start, stop, strand, vals = rec
        k = "%d%s%d%d%s" % (sid, chrom, start, stop, strand)
        for pos in fields:
            k += "\u00a7" + gvalue_to_string(vals[pos])
        if k not in groups:
            groups[k] = ((sid, chrom, start, stop, strand), [[v] for v in vals])
        else:
            for lst, v in zip(groups[k][1], vals):
                lst.append(v)
    out = []
    for key, lists in groups.values():
        aggregated = []
        for a in (aggs or []):
            line = lists[a.index]
            aggregated.append(a.fun_out(a.fun(line), (len(line), sum(1 for v in line if v is not None))))
        out.append(key + (tuple(lists[pos][0] for pos in fields) + tuple(aggregated),))
    return out


def merge_reference(ds, groups=None):
    """MergeRD.apply (RegionsOperators/MergeRD.scala:20-45): id := group of the sample (inner join, :41-44) or 1L (:33-36)."""
    if groups is None:
        return [(1, c, s, e, st, tuple(v)) for (_sid, c, s, e, st, v) in ds]
    return [(groups[sid], c, s, e, st, tuple(v)) for (sid, c, s, e, st, v) in ds if sid in groups]


def md5_long_id(v: int) -> int:
    """Hashing.md5().newHasher().putLong(v).hash().asLong() (GenometricDifference.scala:53)."""
    return int.from_bytes(hashlib.md5(struct.pack("<q", v)).digest()[:8], "little", signed=True)


def difference_reference(ref, exp, groups, exact=False, bin_size=50000):
    """GenometricDifference.apply (RegionsOperators/GenometricDifference.scala:23-100), restated line by line.
    `groups`: dict ref_sample_id -> list of exp sample ids (implement_mjd, :33-34).  A reference region survives when no
    region of ANY of its paired experiment samples overlaps it (strictly, strand-compatible; `exact`: with identical
    coordinates, :67).  Reference regions of a sample without a group entry (or with an empty one) produce nothing
    (binRefDS :104-114); identical reference records of one sample collapse into one output record (the reduce key
    :54 is built from the record's content; here the tuple itself stands for the md5 of the unseparated string
    concatenation, whose accidental aliasing is not reproduced).  Output id = md5(ref sample id) (:53)."""
    ref_binned = defaultdict(list)                       # binRefDS (:102-114); the .distinct (:37) cannot collapse
    for (sid, chrom, start, stop, strand, vals) in ref:   # records: Array[GValue] has reference equality
        grp = groups.get(sid)
        if grp is None:
            continue
        for b in range(start // bin_size, stop // bin_size + 1):
            for exp_id in grp:
                ref_binned[(exp_id, chrom, b)].append((sid, start, stop, strand, tuple(vals)))
    exp_binned = defaultdict(list)                       # binExpDS (:116-122)
    for (sid, chrom, start, stop, strand, vals) in exp:
        for b in range(start // bin_size, stop // bin_size + 1):
            exp_binned[(sid, chrom, b)].append((start, stop, strand))
    reduced = {}                                         # cogroup (:44-72) + reduceByKey (:79-81)
    for key, refs in ref_binned.items():
        exps = exp_binned.get(key, ())
        for (sid, rstart, rstop, rstrand, rvals) in refs:
            count = 0
            for (estart, estop, estrand) in exps:
                if ((rstart < estop and estart < rstop) and _strand_ok(rstrand, estrand)
                        and ((rstart // bin_size) == key[2] or (estart // bin_size) == key[2])):
                    if (not exact) or (rstart == estart and rstop == estop):
                        count += 1
            agg_key = (md5_long_id(sid), key[1], rstart, rstop, rstrand, rvals)
            reduced[agg_key] = reduced.get(agg_key, 0) + count
    return [(k[0], k[1], k[2], k[3], k[4], k[5]) for k, c in reduced.items() if c == 0]     # :83-88


def difference_direct(ref, exp, groups, exact=False):
    """Bin-free form of DIFFERENCE: the set of distinct reference records (of samples with a non-empty group) that no
    region of a paired experiment sample overlaps."""
    exp_by = defaultdict(list)
    for e in exp:
        exp_by[(e[0], e[1])].append(e)
    out = {}
    for (sid, chrom, start, stop, strand, vals) in ref:
        grp = groups.get(sid)
        if not grp:
            continue
        hit = False
        for exp_id in grp:
            for (_s, _c, estart, estop, estrand, _v) in exp_by.get((exp_id, chrom), ()):
                if start < estop and estart < stop and _strand_ok(strand, estrand) and ((not exact) or (start == estart and stop == estop)):
                    hit = True
        key = (md5_long_id(sid), chrom, start, stop, strand, tuple(vals))
        out[key] = out.get(key, False) or hit
    return [k for k, h in out.items() if not h]


def _wrap64(v: int) -> int:
    """Java long arithmetic wraps."""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def profile_reference(records, sample_ids=None):
    """Profiler.profile's per-sample fold (GMQL-Profiler/.../Profilers/Profiler.scala:124-153): ProfilerValue(leftMost,
    rightMost, minLength, maxLength, sumLength, sumLengthOfSquares, count) per sample, folded from zeroProfilerValue
    (:69) with reduceFunc (:71-82), then calculateResult (:84-92).  Returns {sample_id: dict}; samples listed in
    sample_ids without regions keep count 0."""
    LMAX, LMIN = 2 ** 63 - 1, -(2 ** 63)
    acc = {}
    for (sid, _chrom, start, stop, _strand, _vals) in records:
        d = stop - start
        l, r, mn, mx, sl, sq, c = acc.get(sid, (LMAX, LMIN, LMAX, LMIN, 0, 0, 0))
        acc[sid] = (min(l, start), max(r, stop), min(mn, d), max(mx, d), _wrap64(sl + d), _wrap64(sq + d * d), c + 1)
    out = {}
    for sid in (sample_ids if sample_ids is not None else acc.keys()):
        l, r, mn, mx, sl, sq, c = acc.get(sid, (LMAX, LMIN, LMAX, LMIN, 0, 0, 0))
        if c > 0:
            mean = float(sl) / c
            var = float(sq) / c - mean * mean
            out[sid] = dict(leftMost=l, rightMost=r, minLength=mn, maxLength=mx, count=c, avgLength=mean, varianceLength=var)
        else:
            out[sid] = dict(leftMost=0, rightMost=0, minLength=0, maxLength=0, count=0, avgLength=0.0, varianceLength=0.0)
    return out


class CoverParam:
    """CoverParam ADT (core/DataStructures/CoverParameters/CoverParam.scala:6-16) and the parser's
    ALL / ALL/k / (ALL+m)/k closures (Compiler GmqlParsers.scala:543-562)."""

    def __init__(self, kind, n=0, add=0, div=1):
        self.kind, self.n, self.add, self.div = kind, n, add, div

    @staticmethod
    def N(n):
        return CoverParam("N", n=n)

    @staticmethod
    def ANY():
        return CoverParam("ANY")

    @staticmethod
    def ALL(add=0, div=1):
        return CoverParam("ALL", add=add, div=div)

    def fun(self, x):
        return (x + self.add) // self.div


def cover_thresholds(min_p, max_p, group_ids, group_sizes, n_samples_with_regions):
    """GenometricCover.scala:57-79.  group_sizes: dict group->size (grouped) or None (ungrouped)."""
    if min_p.kind == "ALL" or max_p.kind == "ALL":
        if group_sizes is not None:
            all_value = dict(group_sizes)
        else:
            all_value = {g: n_samples_with_regions for g in group_ids}
    else:
        all_value = {0: 0}
    if min_p.kind == "ALL":
        minimum = {g: min_p.fun(v) for g, v in all_value.items()}
    elif min_p.kind == "ANY":
        minimum = {g: 1 for g in group_ids}
    else:
        minimum = {g: min_p.n for g in group_ids}
    if max_p.kind == "ALL":
        maximum = {g: max_p.fun(v) for g, v in all_value.items()}
    elif max_p.kind == "ANY":
        maximum = {g: INT_MAX for g in group_ids}
    else:
        maximum = {g: max_p.n for g in group_ids}
    return minimum, maximum


def _assign_groups(ds, groups):
    """assignGroups, GenometricCover.scala:327-343."""
    contains_unknown = any(r[4] == "*" for r in ds)
    out = []
    for (sid, chrom, start, stop, strand, vals) in ds:
        g = groups[sid] if groups is not None else 0
        out.append((g, chrom, start, stop, "*" if contains_unknown else strand, sid, vals))
    return out


def _extractor(grouped, B):
    """extractor, GenometricCover.scala:345-360.  Dict literal order reproduces Scala's HashMap(a->x, b->y):
    a duplicate key keeps the LAST binding."""
    out = []
    for (g, chrom, start, stop, strand, _sid, _vals) in grouped:
        start_bin = start // B
        stop_bin = stop // B
        if start_bin == stop_bin:
            m = {}
            m[start - start_bin * B] = 1
            m[stop - start_bin * B] = -1
            out.append(((chrom, start_bin, g, strand), m))
        else:
            so = start - start_bin * B
            m1 = {}
            m1[so if so != B else B - 1] = 1
            m1[B] = -1
            out.append(((chrom, start_bin, g, strand), m1))
            eo = stop - stop_bin * B
            m2 = {}
            m2[eo if eo != 0 else 1] = -1
            m2[0] = 1
            out.append(((chrom, stop_bin, g, strand), m2))
            for i in range(start_bin + 1, stop_bin):
                out.append(((chrom, i, g, strand), {0: 1, B: -1}))
    return out


def _cover_helper(points, minimum, maximum, gid, chrom, strand, b, B):
    """coverHelper, GenometricCover.scala:172-218 (also used for FLAT, :116)."""
    start = 0
    count = 0
    count_max = 0
    recording = False
    out = []
    n = len(points)
    for idx, (off, delta) in enumerate(points):
        has_next = idx + 1 < n
        new_count = count + delta
        in_range = minimum[gid] <= new_count <= maximum[gid]
        new_recording = in_range and has_next
        if (not new_recording) and recording:
            new_count_max = 0
        elif new_recording and new_count > count_max:
            new_count_max = new_count
        else:
            new_count_max = count_max
        if in_range and not recording:
            new_start = 0 if off < 0 else off
        else:
            new_start = start
        if (not new_recording) and recording:
            end = B
            out.append((gid, chrom, new_start + b * B, (end + b * B) if off > end else (off + b * B),
                        strand, float(count_max)))
        count, start, count_max, recording = new_count, new_start, new_count_max, new_recording
    return out


def _histogram_helper(points, minimum, maximum, gid, chrom, strand, b, B):
    """histogramHelper, GenometricCover.scala:229-258.  `start.equals(current._1)` compares a boxed Long
    with an Int and is always false (:235), so every point takes the else branch."""
    count = 0
    start = 0
    out = []
    for (off, delta) in points:
        new_count = count + delta
        if minimum[gid] <= new_count <= maximum[gid] and new_count != count:
            new_start = off
        else:
            new_start = start
        if minimum[gid] <= count <= maximum[gid] and new_count != count and count != 0:
            out.append((gid, chrom, start + b * B, off + b * B, strand, float(count)))
        start = new_start
        count = new_count
    return out


def _summit_helper(points, minimum, maximum, gid, chrom, strand, b, B):
    """summitHelper, GenometricCover.scala:269-316 (same dead `start.equals` branch, :282)."""
    start = 0
    count = 0
    valid = False
    re_entered = False
    growing = False
    out = []
    mn = minimum.get(gid, 1)
    mx = maximum.get(gid, INT_MAX)
    for (off, delta) in points:
        new_count = count + delta
        new_valid = mn <= new_count <= mx
        new_growing = new_count > count or (growing and new_count == count)
        new_re_entered = (not valid) and new_valid
        if new_valid and new_count != count:
            new_start = 0 if off < 0 else off
        else:
            new_start = start
        if ((valid and growing and ((not new_growing) or (not new_valid))) or
                (re_entered and not new_growing)) and count != 0:
            end = B
            out.append((gid, chrom, start + b * B, (end + b * B) if off > end else (off + b * B),
                        strand, float(count)))
        start, count, valid, re_entered, growing = new_start, new_count, new_valid, new_re_entered, new_growing
    return out


def _gmap4(aggs, flat, cover_regions, grouped, B):
    """GMAP4.apply, RO/GenometricMap/GMAP4.scala:20-100.  cover_regions: (gid, chrom, start, stop, strand, acc).
    grouped: assignGroups output.  Returns (gid, chrom, start, stop, strand, values)."""
    # binDS exp (:102-124)
    exp_binned = defaultdict(list)
    for (g, chrom, start, stop, strand, _sid, vals) in grouped:
        new_val = tuple(vals[a.index] for a in aggs)
        if B > 0:
            for i in range(start // B, stop // B + 1):
                exp_binned[(g, chrom, i)].append((start, stop, strand, new_val))
        else:
            exp_binned[(g, chrom, 0)].append((start, stop, strand, new_val))
    # binDS ref (:126-139) + leftOuterJoin (:27-51)
    joined = []
    for (g, chrom, start, stop, strand, acc) in cover_regions:
        bins = range(start // B, stop // B + 1) if B > 0 else (0,)
        for i in bins:
            key = (g, chrom, i)
            # aggregation key: unseparated string concat (:32); values.mkString("/") of [GDouble(acc)]
            agg_key = md5_string_id(str(g) + chrom + str(start) + str(stop) + strand + gdouble_to_string(acc))
            for e in exp_binned.get(key, ()):
                if (start < e[1] and e[0] < stop) and _strand_ok(strand, e[2]) and \
                        ((start // B) == i or (e[0] // B) == i):
                    joined.append((agg_key, ((g, chrom, start, stop, strand), (acc,), e[3],
                                             e[0], e[1], e[0], e[1],
                                             (1, tuple(0 if v is None else 1 for v in e[3])))))
    # reduceByKey (:57-77)
    reduced = {}
    for k, r in joined:
        if k not in reduced:
            reduced[k] = r
            continue
        l = reduced[k]
        if len(l[2]) > 0 and len(r[2]) > 0:
            values = tuple(a.fun([l[2][i], r[2][i]]) for i, a in enumerate(aggs))
        elif len(r[2]) > 0:
            values = r[2]
        else:
            values = l[2]
        start_min = min(l[3], r[3])
        end_max = max(l[4], r[4])
        start_max = max(l[5], r[5])
        end_min = min(l[6], r[6])
        reduced[k] = (l[0], l[1], values, start_min, end_max,
                      start_max if start_max < end_min else 0,
                      end_min if end_min > start_max else 0,
                      (l[7][0] + r[7][0], tuple(x + y for x, y in zip(l[7][1], r[7][1]))))
    # output (:81-95)
    out = []
    for res in reduced.values():
        start = float(res[3]) if flat else float(res[0][2])
        end = float(res[4]) if flat else float(res[0][3])
        new_val = tuple(
            a.fun_out(res[2][i] if len(res[2]) > 0 else 1e-16,
                      (res[7][0], res[7][1][i] if len(res[2]) > 0 else 0))
            for i, a in enumerate(aggs))
        den = res[4] - res[3]
        j1 = abs((end - start) / den) if den != 0 else 0.0
        j2 = abs((float(res[6]) - res[5]) / den) if den != 0 else 0.0
        out.append((res[0][0], res[0][1], int(start), int(end), res[0][4], res[1] + (j1, j2) + new_val))
    return out


def cover_reference(flag, min_p, max_p, aggs, ds, groups=None, bin_size=2000):
    """GenometricCover.apply (:30-158).  flag in COVER/FLAT/SUMMIT/HISTOGRAM; groups: dict sample->group or None.
    Returns list of (group_id, chrom, start, stop, strand, (AccIndex, JaccardIntersect, JaccardResult, aggs...))."""
    B = bin_size
    group_ids = sorted(set(groups.values())) if groups is not None else [0]
    grouped = _assign_groups(ds, groups)
    extracted = _extractor(grouped, B)
    group_sizes = None
    if groups is not None:
        group_sizes = defaultdict(int)
        for _s, g in groups.items():
            group_sizes[g] += 1
    n_with = len(set(r[5] for r in grouped))  # :64
    minimum, maximum = cover_thresholds(min_p, max_p, group_ids, group_sizes, n_with)
    # reduceByKey: sum coincident points (:98-108)
    ss = {}
    for key, m in extracted:
        if key not in ss:
            ss[key] = dict(m)
        else:
            acc = ss[key]
            for k, v in m.items():
                acc[k] = acc.get(k, 0) + v
    helper = {"COVER": _cover_helper, "FLAT": _cover_helper,
              "HISTOGRAM": _histogram_helper, "SUMMIT": _summit_helper}[flag]
    binned = []
    for (chrom, b, gid, strand), m in ss.items():
        points = sorted(m.items(), key=lambda kv: kv[0])  # :111
        binned.extend(helper(points, minimum, maximum, gid, chrom, strand, b, B))
    # split + boundary merge (:121-152)
    valid = [v for v in binned if v[2] % B != 0 and v[3] % B != 0]
    not_valid = [v for v in binned if v[2] % B == 0 or v[3] % B == 0]
    by = defaultdict(list)
    for v in not_valid:
        by[(v[0], v[1], v[4])].append(v)
    joined = []
    for _k, lst in by.items():
        lst = sorted(lst, key=lambda x: (x[2], x[3]))
        old = lst[0]
        for cur in lst[1:]:
            if old[3] == cur[2] and (flag != "HISTOGRAM" or old[5] == cur[5]):
                old = (old[0], old[1], old[2], cur[3], old[4], max(old[5], cur[5]))
            else:
                joined.append(old)
                old = cur
        joined.append(old)
    pure = valid + joined
    return _gmap4(aggs, flag == "FLAT", pure, grouped, B)


def cover_phase1_direct(flag, minimum, maximum, grouped, B=2000):
    """Bin-free form of phase 1 for COVER/FLAT/HISTOGRAM: a global sweep over
    (start:+1, stop':-1) with stop' = stop+1 when stop % B == 0 and the region crosses a bin edge.
    SUMMIT has no bin-free form (A-C7).  Zero-length regions follow the HashMap quirk (A-C4):
    a single -1 at the point, whose effect ends at the bin edge."""
    ev = defaultdict(lambda: defaultdict(int))
    for (g, chrom, start, stop, strand, _sid, _v) in grouped:
        seg = ev[(g, chrom, strand)]
        if start == stop:
            seg[start] -= 1
            seg[(start // B + 1) * B] += 1
            continue
        stop2 = stop + 1 if (stop % B == 0 and start // B != stop // B) else stop
        seg[start] += 1
        seg[stop2] -= 1
    out = []
    for (g, chrom, strand), m in ev.items():
        pts = sorted(m.items())
        mn, mx = minimum[g], maximum[g]
        depth = 0
        if flag in ("COVER", "FLAT"):
            run_start = None
            run_max = 0
            for (x, d) in pts:
                depth += d
                inr = mn <= depth <= mx
                if inr and run_start is None:
                    run_start, run_max = x, depth
                elif inr:
                    run_max = max(run_max, depth)
                elif run_start is not None:
                    out.append((g, chrom, run_start, x, strand, float(run_max)))
                    run_start = None
        else:  # HISTOGRAM
            prev_x = None
            for (x, d) in pts:
                if d == 0:
                    continue
                if prev_x is not None and mn <= depth <= mx and depth != 0:
                    out.append((g, chrom, prev_x, x, strand, float(depth)))
                depth += d
                prev_x = x
    return out


def gmap4_direct(aggs, flat, cover_regions, grouped):
    """Bin-free form of GMAP4, keyed on the region itself (no string aliasing, A-C10)."""
    by = defaultdict(list)
    for e in grouped:
        by[(e[0], e[1])].append(e)
    out = []
    for (g, chrom, start, stop, strand, acc) in cover_regions:
        S = [e for e in by.get((g, chrom), ())
             if start < e[3] and e[2] < stop and _strand_ok(strand, e[4])]
        if not S:
            continue
        s_min = min(e[2] for e in S)
        e_max = max(e[3] for e in S)
        s_max = max(e[2] for e in S)
        e_min = min(e[3] for e in S)
        if len(S) > 1 and not (s_max < e_min):
            s_max, e_min = 0, 0
        o_start, o_end = (s_min, e_max) if flat else (start, stop)
        den = e_max - s_min
        j1 = abs((float(o_end) - float(o_start)) / den) if den != 0 else 0.0
        j2 = abs((float(e_min) - s_max) / den) if den != 0 else 0.0
        new_val = []
        for a in aggs:
            col = [e[6][a.index] for e in S]
            v = col[0]
            for x in col[1:]:
                v = a.fun([v, x])
            new_val.append(a.fun_out(v, (len(S), sum(1 for c in col if c is not None))))
        out.append((g, chrom, o_start, o_end, strand, (acc, j1, j2) + tuple(new_val)))
    return out


# --------------------------------------------------------------------------------------
# JOIN — literal (RO/GenometricJoin.scala)
# --------------------------------------------------------------------------------------

def distance_calculator(a, b):
    """distanceCalculator, GenometricJoin.scala:375-386."""
    d1 = a[0] - b[1]
    d2 = b[0] - a[1]
    if a[1] < b[0] or b[1] < a[0]:
        return min(abs(d1), abs(d2))
    return -min(abs(d1), abs(d2))


class _JP:
    def __init__(self, mx=None, mn=None, stream=None):
        self.max, self.min, self.stream = mx, mn, stream


def _create_execution_parameters(atoms):
    """createExecutionParameters, GenometricJoin.scala:388-404.  atoms: list of (kind, arg)."""
    p = _JP()
    for kind, arg in atoms:
        if kind == "DISTLESS":
            p = _JP(arg, p.min, p.stream)
        elif kind == "DISTGREATER":
            p = _JP(p.max, arg, p.stream)
        elif kind == "UP":
            p = _JP(p.max, p.min, "+")
        elif kind == "DOWN":
            p = _JP(p.max, p.min, "-")
        else:
            raise ValueError(kind)
    return p


def _stream_ok(stream, lstrand, lstart, lstop, rstart, rstop):
    """The stream predicate shared by checkRegionCondition (:255-279) and the second round (:171-183)."""
    if stream == "+":
        return ((lstrand in "+*") and rstop <= lstart) or (lstrand == "-" and rstart >= lstop)
    if stream == "-":
        return ((lstrand in "+*") and rstart >= lstop) or (lstrand == "-" and rstop <= lstart)
    return True


def _bin_start_ref(l, first, max_d, B):
    """computeBinStartRef, GenometricJoin.scala:284-296."""
    strand = l[4]
    if first.stream is None or first.stream == strand or (strand == "*" and first.stream == "+"):
        v = l[2] - max_d
    else:
        v = l[3]
    return max(0, v) // B


def _bin_stop_ref(l, first, max_d, B):
    """computeBinStopRef, GenometricJoin.scala:298-309."""
    strand = l[4]
    if first.stream is None or first.stream != strand or (strand == "*" and first.stream == "-"):
        v = l[3] + max_d
    else:
        v = l[2]
    return v // B  # Scala long division truncates toward zero; operands are >= 0 here


def join_split_atoms(atoms, max_bin_distance=100000):
    """GenometricJoin.scala:62-79: first/second round and maxDistance."""
    i = 0
    while i < len(atoms) and atoms[i][0] != "MD":
        i += 1
    first = _create_execution_parameters(atoms[:i])
    md = atoms[i][1] if i < len(atoms) else None
    second = _create_execution_parameters(atoms[i + 1:])
    if first.max is not None:
        max_d = max(0, first.max)
    elif second.max is not None:
        max_d = max(second.max, max_bin_distance)
    elif first.min is not None:
        max_d = first.min + max_bin_distance
    else:
        max_d = max_bin_distance
    return first, md, second, max_d


def join_regions(left, right, builder):
    """joinRegions, GenometricJoin.scala:345-372.  left/right: (id, chrom, start, stop, strand, values)."""
    lv, rv = tuple(left[5]), tuple(right[5])
    if builder == "LEFT":
        return (left[0], left[1], left[2], left[3], left[4], lv + rv)
    if builder == "LEFT_DISTINCT":
        return (left[0], left[1], left[2], left[3], left[4], lv)
    if builder == "RIGHT":
        return (left[0], right[1], right[2], right[3], right[4], lv + rv)
    if builder == "RIGHT_DISTINCT":
        return (left[0], right[1], right[2], right[3], right[4], rv)
    if builder == "CONTIG":
        return (left[0], left[1], min(left[2], right[2]), max(left[3], right[3]), left[4], lv + rv)
    if builder == "BOTH":
        return (left[0], left[1], left[2], left[3], left[4],
                lv + (right[1], float(right[2]), float(right[3]), right[4]) + rv)
    if builder == "INTERSECTION":
        return (left[0], left[1], max(left[2], right[2]), min(left[3], right[3]),
                left[4] if left[4] == right[4] else "*", lv + rv)
    raise ValueError(builder)


def join_reference(left, right, pairs, atoms, builder, bin_size=5000, max_bin_distance=100000,
                   join_on=None):
    """GenometricJoin.apply (:21-212) for one JoinQuadruple.  pairs: dict (left_id, right_id) -> new id.
    atoms: list of (kind, arg), kind in DISTLESS/DISTGREATER/MD/UP/DOWN.  join_on: optional list of
    (left_value_index, right_value_index) equi-conditions on value columns (:102-114)."""
    B = bin_size
    first, md, second, max_d = join_split_atoms(atoms, max_bin_distance)
    # binLeftDs2 (:313-331) / binRightDs2 (:334-342) / join on (bin, chrom) (:97-99)
    right_binned = defaultdict(list)
    for r in right:
        for b in range(r[2] // B, r[3] // B + 1):
            right_binned[(b, r[1])].append(r)
    first_pairs = []
    for l in left:
        b0 = _bin_start_ref(l, first, max_d, B)
        b1 = _bin_stop_ref(l, first, max_d, B)
        for b in range(b0, b1 + 1):
            for r in right_binned.get((b, l[1]), ()):
                if join_on and not all(l[5][i] == r[5][j] for (i, j) in join_on):
                    continue
                # first filter (:117-127)
                if not (b == b0 or b == r[2] // B):
                    continue
                if not _strand_ok(l[4], r[4]):
                    continue
                if (l[0], r[0]) not in pairs:
                    continue
                d = distance_calculator((l[2], l[3]), (r[2], r[3]))
                if not ((first.max is None or first.max > d) and (first.min is None or first.min < d)):
                    continue
                if not _stream_ok(first.stream, l[4], l[2], l[3], r[2], r[3]):
                    continue
                first_pairs.append((l, r))
    # MD(k) (:130-157)
    if md is not None:
        groups = defaultdict(list)
        for l, r in first_pairs:
            new_id = pairs[(l[0], r[0])]
            groups[(new_id, l[1], l[2], l[3], l[4], tuple(l[5]))].append(r)
        first_round = []
        for mk, lst in groups.items():
            itr = sorted(((e, distance_calculator((mk[2], mk[3]), (e[2], e[3]))) for e in lst),
                         key=lambda t: t[1])
            count = min(len(itr), md)
            thr = itr[count - 1][1]
            for e, d in itr:
                if d <= thr:
                    first_round.append((mk, e))
    else:
        first_round = [((pairs[(l[0], r[0])], l[1], l[2], l[3], l[4], tuple(l[5])), r) for l, r in first_pairs]
    # second round (:160-188)
    if second.max is not None or second.min is not None or second.stream is not None:
        res = []
        for l, r in first_round:
            d = distance_calculator((l[2], l[3]), (r[2], r[3]))
            if (second.max is None or second.max > d) and (second.min is None or second.min < d) and \
                    _stream_ok(second.stream, l[4], l[2], l[3], r[2], r[3]):
                res.append((l, r))
    else:
        res = first_round
    if builder == "INTERSECTION":  # :190-198
        res = [(l, r) for l, r in res if l[2] < r[3] and r[2] < l[3]]
    out = [join_regions(l, r, builder) for l, r in res]
    if builder in ("RIGHT_DISTINCT", "LEFT_DISTINCT"):  # :205-218
        seen = {}
        for o in out:
            seen.setdefault(o, o)
        out = list(seen.values())
    return out


def join_direct_pairs(left, right, pairs, atoms, bin_size=5000, max_bin_distance=100000):
    """Bin-free statement of the candidate predicate returning surviving (l, r) index pairs,
    MD(k) per (left record, right sample) with ties, then the second round.  Left duplicates are kept as
    separate rows unless identical records exist (then the MapKey grouping pools them, A-J4)."""
    B = bin_size
    first, md, second, max_d = join_split_atoms(atoms, max_bin_distance)
    cands = []
    for li, l in enumerate(left):
        b0 = _bin_start_ref(l, first, max_d, B)
        b1 = _bin_stop_ref(l, first, max_d, B)
        if b0 > b1:
            continue
        for ri, r in enumerate(right):
            if r[1] != l[1] or (l[0], r[0]) not in pairs:
                continue
            if not (r[2] // B <= b1 and r[3] // B >= b0):
                continue
            if not _strand_ok(l[4], r[4]):
                continue
            d = distance_calculator((l[2], l[3]), (r[2], r[3]))
            if not ((first.max is None or first.max > d) and (first.min is None or first.min < d)):
                continue
            if not _stream_ok(first.stream, l[4], l[2], l[3], r[2], r[3]):
                continue
            cands.append((li, ri, d))
    if md is not None:
        groups = defaultdict(list)
        for li, ri, d in cands:
            l = left[li]
            groups[(l[0], right[ri][0], l[1], l[2], l[3], l[4], tuple(l[5]))].append((li, ri, d))
        kept = []
        for _k, lst in groups.items():
            ds = sorted(t[2] for t in lst)
            thr = ds[min(len(ds), md) - 1]
            kept.extend(t for t in lst if t[2] <= thr)
        cands = kept
    out = []
    for li, ri, d in cands:
        l, r = left[li], right[ri]
        if (second.max is None or second.max > d) and (second.min is None or second.min < d) and \
                _stream_ok(second.stream, l[4], l[2], l[3], r[2], r[3]):
            out.append((li, ri))
    return out


# --------------------------------------------------------------------------------------
# text I/O restatement (host side of the boundary; loaders/BedParser.scala)
# --------------------------------------------------------------------------------------

PARSERS = {
    # name: (chrPos, startPos, stopPos, strandPos, [(pos, type)])  — BedParser.scala:295, :346, :525
    "NarrowPeakParser": (0, 1, 2, 5, [(3, "STRING"), (4, "DOUBLE"), (6, "DOUBLE"), (7, "DOUBLE"), (8, "DOUBLE"), (9, "DOUBLE")]),
    "RnaSeqParser": (0, 1, 2, 3, [(4, "STRING"), (5, "DOUBLE")]),
    "BedScoreParser": (0, 1, 2, None, [(3, "DOUBLE")]),
}


def parse_region_line(parser: str, sample_id: int, line: str):
    """BedParser.region_parser (BedParser.scala:112-172) for the 0-based tab formats used here."""
    chr_pos, start_pos, stop_pos, strand_pos, other = PARSERS[parser]
    if line.startswith("#"):
        return None
    s = line.split("\t")
    vals = []
    for pos, typ in other:
        x = s[pos]
        if typ == "STRING":
            vals.append(x.strip())
        else:
            vals.append(None if x.lower() in ("null", ".") else float(x))  # parseRegion :30-48
    strand = "*"
    if strand_pos is not None:
        t = s[strand_pos].strip()
        strand = t if t in ("+", "-") else "*"  # :166-170
    return (sample_id, s[chr_pos].strip(), int(s[start_pos].strip()), int(s[stop_pos].strip()), strand, tuple(vals))
