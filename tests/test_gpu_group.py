"""GROUP by coordinates (GroupRD.scala:22-64) and MERGE (MergeRD.scala:20-45): CUDA path through the C ABI vs the oracle's literal
restatement.  Keys, counts, MIN / MAX / MEDIAN and the BAG strings are exact (GROUP applies every aggregate once to a group's
whole list, so BAG does not depend on a fold order); SUM / AVG within rel 1e-9 (test values are multiples of 0.25: exact)."""
import random

import pytest

from oracle import gmql_oracle as O
from tests.randgen import rand_regions, canon

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _dupes(rng, base, n_extra, n_vals, with_name):
    """copies of existing records with fresh values (same key), some with a changed strand or stop (different key, same start)"""
    out = []
    for _ in range(n_extra):
        sid, chrom, start, stop, strand, vals = rng.choice(base)
        v = []
        if with_name:
            v.append(rng.choice(["geneA", "geneB", "geneC"]))
        for _k in range(n_vals):
            v.append(None if rng.random() < 0.15 else float(rng.randrange(-50, 50)) / 4.0)
        how = rng.random()
        if how < 0.6:
            out.append((sid, chrom, start, stop, strand, tuple(v)))
        elif how < 0.8:
            out.append((sid, chrom, start, stop + rng.randrange(1, 5), strand, tuple(v)))
        else:
            out.append((sid, chrom, start, stop, rng.choice("+-*"), tuple(v)))
    return out


def _no_alias(ds, fields):
    """the reference's grouping string is unseparated (GroupRD.scala:34-38): make sure this input does not alias under it"""
    strs = {("%d%s%d%d%s" % r[:5]) + "".join("§" + O.gvalue_to_string(r[5][p]) for p in fields) for r in ds}
    tups = {r[:5] + tuple(O.gvalue_to_string(r[5][p]) for p in fields) for r in ds}
    return len(strs) == len(tups)


@pytest.mark.parametrize("seed", range(4))
def test_group_random(ex, seed):
    from gmql_b200 import RegionDataset, Agg
    rng = random.Random(7000 + seed)
    ids = [11, 12, 13]
    base = rand_regions(rng, 150, ids, n_vals=2, span=3000, max_len=400, with_name=True)
    ds = base + _dupes(rng, base, 260, 2, True)
    rng.shuffle(ds)
    aggs = [("COUNT", 0), ("SUM", 1), ("AVG", 1), ("MIN", 2), ("MAX", 2), ("MEDIAN", 1), ("BAG", 0), ("BAGD", 1), ("MEDIAN", 2)]
    for fields in (None, [0], [1], [0, 2]):
        assert _no_alias(ds, fields or [])
        d = RegionDataset.from_records(ds, sample_ids=ids)
        got = ex.GroupRD(fields, [Agg(k, i) for k, i in aggs], d)
        want = O.group_reference(ds, fields, [O.Agg(k, i) for k, i in aggs])
        assert len(want) <= len(ds) and (fields or len(want) < len(ds)) and canon(got, 9) == canon(want, 9), fields


def test_group_without_aggregates_and_medians(ex):
    from gmql_b200 import RegionDataset, Agg
    ds = [(1, "chr1", 10, 20, "+", (4.0,)), (1, "chr1", 10, 20, "+", (1.0,)), (1, "chr1", 10, 20, "+", (None,)), (1, "chr1", 10, 20, "+", (3.0,)),
          (1, "chr1", 10, 20, "+", (2.0,)), (1, "chr1", 10, 20, "-", (9.0,)), (2, "chr1", 10, 20, "+", (7.0,)), (1, "chr1", 10, 21, "+", (None,)),
          (1, "chr2", 10, 20, "+", (-0.0,)), (1, "chr2", 10, 20, "+", (0.0,))]
    d = RegionDataset.from_records(ds, sample_ids=[1, 2])
    got = ex.GroupRD(None, None, d)
    assert canon(got) == canon(O.group_reference(ds, None, None)) and len(got) == 5
    aggs = [("MEDIAN", 0), ("COUNT", 0), ("MIN", 0), ("BAGD", 0)]
    got = ex.GroupRD(None, [Agg(k, i) for k, i in aggs], d)
    want = O.group_reference(ds, None, [O.Agg(k, i) for k, i in aggs])
    assert canon(got) == canon(want)
    by = {r[:5]: r[5] for r in got}
    assert by[(1, "chr1", 10, 20, "+")][:2] == (2.0, 5.0)          # four non-null values 1 2 3 4: the lower median; five members
    assert by[(1, "chr1", 10, 21, "+")][0] is None                 # no non-null value: GNull
    import math
    assert math.copysign(1.0, by[(1, "chr2", 10, 20, "+")][0]) < 0   # -0.0 sorts below 0.0 (Double.compare): median of (-0.0, 0.0)


def test_merge(ex):
    from gmql_b200 import RegionDataset
    rng = random.Random(5)
    ids = [1, 2, 3, 4]
    ds = rand_regions(rng, 60, ids, n_vals=1)
    d = RegionDataset.from_records(ds, sample_ids=ids)
    assert canon(ex.MergeRD(None, d)) == canon(O.merge_reference(ds, None))
    groups = {1: 100, 2: 100, 3: 200}                              # sample 4 has no group: dropped by the join
    got = ex.MergeRD(groups, d)
    assert canon(got) == canon(O.merge_reference(ds, groups)) and {r[0] for r in got} == {100, 200}


def test_group_edges(ex):
    """Empty input; coordinates beyond 2^31 (int64 columns); a long run of records that share sample, chromosome and start but
    differ in stop / strand (every row looks back over its whole run: O(run) per row)."""
    from gmql_b200 import RegionDataset, Agg
    d = RegionDataset.from_records([], sample_ids=[1], n_values=1)
    assert ex.GroupRD(None, [Agg("COUNT", 0)], d) == []
    off = 5_000_000_000
    rng = random.Random(99)
    base = rand_regions(rng, 80, [1, 2], n_vals=1, span=2000, max_len=300)
    ds = [(r[0], r[1], r[2] + off, r[3] + off, r[4], r[5]) for r in base + base[:30] + base[10:25]]
    aggs = [("COUNT", 0), ("SUM", 0), ("MEDIAN", 0), ("MAX", 0)]
    got = ex.GroupRD(None, [Agg(k, i) for k, i in aggs], RegionDataset.from_records(ds, sample_ids=[1, 2]))
    assert canon(got, 9) == canon(O.group_reference(ds, None, [O.Agg(k, i) for k, i in aggs]), 9)
    run = [(1, "chr1", 1000, 1001 + (k % 700), "+-*"[k % 3], (float(k % 11),)) for k in range(3000)]
    rng.shuffle(run)
    got = ex.GroupRD(None, [Agg(k, i) for k, i in aggs], RegionDataset.from_records(run, sample_ids=[1]))
    want = O.group_reference(run, None, [O.Agg(k, i) for k, i in aggs])
    assert len(want) == 2100 and canon(got, 9) == canon(want, 9)
