"""COVER / FLAT / SUMMIT / HISTOGRAM parity: CUDA path (through the C ABI) vs the oracle's literal restatement of
GenometricCover + GMAP4 (per-2000bp-bin sweeps, boundary merge, second-phase join).  Coordinates, AccIndex, the two
Jaccard doubles and MIN/MAX/COUNT are exact; SUM/AVG within rel 1e-9 (test values are multiples of 0.25: exact)."""
import random

import pytest

from oracle import gmql_oracle as O
from tests.randgen import rand_regions, canon

pytestmark = pytest.mark.gpu

FLAGS = ["COVER", "FLAT", "SUMMIT", "HISTOGRAM"]


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _params():
    import gmql_b200 as G
    return [
        ((G.N(1), G.N(2)), (O.CoverParam.N(1), O.CoverParam.N(2))),
        ((G.N(2), G.N(2)), (O.CoverParam.N(2), O.CoverParam.N(2))),
        ((G.N(1), G.ANY()), (O.CoverParam.N(1), O.CoverParam.ANY())),
        ((G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY())),
        ((G.N(1), G.ALL()), (O.CoverParam.N(1), O.CoverParam.ALL())),
        ((G.ANY(), G.ANY()), (O.CoverParam.ANY(), O.CoverParam.ANY())),
        ((G.ANY(), G.ALL()), (O.CoverParam.ANY(), O.CoverParam.ALL())),
        ((G.ALL(), G.ANY()), (O.CoverParam.ALL(), O.CoverParam.ANY())),
        ((G.ALL(div=2), G.ANY()), (O.CoverParam.ALL(div=2), O.CoverParam.ANY())),
        ((G.ALL(add=1, div=2), G.ALL()), (O.CoverParam.ALL(add=1, div=2), O.CoverParam.ALL())),
    ]


def _check(ex, flag, ds, sample_ids, groups, gp, op, aggs):
    from gmql_b200 import RegionDataset, Agg
    d = RegionDataset.from_records(ds, sample_ids=sample_ids)
    got = ex.GenometricCover(flag, gp[0], gp[1], [Agg(k, i) for (k, i) in aggs], groups, d)
    want = O.cover_reference(flag, op[0], op[1], [O.Agg(k, i) for (k, i) in aggs], ds, groups, 2000)
    assert canon(got, 9) == canon(want, 9), (flag, gp)


@pytest.mark.parametrize("flag", FLAGS)
@pytest.mark.parametrize("seed", range(4))
def test_cover_random(ex, flag, seed):
    rng = random.Random(300 + seed)
    ids = [1, 2, 3, 4, 5]
    strands = "+-*" if seed % 2 == 0 else "+-"
    ds = rand_regions(rng, 260, ids, strands=strands, n_vals=1, min_len=(0 if seed == 3 else 1))
    groups = None if seed < 2 else {1: 50, 2: 50, 3: 90, 4: 90, 5: 90}
    for gp, op in _params():
        _check(ex, flag, ds, ids, groups, gp, op, [("SUM", 0), ("MIN", 0), ("COUNT", 0), ("AVG", 0), ("MAX", 0)])


@pytest.mark.parametrize("flag", FLAGS)
def test_cover_conf_shape(ex, flag):
    """conf/test_cover.xml / test_Summit_Flat_Histo.xml shape: 10 files x 1 chr x 20 regions, chromlen 10000, len 5-200."""
    rng = random.Random(11)
    ids = list(range(1, 11))
    ds = rand_regions(rng, 200, ids, chroms=("chr1",), strands="*", span=10000, max_len=200, min_len=5, n_vals=1, aligned_p=0.1, null_p=0.0)
    groups = {i: (100 if i <= 5 else 200) for i in ids}
    for gp, op in _params():
        _check(ex, flag, ds, ids, None, gp, op, [("SUM", 0)])
        _check(ex, flag, ds, ids, groups, gp, op, [("SUM", 0), ("MIN", 0)])


@pytest.mark.parametrize("flag", FLAGS)
def test_cover_dense_long_regions(ex, flag):
    """regions spanning many 2000-bp bins, deep pile-ups, many bin-aligned coordinates"""
    rng = random.Random(77)
    ids = list(range(8))
    ds = rand_regions(rng, 600, ids, chroms=("chr1", "chr11", "chrX"), strands="+-", span=60000, max_len=15000, n_vals=1, aligned_p=0.5)
    import gmql_b200 as G
    _check(ex, flag, ds, ids, None, (G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY()), [("SUM", 0)])
    _check(ex, flag, ds, ids, None, (G.N(3), G.N(6)), (O.CoverParam.N(3), O.CoverParam.N(6)), [])


def test_cover_stop_on_bin_edge(ex):
    import gmql_b200 as G
    ds = [(1, "chr1", 3000, 4000, "*", (1.0,)), (2, "chr1", 3500, 6000, "*", (2.0,)), (2, "chr1", 4000, 4500, "*", (None,))]
    for flag in FLAGS:
        _check(ex, flag, ds, [1, 2], None, (G.N(1), G.ANY()), (O.CoverParam.N(1), O.CoverParam.ANY()), [("SUM", 0)])
        _check(ex, flag, ds, [1, 2], None, (G.N(1), G.N(1)), (O.CoverParam.N(1), O.CoverParam.N(1)), [("SUM", 0)])


def test_cover_empty_and_single(ex):
    import gmql_b200 as G
    from gmql_b200 import RegionDataset
    d = RegionDataset.from_records([(1, "chr1", 5, 9, "+", ())], sample_ids=[1])
    got = ex.GenometricCover("COVER", G.N(1), G.ANY(), [], None, d)
    assert got == [(0, "chr1", 5, 9, "+", (1.0, 1.0, 1.0))]
    got = ex.GenometricCover("COVER", G.N(2), G.ANY(), [], None, d)
    assert got == []


def _canon_bag(rows, bag_cols):
    """BAG / BAGD strings are folded pairwise (GMAP4.scala:57-66): their layout depends on the fold order, which in the
    reference is Spark's shuffle order.  Compare them as token multisets (BAG) / sets (BAGD); everything else exactly."""
    out = []
    for (g, chrom, start, stop, strand, vals) in rows:
        v = list(vals)
        for k, distinct in bag_cols:
            if isinstance(v[k], str):
                toks = v[k].split(",")
                v[k] = ("bag", tuple(sorted(set(toks)) if distinct else sorted(toks)))
            elif isinstance(v[k], float):
                v[k] = ("bag", (repr(v[k]),))
        out.append((g, chrom, start, stop, strand, tuple(round(x, 9) if isinstance(x, float) else x for x in v)))
    return sorted(out, key=repr)


@pytest.mark.parametrize("flag", ["COVER", "FLAT", "SUMMIT", "HISTOGRAM"])
@pytest.mark.parametrize("grouped", [False, True])
def test_cover_bag_aggregates(ex, flag, grouped):
    """BAG / BAGD in COVER (DefaultRegionsToRegionFactory.scala:128-168 through GMAP4.scala:57-66,85): the device lists the
    contributors of every output region (one internal MAP of the input, samples relabelled with their groups, onto the COVER
    rows) and the host folds the strings; next to numeric aggregates, with strands, groups and a string column."""
    import random
    import gmql_b200 as G
    from gmql_b200 import RegionDataset, Agg
    from oracle import gmql_oracle as O
    from tests.randgen import rand_regions
    rng = random.Random(77 + len(flag))
    ids = [1, 2, 3, 4]
    ds = rand_regions(rng, 260, ids, strands="+-" if grouped else "+-*", n_vals=2, span=60000, max_len=1500, with_name=True)
    groups = {1: 10, 2: 10, 3: 20, 4: 20} if grouped else None
    aggs = [("BAG", 1), ("SUM", 1), ("BAGD", 0), ("COUNT", 0), ("BAGD", 2)]     # value 0 = the name (string), 1 and 2 numeric with nulls
    d = RegionDataset.from_records(ds, sample_ids=ids)
    got = ex.GenometricCover(flag, G.N(2), G.ANY(), [Agg(k, i) for k, i in aggs], groups, d)
    want = O.cover_reference(flag, O.CoverParam.N(2), O.CoverParam.ANY(), [O.Agg(k, i) for k, i in aggs], ds, groups, 2000)
    assert len(want) > 20
    bag_cols = [(3 + 0, False), (3 + 2, True), (3 + 4, True)]
    assert _canon_bag(got, bag_cols) == _canon_bag(want, bag_cols)
    # with at most two contributors the fold order cannot matter: those rows agree as plain strings
    plain = lambda rows: sorted((r[:5], r[5][3], r[5][5], r[5][7]) for r in rows if r[5][6] <= 2.0)
    assert plain(got) == plain(want) and len(plain(got)) > 0


def _low_params():
    import gmql_b200 as G
    return [
        ((G.N(0), G.ANY()), (O.CoverParam.N(0), O.CoverParam.ANY())),
        ((G.N(0), G.N(2)), (O.CoverParam.N(0), O.CoverParam.N(2))),
        ((G.N(0), G.N(1)), (O.CoverParam.N(0), O.CoverParam.N(1))),
        ((G.ALL(add=-5), G.ANY()), (O.CoverParam.ALL(add=-5), O.CoverParam.ANY())),     # (ALL - 5) <= 0 with five samples
        ((G.ALL(add=-5), G.ALL()), (O.CoverParam.ALL(add=-5), O.CoverParam.ALL())),
    ]


@pytest.mark.parametrize("flag", FLAGS)
@pytest.mark.parametrize("seed", range(4))
def test_cover_min_acc_below_one(ex, flag, seed):
    """minAcc < 1 (GenometricCover.scala:57-79): depth 0 is then "in range", and coverHelper (:172-218) records from the first to
    the last event of every 2000-bp bin that holds events at all — the result depends on the bins, so it comes from the bin-wise
    path (the reference's per-bin state machines, literally), with the boundary merge and GMAP4's drop of empty regions after it."""
    rng = random.Random(900 + seed)
    ids = [1, 2, 3, 4, 5]
    strands = "+-*" if seed % 2 == 0 else "+-"
    ds = rand_regions(rng, 220, ids, strands=strands, n_vals=1, span=30000, max_len=3500, min_len=(0 if seed == 3 else 1))
    groups = None if seed < 2 else {1: 50, 2: 50, 3: 90, 4: 90, 5: 90}
    for gp, op in _low_params():
        if groups is not None and "ALL" in repr(gp[0]):
            continue                                        # (ALL - 5) is negative for groups of two or three samples: covered ungrouped
        _check(ex, flag, ds, ids, groups, gp, op, [("SUM", 0), ("COUNT", 0), ("MAX", 0)])


@pytest.mark.parametrize("flag", ["COVER", "FLAT", "HISTOGRAM"])
def test_cover_binwise_path_equals_oracle_for_ordinary_thresholds(ex, flag):
    """The bin-wise path (option cover.path = bins) is the reference's algorithm bin by bin; with minAcc >= 1 it must give what the
    sweep / tile paths give: a cross-check of both against the literal oracle."""
    rng = random.Random(4242)
    ids = list(range(6))
    ds = rand_regions(rng, 500, ids, chroms=("chr1", "chr7"), strands="+-*", span=50000, max_len=9000, n_vals=1, aligned_p=0.4)
    ex.set_option("cover.path", "bins")
    try:
        for gp, op in _params()[:5]:
            _check(ex, flag, ds, ids, None, gp, op, [("SUM", 0), ("MIN", 0)])
    finally:
        ex.set_option("cover.path", None)


@pytest.mark.parametrize("flag", ["SUMMIT", "HISTOGRAM", "COVER"])
@pytest.mark.parametrize("form", ["merge", "merge-split", "scatter"])
def test_gmap4_merge_and_scatter_forms(ex, flag, form):
    """GMAP4 over many small phase-1 regions without value aggregates: the merge form (sorted phase-1 regions against the sorted
    input index, staged per chunk, no atomics) and the RED scatter form against the oracle — short regions (slices that fit the
    stage) and a deep pile-up of long regions (a slice beyond the stage, read from global memory)."""
    import gmql_b200 as G
    rng = random.Random(31337)
    ids = list(range(6))
    short = rand_regions(rng, 1500, ids, chroms=("chr1", "chr2"), strands="+-", span=40000, max_len=400, n_vals=1, aligned_p=0.2)   # (COUNT reads column 0 in the reference)
    pile = rand_regions(rng, 3600, ids, chroms=("chr3",), strands="*", span=30000, max_len=25000, min_len=15000, n_vals=0, aligned_p=0.1)
    ex.set_option("cover.path", "sort")
    ex.set_option("cover.gmap4", "scatter" if form == "scatter" else None)
    ex.set_option("cover.finish", "split" if form == "merge-split" else None)   # default: straight into the result columns
    try:
        _check(ex, flag, short, ids, None, (G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY()), [("COUNT", 0)])
        _check(ex, flag, short, ids, {i: (i % 2) for i in ids}, (G.N(1), G.N(3)), (O.CoverParam.N(1), O.CoverParam.N(3)), [])
        if flag != "COVER":
            # the pile-up against the C oracle (the literal Python GMAP4 takes minutes on 3 600 mutually overlapping regions)
            from gmql_b200 import RegionDataset
            from oracle import c_oracle as CO
            d = RegionDataset.from_records(pile, sample_ids=ids)
            got = ex.GenometricCover(flag, G.N(2), G.ANY(), [], None, d)
            w = CO.cover_rows(d, {"SUMMIT": 2, "HISTOGRAM": 3}[flag], [2], [2 ** 31 - 1], None, 2000)
            want = [(0, d.chrom_names[w["chrom"][k]], int(w["start"][k]), int(w["stop"][k]), "*+-"[w["strand"][k]],
                     (float(w["acc"][k]), float(w["j1"][k]), float(w["j2"][k]))) for k in range(len(w["start"]))]
            assert len(want) > 100 and canon(got, 12) == canon(want, 12)
    finally:
        ex.set_option("cover.path", None)
        ex.set_option("cover.gmap4", None)
        ex.set_option("cover.finish", None)


@pytest.mark.parametrize("flag", FLAGS)
def test_binwise_machines_in_place(ex, flag):
    """The per-bin state machines without their shared-memory stage (option cover.bins = inplace: the route of ranges beyond the
    stage): SUMMIT, and COVER / FLAT / HISTOGRAM with minAcc < 1, against the literal oracle."""
    import gmql_b200 as G
    rng = random.Random(606)
    ids = [1, 2, 3, 4]
    ds = rand_regions(rng, 400, ids, chroms=("chr1", "chr4"), strands="+-*", span=30000, max_len=5000, n_vals=1, aligned_p=0.3)
    ex.set_option("cover.bins", "inplace")
    try:
        _check(ex, flag, ds, ids, None, (G.N(0), G.N(3)), (O.CoverParam.N(0), O.CoverParam.N(3)), [("SUM", 0)])
        if flag == "SUMMIT":
            _check(ex, flag, ds, ids, None, (G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY()), [])
    finally:
        ex.set_option("cover.bins", None)
