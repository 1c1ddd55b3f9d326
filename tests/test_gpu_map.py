"""MAP parity: CUDA path (through the C ABI) vs the oracle's literal restatement of GenometricMap71.
Counts and coordinates are exact; SUM/AVG within rel 1e-9 (reduction order differs)."""
import random

import pytest

from oracle import gmql_oracle as O
from tests.randgen import rand_regions, canon

pytestmark = pytest.mark.gpu

RTOL_DIGITS = 7  # canon() rounds floats to this many decimals; test values are multiples of 0.25 so sums are exact


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _run(ex, ref, exp, pairs, aggs, ref_ids, exp_ids):
    from gmql_b200 import RegionDataset, Agg
    r = RegionDataset.from_records(ref, sample_ids=ref_ids)
    e = RegionDataset.from_records(exp, sample_ids=exp_ids, n_values=len(exp[0][5]) if exp else 2)
    plist = [(a, b) for a, bs in pairs.items() for b in bs]
    got = ex.GenometricMap71(plist, [Agg(k, i) for (k, i) in aggs], r, e)
    want = O.map_reference(ref, exp, pairs, [O.Agg(k, i) for (k, i) in aggs], 50000)
    assert canon(got, RTOL_DIGITS) == canon(want, RTOL_DIGITS)
    return got


@pytest.mark.parametrize("seed", range(6))
def test_map_random_small(ex, seed):
    rng = random.Random(seed)
    ref = rand_regions(rng, 60, [1, 2], n_vals=1, with_name=True, span=120000, max_len=60000, min_len=(0 if seed % 2 else 1))
    exp = rand_regions(rng, 400, [10, 11, 12], n_vals=2, span=120000, max_len=30000, min_len=(0 if seed % 2 else 1))
    pairs = {1: [10, 11, 12], 2: [10, 12]}
    aggs = [("COUNT", 0), ("SUM", 0), ("AVG", 1), ("MIN", 0), ("MAX", 1), ("MEDIAN", 1)]
    _run(ex, ref, exp, pairs, aggs, [1, 2], [10, 11, 12])


def test_map_no_aggregates_and_empty_exp_sample(ex):
    rng = random.Random(99)
    ref = rand_regions(rng, 30, [1], n_vals=0, span=50000, max_len=4000)
    exp = rand_regions(rng, 100, [10], n_vals=1, span=50000, max_len=4000)
    got = _run(ex, ref, exp, {1: [10, 11]}, [], [1], [10, 11])  # sample 11 has no regions: count-0 rows (A-M2)
    assert len(got) == 2 * len(set(ref))


def test_map_duplicate_reference_collapses(ex):
    ref = [(1, "chr1", 10, 50, "+", ("g", 1.0)), (1, "chr1", 10, 50, "+", ("g", 1.0)), (1, "chr1", 10, 50, "+", ("h", 1.0))]
    exp = [(7, "chr1", 20, 30, "*", (2.0,)), (7, "chr1", 49, 80, "-", (3.0,)), (7, "chr2", 0, 5, "*", (1.0,))]
    got = _run(ex, ref, exp, {1: [7]}, [("SUM", 0)], [1], [7])
    assert sorted(r[5] for r in got) == [("g", 1.0, 2.0, 4.0), ("h", 1.0, 1.0, 2.0)]


def test_map_all_null_column_and_strands(ex):
    ref = [(1, "chr1", 0, 100, "+", ()), (1, "chr1", 0, 100, "-", ()), (1, "chr1", 0, 100, "*", ())]
    exp = [(7, "chr1", 10, 20, "+", (None,)), (7, "chr1", 10, 20, "-", (None,)), (7, "chr1", 100, 120, "*", (1.0,))]
    _run(ex, ref, exp, {1: [7]}, [("SUM", 0), ("AVG", 0), ("MIN", 0), ("COUNT", 0)], [1], [7])


def test_map_conf_test_map_shape(ex):
    """conf/test_map.xml shape, thinned: 1 ref sample x 21 chr, 10 exp samples, bin-invariance against bin 2000."""
    rng = random.Random(7)
    chroms = ["chr%d" % i for i in range(1, 22)]
    ref = rand_regions(rng, 800, [1], chroms=chroms, n_vals=1, with_name=True, span=100000, max_len=500, min_len=20, aligned_p=0.05)
    ref = list(dict.fromkeys(ref))
    exp_ids = list(range(100, 110))
    exp = rand_regions(rng, 6000, exp_ids, chroms=chroms, n_vals=1, strands="*", span=100000, max_len=500, min_len=20, aligned_p=0.05, null_p=0.0)
    pairs = {1: exp_ids}
    from gmql_b200 import RegionDataset, Agg
    r = RegionDataset.from_records(ref, sample_ids=[1])
    e = RegionDataset.from_records(exp, sample_ids=exp_ids)
    got = ex.GenometricMap71(None, [Agg("MAX", 0), Agg("MIN", 0), Agg("SUM", 0), Agg("AVG", 0)], r, e)
    want = O.map_reference(ref, exp, pairs, [O.Agg("MAX", 0), O.Agg("MIN", 0), O.Agg("SUM", 0), O.Agg("AVG", 0)], 2000)
    assert canon(got, 7) == canon(want, 7)


@pytest.mark.parametrize("seed", range(3))
def test_map_bag_and_bagd(ex, seed):
    """BAG / BAGD (DefaultRegionsToRegionFactory.scala:128-168; in the reference's own matrix, conf/test_map.xml:92): the device
    emits the matched experiment rows of every output row, the host folds them pairwise in experiment order — on numeric
    columns (Double.toString forms, nulls as '.') and on string columns, next to COUNT and a numeric aggregate."""
    from gmql_b200 import RegionDataset, Agg
    rng = random.Random(600 + seed)
    ref = rand_regions(rng, 60, [1, 2], n_vals=1, span=60000, max_len=4000)
    ref = list({(r[0], r[1], r[2], r[3], r[4]): r for r in ref}.values())       # no duplicate reference records: their fold order is arbitrary
    exp = rand_regions(rng, 500, [10, 11], n_vals=1, with_name=True, span=60000, max_len=3000, null_p=0.2)
    exp = [(e[0], e[1], e[2], e[3], e[4], (e[5][0], e[5][1] if rng.random() > 0.3 else float(rng.choice([1e7, 1e-4, 12345678.5, 0.001, 3.0])))) for e in exp]
    pairs = {1: [10, 11], 2: [11]}
    aggs = [("BAG", 1), ("BAGD", 0), ("COUNT", 0), ("MAX", 1), ("BAGD", 1)]
    got = ex.GenometricMap71([(a, b) for a, bs in pairs.items() for b in bs], [Agg(k, i) for k, i in aggs],
                             RegionDataset.from_records(ref, [1, 2]), RegionDataset.from_records(exp, [10, 11]))
    want = O.map_direct(ref, exp, pairs, [O.Agg(k, i) for k, i in aggs])
    assert canon(got, 9) == canon(want, 9)
    assert any(isinstance(r[5][-5], str) and "," in r[5][-5] for r in got)        # some multi-element bags were built
