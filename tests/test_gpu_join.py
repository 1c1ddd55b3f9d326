"""Genometric JOIN parity: CUDA path (through the C ABI) vs the oracle's literal restatement of GenometricJoin
(bin-replicated join on (bin, chrom), first-bin de-duplication, MD(k) with ties, second round, builders, DISTINCT).
Everything is integer work: exact."""
import random

import pytest

from oracle import gmql_oracle as O
from tests.randgen import rand_regions, canon

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["auto", "team", "team-nomemo", "lane"])
def ex(request):
    """Every test runs three times: the library's own choice of the MD kernel, the team form (a warp owns 32 consecutive left
    regions and walks the right samples, candidates staged in shared memory; its write pass from the count pass's memo or
    repeating the selection) and the lane form (one lane per group)."""
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    if request.param != "auto":
        e.set_option("join.md", request.param.split("-")[0])
    if request.param.endswith("nomemo"):
        e.set_option("join.md_memo", "0")
    yield e
    e.close()


def _atoms(spec):
    import gmql_b200 as G
    m = {"DISTLESS": G.DistLess, "DISTGREATER": G.DistGreater, "MD": G.MinDistance}
    out = []
    for k, a in spec:
        if k == "UP":
            out.append(G.Upstream())
        elif k == "DOWN":
            out.append(G.DownStream())
        else:
            out.append(m[k](a))
    return out


ATOM_SETS = [
    [("DISTLESS", 1001)],                                   # distance <= 1000  (conf/test_join.xml)
    [("DISTLESS", 1000)],                                   # distance < 1000
    [("MD", 1)],                                            # minDistance(1): pure-MD bin-quantised window
    [("DISTGREATER", 250), ("MD", 1)],                      # distance > 250, minDistance(1)
    [("DISTLESS", 1000), ("DISTGREATER", 100)],
    [("MD", 2), ("DOWN", 0)],
    [("DOWN", 0), ("MD", 1)],
    [("UP", 0), ("MD", 2)],
    [("DISTLESS", 50000), ("UP", 0), ("MD", 3), ("DISTGREATER", 10)],
    [("MD", 1), ("DISTLESS", 3000)],
    [("DISTLESS", 10001), ("MD", 1), ("DOWN", 0)],
    [("DISTLESS", 0)],                                      # only overlapping (negative distance)
    [("DISTLESS", -5)],
]


def _check(ex, left, right, lids, rids, pairs, spec, builder, join_on=None):
    from gmql_b200 import RegionDataset
    l = RegionDataset.from_records(left, sample_ids=lids)
    r = RegionDataset.from_records(right, sample_ids=rids)
    got = ex.GenometricJoin(list(pairs.keys()), _atoms(spec), builder, l, r, join_on)
    want = O.join_reference(left, right, pairs, spec, builder, 5000, 100000, join_on)
    assert canon(got) == canon(want), (spec, builder)
    return got


@pytest.mark.parametrize("seed", range(4))
def test_join_random(ex, seed):
    rng = random.Random(500 + seed)
    left = rand_regions(rng, 50, [1, 2], span=400000, max_len=3000, n_vals=1, bin_mults=(5000,))
    right = rand_regions(rng, 400, [10, 11, 12], span=400000, max_len=3000, n_vals=1, bin_mults=(5000,))
    pairs = {(1, 10): O.md5_pair_id(1, 10), (1, 11): O.md5_pair_id(1, 11), (2, 10): O.md5_pair_id(2, 10), (2, 12): O.md5_pair_id(2, 12)}
    for spec in ATOM_SETS:
        _check(ex, left, right, [1, 2], [10, 11, 12], pairs, spec, "LEFT")


@pytest.mark.parametrize("builder", ["LEFT", "LEFT_DISTINCT", "RIGHT", "RIGHT_DISTINCT", "CONTIG", "INTERSECTION", "BOTH"])
def test_join_builders(ex, builder):
    rng = random.Random(42)
    left = rand_regions(rng, 40, [1, 2], span=60000, max_len=4000, n_vals=1, bin_mults=(5000,))
    right = rand_regions(rng, 200, [10, 11], span=60000, max_len=4000, n_vals=1, bin_mults=(5000,))
    left = left + left[:5]       # identical left records: MapKey pooling / DISTINCT
    right = right + right[:7]
    pairs = {(a, b): O.md5_pair_id(a, b) for a in (1, 2) for b in (10, 11)}
    for spec in ([("DISTLESS", 2001)], [("MD", 1)], [("MD", 2), ("DISTLESS", 30000)], [("DISTLESS", 5000), ("MD", 3)]):
        _check(ex, left, right, [1, 2], [10, 11], pairs, spec, builder)


def test_join_pure_md_window_edges(ex):
    """pure MD(k): candidates are bin-quantised around +-100000 (A-J3): 99 999 / 100 000 / 104 999 / 105 000 bp away."""
    left = [(1, "chr1", 200000, 200100, "+", ())]
    right = []
    for gap in (99999, 100000, 104899, 104999, 105000, 110000):
        right.append((10 + len(right), "chr1", 200100 + gap, 200100 + gap + 50, "*", ()))
    rids = [r[0] for r in right]
    pairs = {(1, b): O.md5_pair_id(1, b) for b in rids}
    _check(ex, left, right, [1], rids, pairs, [("MD", 1)], "RIGHT")
    _check(ex, left, right, [1], rids, pairs, [("MD", 1), ("DOWN", 0)], "RIGHT")
    _check(ex, left, right, [1], rids, pairs, [("DISTGREATER", 100000), ("MD", 1)], "RIGHT")


def test_join_md_ties_and_strands(ex):
    left = [(1, "chr1", 1000, 1100, "+", ()), (1, "chr1", 1000, 1100, "-", ()), (1, "chr1", 1000, 1100, "*", ())]
    right = [(10, "chr1", 800, 900, "*", ()), (10, "chr1", 1200, 1300, "*", ()), (10, "chr1", 1200, 1250, "+", ()),
             (10, "chr1", 1050, 1060, "-", ()), (10, "chr1", 500, 600, "+", ())]
    pairs = {(1, 10): O.md5_pair_id(1, 10)}
    for spec in ([("MD", 1)], [("MD", 2)], [("MD", 1), ("UP", 0)], [("UP", 0), ("MD", 1)], [("DOWN", 0), ("MD", 2)], [("MD", 10)]):
        _check(ex, left, right, [1], [10], pairs, spec, "LEFT")


def test_join_on_attributes(ex):
    rng = random.Random(9)
    left = rand_regions(rng, 30, [1], span=30000, max_len=2000, n_vals=1, null_p=0.0)
    right = rand_regions(rng, 150, [10], span=30000, max_len=2000, n_vals=1, null_p=0.0)
    pairs = {(1, 10): O.md5_pair_id(1, 10)}
    _check(ex, left, right, [1], [10], pairs, [("DISTLESS", 20000)], "LEFT", join_on=[(0, 0)])
    _check(ex, left, right, [1], [10], pairs, [("MD", 2)], "LEFT", join_on=[(0, 0)])


def test_join_conf_test_join_shape(ex):
    """conf/test_join.xml: 5x5 files x 5 chr; distance<1000 with left|contig|int, minDistance(1) right, ..."""
    rng = random.Random(3)
    chroms = ["chr%d" % i for i in range(1, 6)]
    lids, rids = list(range(1, 6)), list(range(11, 16))
    left = rand_regions(rng, 300, lids, chroms=chroms, span=100000, max_len=500, min_len=20, n_vals=1, aligned_p=0.05)
    right = rand_regions(rng, 600, rids, chroms=chroms, span=100000, max_len=500, min_len=20, n_vals=1, aligned_p=0.05)
    pairs = {(a, b): O.md5_pair_id(a, b) for a in lids for b in rids}
    _check(ex, left, right, lids, rids, pairs, [("DISTLESS", 1000)], "LEFT")
    _check(ex, left, right, lids, rids, pairs, [("DISTLESS", 1000)], "CONTIG")
    _check(ex, left, right, lids, rids, pairs, [("DISTLESS", 1000)], "INTERSECTION")
    _check(ex, left, right, lids, rids, pairs, [("MD", 1)], "RIGHT")
    _check(ex, left, right, lids, rids, pairs, [("DISTGREATER", 250), ("MD", 1)], "LEFT")
    _check(ex, left, right, lids, rids, pairs, [("DISTLESS", 1000), ("DISTGREATER", 100)], "LEFT")


def test_join_md_int64_coordinates(ex):
    """Coordinates beyond 2^31 (int64 columns): the MD round then runs on the unpacked index instead of the 16-byte packed
    candidate records; same results as the oracle."""
    rng = random.Random(77)
    off = 3_000_000_000
    left = [(r[0], r[1], r[2] + off, r[3] + off, r[4], r[5]) for r in rand_regions(rng, 60, [1, 2], n_vals=1, span=400000, max_len=3000)]
    right = [(r[0], r[1], r[2] + off, r[3] + off, r[4], r[5]) for r in rand_regions(rng, 400, [10, 11, 12], n_vals=1, span=400000, max_len=3000)]
    pairs = {(a, b): O.md5_pair_id(a, b) for a in (1, 2) for b in (10, 11, 12)}
    for spec, builder in (([("MD", 1)], "RIGHT"), ([("MD", 2), ("DOWN", 0)], "LEFT"), ([("DISTLESS", 20001), ("MD", 3)], "CONTIG")):
        _check(ex, left, right, [1, 2], [10, 11, 12], pairs, spec, builder)


def test_join_md_dense_right_and_many_samples(ex):
    """More right samples than a warp has lanes (the team kernel searches 32 samples at a time), several chromosomes inside one
    team of 32 left regions, and right samples dense enough that a team's candidate list overflows its shared-memory stage
    (every lane then streams its own window)."""
    rng = random.Random(2024)
    chroms = ["chr1", "chr2", "chr3"]
    lids, rids = [1, 2, 3], list(range(10, 10 + 45))
    left = rand_regions(rng, 90, lids, chroms=chroms, span=300000, max_len=2500, n_vals=0, bin_mults=(5000,))
    left = left + left[:6]                                     # identical left records pool their candidate lists
    right = rand_regions(rng, 45 * 12, rids, chroms=chroms, span=300000, max_len=2500, n_vals=0, bin_mults=(5000,))
    dense = rand_regions(rng, 900, [rids[3], rids[40]], chroms=["chr1"], span=120000, max_len=300, n_vals=0)   # > 64 candidates per team
    right = right + dense
    pairs = {(a, b): O.md5_pair_id(a, b) for a in lids for b in rids if (a + b) % 7 != 0}
    for spec, builder in (([("MD", 1)], "RIGHT"), ([("MD", 3), ("DOWN", 0)], "LEFT"), ([("UP", 0), ("MD", 2)], "CONTIG"),
                          ([("DISTLESS", 30001), ("MD", 1), ("DISTGREATER", 50)], "LEFT"), ([("DISTGREATER", 1000), ("MD", 2)], "BOTH")):
        _check(ex, left, right, lids, rids, pairs, spec, builder)
