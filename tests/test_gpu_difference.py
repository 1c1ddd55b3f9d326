"""DIFFERENCE parity: CUDA path (gmql_difference through the C ABI) vs the oracle's literal restatement of
GenometricDifference.scala (50 000-base bins, cogroup, reduce on the record's content)."""
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


def _run(ex, ref, exp, ref_ids, exp_ids, groups, exact):
    from gmql_b200 import RegionDataset
    pairs = [(a, b) for a, bs in groups.items() for b in bs]
    got = ex.GenometricDifference(pairs, RegionDataset.from_records(ref, ref_ids), RegionDataset.from_records(exp, exp_ids), exact=exact)
    want = O.difference_reference(ref, exp, groups, exact=exact)
    assert canon(got, 9) == canon(want, 9)
    return got


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_difference_random(ex, seed, exact):
    rng = random.Random(4000 + seed)
    strands = "+-*" if seed % 2 == 0 else "+-"
    ref = rand_regions(rng, 300, [1, 2, 3], n_vals=1, with_name=(seed % 3 == 0), span=160000, max_len=(60000 if seed % 2 else 3000), min_len=0, strands=strands)
    exp = rand_regions(rng, 200, [10, 11, 12], n_vals=1, span=160000, max_len=4000, min_len=0, strands=strands)
    if exact:      # plant identical coordinates (with compatible and incompatible strands)
        for k in range(0, 120, 3):
            r = ref[k]
            exp.append((rng.choice([10, 11, 12]), r[1], r[2], r[3], rng.choice(strands), (1.0,)))
    groups = {1: [10, 11, 12], 2: [11]}            # sample 3 has no group: its regions produce nothing
    got = _run(ex, ref, exp, [1, 2, 3], [10, 11, 12], groups, exact)
    assert 0 < len(got) < len(ref)


def test_difference_duplicates_and_empty_groups(ex):
    ref = [(1, "chr1", 100, 200, "+", (1.0,)), (1, "chr1", 100, 200, "+", (1.0,)), (1, "chr1", 100, 200, "+", (2.0,)),
           (1, "chr1", 5000, 6000, "*", (3.0,)), (2, "chr1", 100, 200, "+", (1.0,)), (2, "chr2", 0, 10, "-", (4.0,))]
    exp = [(7, "chr1", 150, 160, "-", ()), (7, "chr1", 5999, 6001, "+", ()), (8, "chr2", 9, 20, "-", ())]
    # '+' vs '-' never match; the '*' region is hit by sample 7; sample 2 is paired with 8 only
    got = _run(ex, ref, exp, [1, 2], [7, 8], {1: [7], 2: [8]}, False)
    assert len(got) == 3          # (1,+,1.0) once, (1,+,2.0), (2,chr1,+)
    _run(ex, ref, exp, [1, 2], [7, 8], {1: [7, 8]}, False)
    _run(ex, ref, exp, [1, 2], [7, 8], {1: [], 2: [7]}, False)
    _run(ex, ref, [], [1, 2], [7, 8], {1: [7], 2: [8]}, False)


def test_difference_exact_needs_strict_overlap(ex):
    # identical zero-length regions never satisfy start < stop (GenometricDifference.scala:61)
    ref = [(1, "chr1", 50, 50, "*", ()), (1, "chr1", 60, 70, "*", ()), (1, "chr1", 80, 90, "*", ())]
    exp = [(7, "chr1", 50, 50, "*", ()), (7, "chr1", 60, 70, "*", ()), (7, "chr1", 80, 91, "*", ())]
    got = _run(ex, ref, exp, [1], [7], {1: [7]}, True)
    assert sorted((r[2], r[3]) for r in got) == [(50, 50), (80, 90)]


def test_difference_large_against_direct(ex):
    from gmql_b200 import synth, RegionDataset
    ref = synth.promoter_samples(2, 4000, 6)
    exp = synth.peak_samples(3, 40000, 6, with_values=())
    pairs = [(int(a), int(b)) for a in ref.sample_ids for b in exp.sample_ids]
    got = ex.GenometricDifference(pairs, ref, exp)
    groups = {int(a): [int(b) for b in exp.sample_ids] for a in ref.sample_ids}
    want = O.difference_direct(ref.to_records(), exp.to_records(), groups)
    assert len(want) > 100 and canon(got, 9) == canon(want, 9)
