"""Profiler statistics (gmql_profile) against the oracle's restatement of Profiler.profile's per-sample fold: all integer
fields exact, mean/variance computed from them by the same expressions (calculateResult, Profiler.scala:84-92)."""
import random

import numpy as np
import pytest

from oracle import gmql_oracle as O
from tests.randgen import rand_regions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


@pytest.mark.parametrize("ordered", [False, True])
def test_profile_random(ex, ordered):
    from gmql_b200 import RegionDataset
    rng = random.Random(55)
    ids = [3, 1, 4, 15, 9]                          # sample 9 gets no regions
    recs = rand_regions(rng, 3000, ids[:4], n_vals=0, span=10 ** 8, max_len=10 ** 6, min_len=0)
    if ordered:
        recs.sort(key=lambda r: ids.index(r[0]))   # one run of rows per sample: the warp-uniform path
    got = ex.profile(RegionDataset.from_records(recs, sample_ids=ids))
    want = O.profile_reference(recs, ids)
    assert got == want
    assert want[9]["count"] == 0 and want[3]["count"] > 0


def test_profile_scale_and_wraparound(ex):
    """2e6 regions; lengths near 2^31 make the sum of squares wrap like a Java long."""
    from gmql_b200 import RegionDataset
    rng = np.random.Generator(np.random.PCG64(7))
    n = 2_000_000
    sample = np.sort(rng.integers(0, 7, n)).astype(np.int32)
    start = rng.integers(0, 1000, n)
    stop = start + rng.integers(2 ** 31 - 5000, 2 ** 31 - 1, n)
    ds = RegionDataset(["chr1"], np.zeros(n, np.int32), start, stop, np.zeros(n, np.int8), sample, np.arange(7, dtype=np.int64))
    got = ex.profile(ds)
    ln = (stop - start).astype(np.int64)
    for s in range(7):
        m = sample == s
        with np.errstate(over="ignore"):
            sq = int(np.sum(ln[m].astype(np.uint64) * ln[m].astype(np.uint64), dtype=np.uint64).astype(np.int64))
        sl = int(ln[m].sum())
        c = int(m.sum())
        mean = float(sl) / c
        assert got[s]["count"] == c and got[s]["leftMost"] == int(start[m].min()) and got[s]["rightMost"] == int(stop[m].max())
        assert got[s]["minLength"] == int(ln[m].min()) and got[s]["maxLength"] == int(ln[m].max())
        assert got[s]["avgLength"] == mean and got[s]["varianceLength"] == float(sq) / c - mean * mean
