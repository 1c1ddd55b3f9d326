"""The ABI's compact input encodings (include/gmql_cuda.h: chrom_bits = 8/16, sample_offsets) give the same results as the
plain int32 columns: MAP, COVER and JOIN on sample-ordered inputs, through both executors."""
import random

import numpy as np
import pytest

from tests.randgen import rand_regions, canon

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exs():
    from gmql_b200 import GMQLCudaExecutor
    a, b = GMQLCudaExecutor(), GMQLCudaExecutor(narrow=True)
    yield a, b
    a.close()
    b.close()


def _sorted_by_sample(recs):
    return sorted(recs, key=lambda r: r[0])


def test_narrow_struct_is_used():
    from gmql_b200 import RegionDataset
    rng = random.Random(1)
    d = RegionDataset.from_records(_sorted_by_sample(rand_regions(rng, 50, [1, 2, 3], n_vals=1)), sample_ids=[1, 2, 3])
    r, keep = d.as_c(narrow=True)
    assert r.chrom_bits == 8 and not r.sample and r.sample_offsets
    d2 = RegionDataset.from_records(rand_regions(rng, 50, [3, 1, 2], n_vals=1)[::-1], sample_ids=[1, 2, 3])
    r2, keep2 = d2.as_c(narrow=True)
    assert r2.chrom_bits == 8
    assert bool(r2.sample) == (not bool(np.all(np.diff(d2.sample) >= 0)))


def test_narrow_map_cover_join(exs):
    import gmql_b200 as G
    from gmql_b200 import RegionDataset, Agg
    plain, narrow = exs
    rng = random.Random(42)
    ref = _sorted_by_sample(rand_regions(rng, 80, [1, 2], n_vals=1, span=100000, max_len=5000))
    exp = _sorted_by_sample(rand_regions(rng, 900, [10, 11, 12], n_vals=1, span=100000, max_len=5000))
    dr, de = RegionDataset.from_records(ref, [1, 2]), RegionDataset.from_records(exp, [10, 11, 12])
    aggs = [Agg("SUM", 0), Agg("MAX", 0), Agg("COUNT", 0)]
    assert canon(plain.GenometricMap71(None, aggs, dr, de), 9) == canon(narrow.GenometricMap71(None, aggs, dr, de), 9)
    for flag in ("COVER", "FLAT", "SUMMIT", "HISTOGRAM"):
        a = plain.GenometricCover(flag, G.N(2), G.ANY(), [Agg("SUM", 0)], None, de)
        b = narrow.GenometricCover(flag, G.N(2), G.ANY(), [Agg("SUM", 0)], None, de)
        assert len(a) > 0 and canon(a, 9) == canon(b, 9)
    pairs = [(a, b) for a in (1, 2) for b in (10, 11, 12)]
    ja = plain.GenometricJoin(pairs, [G.DistLess(2000)], "LEFT", dr, de)
    jb = narrow.GenometricJoin(pairs, [G.DistLess(2000)], "LEFT", dr, de)
    assert len(ja) > 0 and canon(ja, 9) == canon(jb, 9)


def test_bad_sample_offsets_rejected(exs):
    import ctypes as C
    from gmql_b200 import RegionDataset, _lib as L
    plain, narrow = exs
    rng = random.Random(3)
    d = RegionDataset.from_records(_sorted_by_sample(rand_regions(rng, 40, [1, 2], n_vals=0)), sample_ids=[1, 2])
    r, keep = d.as_c(narrow=True)
    off = np.array([0, 50, 40], dtype=np.int64)      # decreasing
    r.sample_offsets = off.ctypes.data
    out = C.c_void_p()
    mn, mx = np.array([1], np.int32), np.array([5], np.int32)
    rc = narrow.lib.gmql_cover(narrow.ctx, C.byref(r), None, 1, 0, mn.ctypes.data, mx.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out))
    assert rc < 0 and b"sample_offsets" in narrow.lib.gmql_last_error(narrow.ctx)
