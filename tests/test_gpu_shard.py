"""Coordinate-range sharding of COVER / FLAT (gmql_cover_shard + dist.merge_cover_shards): every shard is computed on its own
(here one after the other on one GPU; each call sees only the rows dist.shard_by_tiles gives it), the few edge runs are stitched
on the host, and the result equals the unsharded operator row for row — runs crossing one cut, runs spanning whole shards, cuts
that fall into gaps, FLAT coordinates and both Jaccard doubles included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ANYV = 2 ** 31 - 1


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


@pytest.fixture(autouse=True, params=["count", "rank"])
def tile_kernel(request):
    import os
    os.environ["GMQL_COVER_KERNEL"] = request.param
    yield request.param
    os.environ.pop("GMQL_COVER_KERNEL", None)


def _unsharded(ex, flag, mn, mx, ds):
    import gmql_b200 as G
    from gmql_b200 import _lib as L
    res, _ = ex.cover_raw(flag, G.N(mn), G.ANY() if mx == ANYV else G.N(mx), [], None, ds)
    out = {k: res.column(c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64),
                                                  ("acc", L.COL_COV_ACC, np.int32), ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64))}
    res.free()
    return out


def _sharded(ex, flag, mn, mx, ds, span, cuts):
    from gmql_b200 import dist as D
    tables = []
    for r in range(len(cuts) - 1):
        part = D.shard_by_tiles(ds, span, cuts[r], cuts[r + 1])
        tables.append(ex.cover_shard_raw(flag, mn, mx, part, span, cuts[r], cuts[r + 1]))
    return D.merge_cover_shards(tables, flat=(flag == "FLAT")), tables


def _same(a, b):
    assert len(a["start"]) == len(b["start"]) and len(a["start"]) > 0
    for k in ("chrom", "start", "stop", "acc", "j1", "j2"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
@pytest.mark.parametrize("world", [2, 5])
def test_range_shards_equal_unsharded(ex, flag, world):
    from gmql_b200 import synth, dist as D
    ds = synth.peak_samples(60, 50000, 12, with_values=())
    span = np.array([int(ds.stop[ds.chrom == c].max(initial=0)) for c in range(len(ds.chrom_names))], dtype=np.int64) + 7     # any upper bound
    cuts = D.plan_tile_cuts(ds, span, world)
    assert cuts[0] == 0 and all(a < b for a, b in zip(cuts, cuts[1:]))
    loads = [D.shard_by_tiles(ds, span, cuts[r], cuts[r + 1]).n for r in range(world)]
    assert max(loads) < 1.05 * ds.n / world + 2000                      # balanced by construction (halo included)
    for mn, mx in ((2, ANYV), (1, 3)):
        got, tables = _sharded(ex, flag, mn, mx, ds, span, cuts)
        _same(got, _unsharded(ex, flag, mn, mx, ds))


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
def test_runs_spanning_whole_shards(ex, flag):
    """A pile of long regions: runs of hundreds of kilobases, cuts every three tiles — most shards hold a single run that
    both continues from the left and goes on to the right."""
    from gmql_b200 import RegionDataset
    rng = np.random.Generator(np.random.PCG64(21))
    n = 8000
    start = np.sort(rng.integers(0, 1_500_000, n))
    ln = rng.integers(2000, 60000, n)
    gap = (start > 400_000) & (start < 470_000)                          # a hole, so that some runs end inside the range
    start, ln = start[~gap], ln[~gap]
    m = len(start)
    ds = RegionDataset(["chr1", "chr2"], np.concatenate([np.zeros(m, np.int32), np.ones(5, np.int32)]), np.concatenate([start, [10, 20, 30, 40, 50]]),
                       np.concatenate([start + ln, [100, 120, 130, 140, 150]]), np.zeros(m + 5, np.int8), rng.integers(0, 4, m + 5).astype(np.int32),
                       np.arange(4, dtype=np.int64))
    span = np.array([1_600_000, 1000], dtype=np.int64)
    n_tiles = ((int(span.sum()) + 2 * 2002) >> 16) + 1
    cuts = list(range(0, n_tiles, 3)) + [n_tiles]
    for mn, mx in ((2, ANYV), (5, 40)):
        got, tables = _sharded(ex, flag, mn, mx, ds, span, cuts)
        assert sum(int((t["edge"] == 3).sum()) for t in tables) >= 2 or mx != ANYV      # runs that span an entire shard exist
        _same(got, _unsharded(ex, flag, mn, mx, ds))


def test_shard_requirements_are_errors(ex):
    from gmql_b200 import synth, RegionDataset, GmqlCudaError
    ds = synth.peak_samples(4, 2000, 3, with_values=())
    span = np.asarray(synth.HG38_SIZES, dtype=np.int64)
    with pytest.raises(GmqlCudaError, match="tile-resident|UNSUPPORTED|needs"):
        ex.cover_shard_raw("SUMMIT", 1, ANYV, ds, span, 0, 10)
    short = np.zeros(len(ds.chrom_names), dtype=np.int64)
    with pytest.raises(GmqlCudaError, match="chrom_span"):
        ex.cover_shard_raw("COVER", 1, ANYV, ds, short, 0, 10)
    stranded = RegionDataset(ds.chrom_names, ds.chrom, ds.start, ds.stop, np.ones(ds.n, np.int8), ds.sample, ds.sample_ids)
    with pytest.raises(GmqlCudaError, match="unified"):
        ex.cover_shard_raw("COVER", 1, ANYV, stranded, span, 0, 10)
