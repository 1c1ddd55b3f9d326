"""Tile-resident COVER / FLAT path (gmql_b200/csrc/cover_tiled.cuh) against the oracle.

The path is chosen by a density heuristic; the tests force it (GMQL_COVER_PATH=tiled) and shrink the tile width
(GMQL_COVER_LOGW) so that small inputs cross many tile edges: runs spanning several tiles, regions of the previous tile
still open at a tile start, events exactly on a tile edge, the one-base extension of GenometricCover.scala:353-354 and
GMAP4's exactly-once accounting (GMAP4.scala:44-45).  Every test checks from the per-kernel profile that the tile kernel
ran and the sort-and-merge sweep did not (no silent fall-back)."""
import ctypes as C
import os
import random

import numpy as np
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


@pytest.fixture(autouse=True, params=["count", "rank"])
def tile_kernel(request):
    """Every test of this module runs with both tile kernels: the word-count kernel (cover_tile_wc.cuh, the default where there
    is one group) and the bitmap-rank kernel (its fall-back, and the only one for several groups)."""
    os.environ["GMQL_COVER_KERNEL"] = request.param
    yield request.param
    os.environ.pop("GMQL_COVER_KERNEL", None)


class forced:
    """Force the tiled path at a given tile width and record which kernels ran."""

    def __init__(self, ex, logw):
        self.ex, self.logw = ex, logw

    def __enter__(self):
        from gmql_b200 import _lib as L
        os.environ["GMQL_COVER_PATH"] = "tiled"
        os.environ["GMQL_COVER_LOGW"] = str(self.logw)
        self.ex.lib.gmql_ctx_set_profiling(self.ex.ctx, 1)
        return self

    def __exit__(self, *a):
        from gmql_b200 import _lib as L
        os.environ.pop("GMQL_COVER_PATH", None)
        os.environ.pop("GMQL_COVER_LOGW", None)
        buf = C.create_string_buffer(1 << 16)
        self.ex.lib.gmql_ctx_profile_report(self.ex.ctx, buf, len(buf))
        self.ex.lib.gmql_ctx_set_profiling(self.ex.ctx, 0)
        self.kernels = buf.value.decode()
        return False

    def assert_tiled(self):
        assert "k_cover_tile" in self.kernels, self.kernels
        assert "k_cover_runs_fused" not in self.kernels, "fell back to the sort-and-merge path"


def _params():
    import gmql_b200 as G
    return [
        ((G.N(1), G.N(2)), (O.CoverParam.N(1), O.CoverParam.N(2))),
        ((G.N(2), G.N(2)), (O.CoverParam.N(2), O.CoverParam.N(2))),
        ((G.N(1), G.ANY()), (O.CoverParam.N(1), O.CoverParam.ANY())),
        ((G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY())),
        ((G.N(3), G.N(6)), (O.CoverParam.N(3), O.CoverParam.N(6))),
        ((G.ANY(), G.ALL()), (O.CoverParam.ANY(), O.CoverParam.ALL())),
        ((G.ALL(div=2), G.ANY()), (O.CoverParam.ALL(div=2), O.CoverParam.ANY())),
    ]


def _clamp(rng, ds, max_len):
    """regions strictly shorter than max_len (the aligned variants of rand_regions can stretch a region), keeping an
    aligned stop aligned"""
    out = []
    for (sid, chrom, start, stop, strand, vals) in ds:
        if stop - start >= max_len:
            ln = rng.randrange(1, max_len)
            if stop % 64 == 0 and stop - ln >= 0:
                start = stop - ln
            else:
                stop = start + ln
        out.append((sid, chrom, start, stop, strand, vals))
    return out


def _check(ex, flag, ds, sample_ids, groups, gp, op):
    from gmql_b200 import RegionDataset
    d = RegionDataset.from_records(ds, sample_ids=sample_ids)
    got = ex.GenometricCover(flag, gp[0], gp[1], [], groups, d)
    want = O.cover_reference(flag, op[0], op[1], [], ds, groups, 2000)
    assert canon(got, 9) == canon(want, 9), (flag, gp)


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
@pytest.mark.parametrize("logw", [8, 10, 12, 16])
@pytest.mark.parametrize("seed", range(3))
def test_tiled_random(ex, flag, logw, seed):
    rng = random.Random(900 + seed)
    ids = [1, 2, 3, 4, 5]
    strands = "+-*" if seed == 0 else "+-"
    max_len = min(4500, (1 << logw) - 3)
    mults = (2000, 5000, 50000) if logw >= 12 else (64, 256, 2000)     # coordinates on tile edges and bin edges
    ds = _clamp(rng, rand_regions(rng, 400, ids, strands=strands, n_vals=0, max_len=max_len, aligned_p=0.4, bin_mults=mults), max_len)
    groups = None if seed < 2 else {1: 50, 2: 50, 3: 90, 4: 90, 5: 90}
    with forced(ex, logw) as f:
        for gp, op in _params():
            _check(ex, flag, ds, ids, groups, gp, op)
    f.assert_tiled()
    if groups is None:
        assert ("k_cover_tile_wc" in f.kernels) == (os.environ["GMQL_COVER_KERNEL"] == "count"), f.kernels


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
def test_tiled_long_regions_keep_full_16_bit_lengths(ex, flag):
    """With every region shorter than 32 768 bases the partition stores the extractor's one-base extension in bit 15 of the 4-byte
    record (test_tiled_random runs that format); a longer region needs all 16 length bits and the tile kernels derive the extension
    from the chromosome offsets themselves.  Same rows as the oracle either way, including bin-aligned stops of the long regions."""
    rng = random.Random(4242)
    ids = [1, 2, 3, 4]
    ds = _clamp(rng, rand_regions(rng, 500, ids, strands="*", n_vals=0, max_len=3000, aligned_p=0.4, bin_mults=(2000, 5000, 50000)), 3000)
    chroms = sorted({r[1] for r in ds})
    for k in range(12):
        st = rng.randrange(0, 200000)
        ln = rng.randrange(33000, 65000)
        stop = st + ln if k % 3 else ((st + ln) // 2000) * 2000
        ds.append((ids[k % 4], chroms[k % len(chroms)], st, stop, "*", ()))
    with forced(ex, 16) as f:
        for gp, op in _params()[:5]:
            _check(ex, flag, ds, ids, None, gp, op)
    f.assert_tiled()


def test_word_count_kernel_hands_over_when_a_tile_has_too_many_words_to_resolve(ex):
    """Two long regions give depth 2 over most of a tile; a chain of short regions, each ending five bases into a 32-position word
    and the next one starting five bases later, adds a third layer that is open at every word boundary.  Every word then holds a
    stop and a start at the plateau's maximum: its upper depth bound (4) exceeds every boundary depth (3), so the word-count kernel
    must resolve ~1 800 words of one tile — more than WC_NCAP = 384.  It raises its flag and the operator is recomputed by the
    bitmap-rank kernel on the same tiles: same rows as the oracle (one run, AccIndex 3), both kernels ran."""
    import gmql_b200 as G
    if os.environ["GMQL_COVER_KERNEL"] != "count":
        pytest.skip("word-count kernel only")
    ds = [(1, "chr1", 100, 60100, "*", ()), (2, "chr1", 100, 60100, "*", ())]
    for k in range(5, 1800):
        ds.append((3, "chr1", 32 * k + 10, 32 * k + 37, "*", ()))
    with forced(ex, 16) as f:
        _check(ex, "COVER", ds, [1, 2, 3], None, (G.N(2), G.ANY()), (O.CoverParam.N(2), O.CoverParam.ANY()))
        _check(ex, "FLAT", ds, [1, 2, 3], None, (G.N(2), G.N(3)), (O.CoverParam.N(2), O.CoverParam.N(3)))    # the bound 4 straddles max = 3
    f.assert_tiled()
    assert "k_cover_tile_wc" in f.kernels and "k_cover_tile<false>" in f.kernels, f.kernels


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
def test_tiled_tile_edge_events(ex, flag):
    """starts, stops and extended stops exactly on tile edges (W = 256: multiples of 256; bin edge 2000 -> stop + 1)"""
    import gmql_b200 as G
    ds = [(1, "chr1", 3, 256, "*", ()), (2, "chr1", 260, 512, "*", ()), (1, "chr1", 256, 300, "*", ()), (1, "chr1", 255, 257, "*", ()), (2, "chr1", 520, 768, "*", ()),
          (1, "chr1", 1800, 2000, "*", ()), (2, "chr1", 1999, 2048, "*", ()), (1, "chr1", 2047, 2049, "*", ()), (2, "chr1", 1900, 2000, "*", ()),
          (1, "chr1", 3840, 4000, "*", ()), (2, "chr1", 3900, 4095, "*", ()), (1, "chr1", 4000, 4096, "*", ()), (2, "chr1", 4096, 4100, "*", ()),
          (1, "chr2", 0, 250, "*", ()), (2, "chr2", 10, 20, "*", ()), (1, "chr2", 100, 300, "*", ())]
    with forced(ex, 8) as f:
        for mn, mx in ((1, None), (2, None), (1, 1), (2, 2), (1, 2)):
            gp = (G.N(mn), G.ANY() if mx is None else G.N(mx))
            op = (O.CoverParam.N(mn), O.CoverParam.ANY() if mx is None else O.CoverParam.N(mx))
            _check(ex, flag, ds, [1, 2], None, gp, op)
    f.assert_tiled()


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
def test_tiled_against_c_oracle_medium(ex, flag):
    """60 narrowPeak samples x 20 000 regions (C2 shape, thinned) at the production tile width and at a narrow one."""
    from gmql_b200 import synth, N, ANY, _lib as L
    from oracle import c_oracle as CO
    sizes = synth.HG38_SIZES[[18, 19, 20, 21]] // 2
    ds = synth.peak_samples(60, 20000, 2, with_values=(), sizes=sizes)
    code = {"COVER": 0, "FLAT": 1}[flag]
    for logw in (16, 13):
        for (mn, mx, gmin, gmax) in ((2, 2 ** 31 - 1, N(2), ANY()), (3, 7, N(3), N(7))):
            with forced(ex, logw) as f:
                res, gids = ex.cover_raw(flag, gmin, gmax, [], None, ds)
            f.assert_tiled()
            g = {k: res.column(c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64),
                                                       ("stop", L.COL_COV_STOP, np.int64), ("acc", L.COL_COV_ACC, np.int32),
                                                       ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64))}
            res.free()
            w = CO.cover_rows(ds, code, [mn], [mx], None, 2000, val=None, valid=None)
            og = np.lexsort((g["j2"], g["j1"], g["acc"], g["stop"], g["start"], g["chrom"]))
            ow = np.lexsort((w["j2"], w["j1"], w["acc"], w["stop"], w["start"], w["chrom"]))
            assert len(og) == len(ow) and len(og) > 50
            for k in ("chrom", "start", "stop", "acc", "j1", "j2"):
                assert np.array_equal(g[k][og], w[k][ow]), (flag, logw, mn, k)


def test_tiled_equals_sort_path_at_scale(ex):
    """4e6 regions: the two device paths agree row for row (COVER(2,ANY) and FLAT(1,3))."""
    from gmql_b200 import synth, N, ANY, _lib as L
    ds = synth.peak_samples(80, 50000, 5, with_values=())
    for flag, gmin, gmax in (("COVER", N(2), ANY()), ("FLAT", N(1), N(3))):
        out = {}
        for path in ("tiled", "sort"):
            os.environ["GMQL_COVER_PATH"] = path
            try:
                res, _ = ex.cover_raw(flag, gmin, gmax, [], None, ds)
            finally:
                os.environ.pop("GMQL_COVER_PATH", None)
            cols = [res.column(c, dt) for c, dt in ((L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64),
                                                    (L.COL_COV_ACC, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64))]
            res.free()
            o = np.lexsort(tuple(reversed(cols)))
            out[path] = [c[o] for c in cols]
        assert len(out["tiled"][0]) > 1000
        for a, b in zip(out["tiled"], out["sort"]):
            assert np.array_equal(a, b), flag


def test_tiled_pile_up_retries_with_narrower_tiles(ex):
    """A hotspot (15 000 regions inside 30 kb on top of a dense background) overflows a 65 536-wide tile's tables; the path
    retries with narrower tiles instead of leaving the input to the sort-and-merge path, and the rows still equal that path's."""
    from gmql_b200 import synth, N, ANY, RegionDataset, _lib as L
    bg = synth.peak_samples(40, 50000, 7, with_values=())
    rng = np.random.Generator(np.random.PCG64(3))
    k = 15000
    hs = 5_000_000 + rng.integers(0, 30000, k)
    hot = RegionDataset(bg.chrom_names, np.concatenate([bg.chrom, np.zeros(k, np.int32)]), np.concatenate([bg.start, hs]),
                        np.concatenate([bg.stop, hs + rng.integers(50, 400, k)]), np.zeros(bg.n + k, np.int8),
                        np.concatenate([bg.sample, rng.integers(0, 40, k).astype(np.int32)]), bg.sample_ids)
    out = {}
    for path in ("tiled", "sort"):
        os.environ["GMQL_COVER_PATH"] = path
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
        try:
            res, _ = ex.cover_raw("COVER", N(3), ANY(), [], None, hot)
        finally:
            os.environ.pop("GMQL_COVER_PATH", None)
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
        if path == "tiled":
            rep = buf.value.decode()
            line = [ln for ln in rep.splitlines() if "k_cover_tile" in ln][0]
            if os.environ["GMQL_COVER_KERNEL"] == "rank":
                assert int(line.rsplit(" ", 2)[1]) >= 2, rep      # the first width overflowed, a narrower one succeeded
            else:
                assert int(line.rsplit(" ", 2)[1]) == 1 and "k_cover_tile_wc" in line, rep   # no per-position table to overflow: one launch
            assert "k_cover_runs_fused" not in rep
        cols = [res.column(c, dt) for c, dt in ((L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64),
                                                (L.COL_COV_ACC, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64))]
        res.free()
        out[path] = cols
    assert len(out["tiled"][0]) > 100
    for a, b in zip(out["tiled"], out["sort"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("flag", ["COVER", "FLAT"])
@pytest.mark.parametrize("logw", [8, 12, 16])
def test_tiled_with_aggregates(ex, flag, logw):
    """Value aggregates on the tile-resident path (region-driven post-pass over the runs): SUM / MIN / MAX / AVG / COUNT with
    nulls, against the literal oracle."""
    import gmql_b200 as G
    from gmql_b200 import RegionDataset, Agg
    rng = random.Random(1200 + logw)
    ids = [1, 2, 3, 4]
    max_len = min(3000, (1 << logw) - 3)
    ds = _clamp(rng, rand_regions(rng, 500, ids, strands="+-*", n_vals=2, max_len=max_len, aligned_p=0.3, null_p=0.15), max_len)
    aggs = [("SUM", 0), ("MIN", 1), ("COUNT", 0), ("AVG", 1), ("MAX", 0)]
    d = RegionDataset.from_records(ds, sample_ids=ids)
    with forced(ex, logw) as f:
        for gp, op in _params()[:4]:
            got = ex.GenometricCover(flag, gp[0], gp[1], [Agg(k, i) for (k, i) in aggs], None, d)
            want = O.cover_reference(flag, op[0], op[1], [O.Agg(k, i) for (k, i) in aggs], ds, None, 2000)
            assert canon(got, 9) == canon(want, 9), (flag, gp)
    f.assert_tiled()
    assert "k_run_agg_scatter" in f.kernels


def test_tiled_aggregates_medium_against_c_oracle(ex):
    from gmql_b200 import synth, Agg, N, ANY, _lib as L
    from oracle import c_oracle as CO
    sizes = synth.HG38_SIZES[[18, 19, 20, 21]] // 2
    ds = synth.peak_samples(60, 20000, 2, with_values=("signalValue",), sizes=sizes)
    with forced(ex, 16) as f:
        res, gids = ex.cover_raw("COVER", N(2), ANY(), [Agg("SUM", 0), Agg("MAX", 0), Agg("COUNT", 0)], None, ds)
    f.assert_tiled()
    g = {k: res.column(c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64))}
    s, m, c = (res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(3))
    res.free()
    w = CO.cover_rows(ds, 0, [2], [2 ** 31 - 1], None, 2000, val=ds.columns[0][1], valid=None)
    og = np.lexsort((g["stop"], g["start"], g["chrom"]))
    ow = np.lexsort((w["stop"], w["start"], w["chrom"]))
    assert len(og) == len(ow) > 50
    for k in ("chrom", "start", "stop"):
        assert np.array_equal(g[k][og], w[k][ow])
    np.testing.assert_allclose(s[og], w["sum"][ow], rtol=1e-9)
    assert np.array_equal(m[og], w["max"][ow]) and np.array_equal(c[og], w["count"][ow].astype(np.float64))


@pytest.mark.parametrize("n_samples,expect_fused", [(10, True), (30, True), (800, False)])
def test_fused_stitch_equals_split_kernels(ex, n_samples, expect_fused):
    """The one-block stitch kernel (<= 262 144 pieces) and the separate stitch kernels give the same rows in the same order;
    with more pieces the fused kernel declines and the separate kernels run."""
    from gmql_b200 import synth, N, ANY, _lib as L
    ds = synth.peak_samples(n_samples, 5000, 7, with_values=())
    for flag, gmin, gmax in (("COVER", N(2), ANY()), ("FLAT", N(1), N(3))):
        out, kern = {}, {}
        for mode in ("fused", "split"):
            os.environ["GMQL_COVER_PATH"] = "tiled"
            if mode == "split":
                os.environ["GMQL_COVER_STITCH"] = "split"
            ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
            try:
                res, _ = ex.cover_raw(flag, gmin, gmax, [], None, ds)
            finally:
                os.environ.pop("GMQL_COVER_PATH", None)
                os.environ.pop("GMQL_COVER_STITCH", None)
            buf = C.create_string_buffer(1 << 16)
            ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
            ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
            kern[mode] = buf.value.decode()
            out[mode] = [res.column(c, dt) for c, dt in ((L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32),
                                                         (L.COL_COV_STRAND, np.int8), (L.COL_COV_GROUP, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64))]
            res.free()
        assert "k_cover_tile" in kern["fused"] and "k_stitch_small" in kern["fused"] and "k_stitch_small" not in kern["split"]
        assert ("k_piece_merge" not in kern["fused"]) == expect_fused, kern["fused"]
        assert len(out["fused"][0]) > 100
        for a, b in zip(out["fused"], out["split"]):
            assert np.array_equal(a, b), (flag, n_samples)
