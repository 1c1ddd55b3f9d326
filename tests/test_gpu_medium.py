"""Medium-scale parity (sizes the literal Python oracle cannot finish): CUDA path vs the C oracle port on the
BASELINE.json config shapes, plus size-independent properties on larger inputs.  Exercises 32-bit linear keys, many sort
tiles with decoupled look-back, multi-tile merge sweeps and the bucket-table search."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def test_c1_map_counts_and_aggregates(ex):
    """C1: MAP of a 10 000-region reference onto 10 x 42 000-region BedScore samples (conf/test_map.xml shape)."""
    from gmql_b200 import synth, Agg, _lib as L
    from oracle import c_oracle as CO
    ref, exp = synth.c1_datasets()
    pairs = [(int(ref.sample_ids[0]), int(b)) for b in exp.sample_ids]
    res, used, ref2 = ex.map_raw(pairs, [Agg("SUM", 0), Agg("MIN", 0), Agg("MAX", 0), Agg("AVG", 0)], ref, exp)
    count = res.column(L.COL_MAP_COUNT, np.int32)
    vals = [res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(4)]
    oks = [res.column(L.COL_AGG_VALID + k, np.uint8) for k in range(4)]
    res.free()
    pidx = np.stack([np.zeros(10, np.int32), np.arange(10, dtype=np.int32)], axis=1)
    want = CO.map_rows(ref, exp, pidx, val=exp.columns[0][1], valid=None, bin_size=50000)
    assert count.shape == want["count"].shape == (100000,)
    assert np.array_equal(count, want["count"])
    has = want["nn"] > 0
    assert np.array_equal(oks[0].astype(bool), has)
    np.testing.assert_allclose(vals[0][has], want["sum"][has], rtol=1e-9)
    assert np.array_equal(vals[1][has], want["min"][has])
    assert np.array_equal(vals[2][has], want["max"][has])
    np.testing.assert_allclose(vals[3][has], want["sum"][has] / want["nn"][has], rtol=1e-9)


@pytest.mark.parametrize("flag", ["COVER", "FLAT", "SUMMIT", "HISTOGRAM"])
def test_c2_shape_cover_family(ex, flag):
    """C2 shape, thinned: 60 narrowPeak samples x 20 000 regions on hg38 chr19-22 sizes (dense pile-ups, 32-bit keys)."""
    from gmql_b200 import synth, Agg, N, ANY, _lib as L
    from oracle import c_oracle as CO
    sizes = synth.HG38_SIZES[[18, 19, 20, 21]] // 2
    ds = synth.peak_samples(60, 20000, 2, with_values=("signalValue",), sizes=sizes)
    code = {"COVER": 0, "FLAT": 1, "SUMMIT": 2, "HISTOGRAM": 3}[flag]
    for (mn, mx, gmin, gmax) in ((2, 2 ** 31 - 1, N(2), ANY()), (3, 7, N(3), N(7))):
        res, gids = ex.cover_raw(flag, gmin, gmax, [Agg("SUM", 0), Agg("MAX", 0), Agg("COUNT", 0)], None, ds)
        g = {k: res.column(c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64),
                                                   ("stop", L.COL_COV_STOP, np.int64), ("acc", L.COL_COV_ACC, np.int32),
                                                   ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64))}
        s, m, c = (res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(3))
        res.free()
        w = CO.cover_rows(ds, code, [mn], [mx], None, 2000, val=ds.columns[0][1], valid=None)
        # FLAT rows of different runs may share (chrom, start): order on the full integer key
        # (many FLAT rows of one pile-up share every coordinate: order on the whole row, the float sum rounded, last)
        og = np.lexsort((np.round(s, 5), m, c, g["j2"], g["j1"], g["acc"], g["stop"], g["start"], g["chrom"]))
        ow = np.lexsort((np.round(w["sum"], 5), w["max"], w["count"], w["j2"], w["j1"], w["acc"], w["stop"], w["start"], w["chrom"]))
        assert len(og) == len(ow) and len(og) > 50
        for k in ("chrom", "start", "stop", "acc", "j1", "j2"):
            assert np.array_equal(g[k][og], w[k][ow]), (flag, mn, k)
        np.testing.assert_allclose(s[og], w["sum"][ow], rtol=1e-9)
        assert np.array_equal(m[og], w["max"][ow])
        assert np.array_equal(c[og], w["count"][ow].astype(np.float64))


def test_join_dle_medium_against_numpy(ex):
    """JOIN DLE(10000) between promoters and peaks (C4 shape, thinned), checked against a vectorised numpy statement of
    the distance predicate (GenometricJoin.scala:375-386, strict comparison :250)."""
    from gmql_b200 import synth, DistLess, _lib as L
    left = synth.promoter_samples(3, 1500, 4)
    right = synth.peak_samples(4, 30000, 4, with_values=())
    pairs = [(int(a), int(b)) for a in left.sample_ids for b in right.sample_ids]
    res, used, l, r = ex.join_raw(pairs, [DistLess(10001)], "LEFT", left, right, dedup=False)
    lrow = res.column(L.COL_JOIN_LEFT_ROW, np.int32)
    rrow = res.column(L.COL_JOIN_RIGHT_ROW, np.int32)
    res.free()
    got = np.sort(lrow.astype(np.int64) * right.n + rrow)
    want = []
    for c in range(24):
        ri = np.nonzero(r.chrom == c)[0]
        if not len(ri):
            continue
        rs, re_ = r.start[ri], r.stop[ri]
        for i in np.nonzero(l.chrom == c)[0]:
            ls, le = l.start[i], l.stop[i]
            d1, d2 = np.abs(ls - re_), np.abs(rs - le)
            m = np.minimum(d1, d2)
            dist = np.where((le < rs) | (re_ < ls), m, -m)
            hit = ri[dist < 10001]          # every strand pair here is compatible ('*' on the right)
            want.append(i * right.n + hit)
    want = np.sort(np.concatenate(want))
    assert len(want) > 1000 and np.array_equal(got, want)


def test_properties_at_scale(ex):
    """2e6 regions: COVER(1,ANY) output is sorted and disjoint, its total length equals the union of the inputs computed by
    numpy, COVER is idempotent on its own output, and MAP counts equal the two-binary-search count."""
    from gmql_b200 import synth, N, ANY, RegionDataset, _lib as L
    ds = synth.peak_samples(40, 50000, 9, with_values=())
    res, _ = ex.cover_raw("COVER", N(1), ANY(), [], None, ds)
    chrom = res.column(L.COL_COV_CHROM, np.int32)
    start = res.column(L.COL_COV_START, np.int64)
    stop = res.column(L.COL_COV_STOP, np.int64)
    res.free()
    assert np.all(stop > start)
    key = chrom.astype(np.int64) * (1 << 32)
    assert np.all(np.diff(key + start) > 0)                       # sorted by (chrom, start)
    same = chrom[1:] == chrom[:-1]
    assert np.all(start[1:][same] > stop[:-1][same])              # maximal runs never touch
    # union length by numpy (stop+1 extension for stops on a 2000 multiple, SURVEY A-C4)
    ext = ds.stop + ((ds.stop % 2000 == 0) & (ds.start // 2000 != ds.stop // 2000))
    o = np.lexsort((ds.start, ds.chrom))
    c, s, e = ds.chrom[o], ds.start[o], ext[o]
    lin_s = c.astype(np.int64) * (1 << 32) + s
    lin_e = c.astype(np.int64) * (1 << 32) + e
    run_end = np.maximum.accumulate(lin_e)
    new_run = np.concatenate([[True], lin_s[1:] > run_end[:-1]])
    n_runs = int(new_run.sum())
    assert n_runs == len(start)
    union_len = int((np.maximum.reduceat(lin_e, np.nonzero(new_run)[0]) - lin_s[new_run]).sum())
    assert union_len == int((stop - start).sum())
    # idempotence
    cov = RegionDataset(ds.chrom_names, chrom, start, stop, np.zeros(len(start), np.int8), np.zeros(len(start), np.int32), np.array([0], np.int64))
    res2, _ = ex.cover_raw("COVER", N(1), ANY(), [], None, cov)
    s2 = res2.column(L.COL_COV_START, np.int64)
    e2 = res2.column(L.COL_COV_STOP, np.int64)
    res2.free()
    # re-covering can only add the one-base extension where a run stops on a bin multiple
    assert len(s2) <= len(start) and np.array_equal(np.unique(s2), np.unique(s2))
    assert int((e2 - s2).sum()) >= union_len
    # MAP count of the cover runs over the inputs == searchsorted count
    pairs = [(0, int(b)) for b in ds.sample_ids[:3]]
    sub = RegionDataset(ds.chrom_names, ds.chrom, ds.start, ds.stop, ds.strand, ds.sample, ds.sample_ids)
    resm, used, _ref = ex.map_raw(pairs, [], cov, sub)
    cnt = resm.column(L.COL_MAP_COUNT, np.int32)
    resm.free()
    n = len(start)
    for p, (_a, b) in enumerate(used):
        bi = int(np.nonzero(ds.sample_ids == b)[0][0])
        m = ds.sample == bi
        ls = np.sort(ds.chrom[m].astype(np.int64) * (1 << 32) + ds.start[m])
        le = np.sort(ds.chrom[m].astype(np.int64) * (1 << 32) + ds.stop[m])
        want = np.searchsorted(ls, key + stop, side="left") - np.searchsorted(le, key + start, side="right")
        assert np.array_equal(cnt[p * n:(p + 1) * n], want)


@pytest.mark.parametrize("flag", ["COVER", "HISTOGRAM", "SUMMIT"])
def test_cover_64bit_linear_keys(ex, flag):
    """Stranded input on hg38-sized chromosomes: 2 strand classes x 3.1e9 positions exceed 2^32, so the sort keys, the index and
    the region-driven GMAP4 run on 64-bit linear coordinates."""
    from gmql_b200 import synth, RegionDataset, Agg, N, ANY, _lib as L
    from oracle import c_oracle as CO
    base = synth.peak_samples(30, 20000, 8, with_values=("signalValue",))
    rng = np.random.Generator(np.random.PCG64(4))
    strand = rng.integers(1, 3, base.n).astype(np.int8)                       # '+' / '-' only: the strands stay separate
    # concentrate the regions so that pile-ups exist
    start = (base.start // 4000).astype(np.int64) * 40 + (base.chrom == 0) * 200_000_000
    ds = RegionDataset(base.chrom_names, base.chrom, start, start + (base.stop - base.start), strand, base.sample, base.sample_ids, base.columns)
    code = {"COVER": 0, "SUMMIT": 2, "HISTOGRAM": 3}[flag]
    res, gids = ex.cover_raw(flag, N(2), ANY(), [Agg("MAX", 0), Agg("COUNT", 0)], None, ds)
    g = {k: res.column(c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64),
                                               ("strand", L.COL_COV_STRAND, np.int8), ("acc", L.COL_COV_ACC, np.int32), ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64))}
    m, c = (res.column(L.COL_AGG_VALUE + k, np.float64) for k in range(2))
    res.free()
    w = CO.cover_rows(ds, code, [2], [2 ** 31 - 1], None, 2000, val=ds.columns[0][1], valid=None)
    og = np.lexsort((g["j2"], g["j1"], g["acc"], g["stop"], g["start"], g["strand"], g["chrom"]))
    ow = np.lexsort((w["j2"], w["j1"], w["acc"], w["stop"], w["start"], w["strand"], w["chrom"]))
    assert len(og) == len(ow) > 1000 and set(np.unique(g["strand"])) == {1, 2}
    for k in ("chrom", "start", "stop", "strand", "acc", "j1", "j2"):
        assert np.array_equal(g[k][og], w[k][ow]), (flag, k)
    assert np.array_equal(m[og], w["max"][ow]) and np.array_equal(c[og], w["count"][ow].astype(np.float64))


def test_map_64bit_linear_keys(ex):
    """MAP with chromosome coordinates beyond 2^31 (int64 columns, a linear genome above 2^32): counts equal the numpy
    two-search count."""
    from gmql_b200 import RegionDataset, _lib as L
    rng = np.random.Generator(np.random.PCG64(9))
    big = 3_000_000_000
    nr, ne = 3000, 200000
    rs = np.sort(rng.integers(0, big, nr) // 1000 * 1000 + rng.integers(0, 2, nr) * big // 2) % big
    ref = RegionDataset(["chrA", "chrB"], rng.integers(0, 2, nr).astype(np.int32), rs, rs + rng.integers(100, 3_000_000, nr), np.zeros(nr, np.int8), np.zeros(nr, np.int32),
                        np.array([0], np.int64))
    es = rng.integers(0, big, ne)
    exp = RegionDataset(["chrA", "chrB"], rng.integers(0, 2, ne).astype(np.int32), es, es + rng.integers(1, 500_000, ne), np.zeros(ne, np.int8),
                        np.sort(rng.integers(0, 3, ne)).astype(np.int32), np.arange(3, dtype=np.int64))
    res, used, ref2 = ex.map_raw([(0, 0), (0, 1), (0, 2)], [], ref, exp)
    cnt = res.column(L.COL_MAP_COUNT, np.int32)
    res.free()
    assert int(cnt.max()) > 0
    for p, (_a, b) in enumerate(used):
        for c in (0, 1):
            m = (exp.sample == b) & (exp.chrom == c)
            ls, le = np.sort(exp.start[m]), np.sort(exp.stop[m])
            rmask = ref.chrom == c
            want = np.searchsorted(ls, ref.stop[rmask], side="left") - np.searchsorted(le, ref.start[rmask], side="right")
            got = cnt[p * nr:(p + 1) * nr][rmask]
            sel = got >= 0                                   # rows collapsed into a duplicate carry -1
            assert np.array_equal(got[sel], want[sel])
