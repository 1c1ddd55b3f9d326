"""BASELINE.json's full sizes (C2: 5e7 regions, C5: 1e8 regions) through size-independent properties: the two independent
device paths of COVER agree row for row, HISTOGRAM's depth-weighted length equals the summed (extended) region lengths
(linearity), COVER output is sorted, disjoint and never touching, and the MAP counts of the C5 pipeline add up to the number of
(region, overlapped run) pairs computed by numpy from the runs alone."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _upload(ex, ds):
    from gmql_b200 import _lib as L
    host, keep = ds.as_c(narrow=True)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    return dset, view, keep


def _cover(ex, view, flag, mn, mx, path=None):
    from gmql_b200 import _lib as L
    a, b = np.array([mn], np.int32), np.array([mx], np.int32)
    if path:
        os.environ["GMQL_COVER_PATH"] = path
    try:
        out = C.c_void_p()
        ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, flag, a.ctypes.data, b.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out)))
    finally:
        os.environ.pop("GMQL_COVER_PATH", None)
    return out


def _cols(ex, res, ids_dtypes):
    outs = []
    for cid, dt in ids_dtypes:
        n = int(ex.lib.gmql_result_column_len(res, cid))
        a = np.empty(n, dtype=dt)
        if n:
            ex._check(ex.lib.gmql_result_column_copy(ex.ctx, res, cid, a.ctypes.data))
        outs.append(a)
    return outs


def test_c2_full_size_properties(ex):
    from gmql_b200 import synth, _lib as L
    ds = synth.peak_samples(1000, 50000, 2, with_values=())
    assert ds.n == 50_000_000
    dset, view, keep = _upload(ex, ds)
    ANYV = 2 ** 31 - 1
    spec = [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)]
    # (1) the tile-resident path and the sort-and-merge path are independent algorithms: identical rows
    rt, rs = _cover(ex, view, 0, 2, ANYV, "tiled"), _cover(ex, view, 0, 2, ANYV, "sort")
    ct, cs = _cols(ex, rt, spec), _cols(ex, rs, spec)
    ex.lib.gmql_result_free(ex.ctx, rt); ex.lib.gmql_result_free(ex.ctx, rs)
    assert len(ct[0]) > 100000
    for a, b in zip(ct, cs):
        assert np.array_equal(a, b)
    chrom, start, stop = ct[0], ct[1], ct[2]
    key = chrom.astype(np.int64) << 32
    assert np.all(stop > start) and np.all(np.diff(key + start) > 0)
    same = chrom[1:] == chrom[:-1]
    assert np.all(start[1:][same] > stop[:-1][same])                 # maximal runs never touch
    assert int(ct[3].min()) >= 2
    # (2) linearity: sum over HISTOGRAM(1,ANY) segments of depth x length == sum of the extended region lengths
    rh = _cover(ex, view, 3, 1, ANYV)
    hs, he, hacc = _cols(ex, rh, [(L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32)])
    ex.lib.gmql_result_free(ex.ctx, rh)
    ext = ds.stop + ((ds.stop % 2000 == 0) & (ds.start // 2000 != ds.stop // 2000))
    assert int(((he - hs) * hacc.astype(np.int64)).sum()) == int((ext - ds.start).sum())
    ex.lib.gmql_dataset_free(ex.ctx, dset)


def test_c5_full_size_pipeline(ex):
    from gmql_b200 import synth, _lib as L
    ds = synth.peak_samples(2000, 50000, 5, with_values=())
    assert ds.n == 100_000_000
    dset, view, keep = _upload(ex, ds)
    cov = _cover(ex, view, 0, 2, 2 ** 31 - 1)
    cview = L.Regions()
    ex._check(ex.lib.gmql_result_as_regions(ex.ctx, cov, C.byref(cview)))
    pairs = (L.Pair * 2000)()
    for b in range(2000):
        pairs[b].left_sample, pairs[b].right_sample = 0, b
    mp = C.c_void_p()
    ex._check(ex.lib.gmql_map(ex.ctx, C.byref(cview), C.byref(view), pairs, 2000, (L.Agg * 1)(), 0, C.byref(mp)))
    chrom, start, stop = _cols(ex, cov, [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64)])
    (count,) = _cols(ex, mp, [(L.COL_MAP_COUNT, np.int32)])
    ex.lib.gmql_result_free(ex.ctx, cov); ex.lib.gmql_result_free(ex.ctx, mp)
    n_runs = len(start)
    assert count.shape[0] == 2000 * n_runs and count.min() >= 0
    # runs are disjoint and sorted: a region overlaps the runs between the first run ending after its start and the last
    # run starting before its stop
    rk_s = (chrom.astype(np.int64) << 32) + start
    rk_e = (chrom.astype(np.int64) << 32) + stop
    ck = ds.chrom.astype(np.int64) << 32
    first = np.searchsorted(rk_e, ck + ds.start, side="right")
    last = np.searchsorted(rk_s, ck + ds.stop, side="left")
    per_region = np.maximum(last - first, 0)
    assert int(count.astype(np.int64).sum()) == int(per_region.sum())
    # per experiment sample as well (rows of pair b are contiguous)
    per_sample = np.bincount(ds.sample, weights=per_region, minlength=2000).astype(np.int64)
    assert np.array_equal(count.reshape(2000, n_runs).astype(np.int64).sum(axis=1), per_sample)
    ex.lib.gmql_dataset_free(ex.ctx, dset)


def test_c4_full_size_join(ex):
    """C4: 200 x 200 samples (1e6 promoters, 1e7 peaks).  DLE(10000): the number of output pairs equals the count numpy
    derives from sorted starts / stops per chromosome, and every emitted pair satisfies the predicate; MD(1): every left
    region keeps, per right sample, only candidates at its minimum distance (checked on the pairs themselves)."""
    from gmql_b200 import synth, _lib as L
    left = synth.promoter_samples(200, 5000, 4)
    right = synth.peak_samples(200, 50000, 4, with_values=())
    dl, vl, k1 = _upload(ex, left)
    dr, vr, k2 = _upload(ex, right)
    pairs = (L.Pair * (200 * 200))()
    for a in range(200):
        for b in range(200):
            pairs[a * 200 + b].left_sample, pairs[a * 200 + b].right_sample = a, b

    def join(atoms, builder, n_pairs=200 * 200):
        at = (L.Atom * len(atoms))()
        for i, (kd, arg) in enumerate(atoms):
            at[i].kind, at[i].arg = kd, arg
        out = C.c_void_p()
        ex._check(ex.lib.gmql_join(ex.ctx, C.byref(vl), C.byref(vr), pairs, n_pairs, at, len(atoms), builder | 0x100, 5000, 100000, None, None, C.byref(out)))
        lrow, rrow = _cols(ex, out, [(L.COL_JOIN_LEFT_ROW, np.int32), (L.COL_JOIN_RIGHT_ROW, np.int32)])
        ex.lib.gmql_result_free(ex.ctx, out)
        return lrow, rrow

    def dist(lrow, rrow):
        ls, le, rs, re_ = left.start[lrow], left.stop[lrow], right.start[rrow], right.stop[rrow]
        m = np.minimum(np.abs(ls - re_), np.abs(rs - le))
        return np.where((le < rs) | (re_ < ls), m, -m)

    lrow, rrow = join([(0, 10001)], 4)                       # DLE(10000), CONTIG
    want = 0
    for c in range(24):
        lm, rm = left.chrom == c, right.chrom == c
        rs, re_ = np.sort(right.start[rm]), np.sort(right.stop[rm])
        # distance < 10001  <=>  r.start < l.stop + 10001  and  r.stop > l.start - 10001 (strands are compatible: peaks are '*')
        want += int((np.searchsorted(rs, left.stop[lm] + 10001, side="left") - np.searchsorted(re_, left.start[lm] - 10001, side="right")).sum())
    assert len(lrow) == want and want > 5 * 10 ** 7
    assert np.array_equal(left.chrom[lrow], right.chrom[rrow]) and int(dist(lrow, rrow).max()) < 10001
    key = lrow.astype(np.int64) * right.n + rrow
    assert len(np.unique(key)) == len(key)                  # no pair twice

    lrow, rrow = join([(2, 1)], 2, 30 * 200)                 # MD(1), RIGHT; the first 30 left samples x every right sample
    d = dist(lrow, rrow)
    grp = lrow.astype(np.int64) * 200 + right.sample[rrow]  # (left region, right sample) group
    o = np.argsort(grp, kind="stable")
    g, dd = grp[o], d[o]
    first = np.concatenate([[True], g[1:] != g[:-1]])
    mins = np.minimum.reduceat(dd, np.nonzero(first)[0])
    assert np.array_equal(dd, np.repeat(mins, np.diff(np.concatenate([np.nonzero(first)[0], [len(g)]]))))   # only minimum-distance candidates survive
    assert len(lrow) >= int(first.sum()) > 10 ** 7 and np.array_equal(left.chrom[lrow], right.chrom[rrow])
    ex.lib.gmql_dataset_free(ex.ctx, dl); ex.lib.gmql_dataset_free(ex.ctx, dr)


# ---- BASELINE.json's configs at FULL size, row for row against the C oracle (oracle/gmql_oracle_c.c: the bin-based plan of
# GenometricCover.scala:98-152 + GMAP4.scala:20-100 and GenometricMap71.scala:81-147), not only through properties -------------
def _sorted_rows(cols, keys=("chrom", "start", "stop")):
    o = np.lexsort(tuple(cols[k] for k in reversed(keys)))
    return {k: v[o] for k, v in cols.items()}


@pytest.mark.parametrize("flag", [0, 1, 2, 3], ids=["COVER", "FLAT", "SUMMIT", "HISTOGRAM"])
def test_c2_full_size_against_the_oracle(ex, flag):
    """C2: (2, ANY) over 1 000 x 50 000 peaks (5e7 regions) — every row of every variant: coordinates, AccIndex and both Jaccard
    columns, bit for bit."""
    from gmql_b200 import synth, _lib as L
    from oracle import c_oracle as CO
    CO.set_threads(os.cpu_count() or 1)
    ds = synth.peak_samples(1000, 50000, 2, with_values=())
    dset, view, keep = _upload(ex, ds)
    res = _cover(ex, view, flag, 2, 2 ** 31 - 1)
    spec = [("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64), ("acc", L.COL_COV_ACC, np.int32),
            ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64)]
    got = dict(zip([s[0] for s in spec], _cols(ex, res, [(s[1], s[2]) for s in spec])))
    ex.lib.gmql_result_free(ex.ctx, res)
    ex.lib.gmql_dataset_free(ex.ctx, dset)
    want = CO.cover_rows(ds, flag, [2], [2 ** 31 - 1], None, 2000)
    want = {k: want[k] for k in got}
    assert len(got["start"]) == len(want["start"]) > 100000
    # FLAT rows are the extents of the contributing regions: neighbouring runs under the same long regions give identical
    # (chrom, start, stop), so the comparison orders the rows by every column
    allk = ("chrom", "start", "stop", "acc", "j1", "j2")
    got, want = _sorted_rows(got, allk), _sorted_rows(want, allk)
    for k in got:
        assert np.array_equal(got[k], want[k]), k


def test_c5_full_size_against_the_oracle(ex):
    """C5: the bench's own step at its own size (2 000 x 50 000 peaks, 1e8 regions, production route on the resident dataset):
    the COVER table and all 2 000 x n_runs MAP counts equal the oracle's."""
    from gmql_b200 import synth, _lib as L
    from gmql_b200.regions import RegionDataset
    from oracle import c_oracle as CO
    CO.set_threads(os.cpu_count() or 1)
    ds = synth.peak_samples(2000, 50000, 5, with_values=())
    dset, view, keep = _upload(ex, ds)
    cov = _cover(ex, view, 0, 2, 2 ** 31 - 1)
    cview = L.Regions()
    ex._check(ex.lib.gmql_result_as_regions(ex.ctx, cov, C.byref(cview)))
    pairs = (L.Pair * 2000)()
    for b in range(2000):
        pairs[b].left_sample, pairs[b].right_sample = 0, b
    mp = C.c_void_p()
    ex._check(ex.lib.gmql_map(ex.ctx, C.byref(cview), C.byref(view), pairs, 2000, (L.Agg * 1)(), 0, C.byref(mp)))
    names = ("chrom", "start", "stop", "acc", "j1", "j2")
    got = dict(zip(names, _cols(ex, cov, [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32),
                                          (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)])))
    (count,) = _cols(ex, mp, [(L.COL_MAP_COUNT, np.int32)])
    ex.lib.gmql_result_free(ex.ctx, cov); ex.lib.gmql_result_free(ex.ctx, mp)
    ex.lib.gmql_dataset_free(ex.ctx, dset)
    r = CO.cover_rows(ds, 0, [2], [2 ** 31 - 1], None, 2000)
    r = _sorted_rows(r, ("chrom", "start"))
    n_runs = len(r["start"])
    assert n_runs == len(got["start"]) > 5000
    for k in names:
        assert np.array_equal(got[k], r[k]), k                   # the device emits (chrom, start) order
    ref = RegionDataset(ds.chrom_names, r["chrom"], r["start"], r["stop"], r["strand"], np.zeros(n_runs, np.int32), np.array([0], np.int64))
    pidx = np.stack([np.zeros(2000, np.int32), np.arange(2000, dtype=np.int32)], axis=1)
    want = CO.map_rows(ref, ds, pidx, None, None, 50000)["count"]
    assert np.array_equal(count, want)


def test_c3_full_size_against_the_oracle(ex):
    """C3: MAP with SUM / AVG / MAX of signalValue, 100 reference x 1 000 experiment samples (1e8 output rows): counts and MAX
    bit-exact, SUM / AVG within 1e-9 relative (the association order of a floating-point sum differs, SURVEY A-M8)."""
    from gmql_b200 import synth, _lib as L
    from oracle import c_oracle as CO
    CO.set_threads(os.cpu_count() or 1)
    nr, ne = 100, 1000
    ref = synth.promoter_samples(nr, 1000, 3)
    exp = synth.peak_samples(ne, 50000, 3, with_values=("signalValue",))
    dr, vr, k1 = _upload(ex, ref)
    de, ve, k2 = _upload(ex, exp)
    pairs = (L.Pair * (nr * ne))()
    for a in range(nr):
        for b in range(ne):
            pairs[a * ne + b].left_sample, pairs[a * ne + b].right_sample = a, b
    aggs = (L.Agg * 3)()
    for i, kind in enumerate((1, 4, 3)):                         # SUM, AVG, MAX
        aggs[i].kind, aggs[i].col = kind, 0
    out = C.c_void_p()
    ex._check(ex.lib.gmql_map(ex.ctx, C.byref(vr), C.byref(ve), pairs, nr * ne, aggs, 3, C.byref(out)))
    count, v_sum, v_avg, v_max, ok_sum, ok_max = _cols(ex, out, [(L.COL_MAP_COUNT, np.int32), (L.COL_AGG_VALUE, np.float64), (L.COL_AGG_VALUE + 1, np.float64),
                                                               (L.COL_AGG_VALUE + 2, np.float64), (L.COL_AGG_VALID, np.uint8), (L.COL_AGG_VALID + 2, np.uint8)])
    ex.lib.gmql_result_free(ex.ctx, out)
    ex.lib.gmql_dataset_free(ex.ctx, dr); ex.lib.gmql_dataset_free(ex.ctx, de)
    pidx = np.array([(a, b) for a in range(nr) for b in range(ne)], dtype=np.int32)
    val = [c for c in exp.columns if c[0] == "num"][0][1]
    want = CO.map_rows(ref, exp, pidx, val=val, valid=None, bin_size=50000)
    assert count.shape[0] == 100_000_000 == want["count"].shape[0]
    assert np.array_equal(count, want["count"]) and int(count.astype(np.int64).sum()) > 10 ** 6
    has = want["nn"] > 0
    assert np.array_equal(ok_sum.astype(bool), has) and np.array_equal(ok_max.astype(bool), has)
    assert np.array_equal(v_max[has], want["max"][has])
    np.testing.assert_allclose(v_sum[has], want["sum"][has], rtol=1e-9)
    np.testing.assert_allclose(v_avg[has], want["sum"][has] / want["nn"][has], rtol=1e-9)


def _md1_reference(left, right, lrows, first_max, down, B=5000, max_bin_distance=100000):
    """Bin-free numpy restatement of JOIN [DLE(first_max - 1),] MD(1) [, DOWN] for the left rows `lrows` against EVERY right sample
    (GenometricJoin.scala:62-81 parameter split, :284-309 bin window, :117-127 first round, :130-151 MD(k) with ties, :160-188 second
    round).  Returns the sorted keys lrow * right.n + rrow.  Right strands are '*' (compatible with everything)."""
    n_chrom = 24
    maxd = first_max if first_max is not None else max_bin_distance          # :75-79
    key = (right.sample.astype(np.int64) * n_chrom + right.chrom) << 32 | right.start.astype(np.int64)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    maxlen = int((right.stop - right.start).max())
    n_rs = len(right.sample_ids)
    out = []
    for b in range(n_rs):
        ls, le, lst, lc = left.start[lrows], left.stop[lrows], left.strand[lrows], left.chrom[lrows]
        b0 = np.maximum(ls - maxd, 0) // B
        b1 = (le + maxd) // B
        wbeg, wend = b0 * B, (b1 + 1) * B
        base = (np.int64(b) * n_chrom + lc) << 32
        lo = np.searchsorted(skey, base + np.maximum(wbeg - maxlen, 0), side="left")
        hi = np.searchsorted(skey, base + wend, side="left")                  # r.start / B <= binEnd
        cnt = hi - lo
        g = np.repeat(np.arange(len(lrows)), cnt)                             # group = position in lrows
        r = order[np.repeat(lo, cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))]
        rs, re_ = right.start[r], right.stop[r]
        gls, gle = ls[g], le[g]
        m = np.minimum(np.abs(gls - re_), np.abs(rs - gle))
        d = np.where((gle < rs) | (re_ < gls), m, -m)                          # distanceCalculator :375-386
        ok = re_ >= wbeg[g]                                                   # r.stop / B >= binStart
        if first_max is not None:
            ok &= d < first_max                                               # :250
        g, r, d, rs, re_ = g[ok], r[ok], d[ok], rs[ok], re_[ok]
        dmin = np.full(len(lrows), np.iinfo(np.int64).max)
        np.minimum.at(dmin, g, d)
        keep = d == dmin[g]                                                   # k = 1: the minimum with its ties
        if down:                                                              # second round, :171-183 (left strand only; '*' as '+')
            keep &= np.where(lst[g] != 2, rs >= le[g], re_ <= ls[g])
        out.append(lrows[g[keep]].astype(np.int64) * right.n + r[keep])
    return np.sort(np.concatenate(out))


def test_c4_md_full_size_exact_on_a_sample_of_left_regions(ex):
    """C4's two MD joins at full size (200 x 200 samples, 2e8 (left region, right sample) groups, the team kernels): for the
    10 000 left regions of two left samples the emitted (left row, right row) pairs are EXACTLY those of an independent bin-free
    numpy restatement — complete and exact, not only the minimum-distance property — and the lane kernel gives the same rows for
    the whole join."""
    from gmql_b200 import synth, _lib as L
    left = synth.promoter_samples(200, 5000, 4)
    right = synth.peak_samples(200, 50000, 4, with_values=())
    dl, vl, k1 = _upload(ex, left)
    dr, vr, k2 = _upload(ex, right)
    pairs = (L.Pair * (200 * 200))()
    for a in range(200):
        for b in range(200):
            pairs[a * 200 + b].left_sample, pairs[a * 200 + b].right_sample = a, b

    def join(atoms, kernel):
        at = (L.Atom * len(atoms))()
        for i, (kd, arg) in enumerate(atoms):
            at[i].kind, at[i].arg = kd, arg
        ex.set_option("join.md", kernel)
        try:
            out = C.c_void_p()
            ex._check(ex.lib.gmql_join(ex.ctx, C.byref(vl), C.byref(vr), pairs, 200 * 200, at, len(atoms), 2 | 0x100, 5000, 100000, None, None, C.byref(out)))
        finally:
            ex.set_option("join.md", None)
        lrow, rrow = _cols(ex, out, [(L.COL_JOIN_LEFT_ROW, np.int32), (L.COL_JOIN_RIGHT_ROW, np.int32)])
        ex.lib.gmql_result_free(ex.ctx, out)
        return np.sort(lrow.astype(np.int64) * right.n + rrow)

    lrows = np.nonzero((left.sample == 3) | (left.sample == 117))[0]
    for atoms, first_max in (([(2, 1), (4, 0)], None), ([(0, 10001), (2, 1), (4, 0)], 10001)):   # MD(1), DOWN;  DLE(10000), MD(1), DOWN
        team = join(atoms, "team")
        want = _md1_reference(left, right, lrows, first_max, down=True)
        got = team[np.isin(team // right.n, lrows)]
        assert len(want) > 10 ** 5 and np.array_equal(got, want)
        assert np.array_equal(team, join(atoms, "lane"))
    ex.lib.gmql_dataset_free(ex.ctx, dl); ex.lib.gmql_dataset_free(ex.ctx, dr)
