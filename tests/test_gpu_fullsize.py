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
