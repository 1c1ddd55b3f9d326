"""The production route of the C5 pipeline (`C = COVER(2,ANY) E; M = MAP() C E` on a resident dataset), pinned DIRECTLY to the
oracle: a C5-shaped input thinned so that every production selector fires without any override — the dataset is uploaded
in the ABI's narrow encodings through gmql_dataset_upload (statistics + partition offsets at upload), COVER starts at the
partition (k_multisplit<MSP_NARROW> -> k_multisplit<MSP_FINE> -> tile kernel -> k_stitch_small), MAP indexes the COVER view
with k_map_index_small and counts with k_map_count_grouped (>= 4 x SMs samples).  The kernel profile proves those kernels
ran; rows are compared one for one with oracle/c_oracle.py (GenometricCover.scala:98-152, GMAP4.scala:20-100,
GenometricMap71.scala:81-147).  Also: inputs the upload-time statistics rule out of the fast route, the deferred validity
check, per-context options, and the 16-bit counter overflow guard of the grouped kernel."""
import ctypes as C
import os

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


def _upload(ex, ds, **kw):
    from gmql_b200 import _lib as L
    host, keep = ds.as_c(narrow=True, **kw)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    return dset, view, keep


def _profile(ex, fn):
    ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
    try:
        out = fn()
    finally:
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
    return out, buf.value.decode()


def _cols(ex, res, spec):
    outs = []
    for cid, dt in spec:
        n = int(ex.lib.gmql_result_column_len(res, cid))
        a = np.empty(n, dtype=dt)
        if n:
            ex._check(ex.lib.gmql_result_column_copy(ex.ctx, res, cid, a.ctypes.data))
        outs.append(a)
    return outs


def _cover(ex, view, flag=0, mn=2, mx=ANYV):
    from gmql_b200 import _lib as L
    a, b = np.array([mn], np.int32), np.array([mx], np.int32)
    out = C.c_void_p()
    ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, flag, a.ctypes.data, b.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out)))
    return out


COV_SPEC = None


def _cov_spec():
    from gmql_b200 import _lib as L
    return [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32),
            (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)]


def _pipeline(ex, view, n_samples):
    """COVER(2,ANY) -> view -> MAP() on device-resident data; returns (cover columns, counts [n_samples, n_runs])."""
    from gmql_b200 import _lib as L
    cov = _cover(ex, view)
    cview = L.Regions()
    ex._check(ex.lib.gmql_result_as_regions(ex.ctx, cov, C.byref(cview)))
    pairs = (L.Pair * n_samples)()
    for b in range(n_samples):
        pairs[b].left_sample, pairs[b].right_sample = 0, b
    mp = C.c_void_p()
    ex._check(ex.lib.gmql_map(ex.ctx, C.byref(cview), C.byref(view), pairs, n_samples, (L.Agg * 1)(), 0, C.byref(mp)))
    cc = _cols(ex, cov, _cov_spec())
    (count,) = _cols(ex, mp, [(L.COL_MAP_COUNT, np.int32)])
    ex.lib.gmql_result_free(ex.ctx, cov)
    ex.lib.gmql_result_free(ex.ctx, mp)
    return cc, count.reshape(n_samples, len(cc[0]))


def _oracle_pipeline(ds, n_samples):
    from oracle import c_oracle as CO
    from gmql_b200.regions import RegionDataset
    r = CO.cover_rows(ds, 0, [2], [ANYV], None, 2000)
    o = np.lexsort((r["start"], r["chrom"]))              # the oracle emits a chromosome's bin-edge pieces after the others
    r = {k: v[o] for k, v in r.items()}
    nc = len(r["start"])
    ref = RegionDataset(ds.chrom_names, r["chrom"], r["start"], r["stop"], r["strand"], np.zeros(nc, np.int32), np.array([0], np.int64))
    pidx = np.stack([np.zeros(n_samples, np.int32), np.arange(n_samples, dtype=np.int32)], axis=1)
    m = CO.map_rows(ref, ds, pidx, None, None, 50000)
    return r, m["count"].reshape(n_samples, nc)


def test_c5_production_route_against_the_oracle(ex):
    """640 samples x 8 000 peaks on hg38 shrunk twenty-fold (the same depth and tile density as C5): the exact kernels the bench
    times, row for row against the C oracle, with no path override."""
    from gmql_b200 import synth
    for k in ("GMQL_COVER_PATH", "GMQL_COVER_LOGW", "GMQL_MAP_GROUPED", "GMQL_COVER_KERNEL", "GMQL_COVER_SPECULATE"):
        assert k not in os.environ
    ds = synth.peak_samples(640, 8000, 5, with_values=(), sizes=synth.HG38_SIZES // 20)
    # the natural incidence of bin-aligned stops is kept, and a few are forced (extractor's one-base extension, :353-354)
    ds.stop[::977] = np.maximum((ds.stop[::977] // 2000) * 2000, ds.start[::977] + 1)
    dset, view, keep = _upload(ex, ds)
    assert view.chrom_bits == 8 and view.stop_is_length == 16 and not view.sample and view.sample_offsets and view.chrom_span_hint
    (got_c, got_m), kern = _profile(ex, lambda: _pipeline(ex, view, 640))
    for name in ("k_multisplit<tiled::MSP_NARROW>", "k_multisplit<tiled::MSP_FINE>", "k_cover_tile", "k_stitch_small", "k_map_index_small", "k_map_count_lean"):
        assert name in kern, (name, kern)
    for name in ("k_prepass_hist", "k_cover_prepass", "k_tile_hist", "k_widen_chrom", "k_stop_from_length", "k_expand_sample_offsets", "k_map_stream", "onesweep"):
        assert name not in kern, (name, kern)
    want_c, want_m = _oracle_pipeline(ds, 640)
    assert 100 < len(want_c["start"]) < 12000
    for a, k in zip(got_c, ("chrom", "start", "stop", "acc", "j1", "j2")):
        assert np.array_equal(a, want_c[k]), k
    assert np.array_equal(got_m, want_m)
    # the same dataset through the generic route (upload-time statistics ignored): identical rows
    os.environ["GMQL_COVER_SPECULATE"] = "0"
    try:
        (gen_c, gen_m), kern2 = _profile(ex, lambda: _pipeline(ex, view, 640))
    finally:
        os.environ.pop("GMQL_COVER_SPECULATE", None)
    assert "k_multisplit<tiled::MSP_NARROW>" not in kern2
    for a, b in zip(gen_c, got_c):
        assert np.array_equal(a, b)
    assert np.array_equal(gen_m, got_m)
    # ... and MAP through the general per-sample kernel (the lean kernel's fall-back for references with strands or duplicates)
    os.environ["GMQL_MAP_LEAN"] = "0"
    try:
        (_, grp_m), kern3 = _profile(ex, lambda: _pipeline(ex, view, 640))
    finally:
        os.environ.pop("GMQL_MAP_LEAN", None)
    assert "k_map_count_grouped<true, true>" in kern3 and "k_map_count_lean" not in kern3
    assert np.array_equal(grp_m, want_m)
    ex.lib.gmql_dataset_free(ex.ctx, dset)


@pytest.mark.parametrize("case", ["zero_length", "long_region", "strand_column", "other_bin"])
def test_upload_statistics_choose_the_route(ex, case):
    """What the upload learns decides the route without a pass over the data: a zero-length region, a region as long as a tile or a
    strand column keep a narrow dataset off the fast route (same rows as the oracle through the generic one); another
    BinSize.Cover only recounts the partition offsets."""
    from gmql_b200 import synth, _lib as L
    from oracle import c_oracle as CO
    ds = synth.peak_samples(80, 20000, 7, with_values=(), sizes=synth.HG38_SIZES // 40)
    bin_size = 2000
    if case == "zero_length":
        ds.stop[4321] = ds.start[4321]
    elif case == "long_region":
        ds.stop[99] = ds.start[99] + 65534
    elif case == "other_bin":
        bin_size = 1000
    if case == "strand_column":
        ds.strand[:] = np.random.Generator(np.random.PCG64(1)).integers(1, 3, ds.n)     # '+' / '-' only: two strand classes
    host, keep = ds.as_c(narrow=True)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))

    def run():
        a, b = np.array([2], np.int32), np.array([ANYV], np.int32)
        out = C.c_void_p()
        ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, 0, a.ctypes.data, b.ctypes.data, bin_size, (L.Agg * 1)(), 0, C.byref(out)))
        cc = _cols(ex, out, _cov_spec())
        ex.lib.gmql_result_free(ex.ctx, out)
        return cc
    got, kern = _profile(ex, run)
    assert ("k_multisplit<tiled::MSP_NARROW>" in kern) == (case == "other_bin"), kern
    want = CO.cover_rows(ds, 0, [2], [ANYV], None, bin_size)
    o = np.lexsort((want["start"], want["chrom"], want["strand"]))
    want = {k: v[o] for k, v in want.items()}
    assert len(want["start"]) > 100
    for a, k in zip(got, ("chrom", "start", "stop", "acc", "j1", "j2")):
        assert np.array_equal(a, want[k]), (case, k)
    # a second call reuses what the first one materialised (wide columns / recounted offsets): same rows
    got2 = run()
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)
    ex.lib.gmql_dataset_free(ex.ctx, dset)


def test_deferred_check_and_options(ex):
    """"defer_check": gmql_map on a COVER view returns without waiting; a broken layout promise then surfaces at the next
    synchronising call instead of in gmql_map.  Unknown options are rejected; options are per context."""
    from gmql_b200 import GMQLCudaExecutor, synth, _lib as L
    assert ex.lib.gmql_ctx_set_option(ex.ctx, b"no.such.option", b"1") < 0
    ds = synth.peak_samples(50, 4000, 3, with_values=(), sizes=synth.HG38_SIZES // 100)
    dset, view, keep = _upload(ex, ds)
    want_c, want_m = _pipeline(ex, view, 50)
    ex.set_option("defer_check", "1")
    try:
        got_c, got_m = _pipeline(ex, view, 50)
        assert np.array_equal(got_m, want_m)
        # a second context is not affected by the first one's options
        ex2 = GMQLCudaExecutor()
        d2, v2, k2 = _upload(ex2, ds)
        c2, m2 = _pipeline(ex2, v2, 50)
        assert np.array_equal(m2, want_m)
        ex2.lib.gmql_dataset_free(ex2.ctx, d2)
        ex2.close()
        # reference rows that break the promised order: reported, late
        cov = _cover(ex, view)
        cview = L.Regions()
        ex._check(ex.lib.gmql_result_as_regions(ex.ctx, cov, C.byref(cview)))
        n = int(ex.lib.gmql_result_rows(cov))
        start = np.empty(n, np.int64)
        ex._check(ex.lib.gmql_result_column_copy(ex.ctx, cov, L.COL_COV_START, start.ctypes.data))
        bad = RegionDatasetFromRuns(ds, ex, cov)[::-1]
        c_bad, kb = bad.as_c(sorted_by_position=True)
        pairs = (L.Pair * 50)()
        for b in range(50):
            pairs[b].left_sample, pairs[b].right_sample = 0, b
        mp = C.c_void_p()
        rc = ex.lib.gmql_map(ex.ctx, C.byref(c_bad), C.byref(view), pairs, 50, (L.Agg * 1)(), 0, C.byref(mp))
        assert rc == 0                                            # not reported yet
        cnt = np.empty(int(ex.lib.gmql_result_column_len(mp, L.COL_MAP_COUNT)), np.int32)
        rc = ex.lib.gmql_result_column_copy(ex.ctx, mp, L.COL_MAP_COUNT, cnt.ctypes.data)
        assert rc < 0 and b"order" in ex.lib.gmql_last_error(ex.ctx)
        ex.lib.gmql_result_free(ex.ctx, mp)
        ex.lib.gmql_result_free(ex.ctx, cov)
    finally:
        ex.set_option("defer_check", None)
    # without the option the same call fails on the spot
    cov = _cover(ex, view)
    bad = RegionDatasetFromRuns(ds, ex, cov)[::-1]
    c_bad, kb = bad.as_c(sorted_by_position=True)
    mp = C.c_void_p()
    rc = ex.lib.gmql_map(ex.ctx, C.byref(c_bad), C.byref(view), pairs, 50, (L.Agg * 1)(), 0, C.byref(mp))
    assert rc < 0 and b"order" in ex.lib.gmql_last_error(ex.ctx)
    ex.lib.gmql_result_free(ex.ctx, cov)
    ex.lib.gmql_dataset_free(ex.ctx, dset)


class RegionDatasetFromRuns:
    """COVER result rows as a host RegionDataset; [::-1] reverses the rows."""

    def __init__(self, ds, ex, cov):
        from gmql_b200 import _lib as L
        from gmql_b200.regions import RegionDataset
        chrom, start, stop = _cols(ex, cov, [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64)])
        self.names = ds.chrom_names
        self.cols = (chrom, start, stop)
        self.cls = RegionDataset

    def __getitem__(self, sl):
        chrom, start, stop = (a[sl] for a in self.cols)
        n = len(chrom)
        return self.cls(self.names, chrom, start, stop, np.zeros(n, np.int8), np.zeros(n, np.int32), np.array([0], np.int64))


@pytest.mark.parametrize("n_ref", [1, 900, 9000])
def test_lean_count_kernel_against_the_oracle(ex, n_ref):
    """k_map_count_lean directly against the C oracle (GenometricMap71.scala:81-147) on what its vector loads and 32-bit pipeline
    could get wrong: samples whose first / last rows share an aligned group of four with their neighbours (sizes 1, 2, 3, 5, empty
    samples, a last group cut by the end of the dataset), a sorted annotation whose regions OVERLAP (running-maximum walk), long
    reference regions, experiment rows beyond the last reference region of their chromosome and zero-length rows."""
    from gmql_b200 import _lib as L
    from gmql_b200.regions import RegionDataset
    from oracle import c_oracle as CO
    rng = np.random.Generator(np.random.PCG64(700 + n_ref))
    span = 2_000_000
    names = ["chr1", "chr2", "chrX", "chrY"]
    chrom = rng.integers(0, 3, n_ref).astype(np.int32)        # chrY holds no reference region
    start = rng.integers(0, span, n_ref)
    ln = rng.integers(1, 1200, n_ref)
    ln[rng.random(n_ref) < 0.03] = 300_000
    o = np.lexsort((start, chrom))
    chrom, start, ln = chrom[o], start[o], ln[o]
    ref = RegionDataset(names, chrom, start, start + ln, np.zeros(n_ref, np.int8), np.zeros(n_ref, np.int32), np.array([5], np.int64))
    sizes = [30001, 1, 20002, 0, 60003, 3, 5, 2, 0, 41]
    smp = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    n_exp = smp.shape[0]
    assert n_exp % 4 != 0
    ec = rng.integers(0, 4, n_exp).astype(np.int32)
    es = rng.integers(0, span + 500_000, n_exp)
    el = rng.integers(0, 3000, n_exp)
    el[::53] = 0
    exp = RegionDataset(names, ec, es, es + el, np.zeros(n_exp, np.int8), smp, np.arange(10, 10 + len(sizes), dtype=np.int64))
    dset, view, keep = _upload(ex, exp)
    c_ref, k1 = ref.as_c(sorted_by_position=True)
    c_ref.strand = None                                       # no strand column: every reference region is '*'
    n_pairs = len(sizes) - 1                                  # sample 8 (empty) is in no pair
    order = [b for b in range(len(sizes)) if b != 8]
    pairs = (L.Pair * n_pairs)()
    for k, b in enumerate(order):
        pairs[k].left_sample, pairs[k].right_sample = 0, b

    def run():
        out = C.c_void_p()
        ex._check(ex.lib.gmql_map(ex.ctx, C.byref(c_ref), C.byref(view), pairs, n_pairs, (L.Agg * 1)(), 0, C.byref(out)))
        (cnt,) = _cols(ex, out, [(L.COL_MAP_COUNT, np.int32)])
        ex.lib.gmql_result_free(ex.ctx, out)
        return cnt
    os.environ["GMQL_MAP_GROUPED"] = "1"
    try:
        got, kern = _profile(ex, run)
        os.environ["GMQL_MAP_LEAN"] = "0"
        got_g, kern_g = _profile(ex, run)
    finally:
        os.environ.pop("GMQL_MAP_GROUPED", None)
        os.environ.pop("GMQL_MAP_LEAN", None)
    assert "k_map_count_lean" in kern and "k_map_count_grouped" in kern_g and "k_map_count_lean" not in kern_g, (kern, kern_g)
    pidx = np.stack([np.zeros(n_pairs, np.int32), np.array(order, np.int32)], axis=1)
    want = CO.map_rows(ref, exp, pidx, None, None, 50000)["count"]
    assert want.sum() > 0
    assert np.array_equal(got, want) and np.array_equal(got_g, want)
    ex.lib.gmql_dataset_free(ex.ctx, dset)


def test_grouped_kernel_is_not_used_with_duplicate_reference_records(ex):
    """duplicate reference records pool their counts on the canonical row (GenometricMap71.scala:186-203), which
    can exceed the 16-bit counters of k_map_count_grouped: a reference with dup_of takes the streaming kernel even when the
    grouped one is forced.  Two identical whole-chromosome reference regions over a sample of 40 000 overlapping rows: the
    canonical row must read 80 000."""
    from gmql_b200 import _lib as L
    from gmql_b200.regions import RegionDataset
    names = ["chr1"]
    ref = RegionDataset(names, np.zeros(3, np.int32), np.array([0, 0, 5_000_000]), np.array([4_000_000, 4_000_000, 5_000_100]), np.zeros(3, np.int8), np.zeros(3, np.int32),
                        np.array([7], np.int64))
    n_exp = 40000
    rng = np.random.Generator(np.random.PCG64(3))
    es = rng.integers(0, 3_900_000, n_exp)
    exp = RegionDataset(names, np.zeros(n_exp, np.int32), es, es + 50, np.zeros(n_exp, np.int8), np.zeros(n_exp, np.int32), np.array([9], np.int64))
    c_ref, k1 = ref.as_c(dup_of=np.array([0, 0, 2], np.int32), sorted_by_position=True)
    c_exp, k2 = exp.as_c(narrow=True)
    pairs = (L.Pair * 1)()
    os.environ["GMQL_MAP_GROUPED"] = "1"
    try:
        def run():
            out = C.c_void_p()
            ex._check(ex.lib.gmql_map(ex.ctx, C.byref(c_ref), C.byref(c_exp), pairs, 1, (L.Agg * 1)(), 0, C.byref(out)))
            (cnt,) = _cols(ex, out, [(L.COL_MAP_COUNT, np.int32)])
            ex.lib.gmql_result_free(ex.ctx, out)
            return cnt
        cnt, kern = _profile(ex, run)
    finally:
        os.environ.pop("GMQL_MAP_GROUPED", None)
    assert "k_map_count_grouped" not in kern
    assert cnt.tolist() == [80000, -1, 0]
