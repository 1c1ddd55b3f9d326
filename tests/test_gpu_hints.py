"""Optional input promises of the ABI (include/gmql_cuda.h): chrom_span_hint and sorted_by_position.  Honoured promises
give the same result as the plain call; broken ones are errors, never wrong results."""
import ctypes as C
import random

import numpy as np
import pytest

from tests.randgen import rand_regions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _map_counts(ex, c_ref, c_exp, n_pairs_samples):
    from gmql_b200 import _lib as L
    pairs = (L.Pair * n_pairs_samples)()
    for b in range(n_pairs_samples):
        pairs[b].left_sample, pairs[b].right_sample = 0, b
    out = C.c_void_p()
    rc = ex.lib.gmql_map(ex.ctx, C.byref(c_ref), C.byref(c_exp), pairs, n_pairs_samples, (L.Agg * 1)(), 0, C.byref(out))
    if rc != 0:
        return rc, ex.lib.gmql_last_error(ex.ctx)
    n = int(ex.lib.gmql_result_column_len(out, L.COL_MAP_COUNT))
    a = np.empty(n, np.int32)
    ex._check(ex.lib.gmql_result_column_copy(ex.ctx, out, L.COL_MAP_COUNT, a.ctypes.data))
    ex.lib.gmql_result_free(ex.ctx, out)
    return 0, a


def test_hints_honoured_and_violations_rejected(ex):
    from gmql_b200 import RegionDataset
    rng = random.Random(9)
    ref = sorted(rand_regions(rng, 300, [1], chroms=("chr1", "chr2", "chr3"), n_vals=0, span=200000, max_len=3000), key=lambda r: (r[1], r[2]))
    exp = rand_regions(rng, 3000, [10, 11], chroms=("chr1", "chr2", "chr3"), n_vals=0, span=200000, max_len=3000)
    dr, de = RegionDataset.from_records(ref, [1]), RegionDataset.from_records(exp, [10, 11])
    c_exp, k2 = de.as_c()
    c_plain, k1 = dr.as_c()
    rc, want = _map_counts(ex, c_plain, c_exp, 2)
    assert rc == 0 and want.sum() > 0
    spans = [int(dr.stop[dr.chrom == c].max()) for c in range(3)]
    for kw in (dict(sorted_by_position=True), dict(chrom_span_hint=[s + 5 for s in spans]), dict(sorted_by_position=True, chrom_span_hint=[250000] * 3)):
        c_ref, k = dr.as_c(**kw)
        rc, got = _map_counts(ex, c_ref, c_exp, 2)
        assert rc == 0 and np.array_equal(got, want), kw
    # a reference region beyond its bound
    c_ref, k = dr.as_c(chrom_span_hint=[spans[0] - 1, spans[1], spans[2]])
    rc, msg = _map_counts(ex, c_ref, c_exp, 2)
    assert rc < 0 and b"chrom_span_hint" in msg
    # rows not in the promised order
    shuffled = RegionDataset.from_records(ref[::-1], [1])
    c_ref, k = shuffled.as_c(sorted_by_position=True)
    rc, msg = _map_counts(ex, c_ref, c_exp, 2)
    assert rc < 0 and b"order" in msg
    # the same data without the promise is fine
    c_ref, k = shuffled.as_c()
    rc, got = _map_counts(ex, c_ref, c_exp, 2)
    assert rc == 0 and int(got.sum()) == int(want.sum())


def test_cover_result_view_carries_the_hints(ex):
    """C = COVER(2,ANY) E; M = MAP() C E on the device-resident view equals MAP on the downloaded COVER rows."""
    import gmql_b200 as G
    from gmql_b200 import RegionDataset, _lib as L
    rng = random.Random(10)
    exp = rand_regions(rng, 2500, [10, 11, 12], chroms=("chr1", "chr2"), strands="*", n_vals=0, span=60000, max_len=900)
    de = RegionDataset.from_records(exp, [10, 11, 12])
    res, _ = ex.cover_raw("COVER", G.N(2), G.ANY(), [], None, de)
    view = L.Regions()
    ex._check(ex.lib.gmql_result_as_regions(ex.ctx, res.h, C.byref(view)))
    assert view.sorted_by_position == 1 and view.chrom_span_hint
    c_exp, k2 = de.as_c()
    rc, got = _map_counts(ex, view, c_exp, 3)
    assert rc == 0
    cov = RegionDataset(de.chrom_names, res.column(L.COL_COV_CHROM, np.int32), res.column(L.COL_COV_START, np.int64), res.column(L.COL_COV_STOP, np.int64),
                        np.zeros(res.rows(), np.int8), np.zeros(res.rows(), np.int32), np.array([0], np.int64))
    c_cov, k3 = cov.as_c()
    rc, want = _map_counts(ex, c_cov, c_exp, 3)
    res.free()
    assert rc == 0 and len(want) > 0 and np.array_equal(got, want)


@pytest.mark.parametrize("n_ref", [1, 7, 1025, 50000])
def test_one_launch_index_equals_separate_index_kernels(ex, n_ref):
    """A pre-sorted single-sample reference is indexed by one cluster kernel (k_map_index_small); same counts as the
    separate index kernels and as the un-hinted call.  Long regions early in a block range make the running maximum
    cross the cluster's block boundaries."""
    import os
    from gmql_b200 import RegionDataset
    rng = np.random.Generator(np.random.PCG64(n_ref))
    span = 5_000_000
    chrom = np.sort(rng.integers(0, 3, n_ref)).astype(np.int32)
    start = rng.integers(0, span, n_ref)
    ln = rng.integers(1, 400, n_ref)
    ln[rng.random(n_ref) < 0.01] = 2_000_000                      # a few regions covering thousands of their successors
    o = np.lexsort((start, chrom))
    chrom, start, ln = chrom[o], start[o], ln[o]
    names = ["chr1", "chr2", "chr3"]
    dr = RegionDataset(names, chrom, start, start + ln, rng.integers(0, 3, n_ref).astype(np.int8), np.zeros(n_ref, np.int32), np.array([5], np.int64))
    n_exp = 200000
    ec = rng.integers(0, 3, n_exp).astype(np.int32)
    es = rng.integers(0, span + 2_000_000, n_exp)
    de = RegionDataset(names, ec, es, es + rng.integers(0, 3000, n_exp), rng.integers(0, 3, n_exp).astype(np.int8), rng.integers(0, 2, n_exp).astype(np.int32), np.array([10, 11], np.int64))
    c_exp, k2 = de.as_c()
    c_plain, k0 = dr.as_c()
    rc, want = _map_counts(ex, c_plain, c_exp, 2)
    assert rc == 0 and want.sum() > 0
    c_ref, k1 = dr.as_c(sorted_by_position=True)
    got = {}
    for mode in ("fused", "split"):
        if mode == "split":
            os.environ["GMQL_MAP_INDEX"] = "split"
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
        try:
            rc, got[mode] = _map_counts(ex, c_ref, c_exp, 2)
        finally:
            os.environ.pop("GMQL_MAP_INDEX", None)
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
        assert rc == 0
        assert ("k_map_index_small" in buf.value.decode()) == (mode == "fused")
    assert np.array_equal(got["fused"], want) and np.array_equal(got["split"], want)


def _profile(ex, fn):
    ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
    try:
        out = fn()
    finally:
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
    return out, buf.value.decode()


@pytest.mark.parametrize("n_ref,presorted", [(1, False), (700, False), (9000, True), (14000, True)])
def test_grouped_count_kernel_equals_streaming_kernel(ex, n_ref, presorted):
    """COUNT onto a small single-sample reference with the experiment rows grouped by sample runs in k_map_count_grouped
    (everything random-access in shared memory); same counts as the streaming kernel.  14 000 reference regions do not
    fit shared memory: the streaming kernel runs.  Sample 2 is in no pair, sample 3 is empty."""
    import os
    from gmql_b200 import RegionDataset, _lib as L
    rng = np.random.Generator(np.random.PCG64(100 + n_ref))
    span = 3_000_000
    names = ["chr1", "chr2", "chrX"]
    chrom = rng.integers(0, 3, n_ref).astype(np.int32)
    start = rng.integers(0, span, n_ref)
    ln = rng.integers(0, 900, n_ref)
    ln[rng.random(n_ref) < 0.02] = 500_000
    if presorted:
        o = np.lexsort((start, chrom))
        chrom, start, ln = chrom[o], start[o], ln[o]
    dr = RegionDataset(names, chrom, start, start + ln, rng.integers(0, 3, n_ref).astype(np.int8), np.zeros(n_ref, np.int32), np.array([5], np.int64))
    sizes = [30000, 1, 20000, 0, 60000]
    smp = np.repeat(np.arange(5), sizes).astype(np.int32)
    n_exp = smp.shape[0]
    ec = rng.integers(0, 3, n_exp).astype(np.int32)
    es = rng.integers(0, span + 600_000, n_exp)
    de = RegionDataset(names, ec, es, es + rng.integers(0, 2500, n_exp), rng.integers(0, 3, n_exp).astype(np.int8), smp, np.arange(10, 15, dtype=np.int64))
    pairs = (L.Pair * 4)()
    for k, b in enumerate((0, 1, 3, 4)):
        pairs[k].left_sample, pairs[k].right_sample = 0, b

    def counts(c_ref, c_exp):
        out = C.c_void_p()
        rc = ex.lib.gmql_map(ex.ctx, C.byref(c_ref), C.byref(c_exp), pairs, 4, (L.Agg * 1)(), 0, C.byref(out))
        if rc != 0:
            return rc, ex.lib.gmql_last_error(ex.ctx)
        a = np.empty(int(ex.lib.gmql_result_column_len(out, L.COL_MAP_COUNT)), np.int32)
        ex._check(ex.lib.gmql_result_column_copy(ex.ctx, out, L.COL_MAP_COUNT, a.ctypes.data))
        ex.lib.gmql_result_free(ex.ctx, out)
        return 0, a

    c_ref, k0 = dr.as_c(sorted_by_position=presorted)
    c_plain, k1 = de.as_c()
    (rc, want), kern = _profile(ex, lambda: counts(c_ref, c_plain))
    assert rc == 0 and want.shape[0] == 4 * n_ref and "k_map_count_grouped" not in kern
    if n_ref > 1:
        assert want.sum() > 0
    c_grp, k2 = de.as_c(narrow=True)
    assert c_grp.sample_offsets and not c_grp.sample
    # five samples would not fill the GPU: the kernel is chosen from 4 x SMs samples on; "1" forces it
    (rc, got), kern = _profile(ex, lambda: counts(c_ref, c_grp))
    assert rc == 0 and np.array_equal(got, want) and "k_map_count_grouped" not in kern
    os.environ["GMQL_MAP_GROUPED"] = "1"
    try:
        (rc, got), kern = _profile(ex, lambda: counts(c_ref, c_grp))
        assert rc == 0 and np.array_equal(got, want)
        assert ("k_map_count_grouped" in kern) == (n_ref <= 12000), kern
        # a sample column that contradicts the offsets is an error, never a wrong result
        c_bad, k3 = de.as_c()
        off = np.array([0, 30000, 30001, 50001, 50001, n_exp], np.int64)
        off_bad = off.copy()
        off_bad[1] = 29990
        c_bad.sample_offsets = off_bad.ctypes.data
        rc, msg = counts(c_ref, c_bad)
        if n_ref <= 12000:
            assert rc < 0 and b"sample_offsets" in msg
        c_bad.sample_offsets = off.ctypes.data             # consistent offsets next to the column: fine
        rc, got = counts(c_ref, c_bad)
        assert rc == 0 and np.array_equal(got, want)
    finally:
        os.environ.pop("GMQL_MAP_GROUPED", None)


def _cover_cols(ex, c_ds, flag, mn, mx):
    from gmql_b200 import _lib as L
    mn_a, mx_a = np.array([mn], np.int32), np.array([mx], np.int32)
    out = C.c_void_p()
    rc = ex.lib.gmql_cover(ex.ctx, C.byref(c_ds), None, 1, flag, mn_a.ctypes.data, mx_a.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out))
    if rc != 0:
        return rc, ex.lib.gmql_last_error(ex.ctx)
    cols = []
    for col, dt in ((L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32),
                    (L.COL_COV_STRAND, np.int8), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)):
        a = np.empty(int(ex.lib.gmql_result_column_len(out, col)), dt)
        if a.shape[0]:
            ex._check(ex.lib.gmql_result_column_copy(ex.ctx, out, col, a.ctypes.data))
        cols.append(a)
    ex.lib.gmql_result_free(ex.ctx, out)
    return 0, cols


@pytest.mark.parametrize("case", ["eligible", "zero_length", "long_region", "beyond_hint", "bad_row"])
def test_speculative_cover_with_span_hint(ex, case):
    """With chrom_span_hint (and no strand column) gmql_cover queues the whole tile pipeline behind a fused prepass + histogram
    kernel without waiting for the prepass flags.  Eligible input: same rows as the ordinary call, and k_prepass_hist ran.
    Inputs that turn out not to be eligible (a zero-length region, a region as long as a tile, a region beyond its hint) are
    caught by the device-side guard and recomputed the ordinary way: same rows as without the hint.  Invalid rows are errors."""
    import os
    from gmql_b200 import synth, _lib as L
    ds = synth.peak_samples(60, 20000, 9, with_values=())
    spans = [int(ds.stop[ds.chrom == c].max(initial=0)) for c in range(len(ds.chrom_names))]
    if case == "zero_length":
        ds.stop[1234] = ds.start[1234]
    elif case == "long_region":
        ds.stop[777] = ds.start[777] + 70000
        spans = [int(ds.stop[ds.chrom == c].max(initial=0)) for c in range(len(ds.chrom_names))]
    elif case == "bad_row":
        ds.start[55] = ds.stop[55] + 10
    hint = list(spans)
    if case == "beyond_hint":
        hint[0] -= 5
    os.environ["GMQL_COVER_PATH"] = "tiled"
    try:
        for flag, (mn, mx) in ((0, (2, 2 ** 31 - 1)), (1, (1, 3))):
            c_plain, k0 = ds.as_c()
            c_plain.strand = None
            rc0, want = _cover_cols(ex, c_plain, flag, mn, mx)
            c_hint, k1 = ds.as_c(chrom_span_hint=hint)
            c_hint.strand = None
            (rc, got), kern = _profile(ex, lambda: _cover_cols(ex, c_hint, flag, mn, mx))
            if case == "bad_row":
                assert rc0 < 0 and rc < 0
                continue
            assert rc0 == 0 and rc == 0 and "k_prepass_hist" in kern
            assert ("k_cover_prepass" in kern) == (case != "eligible"), kern      # the ordinary prepass only runs after a tripped guard
            assert len(want[0]) > 100
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (case, flag)
            os.environ["GMQL_COVER_SPECULATE"] = "0"
            try:
                (rc, got), kern = _profile(ex, lambda: _cover_cols(ex, c_hint, flag, mn, mx))
            finally:
                os.environ.pop("GMQL_COVER_SPECULATE", None)
            assert rc == 0 and "k_prepass_hist" not in kern
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (case, flag)
    finally:
        os.environ.pop("GMQL_COVER_PATH", None)
