"""Full-size runs of BASELINE.json configs[0..3] (C1..C4) through the C ABI on one GPU.

Not the contract bench (that is bench.py on C5): these are the parity-test shapes at their stated sizes, timed once each
(device-resident inputs, CUDA events via gmql_last_device_ms, best of --reps) with size-independent checks of the results.
Prints one JSON line per case.  usage: python bench_configs.py [--cases c1,c2,c3,c4] [--reps 3]"""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np


def upload(ex, L, ds, dup_of=None):
    host, keep = ds.as_c(dup_of=dup_of, narrow=True)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    return dset, view, keep


PROFILE = False


def timed(ex, fn, reps):
    best, out = None, None
    each = []
    for _ in range(reps):
        if out is not None:
            ex.lib.gmql_result_free(ex.ctx, out)
        out = fn()
        ms = ex.last_device_ms()
        each.append(round(ms, 2))
        best = ms if best is None else min(best, ms)
    sys.stderr.write("reps (device ms): %s\n" % each)
    if PROFILE:      # one more run with CUDA events around every launch: per-kernel totals on stderr
        ex.lib.gmql_result_free(ex.ctx, out)
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
        out = fn()
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
        sys.stderr.write("--- kernels\n" + "\n".join(buf.value.decode().splitlines()[:10]) + "\n")
    return best, out


def col(ex, res, c, dt):
    n = int(ex.lib.gmql_result_column_len(res, c))
    a = np.empty(n, dtype=dt)
    if n:
        ex._check(ex.lib.gmql_result_column_copy(ex.ctx, res, c, a.ctypes.data))
    return a


def emit(case, n_in, ms, **kw):
    if PROFILE:
        sys.stderr.write("^^^ %s: %.3f ms\n" % (case, ms))
    print(json.dumps(dict(case=case, input_regions=int(n_in), device_ms=round(ms, 3), regions_per_s=round(n_in / (ms * 1e-3), 1), **kw)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c1,c2,c3,c4")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="scales the sample counts (quick runs)")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    global PROFILE
    PROFILE = args.profile
    cases = args.cases.split(",")
    from gmql_b200 import GMQLCudaExecutor, synth, _lib as L
    ex = GMQLCudaExecutor()
    lib, ctx = ex.lib, ex.ctx
    no_aggs = (L.Agg * 1)()

    if "parse" in cases:
        # narrowPeak text of 2e6 regions (10 columns) -> device columns; host buffer in, H2D inside the timed region
        rng = np.random.Generator(np.random.PCG64(11))
        n = int(2_000_000 * args.scale)
        chrom = rng.integers(1, 23, n)
        start = rng.integers(0, 10 ** 8, n)
        ln = rng.integers(50, 5000, n)
        sc = rng.integers(0, 1000, n)
        sv, pv, qv = rng.lognormal(1.5, 0.8, n), rng.uniform(0, 300, n), rng.uniform(0, 300, n)
        pk = rng.integers(0, 50, n)
        text = "".join("chr%d\t%d\t%d\tp%d\t%d\t.\t%r\t%r\t%r\t%d\n" % (chrom[i], start[i], start[i] + ln[i], i, sc[i], float(sv[i]), float(pv[i]), float(qv[i]), pk[i]) for i in range(n)).encode()
        names = ["chr%d" % c for c in range(1, 23)]
        best = None
        for _ in range(args.reps):
            t0 = time.perf_counter()
            got = ex.parse_text(text, "NarrowPeakParser", chrom_names=names)
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        dev_ms = ex.last_device_ms()
        ok = bool(not got["status"].any() and np.array_equal(got["start"], start) and np.array_equal(got["values"][1][0], sv))
        print(json.dumps(dict(case="PARSE narrowPeak text -> columns", lines=n, text_mb=round(len(text) / 1e6, 1), device_ms=round(dev_ms, 3),
                              wall_ms_with_copies=round(best, 1), gb_per_s_device=round(len(text) / dev_ms / 1e6, 2), bit_exact=ok)), flush=True)

    if "c1" in cases:
        ref, exp = synth.c1_datasets()
        dr, vr, k1 = upload(ex, L, ref, dup_of=None)
        de, ve, k2 = upload(ex, L, exp)
        pairs = (L.Pair * exp.sample_ids.shape[0])()
        for b in range(exp.sample_ids.shape[0]):
            pairs[b].left_sample, pairs[b].right_sample = 0, b

        def run():
            out = C.c_void_p()
            ex._check(lib.gmql_map(ctx, C.byref(vr), C.byref(ve), pairs, len(pairs), no_aggs, 0, C.byref(out)))
            return out
        ms, res = timed(ex, run, args.reps)
        cnt = col(ex, res, L.COL_MAP_COUNT, np.int32)
        emit("C1 MAP COUNT 10k ref x 10 samples", ref.n + exp.n, ms, rows=int(cnt.shape[0]), count_sum=int(cnt.sum()))
        lib.gmql_result_free(ctx, res)
        lib.gmql_dataset_free(ctx, dr); lib.gmql_dataset_free(ctx, de)

    if "c2" in cases:
        ns = max(2, int(1000 * args.scale))
        t0 = time.time()
        ds = synth.peak_samples(ns, 50000, 2, with_values=("signalValue",))
        d, v, k = upload(ex, L, ds)
        mn, mx = np.array([2], np.int32), np.array([2 ** 31 - 1], np.int32)
        sys.stderr.write("c2 data ready in %.1f s\n" % (time.time() - t0))
        agg_sum = (L.Agg * 1)()
        agg_sum[0].kind, agg_sum[0].col = 1, 0
        for flag, name in ((0, "COVER"), (1, "FLAT"), (2, "SUMMIT"), (3, "HISTOGRAM")):
            for (aggs, na, tag) in ((no_aggs, 0, ""), (agg_sum, 1, " + SUM(signalValue)")):
                if na and flag in (2, 3):
                    continue

                def run():
                    out = C.c_void_p()
                    ex._check(lib.gmql_cover(ctx, C.byref(v), None, 1, flag, mn.ctypes.data, mx.ctypes.data, 2000, aggs, na, C.byref(out)))
                    return out
                ms, res = timed(ex, run, args.reps)
                start, stop = col(ex, res, L.COL_COV_START, np.int64), col(ex, res, L.COL_COV_STOP, np.int64)
                acc = col(ex, res, L.COL_COV_ACC, np.int32)
                ok = bool(np.all(stop >= start)) and (len(acc) == 0 or int(acc.min()) >= (1 if flag == 2 else 2))
                emit("C2 %s(2,ANY)%s %d samples x 50k" % (name, tag, ns), ds.n, ms, out_regions=int(start.shape[0]), covered_bp=int((stop - start).sum()), sane=ok)
                lib.gmql_result_free(ctx, res)
        lib.gmql_dataset_free(ctx, d)

    if "c3" in cases:
        nr, ne = max(2, int(100 * args.scale)), max(2, int(1000 * args.scale))
        ref = synth.promoter_samples(nr, 1000, 3)
        exp = synth.peak_samples(ne, 50000, 3, with_values=("signalValue",))
        dr, vr, k1 = upload(ex, L, ref)
        de, ve, k2 = upload(ex, L, exp)
        pairs = (L.Pair * (nr * ne))()
        for a in range(nr):
            for b in range(ne):
                pairs[a * ne + b].left_sample, pairs[a * ne + b].right_sample = a, b
        aggs = (L.Agg * 3)()
        for i, kind in enumerate((1, 4, 3)):         # SUM, AVG, MAX
            aggs[i].kind, aggs[i].col = kind, 0

        def run():
            out = C.c_void_p()
            ex._check(lib.gmql_map(ctx, C.byref(vr), C.byref(ve), pairs, len(pairs), aggs, 3, C.byref(out)))
            return out
        ms, res = timed(ex, run, args.reps)
        cnt = col(ex, res, L.COL_MAP_COUNT, np.int32)
        s = col(ex, res, L.COL_AGG_VALUE + 0, np.float64)
        okv = col(ex, res, L.COL_AGG_VALID + 0, np.uint8)
        emit("C3 MAP SUM/AVG/MAX %d ref x %d exp samples" % (nr, ne), ref.n + exp.n, ms, rows=int(cnt.shape[0]), count_sum=int(cnt.astype(np.int64).sum()),
             sum_of_sums=float(s[okv != 0].sum()), sane=bool(np.array_equal(okv != 0, cnt > 0)))
        lib.gmql_result_free(ctx, res)
        lib.gmql_dataset_free(ctx, dr); lib.gmql_dataset_free(ctx, de)

    if "c4" in cases:
        nl = nrg = max(2, int(200 * args.scale))
        left = synth.promoter_samples(nl, 5000, 4)
        right = synth.peak_samples(nrg, 50000, 4, with_values=())
        dl, vl, k1 = upload(ex, L, left)
        dr, vr, k2 = upload(ex, L, right)
        pairs = (L.Pair * (nl * nrg))()
        for a in range(nl):
            for b in range(nrg):
                pairs[a * nrg + b].left_sample, pairs[a * nrg + b].right_sample = a, b
        # atoms: DLE(10000) = DISTLESS 10001 (GmqlParsers: DLE(n) -> DistLess(n+1)); MD(1); DOWN
        variants = (("DLE(10000) CAT", [(0, 10001)], 4), ("MD(1),DOWN RIGHT", [(2, 1), (4, 0)], 2), ("DLE(10000),MD(1),DOWN RIGHT", [(0, 10001), (2, 1), (4, 0)], 2))
        for name, atoms, builder in variants:
            at = (L.Atom * len(atoms))()
            for i, (kd, arg) in enumerate(atoms):
                at[i].kind, at[i].arg = kd, arg

            def run():
                out = C.c_void_p()
                ex._check(lib.gmql_join(ctx, C.byref(vl), C.byref(vr), pairs, len(pairs), at, len(atoms), builder | 0x100, 5000, 100000, None, None, C.byref(out)))
                return out
            ms, res = timed(ex, run, args.reps)
            emit("C4 JOIN %s %d x %d samples" % (name, nl, nrg), left.n + right.n, ms, out_pairs=int(lib.gmql_result_rows(res)))
            lib.gmql_result_free(ctx, res)
        lib.gmql_dataset_free(ctx, dl); lib.gmql_dataset_free(ctx, dr)
    ex.close()


if __name__ == "__main__":
    main()
