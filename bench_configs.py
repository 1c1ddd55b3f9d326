"""bench.py --workload c2 | c3 | c4: BASELINE.json configs[1..3] at their full sizes on ONE GPU, in the
same contract form as the default C5 line (`value` device-resident, `e2e` from pinned host buffers with every result column
copied back, `roofline` for the dominant kernel and for the whole step against SURVEY §8(d)'s algorithmic bytes, `cpu_baseline`
= the oracle restatement on a bounded sample).  A step runs every operator of the config once:

    c2   COVER(2,ANY), FLAT(2,ANY), SUMMIT(2,ANY), HISTOGRAM(2,ANY) over 1 000 narrowPeak samples x 50 000 regions
    c3   MAP(SUM, AVG, MAX of signalValue) of 100 reference samples x 1 000 experiment samples (1e8 output rows)
    c4   JOIN DLE(10000) CAT;  JOIN MD(1), DOWN RIGHT;  JOIN DLE(10000), MD(1), DOWN RIGHT between 200 x 200 samples

`input regions` of a step = the rows of all region inputs of every operator it runs.  The driver only runs the
default workload; the lines of these are committed under profiles/.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ANYV = 2 ** 31 - 1


def _upload(ex, L, host):
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    return dset, view


class Workload:
    name = ""
    config = {}

    def hosts(self):          # [(gmql_regions over pinned host memory, keepalive, h2d bytes)]
        raise NotImplementedError

    def ops(self, views):     # list of (label, callable -> result handle, fetch spec [(col id, dtype)], input regions)
        raise NotImplementedError

    def alg_bytes(self, label, n_out):
        raise NotImplementedError

    def cpu(self):            # cpu_baseline dict
        raise NotImplementedError


def run(args, stdout):
    import torch
    import bench as B
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: this back end has no CPU fallback"
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        if int(os.environ.get("RANK", 0)) == 0:
            print(json.dumps({"unavailable": "--workload %s is a single-GPU line (the scaling workload is the default C5)" % args.workload}), file=stdout, flush=True)
        return
    torch.cuda.set_device(0)
    from gmql_b200 import GMQLCudaExecutor, _lib as L
    from gmql_b200.testing import install_env_hooks
    numa = B.bind_to_gpu_numa_node(torch, 0)
    stream = torch.cuda.current_stream()
    ex = GMQLCudaExecutor(device=0, stream=stream.cuda_stream)
    install_env_hooks(ex.lib)
    lib, ctx = ex.lib, ex.ctx
    wl = {"c2": C2, "c3": C3, "c4": C4}[args.workload](args, ex, L, torch)
    if args.impl == "reference":
        B.restore_all_cpus()
        cpu = wl.cpu()
        out = {"impl": "reference", "metric": B.METRIC, "value": cpu["value"], "unit": "input regions/s", "n_gpus": 1, "steps": 1, "warmup": 0,
               "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64 coordinates / f64 values", "data": "synthetic",
               "config": wl.config, "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "input regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), file=stdout, flush=True)
        return
    hosts = wl.hosts()
    h2d_bytes = sum(h[2] for h in hosts)
    ups = [_upload(ex, L, h[0]) for h in hosts]
    views = [u[1] for u in ups]
    ops = wl.ops(views)
    n_in_step = sum(o[3] for o in ops)

    def step(collect=None):
        for label, fn, spec, _n in ops:
            res = fn()
            if collect is not None:
                collect[label] = (int(lib.gmql_result_rows(res)), float(lib.gmql_last_device_ms(ctx)))
            lib.gmql_result_free(ctx, res)

    info = {}
    for _ in range(args.warmup):
        step(info)
    sampler = B.ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    l0 = ex.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    per_op = {o[0]: [] for o in ops}
    for k in range(args.steps):
        got = {}
        step(got)
        for lab, (_r, ms) in got.items():
            per_op[lab].append(ms)
        evs[k + 1].record(stream)
    torch.cuda.synchronize()
    launches = ex.launch_count() - l0
    clocks = sampler.stop()
    ms_dev = evs[0].elapsed_time(evs[-1]) / args.steps
    each = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps))
    prof = B.profile_kernels(ex, lambda: step())
    for d, _v in ups:
        lib.gmql_dataset_free(ctx, d)

    # ---- end to end: pinned host buffers in, every result column out ---------------------------------------------------
    outs = {}
    d2h = 0
    for label, fn, spec, _n in ops:
        rows = info[label][0]
        outs[label] = [torch.empty(max(rows, 1) * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory() for (_c, dt) in spec]
        d2h += sum(rows * np.dtype(dt).itemsize for (_c, dt) in spec)

    def e2e_step():
        ups2 = [_upload(ex, L, h[0]) for h in hosts]
        ops2 = wl.ops([u[1] for u in ups2])
        for (label, fn, spec, _n) in ops2:
            res = fn()
            for (cid, dt), buf in zip(spec, outs[label]):
                if int(lib.gmql_result_column_len(res, cid)) > 0:
                    ex._check(lib.gmql_result_column_copy(ctx, res, cid, buf.data_ptr()))
            lib.gmql_result_free(ctx, res)
        for d, _v in ups2:
            lib.gmql_dataset_free(ctx, d)

    e2e_step()
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps

    peak, peak_kind = B.measured_peak()
    top = prof[0] if prof else ("none", 1, 0.0)
    top_ms = top[2] / max(1, top[1])
    step_bytes = sum(wl.alg_bytes(label, info[label][0]) for (label, _f, _s, _n) in ops)
    per = {lab: {"device_ms": round(float(np.median(v)), 3), "out_rows": info[lab][0], "algorithmic_bytes": wl.alg_bytes(lab, info[lab][0]),
                 "roofline_frac": wl.alg_bytes(lab, info[lab][0]) / (float(np.median(v)) * 1e-3) / 1e9 / peak} for lab, v in per_op.items()}
    top_alg = max((wl.alg_bytes(lab, info[lab][0]) for lab in per_op), default=0) if len(ops) > 1 else step_bytes
    out = {
        "metric": B.METRIC, "value": n_in_step / (ms_dev * 1e-3), "unit": "input regions/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "ms_per_step_median": each[len(each) // 2], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 coordinates / f64 values", "data": "synthetic", "config": wl.config,
        "operators": per,
        "e2e": {"value": n_in_step / (ms_e2e * 1e-3), "unit": "input regions/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "host_binding": numa},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": top[0], "launches_per_step": top[1], "ms_per_launch": top_ms, "achieved": top_alg / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0,
                     "peak": peak, "unit": "GB/s", "frac": (top_alg / (top_ms * 1e-3) / 1e9 / peak) if top_ms > 0 else 0.0, "traffic": None, "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": top_alg, "note": "algorithmic bytes of the operator this kernel belongs to (SURVEY 8d) / the kernel's own time"},
        "roofline_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (ms_dev * 1e-3) / 1e9, "frac_per_gpu": step_bytes / (ms_dev * 1e-3) / 1e9 / peak},
        "kernels": [{"name": n, "launches": c, "ms": round(m, 4)} for (n, c, m) in prof[:14]],
    }
    if not args.no_cpu:
        B.restore_all_cpus()
        out["cpu_baseline"] = wl.cpu()
    print(json.dumps(out), file=stdout, flush=True)
    ex.close()


# ---------------------------------------------------------------------------------------------------------------------
class C2(Workload):
    def __init__(self, args, ex, L, torch):
        from gmql_b200 import synth
        self.args, self.ex, self.L, self.torch = args, ex, L, torch
        self.ns = max(2, int(1000 * args.scale))
        self.config = {"workload": "C2: COVER(2,ANY), FLAT(2,ANY), SUMMIT(2,ANY), HISTOGRAM(2,ANY)  (BASELINE.json configs[1])", "samples": self.ns, "regions_per_sample": 50000,
                       "input_regions": 4 * self.ns * 50000, "sharding": "single GPU",
                       "l2": "inputs resident in HBM (0.35 GB), far larger than the 126 MB L2; nothing is reused between operators"}
        self.ds = None
        self.synth = synth

    def data(self):
        if self.ds is None:
            self.ds = self.synth.peak_samples(self.ns, 50000, 2, with_values=())
        return self.ds

    def hosts(self):
        import bench as B
        return [B.narrow_host_struct(self.L, self.torch, self.data())]

    def ops(self, views):
        L, ex = self.L, self.ex
        v = views[0]
        mn, mx = np.array([2], np.int32), np.array([ANYV], np.int32)
        self._keep = (mn, mx)
        no_aggs = (L.Agg * 1)()
        spec = [(L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64), (L.COL_COV_ACC, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)]

        def mk(flag):
            def run():
                out = C.c_void_p()
                ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(v), None, 1, flag, mn.ctypes.data, mx.ctypes.data, 2000, no_aggs, 0, C.byref(out)))
                return out
            return run
        return [(name, mk(flag), spec, self.data().n) for flag, name in ((0, "COVER"), (1, "FLAT"), (2, "SUMMIT"), (3, "HISTOGRAM"))]

    def alg_bytes(self, label, n_out):
        return 9 * self.data().n + 28 * n_out

    def cpu(self):
        from oracle import c_oracle as CO
        import bench as B
        cores = CO.set_threads(B.host_cores())
        order = np.argsort(self.synth.HG38_SIZES)
        chroms = sorted(int(c) for c in order[:4])
        ds = self.synth.peak_samples(self.ns, 50000, 2, with_values=(), chrom_subset=chroms)
        t0 = time.perf_counter()
        rows = {}
        for flag, name in ((0, "COVER"), (1, "FLAT"), (2, "SUMMIT"), (3, "HISTOGRAM")):
            t1 = time.perf_counter()
            r = CO.cover_rows(ds, flag, [2], [ANYV], None, 2000)
            rows[name] = {"rows": int(len(r["start"])), "s": round(time.perf_counter() - t1, 2)}
        dt = time.perf_counter() - t0
        return {"value": 4 * ds.n / dt, "unit": "input regions/s", "cores": cores, "kind": "port",
                "sample": "C2 sample: %d samples x 50000 regions restricted to the 4 smallest hg38 chromosomes (%d regions), all four flags, %.1f s; C/OpenMP restatement of the bin-based algorithm, not GMQL-Spark" % (self.ns, ds.n, dt),
                "operators": rows}


class C3(Workload):
    def __init__(self, args, ex, L, torch):
        from gmql_b200 import synth
        self.args, self.ex, self.L, self.torch, self.synth = args, ex, L, torch, synth
        self.nr, self.ne = max(2, int(100 * args.scale)), max(2, int(1000 * args.scale))
        self.config = {"workload": "C3: MAP(SUM, AVG, MAX of signalValue) R E  (BASELINE.json configs[2])", "ref_samples": self.nr, "ref_regions_per_sample": 1000,
                       "exp_samples": self.ne, "exp_regions_per_sample": 50000, "input_regions": self.nr * 1000 + self.ne * 50000, "sharding": "single GPU",
                       "l2": "inputs resident in HBM (0.75 GB), outputs 3.1 GB: far larger than the 126 MB L2"}
        self.ref = self.exp = None

    def data(self):
        if self.ref is None:
            self.ref = self.synth.promoter_samples(self.nr, 1000, 3)
            self.exp = self.synth.peak_samples(self.ne, 50000, 3, with_values=("signalValue",))
        return self.ref, self.exp

    def hosts(self):
        import bench as B
        ref, exp = self.data()
        return [B.narrow_host_struct(self.L, self.torch, ref, keep_strand=True), B.narrow_host_struct(self.L, self.torch, exp, with_cols=True)]

    def ops(self, views):
        L, ex = self.L, self.ex
        vr, ve = views
        nr, ne = self.nr, self.ne
        pairs = (L.Pair * (nr * ne))()
        for a in range(nr):
            for b in range(ne):
                pairs[a * ne + b].left_sample, pairs[a * ne + b].right_sample = a, b
        aggs = (L.Agg * 3)()
        for i, kind in enumerate((1, 4, 3)):         # SUM, AVG, MAX
            aggs[i].kind, aggs[i].col = kind, 0
        self._keep = (pairs, aggs)
        spec = [(L.COL_MAP_COUNT, np.int32)] + [(L.COL_AGG_VALUE + k, np.float64) for k in range(3)] + [(L.COL_AGG_VALID + k, np.uint8) for k in range(3)]

        def run():
            out = C.c_void_p()
            ex._check(ex.lib.gmql_map(ex.ctx, C.byref(vr), C.byref(ve), pairs, nr * ne, aggs, 3, C.byref(out)))
            return out
        ref, exp = self.data()
        return [("MAP", run, spec, ref.n + exp.n)]

    def alg_bytes(self, label, n_out):
        ref, exp = self.data()
        return (9 + 8) * exp.n + 9 * ref.n + (4 + 8 * 3) * n_out

    def cpu(self):
        from oracle import c_oracle as CO
        import bench as B
        cores = CO.set_threads(B.host_cores())
        nr, ne = max(2, self.nr // 10), max(2, self.ne // 10)
        ref = self.synth.promoter_samples(nr, 1000, 3)
        exp = self.synth.peak_samples(ne, 50000, 3, with_values=("signalValue",))
        pidx = np.array([(a, b) for a in range(nr) for b in range(ne)], dtype=np.int32)
        t0 = time.perf_counter()
        m = CO.map_rows(ref, exp, pidx, val=exp.columns[0][1], valid=None, bin_size=50000)
        dt = time.perf_counter() - t0
        return {"value": (ref.n + exp.n) / dt, "unit": "input regions/s", "cores": cores, "kind": "port",
                "sample": "C3 sample: %d reference x %d experiment samples (%d + %d regions, %d output rows), %.1f s; C/OpenMP restatement of the bin-based algorithm, not GMQL-Spark" % (
                    nr, ne, ref.n, exp.n, len(m["count"]), dt)}


class C4(Workload):
    VARIANTS = (("DLE(10000) CAT", [(0, 10001)], 4), ("MD(1),DOWN RIGHT", [(2, 1), (4, 0)], 2), ("DLE(10000),MD(1),DOWN RIGHT", [(0, 10001), (2, 1), (4, 0)], 2))

    def __init__(self, args, ex, L, torch):
        from gmql_b200 import synth
        self.args, self.ex, self.L, self.torch, self.synth = args, ex, L, torch, synth
        self.nl = self.nrg = max(2, int(200 * args.scale))
        self.config = {"workload": "C4: JOIN DLE(10000) CAT; JOIN MD(1),DOWN RIGHT; JOIN DLE(10000),MD(1),DOWN RIGHT  (BASELINE.json configs[3])", "left_samples": self.nl,
                       "left_regions_per_sample": 5000, "right_samples": self.nrg, "right_regions_per_sample": 50000,
                       "input_regions": 3 * (self.nl * 5000 + self.nrg * 50000), "sharding": "single GPU",
                       "l2": "outputs of 0.3 - 1.1 GB per join, far larger than the 126 MB L2"}
        self.left = self.right = None

    def data(self):
        if self.left is None:
            self.left = self.synth.promoter_samples(self.nl, 5000, 4)
            self.right = self.synth.peak_samples(self.nrg, 50000, 4, with_values=())
        return self.left, self.right

    def hosts(self):
        import bench as B
        left, right = self.data()
        return [B.narrow_host_struct(self.L, self.torch, left, keep_strand=True), B.narrow_host_struct(self.L, self.torch, right)]

    def ops(self, views):
        L, ex = self.L, self.ex
        vl, vr = views
        nl, nrg = self.nl, self.nrg
        pairs = (L.Pair * (nl * nrg))()
        for a in range(nl):
            for b in range(nrg):
                pairs[a * nrg + b].left_sample, pairs[a * nrg + b].right_sample = a, b
        self._keep = [pairs]
        spec = [(L.COL_JOIN_LEFT_ROW, np.int32), (L.COL_JOIN_RIGHT_ROW, np.int32), (L.COL_JOIN_PAIR, np.int32)]
        left, right = self.data()
        out = []
        for name, atoms, builder in self.VARIANTS:
            at = (L.Atom * len(atoms))()
            for i, (kd, arg) in enumerate(atoms):
                at[i].kind, at[i].arg = kd, arg
            self._keep.append(at)

            def run(at=at, n_at=len(atoms), builder=builder):
                res = C.c_void_p()
                ex._check(ex.lib.gmql_join(ex.ctx, C.byref(vl), C.byref(vr), pairs, nl * nrg, at, n_at, builder | 0x100, 5000, 100000, None, None, C.byref(res)))
                return res
            out.append((name, run, spec, left.n + right.n))
        return out

    def alg_bytes(self, label, n_out):
        left, right = self.data()
        return 9 * (left.n + right.n) + 8 * n_out

    def cpu(self):
        """The C port has no JOIN: the literal Python restatement (oracle/gmql_oracle.py) on a small sample, one core."""
        from oracle import gmql_oracle as O
        left = self.synth.promoter_samples(2, 5000, 4)
        right = self.synth.peak_samples(2, 50000, 4, with_values=())
        keep_l = left.chrom >= 20
        keep_r = right.chrom >= 20
        from gmql_b200 import dist as D
        l2, r2 = D._take(left, keep_l), D._take(right, keep_r)
        lrec, rrec = l2.to_records(), r2.to_records()
        pairs = {(int(a), int(b)): O.md5_pair_id(int(a), int(b)) for a in l2.sample_ids for b in r2.sample_ids}
        t0 = time.perf_counter()
        n_out = 0
        for atoms, builder in (([("DISTLESS", 10001)], "CONTIG"), ([("MD", 1), ("DOWN", 0)], "RIGHT"), ([("DISTLESS", 10001), ("MD", 1), ("DOWN", 0)], "RIGHT")):
            n_out += len(O.join_reference(lrec, rrec, pairs, atoms, builder))
        dt = time.perf_counter() - t0
        return {"value": 3 * (l2.n + r2.n) / dt, "unit": "input regions/s", "cores": 1, "kind": "port",
                "sample": "C4 sample: 2 x 2 samples restricted to chr21, chr22, chrX, chrY (%d + %d regions), the three joins, %.1f s, %d output records; literal Python restatement (the C port has no JOIN), one core, not GMQL-Spark" % (
                    l2.n, r2.n, dt, n_out)}
