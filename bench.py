#!/usr/bin/env python
"""bench.py — GMQL region-operator hot path on B200: input regions/s.

Default workload = BASELINE.json configs[4] (the one `metric` is quoted on: "... at 1/2/4/8 B200"):
    C5:  C = COVER(2, ANY) E;  M = MAP() C E;   E = 2 000 synthetic narrowPeak samples x 50 000 regions (hg38), 1e8 regions,
sharded across ranks with no data-path collective.  A step = one pass of the pipeline over the rank's shard.  `value` =
input regions of ALL ranks / max-over-ranks step time, inputs resident in HBM (as uploaded: 7 bytes per region).  `e2e` = the
same through the C ABI with HOST (pinned) buffers: H2D of the SoA columns, COVER, MAP, D2H of the COVER table and the MAP
counts, all inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c2|c3|c4]

--workload c2 | c3 | c4 time the other BASELINE.json configs at their full sizes on ONE GPU (COVER / FLAT / SUMMIT / HISTOGRAM
over 1 000 x 50 000 peaks; MAP SUM / AVG / MAX of 100 x 1 000 samples; JOIN DLE / MD(1),DOWN between 200 x 200 samples) with the
same keys (`roofline`, `cpu_baseline`, `e2e`); the committed lines live under profiles/.

`--impl reference` times the reference's CPU algorithm (oracle/gmql_oracle_c.c, a bin-based C/OpenMP port; GMQL-Spark itself
cannot run: the probe for a JVM / spark-submit is recorded in the line) on a bounded sample of the same workload with all host
cores.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
METRIC = "MAP/COVER/JOIN input regions/s at 1/2/4/8 B200 vs GMQL-Spark local[N] on host"
ANYV = 2 ** 31 - 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (the recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_ALL_CPUS = None


def bind_to_gpu_numa_node(torch, local):
    """Pins this process to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function), so that the pinned
    host buffers of the end-to-end arm are allocated on the GPU's NUMA node; returns a short description.  The CPU arm
    restores the full CPU set (restore_all_cpus)."""
    global _ALL_CPUS
    try:
        _ALL_CPUS = os.sched_getaffinity(0)
        rt = C.CDLL("libcudart.so")
        buf = C.create_string_buffer(32)
        if rt.cudaDeviceGetPCIBusId(buf, 32, local) != 0:
            return "unbound"
        addr = buf.value.decode().lower()
        path = "/sys/bus/pci/devices/%s/local_cpulist" % addr
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= _ALL_CPUS
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%s: %d local CPUs" % (addr, len(cpus))
    except Exception as e:   # no sysfs / no permission: run unbound
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def restore_all_cpus():
    if _ALL_CPUS:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)
        except Exception:
            pass


def host_cores():
    try:
        return len(_ALL_CPUS or os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def spark_probe():
    """is a real GMQL-Spark baseline possible on this box?  (java, spark-submit, GMQL jars)"""
    def ver(cmd):
        exe = shutil.which(cmd[0])
        if not exe:
            return None
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=20)
            return (p.stderr or p.stdout).strip().splitlines()[0][:80]
        except Exception:
            return "present"
    jars = any(os.path.exists(p) for p in ("/opt/gmql", "/usr/share/gmql", os.path.join(ROOT, "baseline", "_ref")))
    return {"java": ver(["java", "-version"]), "spark_submit": ver(["spark-submit", "--version"]), "gmql_jars": jars,
            "conclusion": "GMQL-Spark local[N] not runnable here: the CPU arm is the C/OpenMP restatement of its bin-based algorithm (oracle/)"}


_REAL_STDOUT = sys.stdout


def pinned(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def narrow_host_struct(L, torch, ds, with_cols=False, keep_strand=False):
    """gmql_regions over pinned host columns in the ABI's compact encodings: 8-bit chromosome index, int32 start, region
    lengths as uint16 instead of stops, sample_offsets (one run of rows per sample file) and no strand column when every region
    is '*' (narrowPeak '.'): 7 bytes per region (+ 8 per value column).  Returns (struct, keepalive, h2d bytes)."""
    n = ds.n
    assert len(ds.chrom_names) <= 256 and np.all(np.diff(ds.sample) >= 0) and int((ds.stop - ds.start).max(initial=0)) < 65536
    keep = {"chrom": pinned(torch, ds.chrom.astype(np.uint8)), "start": pinned(torch, ds.start.astype(np.int32)),
            "len": pinned(torch, (ds.stop - ds.start).astype(np.uint16))}
    n_s = int(ds.sample_ids.shape[0])
    off = np.searchsorted(ds.sample, np.arange(n_s + 1), side="left").astype(np.int64)
    keep["off"], keep["ids"] = off, ds.sample_ids
    r = L.Regions()
    r.n, r.location, r.coord_bits = n, 0, 32
    r.chrom, r.chrom_bits = keep["chrom"].data_ptr(), 8
    r.start = keep["start"].data_ptr()
    r.stop, r.stop_is_length = keep["len"].data_ptr(), 16
    r.strand = None
    nbytes = 7 * n + int(off.nbytes)
    if keep_strand and ds.strand.any():
        keep["strand"] = pinned(torch, ds.strand)
        r.strand = keep["strand"].data_ptr()
        nbytes += n
    r.sample = None
    r.sample_offsets = off.ctypes.data
    r.n_chrom, r.n_samples = len(ds.chrom_names), n_s
    r.sample_ids = ds.sample_ids.ctypes.data
    r.n_cols = 0
    if with_cols:
        nums = [c for c in ds.columns if c[0] == "num"]
        cols = (C.c_void_p * len(nums))()
        for k, c in enumerate(nums):
            keep["col%d" % k] = pinned(torch, c[1].astype(np.float64))
            cols[k] = keep["col%d" % k].data_ptr()
            nbytes += 8 * n
        keep["cols"] = cols
        r.n_cols = len(nums)
        r.cols = C.cast(cols, C.POINTER(C.c_void_p))
    return r, keep, nbytes


def result_cols(ex, L, res, spec):
    out = []
    for cid, dt in spec:
        n = int(ex.lib.gmql_result_column_len(res, cid))
        a = np.empty(max(n, 0), dtype=dt)
        if n > 0:
            ex._check(ex.lib.gmql_result_column_copy(ex.ctx, res, cid, a.ctypes.data))
        out.append(a)
    return out


def profile_kernels(ex, fn):
    lib, ctx = ex.lib, ex.ctx
    lib.gmql_ctx_set_profiling(ctx, 1)
    fn()
    buf = C.create_string_buffer(1 << 16)
    lib.gmql_ctx_profile_report(ctx, buf, len(buf))
    lib.gmql_ctx_set_profiling(ctx, 0)
    prof = []
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.rsplit(" ", 2)
        prof.append((name, int(cnt), float(ms)))
    return prof


def kernel_traffic(name):
    """DRAM bytes per launch of `name` from this round's committed `ncu --set full` capture of the same workload, or None when the
    capture does not hold that kernel (a stale number is worse than none)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_kernel_traffic.json")))
        return d.get(name)
    except (OSError, ValueError):
        return None


# ---------------------------------------------------------------------------------------------------------------------
# C5
def make_c5(n_samples, n_regions, parts):
    from gmql_b200 import synth
    return synth.peak_samples(n_samples, n_regions, 5, with_values=(), chrom_subset=parts)


def c5_config(args, world):
    """The workload, identically worded by both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": "C5: C = COVER(2,ANY) E; M = MAP() C E  (BASELINE.json configs[4])", "samples": args.samples, "regions_per_sample": args.regions,
            "input_regions": args.samples * args.regions,
            "sharding": ("by chromosome (LPT packing), no collective on the data path" if world > 1 else "single GPU"),
            "l2": "inputs resident in HBM (0.7 GB per 1e8 regions) and far larger than the 126 MB L2; nothing is reused between steps"}


def c5_algorithmic_bytes(n_in, n_cover, n_samples):
    # COVER (9 + 8*A_in)*N_in + (28 + 8*A)*N_out ; MAP 9*N_exp + 9*N_ref + 4*N_rows
    return 9 * n_in + 28 * n_cover + 9 * n_in + 9 * n_cover + 4 * n_cover * n_samples


class Pipeline:
    """C = COVER(2, ANY) E; M = MAP() C E through the C ABI."""

    def __init__(self, ex, n_samples):
        from gmql_b200 import _lib as L
        self.L, self.ex, self.lib = L, ex, ex.lib
        self.n_samples = n_samples
        self.mn = np.array([2], dtype=np.int32)
        self.mx = np.array([ANYV], dtype=np.int32)
        self.pairs = (L.Pair * n_samples)()
        for b in range(n_samples):
            self.pairs[b].left_sample = 0
            self.pairs[b].right_sample = b
        self.no_aggs = (L.Agg * 1)()

    def run_on_view(self, view):
        """COVER -> MAP on a device-located view; returns (cover result, map result) handles."""
        lib, ctx = self.lib, self.ex.ctx
        cov = C.c_void_p()
        self.ex._check(lib.gmql_cover(ctx, C.byref(view), None, 1, 0, self.mn.ctypes.data, self.mx.ctypes.data, 2000, self.no_aggs, 0, C.byref(cov)))
        cview = self.L.Regions()
        self.ex._check(lib.gmql_result_as_regions(ctx, cov, C.byref(cview)))
        mp = C.c_void_p()
        self.ex._check(lib.gmql_map(ctx, C.byref(cview), C.byref(view), self.pairs, self.n_samples, self.no_aggs, 0, C.byref(mp)))
        return cov, mp

    def free(self, *handles):
        for h in handles:
            self.lib.gmql_result_free(self.ex.ctx, h)


def cpu_slice(args):
    """Bounded sample of C5 for the CPU arm: all samples, the smallest chromosomes only (keeps the coverage depth)."""
    from gmql_b200 import synth
    order = np.argsort(synth.HG38_SIZES)
    chroms = sorted(int(c) for c in order[: args.cpu_chroms])
    n_samples = min(args.samples, args.cpu_samples)
    ds = make_c5(n_samples, args.regions, chroms)
    frac = float(synth.HG38_SIZES[chroms].sum()) / float(synth.HG38_SIZES.sum())
    return ds, n_samples, "C5 sample: %d samples x %d regions restricted to the %d smallest hg38 chromosomes (%.1f%% of the genome, %d regions)" % (
        n_samples, args.regions, len(chroms), 100 * frac, ds.n)


def cpu_pipeline_once(ds, n_samples):
    """The C/OpenMP port on the host: returns (seconds, cover rows sorted by (chrom, start), MAP counts [n_samples, n_cover])."""
    from oracle import c_oracle as CO
    from gmql_b200.regions import RegionDataset
    t0 = time.perf_counter()
    r = CO.cover_rows(ds, 0, [2], [ANYV], None, 2000)
    o = np.lexsort((r["start"], r["chrom"]))
    r = {k: v[o] for k, v in r.items()}
    nc = len(r["start"])
    ref = RegionDataset(ds.chrom_names, r["chrom"], r["start"], r["stop"], r["strand"], np.zeros(nc, np.int32), np.array([0], np.int64))
    pidx = np.stack([np.zeros(n_samples, np.int32), np.arange(n_samples, dtype=np.int32)], axis=1)
    m = CO.map_rows(ref, ds, pidx, None, None, 50000)
    return time.perf_counter() - t0, r, m["count"].reshape(n_samples, nc)


def c5_parity_and_cpu(args, ex, L, torch):
    """Outside every timed region: the CPU port on the bounded sample (cpu_baseline) and the GPU on the SAME sample, compared row
    for row (COVER table + per-row MAP counts), plus GMAP4's key-aliasing count on the GPU's table (SURVEY A-C10, expected 0)."""
    from oracle import c_oracle as CO
    from gmql_b200.alias import count_aliased_groups
    restore_all_cpus()
    cores = CO.set_threads(host_cores())
    ds, n_samples, what = cpu_slice(args)
    dt, want, want_m = cpu_pipeline_once(ds, n_samples)
    # the GPU on the same sample
    host, keep, _ = narrow_host_struct(L, torch, ds)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    pipe = Pipeline(ex, n_samples)
    gpu_ms = []
    for _ in range(4):
        t0 = time.perf_counter()
        cov, mp = pipe.run_on_view(view)
        ex.lib.gmql_ctx_sync(ex.ctx)
        gpu_ms.append((time.perf_counter() - t0) * 1e3)
        last = (cov, mp)
        if _ < 3:
            pipe.free(cov, mp)
    cov, mp = last
    g = result_cols(ex, L, cov, [(L.COL_COV_GROUP, np.int32), (L.COL_COV_CHROM, np.int32), (L.COL_COV_START, np.int64), (L.COL_COV_STOP, np.int64),
                                 (L.COL_COV_STRAND, np.int8), (L.COL_COV_ACC, np.int32), (L.COL_COV_J1, np.float64), (L.COL_COV_J2, np.float64)])
    (cnt,) = result_cols(ex, L, mp, [(L.COL_MAP_COUNT, np.int32)])
    pipe.free(cov, mp)
    ex.lib.gmql_dataset_free(ex.ctx, dset)
    ok = len(g[1]) == len(want["start"])
    if ok:
        for a, k in zip(g[1:], ("chrom", "start", "stop", "strand", "acc", "j1", "j2")):
            ok = ok and bool(np.array_equal(a, want[k]))
        ok = ok and bool(np.array_equal(cnt.reshape(n_samples, -1), want_m))
    aliased = count_aliased_groups([0], ds.chrom_names, *g[:6])
    cpu = {"value": ds.n / dt, "unit": "input regions/s", "cores": cores, "kind": "port",
           "sample": what + "; %.1f s; oracle-restatement (C/OpenMP port of the bin-based algorithm), not GMQL-Spark" % dt,
           "cover_regions": int(len(want["start"])), "map_count_checksum": int(want_m.astype(np.int64).sum()),
           "gpu_same_sample": {"ms_per_step_wall": round(float(np.median(gpu_ms[1:])), 3), "input_regions_per_s": ds.n / (float(np.median(gpu_ms[1:])) * 1e-3),
                               "what": "this back end on the identical sample, device-resident, wall clock around COVER + MAP (same config for both sides)"}}
    return cpu, ok, aliased


def run_c5(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: this back end has no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gmql_b200 import GMQLCudaExecutor, synth, _lib as L
    from gmql_b200.testing import install_env_hooks
    numa = bind_to_gpu_numa_node(torch, local)      # pinned staging buffers on the GPU's own NUMA node (host -> device copies)

    n_samples, n_regions = args.samples, args.regions
    parts = synth.chrom_partition(world)[rank] if world > 1 else None
    if world == 1 and args.emulate_shard:                 # one GPU running rank r's shard of a w-GPU job (tuning aid, not a bench line)
        r, w = (int(x) for x in args.emulate_shard.split("/"))
        parts = synth.chrom_partition(w)[r]
    ds = make_c5(n_samples, n_regions, parts)
    n_in = ds.n
    stream = torch.cuda.current_stream()
    ex = GMQLCudaExecutor(device=local, stream=stream.cuda_stream)
    install_env_hooks(ex.lib)                             # GMQL_* tuning variables -> per-context options (off unless set)
    ex.set_option("defer_check", "1")                     # one host synchronisation per COVER -> MAP step (include/gmql_cuda.h)
    lib, ctx = ex.lib, ex.ctx
    pipe = Pipeline(ex, n_samples)
    host, keep, h2d_bytes = narrow_host_struct(L, torch, ds)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------------------------------
    dset, view = C.c_void_p(), L.Regions()
    ex._check(lib.gmql_dataset_upload(ctx, C.byref(host), C.byref(dset), C.byref(view)))
    n_cover = n_rows = 0
    for _ in range(args.warmup):
        cov, mp = pipe.run_on_view(view)
        n_cover, n_rows = int(lib.gmql_result_rows(cov)), int(lib.gmql_result_rows(mp))
        pipe.free(cov, mp)
    ex._check(lib.gmql_ctx_sync(ctx))
    gpu_index = local if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
    sampler = ClockSampler(gpu_index) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = ex.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        cov, mp = pipe.run_on_view(view)
        pipe.free(cov, mp)
        evs[k + 1].record(stream)
    barrier()
    ex._check(lib.gmql_ctx_sync(ctx))                     # deferred validity flags of the timed steps
    launches = ex.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_dev = evs[0].elapsed_time(evs[-1]) / args.steps
    each = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps))
    ms_median = each[len(each) // 2]

    # ---- end-to-end arm: host buffers in, result tables out --------------------------------------------------------------
    out_cov = {k: torch.empty(max(n_cover, 1), dtype=dt).pin_memory() for k, dt in
               (("chrom", torch.int32), ("start", torch.int64), ("stop", torch.int64), ("acc", torch.int32), ("j1", torch.float64), ("j2", torch.float64))}
    out_cnt = torch.empty(max(n_rows, 1), dtype=torch.int32).pin_memory()
    col_of = {"chrom": L.COL_COV_CHROM, "start": L.COL_COV_START, "stop": L.COL_COV_STOP, "acc": L.COL_COV_ACC, "j1": L.COL_COV_J1, "j2": L.COL_COV_J2}
    h2d_ms, e2e_phases = [], []          # wall ms per phase of every end-to-end step: upload (+sync), COVER + MAP, result copies, frees

    def e2e_step():
        t_a = time.perf_counter()
        d, v = C.c_void_p(), L.Regions()
        ex._check(lib.gmql_dataset_upload(ctx, C.byref(host), C.byref(d), C.byref(v)))
        h2d_ms.append(float(lib.gmql_last_h2d_ms(ctx)))
        t_b = time.perf_counter()
        cov, mp = pipe.run_on_view(v)
        assert int(lib.gmql_result_rows(cov)) == n_cover and int(lib.gmql_result_rows(mp)) == n_rows
        t_c = time.perf_counter()
        for k, t in out_cov.items():
            ex._check(lib.gmql_result_column_copy(ctx, cov, col_of[k], t.data_ptr()))
        ex._check(lib.gmql_result_column_copy(ctx, mp, L.COL_MAP_COUNT, out_cnt.data_ptr()))
        t_d = time.perf_counter()
        pipe.free(cov, mp)
        lib.gmql_dataset_free(ctx, d)
        t_e = time.perf_counter()
        e2e_phases.append([round((y - x) * 1e3, 2) for x, y in ((t_a, t_b), (t_b, t_c), (t_c, t_d), (t_d, t_e))])

    e2e_steps = max(1, min(args.steps, 10))
    e2e_step()                            # one untimed step (first use of this shape of buffers)
    h2d_ms[:], e2e_phases[:] = [], []
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_each = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        e2e_step()
        e2e_each.append(round((time.perf_counter() - ts) * 1e3, 2))     # every step ends with synchronous column copies
    e1.record(stream)
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    d2h_bytes = sum(int(n_cover * t.element_size()) for t in out_cov.values()) + 4 * n_rows
    checksum = int(out_cnt[:n_rows].to(torch.int64).sum().item())

    # ---- per-kernel pass (CUDA events around every launch), one extra step ---------------------------------------------
    def one():
        cov, mp = pipe.run_on_view(view)
        pipe.free(cov, mp)
    prof = profile_kernels(ex, one)
    lib.gmql_dataset_free(ctx, dset)

    # ---- reduce over ranks ---------------------------------------------------------------------------------------------
    stats = torch.tensor([ms_dev, ms_e2e, ms_median], dtype=torch.float64, device="cuda")
    sums = torch.tensor([n_in, n_cover, n_rows, h2d_bytes, d2h_bytes, launches, checksum], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max, ms_median_max = [float(x) for x in stats.tolist()]
    tot_in, tot_cover, tot_rows, tot_h2d, tot_d2h, tot_launch, tot_check = [float(x) for x in sums.tolist()]

    if rank == 0:
        peak, peak_kind = measured_peak()
        top = prof[0] if prof else ("none", 1, 0.0)
        top_ms = top[2] / max(1, top[1])
        # algorithmic bytes of one launch of the dominant kernel: SURVEY §8(d)'s per-unit figure (9 B per input region of
        # COVER or of MAP's experiment side) x the regions one launch processes (every launch sweeps this rank's whole shard)
        alg_launch = 9.0 * n_in
        achieved = alg_launch / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        step_bytes = c5_algorithmic_bytes(tot_in, tot_cover, n_samples)
        cfg = c5_config(args, world)
        out = {
            "metric": METRIC, "value": tot_in / (ms_dev_max * 1e-3), "unit": "input regions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev_max, "ms_per_step_median": ms_median_max,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32 coordinates / f64 values",
            "data": "synthetic", "config": cfg,
            "result": {"input_regions": int(tot_in), "cover_regions": int(tot_cover), "map_rows": int(tot_rows), "map_count_checksum": int(tot_check)},
            "e2e": {"value": tot_in / (ms_e2e_max * 1e-3), "unit": "input regions/s", "ms_per_step": ms_e2e_max,
                    "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h), "steps": e2e_steps,
                    "h2d_ms_rank0": round(float(np.median(h2d_ms)), 3), "step_ms_rank0": e2e_each,
                    "phase_ms_rank0": {"order": ["upload", "cover+map", "result copies", "frees"], "steps": e2e_phases}, "host_binding_rank0": numa},
            "gpu_launches": int(tot_launch),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top[0], "launches_per_step": top[1], "ms_per_launch": top_ms,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": kernel_traffic(top[0]) if world == 1 else None,
                         "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg_launch},
            "roofline_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (ms_dev_max * 1e-3) / 1e9 / world,
                              "frac_per_gpu": step_bytes / (ms_dev_max * 1e-3) / 1e9 / world / peak},
            "kernels": [{"name": n, "launches": c, "ms": round(m, 4)} for (n, c, m) in prof[:12]],
            "host_gap_ms_rank0": round(ms_dev - sum(m for (_n, _c, m) in prof), 4),
        }
        if world == 1 and not args.no_cpu:
            cpu, ok, aliased = c5_parity_and_cpu(args, ex, L, torch)
            out["cpu_baseline"] = cpu
            out["parity_checked"] = ok
            out["gmap4_aliased_groups"] = aliased
            out["gmql_spark_probe"] = spark_probe()
        print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    ex.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference_c5(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    from oracle import c_oracle as CO
    cores = CO.set_threads(host_cores())
    ds, n_samples, what = cpu_slice(args)
    for _ in range(args.warmup):
        cpu_pipeline_once(ds, n_samples)
    times = []
    for _ in range(args.steps):
        dt, r, m = cpu_pipeline_once(ds, n_samples)
        times.append(dt)
    dt = float(np.mean(times))
    val = ds.n / dt
    out = {
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "input regions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64 coordinates / f64 values",
        "data": "synthetic", "config": c5_config(args, world),
        "result": {"sample_input_regions": int(ds.n), "cover_regions": int(len(r["start"])), "map_count_checksum": int(m.astype(np.int64).sum())},
        "cpu_baseline": {"value": val, "unit": "input regions/s", "cores": cores, "kind": "port",
                         "sample": what + "; every step is one pass of COVER + MAP over this sample; oracle-restatement (C/OpenMP port of the bin-based algorithm), not GMQL-Spark"},
        "gmql_spark_probe": spark_probe(),
        "e2e": {"value": val, "unit": "input regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), file=_REAL_STDOUT, flush=True)


def main():
    # NCCL and friends print banners on stdout; the contract is ONE JSON line there.  Everything else goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c2", "c3", "c4"])
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--regions", type=int, default=50000)
    ap.add_argument("--cpu-chroms", type=int, default=8, help="CPU arm: number of smallest chromosomes in the bounded sample")
    ap.add_argument("--cpu-samples", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0, help="c2/c3/c4: scales the sample counts (quick runs)")
    ap.add_argument("--emulate-shard", default="", help="r/w: run rank r's chromosome shard of a w-GPU job on one GPU (tuning aid)")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    if args.workload != "c5":
        import bench_configs
        bench_configs.run(args, _REAL_STDOUT)
    elif args.impl == "reference":
        run_reference_c5(args)
    else:
        run_c5(args)


if __name__ == "__main__":
    main()
