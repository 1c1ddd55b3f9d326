#!/usr/bin/env python
"""bench.py — GMQL region-operator hot path on B200: input regions/s.

Default workload = BASELINE.json configs[4] (the one `metric` is quoted on: "... at 1/2/4/8 B200"):
    C5:  C = COVER(2, ANY) E;  M = MAP() C E;   E = 2 000 synthetic narrowPeak samples x 50 000 regions (hg38), 1e8 regions,
sharded across ranks by chromosome (north_star: "partition naturally by chromosome"), no data-path collective.
A step = one pass of the pipeline over the rank's shard.  `value` = input regions of ALL ranks / max-over-ranks step time,
inputs resident in HBM.  `e2e` = the same through the C ABI with HOST (pinned) buffers: H2D of the SoA columns, COVER, MAP,
D2H of the COVER table and the MAP counts, all inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c2|c3|c4] [--samples S]

`--impl reference` times the reference's CPU algorithm (oracle/gmql_oracle_c.c, a bin-based C/OpenMP port; GMQL-Spark
itself cannot run: no JVM in the image) on a bounded slice of the same workload with all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


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
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bus is None:
            import ctypes
            rt = ctypes.CDLL("libcudart.so")
            buf = ctypes.create_string_buffer(32)
            if rt.cudaDeviceGetPCIBusId(buf, 32, local) != 0:
                return "unbound"
            addr = buf.value.decode().lower()
        else:
            addr = "0000:%02x:00.0" % int(bus)
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


# ---------------------------------------------------------------------------------------------------------------------
def make_c5(n_samples, n_regions, parts):
    from gmql_b200 import synth
    return synth.peak_samples(n_samples, n_regions, 5, with_values=(), chrom_subset=parts)


_REAL_STDOUT = sys.stdout


def c5_algorithmic_bytes(n_in, n_cover, n_samples):
    # COVER (9 + 8*A_in)*N_in + (28 + 8*A)*N_out ; MAP 9*N_exp + 9*N_ref + 4*N_rows
    return 9 * n_in + 28 * n_cover + 9 * n_in + 9 * n_cover + 4 * n_cover * n_samples


class Pipeline:
    """C = COVER(2, ANY) E; M = MAP() C E through the C ABI."""

    def __init__(self, ex, ds, n_samples):
        from gmql_b200 import _lib as L
        self.L, self.ex, self.lib, self.ds = L, ex, ex.lib, ds
        self.n_samples = n_samples
        self.mn = np.array([2], dtype=np.int32)
        self.mx = np.array([2 ** 31 - 1], dtype=np.int32)
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


def host_struct(L, cols, n, n_chrom, n_samples, sample_ids, sample_offsets):
    """gmql_regions over pinned host columns, in the ABI's compact encodings: 8-bit chromosome index, int32 start/stop,
    region lengths as uint16 instead of stops (narrowPeak lengths stay far below 65 536), no strand column (narrowPeak
    '.' -> every region '*') and sample_offsets (one run of rows per sample file): 7 bytes per region."""
    r = L.Regions()
    r.n = n
    r.location = 0
    r.coord_bits = 32
    r.chrom = cols["chrom"].data_ptr()
    r.chrom_bits = 8
    r.start = cols["start"].data_ptr()
    r.stop = cols["len"].data_ptr()
    r.stop_is_length = 16
    r.strand = None
    r.sample = None
    r.sample_offsets = sample_offsets.ctypes.data
    r.n_chrom = n_chrom
    r.n_samples = n_samples
    r.sample_ids = sample_ids.ctypes.data
    r.n_cols = 0
    return r


def run_ours(args):
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
    numa = bind_to_gpu_numa_node(torch, local)      # pinned staging buffers on the GPU's own NUMA node (host -> device copies)

    n_samples, n_regions = args.samples, args.regions
    parts = synth.chrom_partition(world)[rank] if world > 1 else None
    if world == 1 and args.emulate_shard:                 # one GPU running rank r's shard of a w-GPU job (tuning aid, not a bench line)
        r, w = (int(x) for x in args.emulate_shard.split("/"))
        parts = synth.chrom_partition(w)[r]
    ds = make_c5(n_samples, n_regions, parts)
    n_in = ds.n
    assert len(ds.chrom_names) <= 256 and np.all(np.diff(ds.sample) >= 0) and int((ds.stop - ds.start).max()) < 65536
    cols = {
        "chrom": torch.from_numpy(ds.chrom.astype(np.uint8)).pin_memory(),
        "start": torch.from_numpy(ds.start.astype(np.int32)).pin_memory(),
        "len": torch.from_numpy((ds.stop - ds.start).astype(np.uint16)).pin_memory(),
    }
    sample_offsets = np.searchsorted(ds.sample, np.arange(n_samples + 1), side="left").astype(np.int64)
    h2d_bytes = sum(int(t.numel() * t.element_size()) for t in cols.values()) + int(sample_offsets.nbytes)
    stream = torch.cuda.current_stream()
    ex = GMQLCudaExecutor(device=local, stream=stream.cuda_stream)
    lib, ctx = ex.lib, ex.ctx
    pipe = Pipeline(ex, ds, n_samples)
    host = host_struct(L, cols, n_in, len(ds.chrom_names), n_samples, ds.sample_ids, sample_offsets)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------------------------------
    dset = C.c_void_p()
    view = L.Regions()
    ex._check(lib.gmql_dataset_upload(ctx, C.byref(host), C.byref(dset), C.byref(view)))
    n_cover = n_rows = 0
    for _ in range(args.warmup):
        cov, mp = pipe.run_on_view(view)
        n_cover, n_rows = int(lib.gmql_result_rows(cov)), int(lib.gmql_result_rows(mp))
        pipe.free(cov, mp)
    gpu_index = local if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
    sampler = ClockSampler(gpu_index) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = ex.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        cov, mp = pipe.run_on_view(view)
        pipe.free(cov, mp)
    e1.record(stream)
    barrier()
    launches = ex.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_dev = e0.elapsed_time(e1) / args.steps

    # ---- end-to-end arm: host buffers in, result tables out --------------------------------------------------------------
    out_cov = {k: torch.empty(max(n_cover, 1), dtype=dt).pin_memory() for k, dt in
               (("chrom", torch.int32), ("start", torch.int64), ("stop", torch.int64), ("acc", torch.int32), ("j1", torch.float64), ("j2", torch.float64))}
    out_cnt = torch.empty(max(n_rows, 1), dtype=torch.int32).pin_memory()
    col_of = {"chrom": L.COL_COV_CHROM, "start": L.COL_COV_START, "stop": L.COL_COV_STOP, "acc": L.COL_COV_ACC, "j1": L.COL_COV_J1, "j2": L.COL_COV_J2}

    h2d_ms = []

    e2e_phases = []          # wall ms per phase of every end-to-end step: upload (+sync), COVER + MAP, result copies, frees

    def e2e_step():
        t_a = time.perf_counter()
        d = C.c_void_p()
        v = L.Regions()
        ex._check(lib.gmql_dataset_upload(ctx, C.byref(host), C.byref(d), C.byref(v)))
        lib.gmql_ctx_sync(ctx)
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

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(3):        # warm-up: the stream-ordered pool needs a few rounds of upload / operate / free to stop growing
        e2e_step()
    h2d_ms[:] = h2d_ms[-1:]
    e2e_phases[:] = e2e_phases[-1:]
    barrier()
    t0 = time.perf_counter()
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
    lib.gmql_ctx_set_profiling(ctx, 1)
    cov, mp = pipe.run_on_view(view)
    pipe.free(cov, mp)
    buf = C.create_string_buffer(1 << 16)
    lib.gmql_ctx_profile_report(ctx, buf, len(buf))
    lib.gmql_ctx_set_profiling(ctx, 0)
    prof = []
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.rsplit(" ", 2)
        prof.append((name, int(cnt), float(ms)))
    lib.gmql_dataset_free(ctx, dset)

    # ---- reduce over ranks ---------------------------------------------------------------------------------------------
    stats = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    sums = torch.tensor([n_in, n_cover, n_rows, h2d_bytes, d2h_bytes, launches, checksum], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max = [float(x) for x in stats.tolist()]
    tot_in, tot_cover, tot_rows, tot_h2d, tot_d2h, tot_launch, tot_check = [float(x) for x in sums.tolist()]

    if rank == 0:
        peak, peak_kind = measured_peak()
        top = prof[0] if prof else ("none", 1, 0.0)
        top_ms = top[2] / max(1, top[1])
        # algorithmic bytes of one launch of the dominant kernel: SURVEY §8(d)'s COVER per-unit figure (9 B per input
        # region) x the regions one launch processes (every launch of these kernels sweeps this rank's whole shard)
        alg_launch = 9.0 * n_in
        achieved = alg_launch / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        step_bytes = c5_algorithmic_bytes(tot_in, tot_cover, n_samples)
        traffic = None
        try:   # DRAM bytes per launch of that kernel from the committed ncu capture of this workload (N = 1 only)
            if world == 1 and n_samples == 2000 and n_regions == 50000:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_kernel_traffic.json"))).get(top[0])
        except (OSError, ValueError):
            traffic = None
        out = {
            "metric": "MAP/COVER/JOIN input regions/s at 1/2/4/8 B200 vs GMQL-Spark local[N] on host",
            "value": tot_in / (ms_dev_max * 1e-3),
            "unit": "input regions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev_max,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32 coordinates / f64 values",
            "data": "synthetic",
            "config": {"workload": "C5: C = COVER(2,ANY) E; M = MAP() C E  (BASELINE.json configs[4])", "samples": n_samples,
                       "regions_per_sample": n_regions, "input_regions": int(tot_in), "cover_regions": int(tot_cover),
                       "map_rows": int(tot_rows), "sharding": "by chromosome (LPT packing), no collective on the data path" if world > 1 else "single GPU",
                       "l2": "device-resident inputs (%.2f GB per rank) larger than the 126 MB L2" % (16.0 * n_in / 1e9), "map_count_checksum": int(tot_check)},
            "e2e": {"value": tot_in / (ms_e2e_max * 1e-3), "unit": "input regions/s", "ms_per_step": ms_e2e_max,
                    "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h), "steps": e2e_steps,
                    "h2d_ms_rank0": round(float(np.median(h2d_ms[1:] or h2d_ms)), 3), "step_ms_rank0": e2e_each, "h2d_each_ms_rank0": [round(float(x), 2) for x in h2d_ms[1:]],
                    "phase_ms_rank0": {"order": ["upload", "cover+map", "result copies", "frees"], "steps": e2e_phases[1:]}, "host_binding_rank0": numa},
            "gpu_launches": int(tot_launch),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top[0], "launches_per_step": top[1], "ms_per_launch": top_ms,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg_launch},
            "roofline_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (ms_dev_max * 1e-3) / 1e9 / world,
                              "frac_per_gpu": step_bytes / (ms_dev_max * 1e-3) / 1e9 / world / peak},
            "kernels": [{"name": n, "launches": c, "ms": round(m, 4)} for (n, c, m) in prof[:12]],
        }
        if world == 1 and not args.no_cpu:
            restore_all_cpus()
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    ex.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def cpu_slice(args):
    """Bounded sample of C5 for the CPU arm: all samples, the smallest chromosomes only (keeps the coverage depth)."""
    from gmql_b200 import synth
    order = np.argsort(synth.HG38_SIZES)
    chroms = sorted(int(c) for c in order[: args.cpu_chroms])
    n_samples = min(args.samples, args.cpu_samples)
    ds = make_c5(n_samples, args.regions, chroms)
    frac = float(synth.HG38_SIZES[chroms].sum()) / float(synth.HG38_SIZES.sum())
    return ds, n_samples, "C5 slice: %d samples x %d regions restricted to the %d smallest hg38 chromosomes (%.1f%% of the genome, %d regions)" % (
        n_samples, args.regions, len(chroms), 100 * frac, ds.n)


def cpu_pipeline_once(ds, n_samples):
    from oracle import c_oracle as CO
    from gmql_b200.regions import RegionDataset
    t0 = time.perf_counter()
    r = CO.cover_rows(ds, 0, [2], [2 ** 31 - 1], None, 2000)
    nc = len(r["start"])
    ref = RegionDataset(ds.chrom_names, r["chrom"], r["start"], r["stop"], r["strand"], np.zeros(nc, np.int32), np.array([0], np.int64))
    pidx = np.stack([np.zeros(n_samples, np.int32), np.arange(n_samples, dtype=np.int32)], axis=1)
    m = CO.map_rows(ref, ds, pidx, None, None, 50000)
    return time.perf_counter() - t0, nc, int(m["count"].astype(np.int64).sum())


def cpu_baseline(args):
    from oracle import c_oracle as CO
    ds, n_samples, what = cpu_slice(args)
    dt, nc, chk = cpu_pipeline_once(ds, n_samples)
    return {"value": ds.n / dt, "unit": "input regions/s", "cores": CO.num_threads(), "kind": "port",
            "sample": what + "; %.1f s; oracle-restatement (C/OpenMP port of the bin-based algorithm), not GMQL-Spark" % dt,
            "cover_regions": nc, "map_count_checksum": chk}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import c_oracle as CO
    ds, n_samples, what = cpu_slice(args)
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline_once(ds, n_samples)
    times = []
    for _ in range(args.steps):
        dt, nc, chk = cpu_pipeline_once(ds, n_samples)
        times.append(dt)
    dt = float(np.mean(times))
    val = ds.n / dt
    out = {
        "impl": "reference",
        "metric": "MAP/COVER/JOIN input regions/s at 1/2/4/8 B200 vs GMQL-Spark local[N] on host",
        "value": val, "unit": "input regions/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64 coordinates / f64 values",
        "data": "synthetic",
        "config": {"workload": "C5: C = COVER(2,ANY) E; M = MAP() C E  (BASELINE.json configs[4])", "samples": n_samples, "regions_per_sample": args.regions,
                   "input_regions": int(ds.n), "cover_regions": nc, "map_count_checksum": chk},
        "cpu_baseline": {"value": val, "unit": "input regions/s", "cores": CO.num_threads(), "kind": "port",
                         "sample": what + "; oracle-restatement (C/OpenMP port of the bin-based algorithm), not GMQL-Spark (no JVM in the image)"},
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--regions", type=int, default=50000)
    ap.add_argument("--cpu-chroms", type=int, default=4, help="CPU arm: number of smallest chromosomes in the bounded slice")
    ap.add_argument("--cpu-samples", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--emulate-shard", default="", help="r/w: run rank r's chromosome shard of a w-GPU job on one GPU (tuning aid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
