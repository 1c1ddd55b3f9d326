"""C5's pipeline (C = COVER(2, ANY) E; M = MAP() C E) sharded by GENOMIC COORDINATE RANGE with halo handling across the GPUs of
one node — north_star's second sharding mode, on hardware (the bench shards by chromosome):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/range_shard_pipeline.py [samples regions]

  * every rank plans the same cuts from the tile histogram (dist.plan_tile_cuts: equal region counts, any number of ranks);
  * a rank uploads the regions that start in its tiles plus the tile before (the halo) and runs gmql_cover_shard on them;
  * the per-shard tables travel over NCCL (dist.gather_table -> all ranks); runs that touch at a cut are fused by
    dist.merge_cover_shards — the reference's boundary merge, GenometricCover.scala:121-152, at shard granularity;
  * MAP: a rank counts ITS OWN regions (halo excluded: every region is owned by exactly one rank) onto the merged runs; a run
    near a cut receives counts from both sides, so the count table is summed over the ranks with one NCCL all-reduce
    (GenometricMap71.scala:110-147's reduceByKey over the bins of a region, at shard granularity).
Rank 0 then computes the unsharded pipeline on one GPU and compares both tables row for row.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ANYV = 2 ** 31 - 1


def main():
    import torch
    import torch.distributed as dist
    from gmql_b200 import GMQLCudaExecutor, synth, dist as D, _lib as L
    from gmql_b200.regions import RegionDataset
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ns, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 50000)
    ds = synth.peak_samples(ns, nr, 5, with_values=())
    span = np.array([int(ds.stop[ds.chrom == c].max(initial=0)) for c in range(len(ds.chrom_names))], dtype=np.int64)
    cuts = D.plan_tile_cuts(ds, span, world)
    coff, _ = D.linear_layout(span)
    tile = (coff[ds.chrom] + ds.start) >> D.TILE_BITS
    part = D.shard_by_tiles(ds, span, cuts[rank], cuts[rank + 1])                         # own tiles + the halo tile
    own = D._take(ds, (tile >= cuts[rank]) & (tile < cuts[rank + 1]))                      # own tiles only (MAP's experiment side)
    ex = GMQLCudaExecutor(device=local)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    table = ex.cover_shard_raw("COVER", 2, ANYV, part, span, cuts[rank], cuts[rank + 1])
    t1 = time.perf_counter()
    # every rank needs the merged table (it is MAP's reference): gather onto each rank in turn is wasteful, all ranks merge
    tables = []
    gathered = {}
    for k in sorted(table):
        a = np.ascontiguousarray(table[k])
        n_loc = torch.tensor([len(a)], dtype=torch.int64, device="cuda")
        counts = torch.zeros(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(counts, n_loc)
        counts_h = counts.cpu().numpy()
        mx = int(counts_h.max())
        t = torch.zeros(mx, dtype=torch.from_numpy(a[:0]).dtype, device="cuda")
        t[: len(a)] = torch.from_numpy(a).cuda()
        buf = torch.zeros(world * mx, dtype=t.dtype, device="cuda")
        dist.all_gather_into_tensor(buf, t)
        h = buf.cpu().numpy().reshape(world, mx)
        gathered[k] = [h[r, : counts_h[r]] for r in range(world)]
    for r in range(world):
        tables.append({k: gathered[k][r] for k in gathered})
    merged = D.merge_cover_shards(tables, flat=False)
    t2 = time.perf_counter()
    n_runs = len(merged["start"])
    ref = RegionDataset(ds.chrom_names, merged["chrom"], merged["start"], merged["stop"], np.zeros(n_runs, np.int8), np.zeros(n_runs, np.int32), np.array([0], np.int64))
    c_ref, k1 = ref.as_c(sorted_by_position=True)
    c_ref.strand = None
    c_own, k2 = own.as_c(narrow=True)
    pairs = (L.Pair * ns)()
    for b in range(ns):
        pairs[b].left_sample, pairs[b].right_sample = 0, b
    mp = C.c_void_p()
    ex._check(ex.lib.gmql_map(ex.ctx, C.byref(c_ref), C.byref(c_own), pairs, ns, (L.Agg * 1)(), 0, C.byref(mp)))
    cnt_dev = torch.empty(ns * n_runs, dtype=torch.int32, device="cuda")
    src = ex.lib.gmql_result_column_device(mp, L.COL_MAP_COUNT)
    C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(cnt_dev.data_ptr()), C.c_void_p(src), C.c_size_t(4 * ns * n_runs), 3)
    ex.lib.gmql_result_free(ex.ctx, mp)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    dist.all_reduce(cnt_dev)                                                              # NCCL: partial counts of the runs near the cuts
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    counts = cnt_dev.cpu().numpy().reshape(ns, n_runs)
    loads = torch.tensor([part.n, own.n], dtype=torch.int64, device="cuda")
    all_loads = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(all_loads, loads)
    edge_rows = int(sum(int((t["edge"] != 0).sum()) for t in tables))
    ok_cover = ok_map = None
    if rank == 0:
        # the unsharded operators on one GPU: the oracle-pinned route of tests/test_gpu_resident.py
        host, keep = ds.as_c(narrow=True)
        dset, view = C.c_void_p(), L.Regions()
        ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
        mn, mxv = np.array([2], np.int32), np.array([ANYV], np.int32)
        cov = C.c_void_p()
        ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, 0, mn.ctypes.data, mxv.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(cov)))
        cview = L.Regions()
        ex._check(ex.lib.gmql_result_as_regions(ex.ctx, cov, C.byref(cview)))
        mp = C.c_void_p()
        ex._check(ex.lib.gmql_map(ex.ctx, C.byref(cview), C.byref(view), pairs, ns, (L.Agg * 1)(), 0, C.byref(mp)))

        def col(res, cid, dt):
            n = int(ex.lib.gmql_result_column_len(res, cid))
            a = np.empty(n, dtype=dt)
            if n:
                ex._check(ex.lib.gmql_result_column_copy(ex.ctx, res, cid, a.ctypes.data))
            return a
        want = {k: col(cov, c, dt) for k, c, dt in (("chrom", L.COL_COV_CHROM, np.int32), ("start", L.COL_COV_START, np.int64), ("stop", L.COL_COV_STOP, np.int64),
                                                     ("acc", L.COL_COV_ACC, np.int32), ("j1", L.COL_COV_J1, np.float64), ("j2", L.COL_COV_J2, np.float64))}
        want_cnt = col(mp, L.COL_MAP_COUNT, np.int32).reshape(ns, -1)
        ok_cover = all(len(want[k]) == len(merged[k]) and bool(np.array_equal(want[k], merged[k])) for k in want)
        ok_map = want_cnt.shape == counts.shape and bool(np.array_equal(want_cnt, counts))
        al = all_loads.cpu().numpy().reshape(world, 2)
        print(json.dumps({
            "what": "C5 pipeline sharded by genomic coordinate range with halo (gmql_cover_shard + NCCL gather + boundary-run merge; MAP partial counts all-reduced over NCCL)",
            "n_gpus": world, "samples": ns, "regions_per_sample": nr, "input_regions": int(ds.n), "tile_cuts": [int(c) for c in cuts],
            "rows_per_rank_with_halo": [int(x) for x in al[:, 0]], "own_rows_per_rank": [int(x) for x in al[:, 1]],
            "imbalance_own_rows": float(al[:, 1].max() / (ds.n / world)),
            "cover_runs": int(n_runs), "edge_rows_merged": edge_rows, "map_count_checksum": int(counts.astype(np.int64).sum()),
            "equals_unsharded_cover": ok_cover, "equals_unsharded_map": ok_map,
            "rank0_wall_ms": {"cover_shard (host columns in, H2D included)": round((t1 - t0) * 1e3, 2), "gather + merge": round((t2 - t1) * 1e3, 2),
                              "map of own rows (H2D included)": round((t3 - t2) * 1e3, 2), "all-reduce of the count table": round((t4 - t3) * 1e3, 2)}}), flush=True)
    ex.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not (ok_cover and ok_map):
        sys.exit(1)


if __name__ == "__main__":
    main()
