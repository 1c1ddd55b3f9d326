"""N>1 host logic on CPU: world_size-2 gloo processes shard a dataset by chromosome / experiment sample, each computes
its shard (the oracle stands in for the GPU kernels here: there is no GPU in this container), the result tables are
gathered with the same gather_table the NCCL path uses, and rank 0 compares with the unsharded result."""
import os
import random
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import gmql_oracle as O
        from gmql_b200.regions import RegionDataset
        from gmql_b200 import dist as D
        from tests.randgen import rand_regions, canon
        rng = random.Random(123)
        ids = [1, 2, 3, 4]
        recs = rand_regions(rng, 400, ids, chroms=("chr1", "chr2", "chr3", "chrX", "chrY"), n_vals=1, strands="+-*")
        full = RegionDataset.from_records(recs, sample_ids=ids)
        # ---- COVER sharded by chromosome
        parts = D.partition_weights(D.chromosome_weights(full), world)
        assert sorted(sum(parts, [])) == list(range(len(full.chrom_names)))
        shard = D.shard_by_chromosome(full, world, rank, parts)
        # strand unification is a GLOBAL property of the input (assignGroups :328): decide it on the full dataset
        has_star = bool((full.strand == 0).any())
        local = O.cover_reference("COVER", O.CoverParam.N(2), O.CoverParam.ANY(), [O.Agg("SUM", 0)], shard.to_records() +
                                  ([(ids[0], "__none__", 0, 1, "*", (None,))] if has_star and not (shard.strand == 0).any() else []), None, 2000)
        local = [r for r in local if r[1] != "__none__"]
        names = full.chrom_names
        cols = {
            "chrom": np.array([names.index(r[1]) for r in local], dtype=np.int32),
            "start": np.array([r[2] for r in local], dtype=np.int64),
            "stop": np.array([r[3] for r in local], dtype=np.int64),
            "strand": np.array(["*+-".index(r[4]) for r in local], dtype=np.int8),
            "acc": np.array([r[5][0] for r in local], dtype=np.float64),
            "sum": np.array([np.nan if r[5][3] is None else r[5][3] for r in local], dtype=np.float64),
        }
        got = D.gather_table(cols)
        # ---- MAP sharded by experiment sample
        ref = rand_regions(rng, 40, [100], chroms=("chr1", "chr2"), n_vals=0)
        ref = list(dict.fromkeys(ref))
        exp_shard, sparts = D.shard_by_sample(full, world, rank)
        pairs = [(100, b) for b in ids]
        my_pairs = D.shard_pairs(pairs, full.sample_ids, sparts, rank)
        pd = {}
        for a, b in my_pairs:
            pd.setdefault(a, []).append(b)
        mloc = O.map_reference(ref, exp_shard.to_records(), pd, [O.Agg("MAX", 0)])
        mcols = {"id": np.array([r[0] for r in mloc], dtype=np.int64), "start": np.array([r[2] for r in mloc], dtype=np.int64),
                 "count": np.array([r[5][-2] for r in mloc], dtype=np.float64)}
        mgot = D.gather_table(mcols)
        if rank == 0:
            want = O.cover_reference("COVER", O.CoverParam.N(2), O.CoverParam.ANY(), [O.Agg("SUM", 0)], recs, None, 2000)
            g = sorted(zip(got["chrom"].tolist(), got["start"].tolist(), got["stop"].tolist(), got["strand"].tolist(), got["acc"].tolist(),
                           [None if v != v else round(v, 7) for v in got["sum"].tolist()]))
            w = sorted((names.index(r[1]), r[2], r[3], "*+-".index(r[4]), r[5][0], None if r[5][3] is None else round(r[5][3], 7)) for r in want)
            mwant = O.map_reference(ref, recs, {100: ids}, [O.Agg("MAX", 0)])
            mg = sorted(zip(mgot["id"].tolist(), mgot["start"].tolist(), mgot["count"].tolist()))
            mw = sorted((r[0], r[2], r[5][-2]) for r in mwant)
            q.put(("ok", g == w and len(g) > 0, mg == mw and len(mg) == len(ref) * len(ids)))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e), False))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res
    assert res[1], "sharded COVER != unsharded COVER"
    assert res[2], "sharded MAP != unsharded MAP"


def test_partition_is_balanced():
    from gmql_b200 import dist as D
    from gmql_b200.synth import HG38_SIZES
    for world in (2, 4, 8):
        parts = D.partition_weights(HG38_SIZES, world)
        loads = [HG38_SIZES[p].sum() for p in parts]
        assert sorted(sum(parts, [])) == list(range(24))
        assert max(loads) / (HG38_SIZES.sum() / world) < 1.12   # chromosome packing alone keeps the imbalance under 12 %


def test_range_shard_planning_and_edge_merge():
    """Host side of coordinate-range sharding (dist.plan_tile_cuts / shard_by_tiles / merge_cover_shards) on hand-made tables:
    a run crossing one cut, a run spanning a whole shard, an edge run whose contributors all live on the neighbour, and an
    edge run nobody overlaps (dropped, GMAP4.scala:33-34,48)."""
    import numpy as np
    from gmql_b200 import dist as D
    from gmql_b200.regions import RegionDataset
    rng = np.random.Generator(np.random.PCG64(1))
    n = 20000
    span = np.array([900000, 500000], dtype=np.int64)
    chrom = rng.integers(0, 2, n).astype(np.int32)
    start = (rng.random(n) * (span[chrom] - 1000)).astype(np.int64)
    ds = RegionDataset(["chr1", "chr2"], chrom, start, start + 500, np.zeros(n, np.int8), np.zeros(n, np.int32), np.array([0], np.int64))
    coff, total = D.linear_layout(span)
    assert coff.tolist() == [0, 900000 + 2002] and total == 1400000 + 2 * 2002
    cuts = D.plan_tile_cuts(ds, span, 4)
    assert cuts[0] == 0 and cuts[-1] == (total >> 16) + 1 and cuts == sorted(cuts)
    own = [int((((coff[ds.chrom] + ds.start) >> 16 >= cuts[r]) & ((coff[ds.chrom] + ds.start) >> 16 < cuts[r + 1])).sum()) for r in range(4)]
    assert sum(own) == n and max(own) < 1.6 * n / 4
    part = D.shard_by_tiles(ds, span, cuts[1], cuts[2])
    t = (coff[part.chrom] + part.start) >> 16
    assert int(t.min()) >= cuts[1] - 1 and int(t.max()) < cuts[2] and part.n >= own[1]

    def table(rows):
        keys = ("chrom", "start", "stop", "acc", "j1", "j2", "edge", "cnt", "rstart", "rstop", "smin", "emax", "smax", "emin")
        dt = dict(chrom=np.int32, acc=np.int32, edge=np.int32, cnt=np.int32, j1=np.float64, j2=np.float64)
        return {k: np.array([r[i] for r in rows], dtype=dt.get(k, np.int64)) for i, k in enumerate(keys)}

    s0 = table([(0, 10, 20, 2, 0.5, 0.25, 0, 3, 10, 20, 5, 30, 12, 18),                 # final row
                (0, 100, 200, 3, 0, 0, 2, 2, 100, 200, 90, 260, 150, 190)])             # reaches the cut at 200
    s1 = table([(0, 200, 300, 5, 0, 0, 3, 0, 200, 300, -1, -1, -1, -1)])                # spans the whole shard, no contributor here
    s2 = table([(0, 300, 350, 4, 0, 0, 1, 4, 300, 350, 280, 400, 340, 310),             # ends inside
                (0, 500, 600, 2, 0, 0, 2, 0, 500, 600, -1, -1, -1, -1),                 # open to the right, nobody overlaps it at all
                ])
    s3 = table([(0, 600, 640, 2, 0, 0, 1, 0, 600, 640, -1, -1, -1, -1),
                (1, 5, 9, 7, 1.0, 1.0, 0, 1, 5, 9, 5, 9, 5, 9)])
    for flat in (False, True):
        m = D.merge_cover_shards([s0, s1, s2, s3], flat)
        assert m["chrom"].tolist() == [0, 0, 1]
        # merged run 100..350: acc 5, cnt 6, smin 90, emax 400, smax 340, emin 190 -> (startMax >= endMin: zeroed numerator)
        assert m["acc"].tolist() == [2, 5, 7]
        assert (m["start"][1], m["stop"][1]) == ((90, 400) if flat else (100, 350))
        den = 400 - 90
        assert m["j1"][1] == abs((float(m["stop"][1]) - float(m["start"][1])) / den) and m["j2"][1] == 0.0
        assert m["j1"][0] == 0.5 and m["j2"][2] == 1.0


def test_range_shard_planner_refuses_long_regions():
    """the one-tile halo of coordinate-range sharding only holds for regions shorter than a tile; the planner says so
    on the full dataset, before any rank uploads anything, and points at chromosome sharding."""
    import numpy as np
    import pytest
    from gmql_b200 import dist as D
    from gmql_b200.regions import RegionDataset
    span = np.array([900000], dtype=np.int64)
    start = np.array([10, 5000, 200000], dtype=np.int64)
    ok = RegionDataset(["chr1"], np.zeros(3, np.int32), start, start + np.array([500, 65533, 10]), np.zeros(3, np.int8), np.zeros(3, np.int32), np.array([0], np.int64))
    assert D.plan_tile_cuts(ok, span, 2)[0] == 0
    long_ = RegionDataset(["chr1"], np.zeros(3, np.int32), start, start + np.array([500, 65534, 10]), np.zeros(3, np.int8), np.zeros(3, np.int32), np.array([0], np.int64))
    with pytest.raises(ValueError, match="shard_by_chromosome"):
        D.plan_tile_cuts(long_, span, 2)
