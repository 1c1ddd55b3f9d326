"""Multi-GPU host logic: one process per GPU, the region operators shard with NO data-path collective.

north_star / SURVEY.md §8(e): COVER partitions by chromosome (each (group, strand, chrom) segment is independent —
GenometricCover.scala keys everything by (chrom, bin, group, strand), :51,98-108); MAP additionally by experiment sample
(every (ref sample, exp sample) output is independent, GenometricMap71.scala:82-83 keys by expSample); JOIN by
chromosome.  NCCL (or gloo in the CPU tests) is used only to gather the compacted result tables.
"""
import numpy as np

from .regions import RegionDataset


def partition_weights(weights, world_size):
    """Longest-processing-time packing of weighted items (chromosomes, samples) onto ranks; returns list of index lists."""
    order = np.argsort(-np.asarray(weights, dtype=np.float64), kind="stable")
    load = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for c in order:
        r = int(np.argmin(load))
        parts[r].append(int(c))
        load[r] += float(weights[c])
    return [sorted(p) for p in parts]


def chromosome_weights(ds):
    """regions per chromosome (the work of every operator here is linear in its regions)"""
    return np.bincount(ds.chrom, minlength=len(ds.chrom_names)).astype(np.float64)


def _take(ds, mask):
    cols = []
    for c in ds.columns:
        if c[0] == "str":
            cols.append(("str", c[1][mask]))
        else:
            cols.append(("num", c[1][mask], None if c[2] is None else c[2][mask]))
    return RegionDataset(ds.chrom_names, ds.chrom[mask], ds.start[mask], ds.stop[mask], ds.strand[mask], ds.sample[mask], ds.sample_ids, cols)


def shard_by_chromosome(ds, world_size, rank, parts=None):
    """Rows of `ds` whose chromosome belongs to `rank`.  `parts` must be identical on every rank (derive it from the
    full dataset, or from a shared weight table such as chromosome sizes)."""
    if parts is None:
        parts = partition_weights(chromosome_weights(ds), world_size)
    keep = np.zeros(len(ds.chrom_names), dtype=bool)
    keep[parts[rank]] = True
    return _take(ds, keep[ds.chrom])


def shard_by_sample(ds, world_size, rank, parts=None):
    """Rows of `ds` whose sample belongs to `rank` (MAP: experiment side).  Sample ids / indices are kept global."""
    if parts is None:
        parts = partition_weights(np.bincount(ds.sample, minlength=len(ds.sample_ids)), world_size)
    keep = np.zeros(len(ds.sample_ids), dtype=bool)
    keep[parts[rank]] = True
    return _take(ds, keep[ds.sample]), parts


def shard_pairs(pairs, exp_sample_ids, parts, rank):
    """(ref id, exp id) pairs whose experiment sample lives on `rank`."""
    mine = {int(exp_sample_ids[i]) for i in parts[rank]}
    return [(a, b) for (a, b) in pairs if b in mine]


# ---- coordinate-range sharding of COVER / FLAT (north_star: "by genomic coordinate range with halo handling") -----------------
TILE_BITS = 16      # gmql_cover_shard cuts the linear coordinate into tiles of 65 536 positions


def linear_layout(chrom_span, bin_size=2000):
    """coff[c] of the library's linear coordinate lin(chrom, pos) = coff[chrom] + pos (include/gmql_cuda.h, gmql_cover_shard):
    exclusive prefix sum of (chrom_span[c] + bin_size + 2); returns (coff, total)."""
    w = np.asarray(chrom_span, dtype=np.int64) + bin_size + 2
    coff = np.concatenate([[0], np.cumsum(w)[:-1]]).astype(np.int64)
    return coff, int(w.sum())


def plan_tile_cuts(ds, chrom_span, world_size, bin_size=2000):
    """Tile boundaries [t_0 = 0 < t_1 < ... < t_world = n_tiles] that give every rank about the same number of regions
    (by the tile of the region's start).  Must be computed from the FULL dataset (or an all-reduced tile histogram)."""
    # the one-tile halo only holds for regions shorter than a tile (gmql_cover_shard's own bound): say so here, before any rank has
    # uploaded anything, instead of GMQL_ERR_UNSUPPORTED from whichever shard happens to hold the long region
    if ds.n and int((np.asarray(ds.stop, dtype=np.int64) - np.asarray(ds.start, dtype=np.int64)).max()) >= (1 << TILE_BITS) - 2:
        raise ValueError("coordinate-range sharding needs every region shorter than %d bases (one-tile halo); shard this dataset by "
                         "chromosome instead (dist.shard_by_chromosome)" % ((1 << TILE_BITS) - 2))
    coff, total = linear_layout(chrom_span, bin_size)
    n_tiles = (total >> TILE_BITS) + 1
    tile = (coff[ds.chrom] + ds.start) >> TILE_BITS
    cum = np.cumsum(np.bincount(tile, minlength=n_tiles))
    cuts = [0]
    for r in range(1, world_size):
        t = int(np.searchsorted(cum, ds.n * r / world_size, side="left")) + 1
        cuts.append(min(max(t, cuts[-1]), n_tiles))
    cuts.append(n_tiles)
    return cuts


def shard_by_tiles(ds, chrom_span, first_tile, end_tile, bin_size=2000):
    """Rows a coordinate-range shard needs: the regions starting in its tiles plus the tile before (the halo: regions are
    shorter than a tile, so nothing further left can reach in)."""
    coff, _ = linear_layout(chrom_span, bin_size)
    tile = (coff[ds.chrom] + ds.start) >> TILE_BITS
    return _take(ds, (tile >= first_tile - 1) & (tile < end_tile))


def merge_cover_shards(shards, flat):
    """Stitches the per-shard tables of gmql_cover_shard (dicts of numpy columns, in shard order) into the result of the
    unsharded operator: rows without edge flags are final; runs that touch at a cut are one run — AccIndex = max, contributors
    = sum, startMin / endMax / startMax / endMin = min / max / max / min — finished with GMAP4's formulas
    (GMAP4.scala:68-93, the same expressions as the device's k_runs_finalize).  Only the edge rows are touched on the host."""
    out = {k: [] for k in ("chrom", "start", "stop", "acc", "j1", "j2")}
    pending = None      # an open run: [chrom, rstart, rstop, acc, cnt, smin, emax, smax, emin]

    def combine(p, q):
        p[2] = q[2]; p[3] = max(p[3], q[3])
        if q[4] > 0:
            if p[4] > 0:
                p[5] = min(p[5], q[5]); p[6] = max(p[6], q[6]); p[7] = max(p[7], q[7]); p[8] = min(p[8], q[8])
            else:
                p[5:9] = q[5:9]
            p[4] += q[4]

    def finish(p):
        chrom, cs, ce, acc, cnt, smin, emax, smax, emin = p
        if cnt <= 0:
            return                                              # dropped: nothing overlaps it (GMAP4.scala:33-34,48)
        os_, oe = (smin, emax) if flat else (cs, ce)
        den = emax - smin
        num2 = float(emin - smax) if cnt == 1 else (float(emin - smax) if smax < emin else 0.0)
        out["chrom"].append([chrom]); out["start"].append([os_]); out["stop"].append([oe]); out["acc"].append([acc])
        out["j1"].append([abs((float(oe) - float(os_)) / float(den)) if den != 0 else 0.0])
        out["j2"].append([abs(num2 / float(den)) if den != 0 else 0.0])

    for t in shards:
        edge = t["edge"]
        idx = np.nonzero(edge)[0]
        lo = 0
        for i in idx:                                           # final rows between edge rows pass through untouched
            if i > lo:
                if pending is not None:
                    finish(pending); pending = None
                for k in out:
                    out[k].append(t[k][lo:i])
            row = [int(t["chrom"][i]), int(t["rstart"][i]), int(t["rstop"][i]), int(t["acc"][i]), int(t["cnt"][i]),
                   int(t["smin"][i]), int(t["emax"][i]), int(t["smax"][i]), int(t["emin"][i])]
            if pending is not None and (int(edge[i]) & 1) and pending[0] == row[0] and pending[2] == row[1]:
                combine(pending, row)
            else:
                if pending is not None:
                    finish(pending)
                pending = row
            if not (int(edge[i]) & 2):
                finish(pending); pending = None
            lo = i + 1
        if lo < len(edge):
            if pending is not None:
                finish(pending); pending = None
            for k in out:
                out[k].append(t[k][lo:])
    if pending is not None:
        finish(pending)
    dt = dict(chrom=np.int32, start=np.int64, stop=np.int64, acc=np.int32, j1=np.float64, j2=np.float64)
    return {k: (np.concatenate([np.asarray(a, dtype=dt[k]) for a in v]) if v else np.zeros(0, dt[k])) for k, v in out.items()}


def gather_table(columns, dst=0, group=None):
    """Gathers a dict of equal-length 1-D numpy columns from every rank onto `dst` (concatenated in rank order).
    Uses the default process group: NCCL moves device tensors over NVLink, gloo (CPU tests) host tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    names = sorted(columns)
    n = len(next(iter(columns.values()))) if columns else 0
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_h = counts.cpu().numpy()
    mx = int(counts_h.max()) if world else 0
    out = {}
    for name in names:
        a = np.ascontiguousarray(columns[name])
        t = torch.zeros(mx, dtype=torch.from_numpy(a[:0]).dtype, device=dev)
        if n:
            t[:n] = torch.from_numpy(a).to(dev)
        buf = torch.zeros(world * mx, dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(buf, t, group=group)
        if rank == dst:
            h = buf.cpu().numpy().reshape(world, mx)
            out[name] = np.concatenate([h[r, : counts_h[r]] for r in range(world)]) if world else h.reshape(-1)
    return out if rank == dst else None
