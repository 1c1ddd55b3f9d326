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
