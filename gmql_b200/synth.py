"""Synthetic inputs for BASELINE.json's configs, generated straight into the SoA columns.

RNG: NumPy Generator(PCG64(seed)), seed = 20261000 + 100*config + sampleIndex, one generator per sample.  Rows are in
generation order (uniform random positions => unsorted within a sample, as the in-repo fixture is).  Lengths >= 1,
natural incidence of stop % 2000 == 0 kept.  hg38 primary chromosome sizes are an external constant.
"""
import numpy as np

from .regions import RegionDataset

HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
        ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]
HG38_NAMES = [c for c, _ in HG38]
HG38_SIZES = np.array([s for _, s in HG38], dtype=np.int64)


def _seed(config, sample_index):
    return 20261000 + 100 * config + sample_index


def _sample_ids(prefix, n):
    from .ids import sample_id_from_filename
    return np.array([sample_id_from_filename("%s_%05d.bed" % (prefix, i)) for i in range(n)], dtype=np.int64)


def peak_samples(n_samples, n_regions, config, with_values=("signalValue",), chrom_subset=None, prefix="peaks",
                 first_sample=0, sizes=None):
    """Appendix B "peak sample" (narrowPeak): chrom ~ size; len = clip(round(lognormal(ln 300, 0.6)), 50, 5000);
    start ~ U{0..size-len}; strand '.' -> '*'; signalValue ~ lognormal(1.5, 0.8).  chrom_subset restricts the
    chromosomes (multi-GPU sharding by chromosome: every rank draws the same stream and keeps its chromosomes)."""
    sizes = HG38_SIZES if sizes is None else np.asarray(sizes, dtype=np.int64)
    cum = np.cumsum(sizes / sizes.sum())
    keep = None
    if chrom_subset is not None:
        keep = np.zeros(len(sizes), dtype=bool)
        keep[list(chrom_subset)] = True
    chrom_l, start_l, stop_l, sample_l, val_l = [], [], [], [], []
    for s in range(n_samples):
        rng = np.random.Generator(np.random.PCG64(_seed(config, first_sample + s)))
        chrom = np.searchsorted(cum, rng.random(n_regions), side="right").astype(np.int32)
        np.minimum(chrom, len(sizes) - 1, out=chrom)
        ln = np.clip(np.rint(rng.lognormal(np.log(300.0), 0.6, n_regions)), 50, 5000).astype(np.int64)
        start = (rng.random(n_regions) * (sizes[chrom] - ln + 1)).astype(np.int64)
        sv = rng.lognormal(1.5, 0.8, n_regions) if with_values else None
        if keep is not None:
            m = keep[chrom]
            chrom, ln, start = chrom[m], ln[m], start[m]
            if sv is not None:
                sv = sv[m]
        chrom_l.append(chrom)
        start_l.append(start)
        stop_l.append(start + ln)
        sample_l.append(np.full(chrom.shape[0], s, dtype=np.int32))
        if sv is not None:
            val_l.append(sv)
    chrom = np.concatenate(chrom_l)
    cols = [("num", np.concatenate(val_l), None)] if with_values else []
    return RegionDataset(HG38_NAMES if sizes is HG38_SIZES or len(sizes) == len(HG38_NAMES) else ["chr%d" % (i + 1) for i in range(len(sizes))],
                         chrom, np.concatenate(start_l), np.concatenate(stop_l), np.zeros(chrom.shape[0], np.int8),
                         np.concatenate(sample_l), _sample_ids(prefix, n_samples), cols)


def promoter_samples(n_samples, n_regions, config, prefix="prom", length=2000):
    """Appendix B C3/C4 reference/left: `len = 2000`, strand +/-, one numeric score column."""
    sizes = HG38_SIZES
    cum = np.cumsum(sizes / sizes.sum())
    chrom_l, start_l, strand_l, sample_l, val_l = [], [], [], [], []
    for s in range(n_samples):
        rng = np.random.Generator(np.random.PCG64(_seed(config, 5000 + s)))
        chrom = np.minimum(np.searchsorted(cum, rng.random(n_regions), side="right"), len(sizes) - 1).astype(np.int32)
        start = (rng.random(n_regions) * (sizes[chrom] - length + 1)).astype(np.int64)
        chrom_l.append(chrom)
        start_l.append(start)
        strand_l.append(rng.integers(1, 3, n_regions).astype(np.int8))
        sample_l.append(np.full(n_regions, s, dtype=np.int32))
        val_l.append(rng.random(n_regions))
    start = np.concatenate(start_l)
    return RegionDataset(HG38_NAMES, np.concatenate(chrom_l), start, start + length, np.concatenate(strand_l), np.concatenate(sample_l),
                         _sample_ids(prefix, n_samples), [("num", np.concatenate(val_l), None)])


def c1_datasets(n_ref_per_chrom=500, n_exp_per_chrom=2000):
    """C1 (conf/test_map.xml shape): ref 1 RnaSeq-type sample, chr1..chr20 x 500, chromlen 100000, len U{20..500},
    strand {+,-,*} p=(.4,.4,.2); exp 10 BedScore samples, chr1..chr21 x 2000, score U[0,1)."""
    names = ["chr%d" % i for i in range(1, 22)]
    rng = np.random.Generator(np.random.PCG64(_seed(1, 0)))
    n = 20 * n_ref_per_chrom
    chrom = np.repeat(np.arange(20, dtype=np.int32), n_ref_per_chrom)
    ln = rng.integers(20, 501, n)
    start = (rng.random(n) * (100000 - ln + 1)).astype(np.int64)
    strand = rng.choice(np.array([1, 2, 0], dtype=np.int8), size=n, p=[0.4, 0.4, 0.2])
    perm = rng.permutation(n)
    ref = RegionDataset(names, chrom[perm], start[perm], (start + ln)[perm], strand[perm], np.zeros(n, np.int32), _sample_ids("c1ref", 1),
                        [("num", rng.random(n), None)])
    ec, es, ee, esm, ev = [], [], [], [], []
    for s in range(10):
        rng = np.random.Generator(np.random.PCG64(_seed(1, 1 + s)))
        m = 21 * n_exp_per_chrom
        c = np.repeat(np.arange(21, dtype=np.int32), n_exp_per_chrom)
        l2 = rng.integers(20, 501, m)
        st = (rng.random(m) * (100000 - l2 + 1)).astype(np.int64)
        perm = rng.permutation(m)
        ec.append(c[perm]); es.append(st[perm]); ee.append((st + l2)[perm]); esm.append(np.full(m, s, np.int32)); ev.append(rng.random(m))
    exp = RegionDataset(names, np.concatenate(ec), np.concatenate(es), np.concatenate(ee), np.zeros(10 * 21 * n_exp_per_chrom, np.int8),
                        np.concatenate(esm), _sample_ids("c1exp", 10), [("num", np.concatenate(ev), None)])
    return ref, exp


def chrom_partition(n_ranks, sizes=None):
    """Longest-processing-time packing of chromosomes onto ranks (north_star: partition naturally by chromosome)."""
    sizes = HG38_SIZES if sizes is None else np.asarray(sizes)
    order = np.argsort(-sizes)
    load = [0] * n_ranks
    parts = [[] for _ in range(n_ranks)]
    for c in order:
        r = int(np.argmin(load))
        parts[r].append(int(c))
        load[r] += int(sizes[c])
    return parts
