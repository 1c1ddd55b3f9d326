"""Small randomised region sets for oracle/CUDA parity tests (tuples in the oracle's record model)."""
import random


def rand_regions(rng, n, sample_ids, chroms=("chr1", "chr2"), span=12000, max_len=4500,
                 strands="+-*", n_vals=1, null_p=0.1, aligned_p=0.3, bin_mults=(2000, 5000, 50000),
                 min_len=1, with_name=False):
    out = []
    for _ in range(n):
        sid = rng.choice(sample_ids)
        chrom = rng.choice(chroms)
        start = rng.randrange(0, span)
        ln = rng.randrange(min_len, max_len)
        stop = start + ln
        if rng.random() < aligned_p:
            m = rng.choice(bin_mults)
            if rng.random() < 0.5:
                start = (start // m) * m
                stop = max(stop, start + min_len)
            else:
                stop = max((stop // m) * m, start + min_len)
        strand = rng.choice(strands)
        vals = []
        if with_name:
            vals.append(rng.choice(["geneA", "geneB", "geneC"]))
        for _k in range(n_vals):
            vals.append(None if rng.random() < null_p else float(rng.randrange(-50, 50)) / 4.0)
        out.append((sid, chrom, start, stop, strand, tuple(vals)))
    return out


def canon(rows, ndigits=9):
    """Multiset-comparable form: floats rounded, sorted."""
    def cv(v):
        if isinstance(v, float):
            return ("f", round(v, ndigits))
        if v is None:
            return ("n", 0)
        return ("s", v)
    return sorted((r[0], r[1], r[2], r[3], r[4], tuple(cv(v) for v in r[5])) for r in rows)
