"""The bin-free numpy restatement of JOIN [DLE,] MD(1) [, DOWN] that tests/test_gpu_fullsize.py uses at C4's full size
(_md1_reference) is itself pinned to the literal oracle (oracle/gmql_oracle.py: join_reference, the bin-replicated plan of
GenometricJoin.scala:62-218) on a small random input: same (left, right) pairs for pure MD, MD + DOWN and DLE + MD + DOWN."""
import numpy as np

from oracle import gmql_oracle as O
from tests.test_gpu_fullsize import _md1_reference


def test_numpy_md_reference_equals_the_literal_oracle():
    from gmql_b200.regions import RegionDataset
    rng = np.random.Generator(np.random.PCG64(5))
    names = ["chr%d" % i for i in range(24)]

    def mk(ns, n, strands, maxlen, fixed=None):
        chrom = rng.integers(0, 24, ns * n).astype(np.int32)
        start = rng.integers(0, 300000, ns * n).astype(np.int64)
        ln = np.full(ns * n, fixed) if fixed else rng.integers(50, maxlen, ns * n)
        strand = rng.choice(np.array(strands, dtype=np.int8), ns * n)
        return RegionDataset(names, chrom, start, start + ln, strand, np.repeat(np.arange(ns, dtype=np.int32), n), np.arange(100, 100 + ns, dtype=np.int64))

    left, right = mk(3, 150, [1, 2], 0, fixed=2000), mk(4, 500, [0], 5000)
    lrec, rrec = left.to_records(), right.to_records()
    pairs = {(int(a), int(b)): O.md5_pair_id(int(a), int(b)) for a in left.sample_ids for b in right.sample_ids}
    for spec, first_max in (([("MD", 1), ("DOWN", 0)], None), ([("DISTLESS", 10001), ("MD", 1), ("DOWN", 0)], 10001), ([("MD", 1)], None)):
        want = sorted(x[:5] for x in O.join_reference(lrec, rrec, pairs, spec, "RIGHT", 5000, 100000, None))
        keys = _md1_reference(left, right, np.arange(left.n), first_max, any(k == "DOWN" for k, _ in spec))
        got = sorted((pairs[(int(left.sample_ids[left.sample[a]]), int(right.sample_ids[right.sample[b]]))], names[right.chrom[b]],
                      int(right.start[b]), int(right.stop[b]), "*") for a, b in zip(keys // right.n, keys % right.n))
        assert len(want) > 400 and got == want
