"""columns -> text on the device (gmql_format_text) against the oracle's restatement of StoreTABRD's line construction
(SelectRegions/StoreTABRD.scala:62-80) and of GDouble.toString = DecimalFormat("#.########") with RoundingMode.FLOOR
(GMQL-Core/.../DataTypes.scala:139-150).  Byte for byte."""
import math
import random
import struct

import numpy as np
import pytest

from oracle import gmql_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


def _nextafter(x, up, k=1):
    for _ in range(k):
        x = math.nextafter(x, math.inf if up else -math.inf)
    return x


def _doubles(rng, n):
    """What the formatter could get wrong: values on, just below and just above multiples of 1e-8 (FLOOR on the shortest decimal,
    not on the binary value), negatives (FLOOR is towards minus infinity, not truncation), tails below 1e-8, -0.0, integers,
    subnormals, powers of two around the 2^26 / 2^53 hand-over points."""
    out = [0.0, -0.0, 0.3, -0.3, 0.1 + 0.2, 1e-9, -1e-9, 1e-8, -1e-8, 0.99999999, 0.999999999, -0.999999999, 5e-324, -5e-324, 2.2250738585072014e-308,
           1.0, -1.0, 123456.0, -98765432.0, 2.0 ** 26 - 0.5, 2.0 ** 26 + 0.5, 2.0 ** 52 + 0.5, 2.0 ** 53 - 1, 2.0 ** 53, -2.0 ** 53 - 2, 1e15, 1e16, 1.5e300,
           67108863.99999999, -67108863.99999999, 0.000000015, 0.00000001500000001, 12.345678905, 12.345678915, 1 / 3, -2 / 3, 1e-7, 123.456]
    while len(out) < n:
        kind = rng.randrange(8)
        if kind == 0:
            v = rng.uniform(-1000, 1000)
        elif kind == 1:
            v = rng.randrange(-10 ** 9, 10 ** 9) / 1e8                 # a multiple of 1e-8 (as close as a double gets)
        elif kind == 2:
            v = _nextafter(rng.randrange(-10 ** 9, 10 ** 9) / 1e8, rng.random() < 0.5, rng.randrange(1, 4))
        elif kind == 3:
            v = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
            if v != v or math.isinf(v):
                continue
        elif kind == 4:
            v = float(rng.randrange(-2 ** 40, 2 ** 40))
        elif kind == 5:
            v = rng.uniform(-1, 1) * 10.0 ** rng.randrange(-12, 9)
        elif kind == 6:
            v = round(rng.uniform(0, 100), rng.randrange(0, 10))       # typical score / p-value columns
        else:
            v = rng.randrange(0, 2 ** 26) + rng.randrange(0, 10 ** 8) / 1e8
        out.append(v)
    return out


def test_lines_equal_the_oracle_byte_for_byte(ex):
    from gmql_b200.regions import RegionDataset
    rng = random.Random(31)
    n = 6000
    vals = _doubles(rng, n)
    other = [None if rng.random() < 0.1 else rng.uniform(0, 50) for _ in range(n)]
    names = ["chr1", "chr10", "chrX", "chrUn_KI270742v1"]
    recs = []
    for i in range(n):
        st = rng.randrange(0, 2 ** 31 - 5000)
        recs.append((7 + i % 3, names[rng.randrange(4)], st, st + rng.randrange(0, 4000), "*+-"[rng.randrange(3)], (vals[i], other[i])))
    ds = RegionDataset.from_records(recs, sample_ids=[7, 8, 9])
    order = list(range(n))
    for one_based in (False, True):
        text, off, status = ex.format_text(ds, one_based=one_based, host_format=O.gdouble_to_string)
        want = "".join(O.store_tab_line(recs[i], one_based) + "\n" for i in order)
        assert text.decode() == want
        assert off[0] == 0 and off[-1] == len(text) and len(off) == n + 1
    # the raw device text: the hand-over to the host is exactly the documented set (placeholder byte, status bit)
    raw, off, status = ex.format_text(ds)
    for k, i in enumerate(order):
        v = vals[i]
        big = (abs(v) >= 2.0 ** 53) or (v != math.floor(v) and abs(v) >= 2.0 ** 26)
        assert bool(status[k] & 1) == big, v
        assert (b"\x1a" in raw[off[k]:off[k + 1]]) == big


def test_null_only_and_no_value_columns(ex):
    from gmql_b200.regions import RegionDataset
    recs = [(1, "chr2", 10, 20, "+", (None,)), (1, "chr2", 0, 0, "*", (None,)), (1, "chr1", 5, 6, "-", (-0.0,))]
    ds = RegionDataset.from_records(recs, sample_ids=[1])
    text, off, status = ex.format_text(ds)
    assert text.decode() == "chr2\t10\t20\t+\tnull\nchr2\t0\t0\t*\tnull\nchr1\t5\t6\t-\t-0\n"
    text, off, status = ex.format_text(ds, value_cols=[])
    assert text.decode() == "chr2\t10\t20\t+\nchr2\t0\t0\t*\nchr1\t5\t6\t-\n"
    empty = RegionDataset.from_records([], sample_ids=[1])
    text, off, status = ex.format_text(empty, value_cols=[])
    assert text == b"" and list(off) == [0]


def test_large_dataset_spans_many_blocks_and_reports_throughput(ex):
    """2e6 narrowPeak-like rows (the staging path of k_fmt_write: 256 rows per block through shared memory, 16-byte stores):
    sampled lines against the oracle, every line offset consistent, total length equal to the sum of the oracle's line lengths on
    a slice."""
    from gmql_b200 import synth
    ds = synth.peak_samples(40, 50000, 2)
    text, off, status = ex.format_text(ds)
    n = ds.n
    assert len(off) == n + 1 and off[-1] == len(text) and not status.any()
    recs = ds.to_records()
    rng = random.Random(5)
    for i in [0, 1, 255, 256, 257, n - 1] + [rng.randrange(n) for _ in range(3000)]:
        assert text[off[i]:off[i + 1]].decode() == O.store_tab_line(recs[i]) + "\n"
    assert sum(len(O.store_tab_line(recs[i])) + 1 for i in range(20000)) == off[20000]
