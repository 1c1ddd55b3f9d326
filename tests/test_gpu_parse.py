"""gmql_parse_text (device text -> columns) against the oracle's restatement of BedParser.region_parser and against Python's
float() bit for bit."""
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


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def _narrowpeak_lines(rng, n):
    lines = []
    for i in range(n):
        chrom = rng.choice(["chr1", "chr2", "chrX", "chr10_random"])
        s = rng.randrange(0, 10 ** 8)
        e = s + rng.randrange(0, 5000)
        strand = rng.choice(["+", "-", ".", "*", " + "])
        vals = []
        for k in range(5):
            r = rng.random()
            if r < 0.08:
                vals.append(rng.choice(["null", "NULL", ".", "Null"]))
            elif r < 0.5:
                vals.append(repr(rng.lognormvariate(1.5, 0.8)))
            elif r < 0.7:
                vals.append("%d" % rng.randrange(0, 1000))
            elif r < 0.85:
                vals.append("%.3e" % rng.uniform(-1e20, 1e20))
            else:
                vals.append(" %s " % repr(rng.uniform(-1, 1)))
        lines.append("\t".join([chrom if rng.random() > 0.05 else " " + chrom + " ", " %d" % s if rng.random() < 0.1 else str(s), str(e), "p%d" % i, vals[0], strand] + vals[1:]))
    return lines


def test_parse_narrowpeak_against_oracle(ex):
    rng = random.Random(5)
    lines = _narrowpeak_lines(rng, 4000)
    lines.insert(3, "# a comment line")
    lines.insert(10, "#chr1\tnot a number\t5")
    text = "\n".join(lines) + "\n"
    got = ex.parse_text(text, "NarrowPeakParser")
    want = [O.parse_region_line("NarrowPeakParser", 7, ln) for ln in lines]
    want = [w for w in want if w is not None]
    assert len(got["start"]) == len(want) == 4000
    assert not got["status"].any()
    names = got["chrom_names"]
    assert [names[c] for c in got["chrom"]] == [w[1] for w in want]
    assert got["start"].tolist() == [w[2] for w in want] and got["stop"].tolist() == [w[3] for w in want]
    assert ["*+-"[s] for s in got["strand"]] == [w[4] for w in want]
    # NarrowPeakParser values: name(STRING, skipped), score, signalValue, pValue, qValue, peak
    for k in range(5):
        v, ok = got["values"][k]
        exp = [w[5][k + 1] for w in want]
        assert ok.tolist() == [0 if x is None else 1 for x in exp]
        m = ok.astype(bool)
        assert np.array_equal(bits(v[m]), bits([x for x in exp if x is not None]))      # bit-identical to float()


def test_parse_double_bit_exact_random(ex):
    rng = random.Random(6)
    vals = []
    for _ in range(30000):
        x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        if x != x or x in (float("inf"), float("-inf")):
            continue
        vals.append(rng.choice([repr(x), "%.17g" % x, "%.12g" % x, "%.6f" % (x % 1e6)]))
    vals += ["0", "-0.0", "1e-400", "1e400", "4.9e-324", "2.4703282292062328e-324", "1.7976931348623157e308", "9007199254740993", "5e-324",
             "123456789012345678901234567890", "0.1000000000000000055511151231257827", "1.00000000000000011102230246251565404236316680908203125"]
    text = "".join("chr1\t1\t2\t%s\n" % v for v in vals)
    got = ex.parse_text(text, (0, 1, 2, None, [(3, "DOUBLE")]))
    v, ok = got["values"][0]
    decided = got["status"] == 0
    assert decided.sum() > len(vals) - 50                     # a handful of >19-digit ties are left to the host
    want = np.array([float(s) for s in vals])
    assert np.array_equal(bits(v[decided]), bits(want[decided]))
    assert ok[decided].all()


def test_parse_status_bits_and_layouts(ex):
    text = "\n".join([
        "chr1\t10\t20\t0.5",            # ok
        "chrZ\t10\t20\t0.5",            # unknown chromosome
        "chr1\tten\t20\t0.5",           # bad integer
        "chr1\t10",                       # missing columns
        "chr1\t10\t20\t0x1p3",          # hex float: host decides
        "chr1\t10\t20\tNaN",            # host decides
        "",                                # empty line: malformed in the reference (it throws)
        "chr1\t5\t6\t1e3\r",           # carriage return stripped
        "chr2\t7\t8\t.5",
    ])
    got = ex.parse_text(text, (0, 1, 2, None, [(3, "DOUBLE")]), chrom_names=["chr1", "chr2"])
    assert got["status"].tolist() == [0, 4, 2, 2, 1, 1, 2, 0, 0]
    assert got["chrom"].tolist()[:2] == [0, -1] and got["strand"].tolist() == [0] * 9
    assert got["values"][0][0][7] == 1000.0 and got["values"][0][0][8] == 0.5
    # 1-based schema, strand column, no values; no trailing newline
    got = ex.parse_text("chr1\t11\t20\t-\nchr1\t1\t5\t+", (0, 1, 2, 3, []), start_minus_one=True)
    assert got["start"].tolist() == [10, 0] and got["strand"].tolist() == [2, 1] and got["line_offset"].tolist() == [0, 13]


def test_parse_feeds_the_operators(ex):
    """text -> device columns -> COVER equals the oracle on the parsed records (ann-style layout)."""
    import gmql_b200 as G
    from tests.randgen import rand_regions, canon
    rng = random.Random(8)
    recs = rand_regions(rng, 800, [1], chroms=("chr1", "chr2"), strands="+-*", n_vals=1, null_p=0.0)
    text = "".join("%s\t%d\t%d\t%s\tg\t%r\n" % (r[1], r[2], r[3], r[4] if r[4] != "*" else ".", r[5][0]) for r in recs)
    got = ex.parse_text(text, "RnaSeqParser")
    assert not got["status"].any()
    ds = G.RegionDataset(got["chrom_names"], got["chrom"], got["start"], got["stop"], got["strand"], np.zeros(len(recs), np.int32), np.array([1], np.int64),
                         [("num", got["values"][0][0], None)])
    a = ex.GenometricCover("COVER", G.N(2), G.ANY(), [G.Agg("SUM", 0)], None, ds)
    want = O.cover_reference("COVER", O.CoverParam.N(2), O.CoverParam.ANY(), [O.Agg("SUM", 0)], [(r[0], r[1], r[2], r[3], r[4], (r[5][0],)) for r in recs], None, 2000)
    assert canon(a, 9) == canon(want, 9)
