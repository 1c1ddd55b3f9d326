"""The decimal-to-double conversion used by gmql_parse_text (Eisel-Lemire, tools/gen_pow5_table.py holds the table generator
and a pure-Python statement of the algorithm) against Python's float(), which is correctly rounded like Java's
Double.parseDouble (the reference parses numeric fields with String.toDouble, BedParser.scala:30-48).  CPU only."""
import os
import random
import struct
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_pow5_table as G   # noqa: E402


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_table_header_is_current():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(here, "gmql_b200", "csrc", "pow5_table.h")).read()
    hi, lo = G.TABLE[0]
    assert "0x%016xull, 0x%016xull," % (hi, lo) in txt
    hi, lo = G.TABLE[-1]
    assert "0x%016xull, 0x%016xull," % (hi, lo) in txt and txt.count("ull,") == 2 * len(G.TABLE)


def test_conversion_matches_float():
    rng = random.Random(2026)
    undecided = 0
    for _ in range(60000):
        x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        if x != x or x in (float("inf"), float("-inf")):
            continue
        for s in (repr(x), "%.17g" % x, "%.15g" % x, "%.8f" % (x % 1000.0)):
            b = G.to_bits(s)
            if b is None:
                undecided += 1
            else:
                assert b == bits(float(s)), s
    for s in ["0", "-0.0", "1e-400", "1e400", "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "1.7976931348623157e308",
              "1.7976931348623159e308", "0.765730380838354", "9007199254740993", "9007199254740993e0", "2.2250738585072011e-308", "123456789012345678901234567890"]:
        b = G.to_bits(s)
        assert b is None or b == bits(float(s)), s
    assert undecided < 100
    for s in ["", ".", "1e", "0x1p3", "NaN", "Infinity", "1.0d", "1 2", "--1"]:
        assert G.to_bits(s) is None
