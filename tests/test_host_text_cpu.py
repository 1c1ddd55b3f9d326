"""Host-side text forms of the product mirror (gmql_b200/javafmt.py) against the oracle's restatements: GValue.toString as GROUP's
grouping fields see it (DecimalFormat("#.########"), RoundingMode.FLOOR; core/DataTypes.scala:139-150,185) and BAG / BAGD applied to a
whole list (DefaultRegionsToRegionFactory.scala:128-168 as GroupRD.scala:50 calls them)."""
import math
import random
import struct

from gmql_b200 import javafmt as J
from oracle import gmql_oracle as O


def test_gvalue_text_equals_the_oracle():
    rng = random.Random(12)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-9, -1e-9, 0.123456789, -0.123456789, 123456.000000019, 2.0 ** 53, -2.0 ** 60, 1e300, 5e-324,
            0.7657303808, 99999999.99999999, -99999999.99999999, float("inf"), float("-inf"), float("nan")]
    vals += [rng.uniform(-1e6, 1e6) for _ in range(2000)]
    vals += [rng.uniform(-1, 1) * 10.0 ** rng.randrange(-12, 12) for _ in range(2000)]
    vals += [struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0] for _ in range(2000)]
    for v in vals:
        assert J.gvalue_text(v) == O.gvalue_to_string(v), repr(v)
    assert J.gvalue_text(None) == O.gvalue_to_string(None) == "null"
    assert J.gvalue_text("geneA") == "geneA"


def test_bag_text_equals_the_oracle():
    rng = random.Random(3)
    pool = [None, "geneA", "geneB", "x", 1.5, -0.25, 1e7, 1e-4, 3.0, float("nan")]
    for _ in range(300):
        items = [rng.choice(pool) for _ in range(rng.randrange(0, 9))]
        items = [x for x in items if not (isinstance(x, float) and math.isnan(x))]      # NaN != NaN: `distinct` on it is not well defined
        for kind, distinct in (("BAG", False), ("BAGD", True)):
            assert J.bag_text(items, distinct) == O.Agg(kind, 0).fun(items)
