import sys
sys.path.insert(0, "/root/repo")
from oracle import gmql_oracle as O
from tests.randgen import canon
import gmql_b200 as G
from gmql_b200 import GMQLCudaExecutor, RegionDataset
left = [(1, "chr1", 1000, 1100, "+", ()), (1, "chr1", 1000, 1100, "-", ()), (1, "chr1", 1000, 1100, "*", ())]
right = [(10, "chr1", 800, 900, "*", ()), (10, "chr1", 1200, 1300, "*", ()), (10, "chr1", 1200, 1250, "+", ()),
         (10, "chr1", 1050, 1060, "-", ()), (10, "chr1", 500, 600, "+", ())]
pairs = {(1, 10): O.md5_pair_id(1, 10)}
for mode in ("lane", "team", "team-nomemo"):
    ex = GMQLCudaExecutor()
    ex.set_option("join.md", mode.split("-")[0])
    if mode.endswith("nomemo"): ex.set_option("join.md_memo", "0")
    l = RegionDataset.from_records(left, sample_ids=[1]); r = RegionDataset.from_records(right, sample_ids=[10])
    got = ex.GenometricJoin(list(pairs.keys()), [G.MinDistance(1)], "LEFT", l, r, None)
    want = O.join_reference(left, right, pairs, [("MD", 1)], "LEFT", 5000, 100000, None)
    print(mode, len(got), len(want), canon(got) == canon(want))
    ex.close()
