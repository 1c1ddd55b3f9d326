"""Generates tests/golden/ann_fixture.json from the reference's own in-repo fixture
(GMQL-Repository/src/main/resources/ann/file{1,2}.bed, schema test.schema: chr, start, stop, strand, name, score =
RnaSeqParser layout, loaders/BedParser.scala:346-348) and the oracle (oracle/gmql_oracle.py).

The reference has no expected outputs for these files, so the vectors pin the ORACLE on the reference's
own inputs: tests/test_golden.py checks that the oracle still reproduces them (CPU) and that the CUDA path equals them
(GPU).  Run in the build container only (reads /root/reference):  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gmql_oracle as O          # noqa: E402
from tests.randgen import canon              # noqa: E402

ANN = "/root/reference/GMQL-Repository/src/main/resources/ann"


def load(name):
    sid = O.sample_id_from_filename(name)
    recs = []
    with open(os.path.join(ANN, name)) as f:
        for line in f:
            r = O.parse_region_line("RnaSeqParser", sid, line.rstrip("\n"))
            if r is not None:
                recs.append(r)
    return sid, recs


def enc(rows):
    return [[r[0], r[1], r[2], r[3], r[4], [v[1] if v[0] != "n" else None for v in r[5]]] for r in canon(rows, 9)]


def main():
    id1, f1 = load("file1.bed")
    id2, f2 = load("file2.bed")
    both = f1 + f2
    out = {"inputs": {"file1.bed": [list(r[:5]) + [list(r[5])] for r in f1], "file2.bed": [list(r[:5]) + [list(r[5])] for r in f2]},
           "sample_ids": {"file1.bed": id1, "file2.bed": id2}, "cases": {}}
    A = O.Agg
    cases = out["cases"]
    cases["map_count_sum_min_max_avg"] = enc(O.map_reference(f1, f2, {id1: [id2]}, [A("SUM", 1), A("MIN", 1), A("MAX", 1), A("AVG", 1)]))
    cases["map_self"] = enc(O.map_reference(f2, both, {id2: [id1, id2]}, [A("COUNT")]))
    for flag in ("COVER", "FLAT", "SUMMIT", "HISTOGRAM"):
        cases["cover_%s_2_any" % flag] = enc(O.cover_reference(flag, O.CoverParam.N(2), O.CoverParam.ANY(), [A("SUM", 1)], both, None, 2000))
        cases["cover_%s_1_2" % flag] = enc(O.cover_reference(flag, O.CoverParam.N(1), O.CoverParam.N(2), [], both, None, 2000))
    pairs = {(id1, id2): O.md5_pair_id(id1, id2)}
    cases["join_dle20_left"] = enc(O.join_reference(f1, f2, pairs, [("DISTLESS", 21)], "LEFT"))
    cases["join_md1_right"] = enc(O.join_reference(f1, f2, pairs, [("MD", 1)], "RIGHT"))
    cases["join_dle500_md2_up_contig"] = enc(O.join_reference(f1, f2, pairs, [("DISTLESS", 501), ("MD", 2), ("UP", 0)], "CONTIG"))
    cases["difference"] = enc(O.difference_reference(f1, f2, {id1: [id2]}))
    cases["difference_exact"] = enc(O.difference_reference(both, f2, {id1: [id2], id2: [id2]}, exact=True))
    prof = O.profile_reference(both, [id1, id2])
    cases["profile"] = {str(k): v for k, v in prof.items()}
    with open(os.path.join(HERE, "ann_fixture.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: (len(v) if isinstance(v, list) else "dict") for k, v in cases.items()})


if __name__ == "__main__":
    main()
