"""Golden vectors on the reference's own in-repo fixture (GMQL-Repository/src/main/resources/ann/file{1,2}.bed), generated
by tests/golden/make_golden.py with the oracle.  CPU: the oracle still reproduces them (guards the checker itself).
GPU: the CUDA path, through the C ABI, equals them — MAP, the COVER family, JOIN, DIFFERENCE and the profiler fold."""
import json
import os

import pytest

from oracle import gmql_oracle as O
from tests.randgen import canon

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "ann_fixture.json")) as f:
        g = json.load(f)
    recs = {k: [(r[0], r[1], r[2], r[3], r[4], tuple(r[5])) for r in v] for k, v in g["inputs"].items()}
    return g, recs["file1.bed"], recs["file2.bed"], g["sample_ids"]["file1.bed"], g["sample_ids"]["file2.bed"]


def dec(rows):
    return canon([(r[0], r[1], r[2], r[3], r[4], tuple(r[5])) for r in rows], 9)


def test_fixture_ids_follow_the_loader(gold):
    g, f1, f2, id1, id2 = gold
    assert id1 == O.sample_id_from_filename("file1.bed") and id2 == O.sample_id_from_filename("file2.bed.meta")
    assert len(f1) == 500 and len(f2) == 501 and all(r[0] == id1 for r in f1)


def test_oracle_reproduces_golden(gold):
    g, f1, f2, id1, id2 = gold
    c = g["cases"]
    A = O.Agg
    assert canon(O.map_reference(f1, f2, {id1: [id2]}, [A("SUM", 1), A("MIN", 1), A("MAX", 1), A("AVG", 1)]), 9) == dec(c["map_count_sum_min_max_avg"])
    for flag in ("COVER", "FLAT", "SUMMIT", "HISTOGRAM"):
        assert canon(O.cover_reference(flag, O.CoverParam.N(2), O.CoverParam.ANY(), [A("SUM", 1)], f1 + f2, None, 2000), 9) == dec(c["cover_%s_2_any" % flag])
    pairs = {(id1, id2): O.md5_pair_id(id1, id2)}
    assert canon(O.join_reference(f1, f2, pairs, [("MD", 1)], "RIGHT"), 9) == dec(c["join_md1_right"])
    assert canon(O.difference_reference(f1, f2, {id1: [id2]}), 9) == dec(c["difference"])


@pytest.fixture(scope="module")
def ex():
    from gmql_b200 import GMQLCudaExecutor
    e = GMQLCudaExecutor()
    yield e
    e.close()


@pytest.mark.gpu
def test_cuda_map_matches_golden(gold, ex):
    from gmql_b200 import RegionDataset, Agg
    g, f1, f2, id1, id2 = gold
    d1, d2, both = RegionDataset.from_records(f1, [id1]), RegionDataset.from_records(f2, [id2]), RegionDataset.from_records(f1 + f2, [id1, id2])
    got = ex.GenometricMap71([(id1, id2)], [Agg("SUM", 1), Agg("MIN", 1), Agg("MAX", 1), Agg("AVG", 1)], d1, d2)
    assert canon(got, 9) == dec(g["cases"]["map_count_sum_min_max_avg"])
    got = ex.GenometricMap71([(id2, id1), (id2, id2)], [Agg("COUNT", 0)], d2, both)
    assert canon(got, 9) == dec(g["cases"]["map_self"])


@pytest.mark.gpu
@pytest.mark.parametrize("flag", ["COVER", "FLAT", "SUMMIT", "HISTOGRAM"])
def test_cuda_cover_matches_golden(gold, ex, flag):
    import gmql_b200 as G
    g, f1, f2, id1, id2 = gold
    both = G.RegionDataset.from_records(f1 + f2, [id1, id2])
    got = ex.GenometricCover(flag, G.N(2), G.ANY(), [G.Agg("SUM", 1)], None, both)
    assert canon(got, 9) == dec(g["cases"]["cover_%s_2_any" % flag])
    got = ex.GenometricCover(flag, G.N(1), G.N(2), [], None, both)
    assert canon(got, 9) == dec(g["cases"]["cover_%s_1_2" % flag])


@pytest.mark.gpu
def test_cuda_join_difference_profile_match_golden(gold, ex):
    import gmql_b200 as G
    g, f1, f2, id1, id2 = gold
    c = g["cases"]
    d1, d2, both = G.RegionDataset.from_records(f1, [id1]), G.RegionDataset.from_records(f2, [id2]), G.RegionDataset.from_records(f1 + f2, [id1, id2])
    assert canon(ex.GenometricJoin([(id1, id2)], [G.DistLess(21)], "LEFT", d1, d2), 9) == dec(c["join_dle20_left"])
    assert canon(ex.GenometricJoin([(id1, id2)], [G.MinDistance(1)], "RIGHT", d1, d2), 9) == dec(c["join_md1_right"])
    assert canon(ex.GenometricJoin([(id1, id2)], [G.DistLess(501), G.MinDistance(2), G.Upstream()], "CONTIG", d1, d2), 9) == dec(c["join_dle500_md2_up_contig"])
    assert canon(ex.GenometricDifference([(id1, id2)], d1, d2), 9) == dec(c["difference"])
    assert canon(ex.GenometricDifference([(id1, id2), (id2, id2)], both, d2, exact=True), 9) == dec(c["difference_exact"])
    assert {str(k): v for k, v in ex.profile(both).items()} == c["profile"]
