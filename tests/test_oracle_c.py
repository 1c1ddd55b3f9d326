"""Pins the C oracle port (oracle/gmql_oracle_c.c) against the literal Python restatement on randomised inputs.  CPU only."""
import random

import numpy as np
import pytest

from oracle import gmql_oracle as O
from oracle import c_oracle as CO
from gmql_b200.regions import RegionDataset
from tests.randgen import rand_regions, canon


def _c_map_records(ref_recs, exp_recs, ref_ids, exp_ids, pairs, col):
    ref = RegionDataset.from_records(ref_recs, sample_ids=ref_ids)
    exp = RegionDataset.from_records(exp_recs, sample_ids=exp_ids)
    names = sorted(set(ref.chrom_names) | set(exp.chrom_names))
    ref, exp = ref.with_chrom_names(names), exp.with_chrom_names(names)
    pidx = np.array([(ref_ids.index(a), exp_ids.index(b)) for a, bs in pairs.items() for b in bs], dtype=np.int32)
    c = exp.columns[col]
    res = CO.map_rows(ref, exp, pidx, val=c[1], valid=c[2])
    out = []
    row = 0
    recs = ref.to_records()
    for (a, b) in pidx:
        rows = np.nonzero(ref.sample == a)[0]
        for r in rows:
            cnt, nn = int(res["count"][row]), int(res["nn"][row])
            vals = (float(cnt), res["sum"][row] if nn else None, (res["sum"][row] / nn) if nn else None,
                    res["min"][row] if nn else None, res["max"][row] if nn else None)
            rec = recs[r]
            out.append((O.md5_pair_id(ref_ids[a], exp_ids[b]), rec[1], rec[2], rec[3], rec[4], tuple(rec[5]) + vals))
            row += 1
    return out


@pytest.mark.parametrize("seed", range(8))
def test_c_map_equals_literal(seed):
    rng = random.Random(seed)
    ref = list(dict.fromkeys(rand_regions(rng, 60, [1, 2], n_vals=0, span=150000, max_len=70000, min_len=seed % 2)))
    exp = rand_regions(rng, 300, [10, 11, 12], n_vals=1, span=150000, max_len=40000, min_len=seed % 2)
    pairs = {1: [10, 11, 12], 2: [11]}
    aggs = [O.Agg("SUM", 0), O.Agg("AVG", 0), O.Agg("MIN", 0), O.Agg("MAX", 0)]
    want = O.map_reference(ref, exp, pairs, aggs, 50000)
    got = _c_map_records(ref, exp, [1, 2], [10, 11, 12], pairs, 0)
    assert canon(got, 7) == canon(want, 7)


FLAGC = {"COVER": 0, "FLAT": 1, "SUMMIT": 2, "HISTOGRAM": 3}


@pytest.mark.parametrize("flag", ["COVER", "FLAT", "SUMMIT", "HISTOGRAM"])
@pytest.mark.parametrize("seed", range(6))
def test_c_cover_equals_literal(flag, seed):
    rng = random.Random(100 + seed)
    ids = [1, 2, 3, 4]
    strands = "+-*" if seed % 2 else "+-"
    ds = rand_regions(rng, 220, ids, strands=strands, n_vals=1, min_len=(0 if seed == 5 else 1), max_len=(9000 if seed % 3 == 0 else 4500))
    groups = None if seed % 3 else {1: 7, 2: 7, 3: 8, 4: 8}
    d = RegionDataset.from_records(ds, sample_ids=ids)
    gids = sorted(set(groups.values())) if groups else [0]
    sg = np.array([gids.index(groups[s]) for s in ids], dtype=np.int32) if groups else None
    for (mn, mx) in ((1, 2), (2, 2 ** 31 - 1), (1, 2 ** 31 - 1), (3, 5)):
        omn = O.CoverParam.N(mn)
        omx = O.CoverParam.ANY() if mx > 10 ** 6 else O.CoverParam.N(mx)
        want = O.cover_reference(flag, omn, omx, [O.Agg("SUM", 0), O.Agg("MIN", 0), O.Agg("MAX", 0), O.Agg("COUNT", 0)], ds, groups, 2000)
        c = d.columns[0]
        r = CO.cover_rows(d, FLAGC[flag], [mn] * len(gids), [mx] * len(gids), sg, 2000, val=c[1], valid=c[2])
        got = []
        for i in range(len(r["start"])):
            nn = int(r["nn"][i])
            got.append((gids[r["group"][i]], d.chrom_names[r["chrom"][i]], int(r["start"][i]), int(r["stop"][i]), "*+-"[r["strand"][i]],
                        (float(r["acc"][i]), float(r["j1"][i]), float(r["j2"][i]), r["sum"][i] if nn else None,
                         r["min"][i] if nn else None, r["max"][i] if nn else None, float(r["count"][i]))))
        assert canon(got, 9) == canon(want, 9), (flag, mn, mx)
