"""ctypes wrapper of the C oracle port (oracle/gmql_oracle_c.c).  TEST INFRASTRUCTURE ONLY — see that file's header.
Works on the numpy columns of a RegionDataset-like object (chrom/start/stop/strand/sample arrays)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = C.CDLL(_LIB)
        _lib.oracle_map.restype = C.c_int64
        _lib.oracle_cover.restype = C.c_int64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def set_threads(n):
    """All host cores for the CPU arm of bench.py, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    lib().oracle_set_threads(C.c_int(int(n)))
    return num_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def map_rows(ref, exp, pairs_idx, val=None, valid=None, bin_size=50000):
    """pairs_idx: int32 array [n_pairs, 2] of (ref sample index, exp sample index).
    Returns dict(count, nn, sum, min, max) in the CUDA result's row layout."""
    L = lib()
    pl = np.ascontiguousarray(pairs_idx[:, 0], dtype=np.int32)
    pr = np.ascontiguousarray(pairs_idx[:, 1], dtype=np.int32)
    per = np.bincount(ref.sample, minlength=len(ref.sample_ids))
    n_rows = int(per[pl].sum())
    count = np.zeros(n_rows, np.int32)
    has = val is not None
    nn = np.zeros(n_rows, np.int32) if has else None
    s = np.zeros(n_rows, np.float64) if has else None
    mn = np.zeros(n_rows, np.float64) if has else None
    mx = np.zeros(n_rows, np.float64) if has else None
    r64s, r64e = ref.start.astype(np.int64), ref.stop.astype(np.int64)
    e64s, e64e = exp.start.astype(np.int64), exp.stop.astype(np.int64)
    v = np.ascontiguousarray(val, dtype=np.float64) if has else None
    vv = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    rc = L.oracle_map(C.c_int64(ref.n), _p(ref.chrom), _p(r64s), _p(r64e), _p(ref.strand), _p(ref.sample), C.c_int32(len(ref.sample_ids)),
                      C.c_int64(exp.n), _p(exp.chrom), _p(e64s), _p(e64e), _p(exp.strand), _p(exp.sample), C.c_int32(len(exp.sample_ids)),
                      _p(v), _p(vv), C.c_int64(len(pl)), _p(pl), _p(pr), C.c_int64(bin_size),
                      _p(count), _p(nn), _p(s), _p(mn), _p(mx))
    assert rc == n_rows, rc
    return dict(count=count, nn=nn, sum=s, min=mn, max=mx)


def cover_rows(ds, flag, mn, mx, sample_group=None, bin_size=2000, val=None, valid=None):
    """flag: 0 COVER, 1 FLAT, 2 SUMMIT, 3 HISTOGRAM; mn/mx: int32 per group.  Returns dict of numpy columns."""
    L = lib()
    mn = np.ascontiguousarray(mn, dtype=np.int32)
    mx = np.ascontiguousarray(mx, dtype=np.int32)
    sg = np.ascontiguousarray(sample_group, dtype=np.int32) if sample_group is not None else None
    s64, e64 = ds.start.astype(np.int64), ds.stop.astype(np.int64)
    v = np.ascontiguousarray(val, dtype=np.float64) if val is not None else None
    vv = np.ascontiguousarray(valid, dtype=np.uint8) if valid is not None else None
    n = L.oracle_cover(C.c_int64(ds.n), _p(ds.chrom), _p(s64), _p(e64), _p(ds.strand), _p(ds.sample), _p(sg), C.c_int32(len(mn)),
                       C.c_int32(flag), _p(mn), _p(mx), C.c_int64(bin_size), _p(v), _p(vv))
    assert n >= 0
    out = dict(group=np.zeros(n, np.int32), chrom=np.zeros(n, np.int32), start=np.zeros(n, np.int64), stop=np.zeros(n, np.int64),
               strand=np.zeros(n, np.int8), acc=np.zeros(n, np.int32), j1=np.zeros(n, np.float64), j2=np.zeros(n, np.float64),
               count=np.zeros(n, np.int32), nn=np.zeros(n, np.int32), sum=np.zeros(n, np.float64), min=np.zeros(n, np.float64), max=np.zeros(n, np.float64))
    L.oracle_cover_fetch(_p(out["group"]), _p(out["chrom"]), _p(out["start"]), _p(out["stop"]), _p(out["strand"]), _p(out["acc"]), _p(out["j1"]), _p(out["j2"]),
                         _p(out["count"]), _p(out["nn"]), _p(out["sum"]), _p(out["min"]), _p(out["max"]))
    return out
