"""Profiling aid: time of the word-count tile kernel on the C5 input when it is cut short after a phase (option "cover.dbg";
results are wrong then — only the kernel times mean anything).  usage: python tools/tile_phase_probe.py [samples regions]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmql_b200 import GMQLCudaExecutor, synth, _lib as L
ns, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 50000)
ds = synth.peak_samples(ns, nr, 5, with_values=())
ex = GMQLCudaExecutor()
host, keep = ds.as_c(narrow=True)
host.strand = None
dset, view = C.c_void_p(), L.Regions()
ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
mn, mx = np.array([2], np.int32), np.array([2 ** 31 - 1], np.int32)
for dbg in (b"", b"8", b"9"):
    ex.lib.gmql_ctx_set_option(ex.ctx, b"cover.dbg", dbg or None)
    ms = []
    for rep in range(3):
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
        out = C.c_void_p()
        rc = ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, 0, mn.ctypes.data, mx.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out))
        buf = C.create_string_buffer(1 << 16)
        ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
        ex.lib.gmql_ctx_set_profiling(ex.ctx, 0)
        if rc == 0:
            ex.lib.gmql_result_free(ex.ctx, out)
        for ln in buf.value.decode().splitlines():
            if "k_cover_tile" in ln:
                ms.append(ln.rsplit(" ", 2))
    print("dbg=%s rc=%d" % (dbg.decode() or "-", rc), [(m[0][-30:], m[2]) for m in ms[-2:]], flush=True)
