"""Per-kernel device times of ONE operator at BASELINE size (the context's own kernel profile: CUDA events around every launch):
    python tools/op_profile.py summit|histogram|cover|flat [samples regions]
A tuning aid: bench.py --workload c2 reports the four operators' kernels together."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from gmql_b200 import GMQLCudaExecutor, synth, _lib as L
    op = sys.argv[1] if len(sys.argv) > 1 else "histogram"
    ns, nr = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 50000)
    flag = {"cover": 0, "flat": 1, "summit": 2, "histogram": 3}[op]
    ds = synth.peak_samples(ns, nr, 2, with_values=())
    ex = GMQLCudaExecutor()
    host, keep = ds.as_c(narrow=True)
    dset, view = C.c_void_p(), L.Regions()
    ex._check(ex.lib.gmql_dataset_upload(ex.ctx, C.byref(host), C.byref(dset), C.byref(view)))
    a, b = np.array([2], np.int32), np.array([2 ** 31 - 1], np.int32)

    def run():
        out = C.c_void_p()
        ex._check(ex.lib.gmql_cover(ex.ctx, C.byref(view), None, 1, flag, a.ctypes.data, b.ctypes.data, 2000, (L.Agg * 1)(), 0, C.byref(out)))
        rows = int(ex.lib.gmql_result_rows(out))
        ms = ex.last_device_ms()
        ex.lib.gmql_result_free(ex.ctx, out)
        return rows, ms

    for _ in range(3):
        rows, ms = run()
    print(op, "rows", rows, "device ms", round(ms, 3))
    ex.lib.gmql_ctx_set_profiling(ex.ctx, 1)
    run()
    buf = C.create_string_buffer(1 << 16)
    ex.lib.gmql_ctx_profile_report(ex.ctx, buf, len(buf))
    tot = 0.0
    rowsl = []
    for ln in buf.value.decode().splitlines():
        name, cnt, t = ln.rsplit(" ", 2)
        rowsl.append((float(t), int(cnt), name))
        tot += float(t)
    for t, cnt, name in sorted(rowsl, reverse=True):
        print("%8.3f ms  x%-3d %s" % (t, cnt, name))
    print("sum of kernels %.3f ms" % tot)


if __name__ == "__main__":
    main()
