#!/usr/bin/env python
"""Times the fused passes alone at a given N with CUDA events on the library's stream:
smoothing pass (S sweeps + error), -1 node (mgDownLeg) and 1 node (mgUpLeg).
Prints achieved GB/s on the compulsory bytes of SURVEY.md 8(d)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import multigrid_poisson_solver_b200 as mg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--steps", type=int, nargs="*", default=[3])
    ap.add_argument("--warm", type=int, default=3, help="untimed calls per leg (0 for a profiler capture)")
    a = ap.parse_args()
    lib = mg.init(0)
    stream = torch.cuda.ExternalStream(lib.mgStream(), device=0)
    N, M = a.n, a.n // 2
    n, m = N * N, M * M
    U, W, F, Fc, Uc = mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(M), mg.DeviceGrid(M)
    lib.getSource(N, 1.0, F.ptr, 0.0, 0.0)
    lib.getSource(M, 1.0, Uc.ptr, 0.0, 0.0)
    lib.mgGridZero(N, U.ptr)
    slot = lib.mgScalarSlot(100)

    def timeit(fn):
        for _ in range(a.warm):
            fn()
        lib.mgSync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.reps):
            fn()
        e1.record(stream)
        lib.mgSync()
        return e0.elapsed_time(e1) / a.reps

    out = {"N": N, "H": os.environ.get("MG_STREAM_H", "auto")}
    for s in a.steps:
        ms = timeit(lambda: lib.mgSmooth(N, 1.0, U.ptr, F.ptr, s, W.ptr, slot))
        out["smooth_S%d" % s] = {"ms": ms, "GBs_compulsory": 24.0 * n / ms / 1e6, "GBs_unfused_equiv": (24.0 * s + 16) * n / ms / 1e6}
        ms = timeit(lambda: lib.mgDownLeg(N, 1.0, U.ptr, W.ptr, F.ptr, s, 1, M, Fc.ptr, slot))
        out["down_zero_S%d" % s] = {"ms": ms, "GBs_compulsory": (16.0 * n + 8.0 * m) / ms / 1e6}
        ms = timeit(lambda: lib.mgDownLeg(N, 1.0, U.ptr, W.ptr, F.ptr, s, 0, M, Fc.ptr, slot))
        out["down_load_S%d" % s] = {"ms": ms, "GBs_compulsory": (24.0 * n + 8.0 * m) / ms / 1e6}
        ms = timeit(lambda: lib.mgUpLeg(M, Uc.ptr, N, 1.0, U.ptr, W.ptr, F.ptr, s, slot))
        out["up_S%d" % s] = {"ms": ms, "GBs_compulsory": (24.0 * n + 8.0 * m) / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
