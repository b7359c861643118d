#!/usr/bin/env python
"""Times the fused legs on ODD grids (tile kernel) against the one-kernel-per-operator path (mgSetTileMaxN(0))."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import multigrid_poisson_solver_b200 as mg  # noqa: E402

lib = mg.init(0)
stream = torch.cuda.ExternalStream(lib.mgStream(), device=0)
for N in [int(a) for a in sys.argv[1:]] or [2049, 8193]:
    M = N - 1
    U, W, F, Fc, Uc = mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(M), mg.DeviceGrid(M)
    lib.getSource(N, 1.0, F.ptr, 0.0, 0.0)
    lib.getSource(M, 1.0, Uc.ptr, 0.0, 0.0)
    lib.mgGridZero(N, U.ptr)
    slot = lib.mgScalarSlot(100)

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        lib.mgSync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        lib.mgSync()
        return e0.elapsed_time(e1) / reps

    out = {"N": N, "M": M}
    for name, tile in (("tile", -1), ("per_operator", 0)):
        lib.mgSetTileMaxN(tile)
        out[name] = {"smooth3_ms": timeit(lambda: lib.mgSmooth(N, 1.0, U.ptr, F.ptr, 3, W.ptr, slot)),
                     "down3_ms": timeit(lambda: lib.mgDownLeg(N, 1.0, U.ptr, W.ptr, F.ptr, 3, 1, M, Fc.ptr, slot)),
                     "up3_ms": timeit(lambda: lib.mgUpLeg(M, Uc.ptr, N, 1.0, U.ptr, W.ptr, F.ptr, 3, slot))}
    lib.mgSetTileMaxN(-1)
    print(json.dumps(out))
