#!/usr/bin/env python
"""Debug aid: one fused pass of every kind at size N against the oracle; prints where the grids differ."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_poisson_solver_b200 as mg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def report(name, a, b, N):
    bad = np.flatnonzero(a != b)
    if bad.size == 0:
        print("%-28s ok" % name)
        return
    rows, cols = bad // N, bad % N
    ur = np.unique(rows)
    print("%-28s %d bad values in %d rows; rows %s ; cols of first bad row: %d..%d (%d values)" % (
        name, bad.size, ur.size, list(ur[:24]), cols[rows == ur[0]].min(), cols[rows == ur[0]].max(), (rows == ur[0]).sum()))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    M = N // 2
    lib = mg.init(0)
    g, orc = mg.GpuOps(), po.oracle_ops()
    rng = np.random.default_rng(5)
    U = rng.random((N, N)) - 0.25
    U[0, :] = U[-1, :] = 0
    U[:, 0] = U[:, -1] = 0
    U = U.reshape(-1)
    F = rng.random(N * N) * 3 - 1
    Uc = rng.random(M * M) - 0.5
    import ctypes as C
    for s in (1, 2, 3):
        dU, dF, dO = mg.DeviceGrid(N, U), mg.DeviceGrid(N, F), mg.DeviceGrid(N)
        lib.mgSmooth(N, 1.0, dU.ptr, dF.ptr, s, dO.ptr, None)      # plain pass (no error)
        lib.mgSync()
        ref, _ = orc.doSmoothing(N, 1.0, U, F, s)
        report("plain S=%d" % s, dO.numpy(), ref, N)
        a, _ = g.smooth(N, 1.0, U, F, s)
        report("err   S=%d" % s, a, ref, N)
        for z in (1, 0):
            a, _, fa = g.down_leg(N, 1.0, U, F, s, z, M)
            b, _ = orc.doSmoothing(N, 1.0, np.zeros(N * N) if z else U, F, s)
            fb = orc.doRestriction(N, -orc.getResidual(N, 1.0, b, F), M)
            report("down  S=%d zero=%d U" % (s, z), a, b, N)
            report("down  S=%d zero=%d Fc" % (s, z), fa, fb, M)
        a, _ = g.up_leg(M, Uc, N, 1.0, U, F, s)
        b = orc.doGridAddition(N, U, orc.doProlongation(M, Uc, N))
        b, _ = orc.doSmoothing(N, 1.0, b, F, s)
        report("up    S=%d" % s, a, b, N)


if __name__ == "__main__":
    main()
