#!/usr/bin/env python
"""Debug aid: repeats plain / error smoothing passes at size N and lists every value that differs from the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_poisson_solver_b200 as mg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lib = mg.init(0)
    orc = po.oracle_ops()
    rng = np.random.default_rng(5)
    U = rng.random((N, N)) - 0.25
    U[0, :] = U[-1, :] = 0
    U[:, 0] = U[:, -1] = 0
    U = U.reshape(-1)
    F = rng.random(N * N) * 3 - 1
    dU, dF, dO = mg.DeviceGrid(N, U), mg.DeviceGrid(N, F), mg.DeviceGrid(N)
    for s in (1, 2, 3):
        ref, _ = orc.doSmoothing(N, 1.0, U, F, s)
        W = 120
        for err in (0, 1):
            nbad = 0
            for it in range(reps):
                lib.mgGridZero(N, dO.ptr)
                lib.mgSmooth(N, 1.0, dU.ptr, dF.ptr, s, dO.ptr, lib.mgScalarSlot(5) if err else None)
                lib.mgSync()
                a = dO.numpy()
                bad = np.flatnonzero(a != ref)
                if bad.size:
                    nbad += 1
                    pts = [(int(b // N), int(b % N)) for b in bad[:40]]
                    print("S=%d err=%d rep %d: %d bad: %s" % (s, err, it, bad.size,
                          " ".join("(%d,%d|strip %d lane %d q %d|got %.3g want %.3g)" % (r, c, c // W, (c % W) // 4 + 1, c % 4, a[r * N + c], ref[r * N + c])
                                   for r, c in pts[:12])))
            print("S=%d err=%d: %d of %d repetitions wrong" % (s, err, nbad, reps))


if __name__ == "__main__":
    main()
