#!/usr/bin/env python
"""Debug aid (library built with -DMG_SP_DEBUG): what the copy ring of k_strip handed out compared with global memory."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_poisson_solver_b200 as mg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = mg.init(0)
rng = np.random.default_rng(5)
U = rng.random(N * N)
F = rng.random(N * N)
dU, dF, dO = mg.DeviceGrid(N, U), mg.DeviceGrid(N, F), mg.DeviceGrid(N)
out = (C.c_ulonglong * 64)()
lib_c = C.CDLL(mg.lib_path())
lib_c.mgStripDebug(out, 1)
for s in (1, 2, 3):
    for it in range(reps):
        lib.mgSmooth(N, 1.0, dU.ptr, dF.ptr, s, dO.ptr, None)
    lib.mgSync()
    lib_c.mgStripDebug(out, 1)
    v = list(out)
    print("S=%d: mismatches %d | value is: previous occupant %d, next occupant %d, other %d | re-read after 2us: now right %d, unchanged %d, other %d | F mismatches %d | "
          "first chunk %d later %d | slot k: %s | sample task %d r %d lane %d r_first %d r_end %d phase %d" % (
              s, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10:14], v[16], v[17], v[18], v[19], v[20], v[21]))
