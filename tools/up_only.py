#!/usr/bin/env python
"""Runs mgUpLeg alone a few times (profiling target): python tools/up_only.py N reps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import multigrid_poisson_solver_b200 as mg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = mg.init(0)
stream = torch.cuda.ExternalStream(lib.mgStream(), device=0)
M = N // 2
U, W, F, Uc = mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(M)
lib.getSource(N, 1.0, F.ptr, 0.0, 0.0)
lib.getSource(M, 1.0, Uc.ptr, 0.0, 0.0)
lib.mgGridZero(N, U.ptr)
slot = lib.mgScalarSlot(100)
for _ in range(2):
    lib.mgUpLeg(M, Uc.ptr, N, 1.0, U.ptr, W.ptr, F.ptr, 3, slot)
lib.mgSync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    lib.mgUpLeg(M, Uc.ptr, N, 1.0, U.ptr, W.ptr, F.ptr, 3, slot)
e1.record(stream)
lib.mgSync()
print("up N=%d H=%s G=%s: %.4f ms" % (N, os.environ.get("MG_STREAM_H", "auto"), os.environ.get("MG_SCHED_G", "-"), e0.elapsed_time(e1) / reps))
