#!/usr/bin/env python
"""Runs mgUpLeg alone a few times (profiling target, A/B of library variants with MG_LIB_NAME): python tools/up_only.py N reps"""
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
clk, pw = [], []
try:                                   # SM clock / power while the launches run (the 1 node is issue bound: clock sensitive)
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not e1.query():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
except Exception:
    pass
lib.mgSync()
clk.sort()
print("up N=%d lib=%s H=%s: %.4f ms  sm_mhz median %s min %s  power max %s W" % (
    N, os.environ.get("MG_LIB_NAME", "libmgb200.so"), os.environ.get("MG_STREAM_H", "auto"), e0.elapsed_time(e1) / reps,
    clk[len(clk) // 2] if clk else "?", clk[0] if clk else "?", ("%.0f" % max(pw)) if pw else "?"))
