#!/usr/bin/env python
"""BASELINE config 5 (the feasible half): smoothing-only stress on row slabs, N up to 65536.

    python tools/smooth_stress.py                        # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/smooth_stress.py

1. check at N=4096: the slab result (owned rows) equals doSmoothing on one GPU bit for bit;
2. sweep step in {1, 3, 10, 100} at the stress size (65536 on >= 4 GPUs, 32768 on 2, 16384 on 1):
   ms per doSmoothing call, aggregate algorithmic GB/s (24 B/point/sweep) and per GPU.
The con_N=2 ladder half of config 5 needs odd level sizes on slabs, which the slab driver does
not support yet."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import multigrid_poisson_solver_b200 as mg  # noqa: E402


def stress(lib, N, step, reps, want_U=False):
    ms, err, lo, hi = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
    U = np.empty(N * N) if want_U else None
    rc = lib.mgDistSmoothStress(N, 1.0, step, reps, C.byref(ms), C.byref(err), U.ctypes.data if want_U else None, C.byref(lo), C.byref(hi))
    if rc != 0:
        raise SystemExit("mgDistSmoothStress failed: %d %s" % (rc, lib.mgLastError().decode()))
    return ms.value, err.value, (lo.value, hi.value), U


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = mg.init(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        mg.dist_init(rank, world, bcast)

    # 1. bit-identity of the slab path against doSmoothing on this GPU
    N = 4096
    ok = True
    for step in (3, 7):
        _, err, (lo, hi), U = stress(lib, N, step, 1, want_U=True)
        g = mg.GpuOps()
        ref, eref = g.doSmoothing(N, 1.0, np.zeros(N * N), g.getSource(N), step)
        same = np.array_equal(U[: (hi - lo) * N], ref[lo * N:hi * N])
        ok = ok and same and abs(err - eref) <= 1e-10 * eref
        if rank == 0:
            print("check N=%d step=%d rows[%d,%d) bit_identical=%s err=%.15g (single %.15g)" % (N, step, lo, hi, same, err, eref), flush=True)
    # 2. stress
    N = 65536 if world >= 4 else 32768 if world == 2 else 16384
    out = []
    for step in (1, 3, 10, 100):
        reps = 3 if step <= 10 else 1
        ms, err, _, _ = stress(lib, N, step, reps)
        gbs = 24.0 * N * N * step / (ms * 1e6)
        out.append(dict(N=N, world=world, step=step, ms=ms, error=err, algorithmic_GBs=gbs, per_gpu_GBs=gbs / world))
        if rank == 0:
            print(json.dumps(out[-1]), flush=True)
    if rank == 0:
        print("SMOOTH_STRESS", "PASS" if ok else "FAIL", flush=True)
    if world > 1:
        lib.mgDistShutdown()
        torch.distributed.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
