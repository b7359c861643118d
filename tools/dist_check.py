#!/usr/bin/env python
"""Multi-process check of the NCCL slab driver: run under torchrun with one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank also runs the same cycle alone on its own GPU and compares its owned rows of the
distributed solution with the single-GPU solution, bit for bit."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import multigrid_poisson_solver_b200 as mg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mg.init(local)

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    mg.dist_init(rank, world, bcast)
    ok = True
    cases = [("V 2048", mg.cycles.v_cycle(2048, 8), 256), ("V 4096", mg.cycles.v_cycle(4096, 8), 1024),
             ("W 1024", mg.cycles.w_cycle(1024, 8, levels=4, step=2, tol=1e-7), 256),
             ("trigger 1024", mg.cycles.v_cycle(1024, 8, step=-1), 256), ("V step5 1024", mg.cycles.v_cycle(1024, 16, step=5), 256),
             ("W 2048 x2 gathers", mg.cycles.w_cycle(2048, 8, levels=5, step=1, tol=1e-7), 1024), ("V 8192", mg.cycles.v_cycle(8192, 8), 1024)]
    for name, text, thr in cases:
        f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
        f.write(text)
        f.close()
        one = mg.run_cycle_host(f.name, mg.RUN_FUSED | mg.RUN_QUIET)
        d = mg.run_cycle_dist(f.name, thr, mg.RUN_FUSED | mg.RUN_QUIET, want_U=True)
        os.unlink(f.name)
        N = one["N"]
        lo, hi = d["own"]
        same_U = np.array_equal(d["U_own"], one["U"][lo * N:hi * N])
        errs = all(abs(a["err"] - b["err"]) <= 1e-10 * max(abs(b["err"]), 1e-300) and a["steps"] == b["steps"]
                   for a, b in zip(d["trace"], one["trace"]) if b["node"] != 0)
        mge = abs(d["mg_error"] - one["mg_error"]) <= 1e-10 * one["mg_error"]
        good = same_U and errs and mge and len(d["trace"]) == len(one["trace"])      # every rank holds the complete trace
        ok = ok and good
        print("rank %d %-14s rows [%d,%d) U_bit_identical=%s errors_ok=%s mg_error=%.12g (single %.12g) dist %.3f ms single %.3f ms"
              % (rank, name, lo, hi, same_U, errs, d["mg_error"], one["mg_error"], d["time_ms"], one["time_ms"]), flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    mg.lib().mgDistShutdown()
    dist.destroy_process_group()
    return 0 if int(t.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
