#!/usr/bin/env python
"""Runs the BASELINE.json configurations other than the headline one as parity/timing cases.

single GPU :  python tools/config_runs.py
multi GPU  :  python -m torch.distributed.run --nproc-per-node G ... tools/config_runs.py --dist

config 2  V-cycle con_N=1, N_max=8192, fixed 3 sweeps, 1 GPU
config 3  W-cycle (Wcycle.txt recursion) at N_max=16384, full ladder to N=16, 1 GPU and slabs on 2 GPUs
config 4  error-trigger V-cycle (con_step=-1) at N_max=32768, slabs on 8 GPUs, coarse agglomeration
At these sizes the CPU oracle is out of reach (hours), so the checks are the size-independent
ones: fused driver == unfused driver bit for bit (1 GPU), slab result == single-GPU result bit
for bit where one GPU can hold the grid, and the final error against the analytic solution."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_poisson_solver_b200 as mg  # noqa: E402


def cyc(text):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    f.write(text)
    f.close()
    return f.name


def summary(r):
    smooth = [t for t in r["trace"] if t["node"] != 0]
    return dict(nodes=len(r["trace"]), sweeps=sum(max(t["steps"], 0) for t in smooth), device_ms=r["time_ms"], wall_ms=r["wall_ms"],
                launches=r["launches"], mg_error=r["mg_error"], last_error=smooth[-1]["err"] if smooth else None)


def single():
    mg.init(0)
    out = {}
    cases = {
        "config2_V_8192": mg.cycles.v_cycle(8192, 8),
        "config3_W_16384_to_16": mg.cycles.w_cycle(16384, 16, step=3, tol=1e-8),
        "config4shape_trigger_V_16384": mg.cycles.v_cycle(16384, 8, step=-1),
    }
    for name, text in cases.items():
        path = cyc(text)
        mg.run_cycle(path)                                     # warm-up (tables, pools)
        a = mg.run_cycle_host(path, mg.RUN_FUSED | mg.RUN_QUIET)
        b = mg.run_cycle_host(path, mg.RUN_UNFUSED | mg.RUN_QUIET)
        os.unlink(path)
        same = bool(np.array_equal(a["U"], b["U"]))
        steps_same = [t["steps"] for t in a["trace"]] == [t["steps"] for t in b["trace"]]
        errs = max(abs(x["err"] - y["err"]) / max(abs(y["err"]), 1e-300) for x, y in zip(a["trace"], b["trace"]) if y["node"] != 0)
        out[name] = dict(fused=summary(a), unfused=summary(b), U_bit_identical=same, steps_identical=steps_same, max_rel_error_diff=errs)
        print(name, json.dumps(out[name]), flush=True)
    return out


def distributed():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mg.init(local)

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    mg.dist_init(rank, world, bcast)
    cases = {}
    if world == 2:
        cases["config3_W_16384_to_16_2gpu"] = (mg.cycles.w_cycle(16384, 16, step=3, tol=1e-8), 2048, True)
    if world == 8:
        cases["config4_trigger_V_32768_8gpu"] = (mg.cycles.v_cycle(32768, 8, step=-1), 2048, False)
        cases["config3_W_16384_to_16_8gpu"] = (mg.cycles.w_cycle(16384, 16, step=3, tol=1e-8), 2048, True)
    for name, (text, thr, compare) in cases.items():
        path = cyc(text)
        mg.run_cycle_dist(path, thr)                           # warm-up
        d = mg.run_cycle_dist(path, thr, want_U=True)
        res = dict(world=world, dist=summary(d), steps=[t["steps"] for t in d["trace"] if t["node"] != 0][:40])
        if compare:                                            # one GPU can hold N=16384: compare the owned rows bit for bit
            one = mg.run_cycle_host(path, mg.RUN_FUSED | mg.RUN_QUIET)
            lo, hi = d["own"]
            N = one["N"]
            same = bool(np.array_equal(d["U_own"], one["U"][lo * N:hi * N]))
            t = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            res["U_bit_identical_all_ranks"] = bool(int(t.item()))
            res["single_gpu"] = summary(one)
        os.unlink(path)
        if rank == 0:
            print(name, json.dumps(res), flush=True)
    mg.lib().mgDistShutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dist", action="store_true")
    a = ap.parse_args()
    distributed() if a.dist else single()
