"""CPU suite for the N>1 path: the host-side logic of the row-slab driver -- the slab geometry
plan every rank derives independently, and the NCCL-id rendezvous -- run with world_size 2 over
the gloo backend, plus invariants of the plan checked against the oracle's transfer maps."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import multigrid_poisson_solver_b200 as mg

HALO = 8   # mg_dist.cu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. every rank plans the geometry on its own; the plans must be identical
        ladder = mg.cycles.ladder(23168, 8)
        plan = mg.dist_plan(ladder, 8, 2048)
        gathered = [None] * world
        dist.all_gather_object(gathered, plan)
        assert all(g == gathered[0] for g in gathered)
        # 2. rendezvous: rank 0's 128-byte id reaches every rank unchanged
        payload = bytes(range(128)) if rank == 0 else None
        obj = [payload]
        dist.broadcast_object_list(obj, src=0)
        assert obj[0] == bytes(range(128))
        # 3. the product path has no CPU fallback: initialising the slab driver without a GPU fails loudly
        if not os.path.exists("/dev/nvidiactl"):
            with pytest.raises(mg.MGLibraryError):
                mg.dist_init(rank, world, lambda b: bytes(128))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_plan_and_rendezvous(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def fine_of_coarse(N, M):
    """floor map of doRestriction (MG_solver_CPU.cpp:661-662) with the boundary rows pinned."""
    h_f, h_c = 1.0 / (N - 1), 1.0 / (M - 1)
    f = [int(math.floor(c * h_c / h_f)) for c in range(M)]
    f[0], f[M - 1] = 0, N - 2
    return f


@pytest.mark.parametrize("N_max,world,threshold", [(16384, 8, 2048), (23168, 2, 1024), (46336, 8, 1024), (32768, 4, 1024), (4096, 3, 256),
                                                   (1024, 4, 128), (32768, 4, 4096)])
def test_plan_invariants(N_max, world, threshold):
    ladder = mg.cycles.ladder(N_max, 8)
    plan = mg.dist_plan(ladder, world, threshold)
    assert [lv["N"] for lv in plan] == ladder
    seen_agglomerated = False
    for fine, coarse in zip(plan, plan[1:]):
        N, M = fine["N"], coarse["N"]
        if not fine["dist"]:
            seen_agglomerated = True
            assert not coarse["dist"]                      # once agglomerated, always agglomerated
            continue
        assert not seen_agglomerated
        b = fine["bounds"]
        assert b[0] == 0 and b[-1] == N and all(y - x >= 2 * HALO for x, y in zip(b, b[1:]))
        foc = fine_of_coarse(N, M)
        assert all(y > x for x, y in zip(foc, foc[1:]))    # injective: a fused restriction exists
        if coarse["dist"]:
            cb = coarse["bounds"]
            assert cb[0] == 0 and cb[-1] == M
            for k in range(world):
                rows = range(cb[k], cb[k + 1])
                # a rank owns exactly the coarse rows whose lower fine row it owns: restriction is local
                assert all(b[k] <= foc[c] < b[k + 1] for c in rows)
                assert cb[k + 1] - cb[k] >= 2 * HALO
                # prolongation: the coarse cells of the fine rows a pass reads lie inside the coarse halo
                ratio = (1.0 / (M - 1)) / (1.0 / (N - 1))
                lo_cell = int(math.floor(max(b[k] - 5, 0) / ratio))
                hi_cell = int(math.floor(min(b[k + 1] + 5, N - 1) / ratio)) + 1
                assert lo_cell >= cb[k] - HALO and hi_cell < cb[k + 1] + HALO


def test_plan_agglomerates_odd_levels_and_rejects_unfusable_pairs():
    plan = mg.dist_plan([4098, 2049, 1024], 2, 1024)       # 2049 is odd: it and everything below run agglomerated
    assert [lv["dist"] for lv in plan] == [True, False, False]
    with pytest.raises(mg.MGLibraryError):
        mg.dist_plan([4098, 4097], 2, 1024)                # N -> N-1 from a distributed level: no fused transfer (ratio < 1.2)


def test_weak_scaling_sizes_are_servable():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for world, N in bench.WEAK_N.items():
        plan = mg.dist_plan(mg.cycles.ladder(N, 8), world, 1024)
        assert plan[0]["dist"] == (world > 1)
        assert abs(N * N / world / 16384 ** 2 - 1.0) < 0.01  # 16384^2 fine points per GPU
