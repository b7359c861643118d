"""GPU parity at the sizes the headline numbers are made at (-m gpu): the CUDA path against the
CPU ORACLE -- not against itself -- for N = 2048 ... 16384, odd sizes above 1024, the weak-scaling
ladder sizes (1448, 2896), whole V / W / trigger cycles (BASELINE config 2 = N_max 8192 exactly) and
emulated row-slab runs.

From N ~ 2048 on the fused passes use the guided segment schedule (tasks >= resident warps), odd
sizes above 1024 leave the tile kernel, and slabs go through the halo logic: every one of those
paths is compared here with the oracle directly.  Bars as in test_gpu_parity.py: grids
bit-identical, reductions <= 1e-10 relative.
"""
import functools
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402  (checker only)

ERR_RTOL = 1e-10
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def gpu():
    import multigrid_poisson_solver_b200 as mg
    mg.init(0)
    return mg.GpuOps()


@pytest.fixture(scope="module")
def orc():
    import ctypes as C
    lib = po.oracle_lib()
    try:                                     # the oracle's sweeps are OpenMP loops: use the box's cores
        C.CDLL("libgomp.so.1").omp_set_num_threads(THREADS)
    except OSError:
        pass
    return po.Ops(lib, "orc_")


def grids(N, seed):
    rng = np.random.default_rng(seed)
    U = (rng.random((N, N)) - 0.25)
    U[0, :] = U[-1, :] = 0
    U[:, 0] = U[:, -1] = 0
    F = rng.random(N * N) * 3.0 - 1.0
    return U.reshape(-1), F


def assert_same(a, b, what=""):
    if not np.array_equal(a, b):
        bad = np.flatnonzero(a != b)
        rel = np.max(np.abs(a[bad] - b[bad]) / np.maximum(np.abs(b[bad]), 1e-300))
        raise AssertionError("%s: %d of %d values differ, first at %d, max rel %.3e" % (what, bad.size, a.size, bad[0], rel))


def oracle_down_leg(orc, N, L, U, F, step, zero_init, M):
    U0 = np.zeros(N * N) if zero_init else U
    Us, err = orc.doSmoothing(N, L, U0, F, step)
    D = -orc.getResidual(N, L, Us, F)
    return Us, err, orc.doRestriction(N, D, M)


# ------------------------------------------------------------------ operators at BASELINE sizes
@pytest.mark.parametrize("N,step", [(2048, 1), (2048, 3), (2048, 7), (4096, 1), (4096, 3), (4096, 7), (8192, 3), (8192, 7),
                                    (1025, 3), (2049, 1), (2049, 3), (2049, 7), (1448, 3), (2896, 3), (5792, 3)])
def test_smoothing_vs_oracle(N, step, gpu, orc):
    U, F = grids(N, 7 * N + step)
    a, ea = gpu.doSmoothing(N, 1.0, U, F, step)
    b, eb = orc.doSmoothing(N, 1.0, U, F, step)
    assert_same(a, b, "doSmoothing N=%d step=%d" % (N, step))
    assert ea == pytest.approx(eb, rel=ERR_RTOL)


@pytest.mark.parametrize("N,M,step", [(2048, 1024, 1), (2048, 1024, 3), (2048, 1024, 7), (4096, 2048, 3), (4096, 2048, 7),
                                      (8192, 4096, 3), (1025, 513, 3), (2049, 1025, 3), (2049, 1024, 2), (2050, 1025, 3),
                                      (2896, 1448, 3), (1448, 724, 3), (3000, 2999, 3), (2049, 2048, 2)])
def test_down_leg_vs_oracle(N, M, step, gpu, orc):
    U, F = grids(N, 31 * N + step)
    for zero_init in (True, False):
        a, ea, fa = gpu.down_leg(N, 1.0, U, F, step, zero_init, M)
        b, eb, fb = oracle_down_leg(orc, N, 1.0, U, F, step, zero_init, M)
        assert_same(a, b, "down_leg U N=%d step=%d zero=%s" % (N, step, zero_init))
        assert ea == pytest.approx(eb, rel=ERR_RTOL)
        assert_same(fa, fb, "down_leg F_c %d->%d step=%d zero=%s" % (N, M, step, zero_init))


@pytest.mark.parametrize("Nc,N,step", [(1024, 2048, 1), (1024, 2048, 3), (1024, 2048, 7), (2048, 4096, 3), (2048, 4096, 0),
                                       (4096, 8192, 3), (513, 1025, 3), (1025, 2049, 3), (1024, 2049, 2), (1025, 2050, 3),
                                       (1448, 2896, 3), (724, 1448, 3), (181, 362, 3), (2999, 3000, 3), (2048, 2049, 2)])
def test_up_leg_vs_oracle(Nc, N, step, gpu, orc):
    U, F = grids(N, 17 * N + step)
    Uc = np.random.default_rng(Nc).random(Nc * Nc) - 0.5
    a, ea = gpu.up_leg(Nc, Uc, N, 1.0, U, F, step)
    b = orc.doGridAddition(N, U, orc.doProlongation(Nc, Uc, N))
    if step > 0:
        b, eb = orc.doSmoothing(N, 1.0, b, F, step)
        assert ea == pytest.approx(eb, rel=ERR_RTOL)
    assert_same(a, b, "up_leg %d->%d step=%d" % (Nc, N, step))


# ------------------------------------------------------------------ whole cycles
def cycle_text(name):
    import multigrid_poisson_solver_b200 as mg
    c = mg.cycles
    return {
        "V2048": c.v_cycle(2048, 8),
        "V4096": c.v_cycle(4096, 8),
        "V8192": c.v_cycle(8192, 8),                                   # BASELINE config 2
        "V16384": c.v_cycle(16384, 8),                                 # the headline workload
        "W2048": c.w_cycle(2048, 8, tol=1e-8),                         # config 3's shape: full ladder, 128 exact solves
        "trigger4096": c.v_cycle(4096, 8, step=-1),                    # config 4's shape
        "V2049_minus1": c.v_cycle(2049, 2046, step=2, tol=1e3, con_N=2),   # config 5's ladder shape (N -> N-1), odd and even levels
        "V1025": c.v_cycle(1025, 9, step=3),                           # odd ladder 1025, 512, 256 ...
        "V2896": c.v_cycle(2896, 8),                                   # weak-scaling ladder through 181
    }[name]


@functools.lru_cache(maxsize=None)
def oracle_cycle(name):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    f.write(cycle_text(name))
    f.close()
    try:
        return po.run_cycle(f.name, threads=THREADS)
    finally:
        os.unlink(f.name)


def gpu_cycle(name, runner):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    f.write(cycle_text(name))
    f.close()
    try:
        return runner(f.name)
    finally:
        os.unlink(f.name)


def check_against_oracle(r, ref, what):
    assert [(t["node"], t["N"]) for t in r["trace"]] == [(t["node"], t["N"]) for t in ref["trace"]], what
    for mine, want in zip(r["trace"], ref["trace"]):
        if want["node"] != 0:
            assert mine["steps"] == want["steps"], (what, mine, want)
            assert mine["err"] == pytest.approx(want["err"], rel=ERR_RTOL), (what, mine, want)
    assert_same(r["U"], ref["U"], "final U of %s" % what)
    assert r["mg_error"] == pytest.approx(ref["mg_error"], rel=ERR_RTOL)


@pytest.mark.parametrize("name", ["V2048", "V4096", "V8192", "W2048", "trigger4096", "V2049_minus1", "V1025", "V2896"])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
def test_cycle_vs_oracle(name, mode, orc):
    """U bit-identical to the oracle when both start from the oracle's source grid (isolates the
    <= 2 ulp exp difference of getSource); errors <= 1e-10 relative; same nodes, sizes, sweep counts."""
    import multigrid_poisson_solver_b200 as mg
    if mode == "unfused" and name in ("V8192", "W2048"):
        pytest.skip("covered by the fused run; fused == unfused is asserted at this size in test_gpu_parity.py")
    ref = oracle_cycle(name)
    F = orc.getSource(ref["N"])
    flags = (mg.RUN_FUSED if mode == "fused" else mg.RUN_UNFUSED) | mg.RUN_QUIET
    r = gpu_cycle(name, lambda p: mg.run_cycle_host(p, flags, F_host=F))
    check_against_oracle(r, ref, "%s (%s)" % (name, mode))


@pytest.mark.slow
def test_headline_vcycle_16384_vs_oracle(orc):
    """The N = 16384 V-cycle of bench.py against the oracle (about a minute of CPU work, ~14 GB of host memory)."""
    import multigrid_poisson_solver_b200 as mg
    ref = oracle_cycle("V16384")
    F = orc.getSource(16384)
    r = gpu_cycle("V16384", lambda p: mg.run_cycle_host(p, mg.RUN_FUSED | mg.RUN_QUIET, F_host=F))
    check_against_oracle(r, ref, "V16384")
    oracle_cycle.cache_clear()


# ------------------------------------------------------------------ row slabs against the oracle
@pytest.mark.parametrize("name,world,threshold", [("V4096", 4, 512), ("V4096", 8, 256), ("V2896", 8, 300), ("trigger4096", 4, 1024),
                                                  ("W2048", 2, 256), ("V8192", 8, 1024)])
def test_slabs_vs_oracle(name, world, threshold, orc):
    """All ranks emulated on this GPU (the same arenas, peer stores, flags, gather and redundant coarse sub-cycles as
    the multi-process run), started from the oracle's source grid: U bit-identical to the oracle."""
    import multigrid_poisson_solver_b200 as mg
    ref = oracle_cycle(name)
    F = orc.getSource(ref["N"])
    r = gpu_cycle(name, lambda p: mg.run_cycle_dist_emulated(p, world, threshold, F_host=F))
    check_against_oracle(r, ref, "%s on %d emulated slabs" % (name, world))
