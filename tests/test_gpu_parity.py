"""GPU parity suite (-m gpu): every operator of libmgb200 is called through the C ABI and
compared with the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * grids (U, D, F_c, U_f): bit-identical to the oracle (np.array_equal; the tolerance the
    north star states is 1e-12 relative, the kernels are built to hit 0)
  * scalars that are order-dependent reductions (smoothing error, mg_error): <= 1e-10 relative
  * getSource / getAnalytic: <= 2 ulp (CUDA exp vs glibc exp)
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402  (checker only)

ERR_RTOL = 1e-10


@pytest.fixture(scope="module")
def gpu():
    import multigrid_poisson_solver_b200 as mg
    mg.init(0)
    return mg.GpuOps()


@pytest.fixture(scope="module")
def orc():
    return po.oracle_ops()


@pytest.fixture(autouse=True, params=["tile", "stream"])
def fused_kernel(request):
    """Every test runs twice: with every whole grid up to N = 1024 routed through the
    shared-memory tile kernel, and with that kernel switched off, so that the streaming kernel
    (even N) or the one-kernel-per-operator path (odd N) serves the same sizes.  All three must
    give the same bits.  (Default outside the tests: tile kernel for odd sizes only.)"""
    import multigrid_poisson_solver_b200 as mg
    lib = mg.init(0)
    lib.mgSetTileMaxN(1024 if request.param == "tile" else 0)
    yield request.param
    lib.mgSetTileMaxN(-1)


def grids(N, seed, zero_boundary=False):
    rng = np.random.default_rng(seed)
    U = rng.random(N * N) - 0.25
    F = rng.random(N * N) * 3.0 - 1.0
    if zero_boundary:
        U = U.reshape(N, N)
        U[0, :] = U[-1, :] = 0
        U[:, 0] = U[:, -1] = 0
        U = U.reshape(-1)
    return U, F


def assert_same(a, b, what=""):
    a = np.asarray(a); b = np.asarray(b)
    if not np.array_equal(a, b):
        bad = np.flatnonzero(a != b)
        rel = np.max(np.abs(a[bad] - b[bad]) / np.maximum(np.abs(b[bad]), 1e-300))
        raise AssertionError("%s: %d of %d values differ, first at %d, max rel %.3e" % (what, bad.size, a.size, bad[0], rel))


SIZES = [3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 64, 100, 255, 256, 257, 510, 1000, 1024]


@pytest.mark.parametrize("N", SIZES)
def test_residual_add_negate(N, gpu, orc):
    U, F = grids(N, N)
    for L in (1.0, 0.7):
        assert_same(gpu.getResidual(N, L, U, F), orc.getResidual(N, L, U, F), "getResidual N=%d" % N)
    assert_same(gpu.doGridAddition(N, U, F), orc.doGridAddition(N, U, F), "doGridAddition")
    assert_same(gpu.negate(N, U), -U, "negate")
    assert_same(gpu.getBoundary(N), np.zeros(N * N), "getBoundary")


@pytest.mark.parametrize("N", SIZES)
@pytest.mark.parametrize("step", [1, 2, 3, 4, 7])
def test_smoothing(N, step, gpu, orc):
    U, F = grids(N, 100 + N, zero_boundary=(N % 2 == 0))
    a, ea = gpu.doSmoothing(N, 1.0, U, F, step)
    b, eb = orc.doSmoothing(N, 1.0, U, F, step)
    assert_same(a, b, "doSmoothing N=%d step=%d" % (N, step))
    assert ea == pytest.approx(eb, rel=ERR_RTOL, abs=1e-300)


@pytest.mark.parametrize("N", [8, 17, 64, 255, 512])
def test_smoothing_other_domain_and_out_of_place(N, gpu, orc):
    U, F = grids(N, 5 + N)
    for L in (2.0, 0.3):
        a, ea = gpu.doSmoothing(N, L, U, F, 3)
        b, eb = orc.doSmoothing(N, L, U, F, 3)
        assert_same(a, b, "doSmoothing L=%g" % L)
        assert ea == pytest.approx(eb, rel=ERR_RTOL)
    for step in (1, 2, 3, 5):
        a, ea = gpu.smooth(N, 1.0, U, F, step)
        b, eb = orc.doSmoothing(N, 1.0, U, F, step)
        assert_same(a, b, "mgSmooth step=%d" % step)
        assert ea == pytest.approx(eb, rel=ERR_RTOL)


def test_smoothing_zero_steps_and_empty_interior(gpu, orc):
    U, F = grids(3, 1)
    a, ea = gpu.doSmoothing(3, 1.0, U, F, 2)
    b, eb = orc.doSmoothing(3, 1.0, U, F, 2)
    assert_same(a, b)
    assert ea == pytest.approx(eb, rel=ERR_RTOL)
    U, F = grids(16, 2)
    a, ea = gpu.doSmoothing(16, 1.0, U, F, 0)       # step 0: U untouched, error of the input
    b, eb = orc.doSmoothing(16, 1.0, U, F, 0)
    assert_same(a, b)
    assert ea == pytest.approx(eb, rel=ERR_RTOL)


def test_source_and_analytic(gpu, orc):
    for N in (8, 17, 256, 1000):
        for args in ((N,), (N, 2.0, -0.5, 0.25)):
            for name in ("getSource", "getAnalytic"):
                a = getattr(gpu, name)(*args)
                b = getattr(orc, name)(*args)
                ulp = np.spacing(np.maximum(np.abs(b), 1e-300))
                assert np.all(np.abs(a - b) <= (2 if name == "getSource" else 8) * ulp), name
                assert np.array_equal(a == 0, b == 0)


def transfer_pairs():
    pairs = []
    for N in list(range(3, 40)) + [63, 64, 65, 100, 127, 128, 129, 255, 256, 257, 511, 512, 1000, 1024]:
        for M in sorted({N // 2, N - 1, (N + 1) // 2, N // 3 + 2}):
            if 3 <= M < N:
                pairs.append((N, M))
    return pairs


def test_restriction_many_pairs(gpu, orc):
    rng = np.random.default_rng(11)
    for N, M in transfer_pairs():
        g = rng.random(N * N) - 0.5
        assert_same(gpu.doRestriction(N, g, M), orc.doRestriction(N, g, M), "doRestriction %d->%d" % (N, M))


def test_prolongation_many_pairs(gpu, orc):
    rng = np.random.default_rng(12)
    for M, N in transfer_pairs():           # coarse N, fine M
        g = rng.random(N * N) - 0.5
        assert_same(gpu.doProlongation(N, g, M), orc.doProlongation(N, g, M), "doProlongation %d->%d" % (N, M))
    for N in (3, 5, 17, 100, 256):
        g = rng.random(N * N) - 0.5
        for M in (2 * N, 2 * N - 1, 2 * N + 1, 3 * N - 2):
            assert_same(gpu.doProlongation(N, g, M), orc.doProlongation(N, g, M), "doProlongation %d->%d" % (N, M))


def test_transfer_linear_ramp_fixture(gpu):
    """testFunction/Test_doRestriction_GPU.cu:189-193 -- U[i+N*j] = i+j, N=16 -> M=8."""
    N, M = 16, 8
    ramp = np.add.outer(np.arange(N), np.arange(N)).astype(float).reshape(-1)
    c = gpu.doRestriction(N, ramp, M).reshape(M, M)
    k = np.arange(1, M - 1) * (N - 1) / (M - 1)
    assert np.allclose(c[1:-1, 1:-1], np.add.outer(k, k), rtol=1e-13)
    assert (c[0] == 0).all() and (c[:, 0] == 0).all() and (c[-1] == 0).all() and (c[:, -1] == 0).all()


@pytest.mark.parametrize("N,tol", [(3, 1e-7), (4, 1e-7), (8, 1e-7), (9, 1e-6), (16, 1e-7), (21, 1e-9), (32, 1e-8), (33, 1e-7),
                                   (64, 1e-5), (100, 1e-3), (130, 1e-2)])
def test_gauss_seidel_exact_solver(N, tol, gpu, orc):
    F = np.random.default_rng(N).random(N * N) - 0.3
    b = orc.doExactSolver(N, 1.0, F, tol, 1)
    iters = po.oracle_lib().orc_last_gs_iterations()
    a = gpu.doExactSolver(N, 1.0, F, tol, 1)
    assert gpu.last_exact_solver_iterations() == iters
    assert_same(a, b, "GaussSeidel N=%d" % N)


@pytest.mark.parametrize("N", [3, 4, 5, 8, 12])
def test_inverse_matrix_exact_solver(N, gpu, orc):
    F = np.random.default_rng(N).random(N * N) - 0.3
    assert_same(gpu.doExactSolver(N, 1.0, F, 0.0, 0), orc.doExactSolver(N, 1.0, F, 0.0, 0), "InverseMatrix N=%d" % N)


def oracle_down_leg(orc, N, L, U, F, step, zero_init, M):
    U0 = np.zeros(N * N) if zero_init else U
    if step > 0:
        Us, err = orc.doSmoothing(N, L, U0, F, step)
    else:
        Us, err = U0.copy(), None
    D = -orc.getResidual(N, L, Us, F)
    return Us, err, orc.doRestriction(N, D, M)


@pytest.mark.parametrize("N,M", [(8, 4), (16, 8), (17, 9), (33, 16), (64, 32), (100, 50), (255, 127), (256, 128), (257, 128),
                                 (512, 256), (1000, 500), (1024, 512), (45, 21), (40, 39)])
@pytest.mark.parametrize("step", [1, 2, 3, 4, 6])
def test_down_leg(N, M, step, gpu, orc):
    U, F = grids(N, 31 * N + step, zero_boundary=True)
    for zero_init in (True, False):
        a, ea, fa = gpu.down_leg(N, 1.0, U, F, step, zero_init, M)
        b, eb, fb = oracle_down_leg(orc, N, 1.0, U, F, step, zero_init, M)
        assert_same(a, b, "down_leg U N=%d step=%d zero=%s" % (N, step, zero_init))
        assert ea == pytest.approx(eb, rel=ERR_RTOL)
        assert_same(fa, fb, "down_leg F_c N=%d->%d step=%d zero=%s" % (N, M, step, zero_init))


def test_down_leg_without_sweeps(gpu, orc):
    N, M = 64, 32
    U, F = grids(N, 77)
    a, _, fa = gpu.down_leg(N, 1.0, U, F, 0, False, M)
    b, _, fb = oracle_down_leg(orc, N, 1.0, U, F, 0, False, M)
    assert_same(a, b)
    assert_same(fa, fb)


@pytest.mark.parametrize("Nc,N", [(4, 8), (8, 16), (9, 17), (16, 33), (32, 64), (50, 100), (127, 255), (128, 256), (128, 257),
                                  (256, 512), (500, 1000), (512, 1024), (21, 45), (39, 40)])
@pytest.mark.parametrize("step", [0, 1, 2, 3, 4, 6])
def test_up_leg(Nc, N, step, gpu, orc):
    U, F = grids(N, 17 * N + step)
    Uc = np.random.default_rng(Nc).random(Nc * Nc) - 0.5
    a, ea = gpu.up_leg(Nc, Uc, N, 1.0, U, F, step)
    b = orc.doGridAddition(N, U, orc.doProlongation(Nc, Uc, N))
    if step > 0:
        b, eb = orc.doSmoothing(N, 1.0, b, F, step)
        assert ea == pytest.approx(eb, rel=ERR_RTOL)
    assert_same(a, b, "up_leg %d->%d step=%d" % (Nc, N, step))


# ------------------------------------------------------------------ whole cycles
MODE_CYCLES = ["mode_nodestep_autoN", "mode_fixedstep_manualN", "mode_nodestep_minus1", "mode_manual_step0"]
CYCLES = ["test", "Vcycle", "VcycleTrigger", "Wcycle", "V_minus_one_ladder", "V_restart_x2", "manual_nonnested", "W_full_64",
          "V_lu_coarse", "V_offset_domain"] + MODE_CYCLES   # the last four: parser modes (0,!=0), (!=0,0), step 0 and negative steps


@pytest.fixture(scope="module")
def goldens(golden_dir):
    with open(os.path.join(golden_dir, "cycles.json")) as f:
        return json.load(f)


def header(path):
    t = open(path).read().split()
    return float(t[0]), float(t[1]), float(t[2]), int(t[5])


@pytest.mark.parametrize("name", CYCLES)
@pytest.mark.parametrize("mode", ["fused", "unfused"])
def test_cycle_against_reference_golden(name, mode, goldens, golden_dir, orc):
    """Same node/level sequence, same sweep counts, errors <= 1e-10 rel, U bit-identical when
    the source grid is the oracle's (isolates the <=1 ulp exp difference of getSource)."""
    import multigrid_poisson_solver_b200 as mg
    path = os.path.join(golden_dir, "cycle_%s.txt" % name)
    L, mx, my, N = header(path)
    flags = (mg.RUN_FUSED if mode == "fused" else mg.RUN_UNFUSED) | mg.RUN_QUIET
    r = mg.run_cycle_host(path, flags, F_host=orc.getSource(N, L, mx, my))
    g = goldens[name]
    assert [(t["node"], t["N"]) for t in r["trace"]] == [(t["node"], t["N"]) for t in g["trace"]]
    for mine, ref in zip(r["trace"], g["trace"]):
        if ref["node"] != 0:
            assert mine["steps"] == ref["steps"]
            if ref["steps"] != 0:            # a "1" node with step 0 evaluates no error (the oracle's record keeps the level's last one)
                assert mine["err"] == pytest.approx(ref["err"], rel=ERR_RTOL)
    assert r["mg_error"] == pytest.approx(g["mg_error"], rel=ERR_RTOL)
    upath = os.path.join(golden_dir, "cycle_U_%s.npy" % name)
    ref_U = np.load(upath) if os.path.exists(upath) else po.run_cycle(path)["U"]
    assert_same(r["U"], ref_U, "final U of %s (%s)" % (name, mode))


@pytest.mark.parametrize("name,n", [("Vcycle", 1), ("Vcycle", 5), ("Wcycle", 3), ("VcycleTrigger", 2)])
def test_host_batch_equals_single_calls(name, n, golden_dir, orc):
    """mgRunCycleFileHostBatch (double-buffered uploads / cycles / downloads) returns, problem by
    problem, the bits of mgRunCycleFileHost -- different sources so that a mixed-up buffer shows."""
    import multigrid_poisson_solver_b200 as mg
    path = os.path.join(golden_dir, "cycle_%s.txt" % name)
    L, mx, my, N = header(path)
    F0 = orc.getSource(N, L, mx, my)
    Fs = [F0 * (1.0 + 0.25 * k) for k in range(n)]
    flags = mg.RUN_FUSED | mg.RUN_QUIET
    batch = mg.run_cycle_host_batch(path, Fs, flags)
    for k in range(n):
        one = mg.run_cycle_host(path, flags, F_host=Fs[k])
        assert_same(batch["U"][k], one["U"], "problem %d of %d (%s)" % (k, n, name))
        assert batch["mg_error"][k] == one["mg_error"]


@pytest.mark.parametrize("name", ["Vcycle", "Wcycle", "VcycleTrigger"])
def test_cycle_with_device_source(name, goldens, golden_dir):
    """End to end with getSource on the device: U within 1e-12 relative of the reference."""
    import multigrid_poisson_solver_b200 as mg
    path = os.path.join(golden_dir, "cycle_%s.txt" % name)
    r = mg.run_cycle_host(path, mg.RUN_FUSED | mg.RUN_QUIET)
    ref = po.run_cycle(path)
    scale = np.max(np.abs(ref["U"]))
    assert np.max(np.abs(r["U"] - ref["U"])) <= 1e-12 * scale
    assert r["mg_error"] == pytest.approx(goldens[name]["mg_error"], rel=1e-10)


def test_gs_iteration_counts_in_cycles(golden_dir):
    """GS iterations N=8 @1e-7 -> 76 (Vcycle.txt), N=32 @1e-8 -> 1739 (Wcycle.txt)."""
    import multigrid_poisson_solver_b200 as mg
    r = mg.run_cycle_host(os.path.join(golden_dir, "cycle_Vcycle.txt"), mg.RUN_FUSED | mg.RUN_QUIET)
    assert [t["steps"] for t in r["trace"] if t["node"] == 0] == [76]
    r = mg.run_cycle_host(os.path.join(golden_dir, "cycle_Wcycle.txt"), mg.RUN_FUSED | mg.RUN_QUIET)
    assert [t["steps"] for t in r["trace"] if t["node"] == 0] == [1739] * 4


def test_cli_log_matches_reference_binary(golden_dir, tmp_path):
    """MG_GPU prints the same blocks and %lf values as the real MG_CPU."""
    import shutil
    import subprocess
    import multigrid_poisson_solver_b200 as mg
    exe = os.path.join(os.path.dirname(mg.lib_path()), "MG_GPU")
    for name in ["test", "Vcycle", "VcycleTrigger", "Wcycle"] + MODE_CYCLES:
        shutil.copy(os.path.join(golden_dir, "cycle_%s.txt" % name), tmp_path / ("cycle_%s.txt" % name))
        p = subprocess.run([exe, "1", "cycle_%s.txt" % name], cwd=tmp_path, capture_output=True, text=True, check=True)
        mine = [l for l in p.stdout.splitlines() if not l.startswith("Time Used")]
        ref = open(os.path.join(golden_dir, "MG_CPU_%s.log" % name)).read().splitlines()
        ref = [l.replace("Sol_CPU_", "Sol_GPU_") for l in ref]
        assert mine == ref
    for name in ["test"] + MODE_CYCLES:      # the solution file, byte for byte (doPrint2File, %lf)
        assert open(tmp_path / ("Sol_GPU_cycle_%s.txt" % name)).read() == open(os.path.join(golden_dir, "MG_CPU_%s.csv" % name)).read()


def test_problem_plug_point_nonzero_dirichlet(orc, tmp_path):
    """caller-supplied source, initial grid carrying NON-ZERO Dirichlet boundary values, and reference
    solution.  Checked against the same node sequence composed from the oracle's operators (the reference compiles
    its problem in, :488 / :509-519 / :544, so its main() cannot run this): U bit-identical, errors <= 1e-10."""
    import multigrid_poisson_solver_b200 as mg
    N, M = 64, 32
    rng = np.random.default_rng(9)
    x = np.linspace(0.0, 1.0, N)
    exact = np.add.outer(np.sin(2.0 * x), np.cos(3.0 * x))          # u(x, y); Laplace(u) = -(4 sin 2y' ... ) supplied as F below
    F = -(4.0 * np.sin(2.0 * x)[:, None] + 9.0 * np.cos(3.0 * x)[None, :]) + 0.0 * exact
    U0 = np.zeros((N, N))
    U0[0, :], U0[-1, :], U0[:, 0], U0[:, -1] = exact[0, :], exact[-1, :], exact[:, 0], exact[:, -1]   # Dirichlet data, zero interior
    F, U0, exact = F.reshape(-1), U0.reshape(-1), exact.reshape(-1)
    path = tmp_path / "two_grid.txt"
    path.write_text(mg.cycles.two_grid(N, M, step=3, tol=1e-6))
    for flags in (mg.RUN_FUSED | mg.RUN_QUIET, mg.RUN_UNFUSED | mg.RUN_QUIET):
        r = mg.run_cycle_problem(str(path), flags, F_host=F, U0_host=U0, analytic_host=exact)
        # the same nodes with the oracle's operators: -1 (U kept: restart semantics), 0, 1
        U, e_down = orc.doSmoothing(N, 1.0, U0, F, 3)
        Fc = orc.doRestriction(N, -orc.getResidual(N, 1.0, U, F), M)
        Uc = orc.doExactSolver(M, 1.0, Fc, 1e-6, 1)
        U = orc.doGridAddition(N, U, orc.doProlongation(M, Uc, N))
        U, e_up = orc.doSmoothing(N, 1.0, U, F, 3)
        assert_same(r["U"], U, "plug point U")
        errs = [t["err"] for t in r["trace"] if t["node"] != 0]
        assert errs[0] == pytest.approx(e_down, rel=ERR_RTOL) and errs[1] == pytest.approx(e_up, rel=ERR_RTOL)
        assert r["mg_error"] == pytest.approx(np.mean(np.abs(exact - U)), rel=1e-12)
        assert np.array_equal(r["U"].reshape(N, N)[0], exact.reshape(N, N)[0])      # the boundary data came through untouched


# ------------------------------------------------------------------ full-size properties (no oracle needed)
@pytest.mark.parametrize("N", [4096, 16384])
def test_full_size_properties(N):
    """At BASELINE sizes: fused == unfused bit for bit, and the V-cycle reaches the analytic
    error the reference reaches (~8.83e-4 for N = 2048 ... 16384)."""
    import tempfile
    import multigrid_poisson_solver_b200 as mg
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(mg.cycles.v_cycle(N, 8))
        path = f.name
    a = mg.run_cycle_host(path, mg.RUN_FUSED | mg.RUN_QUIET)
    b = mg.run_cycle_host(path, mg.RUN_UNFUSED | mg.RUN_QUIET)
    os.unlink(path)
    assert np.array_equal(a["U"], b["U"])
    assert [t["steps"] for t in a["trace"]] == [t["steps"] for t in b["trace"]]
    for x, y in zip(a["trace"], b["trace"]):
        assert x["err"] == pytest.approx(y["err"], rel=ERR_RTOL)
    assert a["mg_error"] == pytest.approx(8.83e-4, rel=2e-2)
    U = a["U"].reshape(N, N)
    assert np.max(np.abs(U[0])) < 1e-9 and np.max(np.abs(U[:, -1])) < 1e-9
