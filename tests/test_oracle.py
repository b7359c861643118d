"""CPU suite: pins the oracle (oracle/mg_oracle*.c) against
  (1) the committed golden vectors generated from the unmodified reference
      (tests/golden/make_golden.py), which include the SURVEY.md 8(c) values, and
  (2) when oracle/_ref is present, the unmodified reference operators directly,
      bit for bit, on seeded inputs and many transfer-size pairs.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

SURVEY_8C = {  # values printed in SURVEY.md 8(c), captured from the unmodified reference
    "test": dict(errs=[0.52178714565614204, 0.014560224318122416], mg=0.00066580956860994771,
                 sumU=6.3347907458743862, maxU=0.068350537517844742),
    "Vcycle": dict(errs=[0.75504134183443272, 0.73808340673350903, 0.70466404745011391, 0.63809513888352032,
                         0.50816677423068746, 0.014092698296060024, 0.019296718366329537, 0.021170848670371212,
                         0.021852498889273535, 0.022120785695616058],
                   mg=0.00087564990470026624, sumU=1841.0710375080112, maxU=0.068171195385991118),
    "VcycleTrigger": dict(errs=[0.75643250762339986, 0.74142147770217215, 0.7116964802851411, 0.63139712614503352,
                                0.36072134327974981, 0.010324679693711892, 0.015739912572795941,
                                0.017571158738804705, 0.018211897197471849, 0.018477751736314368],
                          steps=[2, 2, 2, 4, 14, 2, 2, 2, 2, 2],
                          mg=0.00078392043391262991, sumU=1847.0826201055577, maxU=0.06838780181654075),
    "Wcycle": dict(mg=5.0079485211558038e-05, sumU=1895.1756205195968, maxU=0.070433019807353864),
}


@pytest.fixture(scope="module")
def goldens(golden_dir):
    with open(os.path.join(golden_dir, "cycles.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def orc():
    return po.oracle_ops()


def same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


MODE_CYCLES = ["mode_nodestep_autoN", "mode_fixedstep_manualN", "mode_nodestep_minus1", "mode_manual_step0"]


@pytest.mark.parametrize("name", ["test", "Vcycle", "VcycleTrigger", "Wcycle", "V_minus_one_ladder", "V_restart_x2",
                                  "manual_nonnested", "W_full_64", "V_lu_coarse", "V_offset_domain"] + MODE_CYCLES)
def test_cycle_matches_reference_golden(name, goldens, golden_dir):
    g = goldens[name]
    r = po.run_cycle(os.path.join(golden_dir, "cycle_%s.txt" % name))
    assert r["trace"] == g["trace"]          # bit-identical errors, step counts, sum/max fingerprints
    assert r["mg_error"] == g["mg_error"] and r["sumU"] == g["sumU"] and r["maxabsU"] == g["maxabsU"]
    upath = os.path.join(golden_dir, "cycle_U_%s.npy" % name)
    if os.path.exists(upath):
        assert same(r["U"], np.load(upath))


@pytest.mark.parametrize("name", sorted(SURVEY_8C))
def test_cycle_matches_survey_values(name, golden_dir):
    s = SURVEY_8C[name]
    r = po.run_cycle(os.path.join(golden_dir, "cycle_%s.txt" % name))
    smooth = [t for t in r["trace"] if t["node"] != 0]
    if "errs" in s:
        assert [t["err"] for t in smooth] == s["errs"]
    if "steps" in s:
        assert [t["steps"] for t in smooth] == s["steps"]
    assert r["mg_error"] == s["mg"] and r["sumU"] == s["sumU"] and r["maxabsU"] == s["maxU"]


def test_wcycle_rezero_quirk(golden_dir):
    """Every re-descent wipes the level's correction (MG_solver_CPU.cpp:252-257)."""
    r = po.run_cycle(os.path.join(golden_dir, "cycle_Wcycle.txt"))
    down64 = [t["err"] for t in r["trace"] if t["node"] == -1 and t["N"] == 64]
    assert len(down64) == 4 and len(set(down64)) == 1 and down64[0] == 0.70466404745011391


@pytest.mark.parametrize("N", [17, 32, 45])
def test_operators_match_reference_vectors(N, orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "ops_N%d.npz" % N))
    _, Mc, Mf = (int(v) for v in g["shape"])
    U, F, Uz = g["U"], g["F"], g["Uz"]
    assert same(orc.getSource(N), g["source"])
    assert same(orc.getSource(N, 2.0, -0.5, 0.25), g["source_shift"])
    assert same(orc.getAnalytic(N), g["analytic"])
    assert same(orc.getResidual(N, 1.0, U, F), g["residual"])
    assert same(orc.doGridAddition(N, U, F), g["add"])
    for s in (1, 2, 3, 4, 7):
        a, e = orc.doSmoothing(N, 1.0, Uz, F, s)
        assert same(a, g["smooth%d" % s]) and e == float(g["smooth%d_err" % s][0])
    assert same(orc.doRestriction(N, U, Mc), g["restrict"])
    assert same(orc.doProlongation(N, U, Mf), g["prolong"])
    assert same(orc.doExactSolver(N, 1.0, F, 1e-7, 1), g["gs"])


def test_smoother_is_jacobi_and_error_is_red_only(orc):
    """both half sweeps read the snapshot; error counts red points twice."""
    N = 12
    rng = np.random.default_rng(3)
    U = rng.random((N, N)); F = rng.random((N, N))
    out, err = orc.doSmoothing(N, 1.0, U.reshape(-1), F.reshape(-1), 1)
    h2 = (1.0 / (N - 1)) ** 2
    jac = U.copy()
    jac[1:-1, 1:-1] = U[1:-1, 1:-1] + 0.25 * (U[2:, 1:-1] + U[:-2, 1:-1] + U[1:-1, 2:] + U[1:-1, :-2]
                                               - 4 * U[1:-1, 1:-1] - h2 * F[1:-1, 1:-1])
    assert np.allclose(out.reshape(N, N), jac, rtol=1e-14, atol=0)
    V = out.reshape(N, N)
    res = np.zeros((N, N))
    res[1:-1, 1:-1] = (V[2:, 1:-1] + V[:-2, 1:-1] + V[1:-1, 2:] + V[1:-1, :-2] - 4 * V[1:-1, 1:-1]) / h2 - F[1:-1, 1:-1]
    ii, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    red = ((ii + jj) % 2 == 0)
    assert err == pytest.approx(2 * np.abs(res[red]).sum() / N / N, rel=1e-13)


def test_transfers_reproduce_linear_ramp(orc):
    """Fixture of testFunction/Test_doRestriction_GPU.cu:189-193: bilinear transfer is exact on i+j."""
    N, M = 16, 8
    ramp = np.add.outer(np.arange(N), np.arange(N)).astype(float).reshape(-1)
    c = orc.doRestriction(N, ramp, M).reshape(M, M)
    k = np.arange(1, M - 1) * (N - 1) / (M - 1)
    assert np.allclose(c[1:-1, 1:-1], np.add.outer(k, k), rtol=1e-13)
    assert (c[0] == 0).all() and (c[:, 0] == 0).all() and (c[-1] == 0).all() and (c[:, -1] == 0).all()
    f = orc.doProlongation(M, np.add.outer(np.arange(M), np.arange(M)).astype(float).reshape(-1), N).reshape(N, N)
    k = np.arange(N) * (M - 1) / (N - 1)
    assert np.allclose(f, np.add.outer(k, k), rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------ against the live reference build
needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (reference sources absent)")


@needs_ref
@pytest.mark.parametrize("N", [3, 4, 5, 9, 16, 33, 100, 256])
def test_oracle_equals_reference_operators(N, orc):
    ref = po.ref_ops(1)
    rng = np.random.default_rng(N)
    U = rng.random(N * N); F = rng.random(N * N)
    assert same(orc.getSource(N), ref.getSource(N))
    assert same(orc.getAnalytic(N, 1.5, 0.1, -0.2), ref.getAnalytic(N, 1.5, 0.1, -0.2))
    assert same(orc.getBoundary(N), ref.getBoundary(N))
    assert same(orc.getResidual(N, 0.7, U, F), ref.getResidual(N, 0.7, U, F))
    assert same(orc.doGridAddition(N, U, F), ref.doGridAddition(N, U, F))
    for s in (1, 3, 6):
        a, ea = orc.doSmoothing(N, 1.0, U, F, s)
        b, eb = ref.doSmoothing(N, 1.0, U, F, s)
        assert same(a, b) and ea == eb


@needs_ref
def test_oracle_equals_reference_transfers(orc):
    ref = po.ref_ops(1)
    rng = np.random.default_rng(7)
    for N in list(range(3, 70)) + [127, 128, 129, 255, 256, 512]:
        g = rng.random(N * N) - 0.5
        for M in sorted({N // 2, N - 1, (N + 1) // 2, N // 3 + 2}):
            if 3 <= M < N:
                assert same(orc.doRestriction(N, g, M), ref.doRestriction(N, g, M)), (N, M)
        for M in sorted({2 * N, N + 1, 2 * N - 1, 2 * N + 1}):
            assert same(orc.doProlongation(N, g, M), ref.doProlongation(N, g, M)), (N, M)


@needs_ref
@pytest.mark.parametrize("N,tol,opt", [(8, 1e-7, 1), (16, 1e-7, 1), (32, 1e-8, 1), (21, 1e-9, 1), (5, 0, 0), (8, 0, 0)])
def test_oracle_equals_reference_exact_solver(N, tol, opt, orc):
    ref = po.ref_ops(1)
    F = np.random.default_rng(N).random(N * N)
    assert same(orc.doExactSolver(N, 1.0, F, tol, opt), ref.doExactSolver(N, 1.0, F, tol, opt))


@needs_ref
def test_driver_matches_real_binary_log(golden_dir, tmp_path):
    """The oracle's driver must print-equal the real ./MG_CPU at %lf precision."""
    import subprocess
    for name in ["Vcycle", "VcycleTrigger", "Wcycle"] + MODE_CYCLES:
        log = open(os.path.join(golden_dir, "MG_CPU_%s.log" % name)).read().splitlines()
        errs = [l.split("=")[1].strip() for l in log if l.strip().startswith("Error =")]
        steps = [int(l.split("=")[1]) for l in log if "Smoothing Steps" in l]
        sizes = [int(l.split("=")[1]) for l in log if "Current Grid Size" in l]
        r = po.run_cycle(os.path.join(golden_dir, "cycle_%s.txt" % name))
        # a "1" node with step 0 prints no smoothing block (:410-412); a step-0 "-1" node does nothing at all
        smooth = [t for t in r["trace"] if t["node"] != 0 and not (t["node"] == 1 and t["steps"] == 0)]
        assert ["%f" % t["err"] for t in smooth] + ["%f" % r["mg_error"]] == errs
        assert [t["steps"] for t in smooth] == steps
        assert [t["N"] for t in r["trace"] if not (t["node"] == 1 and t["steps"] == 0)] == sizes


def csv_text(U, N):
    """doPrint2File's layout (MG_solver_CPU.cpp:735-754): %lf, row j = N-1 first."""
    g = np.asarray(U).reshape(N, N)
    return "".join(",".join("%f" % v for v in g[j]) + "\n" for j in range(N - 1, -1, -1))


@pytest.mark.parametrize("name", ["test"] + MODE_CYCLES)
def test_final_grid_matches_real_binary_csv(name, golden_dir):
    """Control flow of the remaining parser modes (step 0 / negative steps, (0,!=0), (!=0,0)) is pinned by
    the solution file the real ./MG_CPU wrote for the same cycle file (6 decimals: any wrong branch shows)."""
    r = po.run_cycle(os.path.join(golden_dir, "cycle_%s.txt" % name))
    assert csv_text(r["U"], r["N"]) == open(os.path.join(golden_dir, "MG_CPU_%s.csv" % name)).read()
