"""GPU suite for the row-slab multi-GPU driver, exercised on ONE GPU: all ranks are emulated in
one process (mgDistEmuRunCycleFile) with the machinery of the multi-process run -- one arena per
rank, halo rows stored into the neighbours' arenas by the fused kernels, flag words and stream
waits, broadcast of the agglomerated source, redundant coarse sub-cycles.  The slab path must
reproduce the single-GPU result bit for bit (same arithmetic per point), and its all-reduced
errors to <= 1e-10 relative."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import multigrid_poisson_solver_b200 as m
    m.init(0)
    return m


def cycle_file(text):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    f.write(text)
    f.close()
    return f.name


def compare(mg, text, world, threshold):
    path = cycle_file(text)
    try:
        one = mg.run_cycle_host(path, mg.RUN_FUSED | mg.RUN_QUIET)
        emu = mg.run_cycle_dist_emulated(path, world, threshold)
    finally:
        os.unlink(path)
    assert [(t["node"], t["N"]) for t in emu["trace"]] == [(t["node"], t["N"]) for t in one["trace"]]
    for a, b in zip(emu["trace"], one["trace"]):
        assert a["steps"] == b["steps"], (a, b)
        if b["node"] != 0:
            assert a["err"] == pytest.approx(b["err"], rel=1e-10), (a, b)
    bad = np.flatnonzero(emu["U"] != one["U"])
    assert bad.size == 0, "U differs at %d points, first %d (row %d)" % (bad.size, bad[0], bad[0] // one["N"])
    assert emu["mg_error"] == pytest.approx(one["mg_error"], rel=1e-10)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_vcycle_slabs_match_single_gpu(mg, world):
    compare(mg, mg.cycles.v_cycle(1024, 8), world, 256)


@pytest.mark.parametrize("world,threshold", [(2, 64), (4, 128), (8, 1024), (2, 2048)])
def test_vcycle_thresholds(mg, world, threshold):
    compare(mg, mg.cycles.v_cycle(1024, 8), world, threshold)   # (2, 2048): nothing is distributed


def test_wcycle_slabs(mg):
    compare(mg, mg.cycles.w_cycle(512, 8, levels=4, step=2, tol=1e-7), 4, 128)


def test_trigger_cycle_slabs(mg):
    compare(mg, mg.cycles.v_cycle(512, 8, step=-1), 4, 128)


@pytest.mark.parametrize("step", [1, 2, 4, 5, 7])
def test_multi_pass_steps(mg, step):
    compare(mg, mg.cycles.v_cycle(512, 16, step=step), 3, 128)


def test_restart_chain_slabs(mg):
    compare(mg, mg.cycles.v_cycle(512, 8, step=2, cycles=2), 4, 128)


def test_non_power_of_two_ladder(mg):
    compare(mg, mg.cycles.v_cycle(1448, 8), 4, 300)        # 1448 -> 724 -> 362 -> 181 (odd, agglomerated)


def test_two_gathers_back_to_back(mg):
    """A W-cycle whose agglomeration boundary is crossed twice in a row: the gather buffers alternate."""
    compare(mg, mg.cycles.w_cycle(1024, 8, levels=5, step=1, tol=1e-7), 4, 512)


def test_top_level_below_threshold(mg):
    compare(mg, mg.cycles.v_cycle(256, 8), 4, 1024)        # nothing distributed: every rank runs the whole cycle


@pytest.mark.parametrize("name", ["mode_nodestep_autoN", "mode_nodestep_minus1", "mode_manual_step0", "V_restart_x2"])
def test_parser_modes_on_slabs(mg, name, golden_dir):
    """step 0 / negative steps / per-node options through the slab driver (thresholds chosen so that the top levels
    are distributed where their sizes allow it)."""
    path = os.path.join(golden_dir, "cycle_%s.txt" % name)
    compare(mg, open(path).read(), 2, 16)


def test_large_grid_slabs(mg):
    compare(mg, mg.cycles.v_cycle(4096, 8), 8, 1024)


@pytest.mark.parametrize("N,step", [(256, 1), (512, 3), (1024, 7), (1000, 100)])
def test_smoothing_stress_matches_do_smoothing(mg, N, step):
    """mgDistSmoothStress on one GPU (one slab = the whole grid) == doSmoothing from U = 0."""
    import ctypes as C
    lib = mg.lib()
    ms, err, lo, hi = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
    U = np.empty(N * N)
    rc = lib.mgDistSmoothStress(N, 1.0, step, 1, C.byref(ms), C.byref(err), U.ctypes.data, C.byref(lo), C.byref(hi))
    assert rc == 0 and (lo.value, hi.value) == (0, N)
    g = mg.GpuOps()
    ref, eref = g.doSmoothing(N, 1.0, np.zeros(N * N), g.getSource(N), step)
    assert np.array_equal(U, ref)
    assert err.value == pytest.approx(eref, rel=1e-10)


def test_split_launch_path_on_small_slabs():
    """The edge + interior split of a slab pass (normally only from 32 M owned points on) forced on small slabs with
    MG_SPLIT_MIN_POINTS=1 (read when the library is loaded, hence the subprocess): the edge launch with peer stores and
    flags, then the interior launch adding its error sums -- same bits as the single-GPU run, V / trigger / W / restart cycles."""
    import subprocess
    import sys
    code = r'''
import os, sys, tempfile
import numpy as np
import multigrid_poisson_solver_b200 as mg
mg.init(0)
cases = [(mg.cycles.v_cycle(1024, 8), 2, 256), (mg.cycles.v_cycle(1024, 8), 4, 256), (mg.cycles.v_cycle(1024, 8, step=-1), 3, 256),
         (mg.cycles.w_cycle(1024, 8, levels=4, step=2, tol=1e-7), 2, 256), (mg.cycles.v_cycle(1024, 8, step=7, cycles=2), 2, 512)]
for text, world, thr in cases:
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False); f.write(text); f.close()
    one = mg.run_cycle_host(f.name, mg.RUN_FUSED | mg.RUN_QUIET)
    emu = mg.run_cycle_dist_emulated(f.name, world, thr)
    os.unlink(f.name)
    assert np.array_equal(emu["U"], one["U"]), (world, thr)
    assert [t["steps"] for t in emu["trace"]] == [t["steps"] for t in one["trace"]]
    for a, b in zip(emu["trace"], one["trace"]):
        if b["node"] != 0:
            assert abs(a["err"] - b["err"]) <= 1e-10 * abs(b["err"]), (a, b)
print("SPLIT_OK")
'''
    env = dict(os.environ, MG_SPLIT_MIN_POINTS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert r.returncode == 0 and "SPLIT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
