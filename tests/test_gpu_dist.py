"""GPU suite for the row-slab multi-GPU driver, exercised on ONE GPU: all ranks are emulated in
one process (mgDistEmuRunCycleFile), with device-to-device copies in place of NCCL.  The slab
path must reproduce the single-GPU result bit for bit (same arithmetic per point), and its
all-reduced errors to <= 1e-10 relative."""
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


@pytest.fixture
def overlapped(monkeypatch):
    """MG_DIST_OVERLAP=1: passes launched as edge + interior, communication on its own stream."""
    monkeypatch.setenv("MG_DIST_OVERLAP", "1")
    yield
    monkeypatch.delenv("MG_DIST_OVERLAP", raising=False)


@pytest.mark.parametrize("text,world,threshold", [
    ("v1024", 2, 256), ("v1024", 8, 256), ("w512", 4, 128), ("trigger512", 4, 128), ("step5", 3, 128), ("restart", 4, 128),
    ("v4096", 4, 1024)])
def test_overlapped_exchange_matches_single_gpu(mg, overlapped, text, world, threshold):
    """The split-launch / two-stream protocol must give the same bits as the in-line one."""
    texts = {"v1024": mg.cycles.v_cycle(1024, 8), "w512": mg.cycles.w_cycle(512, 8, levels=4, step=2, tol=1e-7),
             "trigger512": mg.cycles.v_cycle(512, 8, step=-1), "step5": mg.cycles.v_cycle(512, 16, step=5),
             "restart": mg.cycles.v_cycle(512, 8, step=2, cycles=2), "v4096": mg.cycles.v_cycle(4096, 8)}
    compare(mg, texts[text], world, threshold)


@pytest.fixture
def staged(monkeypatch):
    """MG_DIST_TRANSPORT=staged: row transfers through staging buffers, copy engines and stream
    memory operations (the emulated ranks run the protocol of the real IPC transport in one process)."""
    monkeypatch.setenv("MG_DIST_TRANSPORT", "staged")
    yield
    monkeypatch.delenv("MG_DIST_TRANSPORT", raising=False)


@pytest.mark.parametrize("text,world,threshold,overlap", [
    ("v1024", 2, 256, False), ("v1024", 8, 256, False), ("w512", 4, 128, False), ("trigger512", 4, 128, False),
    ("step5", 3, 128, False), ("restart", 4, 128, False), ("v4096", 4, 1024, False),
    ("v1024", 4, 256, True), ("w512", 4, 128, True), ("v4096", 8, 1024, True)])
def test_staged_transport_matches_single_gpu(mg, staged, monkeypatch, text, world, threshold, overlap):
    """Sequence flags, slot parity, offsets and acknowledgements of the staged transport: same bits
    as the single-GPU run, in line and with the exchange on its own stream behind the edge launch."""
    if overlap:
        monkeypatch.setenv("MG_DIST_OVERLAP", "1")
    texts = {"v1024": mg.cycles.v_cycle(1024, 8), "w512": mg.cycles.w_cycle(512, 8, levels=4, step=2, tol=1e-7),
             "trigger512": mg.cycles.v_cycle(512, 8, step=-1), "step5": mg.cycles.v_cycle(512, 16, step=5),
             "restart": mg.cycles.v_cycle(512, 8, step=2, cycles=2), "v4096": mg.cycles.v_cycle(4096, 8)}
    compare(mg, texts[text], world, threshold)


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
