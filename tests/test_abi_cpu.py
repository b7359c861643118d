"""CPU suite: the C-ABI library builds, loads without a GPU and exports every symbol that
include/mg_abi.h declares; the product path fails loudly (no CPU fallback) without a device;
host-side cycle-file logic."""
import ctypes
import os
import re

import numpy as np
import pytest

import multigrid_poisson_solver_b200 as mg
from multigrid_poisson_solver_b200 import api, cycles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mg_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct.*?\}\s*\w+;", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built():
    from multigrid_poisson_solver_b200 import build
    build.build()
    return mg.lib_path()


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 30
    assert sorted(api.ABI) == names


def test_library_exports_every_declared_symbol(built):
    raw = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(raw, name), name
    for name in ("getSource", "getBoundary", "getResidual", "doSmoothing", "doRestriction", "doProlongation",
                 "doGridAddition", "doExactSolver"):
        assert hasattr(raw, name)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "multigrid_poisson_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, fn)).read()
                code = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("//", "#!", "*", "/*")))
                assert not re.search(r"^\s*(from|import)\s+oracle", code, flags=re.M), fn
                assert not re.search(r"#include\s*[<\"][^>\"]*oracle", code), fn
                assert "liboracle" not in code and "libmgref" not in code and "_ref/" not in code, fn


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_fails_loudly_without_a_gpu(built):
    with pytest.raises(mg.MGLibraryError):
        mg.init(0)
    raw = mg.lib()
    raw.mgClearError()
    raw.getSource(8, 1.0, None, 0.0, 0.0)        # no context: must report, not compute
    assert raw.mgLastErrorCode() != 0
    raw.mgClearError()


def test_cycle_generators_reproduce_shipped_shapes(golden_dir):
    shipped = {
        "Vcycle": "1.0 0.0 0.0 3 1 256 8 -1 -1 -1 -1 -1 0 0.0000001 1 1 1 1 1 1 2",
        "VcycleTrigger": "1.0 0.0 0.0 -1 1 256 8 -1 -1 -1 -1 -1 0 0.0000001 1 1 1 1 1 1 2",
        "test": "1.0 0.0 0.0 3 1 16 8 -1 0 0.00000001 1 1 2",
        "Wcycle": "1.0 0.0 0.0 3 1 256 8 -1 -1 -1 0 0.00000001 1 1 -1 0 0.00000001 1 1 1 -1 -1 0 0.00000001 1 1 -1 0 "
                  "0.00000001 1 1 1 1 2",
    }
    made = {"Vcycle": cycles.v_cycle(256, 8), "VcycleTrigger": cycles.v_cycle(256, 8, step=-1),
            "test": cycles.two_grid(16, 8), "Wcycle": cycles.w_cycle(256, 8, levels=3)}
    for name, text in shipped.items():
        assert cycles.tokens(made[name]) == cycles.tokens(text)
        assert cycles.tokens(open(os.path.join(golden_dir, "cycle_%s.txt" % name)).read()) == cycles.tokens(text)
        assert not made[name].endswith("\n")
    if os.path.isdir("/root/reference/src"):
        for name in shipped:
            assert cycles.tokens(open("/root/reference/src/%s.txt" % name).read()) == cycles.tokens(made[name])


def test_ladders():
    assert cycles.ladder(16384, 8) == [16384 >> k for k in range(12)]
    assert cycles.ladder(40, 33, con_N=2) == list(range(40, 32, -1))
    w = cycles.tokens(cycles.w_cycle(64, 8))
    assert w.count(0.0) >= 4    # 2^(levels-1) exact solves (+ the 0.0 origin fields)


# ----------------------------------------------------------------------------- task geometry (host-only)
P_RES, P_PLAIN = 148 * 2 * 8, 148 * 3 * 4      # resident warps of the two CTA shapes (mg_stream.cuh)


@pytest.mark.parametrize("rows,n_strips,warps", [
    (16384, 293, P_PLAIN), (16384, 304, P_RES), (8192, 147, P_PLAIN), (11584, 414, P_RES), (5792, 828, P_RES),
    (4096, 74, P_PLAIN), (2048, 37, P_RES), (1024, 19, P_PLAIN), (256, 5, P_PLAIN), (64, 2, P_RES), (7, 1, P_RES),
    (65536, 1171, P_PLAIN), (1447, 52, P_PLAIN)])
def test_segment_plan_partitions_the_owned_rows(rows, n_strips, warps):
    """The row segments of a fused pass tile [0, rows) exactly once, in order, without empty ones;
    large grids get tall segments first and short ones last (no tail), small grids uniform ones."""
    import multigrid_poisson_solver_b200 as mg
    segs = mg.api.segment_plan(rows, n_strips, warps, 9, 0)
    assert segs[0][0] == 0 and segs[-1][1] == rows
    for (a, b), (c, d) in zip(segs, segs[1:]):
        assert a < b and b == c
    heights = [b - a for a, b in segs]
    assert max(heights) <= 256
    assert len(segs) <= 16384
    if len(segs) * n_strips >= 4 * warps:                      # throughput regime: guided schedule
        assert heights[0] == max(heights)
        assert all(h1 >= h2 for h1, h2 in zip(heights[:-2], heights[1:-1]))   # non-increasing (last one may absorb a sliver)
        assert heights[-1] <= max(32, heights[0] // 4)
    # deterministic: the plan is a pure function of its arguments
    assert segs == mg.api.segment_plan(rows, n_strips, warps, 9, 0)


@pytest.mark.parametrize("rows", [40, 97, 120, 121, 500, 5792, 11584, 16384])
def test_split_pass_edge_plus_interior_is_the_whole_pass(rows):
    """Slab passes launched in two parts (mg_dist.cu, MG_DIST_OVERLAP): the edge launch and the
    interior launch together cover every owned row exactly once, and the edge launch alone holds
    the first and last 24 rows -- all a neighbour's halo (8 rows, <= 22 fine rows for the
    restricted grid) can need."""
    import multigrid_poisson_solver_b200 as mg
    edge = mg.api.segment_plan(rows, 100, P_RES, 9, 1)
    inner = mg.api.segment_plan(rows, 100, P_RES, 9, 2)
    cover = np.zeros(rows, dtype=int)
    for a, b in edge + inner:
        assert 0 <= a < b <= rows
        cover[a:b] += 1
    assert np.all(cover == 1)
    in_edge = np.zeros(rows, dtype=bool)
    for a, b in edge:
        in_edge[a:b] = True
    assert np.all(in_edge[:min(24, rows)]) and np.all(in_edge[max(0, rows - 24):])
    if rows >= 24 * 5:
        assert len(edge) == 4 and inner


# ----------------------------------------------------------------------------- bench contract (CPU arm)
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU implementation timed on the host) prints
    ONE JSON line with the keys of the bench contract; bounded sample so that it runs in a second."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--nmax", "128",
                          "--steps", "5", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("V-cycles/sec") and d["unit"] == "V-cycles/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the stock main() is capped at 1 warm-up + 2 timed cycles whatever was asked for, and the line says what ran
    assert d["steps"] == 2 and d["warmup"] == 1 and "requested --steps 5 --warmup 3" in d["config"]["steps_note"]
    if d["cpu_baseline"]["kind"] == "reference":
        assert "stock MG_CPU main()" in d["cpu_baseline"]["sample"] and d["same_config"] is True


def test_segment_plan_properties_random():
    """Property check over random shapes (hypothesis): whatever the grid, strip count and CTA shape,
    the segments tile the owned rows exactly once and the split launches add up to the whole pass."""
    from hypothesis import given, settings, strategies as st
    import multigrid_poisson_solver_b200 as mg

    @settings(max_examples=300, deadline=None)
    @given(rows=st.integers(1, 70000), n_strips=st.integers(1, 1200), warps=st.sampled_from([P_RES, P_PLAIN, 148 * 2 * 4]),
           lead=st.integers(3, 11))
    def check(rows, n_strips, warps, lead):
        whole = mg.api.segment_plan(rows, n_strips, warps, lead, 0)
        assert whole[0][0] == 0 and whole[-1][1] == rows
        assert all(a < b for a, b in whole) and all(x[1] == y[0] for x, y in zip(whole, whole[1:]))
        assert max(b - a for a, b in whole) <= 256 and len(whole) <= 16384
        cover = np.zeros(rows, dtype=np.int8)
        for a, b in mg.api.segment_plan(rows, n_strips, warps, lead, 1) + mg.api.segment_plan(rows, n_strips, warps, lead, 2):
            cover[a:b] += 1
        assert np.all(cover == 1)

    check()
