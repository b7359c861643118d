"""CPU suite: the C-ABI library builds, loads without a GPU and exports every symbol that
include/mg_abi.h declares; the product path fails loudly (no CPU fallback) without a device;
host-side cycle-file logic."""
import ctypes
import os
import re

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
