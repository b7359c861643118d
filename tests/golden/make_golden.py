"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference/src).  Run in the build container only:

    python tests/golden/make_golden.py

Writes
  cycles.json       per-node records (full-precision smoothing errors, sum/max of U)
                    and final results of the reference operators driven node by node,
                    for the shipped cycle shapes and a few extra modes
  cycle_U_*.npy     final U of the small cycles
  ops_N*.npz        seeded inputs + reference outputs of every operator
  MG_CPU_*.log      stdout of the real ./MG_CPU binary (format + %lf values)
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from multigrid_poisson_solver_b200 import cycles as cy  # noqa: E402

CYCLES = {
    "test": cy.two_grid(16, 8),
    "Vcycle": cy.v_cycle(256, 8),
    "VcycleTrigger": cy.v_cycle(256, 8, step=-1),
    "Wcycle": cy.w_cycle(256, 8, levels=3),
    # extra modes of the format
    "V_minus_one_ladder": cy.v_cycle(40, 33, step=2, tol=1e-6, con_N=2),
    "V_restart_x2": cy.v_cycle(64, 8, step=2, tol=1e-7, cycles=2),
    "manual_nonnested": cy.manual([(-1, 2, 21), (-1, 3, 11), (0, 1e-8, 1), (1, 1), (1, 4)], 45, 11),
    "W_full_64": cy.w_cycle(64, 8, step=1, tol=1e-7),
    "V_lu_coarse": cy.v_cycle(32, 8, step=2, tol=0.0, option=0),
    "V_offset_domain": cy.v_cycle(64, 8, step=3, tol=1e-7, L=2.0, min_x=-0.5, min_y=0.25),
    # the remaining (con_step, con_N) parser modes and the step == 0 / negative-step nodes
    # (MG_solver_CPU.cpp:171-189, :241-243, :296-299, :331-344, :410-421)
    # (0, 1): per-node steps on the automatic ladder; the step-0 "-1" node only advances the ladder
    # position (:174-179), so the next restriction goes 32 -> 8; the last "1" node skips its smoothing
    "mode_nodestep_autoN": cy.manual([(-1, 2, 0), (-1, 0, 0), (-1, 3, 0), (0, 1e-7, 1), (1, 2), (1, 0)], 64, 8, con_step=0, con_N=1),
    # (3, 0): fixed steps, per-node next_N (non-nested sizes)
    "mode_fixedstep_manualN": cy.manual([(-1, 0, 21), (-1, 0, 11), (0, 1e-8, 1), (1, 0), (1, 0)], 45, 11, con_step=3, con_N=0),
    # (0, 2): per-node steps on the N -> N-1 ladder, with a step of -2 (zero sweeps, error still evaluated)
    "mode_nodestep_minus1": cy.manual([(-1, 1, 0), (-1, -2, 0), (-1, 2, 0), (0, 1e-6, 1), (1, 1), (1, -2), (1, 3)], 24, 21, con_step=0, con_N=2),
    # (0, 0) with step 0 on both node kinds: the step-0 "-1" node reads its next_N and does nothing
    "mode_manual_step0": cy.manual([(-1, 2, 16), (-1, 0, 12), (-1, 1, 8), (0, 1e-7, 1), (1, 0), (1, 2)], 32, 8),
}
BINARY_LOGS = ("test", "Vcycle", "VcycleTrigger", "Wcycle", "mode_nodestep_autoN", "mode_fixedstep_manualN", "mode_nodestep_minus1",
               "mode_manual_step0")


def main():
    assert po.have_ref(), "build oracle/_ref first (make -C oracle)"
    ref = po.ref_ops(1)
    out = {}
    for name, text in CYCLES.items():
        path = os.path.join(HERE, "cycle_%s.txt" % name)
        with open(path, "w") as f:
            f.write(text)
        r = po.run_cycle(path, ops=ref, threads=1)
        out[name] = dict(
            trace=r["trace"], N=r["N"], mg_error=r["mg_error"], sumU=r["sumU"], maxabsU=r["maxabsU"],
            sha256_U=hashlib.sha256(r["U"].tobytes()).hexdigest())
        if r["N"] <= 64:
            np.save(os.path.join(HERE, "cycle_U_%s.npy" % name), r["U"])
    with open(os.path.join(HERE, "cycles.json"), "w") as f:
        json.dump(out, f, indent=1)

    # real binary logs (run from tests/golden so that argv[2] is a bare file name)
    for name in BINARY_LOGS:
        p = subprocess.run([po.REF_BIN, "1", "cycle_%s.txt" % name], cwd=HERE, capture_output=True, text=True, check=True)
        log = "\n".join(l for l in p.stdout.splitlines() if not l.startswith("Time Used"))
        with open(os.path.join(HERE, "MG_CPU_%s.log" % name), "w") as f:
            f.write(log + "\n")
        csv = os.path.join(HERE, "Sol_CPU_cycle_%s.txt" % name)
        if name == "test" or name.startswith("mode_"):
            os.replace(csv, os.path.join(HERE, "MG_CPU_%s.csv" % name))
        else:
            os.remove(csv)

    # operator vectors
    for N, M_c, M_f in ((17, 9, 33), (32, 16, 64), (45, 21, 46)):
        rng = np.random.default_rng(1000 + N)
        U = rng.random(N * N)
        F = rng.random(N * N)
        Uz = U.copy().reshape(N, N)
        Uz[0, :] = Uz[-1, :] = 0
        Uz[:, 0] = Uz[:, -1] = 0
        Uz = Uz.reshape(-1)
        d = dict(U=U, F=F, Uz=Uz)
        d["source"] = ref.getSource(N)
        d["source_shift"] = ref.getSource(N, 2.0, -0.5, 0.25)
        d["analytic"] = ref.getAnalytic(N)
        d["residual"] = ref.getResidual(N, 1.0, U, F)
        d["add"] = ref.doGridAddition(N, U, F)
        for s in (1, 2, 3, 4, 7):
            a, e = ref.doSmoothing(N, 1.0, Uz, F, s)
            d["smooth%d" % s] = a
            d["smooth%d_err" % s] = np.array([e])
        d["restrict"] = ref.doRestriction(N, U, M_c)
        d["prolong"] = ref.doProlongation(N, U, M_f)
        d["gs"] = ref.doExactSolver(N, 1.0, F, 1e-7, 1)
        d["shape"] = np.array([N, M_c, M_f])
        np.savez_compressed(os.path.join(HERE, "ops_N%d.npz" % N), **d)
    print("golden written")


if __name__ == "__main__":
    main()
