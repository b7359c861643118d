"""ctypes access to the CPU checker (oracle/liboracle.so) and, when present, the
unmodified reference build (oracle/_ref/libmgref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
multigrid_poisson_solver_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmgref.so")
REF_BIN = os.path.join(HERE, "_ref", "MG_CPU")

_dp = C.POINTER(C.c_double)


def build(quiet=True):
    """(Re)build liboracle.so and, if /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class TraceRec(C.Structure):
    _fields_ = [("node", C.c_int), ("N", C.c_int), ("steps", C.c_int),
                ("err", C.c_double), ("sumU", C.c_double), ("maxabsU", C.c_double)]


class CycleResult(C.Structure):
    _fields_ = [("n_recs", C.c_int), ("N", C.c_int), ("mg_error", C.c_double),
                ("time_ms", C.c_double), ("sumU", C.c_double), ("maxabsU", C.c_double)]


_fn5 = C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, C.c_double, C.c_double)


class MgOps(C.Structure):
    _fields_ = [
        ("getSource", _fn5),
        ("getAnalytic", _fn5),
        ("getResidual", C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, _dp)),
        ("doGridAddition", C.CFUNCTYPE(None, C.c_int, _dp, _dp)),
        ("doSmoothing", C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, C.c_int, _dp)),
        ("doExactSolver", C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, C.c_double, C.c_int)),
        ("doRestriction", C.CFUNCTYPE(None, C.c_int, _dp, C.c_int, _dp)),
        ("doProlongation", C.CFUNCTYPE(None, C.c_int, _dp, C.c_int, _dp)),
    ]


SNAP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, _dp)


class Ops:
    """The eight operators (+getAnalytic) on numpy arrays, backed by a C library
    whose symbols are `<prefix>getSource`, ...  prefix 'orc_' = restatement,
    'ref_' = the unmodified reference."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        sig = {
            "getSource": (C.c_int, C.c_double, _dp, C.c_double, C.c_double),
            "getBoundary": (C.c_int, C.c_double, _dp, C.c_double, C.c_double),
            "getAnalytic": (C.c_int, C.c_double, _dp, C.c_double, C.c_double),
            "getResidual": (C.c_int, C.c_double, _dp, _dp, _dp),
            "doGridAddition": (C.c_int, _dp, _dp),
            "doSmoothing": (C.c_int, C.c_double, _dp, _dp, C.c_int, _dp),
            "doExactSolver": (C.c_int, C.c_double, _dp, _dp, C.c_double, C.c_int),
            "doRestriction": (C.c_int, _dp, C.c_int, _dp),
            "doProlongation": (C.c_int, _dp, C.c_int, _dp),
        }
        self._f = {}
        for name, args in sig.items():
            f = getattr(lib, prefix + name)
            f.argtypes, f.restype = args, None
            self._f[name] = f

    # -- numpy-level wrappers; all arrays are flat float64 of N*N ------------
    def getSource(self, N, L=1.0, min_x=0.0, min_y=0.0):
        F = np.empty(N * N)
        self._f["getSource"](N, L, _ptr(F), min_x, min_y)
        return F

    def getBoundary(self, N, L=1.0, min_x=0.0, min_y=0.0):
        F = np.full(N * N, 7.0)
        self._f["getBoundary"](N, L, _ptr(F), min_x, min_y)
        return F

    def getAnalytic(self, N, L=1.0, min_x=0.0, min_y=0.0):
        U = np.empty(N * N)
        self._f["getAnalytic"](N, L, _ptr(U), min_x, min_y)
        return U

    def getResidual(self, N, L, U, F):
        D = np.empty(N * N)
        self._f["getResidual"](N, L, _ptr(U), _ptr(F), _ptr(D))
        return D

    def doGridAddition(self, N, U1, U2):
        out = U1.copy()
        self._f["doGridAddition"](N, _ptr(out), _ptr(U2))
        return out

    def doSmoothing(self, N, L, U, F, step):
        out = U.copy()
        err = C.c_double(0.0)
        self._f["doSmoothing"](N, L, _ptr(out), _ptr(F), step, C.byref(err))
        return out, err.value

    def doExactSolver(self, N, L, F, target, option):
        U = np.full(N * N, 3.0)
        self._f["doExactSolver"](N, L, _ptr(U), _ptr(F), target, option)
        return U

    def doRestriction(self, N, U_f, M):
        U_c = np.full(M * M, 5.0)
        self._f["doRestriction"](N, _ptr(U_f), M, _ptr(U_c))
        return U_c

    def doProlongation(self, N, U_c, M, fill=np.nan):
        U_f = np.full(M * M, fill)
        self._f["doProlongation"](N, _ptr(U_c), M, _ptr(U_f))
        return U_f

    def ops_table(self):
        t = MgOps()
        for name, ftype in MgOps._fields_:
            addr = C.cast(getattr(self.lib, self.prefix + name), C.c_void_p).value
            setattr(t, name, ftype(addr))
        return t


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_run_cycle.argtypes = [C.c_char_p, C.POINTER(MgOps), C.c_int, C.POINTER(TraceRec), C.c_int,
                                      SNAP_FN, C.c_void_p, C.POINTER(_dp), C.POINTER(CycleResult)]
        lib.orc_run_cycle.restype = C.c_int
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_last_gs_iterations.restype = C.c_int
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_set_threads.argtypes = [C.c_int]
        lib.ref_main.argtypes = [C.c_int, C.c_char_p]
        lib.ref_main.restype = C.c_int
        _ref = lib
    return _ref


def oracle_ops():
    return Ops(oracle_lib(), "orc_")


def ref_ops(threads=1):
    lib = ref_lib()
    lib.ref_set_threads(threads)
    return Ops(lib, "ref_")


def run_cycle(path, ops=None, threads=1, max_recs=4096, want_U=True, snapshots=False):
    """Run a Cycle.txt with the oracle's driver.  ops=None -> oracle operators,
    else an Ops (e.g. ref_ops()).  Returns dict(trace=[...], U=ndarray|None, ...)."""
    lib = oracle_lib()
    recs = (TraceRec * max_recs)()
    res = CycleResult()
    uptr = _dp()
    snaps = []

    def _snap(ctx, idx, node, N, U):
        snaps.append((idx, node, N, np.ctypeslib.as_array(U, shape=(N * N,)).copy()))

    cb = SNAP_FN(_snap) if snapshots else SNAP_FN(0)
    table = ops.ops_table() if ops is not None else None
    rc = lib.orc_run_cycle(os.fsencode(path), C.byref(table) if table is not None else None, threads,
                           recs, max_recs, cb, None, C.byref(uptr) if want_U else None, C.byref(res))
    if rc != 0:
        raise RuntimeError("orc_run_cycle(%s) failed with code %d" % (path, rc))
    U = None
    if want_U:
        U = np.ctypeslib.as_array(uptr, shape=(res.N * res.N,)).copy()
        lib.orc_free(uptr)
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err, sumU=r.sumU, maxabsU=r.maxabsU)
             for r in recs[:res.n_recs]]
    return dict(trace=trace, U=U, N=res.N, mg_error=res.mg_error, time_ms=res.time_ms,
                sumU=res.sumU, maxabsU=res.maxabsU, snapshots=snaps)
