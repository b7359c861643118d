"""ctypes binding of include/mg_abi.h (libmgb200.so)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, os.environ.get("MG_LIB_NAME", "libmgb200.so"))   # MG_LIB_NAME: a tuning variant built next to the product library

RUN_UNFUSED, RUN_FUSED, RUN_QUIET, RUN_SKIP_SOURCE, RUN_NO_FINAL_ERROR = 0, 1, 2, 4, 8
SCALAR_SLOTS = 4096

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class MGLibraryError(RuntimeError):
    pass


class TraceRec(C.Structure):
    _fields_ = [("node", C.c_int), ("N", C.c_int), ("steps", C.c_int), ("err", C.c_double)]


class CycleResult(C.Structure):
    _fields_ = [("n_recs", C.c_int), ("N", C.c_int), ("mg_error", C.c_double), ("time_ms", C.c_double),
                ("wall_ms", C.c_double), ("launches", C.c_int)]


# every symbol include/mg_abi.h declares: name -> (restype, argtypes)
ABI = {
    "mgInit": (C.c_int, [C.c_int]),
    "mgShutdown": (None, []),
    "mgLastErrorCode": (C.c_int, []),
    "mgLastError": (C.c_char_p, []),
    "mgClearError": (None, []),
    "mgSync": (None, []),
    "mgStream": (_vp, []),
    "mgKernelLaunches": (C.c_int, []),
    "mgGridAlloc": (_vp, [C.c_int]),
    "mgGridFree": (None, [_vp]),
    "mgGridZero": (None, [C.c_int, _vp]),
    "mgGridNegate": (None, [C.c_int, _vp]),
    "mgGridUpload": (None, [C.c_int, _vp, _vp]),
    "mgGridDownload": (None, [C.c_int, _vp, _vp]),
    "mgGridCopy": (None, [C.c_int, _vp, _vp]),
    "getSource": (None, [C.c_int, C.c_double, _vp, C.c_double, C.c_double]),
    "getBoundary": (None, [C.c_int, C.c_double, _vp, C.c_double, C.c_double]),
    "getAnalytic": (None, [C.c_int, C.c_double, _vp, C.c_double, C.c_double]),
    "getResidual": (None, [C.c_int, C.c_double, _vp, _vp, _vp]),
    "doSmoothing": (None, [C.c_int, C.c_double, _vp, _vp, C.c_int, _dp]),
    "doRestriction": (None, [C.c_int, _vp, C.c_int, _vp]),
    "doProlongation": (None, [C.c_int, _vp, C.c_int, _vp]),
    "doGridAddition": (None, [C.c_int, _vp, _vp]),
    "doExactSolver": (None, [C.c_int, C.c_double, _vp, _vp, C.c_double, C.c_int]),
    "mgAnalyticError": (C.c_double, [C.c_int, C.c_double, _vp, C.c_double, C.c_double]),
    "mgLastExactSolverIterations": (C.c_int, []),
    "mgScalarSlot": (_dp, [C.c_int]),
    "mgSmooth": (None, [C.c_int, C.c_double, _vp, _vp, C.c_int, _vp, _dp]),
    "mgDownLeg": (_vp, [C.c_int, C.c_double, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _dp]),
    "mgExactSolve": (None, [C.c_int, C.c_double, _vp, _vp, C.c_double, C.c_int, _dp]),
    "mgUpLeg": (_vp, [C.c_int, _vp, C.c_int, C.c_double, _vp, _vp, _vp, C.c_int, _dp]),
    "mgDownLegTrigger": (_vp, [C.c_int, C.c_double, _vp, _vp, _vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_int), _dp]),
    "mgUpLegTrigger": (_vp, [C.c_int, _vp, C.c_int, C.c_double, _vp, _vp, _vp, C.POINTER(C.c_int), _dp]),
    "mgCoarseTailMaxN": (C.c_int, []),
    "mgCoarseTailMaxOps": (C.c_int, []),
    "mgCoarseTail": (C.c_int, [C.c_double, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dp]),
    "mgRunCycleFile": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, C.POINTER(TraceRec), C.c_int, C.POINTER(CycleResult)]),
    "mgMeanAbsDiff": (C.c_double, [C.c_int, _vp, _vp]),
    "mgRunCycleFileEx": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, C.POINTER(TraceRec), C.c_int, C.POINTER(CycleResult)]),
    "mgRunCycleFileHostEx": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(TraceRec), C.c_int, C.POINTER(CycleResult)]),
    "mgRunSubcycle": (C.c_int, [_dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                C.c_double, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _vp, C.c_int, C.POINTER(C.c_int), C.c_int,
                                C.POINTER(TraceRec), C.c_int, C.POINTER(C.c_int), C.c_int]),
    "mgSubcycleHarvest": (None, []),
    "mgRunCycleFileHost": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, C.POINTER(TraceRec), C.c_int, C.POINTER(CycleResult)]),
    "mgRunCycleFileHostBatch": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(CycleResult)]),
    "mgPrint2File": (C.c_int, [C.c_int, _vp, C.c_char_p]),
    "mgSetTileMaxN": (C.c_int, [C.c_int]),
    "mgSegmentPlan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "mgDistPlan": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "mgDistUniqueId": (C.c_int, [_vp]),
    "mgDistInit": (C.c_int, [C.c_int, C.c_int, _vp]),
    "mgDistShutdown": (None, []),
    "mgDistSmoothStress": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mgDistSourceSlab": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mgDistUploadSource": (C.c_int, [C.c_int, C.c_int, _vp]),
    "mgDistDownloadSource": (C.c_int, [C.c_int, _vp]),
    "mgDistRunCycleFile": (C.c_int, [C.c_char_p, C.c_int, C.c_int, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(TraceRec), C.c_int, C.POINTER(CycleResult)]),
    "mgDistEmuRunCycleFile": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, _vp, _vp, C.POINTER(TraceRec), C.c_int,
                                        C.POINTER(CycleResult)]),
    "mgDistRunCycleFileHostBatch": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp),
                                              C.POINTER(CycleResult)]),
}

_lib = None
_ready = False


def lib_path():
    return _LIB_PATH


def lib():
    """The loaded library (symbols typed).  Loading needs no GPU; computing does."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise MGLibraryError("%s is missing: run `python -m multigrid_poisson_solver_b200.build` "
                                 "(there is no CPU fallback)" % _LIB_PATH)
        l = C.CDLL(_LIB_PATH)
        for name, (res, args) in ABI.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def _check():
    l = lib()
    if l.mgLastErrorCode() != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError(msg)


def init(device=0):
    """Create the library context on a CUDA device.  Raises if there is none."""
    global _ready
    l = lib()
    if l.mgInit(int(device)) != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("mgInit(%d) failed: %s" % (device, msg))
    _ready = True
    return l


def _need():
    if not _ready:
        init(int(os.environ.get("LOCAL_RANK", "0")))
    return lib()


class DeviceGrid:
    """An N x N fp64 grid resident on the device (pooled by the library)."""

    def __init__(self, N, host=None):
        l = _need()
        self.N = int(N)
        self.ptr = l.mgGridAlloc(self.N)
        _check()
        if host is not None:
            self.upload(host)

    def upload(self, host):
        a = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
        assert a.size == self.N * self.N
        lib().mgGridUpload(self.N, self.ptr, a.ctypes.data)
        lib().mgSync()
        _check()
        return self

    def numpy(self):
        out = np.empty(self.N * self.N)
        lib().mgGridDownload(self.N, self.ptr, out.ctypes.data)
        _check()
        return out

    def free(self):
        if self.ptr:
            lib().mgGridFree(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GpuOps:
    """The reference's operators on numpy arrays, executed by libmgb200 on the GPU.
    Same call shapes as oracle.pyoracle.Ops so parity tests read symmetrically."""

    def __init__(self, device=None):
        if device is not None:
            init(device)
        self.l = _need()

    def getSource(self, N, L=1.0, min_x=0.0, min_y=0.0):
        F = DeviceGrid(N)
        self.l.getSource(N, L, F.ptr, min_x, min_y)
        return F.numpy()

    def getBoundary(self, N, L=1.0, min_x=0.0, min_y=0.0):
        F = DeviceGrid(N, np.full(N * N, 7.0))
        self.l.getBoundary(N, L, F.ptr, min_x, min_y)
        return F.numpy()

    def getAnalytic(self, N, L=1.0, min_x=0.0, min_y=0.0):
        U = DeviceGrid(N)
        self.l.getAnalytic(N, L, U.ptr, min_x, min_y)
        return U.numpy()

    def getResidual(self, N, L, U, F):
        dU, dF, dD = DeviceGrid(N, U), DeviceGrid(N, F), DeviceGrid(N)
        self.l.getResidual(N, L, dU.ptr, dF.ptr, dD.ptr)
        return dD.numpy()

    def doGridAddition(self, N, U1, U2):
        a, b = DeviceGrid(N, U1), DeviceGrid(N, U2)
        self.l.doGridAddition(N, a.ptr, b.ptr)
        return a.numpy()

    def doSmoothing(self, N, L, U, F, step):
        dU, dF = DeviceGrid(N, U), DeviceGrid(N, F)
        err = C.c_double(-1.0)
        self.l.doSmoothing(N, L, dU.ptr, dF.ptr, step, C.byref(err))
        _check()
        return dU.numpy(), err.value

    def doExactSolver(self, N, L, F, target, option):
        dU, dF = DeviceGrid(N, np.full(N * N, 3.0)), DeviceGrid(N, F)
        self.l.doExactSolver(N, L, dU.ptr, dF.ptr, target, option)
        return dU.numpy()

    def last_exact_solver_iterations(self):
        return self.l.mgLastExactSolverIterations()

    def doRestriction(self, N, U_f, M):
        f, c = DeviceGrid(N, U_f), DeviceGrid(M, np.full(M * M, 5.0))
        self.l.doRestriction(N, f.ptr, M, c.ptr)
        return c.numpy()

    def doProlongation(self, N, U_c, M, fill=np.nan):
        c, f = DeviceGrid(N, U_c), DeviceGrid(M, np.full(M * M, fill))
        self.l.doProlongation(N, c.ptr, M, f.ptr)
        return f.numpy()

    def negate(self, N, D):
        d = DeviceGrid(N, D)
        self.l.mgGridNegate(N, d.ptr)
        return d.numpy()

    def analytic_error(self, N, L, U, min_x=0.0, min_y=0.0):
        d = DeviceGrid(N, U)
        return self.l.mgAnalyticError(N, L, d.ptr, min_x, min_y)

    # ---- fused entry points
    def smooth(self, N, L, U, F, step):
        dU, dF, dO = DeviceGrid(N, U), DeviceGrid(N, F), DeviceGrid(N)
        slot = self.l.mgScalarSlot(7)
        self.l.mgSmooth(N, L, dU.ptr, dF.ptr, step, dO.ptr, slot)
        self.l.mgSync()
        _check()
        assert np.array_equal(dU.numpy(), np.asarray(U, dtype=np.float64).reshape(-1)), "mgSmooth touched U_in"
        return dO.numpy(), slot[0]

    def down_leg(self, N, L, U, F, step, zero_init, M):
        dU, dW, dF, dFc = DeviceGrid(N, U), DeviceGrid(N), DeviceGrid(N, F), DeviceGrid(M, np.full(M * M, 9.0))
        slot = self.l.mgScalarSlot(8)
        res = self.l.mgDownLeg(N, L, dU.ptr, dW.ptr, dF.ptr, step, int(zero_init), M, dFc.ptr, slot)
        self.l.mgSync()
        _check()
        out = (dU if res == dU.ptr else dW).numpy()
        return out, slot[0], dFc.numpy()

    def up_leg(self, Nc, U_c, N, L, U_f, F, step):
        dC, dU, dW, dF = DeviceGrid(Nc, U_c), DeviceGrid(N, U_f), DeviceGrid(N), DeviceGrid(N, F)
        slot = self.l.mgScalarSlot(9)
        res = self.l.mgUpLeg(Nc, dC.ptr, N, L, dU.ptr, dW.ptr, dF.ptr, step, slot)
        self.l.mgSync()
        _check()
        out = (dU if res == dU.ptr else dW).numpy()
        return out, slot[0]


def _run(fn_name, path, flags, F, U_out_N, max_recs):
    l = _need()
    recs = (TraceRec * max_recs)()
    res = CycleResult()
    U = np.empty(U_out_N * U_out_N) if U_out_N else None
    rc = getattr(l, fn_name)(os.fsencode(path), flags, F, U.ctypes.data if U is not None else None,
                             recs, max_recs, C.byref(res))
    if rc != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("%s(%s) failed with code %d %s" % (fn_name, path, rc, msg))
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]
    return dict(trace=trace, U=U, N=res.N, mg_error=res.mg_error, time_ms=res.time_ms, wall_ms=res.wall_ms,
                launches=res.launches)


def _n_max(path):
    with open(path) as f:
        return int(f.read().split()[5])


def run_cycle_host(path, flags=RUN_FUSED | RUN_QUIET, F_host=None, want_U=True, max_recs=8192):
    """mgRunCycleFileHost: host F in (optional), host U out."""
    N = _n_max(path)
    Fp = None
    if F_host is not None:
        F_host = np.ascontiguousarray(F_host, dtype=np.float64).reshape(-1)
        assert F_host.size == N * N
        Fp = F_host.ctypes.data
    return _run("mgRunCycleFileHost", path, flags, Fp, N if want_U else 0, max_recs)


def run_cycle_problem(path, flags=RUN_FUSED | RUN_QUIET, F_host=None, U0_host=None, analytic_host=None, max_recs=8192):
    """mgRunCycleFileHostEx: the problem plug point -- caller-supplied source, initial grid with (non-zero) Dirichlet boundary
    data, and reference solution of the final error report."""
    l = _need()
    N = _n_max(path)
    keep = []

    def ptr(a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        assert a.size == N * N
        keep.append(a)
        return a.ctypes.data

    recs = (TraceRec * max_recs)()
    res = CycleResult()
    U = np.empty(N * N)
    rc = l.mgRunCycleFileHostEx(os.fsencode(path), flags, ptr(F_host), ptr(U0_host), ptr(analytic_host), U.ctypes.data, recs, max_recs,
                                C.byref(res))
    if rc != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("mgRunCycleFileHostEx(%s) failed with code %d %s" % (path, rc, msg))
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]
    return dict(trace=trace, U=U, N=res.N, mg_error=res.mg_error, time_ms=res.time_ms, launches=res.launches)


def run_cycle_host_batch(path, F_hosts, flags=RUN_FUSED | RUN_QUIET, U_ptrs=None):
    """mgRunCycleFileHostBatch: independent problems (one host source each) through the same cycle
    file; uploads, cycles and downloads of consecutive problems overlap.  F_hosts: arrays or raw
    host pointers (ints, e.g. of pinned torch tensors); U_ptrs: raw host pointers to fill, else
    numpy arrays are allocated and returned."""
    l = _need()
    N = _n_max(path)
    n = len(F_hosts)
    keep, Fp = [], (C.c_void_p * n)()
    for i, F in enumerate(F_hosts):
        if isinstance(F, int):
            Fp[i] = F
        else:
            a = np.ascontiguousarray(F, dtype=np.float64).reshape(-1)
            assert a.size == N * N
            keep.append(a)
            Fp[i] = a.ctypes.data
    Up, Us = (C.c_void_p * n)(), None
    if U_ptrs is None:
        Us = [np.empty(N * N) for _ in range(n)]
        for i, U in enumerate(Us):
            Up[i] = U.ctypes.data
    else:
        for i, u in enumerate(U_ptrs):
            Up[i] = u
    res = (CycleResult * n)()
    rc = l.mgRunCycleFileHostBatch(os.fsencode(path), flags, n, Fp, Up, res)
    if rc != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("mgRunCycleFileHostBatch(%s) failed with code %d %s" % (path, rc, msg))
    return dict(U=Us, mg_error=[r.mg_error for r in res], time_ms=[r.time_ms for r in res], launches=[r.launches for r in res])


def run_cycle(path, flags=RUN_FUSED | RUN_QUIET, max_recs=8192):
    """mgRunCycleFile with the source generated on the device; the solution stays on the device."""
    return _run("mgRunCycleFile", path, flags, None, 0, max_recs)


def run_cycle_dist_emulated(path, world, threshold, flags=RUN_FUSED | RUN_QUIET, max_recs=8192, F_host=None):
    """mgDistEmuRunCycleFile: the row-slab multi-GPU algorithm with all ranks emulated on this GPU.
    F_host: the source grid (default: getSource on the device)."""
    l = _need()
    N = _n_max(path)
    recs = (TraceRec * max_recs)()
    res = CycleResult()
    U = np.empty(N * N)
    Fp = None
    if F_host is not None:
        F_host = np.ascontiguousarray(F_host, dtype=np.float64).reshape(-1)
        assert F_host.size == N * N
        Fp = F_host.ctypes.data
    rc = l.mgDistEmuRunCycleFile(os.fsencode(path), world, threshold, flags, Fp, U.ctypes.data, recs, max_recs, C.byref(res))
    if rc != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("mgDistEmuRunCycleFile(%s) failed with code %d %s" % (path, rc, msg))
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]
    return dict(trace=trace, U=U, N=res.N, mg_error=res.mg_error, time_ms=res.time_ms, wall_ms=res.wall_ms,
                launches=res.launches)


def dist_init(rank, world, broadcast_bytes):
    """NCCL rendezvous for the slab driver.  `broadcast_bytes(buf_or_None) -> bytes` must return
    rank 0's 128-byte id on every rank (e.g. via torch.distributed.broadcast_object_list)."""
    l = _need()
    buf = (C.c_ubyte * 128)()
    if rank == 0 and l.mgDistUniqueId(buf) != 0:
        _check()
    data = broadcast_bytes(bytes(buf) if rank == 0 else None)
    buf2 = (C.c_ubyte * 128).from_buffer_copy(data)
    if l.mgDistInit(rank, world, buf2) != 0:
        _check()
        raise MGLibraryError("mgDistInit failed")


def run_cycle_dist(path, threshold, flags=RUN_FUSED | RUN_QUIET, want_U=False, max_recs=8192):
    """mgDistRunCycleFile (collective).  Returns this rank's owned rows when want_U."""
    l = _need()
    N = _n_max(path)
    recs = (TraceRec * max_recs)()
    res = CycleResult()
    lo, hi = C.c_int(0), C.c_int(0)
    U = np.empty(N * N) if want_U else None
    rc = l.mgDistRunCycleFile(os.fsencode(path), threshold, flags, U.ctypes.data if want_U else None, C.byref(lo), C.byref(hi),
                              recs, max_recs, C.byref(res))
    if rc != 0:
        msg = l.mgLastError().decode()
        l.mgClearError()
        raise MGLibraryError("mgDistRunCycleFile(%s) failed with code %d %s" % (path, rc, msg))
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]
    out = dict(trace=trace, N=res.N, mg_error=res.mg_error, time_ms=res.time_ms, wall_ms=res.wall_ms, launches=res.launches,
               own=(lo.value, hi.value))
    if want_U:
        out["U_own"] = U[: (hi.value - lo.value) * N].copy()
    return out


def segment_plan(rows, n_strips, resident_warps, lead_rows=9, subset=0):
    """Row segments [(first, past_last), ...] of a fused pass in queue order (host-only)."""
    l = lib()
    out = (C.c_int * (2 * rows + 2))()
    n = l.mgSegmentPlan(rows, n_strips, resident_warps, lead_rows, subset, out, len(out))
    if n < 0:
        raise MGLibraryError("mgSegmentPlan failed with code %d" % n)
    return [(out[2 * k], out[2 * k + 1]) for k in range(n)]


def dist_plan(ladder, world, threshold):
    """Slab geometry of every level of `ladder` (host-only): list of dict(N, dist, bounds)."""
    l = lib()
    n = len(ladder)
    arr = (C.c_int * n)(*ladder)
    out = (C.c_int * (n * (world + 3)))()
    rc = l.mgDistPlan(arr, n, world, threshold, out, len(out))
    if rc < 0:
        raise MGLibraryError("mgDistPlan failed with code %d" % rc)
    plan = []
    for i in range(n):
        o = out[i * (world + 3):(i + 1) * (world + 3)]
        plan.append(dict(N=o[0], dist=bool(o[1]), bounds=list(o[2:])))
    return plan
