#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: V-cycles/sec at N=16384 fp64 (+ smoother HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--nmax 16384]

One "step" = one V-cycle of the shipped Vcycle.txt shape (con_step=3, con_N=1, N_min=8,
Gauss-Seidel 1e-7 at the coarsest grid) over the analytic source grid ("synthetic": generated
by getSource, no dataset), run by the C++ cycle driver of libmgb200 through the C ABI.

 value  whole-job V-cycles/s with the source grid already resident in HBM
 e2e    the same V-cycle through mgRunCycleFileHost with HOST buffers: the source grid is copied
        H2D from pinned memory and the solution D2H inside the timed region, every step
 roofline      the dominant kernel timed alone with CUDA events on the library's stream
 cpu_baseline  the unmodified reference operators (oracle/_ref) on the host cores, bounded sample

--impl reference runs the reference's own CPU implementation (oracle/_ref operators driven by
the oracle's restatement of main()) with all host threads on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "V-cycles/sec at N=16384 fp64"
UNIT = "V-cycles/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def write_cycle(text):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", prefix="cycle_", delete=False)
    f.write(text)
    f.close()
    return f.name


class ClockSampler:
    """SM clock + throttle reasons sampled every ~2 ms WHILE the timed region runs (NVML from a
    thread: the timed loop sits in ctypes calls, which release the GIL).  nvidia-smi -lms is the
    fallback when NVML cannot be loaded; it needs ~100 ms to start, so short regions may get no sample."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None
        self.nvml, self.handle, self.thread, self.stop = None, None, None, False
        self.samples, self.max_mhz = [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                pr = torch.cuda.get_device_properties(gpu)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv, h = self.nvml, self.handle
        while not self.stop:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nvml:
            import threading
            self.stop = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.thread:
            self.stop = True
            self.thread.join(timeout=2)
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.nvml:
            if self.samples:
                bits = 0
                for _, r in self.samples:
                    bits |= r
                out = {"sm_mhz": statistics.median(s for s, _ in self.samples), "sm_max_mhz": self.max_mhz,
                       "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(self.samples),
                       "source": "NVML, every ~2 ms inside the timed region"}
            return out
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi -lms 20 inside the timed region"
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------- workloads
def workloads():
    """BASELINE.json configs as bench workloads.  `N` = N_max, `sample_N` = N_max of the CPU reference run
    (equal to N when the reference can run the real thing in bounded time and memory)."""
    from multigrid_poisson_solver_b200 import cycles
    return {
        # configs[1]'s shape at the size the metric is quoted on: the headline
        "v16384": dict(N=16384, text=lambda n: cycles.v_cycle(n, 8), sample_N=16384, metric=METRIC,
                       what="Vcycle.txt shape (con_step=3, con_N=1, GS 1e-7 opt 1) at N_max=%d N_min=8"),
        "v8192": dict(N=8192, text=lambda n: cycles.v_cycle(n, 8), sample_N=8192, metric="V-cycles/sec at N=8192 fp64",
                      what="BASELINE config 2: Vcycle.txt shape (con_step=3, con_N=1) at N_max=%d N_min=8"),
        "w16384": dict(N=16384, text=lambda n: cycles.w_cycle(n, 16, tol=1e-8), sample_N=4096, metric="W-cycles/sec at N=16384 fp64",
                       what="BASELINE config 3: Wcycle.txt recursion over the whole ladder N_max=%d ... 16 (GS 1e-8 at N=16)"),
        "trigger32768": dict(N=32768, text=lambda n: cycles.v_cycle(n, 8, step=-1), sample_N=8192,
                             metric="error-trigger V-cycles/sec at N=32768 fp64",
                             what="BASELINE config 4: VcycleTrigger.txt shape (con_step=-1) at N_max=%d N_min=8"),
        "trigger16384": dict(N=16384, text=lambda n: cycles.v_cycle(n, 8, step=-1), sample_N=8192,
                             metric="error-trigger V-cycles/sec at N=16384 fp64",
                             what="VcycleTrigger.txt shape (con_step=-1) at N_max=%d N_min=8"),
    }


# ----------------------------------------------------------------------------- reference arm (CPU)
def stock_main_run(text, threads, timeout_s=1200):
    """Runs the UNMODIFIED reference program (oracle/_ref/MG_CPU = g++ -fopenmp, no -O, of
    /root/reference/src/MG_solver_CPU.cpp + linkedlist.cpp, exactly src/Makefile:8) on a cycle file and
    reads ITS OWN report: `Time Used` = omp_get_wtime() span of the node loop (MG_solver_CPU.cpp:156,
    :429-451) and the final `Error`.  The program then dumps the solution as CSV (2.4 GB of fprintf at
    N = 16384, :454-457); that dump is not part of the metric, so the process is stopped once the report
    line has been read (stdout is made line-buffered with stdbuf, or a pty when stdbuf is missing) and
    the output name is pre-linked to /dev/null in a scratch directory that is deleted afterwards."""
    import shutil
    from oracle import pyoracle as po
    if not os.path.exists(po.REF_BIN):
        return None
    d = tempfile.mkdtemp(prefix="mgref_")
    try:
        with open(os.path.join(d, "c.txt"), "w") as f:
            f.write(text)
        os.symlink("/dev/null", os.path.join(d, "Sol_CPU_c.txt"))
        cmd = [po.REF_BIN, str(threads), "c.txt"]
        master = None
        if shutil.which("stdbuf"):
            proc = subprocess.Popen(["stdbuf", "-oL"] + cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            stream = proc.stdout
        else:
            import pty
            master, slave = pty.openpty()
            proc = subprocess.Popen(cmd, cwd=d, stdout=slave, stderr=subprocess.DEVNULL)
            os.close(slave)
            stream = os.fdopen(master, "r", errors="replace")
        out = {"errors": [], "time_ms": None, "mg_error": None}
        t0 = time.time()
        final = False
        try:
            for line in stream:
                line = line.strip()
                if line.startswith("===== Final Result"):
                    final = True
                elif line.startswith("Error ="):
                    v = float(line.split("=")[1])
                    if final:
                        out["mg_error"] = v
                    else:
                        out["errors"].append(v)
                elif line.startswith("Time Used ="):
                    out["time_ms"] = float(line.split("=")[1].split()[0])
                    break
                if time.time() - t0 > timeout_s:
                    break
        except OSError:
            pass
        proc.kill()                      # our own child, by PID: skips the CSV dump
        proc.wait()
        out["wall_s"] = time.time() - t0
        return out if out["time_ms"] is not None else None
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_reference(wl, threads, runs):
    """`runs` timed cycles of workload `wl` on the host cores.  Returns dict(kind, ms=[...] per cycle AT wl['N'],
    sample=text, mg_error).  kind "reference": the stock main(); when the workload is too big for a bounded CPU run the
    same cycle shape is run at sample_N and scaled by the DOF ratio (a cycle's work is linear in N^2), which the sample text says.
    kind "port" (oracle restatement, -O2) only if oracle/_ref is missing."""
    from oracle import pyoracle as po
    N, sN = wl["N"], wl["sample_N"]
    scale = (N / sN) ** 2
    ms, err, kind = [], None, "reference"
    for _ in range(runs):
        r = stock_main_run(wl["text"](sN), threads)
        if r is None:
            kind = "port"
            path = write_cycle(wl["text"](sN))
            q = po.run_cycle(path, threads=threads, want_U=False)
            os.unlink(path)
            r = {"time_ms": q["time_ms"], "mg_error": q["mg_error"]}
        ms.append(r["time_ms"] * scale)
        err = r["mg_error"]
    what = ("stock MG_CPU main() (reference sources compiled -O0 -fopenmp as src/Makefile:8), %d OpenMP threads, its own `Time Used` "
            "(omp_get_wtime span of the node loop, MG_solver_CPU.cpp:156,:429-451); CSV dump skipped" % threads) if kind == "reference" else \
           ("oracle restatement of main() (-O2), %d threads, the reference's timer span" % threads)
    if sN == N:
        sample = "%s; the whole workload: %s, %d cycle(s) run" % (what, wl["what"] % N, runs)
    else:
        sample = "%s; bounded sample: the same cycle shape at N_max=%d (%d cycle(s)), scaled to N_max=%d by the DOF ratio %.0f" % (
            what, sN, runs, N, scale)
    return dict(kind=kind, ms=ms, sample=sample, mg_error=err, same_config=(sN == N))


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the arm's workload.  The stock main()
    at N = 16384 needs ~15-45 s per cycle, so the run is capped at 1 warm-up + 2 timed cycles whatever
    --steps/--warmup ask for (the line says so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    wl = workloads()[args.config]
    if args.nmax:
        wl["N"] = args.nmax
        wl["sample_N"] = min(wl["sample_N"], args.nmax)
    warm, timed = min(args.warmup, 1), max(1, min(args.steps, 2))
    world = max(1, args.gpus)
    r = cpu_reference(wl, threads, warm + timed)
    ms = r["ms"][warm:]
    ms_per_cycle = sum(ms) / len(ms)
    value = 1000.0 / ms_per_cycle
    cfg = {"workload": wl["what"] % wl["N"], "steps_run": timed, "warmup_run": warm,
           "steps_note": "capped at 1 warm-up + 2 timed cycles of the stock main() (each 15-45 s at N=16384); requested --steps %d --warmup %d"
                         % (args.steps, args.warmup)}
    if world > 1:
        cfg["multi_gpu_note"] = ("the %d-GPU arm's value is in units of the 1-GPU workload (16384^2 fine points per GPU, weak scaling); a "
                                 "V-cycle's work is linear in the DOF count, so the reference's value in the same unit is its N=16384 rate" % world)
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
        "warmup": warm, "ms_per_step": ms_per_cycle, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fine_dof_cycles_per_s": value * wl["N"] ** 2, "mg_error": r["mg_error"], "ms_per_cycle_runs": r["ms"],
        "same_config": r["same_config"],
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="v16384", choices=["v16384", "v8192", "w16384", "trigger32768", "trigger16384"],
                    help="workload (BASELINE.json configs); the default is the one the metric is quoted on")
    ap.add_argument("--nmax", type=int, default=0, help="override N_max of the workload")
    ap.add_argument("--unfused", action="store_true", help="one ABI operator per reference call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import multigrid_poisson_solver_b200 as mg
    from multigrid_poisson_solver_b200 import api, cycles

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = mg.init(local)
    stream = torch.cuda.ExternalStream(lib.mgStream(), device=local)

    def barrier():
        lib.mgSync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src = peaks()
    if world > 1:
        return multi_gpu_arm(args, mg, api, cycles, lib, torch, dist, stream, rank, world, local, barrier, max_over_ranks,
                             hbm_peak, peak_src)

    wl = workloads()[args.config]
    if args.nmax:
        wl["N"] = args.nmax
        wl["sample_N"] = min(wl["sample_N"], args.nmax)
    N = wl["N"]
    n = N * N
    path = write_cycle(wl["text"](N))
    max_recs = 8192
    base_flags = (mg.RUN_UNFUSED if args.unfused else mg.RUN_FUSED) | mg.RUN_QUIET | mg.RUN_NO_FINAL_ERROR
    flags = base_flags | mg.RUN_SKIP_SOURCE

    # ---- device-resident run -------------------------------------------------------------
    F = mg.DeviceGrid(N)
    lib.getSource(N, 1.0, F.ptr, 0.0, 0.0)
    recs = (api.TraceRec * max_recs)()
    res = api.CycleResult()

    def one_cycle():
        rc = lib.mgRunCycleFile(os.fsencode(path), flags, F.ptr, None, recs, max_recs, res)
        if rc != 0:
            raise SystemExit("mgRunCycleFile failed: %d %s" % (rc, lib.mgLastError().decode()))
        return res.launches, res.time_ms

    clk = ClockSampler(local)                       # NVML set up before the warm-up, sampling only inside the timed region
    for _ in range(args.warmup):
        one_cycle()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, inner_ms = 0, 0.0
    with clk:
        ev0.record(stream)
        for _ in range(args.steps):
            l, t = one_cycle()
            launches += l
            inner_ms += t
        ev1.record(stream)
        barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = clk.summary()
    ms_per_step = total_ms / args.steps
    value = world * 1000.0 / ms_per_step
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]

    # final error of the last cycle (outside the timed region, like the reference's report)
    check = mg.run_cycle(path, base_flags & ~mg.RUN_NO_FINAL_ERROR)
    mg_error = check["mg_error"]

    # ---- end to end through the host-buffer ABI calls -----------------------------------------
    # Headline: mgRunCycleFileHostBatch, one problem per step -- every step copies its source from
    # pinned host memory to the device, runs the V-cycle and reads the solution back into pinned
    # host memory; consecutive steps are double-buffered (upload i+1 / cycle i / download i-1).
    # `single_call`: the same through one mgRunCycleFileHost call per step (no overlap at all).
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        hF = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
        hU = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
        for h in hF:
            lib.mgGridDownload(N, F.ptr, h.data_ptr())

        def timed(fn, steps):
            fn(2)                                   # warm-up (also first-touch of the staging grids)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t0 = time.perf_counter()
            fn(steps)
            e1.record(stream)
            barrier()
            wall = time.perf_counter() - t0
            return max_over_ranks(max(e0.elapsed_time(e1), 1000.0 * wall)) / steps

        def single(k):
            for i in range(k):
                rc = lib.mgRunCycleFileHost(os.fsencode(path), base_flags, hF[i % 2].data_ptr(), hU[i % 2].data_ptr(), recs, max_recs, res)
                if rc != 0:
                    raise SystemExit("mgRunCycleFileHost failed: %d" % rc)

        def batch(k):
            Fp, Up, rs = (C.c_void_p * k)(), (C.c_void_p * k)(), (api.CycleResult * k)()
            for i in range(k):
                Fp[i], Up[i] = hF[i % 2].data_ptr(), hU[i % 2].data_ptr()
            rc = lib.mgRunCycleFileHostBatch(os.fsencode(path), base_flags, k, Fp, Up, rs)
            if rc != 0:
                raise SystemExit("mgRunCycleFileHostBatch failed: %d" % rc)

        single_ms = timed(single, max(2, min(args.steps, 3)))
        e2e_steps = max(2, min(args.steps, 10))
        e2e_ms = timed(batch, e2e_steps)
        e2e = {"value": world * 1000.0 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
               "ms_per_step": e2e_ms, "steps": e2e_steps,
               "call": "mgRunCycleFileHostBatch(cycle file, one problem per step: pinned F_host -> device, V-cycle, U -> pinned "
                       "U_host; upload of step i+1 and download of step i-1 overlap the cycle of step i)",
               "single_call": {"value": world * 1000.0 / single_ms, "ms_per_step": single_ms,
                               "call": "mgRunCycleFileHost per step (upload, V-cycle, download strictly in sequence)"}}
        del hF, hU

    # ---- roofline of the dominant kernel, timed alone on the library's stream -----------------
    roof = dominant_kernel_roofline(lib, mg, stream, torch, N, hbm_peak, peak_src, args.unfused)

    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (wl["what"] % N) + ", 1 grid per GPU", "name": args.config,
                   "driver": "unfused (8 ABI operators)" if args.unfused else "fused (mgDownLeg/mgUpLeg)",
                   "l2": "inputs exceed L2 (%.1f GiB per grid vs 126 MB)" % (8 * n / 2 ** 30),
                   "parallelism": "1 GPU" if world == 1 else "%d independent replicas (slab partition: see DESIGN.md)" % world},
        "fine_dof_cycles_per_s": value * n,
        "device_ms_per_cycle_node_loop": inner_ms / args.steps,
        "mg_error": mg_error, "trace_errors": [t["err"] for t in trace if t["node"] != 0][:48],
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roof,
        "cycle_roofline": {"unfused_algorithmic_bytes": 357.0 * n, "achieved_GBs": 357.0 * n / (ms_per_step * 1e6),
                           "note": "BASELINE.md 3: 357*N^2 B per unfused V-cycle; effective bandwidth may exceed HBM peak when legs are fused"},
    }

    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = cpu_reference(wl, threads, 1)
        line["cpu_baseline"] = {"value": 1000.0 / r["ms"][0], "unit": UNIT, "cores": threads, "kind": r["kind"], "sample": r["sample"],
                                "ms_per_cycle": r["ms"][0], "mg_error": r["mg_error"], "same_config": r["same_config"]}
    os.unlink(path)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def topology(local):
    """Where this rank's GPU hangs and which CPUs the process may use (NUMA placement of the pinned buffers)."""
    out = {"cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        out["gpu_pci"] = bdf
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            out["gpu_numa_node"] = int(f.read())
    except Exception:
        pass
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        out["numa_nodes"] = len(nodes)
    except Exception:
        pass
    return out


WEAK_N = {1: 16384, 2: 23168, 4: 32768, 8: 46336}   # N^2 per GPU constant (16384^2), even ladders down to the threshold


def multi_gpu_arm(args, mg, api, cycles, lib, torch, dist, stream, rank, world, local, barrier, max_over_ranks, hbm_peak, peak_src):
    """Weak scaling of the row-slab driver: the grid grows with the GPU count so that every GPU keeps
    16384^2 fine points; one process per GPU, halo rows by peer stores of the fused kernels, levels < 1024 rows redundant on every rank."""
    threshold = int(os.environ.get("MG_DIST_THRESHOLD", "1024"))
    wl = workloads()[args.config]
    weak = args.config == "v16384" and not args.nmax     # the headline: weak scaling; any other workload: the same grid on more GPUs
    N = args.nmax if args.nmax else (WEAK_N.get(world, int(round(16384 * world ** 0.5 / 256)) * 256) if weak else wl["N"])
    n = N * N
    base_n = 16384 * 16384 if weak else n

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    mg.dist_init(rank, world, bcast)
    path = write_cycle(wl["text"](N))
    flags = mg.RUN_FUSED | mg.RUN_QUIET | mg.RUN_NO_FINAL_ERROR | mg.RUN_SKIP_SOURCE
    max_recs = 8192
    recs = (api.TraceRec * max_recs)()
    res = api.CycleResult()
    import ctypes as C
    lo, hi = C.c_int(0), C.c_int(0)

    def one_cycle(u_host=None):
        rc = lib.mgDistRunCycleFile(os.fsencode(path), threshold, flags, u_host, C.byref(lo), C.byref(hi), recs, max_recs, res)
        if rc != 0:
            raise SystemExit("mgDistRunCycleFile failed: %d %s" % (rc, lib.mgLastError().decode()))
        return res.launches

    clk = ClockSampler(local)
    for _ in range(max(args.warmup, 1)):
        one_cycle()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with clk:
        ev0.record(stream)
        for _ in range(args.steps):
            launches += one_cycle()
        ev1.record(stream)
        barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = clk.summary()
    ms_per_step = total_ms / args.steps
    value = (n / base_n) * 1000.0 / ms_per_step          # V-cycles/s in units of 16384^2 fine points
    trace = [dict(node=r.node, N=r.N, steps=r.steps, err=r.err) for r in recs[:res.n_recs]]

    # correctness of the distributed result: final error against the analytic solution
    chk = mg.run_cycle_dist(path, threshold, mg.RUN_FUSED | mg.RUN_QUIET | mg.RUN_SKIP_SOURCE)
    mg_error = chk["mg_error"]

    # ---- end to end: mgDistRunCycleFileHostBatch, one problem per step -- every rank uploads its source slab from pinned
    # memory, the ranks run the V-cycle together, every rank reads its rows of the solution back into pinned memory;
    # consecutive steps are double-buffered per rank (upload i+1 / cycle i / download i-1 on two copy streams).
    e2e = None
    if not args.no_e2e:
        r0, rows, olo, ohi = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        lib.mgDistSourceSlab(N, threshold, C.byref(r0), C.byref(rows), C.byref(olo), C.byref(ohi))
        hF = [torch.empty(max(rows.value, 1) * N, dtype=torch.float64).pin_memory() for _ in range(2)]
        hU = [torch.empty(max(ohi.value - olo.value, 1) * N, dtype=torch.float64).pin_memory() for _ in range(2)]
        for h in hF:
            lib.mgDistDownloadSource(N, h.data_ptr())
        e2e_steps = max(2, min(args.steps, 10))
        e2e_flags = mg.RUN_FUSED | mg.RUN_QUIET | mg.RUN_NO_FINAL_ERROR

        def batch(k):
            Fp, Up, rs = (C.c_void_p * k)(), (C.c_void_p * k)(), (api.CycleResult * k)()
            for i in range(k):
                Fp[i], Up[i] = hF[i % 2].data_ptr(), hU[i % 2].data_ptr()
            rc = lib.mgDistRunCycleFileHostBatch(os.fsencode(path), threshold, e2e_flags, k, Fp, Up, rs)
            if rc != 0:
                raise SystemExit("mgDistRunCycleFileHostBatch failed: %d %s" % (rc, lib.mgLastError().decode()))

        batch(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        batch(e2e_steps)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 1000.0 * wall)) / e2e_steps
        # the host-side ceiling of that call: the same two copies per rank and step (pinned -> device, device -> pinned),
        # all ranks at once, nothing else -- what the PCIe links and the host memory of this box give 'world' GPUs together
        dF = torch.empty_like(hF[0], device="cuda")
        dU = torch.empty_like(hU[0], device="cuda")
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        def copies(k):
            for i in range(k):
                with torch.cuda.stream(s_up):
                    dF.copy_(hF[i % 2], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hU[i % 2].copy_(dU, non_blocking=True)
            s_up.synchronize(); s_dn.synchronize()
        copies(1)
        barrier()
        t0 = time.perf_counter()
        copies(4)
        barrier()
        ceil_ms = max_over_ranks(1000.0 * (time.perf_counter() - t0)) / 4
        del dF, dU
        e2e = {"value": (n / base_n) * 1000.0 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": 8 * rows.value * N,
               "d2h_bytes_per_step": 8 * (ohi.value - olo.value) * N, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "host_ceiling": {"ms_per_step": ceil_ms, "value": (n / base_n) * 1000.0 / ceil_ms,
                                "aggregate_GBs_each_way": world * 8 * rows.value * N / (ceil_ms * 1e6),
                                "what": "the step's two copies alone (no compute), all ranks at once, full duplex: the most this box's PCIe links "
                                        "and host memory allow; e2e / this = how much of the copy time the cycles are hidden behind",
                                "topology": topology(local)},
               "call": "mgDistRunCycleFileHostBatch on every rank (one problem per step: pinned source slab -> device, V-cycle on the "
                       "slabs, owned rows of U -> pinned host; upload of step i+1 and download of step i-1 overlap the cycle of step i); "
                       "bytes are per rank"}
        del hF, hU

    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": (wl["what"] % N) + ": one grid row-slab partitioned over %d GPUs%s" % (world, ", 16384^2 fine points per GPU" if weak else ""),
                   "name": args.config,
                   "value_is": "V-cycles/s x (N_max/16384)^2, i.e. in units of the 1-GPU workload" if weak else "cycles/s of the whole grid",
                   "driver": "mgDistRunCycleFile: fused nodes on row slabs; the %d halo rows go into the neighbours' slabs by peer stores "
                             "of the fused kernel itself (CUDA IPC over NVLink, flag words + stream waits, no communication kernel); levels "
                             "< %d rows: source broadcast to every rank, coarse sub-cycle redundant on every rank" % (8, threshold),
                   "l2": "inputs exceed L2", "parallelism": "row slabs x%d" % world},
        "fine_dof_cycles_per_s": n * 1000.0 / ms_per_step, "global_ms_per_cycle": ms_per_step, "mg_error": mg_error,
        "trace_errors": [t["err"] for t in trace if t["node"] != 0][:48], "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": "whole V-cycle on slabs (fused nodes)", "achieved": 44.0 * n * 4 / 3 / (ms_per_step * 1e6) / world,
                     "peak": hbm_peak, "unit": "GB/s", "frac": 44.0 * n * 4 / 3 / (ms_per_step * 1e6) / world / hbm_peak, "traffic": None,
                     "note": "per GPU: compulsory bytes of the fused cycle (44 B per fine point per level, ladder sum 4/3) / time; "
                             "per-kernel rooflines are reported by the 1-GPU run", "peak_source": peak_src},
    }
    os.unlink(path)
    if rank == 0:
        print(json.dumps(line))
    lib.mgDistShutdown()
    dist.destroy_process_group()
    return 0


def dominant_kernel_roofline(lib, mg, stream, torch, N, hbm_peak, peak_src, unfused):
    """Times the kernels that make up the fine level of the V-cycle alone, with CUDA events on the
    stream they are launched on (inputs are 2 GiB grids, far larger than L2), and reports the one
    that takes the largest share of the cycle (profiles/*launches*.csv: the fused 1 node).
    `achieved` uses the compulsory bytes of SURVEY.md 8(d): every distinct array once."""
    n, M = N * N, N // 2
    m = M * M
    U, W, F, Fc, Uc = mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(N), mg.DeviceGrid(M), mg.DeviceGrid(M)
    lib.getSource(N, 1.0, F.ptr, 0.0, 0.0)
    lib.getSource(M, 1.0, Uc.ptr, 0.0, 0.0)
    lib.mgGridZero(N, U.ptr)
    slot = lib.mgScalarSlot(100)
    reps = 10

    def timeit(fn):
        for _ in range(3):
            fn()
        lib.mgSync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        lib.mgSync()
        return e0.elapsed_time(e1) / reps

    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = {k.split()[0]: v["traffic"] for k, v in json.load(f).items()}

    def entry(name, kernel, ms, nbytes, what):
        return {"name": name, "kernel": kernel, "ms_per_launch": ms, "algorithmic_bytes_per_launch": nbytes,
                "achieved": nbytes / (ms * 1e6), "frac": nbytes / (ms * 1e6) / hbm_peak, "bytes_are": what,
                "traffic": traffic.get(name)}

    kernels = []
    if unfused:
        ms = timeit(lambda: lib.mgSmooth(N, 1.0, U.ptr, F.ptr, 1, W.ptr, None))
        kernels.append(entry("sweep", "k_stream<1,IN_LOAD> (one Jacobi sweep)", ms, 24.0 * n, "U in, F in, U out"))
    else:
        ms = timeit(lambda: lib.mgUpLeg(M, Uc.ptr, N, 1.0, U.ptr, W.ptr, F.ptr, 3, slot))
        kernels.append(entry("up_leg", "k_stream<3,IN_PROLONG,ERR> (1 node: prolong+add+3 sweeps+error)", ms, 24.0 * n + 8.0 * m,
                             "U_f in, F in, U_f out, U_c in (unfused sequence moves 122 B/point)"))
        ms = timeit(lambda: lib.mgDownLeg(N, 1.0, U.ptr, W.ptr, F.ptr, 3, 1, M, Fc.ptr, slot))
        kernels.append(entry("down_leg", "k_stream<3,IN_ZERO,ERR,RES> (-1 node: 3 sweeps+error+residual+negate+restrict)", ms,
                             16.0 * n + 8.0 * m, "F in, U out, F_c out (unfused sequence moves 146 B/point)"))
        ms = timeit(lambda: lib.mgSmooth(N, 1.0, U.ptr, F.ptr, 3, W.ptr, slot))
        kernels.append(entry("smooth_S3", "k_strip<3,IN_LOAD,ERR> (doSmoothing step=3: 3 sweeps+error; 4 columns per lane, bulk copies + mbarriers)", ms, 24.0 * n,
                             "U in, F in, U out (unfused sequence moves 88 B/point)"))
    top = kernels[0]
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved"], "peak": hbm_peak, "unit": "GB/s",
            "frac": top["frac"], "traffic": top["traffic"], "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
            "ms_per_launch": top["ms_per_launch"], "peak_source": peak_src, "grid": "N=%d fine level" % N,
            "traffic_source": "profiles/r02_kernel_traffic.json (ncu --set full, dram__bytes_read+write per launch)",
            "kernels": kernels}


if __name__ == "__main__":
    sys.exit(main())
