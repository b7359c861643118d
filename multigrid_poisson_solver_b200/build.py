"""Builds libmgb200.so (CUDA kernels + C ABI + C++ cycle driver) and the MG_GPU
command line for sm_100a, in-tree, with nvcc.  No JIT, no torch extension:

    python -m multigrid_poisson_solver_b200.build [--force] [--verbose]

Every source is compiled to its own object (csrc/_obj/, git-ignored) in parallel; an object is
rebuilt when its source, ANY header under csrc/ or include/, this script or the flags changed.
"""
import concurrent.futures
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, os.environ.get("MG_LIB_NAME", "libmgb200.so"))   # MG_LIB_NAME: tuning variants next to the product library
EXE = os.path.join(HERE, "MG_GPU")
SOURCES = ["mg_abi.cu", "mg_kernels.cu", "mg_fused.cu", "mg_legs.cu", "mg_peer.cu", "mg_exact.cu", "mg_dist.cu", "mg_tail.cu", "mg_driver.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # bit parity with the -O0 CPU reference: never contract to FMA
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-Xptxas", "-v",
    "-I/usr/include",         # nccl.h (types only; the library is bound with dlopen at run time)
]


def headers():
    """Every header a source may include: all of csrc/*.h, csrc/*.cuh and include/*.h."""
    inc = os.path.join(HERE, "..", "include")
    return sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(inc, "*.h")))


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    extra = os.environ.get("MG_NVCC_EXTRA", "").split()      # tuning experiments, e.g. -DMG_STREAM_DEPTH=8
    flags = NVCC_FLAGS + extra
    tag = hashlib.sha1(" ".join(flags).encode()).hexdigest()[:10]
    os.makedirs(OBJ, exist_ok=True)
    common = headers() + [os.path.abspath(__file__)]
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, "%s.%s.o" % (s, tag))
        if not (force or _newer(obj, [src] + common)):
            return obj, "", False
        cmd = [nvcc] + flags + ["-c", "-o", obj, src, "-ccbin", "g++"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed compiling %s" % s)
        return obj, " ".join(cmd) + "\n" + r.stdout + r.stderr, True

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _, _ in results]
    rebuilt = any(b for _, _, b in results)
    if rebuilt or _newer(LIB, objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ccbin", "g++", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libmgb200.so")
        # keep the ptxas -v output of the objects that were rebuilt now, the old log for the others
        logs = {}
        log_path = os.path.join(HERE, "_build.log")
        if os.path.exists(log_path):
            cur = None
            for line in open(log_path):
                if line.startswith("### "):
                    cur = line[4:].strip()
                    logs[cur] = ""
                elif cur:
                    logs[cur] += line
        for s, (_, log, b) in zip(srcs, results):
            if b:
                logs[s] = log
        with open(log_path, "w") as f:
            for s in srcs:
                f.write("### %s\n%s" % (s, logs.get(s, "")))
        if verbose:
            sys.stderr.write(open(log_path).read())
    main_src = os.path.join(CSRC, "mg_main.cpp")
    if force or _newer(EXE, [main_src, LIB]):
        cmd = [nvcc, "-O2", "-std=c++17", "-o", EXE, main_src, "-ccbin", "g++",
               "-L" + HERE, "-lmgb200", "-Xlinker", "-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building MG_GPU")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
