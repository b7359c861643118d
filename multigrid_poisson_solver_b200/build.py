"""Builds libmgb200.so (CUDA kernels + C ABI + C++ cycle driver) and the MG_GPU
command line for sm_100a, in-tree, with nvcc.  No JIT, no torch extension:

    python -m multigrid_poisson_solver_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmgb200.so")
EXE = os.path.join(HERE, "MG_GPU")
SOURCES = ["mg_abi.cu", "mg_kernels.cu", "mg_fused.cu", "mg_exact.cu", "mg_dist.cu", "mg_tail.cu", "mg_driver.cpp"]
HEADERS = ["mg_context.h", "mg_kernels.h", "mg_fused.h", "mg_device.cuh", "mg_stream.cuh", "mg_stream4.cuh", os.path.join("..", "..", "include", "mg_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # bit parity with the -O0 CPU reference: never contract to FMA
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-Xptxas", "-v",
    "-I/usr/include",         # nccl.h (types only; the library is bound with dlopen at run time)
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    extra = os.environ.get("MG_NVCC_EXTRA", "").split()      # tuning experiments, e.g. -DMG_STREAM_WARPS=4
    if force or extra or _newer(LIB, deps):
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-shared", "-o", LIB] + srcs + ["-ccbin", "g++", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libmgb200.so")
        with open(os.path.join(HERE, "_build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
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
