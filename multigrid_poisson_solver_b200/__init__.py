"""B200-native geometric multigrid Poisson hot path.

The product is the C-ABI shared library `libmgb200.so` (hand-written sm_100a CUDA kernels,
C++ cycle driver; see include/mg_abi.h) plus the `MG_GPU` command line.  This package is the
thin Python host mirror of that ABI used by the tests and bench.py: same operator names and
argument meaning as the reference (MG_solver_CPU.cpp:16-34), grids resident on the device.

There is no CPU fallback: importing works anywhere, but every operator needs the built
library and a CUDA device and raises loudly otherwise.
"""
from .api import (run_cycle_dist_emulated, run_cycle_dist, dist_init, dist_plan,  # noqa: F401
                  MGLibraryError, DeviceGrid, GpuOps, init, lib, lib_path, run_cycle, run_cycle_host, run_cycle_host_batch, run_cycle_problem,  # noqa: F401
                  RUN_FUSED, RUN_UNFUSED, RUN_QUIET, RUN_SKIP_SOURCE, RUN_NO_FINAL_ERROR)
from . import cycles  # noqa: F401
