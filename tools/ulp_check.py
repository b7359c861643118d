import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multigrid_poisson_solver_b200 as mg
from oracle import pyoracle as po
mg.init(0); g = mg.GpuOps(); o = po.oracle_ops()
for N in (8, 17, 256, 1000):
    for args in ((N,), (N, 2.0, -0.5, 0.25)):
        for name in ("getSource", "getAnalytic"):
            a = getattr(g, name)(*args); b = getattr(o, name)(*args)
            ulp = np.spacing(np.maximum(np.abs(b), 1e-300))
            d = np.abs(a-b)/ulp
            k = int(np.argmax(d))
            print(name, args, "max ulp", d.max(), "at", k, a[k], b[k], "n>1ulp", int((d>1).sum()))
