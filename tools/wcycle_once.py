import sys, os, tempfile
sys.path.insert(0, '/root/repo')
import multigrid_poisson_solver_b200 as mg
mg.init(0)
f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False); f.write(mg.cycles.w_cycle(16384, 16, step=3, tol=1e-8)); f.close()
mg.run_cycle(f.name)
r = mg.run_cycle(f.name)
print("W-cycle device ms", r["time_ms"], "launches", r["launches"])
