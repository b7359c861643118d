#!/usr/bin/env python
"""Device time of whole V-cycles on small ladders (what the agglomerated part of a slab run costs)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_poisson_solver_b200 as mg  # noqa: E402

mg.init(0)
for N in [int(a) for a in sys.argv[1:]] or [2048, 1448, 1024, 724, 512, 362, 256, 181, 128, 90, 64]:
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    f.write(mg.cycles.v_cycle(N, 8))
    f.close()
    best = 1e9
    for _ in range(6):
        r = mg.run_cycle(f.name)
        best = min(best, r["time_ms"])
    print("V-cycle %5d -> 8: %.3f ms, %d launches" % (N, best, r["launches"]))
    os.unlink(f.name)
