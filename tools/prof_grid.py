import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
spacing = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sc = scenes.lattice_scene(n, spacing=spacing, seed=43, mix=(0.7, 0.2, 0.1))
for mode in ((1,) if len(sys.argv) > 3 and sys.argv[3] == "gridonly" else (1, 0)):
    with npb.World(0) as w:
        w.set_option(w.OPT_BROADPHASE, mode); w.set_option(w.OPT_PHASE_TIMING, 1); w.set_option(w.OPT_COUNTERS, 1)
        w.load_scene(sc)
        for k in range(4):
            w.step(sc["dt"]); ms = w.last_step_ms()
        w.synchronize(); c = w.counters()
        print("mode", "grid" if mode else "exact", n, "step ms [integrate broad narrow resolve snap total]", np.round(ms, 3), "candidates/step", c["candidates"] / 4, "pair tests/step", c["pair_tests"] / 4, "hits/step", c["hits"] / 4)
