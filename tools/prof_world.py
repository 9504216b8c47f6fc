"""Small fixed workload for ncu: a few np_world_step calls on a lattice world."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mix = {"c2": (0.5, 0.5, 0.0), "c3": (0.7, 0.2, 0.1), "box": (1, 0, 0)}[sys.argv[3] if len(sys.argv) > 3 else "c2"]
sc = scenes.lattice_scene(n, spacing=1.2, seed=42, mix=mix)
with npb.World(0) as w:
    w.set_option(w.OPT_PHASE_TIMING, 1)
    w.load_scene(sc)
    for k in range(steps):
        w.step(sc["dt"])
        ms = w.last_step_ms()
    print("world", n, "step ms [integrate broad narrow resolve snap total]", np.round(ms, 4), "body-steps/s %.3e" % (n / (ms[5] * 1e-3)))
