import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sc = scenes.lattice_scene(n, spacing=1.0, seed=43, mix=(0.7, 0.2, 0.1))
with npb.World(0) as w:
    w.set_option(w.OPT_BROADPHASE, 1)
    w.load_scene(sc)
    for k in range(3):
        w.step(sc["dt"]); ms = w.last_step_ms()
    print("grid", n, np.round(ms, 3))
