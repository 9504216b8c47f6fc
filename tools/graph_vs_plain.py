"""C2 world: wall-clock per step of back-to-back np_world_step calls, graph replay (default) vs plain launches (NP_GRAPH=0)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sc = scenes.lattice_scene(n, spacing=1.2, seed=42)
with npb.World(0) as w:
    w.set_option(w.OPT_SNAPSHOT_EACH_STEP, 1)
    w.load_scene(sc)
    for _ in range(10): w.step(sc["dt"])
    w.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(200): w.step(sc["dt"])
        w.synchronize()
        t = (time.perf_counter() - t0) / 200
        print("NP_GRAPH=%s  %d bodies: %.4f ms per step (wall, 200 back-to-back steps); last step by its own events %.4f ms" % (os.environ.get("NP_GRAPH", "1"), n, 1e3 * t, w.last_step_ms()[5]))
