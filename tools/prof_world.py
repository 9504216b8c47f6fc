"""Small fixed workload for ncu: a few np_world_step calls on a lattice world.

    python tools/prof_world.py [bodies] [steps] [c2|c3|box] [spacing] [grid]
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kind = sys.argv[3] if len(sys.argv) > 3 else "c2"
mix = {"c2": (0.5, 0.5, 0.0), "c3": (0.7, 0.2, 0.1), "box": (1, 0, 0)}[kind]
spacing = float(sys.argv[4]) if len(sys.argv) > 4 else (2.0 if kind == "c3" else 2.5)
sc = scenes.lattice_scene(n, spacing=spacing, seed=43 if kind == "c3" else 42, mix=mix)
with npb.World(0) as w:
    w.set_option(w.OPT_PHASE_TIMING, 1)
    w.set_option(w.OPT_COUNTERS, 1)
    if len(sys.argv) > 5 and sys.argv[5] == "grid":
        w.set_option(w.OPT_BROADPHASE, w.GRID)
    w.load_scene(sc)
    acc = []
    for k in range(steps):
        w.step(sc["dt"])
        ms = w.last_step_ms()
        if k >= 2: acc.append(ms)
    ms = np.mean(acc, axis=0) if acc else ms
    st = w.states().view(np.uint32)
    c = w.counters()
    print("world", n, kind, "step ms [integrate broad narrow resolve snap total]", np.round(ms, 4), "body-steps/s %.3e" % (n / (ms[5] * 1e-3)),
          "state-hash %08x %d" % (int(np.bitwise_xor.reduce(st.ravel())), int(st.sum(dtype=np.uint64) & 0x7fffffff)),
          "pair_tests/step %.0f hits/step %.0f" % (c["pair_tests"] / steps, c["hits"] / steps))
