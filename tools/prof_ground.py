"""A carpet of bodies on ONE ground slab (body 0): the slab's candidate list is the whole world (k_cand_large).
    python tools/prof_ground.py [side=512]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = side * side
rng = np.random.RandomState(3)
rows = [scenes.static_data((0, 0, 0), (side * 1.0, -0.45, side * 1.0), (0, 0, 0), 0.0, 0.0)]
sd = np.zeros((n, 32), np.float32)
base = scenes.static_data((0.0, -9.8, 0.0), (0, 0, 0), (0, 0, 0), 1.0, 0.4, 0.4, 0.0, np.float32(1.0) / np.float32(6.0))
sd[:] = base
idx = np.arange(n)
sd[:, 3] = 2.0 * (idx % side) + rng.uniform(-0.3, 0.3, n); sd[:, 4] = 0.5 + rng.uniform(0.0, 0.2, n); sd[:, 5] = 2.0 * (idx // side) + rng.uniform(-0.3, 0.3, n)
sd[:, 6:9] = rng.uniform(-0.2, 0.2, (n, 3))
sc = scenes._scene([("box", 2.2 * side, 2.2 * side, 1.0), ("box", 1.0, 1.0, 1.0)], np.concatenate([[0], np.ones(n, np.int32)]), [rows[0]] + list(sd), np.concatenate([[0], np.ones(n, np.int32)]), "carpet")
with npb.World(0) as w:
    w.set_option(w.OPT_BROADPHASE, w.GRID); w.set_option(w.OPT_PHASE_TIMING, 1); w.set_option(w.OPT_COUNTERS, 1)
    w.load_scene(sc)
    for k in range(4):
        w.step(sc["dt"]); ms = w.last_step_ms()
    w.synchronize(); c = w.counters()
    a, b, d = w.collisions_as_arrays()
    print("carpet of", n, "boxes on one slab: step ms [integrate broad narrow resolve snap total]", np.round(ms, 3), "candidates/step", c["candidates"] / 4, "pair tests/step", c["pair_tests"] / 4, "collisions", len(a), "slab's hit", b[a == 0])
