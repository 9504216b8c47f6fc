"""Tiny workload for compute-sanitizer: every kernel family once (world step exact + grid, pair list 3-D and 2-D, overflow tiers)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
sc = scenes.lattice_scene(600, spacing=1.0, seed=3, mix=(0.4, 0.3, 0.3))
for mode in (0, 1):
    with npb.World(0) as w:
        w.set_option(w.OPT_BROADPHASE, mode); w.set_option(w.OPT_COUNTERS, 1); w.set_option(w.OPT_SNAPSHOT_EACH_STEP, 1)
        w.load_scene(sc)
        for _ in range(2): w.step(sc["dt"])
        w.synchronize(); print("mode", mode, w.counters()["hits"], len(w.collisions()))
specs, sa, pa, sb, pb = scenes.random_pairs(6000, seed=4242, extent=0.6, mix=(0.2, 0.4, 0.4))
for g, e in ((1, 1), (8, 4), (32, 32)):
    os.environ["NP_G_GJK"] = str(g); os.environ["NP_G_EPA"] = str(e)
    with npb.World(0) as w:
        w.set_option(w.OPT_COUNTERS, 1)
        ids = np.array([w.create_box(1, 1, 1), w.create_sphere(0.5), w.create_mesh(specs[2][1])], np.int32)
        hit, out = w.detect_batch(ids[sa], pa, ids[sb], pb)
        h2, o2 = w.detect_batch(ids[sa[:500]], pa[:500], ids[sb[:500]], pb[:500], is3d=False)
        print("G", g, e, int(hit.sum()), w.counters()["epa_overflow"])
