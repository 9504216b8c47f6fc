"""Small fixed workload for ncu: one k_pairs launch over random pairs of a given shape mix."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes

mix = {"box": (1, 0, 0), "mix": (0.5, 0.5, 0), "sphere": (0, 1, 0), "all": (0.4, 0.4, 0.2)}[sys.argv[1] if len(sys.argv) > 1 else "box"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with npb.World(0) as w:
    w.set_option(npb.World.OPT_COUNTERS, int(os.environ.get('NP_COUNT', '0')))
    specs, sa, pa, sb, pb = scenes.random_pairs(n, seed=99, mix=mix)
    ids = np.array([w.create_sphere(s[1]) if s[0] == "sphere" else w.create_box(*s[1:]) if s[0] == "box" else w.create_mesh(s[1]) for s in specs], np.int32)
    d = [w.dev_alloc(4 * n), w.dev_alloc(24 * n), w.dev_alloc(4 * n), w.dev_alloc(24 * n), w.dev_alloc(4 * n), w.dev_alloc(16 * n)]
    w.dev_upload(d[0], ids[sa]); w.dev_upload(d[1], pa); w.dev_upload(d[2], ids[sb]); w.dev_upload(d[3], pb)
    ms = [w.detect_batch_device(n, *d) for _ in range(reps)]
    print("pairs", n, "ms", [round(m, 3) for m in ms], "pairs/s %.3e" % (n / (min(ms) * 1e-3)), w.counters() if os.environ.get("NP_COUNT") else "")
