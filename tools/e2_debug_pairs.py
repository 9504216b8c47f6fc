"""Why k_epa2 passes pairs of a random pair list on (needs a -DNP_E2_DEBUG build: NETPHYS_B200_LIB=build/lib_e2dbg.so)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
L = npb.load_library()
mix = {"box": (1, 0, 0), "mix": (0.5, 0.5, 0), "sphere": (0, 1, 0)}[sys.argv[1] if len(sys.argv) > 1 else "mix"]
n = 1000000
with npb.World(0) as w:
    specs, sa, pa, sb, pb = scenes.random_pairs(n, seed=99, mix=mix)
    ids = np.array([w.create_sphere(s[1]) if s[0] == "sphere" else w.create_box(*s[1:]) if s[0] == "box" else w.create_mesh(s[1]) for s in specs], np.int32)
    d = [w.dev_alloc(4 * n), w.dev_alloc(24 * n), w.dev_alloc(4 * n), w.dev_alloc(24 * n), w.dev_alloc(4 * n), w.dev_alloc(16 * n)]
    w.dev_upload(d[0], ids[sa]); w.dev_upload(d[1], pa); w.dev_upload(d[2], ids[sb]); w.dev_upload(d[3], pb)
    w.detect_batch_device(n, *d)
    out = (C.c_ulonglong * 8)(); L.np_debug_e2(out); v = list(out)
    print(sys.argv[1:], "expansions %d twice %d no-vertex-room %d >8-new-faces %d >16-faces %d | finished %d passed-on %d" % tuple(v[:7]))
