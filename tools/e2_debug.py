"""Why k_epa2 passes pairs on (needs a -DNP_E2_DEBUG build: NETPHYS_B200_LIB=build/lib_e2dbg.so)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
L = npb.load_library()
def dump(tag):
    out = (C.c_ulonglong * 8)(); L.np_debug_e2(out)
    v = list(out)
    print(tag, "expansions %d twice %d no-vertex-room %d >8-new-faces %d >16-faces %d | finished %d passed-on %d" % tuple(v[:7]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
sc = scenes.lattice_scene(n, spacing=2.0, seed=43, mix=(0.7, 0.2, 0.1))
with npb.World(0) as w:
    w.load_scene(sc)
    for _ in range(3): w.step(sc["dt"])
    w.synchronize() if hasattr(w, "synchronize") else w.states()
    dump("world %d x 3 steps:" % n)
