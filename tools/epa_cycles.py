import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, "/root/repo")
os.environ["NETPHYS_B200_LIB"] = "/root/repo/netphys_b200/libdbg.so"
import netphys_b200 as npb
from netphys_b200 import scenes
n = 10000
sc = scenes.lattice_scene(n, spacing=1.2, seed=42)
with npb.World(0) as w:
    w.load_scene(sc)
    for k in range(5): w.step(sc["dt"])
    out = np.zeros((n, 8), np.uint64)
    w.L.np_debug_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    w.L.np_debug_dump(w.w, out.ctypes.data, n)
r = out.astype(np.int64); it = r[:, 6]
for lo, hi in ((1, 2), (2, 6), (6, 12), (12, 21)):
    m = (it >= lo) & (it < hi)
    if not m.any(): continue
    c = r[m][:, 0:4].astype(float); tot = c.sum(1); per_it = tot / np.maximum(it[m], 1)
    print("iters [%2d,%2d): %5d pairs | cycles/iter %7.0f | produce %4.1f%% supports %4.1f%% consume(all) %4.1f%% of which survivors+faces %4.1f%% | final faces mean %.0f" % (lo, hi, m.sum(), per_it.mean(), 100 * c[:, 0].sum() / tot.sum(), 100 * c[:, 1].sum() / tot.sum(), 100 * c[:, 2].sum() / tot.sum(), 100 * c[:, 3].sum() / tot.sum(), r[m][:, 7].mean()))
