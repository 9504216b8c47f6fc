"""Per-body timeline of the narrow phase (needs netphys_b200/libdbg.so built with -DNP_TIMELINE).
slots per body: 0 gjk start, 1 gjk end, 2 gjk trips, 3 widened at, 4 epa start, 5 epa end, 6 epa iterations, 7 (na<<16)|nb of the hit pair"""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NETPHYS_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "netphys_b200", "libdbg.so")
import netphys_b200 as npb
from netphys_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sc = scenes.lattice_scene(n, spacing=1.2, seed=42)
with npb.World(0) as w:
    w.set_option(w.OPT_COUNTERS, 1)
    w.set_option(w.OPT_PHASE_TIMING, 1)
    w.load_scene(sc)
    for k in range(5): w.step(sc["dt"])
    print("step ms", np.round(w.last_step_ms(), 4))
    out = np.zeros((n, 8), np.uint64)
    w.L.np_debug_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    w.L.np_debug_dump(w.w, out.ctypes.data, n)
raw = out.astype(np.int64)
base = raw[:, 0][raw[:, 0] > 0].min()
T = lambda k: np.where(raw[:, k] > 0, (raw[:, k] - base) / 1e3, np.nan)
def pct(x, name):
    x = x[~np.isnan(x)]
    if len(x): print("%-10s n=%5d  p50 %.1f  p90 %.1f  p99 %.1f  p99.9 %.1f  max %.1f" % (name, len(x), *[np.percentile(x, q) for q in (50, 90, 99, 99.9)], x.max()))
g0, g1, gw, e0, e1 = T(0), T(1), T(3), T(4), T(5)
trips = raw[:, 2].astype(float); it = raw[:, 6].astype(float)
pct(g0, "gjk start"); pct(g1, "gjk end"); pct(g1 - g0, "gjk dur"); pct(trips, "gjk trips"); pct((g1 - g0) / np.maximum(trips, 1), "us/trip")
pct(e0, "epa start"); pct(e1, "epa end"); pct(e1 - e0, "epa dur"); pct(it, "epa iters"); pct((e1 - e0) / np.maximum(it, 1), "us/iter")
print("last GJK finishers: body start end trips us/trip widened_at verts(a,b)")
for k in np.argsort(np.nan_to_num(g1))[-8:]: print("   %5d %7.1f %7.1f %4d %6.2f %7.1f   %d,%d" % (k, g0[k], g1[k], trips[k], (g1[k] - g0[k]) / max(trips[k], 1), gw[k], raw[k, 7] >> 16, raw[k, 7] & 0xffff))
print("last EPA finishers: body gjk_end epa_start epa_end iters us/iter verts(a,b)")
for k in np.argsort(np.nan_to_num(e1))[-8:]: print("   %5d %7.1f %7.1f %7.1f %4d %6.2f   %d,%d" % (k, g1[k], e0[k], e1[k], it[k], (e1[k] - e0[k]) / max(it[k], 1), raw[k, 7] >> 16, raw[k, 7] & 0xffff))
