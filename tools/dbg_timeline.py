import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NETPHYS_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "netphys_b200", "libdbg.so")
import netphys_b200 as npb
from netphys_b200 import scenes
n = 10000
sc = scenes.lattice_scene(n, spacing=1.2, seed=42)
with npb.World(0) as w:
    w.load_scene(sc)
    for k in range(5): w.step(sc["dt"])
    print(w.last_step_ms())
    out = np.zeros((n, 4), np.uint64)
    w.L.np_debug_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    w.L.np_debug_dump(w.w, out.ctypes.data, n)
t0 = out[:, 0].astype(np.int64); t1 = out[:, 1].astype(np.int64); sm = out[:, 2]; trips = out[:, 3].astype(np.int64)
ok = t1 > 0
base = t0[ok].min()
print("bodies with tests", ok.sum(), "makespan us", (t1[ok].max() - base) / 1e3)
dur = (t1 - t0)[ok] / 1e3
print("duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur.mean(), np.percentile(dur, 50), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
print("trips: mean %.1f p99 %d max %d" % (trips[ok].mean(), np.percentile(trips[ok], 99), trips[ok].max()))
end = (t1[ok] - base) / 1e3; start = (t0[ok] - base) / 1e3
for q in (50, 90, 99, 99.9, 100): print("  %5.1f%% of bodies finished by %.1f us ; started by %.1f us" % (q, np.percentile(end, q), np.percentile(start, q)))
idx = np.argsort(end)[-10:]
print("last finishers: (start, end, dur, trips)")
for k in idx: print("   %.1f %.1f %.1f %d" % (start[k], end[k], dur[k], trips[ok][k]))
print("us per trip (all):", (dur / np.maximum(trips[ok], 1)).mean())
