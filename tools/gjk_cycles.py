"""Per-phase cycles of the GJK trips (needs netphys_b200/libdbg.so built with -DNP_TIMELINE -DNP_TL_GJK).
slots per body: 0 start ns, 1 end ns, 2 trips, 3 cycles in supports, 4 in consume, 5 in produce (solve + direction)."""
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
    out = np.zeros((n, 8), np.uint64)
    w.L.np_debug_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    w.L.np_debug_dump(w.w, out.ctypes.data, n)
r = out.astype(np.int64)
dur_ns = r[:, 1] - r[:, 0]; trips = r[:, 2]; sup = r[:, 3]; con = r[:, 4]; pro = r[:, 5]
for lo, hi in ((1, 8), (8, 20), (20, 40), (40, 200)):
    m = (trips >= lo) & (trips < hi) & (dur_ns > 0)
    if not m.any(): continue
    cyc_total = dur_ns[m] * 1.965     # ns -> cycles at 1965 MHz
    print("trips [%3d,%3d): %5d bodies | %.0f cycles per trip (wall) | per trip: supports %.0f  consume %.0f  produce %.0f cycles | of the wall time %.0f%% %.0f%% %.0f%%"
          % (lo, hi, m.sum(), (cyc_total / trips[m]).mean(), (sup[m] / trips[m]).mean(), (con[m] / trips[m]).mean(), (pro[m] / trips[m]).mean(),
             100 * sup[m].sum() / cyc_total.sum(), 100 * con[m].sum() / cyc_total.sum(), 100 * pro[m].sum() / cyc_total.sum()))
