#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2g_tests.log
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
python - <<'P'
import sys, time, numpy as np
sys.path.insert(0, '.')
import netphys_b200 as npb
from netphys_b200 import scenes
for n in (10000, 1000000):
    sc = scenes.lattice_scene(n, spacing=2.0 if n > 20000 else 2.5, seed=43, mix=(0.7, 0.2, 0.1) if n > 20000 else (0.5, 0.5, 0))
    for fuse in ("1", "0"):
        import os; os.environ["NP_FUSE"] = fuse
        with npb.World(0) as w:
            w.load_scene(sc); w.step_n(sc["dt"], 4); w.synchronize()
            t0 = time.perf_counter(); w.step_n(sc["dt"], 20); w.synchronize(); t = time.perf_counter() - t0
            print("step_n n=%d fuse=%s: %.4f ms/step" % (n, fuse, 1e3 * t / 20))
P
