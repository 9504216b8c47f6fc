#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/sharded_run.py [bodies] [steps] [grid]
One world, pair search sharded over N GPUs (netphys_b200.dist.ShardedWorld, NCCL all-gather of the collision records).
Checks on every rank that the replicated state is bit-identical to an unsharded world stepped on the same GPU, and prints timing."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb  # noqa: E402
from netphys_b200 import scenes  # noqa: E402
from netphys_b200.dist import ShardedWorld  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
grid = len(sys.argv) > 3 and sys.argv[3] == "grid"
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); ws = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t0 = time.time()
sc = scenes.lattice_scene(n, spacing=1.3, seed=43, mix=(0.7, 0.2, 0.1))
t_scene = time.time() - t0
W = npb.World
with W(lr) as w, W(lr) as ref:
    for x in (w, ref):
        if grid: x.set_option(W.OPT_BROADPHASE, W.GRID)
        x.set_option(W.OPT_PHASE_TIMING, 1)
        x.load_scene(sc)
    sw = ShardedWorld(w, torch.device("cuda", lr), ws, rank)
    dt = sc["dt"]
    for _ in range(3):
        sw.step(dt); ref.step(dt)
    w.synchronize(); ref.synchronize()
    same = np.array_equal(w.states().view(np.uint32), ref.states().view(np.uint32))
    dist.barrier(); torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        w.flush_l2(); sw.step(dt); ms.append(w.last_step_ms())
    ms = np.array(ms); sh = float(np.mean(ms[:, 5]))
    rs = []
    for _ in range(steps):
        ref.flush_l2(); ref.step(dt); rs.append(ref.last_step_ms())
    rs = np.array(rs); un = float(np.mean(rs[:, 5]))
    same &= np.array_equal(w.states().view(np.uint32), ref.states().view(np.uint32))
    a, b, d = w.collisions_as_arrays(); a2, b2, d2 = ref.collisions_as_arrays()
    same &= np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))
    t = torch.tensor([sh, un, 0.0 if same else 1.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sh, un, bad = [float(v) for v in t]
        print("bodies %d ranks %d grid %d | scene %.1fs | unsharded %.3f ms/step | sharded %.3f ms/step (x%.2f) | phases(sharded) %s | collisions %d | bit-identical on all ranks: %s"
              % (n, ws, grid, t_scene, un, sh, un / sh, np.round(ms[:, :5].mean(0), 3).tolist(), len(a), bad == 0.0))
dist.destroy_process_group()
