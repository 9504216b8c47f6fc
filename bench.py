#!/usr/bin/env python
"""bench.py -- headline benchmark of the netphys rigid-body hot path on B200 (driver contract: one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one Physics_Update (integrate -> ordered first-hit GJK/EPA pair search -> impulse response -> replication
snapshot) of BASELINE.json configs[1]: 10 000 boxes/spheres in one world.  With N > 1 every rank steps its own such
world (independent worlds, weak scaling) and the per-step snapshots are all-gathered over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from netphys_b200 import scenes  # noqa: E402

BODIES = 10000
SPACING, JITTER = 1.2, 0.3
WORKLOAD = ("C2 (BASELINE configs[1]): 10000 bodies in one world, unit boxes (8 vertices) and r=0.5 icospheres (162 vertices) alternating, "
            "jittered cubic lattice spacing 1.2 +-0.3, rotations U(-1,1)^3, gravity -9.8, dt 1/30; pair search = reference semantics "
            "(all pairs i<j, stop at the first hit per body)")
# algorithmic flop model of the narrow phase (reference arithmetic, + - * / sqrt cmp = 1 each)
F_SUPPORT_VERT, F_GJK_STEP, F_EPA_TRI, F_EPA_ITER = 27, 120, 110, 60
# algorithmic bytes per body of the integrate+transform pass
B_INTEGRATE = 48 + 56 + 36 + 48
# DRAM traffic of the narrow-phase kernels in one C2 step, from the committed ncu capture (not measured live)
NARROW_DRAM_BYTES_PER_STEP = 13.76e6 + 12.05e6


def make_scene(rank=0):
    return scenes.lattice_scene(BODIES, spacing=SPACING, jitter=JITTER, seed=42 + rank)


def narrow_flops(c0, c1):
    d = {k: c1[k] - c0[k] for k in c1}
    return F_SUPPORT_VERT * d["support_verts"] + F_GJK_STEP * d["gjk_steps"] + F_EPA_TRI * d["epa_tri_tests"] + F_EPA_ITER * d["epa_iters"], d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []; self.p = None; self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append(f)

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15); self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_steps(scene, steps, warmup):
    """The reference's own CPU implementation of the step (oracle/_ref when it was built, else the C port)."""
    from oracle.bind import Oracle, Ref, ref_available
    o = Oracle()
    sc = scenes.resolve(scene, o)
    if ref_available():
        r = Ref(); r.load_scene(sc)
        for _ in range(warmup): r.step(sc["dt"])
        t = r.time_steps(sc["dt"], steps)
        return t, "reference", "Physics_Update of the unmodified reference engine (oracle/_ref), 1 thread: the engine is single-threaded with a process-global body list"
    w = o.world(sc)
    for _ in range(warmup): w.step(sc["dt"])
    t0 = time.perf_counter()
    for _ in range(steps): w.step(sc["dt"])
    return time.perf_counter() - t0, "port", "C restatement of Physics_Update (oracle/netphys_oracle.c), 1 thread"


def run_reference(args, rank):
    if rank != 0:
        return
    scene = make_scene(0)
    t, kind, how = cpu_reference_steps(scene, args.steps, min(args.warmup, 2))
    v = BODIES * args.steps / t
    sample = "%d full steps of the %d-body world (%s)" % (args.steps, BODIES, how)
    print(json.dumps({"impl": "reference", "metric": "body_steps_per_sec", "value": v, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "bodies_per_gpu": BODIES},
                      "cpu_baseline": {"value": v, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_sharded_world(npb, torch, dist, rank, local_rank, world_size):
    """One world of world_size x 1M mixed bodies (C3/C5 generator) replicated on every rank; each rank runs the first-hit search of its
    index range, the collision records are all-gathered (24 B/body, the step's only exchange), every rank applies all responses."""
    from netphys_b200.dist import ShardedWorld
    per_gpu = int(os.environ.get("NP_BENCH_SHARDED_BODIES_PER_GPU", "1000000"))
    n = per_gpu * world_size
    sc = scenes.lattice_scene(n, spacing=1.3, seed=43, mix=(0.7, 0.2, 0.1))
    W = npb.World
    with W(local_rank) as w:
        w.set_option(W.OPT_PHASE_TIMING, 1)
        w.load_scene(sc)
        sw = ShardedWorld(w, torch.device("cuda", local_rank), world_size, rank)
        for _ in range(3): sw.step(sc["dt"])
        w.synchronize(); dist.barrier()
        ms = []
        for _ in range(5):
            w.flush_l2(); sw.step(sc["dt"]); ms.append(w.last_step_ms())
        ms = np.array(ms, np.float64).mean(0)
        st = w.states().view(np.uint32)
        h = np.array([int(np.bitwise_xor.reduce(st.ravel())), int(st.sum(dtype=np.uint64) & 0x7fffffff)], np.int64)
        a, b, _ = w.collisions_as_arrays()
    t = torch.tensor(ms, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.cpu().numpy()
    hs = torch.tensor(h, device="cuda"); lo = hs.clone(); hi = hs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"workload": "ONE world of %d bodies (70%% box, 20%% icosphere, 10%% 16-vertex hull; the configs[2]/[4] generator) replicated on %d GPUs; "
                        "exact first-hit search sharded by body index range, NCCL all-gather of the 24 B/body collision records, responses applied on every rank" % (n, world_size),
            "value": n / (float(ms[5]) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(ms[5]), "scaling": "weak (1M searched bodies per GPU; integrate/world-vertex/response passes are replicated)",
            "phase_ms": {"integrate": float(ms[0]), "world_verts": float(ms[1]), "search_owned": float(ms[2]), "allgather_and_respond": float(ms[3]), "snapshot": float(ms[4])},
            "collisions_per_step": int(len(a)), "state_bit_identical_on_all_ranks": bool(torch.equal(lo, hi)), "l2": "flushed before every timed step", "timing": "CUDA events on the world's stream, mean of 5 steps, max over ranks"}


def run_worlds_batch(npb, torch, dist, rank, local_rank, world_size):
    """BASELINE configs[3]: 4096 independent 256-body worlds sharded over the ranks (contiguous blocks of worlds, islands of one store
    per rank), no data-path collective; the per-step snapshot (36 B/body) is all-gathered for replication."""
    from netphys_b200.dist import SnapshotExchange, shard_range
    lo, hi = shard_range(4096, rank, world_size)
    W = npb.World
    base = scenes.lattice_scene(256, spacing=1.2, seed=1000)
    with W(local_rank) as w:
        w.set_option(W.OPT_SNAPSHOT_EACH_STEP, 1)
        ids = np.array([w.create_box(1.0, 1.0, 1.0), w.create_sphere(0.5)], np.int32)
        for k in range(lo, hi):
            if k > lo: w.begin_island()
            sck = scenes.lattice_scene(256, spacing=1.2, seed=1000 + k) if (k - lo) < 64 else base       # 64 distinct worlds per rank, then repeats (setup time)
            w.create_bodies(ids[sck["body_shape"]], npb.lib.sd_rows_to_struct(sck["sd"], sck["response"]))
        n = w.n
        xchg = SnapshotExchange(36 * 256 * -(-4096 // world_size), torch.device("cuda", local_rank), world_size, rank)
        ext = torch.cuda.ExternalStream(w.L.np_world_stream(w.w), device=torch.device("cuda", local_rank))
        def step():
            buf = xchg.next_send_buffer(); w.set_snapshot_target(buf.data_ptr(), lo * 256)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            w.flush_l2(); e0.record(ext); w.step(base["dt"])
            with torch.cuda.stream(ext): xchg.exchange(buf)
            e1.record(ext); e1.synchronize()
            return e0.elapsed_time(e1), w.last_step_ms()[5]
        for _ in range(3): step()
        dist.barrier()
        r = np.array([step() for _ in range(5)], np.float64).mean(0)
    t = torch.tensor(r, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); r = t.cpu().numpy()
    return {"workload": "BASELINE configs[3]: 4096 independent 256-body worlds (box/icosphere) sharded over %d GPUs, %d worlds per GPU as islands of one store; "
                        "snapshot all-gather (36 B/body) inside the timed region, not overlapped" % (world_size, hi - lo),
            "value": 4096 * 256 / (float(r[0]) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(r[0]), "ms_step_kernels_only": float(r[1]),
            "scaling": "strong (4096 worlds in total)", "timing": "CUDA events on the world's stream around step + all-gather, mean of 5, max over ranks; L2 flushed"}


def run_ours(args, rank, local_rank, world_size):
    import netphys_b200 as npb
    if npb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (netphys_b200 has no CPU path)")
    dist = None; torch = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = make_scene(rank)
    W = npb.World
    w = W(local_rank)
    w.set_option(W.OPT_COUNTERS, 1)
    w.set_option(W.OPT_SNAPSHOT_EACH_STEP, 1)
    w.load_scene(scene)
    dt = scene["dt"]
    xchg = None
    if world_size > 1:
        from netphys_b200.dist import SnapshotExchange
        xchg = SnapshotExchange(36 * BODIES, torch.device("cuda", local_rank), world_size, rank)

    ext = comm = None
    if xchg is not None:
        ext = torch.cuda.ExternalStream(w.L.np_world_stream(w.w), device=torch.device("cuda", local_rank))   # the world's own stream, seen by torch
        comm = torch.cuda.Stream(device=torch.device("cuda", local_rank))

    def one_step(timed):
        ag_ms = 0.0
        w.flush_l2()
        if xchg is not None:
            buf = xchg.next_send_buffer(); w.set_snapshot_target(buf.data_ptr(), rank * BODIES)
        w.step(dt)                      # asynchronous: integrate .. snapshot enqueued on the world's stream
        if xchg is not None:            # the all-gather is enqueued behind the step (no host sync in between) and timed on its own stream
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(comm):
                comm.wait_stream(ext)
                e0.record(comm); xchg.exchange(buf); e1.record(comm)
        ms = w.last_step_ms()           # synchronises the world's stream
        if xchg is not None:
            e1.synchronize(); ag_ms = e0.elapsed_time(e1)
        return ms, ag_ms

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    w.synchronize()
    clocks = ClockSampler(local_rank); clocks.start()
    c0 = w.counters()
    t_wall0 = time.perf_counter()
    per = np.zeros((args.steps, 7), np.float64)
    if xchg is None:
        for k in range(args.steps):
            ms, ag = one_step(True)
            per[k, :6] = ms; per[k, 6] = ag
        w.synchronize()
        step_ms = float(np.mean(per[:, 5]))
    else:
        # N > 1: the all-gather of step k runs on its own stream while step k+1 computes (double-buffered snapshot), as a server
        # would run it.  No host sync inside the loop.  A step is charged max(its kernels, the previous all-gather still in flight),
        # both measured from the step's first kernel with CUDA events; the last all-gather's exposed tail is added once.
        ev = lambda: torch.cuda.Event(enable_timing=True)
        A, B, Cc = [], [], []
        for k in range(args.steps):
            if k >= 2: ext.wait_event(Cc[k - 2])                      # the send buffer about to be rewritten has been gathered
            w.flush_l2()
            buf = xchg.next_send_buffer(); w.set_snapshot_target(buf.data_ptr(), rank * BODIES)
            a = ev(); a.record(ext); w.step(dt); bb = ev(); bb.record(ext)
            with torch.cuda.stream(comm):
                comm.wait_event(bb); xchg.exchange(buf); c = ev(); c.record(comm)
            A.append(a); B.append(bb); Cc.append(c)
        torch.cuda.synchronize(); w.synchronize()
        tot = 0.0
        for k in range(args.steps):
            s = A[k].elapsed_time(B[k]); per[k, 5] = s
            if k >= 1:
                lag = A[k].elapsed_time(Cc[k - 1]); per[k, 6] = max(0.0, lag - s); s = max(s, lag)
            tot += s
        tail = max(0.0, B[-1].elapsed_time(Cc[-1])); per[-1, 6] += tail
        step_ms = (tot + tail) / args.steps
    if dist is not None:
        torch.cuda.synchronize(); dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    c1 = w.counters()
    clk = clocks.stop()
    # Second pass with per-phase events (NP_OPT_PHASE_TIMING): the events cost ~4 us each inside a 0.3 ms step, so the headline pass
    # above runs without them; the phase breakdown and the narrow-phase kernel time of the roofline come from this pass.
    w.set_option(W.OPT_PHASE_TIMING, 1)
    for _ in range(2): one_step(False)
    p0 = w.counters()
    ph = np.array([one_step(True)[0] for _ in range(args.steps if xchg is None else 5)], np.float64)
    p1 = w.counters()
    w.set_option(W.OPT_PHASE_TIMING, 0)
    per[:, :5] = ph[:, :5].mean(0); phase_steps = len(ph); phase_total_ms = float(ph[:, 5].mean())
    if dist is not None:
        t = torch.tensor([step_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); step_ms = float(t.item())
    value = world_size * BODIES / (step_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (upload state, step, download state) --------------------------
    st_in = w.states().copy(); st_out = np.zeros_like(st_in)
    for _ in range(2): w.step_host(dt, st_in, st_out); st_in, st_out = st_out, st_in
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        w.step_host(dt, st_in, st_out); st_in, st_out = st_out, st_in
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([t_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); t_e2e = float(t.item())
    e2e = world_size * BODIES * args.steps / t_e2e

    # ---- N > 1: ONE world with its pair search sharded over the ranks (BASELINE configs[4] shape, 1M bodies per GPU) -----
    sharded = batch = None
    if dist is not None:
        try:
            sharded = run_sharded_world(npb, torch, dist, rank, local_rank, world_size)
        except Exception as e:                                       # never lose the headline line to an auxiliary leg
            sharded = {"error": repr(e)}
        try:
            batch = run_worlds_batch(npb, torch, dist, rank, local_rank, world_size)
        except Exception as e:
            batch = {"error": repr(e)}
    if rank != 0:
        w.close()
        if dist is not None: dist.destroy_process_group()
        return
    # ---- rooflines -----------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp32_peak = w.fp32_peak_tflops()
    _, dcnt = narrow_flops(c0, c1)
    flops, _ = narrow_flops(p0, p1); flops = flops * args.steps / phase_steps          # per args.steps steps, from the phase-timed pass
    narrow_ms = float(np.mean(per[:, 2])); integ_ms = float(np.mean(per[:, 0]))
    ach_tf = flops / args.steps / (narrow_ms * 1e-3) / 1e12
    ach_gbs = BODIES * B_INTEGRATE / (integ_ms * 1e-3) / 1e9
    roof = {"bound": "fp32", "kernel": "k_gjk + k_epa (narrow phase: %.0f%% of the phase-timed step)" % (100 * narrow_ms / phase_total_ms),
            "timing": "CUDA events around the narrow-phase kernels in a second pass of %d steps with NP_OPT_PHASE_TIMING (step %.4f ms with the events, %.4f ms without)" % (phase_steps, phase_total_ms, step_ms),
            "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak if fp32_peak else None,
            "traffic": NARROW_DRAM_BYTES_PER_STEP, "traffic_source": "ncu --set full, dram__bytes_read+write of k_gjk + k_epa in one C2 step (profiles/r1_ncu_full_world_step.txt); the world-vertex cache alone is 13.6 MB",
            "peak_source": "measured in this run: non-FMA FADD/FMUL issue rate (np_selftest_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
            "algorithmic_flop_per_step": flops / args.steps}
    roof_hbm = {"bound": "hbm", "kernel": "k_integrate", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": None,
                "peak_source": hbm_src, "algorithmic_bytes_per_body": B_INTEGRATE}

    # ---- GJK pair-tests/s: explicit pair list of the same two shapes, device resident, larger than L2 ------------------
    npairs = 2_000_000
    specs, sa, pa, sb, pb = scenes.random_pairs(npairs, seed=99, mix=(0.5, 0.5), shapes=[("box", 1.0, 1.0, 1.0), ("sphere", 0.5)])
    with W(local_rank) as w2:
        w2.set_option(W.OPT_COUNTERS, 1)
        ids = np.array([w2.create_box(1.0, 1.0, 1.0), w2.create_sphere(0.5)], np.int32)
        d = [w2.dev_alloc(4 * npairs), w2.dev_alloc(24 * npairs), w2.dev_alloc(4 * npairs), w2.dev_alloc(24 * npairs), w2.dev_alloc(4 * npairs), w2.dev_alloc(16 * npairs)]
        w2.dev_upload(d[0], ids[sa]); w2.dev_upload(d[1], pa); w2.dev_upload(d[2], ids[sb]); w2.dev_upload(d[3], pb)
        for _ in range(3): w2.detect_batch_device(npairs, *d)
        p0 = w2.counters(); reps = 5
        pms = [w2.detect_batch_device(npairs, *d) for _ in range(reps)]
        p1 = w2.counters()
        pflops, _ = narrow_flops(p0, p1)
        pair_ms = float(np.mean(pms))
        pair = {"value": npairs / (pair_ms * 1e-3), "unit": "pair-tests/s", "workload": "%d random box/icosphere pairs (50/50 per side), positions U(-3,3)^3, device-resident inputs (%d MB > L2)" % (npairs, (npairs * 76) >> 20),
                "ms": pair_ms, "roofline": {"bound": "fp32", "achieved": pflops / reps / (pair_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                            "frac": pflops / reps / (pair_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None}}
    # ---- other BASELINE configs on one GPU (reported, not the headline): 1M-body world, 4096 x 256-body worlds -----------
    extra = {}
    if world_size == 1:
        try:
            big = scenes.lattice_scene(1_000_000, spacing=1.3, seed=43, mix=(0.7, 0.2, 0.1))
            with W(local_rank) as w3:
                w3.set_option(W.OPT_COUNTERS, 1); w3.set_option(W.OPT_PHASE_TIMING, 1)
                w3.load_scene(big)
                for _ in range(3): w3.step(big["dt"])
                k0 = w3.counters()
                ms3 = np.array([(w3.flush_l2(), w3.step(big["dt"]), w3.last_step_ms())[2] for _ in range(5)])
                k1 = w3.counters()
                tot = float(np.mean(ms3[:, 5])); integ = float(np.mean(ms3[:, 0])); nar = float(np.mean(ms3[:, 2]))
                f3, _ = narrow_flops(k0, k1)
                extra["world_1m_mixed"] = {"workload": "BASELINE configs[2] size: 1 000 000 bodies (70% box, 20% icosphere, 10% 16-vertex hull), one world, exact pair search",
                                           "value": 1e6 / (tot * 1e-3), "unit": "body-steps/s", "ms_per_step": tot,
                                           "narrow_phase": {"bound": "fp32", "ms": nar, "achieved": f3 / 5 / (nar * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                                            "frac": f3 / 5 / (nar * 1e-3) / 1e12 / fp32_peak if fp32_peak else None, "pair_tests_per_step": (k1["pair_tests"] - k0["pair_tests"]) / 5},
                                           "k_integrate": {"bound": "hbm", "achieved": 1e6 * B_INTEGRATE / (integ * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                           "frac": 1e6 * B_INTEGRATE / (integ * 1e-3) / 1e9 / hbm_peak, "ms": integ}}
            with W(local_rank) as w6:                            # configs[2] as worded: the same 1M world behind the uniform-grid broad phase
                w6.set_option(W.OPT_BROADPHASE, W.GRID); w6.set_option(W.OPT_COUNTERS, 1); w6.set_option(W.OPT_PHASE_TIMING, 1)
                w6.load_scene(big)
                for _ in range(3): w6.step(big["dt"])
                g0 = w6.counters()
                ms6 = np.array([(w6.flush_l2(), w6.step(big["dt"]), w6.last_step_ms())[2] for _ in range(5)])
                g1 = w6.counters()
                tot = float(np.mean(ms6[:, 5])); bp = float(np.mean(ms6[:, 1])); cands = (g1["candidates"] - g0["candidates"]) / 5
                extra["world_1m_mixed_grid"] = {"workload": "BASELINE configs[2]: the same 1 000 000 bodies with NP_BROADPHASE_GRID (AABB culling changes the collision list: the reference has no broad phase, DESIGN.md section 7)",
                                                "value": 1e6 / (tot * 1e-3), "unit": "body-steps/s", "ms_per_step": tot,
                                           "narrow_phase": {"bound": "fp32", "ms": nar, "achieved": f3 / 5 / (nar * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                                            "frac": f3 / 5 / (nar * 1e-3) / 1e12 / fp32_peak if fp32_peak else None, "pair_tests_per_step": (k1["pair_tests"] - k0["pair_tests"]) / 5}, "broadphase_ms": bp, "narrow_ms": float(np.mean(ms6[:, 2])),
                                                "candidate_pairs_per_step": cands, "pair_tests_per_step": (g1["pair_tests"] - g0["pair_tests"]) / 5,
                                                "candidate_pairs_per_sec_broadphase": cands / (bp * 1e-3)}
            wsc = scenes.lattice_scene(256, spacing=1.2, seed=1000)
            with W(local_rank) as w4:
                ids = np.array([w4.create_box(1.0, 1.0, 1.0), w4.create_sphere(0.5)], np.int32)
                for k in range(4096):
                    if k: w4.begin_island()
                    sck = scenes.lattice_scene(256, spacing=1.2, seed=1000 + k) if k < 64 else wsc      # 64 distinct worlds, then repeats (setup time)
                    w4.create_bodies(ids[sck["body_shape"]], npb.lib.sd_rows_to_struct(sck["sd"], sck["response"]))
                for _ in range(3): w4.step(wsc["dt"])
                ms4 = np.array([(w4.flush_l2(), w4.step(wsc["dt"]), w4.last_step_ms())[2] for _ in range(5)])
                tot = float(np.mean(ms4[:, 5]))
                extra["worlds_4096x256"] = {"workload": "BASELINE configs[3] on one GPU: 4096 independent 256-body worlds (islands) in one store, box/icosphere",
                                            "value": 4096 * 256 / (tot * 1e-3), "unit": "body-steps/s", "ms_per_step": tot}
            with W(local_rank, variant="fma") as w5:             # the same C2 world on the FMA-contracted build (poses within 1e-5 relative)
                w5.load_scene(scene)
                for _ in range(3): w5.step(dt)
                ms5 = np.array([(w5.flush_l2(), w5.step(dt), w5.last_step_ms())[2] for _ in range(10)])
                extra["c2_fma_build"] = {"workload": "the headline world on libnetphys_b200_fma.so (-fmad=true; not bit-exact: tests/test_gpu_fma.py holds it to 1e-5 relative on poses)",
                                         "value": BODIES / (float(np.mean(ms5[:, 5])) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(np.mean(ms5[:, 5]))}
        except Exception as e:                                   # never lose the headline line to an auxiliary leg
            extra["error"] = repr(e)
    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------------------------------
    cpu = None
    if world_size == 1:
        cs = 24
        t, kind, how = cpu_reference_steps(scene, cs, 1)
        cpu = {"value": BODIES * cs / t, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": "%d steps of the same %d-body world; %s" % (cs, BODIES, how)}
    out = {"metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "bodies_per_gpu": BODIES, "worlds_per_gpu": 1, "l2": "flushed before every timed step (256 MiB memset on the launching stream)",
                      "timing": "CUDA events on the world's stream around each step, mean over steps, max over ranks" + ("; all-gather of step k overlapped with step k+1, exposed part charged (see bench.py)" if world_size > 1 else ""),
                      "pair_search": "exact (reference semantics)", "collective": "nccl all_gather of the 36 B/body snapshot per step" if world_size > 1 else "none"},
           "clocks": clk, "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
           "e2e": {"value": e2e, "unit": "body-steps/s", "h2d_bytes_per_step": 48 * BODIES, "d2h_bytes_per_step": 48 * BODIES,
                   "how": "np_world_step_host: host state in, step, host state out (copies inside the call), wall clock over %d steps" % args.steps},
           "roofline": roof, "roofline_passes": [roof, roof_hbm], "pair_tests": pair,
           "phase_ms": {"integrate": integ_ms, "broadphase": float(np.mean(per[:, 1])), "narrow": narrow_ms, "resolve": float(np.mean(per[:, 3])),
                        "snapshot": float(np.mean(per[:, 4])), "allgather_exposed": float(np.mean(per[:, 6]))},
           "per_step_counts": {k: v / args.steps for k, v in dcnt.items() if k not in ("steps", "kernel_launches")},
           "wall_s_timed_region": t_wall, "other_configs": extra}
    if sharded is not None:
        out["other_configs"]["sharded_world"] = sharded
        out["other_configs"]["worlds_4096x256_sharded"] = batch
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    w.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
