#!/usr/bin/env python
"""bench.py -- headline benchmark of the netphys rigid-body hot path on B200 (driver contract: one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one Physics_Update (integrate -> ordered first-hit GJK/EPA pair search -> impulse response -> replication
snapshot) of ONE world of 1 000 000 mixed convex bodies (BASELINE.json configs[2] size; SURVEY 8(d) C3 generator), pair
search with the reference's exact semantics, so that `--impl reference` (the reference's own CPU Physics_Update on the
same world) does the same work.  With N > 1 every rank steps its own such world (independent worlds, weak scaling) and
the per-step snapshots (36 B/body) are all-gathered over NCCL, overlapped with the next step.
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

BODIES = int(os.environ.get("NP_BENCH_BODIES", "1000000"))
SPACING, JITTER, SEED, MIX = 2.0, 0.3, 43, (0.7, 0.2, 0.1)
WORKLOAD = ("C3 (BASELINE configs[2] size, SURVEY 8(d)): %d mixed convex bodies in ONE world -- 70%% unit boxes (8 vertices), 20%% r=0.5 icospheres "
            "(162 vertices), 10%% one random 16-vertex hull; jittered cubic lattice spacing 2.0 +-0.3, rotations U(-1,1)^3, seed 43 + rank "
            "(numpy RandomState, not std::mt19937), gravity -9.8, dt 1/30; pair search = reference semantics (all pairs i<j, stop at the "
            "first hit per body: the reference has no broad phase)" % BODIES)
CONFIG = {"workload": WORKLOAD, "bodies_per_gpu": BODIES, "worlds_per_gpu": 1, "pair_search": "exact (reference semantics)"}      # identical in both arms
# algorithmic flop model of the narrow phase (reference arithmetic, + - * / sqrt cmp = 1 each)
F_SUPPORT_VERT, F_GJK_STEP, F_EPA_TRI, F_EPA_ITER = 27, 120, 110, 60
# algorithmic bytes per body of the integrate+transform pass
B_INTEGRATE = 48 + 56 + 36 + 48


def log(msg):
    if os.environ.get("NP_BENCH_VERBOSE"):
        sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() - T0, msg)); sys.stderr.flush()


T0 = time.perf_counter()
_OUT = sys.stdout        # main() swaps in a private copy of the real stdout


def make_scene(rank=0, n=None):
    return scenes.lattice_scene(n or BODIES, spacing=SPACING, jitter=JITTER, seed=SEED + rank, mix=MIX)


def narrow_flops(c0, c1):
    d = {k: c1[k] - c0[k] for k in c1}
    return F_SUPPORT_VERT * d["support_verts"] + F_GJK_STEP * d["gjk_steps"] + F_EPA_TRI * d["epa_tri_tests"] + F_EPA_ITER * d["epa_iters"], d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The timed region is tens of milliseconds
    and nvidia-smi needs longer than that to start (several ranks start one each): it is started BEFORE the warm-up, mark() notes where
    the timed region begins, and the samples from there to 0.15 s after its end are the ones reported."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []; self.p = None; self.gpu = gpu_index; self.t0 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append((time.perf_counter(), f))

    def mark(self):
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15); self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        allrows = list(self.rows)
        rows = [f for (t, f) in allrows if self.t0 is None or t >= self.t0]
        window = "from the start of the timed region to 0.15 s after its end"
        if not rows and allrows:
            rows = [f for (t, f) in allrows[-3:]]; window = "the last samples before the timed region (warm-up steps of the same world; none arrived inside it)"
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


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


def _cpu_world_worker(args):
    """one process = one copy of the reference engine (its body list is a process global): steps `count` 256-body worlds of configs[3] one after the other"""
    first, count, steps = args
    sys.path.insert(0, ROOT)
    from oracle.bind import Oracle, Ref, ref_available
    o = Oracle()
    t = 0.0
    for k in range(first, first + count):
        sc = scenes.resolve(scenes.lattice_scene(256, spacing=2.5, seed=1000 + k), o)
        if ref_available():
            r = Ref(); r.load_scene(sc); r.step(sc["dt"]); t += r.time_steps(sc["dt"], steps)
        else:
            w = o.world(sc); w.step(sc["dt"]); t0 = time.perf_counter()
            for _ in range(steps): w.step(sc["dt"])
            t += time.perf_counter() - t0
    return t


def cpu_reference_worlds(worlds_per_proc=4, steps=20):
    """SURVEY 8(d) CPU baseline (ii): configs[3] is embarrassingly parallel over worlds -- one PROCESS per world set, as many as the box has cores (8 at most)."""
    import multiprocessing as mp
    from oracle.bind import ref_available
    nproc = min(8, os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(nproc) as pool:
        busy = pool.map(_cpu_world_worker, [(p * worlds_per_proc, worlds_per_proc, steps) for p in range(nproc)])
    wall = time.perf_counter() - t0
    # the processes run side by side: throughput = all their body-steps over the longest one's stepping time (process start-up and scene building excluded)
    return {"value": nproc * worlds_per_proc * 256 * steps / max(busy), "unit": "body-steps/s", "cores": nproc, "kind": "reference" if ref_available() else "port",
            "sample": "%d processes x %d worlds x %d steps of the same generator (256 bodies per world, seeds 1000...), one copy of the engine per process; wall %.1f s incl. start-up" % (nproc, worlds_per_proc, steps, wall)}


def cpu_reference_pairs(specs, sa, pa, sb, pb, n1, n8):
    """DetectCollision over a pair list on the host cores: the compiled reference with 1 and with 8 std::threads (SURVEY 8(d) CPU baseline iii)."""
    from oracle.bind import Oracle, Ref, ref_available
    ncpu = os.cpu_count() or 1
    nt = min(8, ncpu)
    if ref_available():
        r = Ref(); r.L.ref_world_clear(); r.L.ref_shapes_clear()
        ids = np.array([r.sphere(float(s[1])) if s[0] == "sphere" else r.box(float(s[1]), float(s[2]), float(s[3])) if s[0] == "box" else r.mesh(s[1]) for s in specs], np.int32)
        t1, _, _ = r.time_detect(ids[sa[:n1]], pa[:n1], ids[sb[:n1]], pb[:n1], 1)
        t8, _, _ = r.time_detect(ids[sa[:n8]], pa[:n8], ids[sb[:n8]], pb[:n8], nt)
        kind = "reference"
    else:
        o = Oracle(); shapes = scenes.resolve_specs(specs, o)
        t0 = time.perf_counter(); o.detect_batch(shapes, sa[:n1], pa[:n1], sb[:n1], pb[:n1], nthreads=1); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); o.detect_batch(shapes, sa[:n8], pa[:n8], sb[:n8], pb[:n8], nthreads=nt); t8 = time.perf_counter() - t0
        kind = "port"
    return {"value": n8 / t8, "unit": "pair-tests/s", "cores": nt, "kind": kind, "sample": "the first %d pairs of the same list, DetectCollision on %d std::threads" % (n8, nt),
            "single_thread": {"value": n1 / t1, "cores": 1, "sample": "the first %d pairs, 1 thread" % n1}}


def run_reference(args, rank):
    if rank != 0:
        return
    scene = make_scene(0)
    t, kind, how = cpu_reference_steps(scene, args.steps, min(args.warmup, 1))
    v = BODIES * args.steps / t
    sample = "%d full steps of the %d-body world (%s)" % (args.steps, BODIES, how)
    _OUT.write(json.dumps({"impl": "reference", "metric": "body_steps_per_sec", "value": v, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": CONFIG,
                      "cpu_baseline": {"value": v, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}) + "\n"); _OUT.flush()


def timed_steps(w, dt, n, flush=True):
    """n steps with the L2 flushed before each; returns the [n, 6] phase/total times of np_world_last_step_ms."""
    out = []
    for _ in range(n):
        if flush: w.flush_l2()
        w.step(dt); out.append(w.last_step_ms())
    return np.array(out, np.float64)


def run_sharded_world(npb, torch, dist, rank, local_rank, world_size):
    """One world of world_size x 1M mixed bodies (C3/C5 generator) replicated on every rank; each rank runs the first-hit search of its
    index range, the collision records are all-gathered (the step's only exchange), every rank applies all responses."""
    from netphys_b200.dist import ShardedWorld
    per_gpu = int(os.environ.get("NP_BENCH_SHARDED_BODIES_PER_GPU", "1000000"))
    if world_size == 8 and "NP_BENCH_SHARDED_BODIES_PER_GPU" not in os.environ:
        per_gpu = 2097152                                            # BASELINE configs[4]: 16 777 216 bodies over 8 GPUs
    n = per_gpu * world_size
    sc = scenes.lattice_scene(n, spacing=SPACING, seed=SEED, mix=MIX)
    W = npb.World
    with W(local_rank) as w:
        w.set_option(W.OPT_PHASE_TIMING, 1)
        w.load_scene(sc)
        sw = ShardedWorld(w, torch.device("cuda", local_rank), world_size, rank)      # np_comm_init + np_world_shard_over_comm; step = np_world_step_sharded
        for _ in range(3): sw.step(sc["dt"])
        w.synchronize(); dist.barrier()
        ms = []
        for _ in range(5):
            w.flush_l2(); sw.step(sc["dt"]); ms.append(w.last_step_ms())
        ms = np.array(ms, np.float64).mean(0)
        st = w.states().view(np.uint32)
        h = np.array([int(np.bitwise_xor.reduce(st.ravel())), int(st.sum(dtype=np.uint64) & 0x7fffffff)], np.int64)
        a, b, _ = w.collisions_as_arrays()
    # the same world and the same eight steps PARTITIONED: every rank integrates / responds for its own bodies only and exchanges halos with
    # its neighbours (np_world_set_shard_halo); its own bodies must end up with the bits the replicated run gave them
    halo = int(os.environ.get("NP_BENCH_SHARD_HALO", "4096"))
    part = None
    with W(local_rank) as w2:
        w2.set_option(W.OPT_PHASE_TIMING, 1)
        w2.load_scene(sc)
        sw2 = ShardedWorld(w2, torch.device("cuda", local_rank), world_size, rank, halo=halo)
        for _ in range(3): sw2.step(sc["dt"])
        w2.synchronize(); dist.barrier()
        ms2 = []
        for _ in range(5):
            w2.flush_l2(); sw2.step(sc["dt"]); ms2.append(w2.last_step_ms())
        w2.synchronize()
        ms2 = np.array(ms2, np.float64).mean(0)
        slot = (n + world_size - 1) // world_size; lo_ = min(n, rank * slot); hi_ = min(n, lo_ + slot)
        same = bool(np.array_equal(w2.states().view(np.uint32)[lo_:hi_], st[lo_:hi_]))
        a2, _, _ = w2.collisions_as_arrays()
        t2 = torch.tensor(list(ms2) + [0.0 if same else 1.0, float(len(a2))], device="cuda", dtype=torch.float64)
        tmax = t2.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); tsum = t2.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tm = tmax.cpu().numpy()
        part = {"workload": "the same world with PARTITIONED steps (np_world_set_shard_halo = %d bodies): each rank runs Physics::Update, the first-hit search and OnCollision for its own "
                            "%d bodies, ncclSend/Recv of the halo transforms (48 B/body) and halo collision records (20 B/body) between neighbours, no all-gather" % (halo, slot),
                "value": n / (float(tm[5]) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(tm[5]),
                "phase_ms": {"integrate_own": float(tm[0]), "halo_transforms": float(tm[1]), "search_owned": float(tm[2]), "halo_records_and_respond_own": float(tm[3]), "snapshot": float(tm[4])},
                "own_bodies_bit_identical_to_the_replicated_run": bool(tm[6] == 0.0), "collisions_per_step_all_ranks": int(tsum.cpu().numpy()[7]),
                "timing": "CUDA events on the world's stream, mean of 5 steps, max over ranks; L2 flushed"}
    t = torch.tensor(ms, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.cpu().numpy()
    hs = torch.tensor(h, device="cuda"); lo = hs.clone(); hi = hs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"workload": "ONE world of %d bodies (the configs[2]/[4] generator, spacing 2.0) replicated on %d GPUs; exact first-hit search sharded by body index range, "
                        "one grouped ncclAllGather of the per-body collision records (hit_j + contact, 20 B/body) inside np_world_step_sharded, responses applied on every rank" % (n, world_size),
            "value": n / (float(ms[5]) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(ms[5]), "scaling": "weak (%d searched bodies per GPU; integrate/response passes are replicated)" % per_gpu,
            "phase_ms": {"integrate": float(ms[0]), "world_verts": float(ms[1]), "search_owned": float(ms[2]), "allgather_and_respond": float(ms[3]), "snapshot": float(ms[4])},
            "collisions_per_step": int(len(a)), "state_bit_identical_on_all_ranks": bool(torch.equal(lo, hi)), "l2": "flushed before every timed step", "timing": "CUDA events on the world's stream, mean of 5 steps, max over ranks",
            "partitioned": part}


def build_worlds_batch(npb, w, lo, hi):
    """worlds lo..hi-1 of BASELINE configs[3] (256-body C2-generator worlds, per-world seed 1000 + k) as islands of one store"""
    ids = np.array([w.create_box(1.0, 1.0, 1.0), w.create_sphere(0.5)], np.int32)
    dt = None
    for k in range(lo, hi):
        if k > lo: w.begin_island()
        sck = scenes.lattice_scene(256, spacing=2.5, seed=1000 + k)
        w.create_bodies(ids[sck["body_shape"]], npb.lib.sd_rows_to_struct(sck["sd"], sck["response"]))
        dt = sck["dt"]
    return dt


def run_worlds_batch(npb, torch, dist, rank, local_rank, world_size):
    """BASELINE configs[3]: 4096 independent 256-body worlds sharded over the ranks (contiguous blocks of worlds, islands of one store
    per rank), no data-path collective; the per-step snapshot (36 B/body) is all-gathered for replication by np_world_step itself."""
    from netphys_b200.dist import init_world_comm, shard_range
    lo, hi = shard_range(4096, rank, world_size)
    W = npb.World
    with W(local_rank) as w:
        dt = build_worlds_batch(npb, w, lo, hi)
        init_world_comm(w, rank, world_size)
        w.set_snapshot_gather(256 * -(-4096 // world_size), lo * 256)
        def run(nsteps):
            tot = 0.0; ker = 0.0
            for k in range(nsteps):
                w.flush_l2(); w.step(dt)
                s = float(w.last_step_ms()[5]); ker += s
                if k >= 1: s = max(s, float(w.last_gather_ms(1)[1]))
                tot += s
            tot += float(w.last_gather_ms(0)[1])
            return tot / nsteps, ker / nsteps
        run(3); dist.barrier()
        r = np.array(run(8), np.float64)
        w.comm_destroy()
    t = torch.tensor(r, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); r = t.cpu().numpy()
    return {"workload": "BASELINE configs[3]: 4096 independent 256-body worlds (C2 generator, seeds 1000..5095) sharded over %d GPUs, %d worlds per GPU as islands of one store; "
                        "snapshot all-gather (36 B/body) launched by np_world_step on the communicator's stream, one step behind, exposed part charged" % (world_size, hi - lo),
            "value": 4096 * 256 / (float(r[0]) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(r[0]), "ms_step_kernels_only": float(r[1]),
            "scaling": "strong (4096 worlds in total)", "timing": "CUDA events inside the library (np_world_last_step_ms / np_world_last_gather_ms), mean of 8, max over ranks; L2 flushed"}


def measured_traffic():
    """DRAM bytes of the narrow-phase kernels in one headline step, from a committed ncu capture taken with --cache-control none (so L2 keeps
    what the step's earlier kernels left there, as in a real step); None when no such capture exists for this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if int(t.get("bodies", 0)) == BODIES:
            return float(t["narrow_dram_bytes_per_step"]), t.get("source", "profiles/r2_traffic.json")
    except Exception:
        pass
    return None, "no ncu capture with --cache-control none for this workload is committed"


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
    log('scene built')
    W = npb.World
    w = W(local_rank)
    w.set_option(W.OPT_COUNTERS, 1)
    w.set_option(W.OPT_SNAPSHOT_EACH_STEP, 1)
    w.load_scene(scene)
    dt = scene["dt"]
    warm = max(args.warmup, 3)
    xchg = None
    if world_size > 1:
        from netphys_b200.dist import init_world_comm
        init_world_comm(w, rank, world_size)                          # the world's own NCCL communicator, behind the C ABI (np_comm_init)
        w.set_snapshot_gather(BODIES, rank * BODIES)                  # every np_world_step now all-gathers its snapshot on a high-priority stream
        xchg = True
    gather_mode = w.gather_mode()

    def overlapped(nsteps):
        """N > 1: np_world_step launches the all-gather of its snapshot on the communicator's own stream; it runs while the next step
        computes (double-buffered gather buffer), as a server would run it.  Step k is charged max(its kernels, the time the all-gather
        of step k-1 needed after step k-1's last kernel) -- step k may not finish before the previous snapshot is out -- all measured
        with CUDA events inside the library; the last all-gather's tail is added once."""
        per = np.zeros((nsteps, 2)); tot = 0.0
        for k in range(nsteps):
            w.flush_l2(); w.step(dt)
            s = float(w.last_step_ms()[5]); per[k, 0] = s          # (waits for the world's stream only)
            if k >= 1:
                lag = float(w.last_gather_ms(1)[1]); per[k, 1] = max(0.0, lag - s); s = max(s, lag)
            tot += s
        tail = float(w.last_gather_ms(0)[1]); per[-1, 1] += tail
        return (tot + tail) / nsteps, per

    log('world loaded, comm ready')
    clocks = ClockSampler(local_rank); clocks.start()              # (nvidia-smi is up and sampling by the time the warm-up is through)
    if xchg is None: timed_steps(w, dt, warm)
    else: overlapped(warm)
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    w.synchronize()
    clocks.mark()
    c0 = w.counters()
    t_wall0 = time.perf_counter()
    if xchg is None:
        per = timed_steps(w, dt, args.steps)
        step_ms = float(np.mean(per[:, 5])); kernels_ms = step_ms; ag_exposed = 0.0
    else:
        step_ms, per2 = overlapped(args.steps)
        kernels_ms = float(per2[:, 0].mean()); ag_exposed = float(per2[:, 1].mean())
    if dist is not None:
        torch.cuda.synchronize(); dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    log('timed region done: %.3f ms/step' % step_ms)
    c1 = w.counters()
    clk = clocks.stop()
    per_rank_ms = [step_ms]
    if dist is not None:
        t = torch.tensor([step_ms, kernels_ms, ag_exposed], device="cuda"); g = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(g, t)
        per_rank_ms = [float(x[0]) for x in g]; per_rank_kernels = [float(x[1]) for x in g]; per_rank_ag = [float(x[2]) for x in g]
        step_ms = max(per_rank_ms)
    value = world_size * BODIES / (step_ms * 1e-3)
    # Second pass with per-phase events (NP_OPT_PHASE_TIMING): the phase breakdown and the narrow-phase kernel time of the roofline.
    w.set_option(W.OPT_PHASE_TIMING, 1)
    timed_steps(w, dt, 2)
    p0 = w.counters()
    phase_steps = min(args.steps, 10)
    ph = timed_steps(w, dt, phase_steps)
    p1 = w.counters()
    w.set_option(W.OPT_PHASE_TIMING, 0)
    phase_total_ms = float(ph[:, 5].mean())

    # ---- end to end through the C ABI with host buffers (pinned): upload state, step, download state, every step ----------------
    st_a = w.host_array((BODIES, 12)); st_b = w.host_array((BODIES, 12))
    st_a[...] = w.states()
    for _ in range(2): w.step_host(dt, st_a, st_b); st_a, st_b = st_b, st_a
    if dist is not None:
        dist.barrier()
    e2e_steps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        w.step_host(dt, st_a, st_b); st_a, st_b = st_b, st_a
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([t_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); t_e2e = float(t.item())
    e2e = world_size * BODIES * e2e_steps / t_e2e
    w.host_free(st_a); w.host_free(st_b)
    log('e2e done')

    # ---- N > 1: ONE world with its pair search sharded over the ranks (BASELINE configs[4]); configs[3] sharded -----------------
    sharded = batch = None
    if dist is not None:
        fp32_peak = max(w.fp32_peaks_tflops()) if rank == 0 else None
        w.close(); w = None                                          # free the headline world before the large legs
        try:
            sharded = run_sharded_world(npb, torch, dist, rank, local_rank, world_size)
        except Exception as e:                                       # never lose the headline line to an auxiliary leg
            sharded = {"error": repr(e)}
        try:
            batch = run_worlds_batch(npb, torch, dist, rank, local_rank, world_size)
        except Exception as e:
            batch = {"error": repr(e)}
    else:
        fp32_peak = max(w.fp32_peaks_tflops())
    log('multi-GPU legs done')
    if rank != 0:
        if dist is not None: dist.destroy_process_group()
        return
    # ---- rooflines -----------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    _, dcnt = narrow_flops(c0, c1)
    flops, _ = narrow_flops(p0, p1); flops_per_step = flops / phase_steps
    narrow_ms = float(ph[:, 2].mean()); integ_ms = float(ph[:, 0].mean())
    ach_tf = flops_per_step / (narrow_ms * 1e-3) / 1e12
    ach_gbs = BODIES * B_INTEGRATE / (integ_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic()
    roof = {"bound": "fp32", "kernel": "narrow phase = k_gjk2 + k_epa_first + k_epa2 + the k_epa tiers (%.0f%% of the phase-timed step)" % (100 * narrow_ms / phase_total_ms),
            "executed_note": "`achieved` counts the REFERENCE's arithmetic (27 flop per vertex its support scan visits): with support-direction tables the kernels "
                             "evaluate only the few vertices that can win, so the executed FP32 work is far below this figure (ncu pipe_fma utilisation is in profiles/)",
            "timing": "CUDA events on the world's stream around the narrow-phase kernels, %d steps with NP_OPT_PHASE_TIMING (step %.4f ms with the events, %.4f ms without)" % (phase_steps, phase_total_ms, kernels_ms),
            "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak if fp32_peak else None,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured in this run: non-FMA FP32 issue rate, the larger of scalar FADD/FMUL chains and packed FADD2/FFMA2-as-multiply chains (np_selftest_fp32_peaks; they come out equal); MEASURED_PEAKS.json has no FP32 figure and the parity build may not contract FMAs",
            "algorithmic_flop_per_step": flops_per_step, "flop_model": "27/support vertex + 120/GJK step + 110/EPA triangle test + 60/EPA iteration, reference arithmetic (SURVEY 8(d)), counted by the kernels"}
    roof_hbm = {"bound": "hbm", "kernel": "k_integrate", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": None,
                "peak_source": hbm_src, "algorithmic_bytes_per_body": B_INTEGRATE, "ms": integ_ms}

    extra = {}
    pair = None
    if world_size == 1:
        # ---- GJK pair-tests/s: explicit pair list, device resident, larger than L2; CPU reference beside it --------------------
        try:
            npairs = 2_000_000
            shp = [("box", 1.0, 1.0, 1.0), ("sphere", 0.5)]
            specs, sa, pa, sb, pb = scenes.random_pairs(npairs, seed=99, mix=(0.5, 0.5), shapes=shp)
            with W(local_rank) as w2:
                w2.set_option(W.OPT_COUNTERS, 1)
                ids = np.array([w2.create_box(1.0, 1.0, 1.0), w2.create_sphere(0.5)], np.int32)
                d = [w2.dev_alloc(4 * npairs), w2.dev_alloc(24 * npairs), w2.dev_alloc(4 * npairs), w2.dev_alloc(24 * npairs), w2.dev_alloc(4 * npairs), w2.dev_alloc(16 * npairs)]
                def run_list(a, b):
                    w2.dev_upload(d[0], ids[a]); w2.dev_upload(d[1], pa); w2.dev_upload(d[2], ids[b]); w2.dev_upload(d[3], pb)
                    for _ in range(3): w2.detect_batch_device(npairs, *d)
                    q0 = w2.counters(); reps = 5
                    pms = float(np.mean([w2.detect_batch_device(npairs, *d) for _ in range(reps)]))
                    q1 = w2.counters()
                    pf, _ = narrow_flops(q0, q1)
                    tf = pf / reps / (pms * 1e-3) / 1e12
                    return {"value": npairs / (pms * 1e-3), "unit": "pair-tests/s", "ms": pms,
                            "roofline": {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak if fp32_peak else None}}
                pair = run_list(sa, sb)
                pair["workload"] = "%d random box/icosphere pairs (50/50 per side), positions U(-3,3)^3, rotations U(-pi,pi)^3, device-resident inputs (%d MB > L2)" % (npairs, (npairs * 76) >> 20)
                zero = np.zeros(npairs, np.int32)
                box = run_list(zero, zero)
                box["workload"] = "the same %d poses, every pair box-box" % npairs
                pair["box_only"] = box
            pair["cpu_baseline"] = cpu_reference_pairs(specs, sa, pa, sb, pb, 60000, 400000)
        except Exception as e:
            pair = {"error": repr(e)}
        # ---- other BASELINE configs on one GPU (reported, not the headline) -----------------------------------------------------
        try:
            c2 = scenes.lattice_scene(10000, spacing=2.5, jitter=0.3, seed=42)
            with W(local_rank) as w5:
                w5.set_option(W.OPT_COUNTERS, 1); w5.set_option(W.OPT_SNAPSHOT_EACH_STEP, 1)
                w5.load_scene(c2)
                timed_steps(w5, dt, 5)
                ms5 = timed_steps(w5, dt, 20)
                w5.set_option(W.OPT_PHASE_TIMING, 1); timed_steps(w5, dt, 2)
                k0 = w5.counters(); ph5 = timed_steps(w5, dt, 10); k1 = w5.counters()
                f5, _ = narrow_flops(k0, k1); nar5 = float(ph5[:, 2].mean())
                extra["c2_10k"] = {"workload": "C2 (BASELINE configs[1], SURVEY 8(d)): 10000 bodies in one world, unit boxes and r=0.5 icospheres alternating, lattice spacing 2.5 +-0.3, seed 42, exact pair search",
                                   "value": 10000 / (float(ms5[:, 5].mean()) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(ms5[:, 5].mean()),
                                   "phase_ms": {"integrate": float(ph5[:, 0].mean()), "world_verts": float(ph5[:, 1].mean()), "narrow": nar5, "resolve": float(ph5[:, 3].mean())},
                                   "narrow_phase": {"bound": "fp32", "achieved": f5 / 10 / (nar5 * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": f5 / 10 / (nar5 * 1e-3) / 1e12 / fp32_peak if fp32_peak else None},
                                   "note": "a latency benchmark on a B200: the step ends on a few dozen lone warps (DESIGN.md section 9)"}
            with W(local_rank, variant="fma") as w7:             # the same C2 world on the FMA-contracted build (poses within 1e-5 relative)
                w7.load_scene(c2)
                timed_steps(w7, dt, 3)
                ms7 = timed_steps(w7, dt, 10)
                extra["c2_10k_fma_build"] = {"workload": "the C2 world on libnetphys_b200_fma.so (-fmad=true; not bit-exact: tests/test_gpu_fma.py holds it to 1e-5 relative on poses)",
                                             "value": 10000 / (float(ms7[:, 5].mean()) * 1e-3), "unit": "body-steps/s", "ms_per_step": float(ms7[:, 5].mean())}
            with W(local_rank) as w6:                            # configs[2] as worded: the headline world behind the uniform-grid broad phase
                w6.set_option(W.OPT_BROADPHASE, W.GRID); w6.set_option(W.OPT_COUNTERS, 1); w6.set_option(W.OPT_PHASE_TIMING, 1)
                w6.load_scene(scene)
                timed_steps(w6, dt, 3)
                g0 = w6.counters(); ms6 = timed_steps(w6, dt, 5); g1 = w6.counters()
                tot = float(ms6[:, 5].mean()); bp = float(ms6[:, 1].mean()); nar6 = float(ms6[:, 2].mean()); cands = (g1["candidates"] - g0["candidates"]) / 5
                f6, _ = narrow_flops(g0, g1)
                extra["world_1m_grid"] = {"workload": "BASELINE configs[2] as worded: the headline world with NP_BROADPHASE_GRID (AABB culling changes the collision list: the reference has no broad phase, DESIGN.md section 7)",
                                          "value": BODIES / (tot * 1e-3), "unit": "body-steps/s", "ms_per_step": tot, "broadphase_ms": bp, "narrow_ms": nar6,
                                          "narrow_phase": {"bound": "fp32", "achieved": f6 / 5 / (nar6 * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": f6 / 5 / (nar6 * 1e-3) / 1e12 / fp32_peak if fp32_peak else None},
                                          "candidate_pairs_per_step": cands, "pair_tests_per_step": (g1["pair_tests"] - g0["pair_tests"]) / 5,
                                          "candidate_pairs_per_sec_broadphase": cands / (bp * 1e-3) if bp > 0 else None}
            with W(local_rank) as w4:
                dt4 = build_worlds_batch(npb, w4, 0, 4096)
                timed_steps(w4, dt4, 3)
                ms4 = timed_steps(w4, dt4, 5)
                tot = float(ms4[:, 5].mean())
                extra["worlds_4096x256"] = {"workload": "BASELINE configs[3] on one GPU: 4096 independent 256-body worlds (C2 generator, seeds 1000..5095) as islands of one store",
                                            "value": 4096 * 256 / (tot * 1e-3), "unit": "body-steps/s", "ms_per_step": tot}
            try:
                extra["worlds_4096x256"]["cpu_baseline"] = cpu_reference_worlds()
            except Exception as e:
                extra["worlds_4096x256"]["cpu_baseline"] = {"error": repr(e)}
        except Exception as e:                                   # never lose the headline line to an auxiliary leg
            extra["error"] = repr(e)
    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------------------------------
    cpu = None
    if world_size == 1:
        cs = 2 if BODIES >= 500000 else max(2, int(2e6 // BODIES))
        t, kind, how = cpu_reference_steps(scene, cs, 1)
        cpu = {"value": BODIES * cs / t, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": "%d steps (after 1 warm-up step) of the same %d-body world; %s" % (cs, BODIES, how)}
    out = {"metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": args.steps, "warmup": warm,
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": CONFIG,
           "method": {"l2": "flushed before every timed step (256 MiB memset on the launching stream); the world's arrays (~0.3 GB) exceed L2 anyway",
                      "timing": "CUDA events on the world's stream around each step, mean over steps, max over ranks" + ("; all-gather of step k overlapped with step k+1 on a high-priority stream, exposed part charged (see bench.py)" if world_size > 1 else ""),
                      "collective": {0: "none", 1: "ncclAllGather of the 36 B/body snapshot per step, issued by np_world_step itself over the world's communicator (np_comm_init; torch.distributed only distributes the unique id)",
                                     2: "all-gather of the 36 B/body snapshot per step on the COPY ENGINES: np_world_step pushes this rank's slot into every peer's gather buffer (mapped through CUDA IPC) with device-to-device copies over NVLink, framed by two 4-byte ncclAllGather exchanges on the world's communicator (np_comm_init; torch.distributed only distributes the unique id) -- no SM is taken from the step it overlaps"}[gather_mode]},
           "clocks": clk, "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
           "e2e": {"value": e2e, "unit": "body-steps/s", "h2d_bytes_per_step": 48 * BODIES, "d2h_bytes_per_step": 48 * BODIES,
                   "how": "np_world_step_host with page-locked host buffers (np_host_alloc): host state in, step, host state out, copies inside the call; wall clock over %d steps" % e2e_steps},
           "roofline": roof, "roofline_passes": [roof, roof_hbm],
           "phase_ms": {"integrate": integ_ms, "world_verts_or_broadphase": float(ph[:, 1].mean()), "narrow": narrow_ms, "resolve": float(ph[:, 3].mean()),
                        "snapshot": float(ph[:, 4].mean()), "allgather_exposed": ag_exposed},
           "per_rank_ms": per_rank_ms,
           "per_step_counts": {k: v / args.steps for k, v in dcnt.items() if k not in ("steps", "kernel_launches")},
           "wall_s_timed_region": t_wall, "other_configs": extra}
    if dist is not None:
        out["per_rank_kernels_ms"] = per_rank_kernels; out["per_rank_allgather_exposed_ms"] = per_rank_ag
        out["other_configs"]["sharded_world"] = sharded
        out["other_configs"]["worlds_4096x256_sharded"] = batch
    if pair is not None:
        out["pair_tests"] = pair
    if cpu is not None:
        out["cpu_baseline"] = cpu
    _OUT.write(json.dumps(out) + "\n"); _OUT.flush()
    if w is not None: w.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world_size = int(os.environ.get("WORLD_SIZE", "1"))
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on file descriptor 1), so the
    # descriptor is handed to stderr for the duration of the run and the line goes to a private copy of the real stdout.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
