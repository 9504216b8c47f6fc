"""First-contact GPU probe: parity of the CUDA path against the C oracle + rough timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netphys_b200 as npb
from netphys_b200 import scenes
from oracle.bind import Oracle

def biteq(a, b): return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
o = Oracle()
w = npb.World(0)
# sincos
g = np.load("tests/golden/sincos_hash.npz")
ok = True
for b in (0, 1, 62, 63, 64, 66, 67, 127, 128, 190, 194, 195, 255):
    hs, hc = w.sincos_hash(b << 24, 1 << 24)
    if hs != int(g["sin"][b]) or hc != int(g["cos"][b]): ok = False; print("sincos block", b, "MISMATCH")
print("sincos sampled blocks:", "ok" if ok else "FAIL")
# solvers KAT
print("tetra T4", w.closest(4, [[0,0,20],[-20,-10,2],[-20,10,2],[20,0,2]]), o.closest_tetra([0,0,0],[0,0,20],[-20,-10,2],[-20,10,2],[20,0,2]))
# shapes
for r in (5.0, 0.5): print("sphere", r, biteq(w.make_sphere(r), o.make_sphere(r)))
print("box", biteq(w.make_box(50, 50, 2), o.make_box(50, 50, 2)))
w.close()
# scenes
for sc, steps in ((scenes.test_physics_scene(), 200), (scenes.test_3d_scene(), 20), (scenes.few_boxes_scene(), 200),
                  (scenes.lattice_scene(300, spacing=1.2, seed=7, mix=(0.4, 0.4, 0.2)), 30), (scenes.lattice_scene(2000, spacing=2.5, seed=42), 5)):
    with npb.World(0) as w:
        w.set_option(npb.World.OPT_COUNTERS, 1)
        w.load_scene(sc); ow = o.world(scenes.resolve(sc, o))
        good = True
        for k in range(steps):
            w.step(sc["dt"]); ao, bo, do = ow.step(sc["dt"], True)
            a, b, d = w.collisions_as_arrays()
            if not (np.array_equal(a, ao) and np.array_equal(b, bo) and biteq(d, do) and biteq(w.states(), ow.state())):
                print("  MISMATCH step", k, "collisions", len(a), len(ao))
                if len(a) == len(ao):
                    bad = np.where((d.view(np.uint32) != do.view(np.uint32)).any(axis=1) | (a != ao) | (b != bo))[0]
                    print("  bad rows", bad[:5], d[bad[:3]], do[bad[:3]])
                sb = np.where((w.states().view(np.uint32) != ow.state().view(np.uint32)).any(axis=1))[0]
                print("  bad states", sb[:5])
                good = False; break
        print(sc["name"], "ok" if good else "FAIL", w.counters(), ow.counters()[:2], "ms", w.last_step_ms())
# pairs
specs, sa, pa, sb, pb = scenes.random_pairs(200000, seed=1234, mix=(0.4, 0.4, 0.2))
shapes = scenes.resolve_specs(specs, o)
with npb.World(0) as w:
    ids = np.array([w.create_mesh(s) for s in shapes], np.int32)
    t = time.time(); hit, out = w.detect_batch(ids[sa], pa, ids[sb], pb); t1 = time.time() - t
    t = time.time(); hit, out = w.detect_batch(ids[sa], pa, ids[sb], pb); t2 = time.time() - t
ho, oo, ctr = o.detect_batch(shapes, sa, pa, sb, pb, nthreads=8)
print("pairs hit mism", int(np.sum(hit != ho)), "out mism", int(np.sum((out.view(np.uint32) != oo.view(np.uint32)).any(axis=1))), "host-call s", t1, t2)
bad = np.where((out.view(np.uint32) != oo.view(np.uint32)).any(axis=1))[0]
for b in bad[:5]: print(b, hit[b], ho[b], out[b], oo[b])
# timing: big pair batch device-resident
with npb.World(0) as w:
    ids = np.array([w.create_mesh(s) for s in shapes], np.int32)
    for mix, name in (((1, 0, 0), "box-box"), ((0.5, 0.5, 0), "box/sphere mix"), ((0, 1, 0), "sphere-sphere")):
        n = 2_000_000
        specs2, sa, pa, sb, pb = scenes.random_pairs(n, seed=99, mix=mix, shapes=specs)
        d_sa = w.dev_alloc(4 * n); d_sb = w.dev_alloc(4 * n); d_pa = w.dev_alloc(24 * n); d_pb = w.dev_alloc(24 * n); d_hit = w.dev_alloc(4 * n); d_out = w.dev_alloc(16 * n)
        w.dev_upload(d_sa, ids[sa]); w.dev_upload(d_sb, ids[sb]); w.dev_upload(d_pa, pa); w.dev_upload(d_pb, pb)
        ms = [w.detect_batch_device(n, d_sa, d_pa, d_sb, d_pb, d_hit, d_out) for _ in range(4)]
        print(name, "ms", ms, "pairs/s %.3e" % (n / (min(ms) * 1e-3)))
        for p in (d_sa, d_sb, d_pa, d_pb, d_hit, d_out): w.dev_free(p)
# timing: world step
for n, spacing in ((10000, 1.2), (100000, 1.2), (1000000, 1.2)):
    sc = scenes.lattice_scene(n, spacing=spacing, seed=42)
    with npb.World(0) as w:
        w.set_option(npb.World.OPT_COUNTERS, 1)
        w.load_scene(sc)
        for _ in range(3): w.step(sc["dt"])
        ms = w.last_step_ms(); c = w.counters()
        w.step_n(sc["dt"], 10); msn = w.last_step_ms()
        print("world", n, "step ms", ms, "graph 10 steps ms", msn[5], "pair tests/step", c["pair_tests"] / 3, "body-steps/s %.3e" % (n * 10 / (msn[5] * 1e-3)))
