"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the C oracle on the same seeded inputs
and against the golden fixtures produced by the reference.  Bit-exact everywhere: the path is IEEE binary32 in the
reference's operation order (no tolerance is needed or used)."""
import numpy as np
import pytest

from conftest import assert_biteq, gold
from netphys_b200 import scenes
from test_oracle import SCENES

pytestmark = pytest.mark.gpu
Z = [0, 0, 0]


# ---- primitives ---------------------------------------------------------------------------------------------
def test_sincos_exhaustive(world):
    """sinf/cosf on the device over ALL 2^32 float bit patterns == the oracle's (== libm's, exhaustively checked)."""
    g = gold("sincos_hash.npz"); blk = int(g["block"])
    for b in range(256):
        hs, hc = world.sincos_hash(b * blk, blk)
        assert hs == int(g["sin"][b]) and hc == int(g["cos"][b]), "block %d" % b


def test_solvers_known_answers(world, oracle):
    t4 = np.array([[0, 0, 20], [-20, -10, 2], [-20, 10, 2], [20, 0, 2]], np.float32)          # engine/test_util.cpp:32-59
    t5 = np.array([[0, 0, -2], [-20, -10, -4], [-20, 10, -4], [20, 0, -4]], np.float32)
    t6 = np.array([[0, 0, 2], [-20, -10, -4], [-20, 10, -4], [20, 0, -4]], np.float32)
    assert world.closest(4, t4)[0] == 14 and world.closest(4, t5)[0] == 1 and world.closest(4, t6)[0] == 0
    r, uvw = world.closest(2, [[0, 4, 0], [2, -4, 0]])                                         # :12-17
    z = np.array([0, 4, 0], np.float32) + (np.array([2, -4, 0], np.float32) - np.array([0, 4, 0], np.float32)) * uvw[0]
    assert_biteq(z, np.array([0.941176474, 0.235294104, 0.0], np.float32), "T2")
    g = gold("solvers.npz")
    for i in range(0, len(g["pts"]), 5):
        p = g["pts"][i]
        r, uvw = world.closest(2, p[:2]); assert r == g["line_region"][i]; assert_biteq(uvw[0], g["line_u"][i], "line")
        r, uvw = world.closest(3, p[:3]); assert r == g["tri_region"][i]; assert_biteq(uvw[:2], g["tri_uv"][i], "triangle")
        r, uvw = world.closest(4, p); assert r == g["tet_region"][i]; assert_biteq(uvw, g["tet_uvw"][i], "tetrahedron")


def test_shape_builders(world, oracle):
    g = gold("shapes.npz")
    assert_biteq(world.make_sphere(5.0), g["sphere5"], "CreateSphere(5)")
    assert_biteq(world.make_sphere(0.5), g["sphere05"], "CreateSphere(.5)")
    assert_biteq(world.make_box(50, 50, 2), g["box_50_50_2"], "CreateBox")
    assert_biteq(world.make_box(2, 3, 4), g["box_2_3_4"], "CreateBox depth ignored")
    assert_biteq(world.make_sphere(1.25), oracle.make_sphere(1.25), "CreateSphere(1.25)")


def test_support_and_transform(world, oracle):
    g = gold("pairs_ref.npz")
    ids = [world.create_box(1, 1, 1), world.create_sphere(0.5), world.create_mesh(g["hull"])]
    for i in range(0, 2000, 40):
        assert_biteq(world.support(ids[g["sa"][i]], g["pb"][i, :3], g["xf"][i]), g["support"][i], "support %d" % i)
    sc = scenes.few_boxes_scene(); world.load_scene(sc)
    for b in range(world.n):
        assert_biteq(world.transform(b), oracle.transform(sc["sd"][b, 3:6], sc["sd"][b, 6:9]), "GetTransform")


# ---- DetectCollision ----------------------------------------------------------------------------------------
def test_detect_collision_golden_pairs(world):
    g = gold("pairs_ref.npz")
    ids = np.array([world.create_box(1, 1, 1), world.create_sphere(0.5), world.create_mesh(g["hull"])], np.int32)
    hit, out = world.detect_batch(ids[g["sa"]], g["pa"], ids[g["sb"]], g["pb"])
    assert np.array_equal(hit, g["hit"])
    ok = g["out"][:, 4] == 1
    assert_biteq(out[ok], g["out"][ok], "converged pairs"); assert_biteq(out[~ok][:, [0, 1, 2, 4]], g["out"][~ok][:, [0, 1, 2, 4]], "non-converged pairs")


@pytest.mark.parametrize("mix,extent,n", [((1, 0, 0), 3.0, 60000), ((0, 1, 0), 3.0, 20000), ((0.3, 0.3, 0.4), 1.0, 60000), ((0.4, 0.4, 0.2), 8.0, 60000)])
def test_detect_collision_random_pairs_vs_oracle(world, oracle, mix, extent, n):
    specs, sa, pa, sb, pb = scenes.random_pairs(n, seed=int(extent * 10) + n, extent=extent, mix=mix)
    shapes = scenes.resolve_specs(specs, oracle)
    ids = np.array([world.create_mesh(s) for s in shapes], np.int32)
    hit, out = world.detect_batch(ids[sa], pa, ids[sb], pb)
    ho, oo, _ = oracle.detect_batch(shapes, sa, pa, sb, pb, nthreads=8)
    assert np.array_equal(hit, ho), "hit/miss set differs on %d pairs" % int(np.sum(hit != ho))
    assert_biteq(out, oo, "CollisionData")


def test_detect_collision_single_and_edge_cases(world, oracle):
    sph = world.create_sphere(5.0); brd = world.create_box(50, 50, 0.1)                    # engine/test_3d.cpp:410-449
    xa = oracle.transform([0, 3, 0], Z); xb = oracle.transform(Z, Z)
    hit, out = world.detect(sph, xa, brd, xb); ho, oo = oracle.detect(oracle.make_sphere(5.0), xa, oracle.make_box(50, 50, 0.1), xb)
    assert hit == ho; assert_biteq(out, oo, "test_3d pair")
    # identical poses, coincident shapes, far apart, NaN pose
    box = world.create_box(1, 1, 1); bo = oracle.make_box(1, 1, 1)
    for pa, pb in (([0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]), ([0, 0, 0, 0, 0, 0], [1e6, 0, 0, 0, 0, 0]), ([0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]),
                   ([0, 0, 0, 0, 0, 0], [np.nan, 0, 0, 0, 0, 0]), ([0, 0, 0, 500.0, -300.0, 1e5], [0.5, 0.5, 0.5, 0, 0, 0])):
        pa = np.array([pa], np.float32); pb = np.array([pb], np.float32)
        hit, out = world.detect_batch([box], pa, [box], pb); ho, oo, _ = oracle.detect_batch([bo], [0], pa, [0], pb)
        assert hit[0] == ho[0]
        assert np.array_equal(np.isnan(out), np.isnan(oo)); assert_biteq(np.nan_to_num(out), np.nan_to_num(oo), "edge pair")
    hit, out = world.detect_batch(np.zeros(0, np.int32), np.zeros((0, 6), np.float32), np.zeros(0, np.int32), np.zeros((0, 6), np.float32))
    assert len(hit) == 0


@pytest.mark.parametrize("g_gjk,g_epa", [(1, 4), (8, 8), (8, 1), (32, 4), (32, 32)])
def test_detect_collision_every_group_width_and_overflow_tier(npb, oracle, g_gjk, g_epa, monkeypatch):
    """Every group width G of the GJK / EPA kernels gives the same bits, and pairs whose EPA polytope outgrows the
    in-kernel arena (32 / 64 / 256 faces) are re-run in the overflow tiers without changing the result."""
    monkeypatch.setenv("NP_G_GJK", str(g_gjk)); monkeypatch.setenv("NP_G_EPA", str(g_epa))       # tuning overrides, read at world creation
    specs, sa, pa, sb, pb = scenes.random_pairs(120000, seed=4242, extent=0.6, mix=(0.0, 0.5, 0.5))
    shapes = scenes.resolve_specs(specs, oracle)
    ho, oo, ctr = oracle.detect_batch(shapes, sa, pa, sb, pb, nthreads=8)
    assert ctr["epa_max_faces"] > 64, "workload no longer exercises the overflow path"
    with npb.World(0) as w:
        ids = np.array([w.create_mesh(s) for s in shapes], np.int32)
        w.set_option(w.OPT_COUNTERS, 1)
        hit, out = w.detect_batch(ids[sa], pa, ids[sb], pb)
        if g_epa != 32:
            assert w.counters()["epa_overflow"] > 0
        c = w.counters()
    assert np.array_equal(hit, ho); assert_biteq(out, oo, "CollisionData incl. overflowed pairs")
    assert c["pair_tests"] == len(sa) and c["hits"] == int(ho.sum()) and c["gjk_steps"] == ctr["gjk_steps"] and c["epa_iters"] == ctr["epa_iters"]
    assert c["support_pairs"] == ctr["gjk_supports"] + ctr["epa_supports"] and c["epa_tri_tests"] == ctr["epa_tri_tests"]


# ---- Physics_Update -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(SCENES))
def test_world_step_golden_scenes(world, name):
    """BASELINE configs[0]: the reference's own scenes, every step's collision list and state == the reference's."""
    g = gold("scene_%s.npz" % name); sc = SCENES[name]()
    world.load_scene(sc); kept = {int(k): i for i, k in enumerate(g["kept_steps"])}
    off = 0
    for k in range(int(g["steps"])):
        world.step(sc["dt"]); n = int(g["col_count"][k])
        if k in kept or k % 50 == 0:
            a, b, d = world.collisions_as_arrays()
            assert np.array_equal(a, g["col_a"][off:off + n]) and np.array_equal(b, g["col_b"][off:off + n]), "collision list step %d" % k
            ok = g["col_data"][off:off + n, 10] == 1
            assert_biteq(d[ok], g["col_data"][off:off + n][ok], "collision data step %d" % k)
        if k in kept:
            assert_biteq(world.states(), g["states"][kept[k]], "state after step %d" % k)
        off += n
    assert_biteq(world.states(), g["final_state"], "final state")


@pytest.mark.parametrize("n,spacing,mix,steps", [(300, 1.2, (0.4, 0.4, 0.2), 25), (2000, 2.5, (0.5, 0.5, 0.0), 6), (5000, 0.9, (0.6, 0.1, 0.3), 4), (1, 1.0, (1, 0, 0), 3), (2, 0.4, (0, 1, 0), 3)])
def test_world_step_vs_oracle(world, oracle, n, spacing, mix, steps):
    sc = scenes.lattice_scene(n, spacing=spacing, seed=n + 1, mix=mix)
    world.set_option(world.OPT_COUNTERS, 1)
    world.load_scene(sc); ow = oracle.world(scenes.resolve(sc, oracle))
    for k in range(steps):
        world.step(sc["dt"]); ao, bo, do = ow.step(sc["dt"], True)
        a, b, d = world.collisions_as_arrays()
        assert np.array_equal(a, ao) and np.array_equal(b, bo), "collision list step %d" % k
        assert_biteq(d, do, "collision data step %d" % k); assert_biteq(world.states(), ow.state(), "state step %d" % k)
    pt, hits, _ = ow.counters(); c = world.counters()
    assert c["pair_tests"] == pt and c["hits"] == hits, "the device ran a different number of DetectCollision calls than the reference algorithm"


def test_world_islands_equal_independent_worlds(world, oracle):
    """BASELINE configs[3] in miniature: a batch of independent worlds in one store == each world stepped alone."""
    worlds = [scenes.lattice_scene(48, spacing=1.1, seed=1000 + k, mix=(0.4, 0.4, 0.2)) for k in range(12)]
    ows = [oracle.world(scenes.resolve(sc, oracle)) for sc in worlds]
    for k, sc in enumerate(worlds):
        if k: world.begin_island()
        world.load_scene(sc)
    for step in range(8):
        world.step(scenes.DT)
        st = world.states(); a, b, d = world.collisions_as_arrays()
        off = 0; cols = []
        for k, ow in enumerate(ows):
            ao, bo, do = ow.step(scenes.DT, True)
            assert_biteq(st[off:off + 48], ow.state(), "island %d step %d" % (k, step))
            cols.append((ao + off, bo + off, do)); off += 48
        assert np.array_equal(a, np.concatenate([c[0] for c in cols])) and np.array_equal(b, np.concatenate([c[1] for c in cols]))
        assert_biteq(d, np.concatenate([c[2] for c in cols]), "island collisions")


def test_world_api_semantics(world, oracle):
    sc = scenes.few_boxes_scene(); o_sc = scenes.resolve(sc, oracle)
    assert world.n == 0
    world.step(sc["dt"])                                         # empty world: a no-op, not an error
    assert len(world.collisions()) == 0
    world.load_scene(sc); ow = oracle.world(o_sc)
    assert_biteq(world.states(), ow.state(), "initial state (momenta zero, physics.cpp:11-18)")
    for _ in range(20): world.step(sc["dt"]); ow.step(sc["dt"])
    # step_n (one CUDA-graph submission) == n x step
    world.step_n(sc["dt"], 30)
    for _ in range(30): ow.step(sc["dt"])
    assert_biteq(world.states(), ow.state(), "step_n")
    # set_state / reset (Physics::Reset, physics.cpp:166-172) and single-body getters
    st = ow.state(); st[1, 0:3] = [0.5, 3.0, -0.25]; st[1, 6:9] = [0.1, -0.2, 0.3]; st[2, 9:12] = [0.01, 0.02, -0.03]
    ow.set_state(st); world.set_body_state(1, st[1]); world.set_body_state(2, st[2])
    world.reset_body(0); st = ow.state(); st[0] = 0; st[0, 0:3] = sc["sd"][0, 3:6]; st[0, 6 - 3:6] = sc["sd"][0, 6:9]; ow.set_state(st)
    assert_biteq(world.body_state(0), st[0], "reset body")
    for _ in range(10): world.step(sc["dt"]); ow.step(sc["dt"])
    assert_biteq(world.states(), ow.state(), "after set_state/reset (non-zero angular momentum -> rotation changes -> sin/cos on the device)")
    # host round trip (the e2e path)
    st_in = ow.state().copy(); out = np.zeros_like(st_in)
    world.step_host(sc["dt"], st_in, out); ow.step(sc["dt"])
    assert_biteq(out, ow.state(), "step_host")
    # bodies added after stepping keep the device state of the old ones
    extra = scenes.lattice_scene(6, spacing=1.0, seed=9, mix=(1, 0, 0))
    before = world.states()
    world.load_scene(extra)
    assert_biteq(world.states()[:len(before)], before, "old bodies after growth")
    world.step(sc["dt"])
    snap = world.snapshot()
    assert len(snap) == world.n and snap.dtype.itemsize == 36 and np.array_equal(snap["guid"], np.arange(world.n))
    assert_biteq(snap["pos"], world.states()[:, 0:3], "snapshot positions")
    assert np.allclose(np.linalg.norm(snap["rot"], axis=1), 1.0, atol=1e-5)


def test_impulse_response_general_normal(world, oracle):
    """ApplyImpulseResponse with a non-zero normal (the reference only ever passes 0, physics.cpp:99,109)."""
    sc = scenes.few_boxes_scene(); world.load_scene(sc)
    for _ in range(5): world.step(sc["dt"])
    st = world.body_state(1).astype(np.float32); st[9:12] = [0.3, -0.1, 0.2]; world.set_body_state(1, st)
    r = np.array([0.2, -0.9, 0.1], np.float32); n = np.array([0.0, 1.0, 0.0], np.float32); depth = np.float32(0.37)
    world.apply_impulse_response(1, r, depth, n)
    got = world.body_state(1)
    # same expression sequence in numpy float32 scalars
    F = np.float32; m = F(sc["sd"][1, 9]); e = F(sc["sd"][1, 10]); I = sc["sd"][1, 23:32].reshape(3, 3)
    L = st[6:9].copy(); H = st[9:12].copy(); pos = st[0:3].copy()
    def dot(a, b): return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))
    def cross(a, b): return np.array([F(F(a[1] * b[2]) - F(a[2] * b[1])), F(F(a[2] * b[0]) - F(a[0] * b[2])), F(F(a[0] * b[1]) - F(a[1] * b[0]))], F)
    def m3(v): return np.array([dot(I[0], v), dot(I[1], v), dot(I[2], v)], F)
    v = (L * F(F(1) / m)).astype(F); w = m3(H); vp = (v + cross(w, r)).astype(F)
    i = F(-(F(1) + e)) * dot(vp, n); k = F(F(F(1) / m) + dot(cross(m3(cross(r, n)), r), n)); j = F(i / k)
    L = (L + n * j).astype(F); H = (H + cross(r, n) * j).astype(F); pos = (pos + n * F(depth * F(1.01))).astype(F)
    assert_biteq(got[0:3], pos, "pos"); assert_biteq(got[6:9], L, "L"); assert_biteq(got[9:12], H, "H")


def test_friction_block_vs_reference_golden(world, oracle):
    """NP_OPT_FRICTION: the block the reference keeps commented out (physics.cpp:49-72) against what the reference computes with it
    uncommented (tests/golden/friction.npz from oracle/_ref/libnetphys_ref_friction.so) -- 512 bodies with their own masses, tensors,
    coefficients (0, NaN, negative), zero / axis / tangent / parallel normals, non-finite momenta -- and against the C port live.
    Option off: the plain build's answers."""
    g = gold("friction.npz"); nb = len(g["sd"])
    sc = scenes._scene([("box", 1.0, 1.0, 1.0)], np.zeros(nb, np.int32), list(g["sd"]), np.ones(nb, np.int32), "friction_cases")
    world.load_scene(sc)
    for on, key in ((1, "out_friction"), (0, "out_plain")):
        world.set_option(world.OPT_FRICTION, on); world.set_states(g["state"])
        ow = oracle.world(scenes.resolve(sc, oracle)); ow.set_state(g["state"]); ow.set_friction(on)
        for i in range(nb):
            world.apply_impulse_response(i, g["r"][i], g["depth"][i], g["n"][i])
            ow.apply_impulse_response(i, g["r"][i], g["depth"][i], g["n"][i])
        got = world.states()
        assert_biteq(got, g[key], key + " (reference)"); assert_biteq(got, ow.state(), key + " (oracle)")
        ow.close()


def test_friction_block_is_inert_in_world_steps(npb, oracle):
    """the world step passes the never-written zero normals (physics.cpp:99,109): with NP_OPT_FRICTION on, 300 steps of the few-boxes
    scene (static_friction_coeff 0.4 on every box) end where the reference's friction build ends -- which is where its plain build ends"""
    g = gold("friction.npz"); sc = scenes.few_boxes_scene()
    for fused in (False, True):
        w = npb.World(0); w.set_option(w.OPT_FRICTION, 1); w.load_scene(sc)
        if fused: w.step_n(sc["dt"], 300)
        else:
            for _ in range(300): w.step(sc["dt"])
        assert_biteq(w.states(), g["few_boxes_300"], "few boxes, 300 steps, friction on, fused=%s" % fused)
        w.close()


# ---- BASELINE sizes: size-independent properties --------------------------------------------------------------
def _check_first_hit_properties(world, oracle, sc, shapes, sample):
    """hit list ascending in a, a < b, one entry per a; on sampled bodies the oracle reproduces exactly the scan:
    every j in (i, hit_j) misses and hit_j hits, with the same CollisionData."""
    st = world.states(); a, b, d = world.collisions_as_arrays()
    assert np.all(np.diff(a) > 0) and np.all(b > a)
    hit_j = np.full(world.n, -1, np.int64); hit_j[a] = b
    rng = np.random.RandomState(0)
    for i in rng.choice(world.n - 1, size=sample, replace=False):
        end = hit_j[i] if hit_j[i] >= 0 else min(world.n - 1, i + 40)
        js = np.arange(i + 1, end + 1)
        if len(js) == 0: continue
        pose = lambda idx: np.concatenate([st[idx, 0:3], st[idx, 3:6]], axis=-1)
        ho, oo, _ = oracle.detect_batch(shapes, sc["body_shape"][np.full(len(js), i)], np.repeat(pose(i)[None], len(js), 0), sc["body_shape"][js], pose(js))
        if hit_j[i] >= 0:
            assert not ho[:-1].any() and ho[-1] == 1, "body %d: first hit differs" % i
            row = np.where(a == i)[0][0]
            assert_biteq(d[row, 6:11], oo[-1], "CollisionData of body %d" % i)
        else:
            assert not ho.any()


def test_baseline_c2_10k_world(world, oracle):
    """BASELINE configs[1]: 10k boxes/spheres in one world."""
    sc = scenes.lattice_scene(10000, spacing=1.2, seed=42)
    world.load_scene(sc); shapes = scenes.resolve_specs(sc["shape_specs"], oracle)
    ow = oracle.world(scenes.resolve(sc, oracle))
    for k in range(3):
        world.step(sc["dt"]); ow.step(sc["dt"])
    ao, bo, do = ow.step(sc["dt"], True); world.step(sc["dt"])
    a, b, d = world.collisions_as_arrays()
    assert np.array_equal(a, ao) and np.array_equal(b, bo); assert_biteq(d, do, "collisions"); assert_biteq(world.states(), ow.state(), "state")
    _check_first_hit_properties(world, oracle, sc, shapes, 200)


def test_baseline_c3_1m_world_properties(world, oracle, npb):
    """BASELINE configs[2] size (1M mixed convex bodies): determinism + sampled re-verification against the oracle."""
    sc = scenes.lattice_scene(1_000_000, spacing=1.3, seed=43, mix=(0.7, 0.2, 0.1))
    shapes = scenes.resolve_specs(sc["shape_specs"], oracle)
    world.load_scene(sc)
    world.step_n(sc["dt"], 2)
    st1 = world.states(); a1, b1, d1 = world.collisions_as_arrays()
    with npb.World(0) as w2:
        w2.load_scene(sc); w2.step(sc["dt"]); w2.step(sc["dt"])
        assert_biteq(w2.states(), st1, "two runs (graph replay vs plain launches) must be bit-identical")
        a2, b2, d2 = w2.collisions_as_arrays()
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2); assert_biteq(d1, d2, "collision data determinism")
    # positions: free fall is collision independent in the reference (normals are never set) -> closed form check on all bodies
    dt = sc["dt"]; g = np.float32(-9.8); L = np.float32(0); y = sc["sd"][:, 4].copy()
    for _ in range(2):
        L = np.float32(L + np.float32(np.float32(g * np.float32(1.0)) * dt)); y = (y + np.float32(np.float32(L * np.float32(1.0)) * dt)).astype(np.float32)
    assert_biteq(st1[:, 1], y, "y after 2 steps")
    _check_first_hit_properties(world, oracle, sc, shapes, 300)


# ---- NP_BROADPHASE_GRID (K2): the extension the reference only lists as a TODO -------------------------------------
def _world_aabbs(oracle, shapes, body_shape, st):
    """AABB of world*p over each body's vertices, float32, in the device's (= the reference support loop's) operation order."""
    lo = np.zeros((len(st), 3), np.float32); hi = np.zeros((len(st), 3), np.float32)
    for i in range(len(st)):
        m = oracle.transform(st[i, 0:3], st[i, 3:6]).reshape(4, 4); p = shapes[body_shape[i]]
        w = np.stack([((m[r, 0] * p[:, 0] + m[r, 1] * p[:, 1]) + m[r, 2] * p[:, 2]) + m[r, 3] for r in range(3)], axis=1).astype(np.float32)
        lo[i] = w.min(axis=0); hi[i] = w.max(axis=0)
    return lo, hi


@pytest.mark.parametrize("n,spacing,with_ground,cell,hashed", [(1500, 0.9, True, 0.0, 0), (3000, 1.1, False, 0.0, 0), (800, 0.8, True, 1.5, 0), (1500, 0.9, True, 0.0, 1), (900, 1.0, False, 0.35, 1),
                                                              (700, 1.0, True, 0.3, 0), (1500, 0.9, "first", 0.0, 0), (600, 0.9, "first", 0.4, 1)])
def test_grid_broadphase_candidate_parity(world, oracle, monkeypatch, n, spacing, with_ground, cell, hashed):
    """Candidate set == brute-force AABB overlap; per candidate pair the GJK/EPA result == the oracle's; first-hit order kept.  Both
    addressings of the grid: dense (cells ordered z, y, x: one run of boxes per x-row) and hashed (NP_GRID_HASH=1, what a world spread over
    more cells than buckets falls back to); cells from a third of a body (everything is "large") to several bodies wide."""
    if hashed: monkeypatch.setenv("NP_GRID_HASH", "1")
    sc = scenes.lattice_scene(n, spacing=spacing, seed=n, mix=(0.5, 0.2, 0.3))
    if with_ground == "first":   # the slab is body 0 and reaches into the lowest layers: ITS candidate list is hundreds long (spill region, k_cand_large's ordered fill)
        sc["shape_specs"] = sc["shape_specs"] + [("box", 40.0, 40.0, 1.0)]
        sc["sd"] = np.concatenate([scenes.static_data((0, 0, 0), (5.0, 0.3, 5.0), (0, 0, 0), 0.0, 0.0)[None], sc["sd"]]).astype(np.float32)
        sc["body_shape"] = np.concatenate([[len(sc["shape_specs"]) - 1], sc["body_shape"]]).astype(np.int32)
        sc["response"] = np.concatenate([[0], sc["response"]]).astype(np.int32)
    elif with_ground:   # a body much wider than a cell exercises the large-body list
        sc["shape_specs"] = sc["shape_specs"] + [("box", 40.0, 40.0, 1.0)]
        sc["sd"] = np.concatenate([sc["sd"], scenes.static_data((0, 0, 0), (5.0, -2.0, 5.0), (0, 0, 0), 0.0, 0.0)[None]]).astype(np.float32)
        sc["body_shape"] = np.concatenate([sc["body_shape"], [len(sc["shape_specs"]) - 1]]).astype(np.int32)
        sc["response"] = np.concatenate([sc["response"], [0]]).astype(np.int32)
    shapes = scenes.resolve_specs(sc["shape_specs"], oracle)
    world.set_option(world.OPT_BROADPHASE, world.GRID); world.set_option(world.OPT_COUNTERS, 1)
    if cell: world.set_option(world.OPT_GRID_CELL, int(np.float32(cell).view(np.uint32)))
    world.load_scene(sc)
    ow = oracle.world(scenes.resolve(sc, oracle))
    for step in range(2):
        world.step(sc["dt"]); world.synchronize()
        st = world.states(); a, b, d = world.collisions_as_arrays()
        ow.step(sc["dt"]); assert_biteq(st, ow.state(), "state (collision independent in the reference: normals are never set)")
        lo, hi = _world_aabbs(oracle, shapes, sc["body_shape"], st)
        N = len(st); exp = []; ncand = 0
        pose = np.concatenate([st[:, 0:3], st[:, 3:6]], axis=1)
        for i in range(N):
            js = np.arange(i + 1, N)
            ov = np.all((hi[i] >= lo[js]) & (hi[js] >= lo[i]), axis=1)
            js = js[ov]; ncand += len(js)
            if len(js) == 0: continue
            ho, oo, _ = oracle.detect_batch(shapes, sc["body_shape"][np.full(len(js), i)], np.repeat(pose[i][None], len(js), 0), sc["body_shape"][js], pose[js])
            h = np.flatnonzero(ho)
            if len(h): exp.append((i, js[h[0]], oo[h[0]]))
        assert len(a) == len(exp), "collision count %d vs %d" % (len(a), len(exp))
        assert np.array_equal(a, [e[0] for e in exp]) and np.array_equal(b, [e[1] for e in exp])
        assert_biteq(d[:, 6:11], np.stack([e[2] for e in exp]), "CollisionData per candidate pair")
        assert world.counters()["candidates"] == ncand * 1 + (0 if step == 0 else prev), "candidate set differs from brute-force AABB overlap"
        prev = world.counters()["candidates"]


# ---- edge cases ------------------------------------------------------------------------------------------------------
def test_world_extreme_states(world, oracle):
    """Huge / tiny / negative-zero Euler angles (both argument-reduction paths of sinf/cosf), NaN and inf positions,
    a massless IMPULSE body (1/0 -> inf*0 -> NaN, physics.cpp:32) -- IEEE semantics must carry through identically."""
    sc = scenes.lattice_scene(64, spacing=1.1, seed=5, mix=(0.4, 0.3, 0.3))
    sd = sc["sd"]
    sd[0, 6:9] = [500.0, -1.0e4, 3.0e5]; sd[1, 6:9] = [1e-20, -0.0, 119.99999]; sd[2, 6:9] = [120.0, 1e9, -3.3e38]
    sd[3, 3:6] = [np.nan, 0.0, 0.0]; sd[4, 3:6] = [np.inf, 1.0, 1.0]; sd[5, 9] = 0.0          # massless, response stays IMPULSE
    sd[6, 6:9] = [np.inf, 0.0, 0.0]; sd[7, 0:3] = [0.0, 0.0, 0.0]                               # no gravity: IsNone branch
    world.set_option(world.OPT_COUNTERS, 1)
    world.load_scene(sc); ow = oracle.world(scenes.resolve(sc, oracle))
    for k in range(6):
        world.step(sc["dt"]); ao, bo, do = ow.step(sc["dt"], True)
        a, b, d = world.collisions_as_arrays()
        assert np.array_equal(a, ao) and np.array_equal(b, bo), "collision list step %d" % k
        assert_biteq(d, do, "collision data step %d" % k); assert_biteq(world.states(), ow.state(), "state step %d" % k)
    assert world.counters()["pair_tests"] == ow.counters()[0]


def test_many_shapes_global_table_and_empty_shape(world, oracle):
    """More vertices than the shared-memory shape table holds (global-memory path) and a shape with no vertices
    (support returns (0,0,0), physics_util.cpp:12)."""
    rng = np.random.RandomState(3)
    shapes = [scenes.random_hull(rng, nverts=64, radius=0.6) for _ in range(40)] + [np.zeros((0, 3), np.float32)]
    n = 600
    sc = scenes.lattice_scene(n, spacing=1.0, seed=8, mix=(1, 0, 0))
    sc["shape_specs"] = [("mesh", s) for s in shapes]
    sc["body_shape"] = rng.randint(0, len(shapes), size=n).astype(np.int32)
    assert sum(len(s) for s in shapes) > 2048
    world.load_scene(sc); ow = oracle.world(scenes.resolve(sc, oracle))
    for k in range(3):
        world.step(sc["dt"]); ao, bo, do = ow.step(sc["dt"], True)
        a, b, d = world.collisions_as_arrays()
        assert np.array_equal(a, ao) and np.array_equal(b, bo); assert_biteq(d, do, "collision data"); assert_biteq(world.states(), ow.state(), "state")
    ids = np.arange(len(shapes), dtype=np.int32)
    specs, sa, pa, sb, pb = scenes.random_pairs(20000, seed=12, extent=1.5, mix=tuple([1] * len(shapes)), shapes=sc["shape_specs"])
    hit, out = world.detect_batch(ids[sa], pa, ids[sb], pb)
    ho, oo, _ = oracle.detect_batch(shapes, sa, pa, sb, pb, nthreads=8)
    assert np.array_equal(hit, ho); assert_biteq(out, oo, "pairs over 41 shapes")


def test_grid_with_islands_and_graph_replay(world, oracle, npb):
    """Grid broad phase + independent islands + step_n (CUDA graph) == the same stepped with plain launches."""
    worlds = [scenes.lattice_scene(200, spacing=0.9, seed=300 + k, mix=(0.5, 0.2, 0.3)) for k in range(6)]
    def build(w):
        w.set_option(w.OPT_BROADPHASE, w.GRID)
        for k, sc in enumerate(worlds):
            if k: w.begin_island()
            w.load_scene(sc)
    build(world)
    world.step_n(scenes.DT, 4); world.synchronize()
    a1, b1, d1 = world.collisions_as_arrays(); st1 = world.states()
    with npb.World(0) as w2:
        build(w2)
        for _ in range(4): w2.step(scenes.DT)
        a2, b2, d2 = w2.collisions_as_arrays()
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2); assert_biteq(d1, d2, "collisions"); assert_biteq(st1, w2.states(), "state")
    island = np.repeat(np.arange(6), 200)
    assert np.all(island[a1] == island[b1]), "a collision crosses islands"


def test_detect_collision_2d_mode(world, oracle):
    """is3D == false (engine/test_2d.cpp's mode) on the device == the oracle, per pair and through the single-query entry point."""
    from test_oracle import _scene_2d
    shapes, sa, pa, sb, pb = _scene_2d(np.random.RandomState(21), 6000)
    ids = np.array([world.create_mesh(s) for s in shapes], np.int32)
    hit, out = world.detect_batch(ids[sa], pa, ids[sb], pb, is3d=False)
    for i in range(len(sa)):
        ho, oo = oracle.detect(shapes[sa[i]], oracle.transform(pa[i, :3], pa[i, 3:]), shapes[sb[i]], oracle.transform(pb[i, :3], pb[i, 3:]), is3d=False)
        assert hit[i] == ho, "pair %d" % i
        assert_biteq(out[i], oo, "2-D CollisionData of pair %d" % i)
    h1, o1 = world.detect(ids[0], oracle.transform(pa[5, :3], pa[5, 3:]), ids[1], oracle.transform(pb[5, :3], pb[5, 3:]), is3d=False)
    h2, o2 = oracle.detect(shapes[0], oracle.transform(pa[5, :3], pa[5, 3:]), shapes[1], oracle.transform(pb[5, :3], pb[5, 3:]), is3d=False)
    assert h1 == h2; assert_biteq(o1, o2, "single 2-D query")


def test_command_frames_and_wire_packet(world, oracle):
    """netphys_server/world_s.cpp: the 5-second command-frame ring and the ClientWorldStateUpdatePacket wire layout."""
    import struct
    sc = scenes.few_boxes_scene(); world.load_scene(sc)
    assert world.fill_packet() == b""                              # no frame yet
    ids = []
    for k in range(8):
        world.step(sc["dt"]); ids.append(world.record_frame(1000.0 * k))
    assert ids == list(range(1, 9))
    assert world.frame_count() == 5                                # frames at 3000..7000 ms are younger than 5000 ms at now = 7000
    pkt = world.fill_packet()
    pid, size, fid = struct.unpack_from("<iii", pkt, 0); t, = struct.unpack_from("<d", pkt, 12); cnt, = struct.unpack_from("<i", pkt, 20)
    assert (pid, fid, t, cnt) == (2342144, 8, 7000.0, world.n) and size == len(pkt) - 8 == 16 + 36 * cnt
    objs = np.frombuffer(pkt, dtype=world.snapshot().dtype, offset=24, count=cnt)
    snap = world.snapshot()
    assert np.array_equal(objs["guid"], snap["guid"]); assert_biteq(objs["pos"], snap["pos"], "packet positions"); assert_biteq(objs["rot"], snap["rot"], "packet rotations")
    assert np.frombuffer(pkt, np.uint8, offset=24 + 32, count=1)[0] == 1          # isEnabled: bool true + padding
    assert world.fill_packet(cap=24 + 36 * cnt - 1) == b""         # Packet::Serialize refuses a short buffer
    assert struct.unpack_from("<i", world.fill_packet(packet_id=2342341), 0)[0] == 2342341


def test_adapter_extension_forces_enable_and_clear(world, oracle, npb):
    """np_body_add_force / np_body_set_enabled / np_world_clear_bodies (the ODE-facade hooks; this project's extension, mirrored by
    the oracle): accumulation order, one-step lifetime, pending forces surviving a store reallocation, graph replay, the snapshot flag."""
    rng = np.random.default_rng(11)
    sc = scenes.lattice_scene(900, spacing=1.2, seed=5, mix=(0.5, 0.3, 0.2))
    ids = world.load_scene(sc); ow = oracle.world(scenes.resolve(sc, oracle))
    n = 900
    for k in range(12):
        for body in rng.integers(0, n, 40):
            f = rng.uniform(-50, 50, 3).astype(np.float32)
            world.add_force(int(body), f); ow.add_force(int(body), f)
        if k == 3:
            for body in (5, 17, 400): world.set_enabled(body, False); ow.set_enabled(body, False)
        if k == 7:
            world.set_enabled(17, True); ow.set_enabled(17, True)
        if k == 5:        # grow the store past its capacity with forces pending: 900 -> 2300 bodies
            extra = scenes.lattice_scene(1400, spacing=1.2, seed=6)
            ex = scenes.resolve(extra, oracle)
            eids = np.array([world.create_sphere(s[1]) if s[0] == "sphere" else world.create_box(*s[1:]) for s in extra["shape_specs"]], np.int32)
            world.create_bodies(eids[extra["body_shape"]], npb.lib.sd_rows_to_struct(extra["sd"], extra["response"]))
            import ctypes as C
            fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
            sh0 = [ow.o.L.npo_world_add_shape(ow.w, fp(s), len(s)) for s in ex["shapes"]]
            for i in range(1400):
                ow.o.L.npo_world_add_body(ow.w, sh0[int(extra["body_shape"][i])], fp(extra["sd"][i]), int(extra["response"][i]))
            ow.n = n = 2300
        if k == 9: world.step_n(sc["dt"], 1)          # the captured graph consumes the pending forces too
        else: world.step(sc["dt"])
        ow.step(sc["dt"])
        assert_biteq(world.states(), ow.state(), "state, step %d" % k)
    snap = world.snapshot()
    assert list(np.flatnonzero(snap["is_enabled"] == 0)) == [5, 400] and world.is_enabled(17) and not world.is_enabled(5)
    world.step(sc["dt"]); ow.step(sc["dt"])          # no force added: nothing lingers
    assert_biteq(world.states(), ow.state(), "state after the forces were consumed")
    world.clear_bodies()
    assert world.n == 0 and len(world.collisions()) == 0
    world.step(sc["dt"])                              # an empty world steps
    few = scenes.few_boxes_scene(); shapes = scenes.resolve(few, oracle)
    fids = np.array([world.create_sphere(s[1]) if s[0] == "sphere" else world.create_box(*s[1:]) for s in few["shape_specs"]], np.int32)
    world.create_bodies(fids[few["body_shape"]], npb.lib.sd_rows_to_struct(few["sd"], few["response"]))
    ow2 = oracle.world(shapes)
    for k in range(5):
        world.step(few["dt"]); ow2.step(few["dt"])
    assert_biteq(world.states(), ow2.state(), "a world rebuilt after np_world_clear_bodies")


@pytest.mark.parametrize("g_gjk,g_epa,coop,grid", [(1, 4, 1, 0), (1, 4, 0, 0), (1, 1, 1, 1), (32, 8, 0, 0), (1, 4, 1, 1)])
def test_world_step_every_kernel_configuration(npb, oracle, g_gjk, g_epa, coop, grid, monkeypatch):
    """World mode with the kernels a LARGE world selects (thread-per-body GJK, 4-lane EPA, warp-cooperative supports for the large
    shapes of a mixed population), forced onto a world small enough for the oracle: collision list, contact data, state and counters."""
    monkeypatch.setenv("NP_G_GJK", str(g_gjk)); monkeypatch.setenv("NP_G_EPA", str(g_epa)); monkeypatch.setenv("NP_COOP", str(coop))
    sc = scenes.lattice_scene(2500, spacing=1.1, seed=77, mix=(0.6, 0.25, 0.15))
    with npb.World(0) as w:
        w.set_option(w.OPT_COUNTERS, 1)
        if grid: w.set_option(w.OPT_BROADPHASE, w.GRID)
        w.load_scene(sc)
        ow = oracle.world(scenes.resolve(sc, oracle))
        monkeypatch.delenv("NP_G_GJK"); monkeypatch.delenv("NP_G_EPA"); monkeypatch.delenv("NP_COOP")   # (read at world creation)
        if True:
            for k in range(5):
                w.step(sc["dt"])
                a, b, d = w.collisions_as_arrays()
                if not grid:
                    ao, bo, do = ow.step(sc["dt"], True)
                    assert np.array_equal(a, ao) and np.array_equal(b, bo), "collision list step %d" % k
                    assert_biteq(d, do, "collision data step %d" % k); assert_biteq(w.states(), ow.state(), "state step %d" % k)
            if grid:                                      # grid mode has no oracle list: the default kernels in grid mode are the reference
                with npb.World(0) as ref:
                    ref.set_option(ref.OPT_BROADPHASE, ref.GRID); ref.load_scene(sc)
                    for k in range(5): ref.step(sc["dt"])
                    a2, b2, d2 = ref.collisions_as_arrays()
                    assert np.array_equal(a, a2) and np.array_equal(b, b2); assert_biteq(d, d2, "collision data"); assert_biteq(w.states(), ref.states(), "state")
            else:
                pt, hits, _ = ow.counters(); c = w.counters()
                assert c["pair_tests"] == pt and c["hits"] == hits


# ---- np_world_step_n: the response of step k fused with the integration of step k+1 ---------------------------------------
@pytest.mark.parametrize("name,n", [("few_boxes", 30), ("test_physics", 30), ("lattice", 7)])
def test_step_n_fused_equals_single_steps(npb, oracle, name, n):
    """np_world_step_n(dt, n) runs k_resolve<FUSE> (OnCollision of step k + Physics::Update of step k+1, physics.cpp:27-46,175-194, in one
    pass): states, collision list and snapshot are bit-identical to n calls of np_world_step, and to the oracle."""
    sc = {"few_boxes": scenes.few_boxes_scene, "test_physics": scenes.test_physics_scene,
          "lattice": lambda: scenes.lattice_scene(3000, spacing=1.0, seed=21, mix=(0.5, 0.3, 0.2))}[name]()
    ow = oracle.world(scenes.resolve(sc, oracle))
    with npb.World(0) as a, npb.World(0) as b:
        for w in (a, b):
            w.set_option(w.OPT_SNAPSHOT_EACH_STEP, 1); w.load_scene(sc)
        for _ in range(n): a.step(sc["dt"])
        b.step_n(sc["dt"], n)
        for _ in range(n): last = ow.step(sc["dt"], True)
        assert_biteq(a.states(), b.states(), "fused vs single steps"); assert_biteq(b.states(), ow.state(), "fused vs oracle")
        ca, cb = a.collisions_as_arrays(), b.collisions_as_arrays()
        assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1]) and np.array_equal(cb[0], last[0]) and np.array_equal(cb[1], last[1])
        assert_biteq(ca[2], cb[2], "collision data"); assert_biteq(cb[2], last[2], "collision data vs oracle")
        assert a.snapshot().tobytes() == b.snapshot().tobytes()
        # a plain step after the fused run continues from the same state (the last step of step_n is not pre-integrated)
        a.step(sc["dt"]); b.step(sc["dt"]); ow.step(sc["dt"])
        assert_biteq(a.states(), b.states(), "step after step_n"); assert_biteq(b.states(), ow.state(), "step after step_n vs oracle")
        # forces and a disabled body go through the fused kernel's EXT variant
        for w in (a, b):
            w.add_force(0, np.array([3.0, -2.0, 1.0], np.float32)); w.set_enabled(1, False)
        ow.add_force(0, np.array([3.0, -2.0, 1.0], np.float32)); ow.set_enabled(1, False)
        for _ in range(3): a.step(sc["dt"]); ow.step(sc["dt"])
        b.step_n(sc["dt"], 3)
        assert_biteq(a.states(), b.states(), "fused EXT vs single steps"); assert_biteq(b.states(), ow.state(), "fused EXT vs oracle")


# ---- the device against the reference's OWN compiled sources (oracle/_ref), when the .so travelled to this box ----------------
def test_device_equals_compiled_reference_directly(world, ref):
    """No C port in between: DetectCollision over random pairs and Physics_Update over a mixed world, CUDA path vs oracle/_ref
    (the reference's engine/*.cpp compiled by oracle/Makefile)."""
    specs, sa, pa, sb, pb = scenes.random_pairs(30000, seed=77, extent=1.5, mix=(0.4, 0.3, 0.3))
    shapes = scenes.resolve_specs(specs, ref_factory(ref))
    ref.L.ref_world_clear(); ref.L.ref_shapes_clear()
    rids = np.array([ref.mesh(s) for s in shapes], np.int32)
    ids = np.array([world.create_mesh(s) for s in shapes], np.int32)
    hit, out = world.detect_batch(ids[sa], pa, ids[sb], pb)
    hr, orf = ref.detect_batch(rids[sa], pa, rids[sb], pb)
    assert np.array_equal(hit, hr)
    ok = (hit == 1) & (orf[:, 4] == 1)                      # converged hits: everything; other hits: depth is uninitialised in the reference
    assert_biteq(out[ok], orf[ok], "CollisionData of the converged hits vs the compiled reference")
    nk = (hit == 1) & (orf[:, 4] != 1)
    assert_biteq(out[nk][:, [0, 1, 2, 4]], orf[nk][:, [0, 1, 2, 4]], "CollisionData of the non-converged hits vs the compiled reference")
    sc = scenes.lattice_scene(400, spacing=1.1, seed=12, mix=(0.5, 0.3, 0.2))
    scr = dict(sc); scr["shapes"] = scenes.resolve_specs(sc["shape_specs"], ref_factory(ref))
    ref.load_scene(scr)
    with type(world)(0) as w:
        w.load_scene(sc)
        for k in range(12):
            w.step(sc["dt"]); a, b, d = w.collisions_as_arrays()
            ar, br, dr = ref.step(sc["dt"], True)
            assert np.array_equal(a, ar) and np.array_equal(b, br), "collision list vs the compiled reference, step %d" % k
            assert_biteq(d[:, 6:11], dr[:, 6:11], "contact data vs the compiled reference, step %d" % k)
            assert_biteq(w.states(), ref.state(), "state vs the compiled reference, step %d" % k)


class ref_factory:
    """scenes.resolve factory over the compiled reference's CreateSphere / CreateBox"""
    def __init__(self, ref): self.ref = ref
    def make_sphere(self, r): return self.ref.shape_verts(self.ref.sphere(r))
    def make_box(self, w, d, h): return self.ref.shape_verts(self.ref.box(w, d, h))


# ---- the wire packet against the REFERENCE's own serialiser bytes ---------------------------------------------
def test_wire_packets_equal_the_reference_serialiser(world):
    """np_world_record_frame + np_world_fill_packet vs tests/golden/packet_few_boxes.npz = the bytes netphys_server/world_s.cpp
    (World_S_Update's 5-second frame ring, World_S_FillWorldUpdateMessage / FillNewConnectionMessage) and netphys_common/common.h
    (Packet::Serialize, CommandFrameObject) produce for the same bodies -- compiled unchanged from the reference by oracle/Makefile,
    generated by oracle/gen_golden_packet.py."""
    g = gold("packet_few_boxes.npz"); sc = scenes.few_boxes_scene()
    world.load_scene(sc)
    world.set_snapshot_target(None, int(g["guid_base"]))
    n = int(g["n"])
    for k, t in enumerate(g["times_ms"]):
        if k < 8: world.step(sc["dt"])
        else: world.set_enabled(2, False)
        assert_biteq(world.states(), g["state_%d" % k], "state of frame %d" % k)
        assert world.record_frame(float(t)) == k + 1                                     # frame ids count from 1 (world_s.cpp:20)
        assert world.frame_count() == int(g["frames_kept_%d" % k]), "5 s frame ring after frame %d" % k
        assert world.fill_packet(2342144) == g["update_%d" % k].tobytes(), "ClientWorldStateUpdatePacket bytes, frame %d" % k
        assert world.fill_packet(2342341) == g["newconn_%d" % k].tobytes(), "ClientNewConnection bytes, frame %d" % k
        assert len(world.fill_packet(2342144, cap=24 + 36 * n - 1)) == int(g["too_small_%d" % k]) == 0


# ---- np_body_destroy / np_shape_destroy (the reference's destructor is a TODO, physics.cpp:20-23) -----------------------------
def test_body_and_shape_destroy(npb, oracle):
    """A destroyed body leaves the world as if it had never been created: the remaining bodies step exactly like an oracle world built
    without it (with its current states), ids of the others stay valid, collisions and snapshot guids report ids."""
    sc = scenes.lattice_scene(600, spacing=1.1, seed=31, mix=(0.5, 0.3, 0.2))
    gone = [0, 17, 18, 333, 599]
    with npb.World(0) as w:
        w.set_option(w.OPT_COUNTERS, 1)
        ids = w.load_scene(sc)
        ow = oracle.world(scenes.resolve(sc, oracle))
        for _ in range(3): w.step(sc["dt"]); ow.step(sc["dt"])
        w.add_force(5, np.array([1.0, 2.0, 3.0], np.float32)); ow.add_force(5, np.array([1.0, 2.0, 3.0], np.float32))      # pending force survives the rebuild
        before = w.states()
        for b in gone: w.destroy_body(b)
        with pytest.raises(npb.NetphysError): w.body_state(17)
        with pytest.raises(npb.NetphysError): w.destroy_body(17)
        keep = np.array([i for i in range(600) if i not in gone])
        assert w.n == 595
        assert_biteq(w.states(), before[keep], "states of the survivors")
        assert_biteq(w.body_state(19), before[19], "a surviving id still addresses its body")
        # the oracle world without those bodies, continued from the same states
        sc2 = dict(sc); sc2["body_shape"] = sc["body_shape"][keep]; sc2["sd"] = sc["sd"][keep]; sc2["response"] = sc["response"][keep]
        ow2 = oracle.world(scenes.resolve(sc2, oracle)); ow2.set_state(before[keep])
        ow2.add_force(int(np.searchsorted(keep, 5)), np.array([1.0, 2.0, 3.0], np.float32))
        for k in range(4):
            w.step(sc["dt"]); a, b, d = w.collisions_as_arrays()
            ao, bo, do = ow2.step(sc["dt"], True)
            assert np.array_equal(a, keep[ao]) and np.array_equal(b, keep[bo]), "collisions report ids, step %d" % k
            assert_biteq(d, do, "collision data"); assert_biteq(w.states(), ow2.state(), "state after destroy, step %d" % k)
        assert list(w.snapshot()["guid"]) == list(keep)
        # shapes: in use -> refused; unused -> gone, its id is not reused
        with pytest.raises(npb.NetphysError): w.destroy_shape(int(ids[0]))
        extra = w.create_box(2.0, 2.0, 2.0)
        w.destroy_shape(extra)
        with pytest.raises(npb.NetphysError): w.shape_vertices(extra)
        assert w.create_box(3.0, 3.0, 3.0) == extra + 1
        w.step(sc["dt"]); ow2.step(sc["dt"])
        assert_biteq(w.states(), ow2.state(), "state after a shape was destroyed")
        # a body created after destroys gets a fresh id
        nb = w.create_bodies(np.array([ids[0]], np.int32), npb.lib.sd_rows_to_struct(sc["sd"][:1], sc["response"][:1]))
        assert nb == 600 and w.n == 596


# ---- swept-AABB broad phase (todo.txt:8) ------------------------------------------------------------------------
def test_swept_broadphase_candidate_parity(npb, oracle):
    """NP_BROADPHASE_SWEPT: candidates = pairs whose SWEPT boxes (this step's AABB united with the previous step's) overlap -- checked
    against a brute force over the same boxes, a superset of the static-box candidates -- each tested at the current poses with the
    oracle's result, first-hit order kept; fast bodies that tunnel past each other between two steps are still candidates."""
    sc = scenes.lattice_scene(1200, spacing=1.0, seed=5, mix=(0.5, 0.2, 0.3))
    sc["sd"][::3, 1] = -400.0                               # strong gravity on a third of the bodies: they move ~0.4+ per step after a few steps
    shapes = scenes.resolve_specs(sc["shape_specs"], oracle)
    with npb.World(0) as w, npb.World(0) as wg:
        w.set_option(w.OPT_BROADPHASE, w.SWEPT); w.set_option(w.OPT_COUNTERS, 1); w.load_scene(sc)
        wg.set_option(wg.OPT_BROADPHASE, wg.GRID); wg.set_option(wg.OPT_COUNTERS, 1); wg.load_scene(sc)
        ow = oracle.world(scenes.resolve(sc, oracle))
        prev_box = None; prev_c = 0; prev_g = 0; grew = False
        for step in range(5):
            w.step(sc["dt"]); w.synchronize(); wg.step(sc["dt"]); wg.synchronize()
            st = w.states(); a, b, d = w.collisions_as_arrays()
            ow.step(sc["dt"]); assert_biteq(st, ow.state(), "state")
            lo, hi = _world_aabbs(oracle, shapes, sc["body_shape"], st)
            slo, shi = (lo, hi) if prev_box is None else (np.minimum(lo, prev_box[0]), np.maximum(hi, prev_box[1]))
            prev_box = (lo, hi)
            N = len(st); exp = []; ncand = 0; nstatic = 0
            pose = np.concatenate([st[:, 0:3], st[:, 3:6]], axis=1)
            for i in range(N):
                js = np.arange(i + 1, N)
                ov = np.all((shi[i] >= slo[js]) & (shi[js] >= slo[i]), axis=1)
                nstatic += int(np.all((hi[i] >= lo[js]) & (hi[js] >= lo[i]), axis=1).sum())
                js = js[ov]; ncand += len(js)
                if len(js) == 0: continue
                ho, oo, _ = oracle.detect_batch(shapes, sc["body_shape"][np.full(len(js), i)], np.repeat(pose[i][None], len(js), 0), sc["body_shape"][js], pose[js])
                h = np.flatnonzero(ho)
                if len(h): exp.append((i, js[h[0]], oo[h[0]]))
            c = w.counters()["candidates"]; cg = wg.counters()["candidates"]
            assert c - prev_c == ncand, "swept candidate set differs from the brute force over swept boxes (step %d)" % step
            assert cg - prev_g == nstatic and ncand >= nstatic
            grew |= ncand > nstatic
            prev_c, prev_g = c, cg
            assert len(a) == len(exp) and np.array_equal(a, [e[0] for e in exp]) and np.array_equal(b, [e[1] for e in exp])
            assert_biteq(d[:, 6:11], np.stack([e[2] for e in exp]), "CollisionData per swept candidate pair")
        assert grew, "the workload no longer produces pairs that only the swept boxes catch"
        w.step_n(sc["dt"], 3); ow.step(sc["dt"]); ow.step(sc["dt"]); ow.step(sc["dt"])        # captured + fused steps in swept mode
        assert_biteq(w.states(), ow.state(), "state after step_n in swept mode")
