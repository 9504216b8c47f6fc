"""GPU suite: support-direction tables (NP_OPT_SUPPORT_TABLES, np_narrow.cuh support_tab / np_suptab.h).

The table path must return the very vertex the reference's scan returns (PhysUtil_GetPointFurthestInDirection, physics_util.cpp:5-35:
strict '>' from -inf, first maximum wins) -- random directions, and the directions where it is hardest: exactly along vertices, edge
midpoints and face normals of the shape (ties in exact arithmetic, decided by rounding), axis directions with zero components, tiny and
huge magnitudes, bodies far from the origin (the rounding the table's margin has to cover grows with |t|), and bodies or directions
outside the table's guarantees, which must fall back to the scan."""
import numpy as np
import pytest

from conftest import assert_biteq
from netphys_b200 import scenes

pytestmark = pytest.mark.gpu


def hard_directions(verts, rng, n_rand):
    v = np.asarray(verts, np.float64)
    d = [rng.normal(size=(n_rand, 3))]
    d.append(v)                                                                  # through every vertex
    i = rng.randint(len(v), size=(3000, 3))
    d.append(v[i[:, 0]] + v[i[:, 1]])                                            # edge midpoints (and chords)
    d.append(v[i[:, 0]] + v[i[:, 1]] + v[i[:, 2]])                               # face centres
    d.append(np.cross(v[i[:, 1]] - v[i[:, 0]], v[i[:, 2]] - v[i[:, 0]]))         # face normals: every vertex of the face ties
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [1, -1, 0], [1, 1, -1], [1, -1, -1]], np.float64)
    d.append(np.concatenate([axes, -axes]))
    d = np.concatenate(d)
    d = np.concatenate([d, d * 1e-9, d * 1e7, d + rng.normal(size=d.shape) * 1e-7])   # scale must not matter; nearly-tied directions
    return d.astype(np.float32)


@pytest.mark.parametrize("spec", [("box", 1.0, 1.0, 1.0), ("sphere", 0.5), ("box", 50.0, 50.0, 2.0), ("sphere", 5.0), ("hull", 16), ("hull", 200), ("box", 2.0, 3.0, 4.0)])
def test_table_support_equals_scan(world, oracle, spec):
    rng = np.random.RandomState(11)
    if spec[0] == "box": sid = world.create_box(*spec[1:])
    elif spec[0] == "sphere": sid = world.create_sphere(spec[1])
    else: sid = world.create_mesh(scenes.random_hull(rng, nverts=spec[1]))
    verts = world.shape_vertices(sid)
    info = world.support_table_info(sid)
    assert info is not None and info["scan_cells"] <= 0.01 * 6 * info["n"] ** 2 and info["mean_candidates"] < 4.0, info
    dirs = hard_directions(verts, rng, 20000)
    n = len(dirs)
    # poses: identity, random rotations near the origin, and far from it (every coordinate up to 99 % of the table's reach)
    pos = rng.uniform(-3, 3, size=(n, 3)); rot = rng.uniform(-np.pi, np.pi, size=(n, 3))
    pos[: n // 8] = 0; rot[: n // 8] = 0
    far = rng.rand(n) < 0.5
    pos[far] = rng.uniform(-1, 1, size=(far.sum(), 3)) * info["reach"] * 0.99
    edge = far & (rng.rand(n) < 0.5)                                              # right at the limit: where the rounding the margin must cover is largest
    pos[edge] = rng.choice([-1.0, 1.0], size=(edge.sum(), 3)) * rng.uniform(0.9, 0.99, size=(edge.sum(), 3)) * info["reach"]
    xf = np.stack([oracle.transform(pos[k].astype(np.float32), rot[k].astype(np.float32)) for k in range(n)]).reshape(n, 16)
    # in shape space the hard directions stay hard: rotate them with the body
    R = xf.reshape(n, 4, 4)[:, :3, :3].astype(np.float64)
    wd = np.einsum("nij,nj->ni", R, dirs.astype(np.float64)).astype(np.float32)
    dirs_all = np.concatenate([dirs, wd]); xf_all = np.concatenate([xf, xf])
    scan, tab, used = world.support_table_check(sid, dirs_all, xf_all)
    assert_biteq(tab, scan, "table vs scan, %s" % (spec,))
    assert used.mean() > 0.95, used.mean()
    for k in rng.randint(len(dirs_all), size=300):                               # the scan itself against the oracle
        assert_biteq(scan[k], oracle.support(verts, dirs_all[k], xf_all[k].reshape(4, 4)), "scan vs oracle")


def test_table_falls_back_outside_its_guarantees(world, oracle):
    sid = world.create_sphere(0.5)
    info = world.support_table_info(sid)
    rng = np.random.RandomState(5)
    n = 4000
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    xf = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (n, 1))
    q = n // 8
    xf[0 * q:1 * q, 3] = info["reach"] * 2.0                                      # beyond the reach
    xf[1 * q:2 * q, 0] = 1.5                                                      # scaled: not a rotation
    xf[2 * q:3 * q, 1] = 0.3                                                      # sheared
    xf[3 * q:4 * q, 7] = np.nan                                                   # NaN translation
    dirs[4 * q:5 * q] = 0.0                                                       # zero direction: vertex 0 wins the scan
    dirs[5 * q:6 * q, 1] = np.inf
    dirs[6 * q:7 * q, 2] = np.nan
    dirs[7 * q:8 * q] *= 1e-30                                                    # below the table's magnitude window
    scan, tab, used = world.support_table_check(sid, dirs, xf)
    assert_biteq(tab, scan, "fallback")
    assert used.sum() == 0
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (n, 1))
    _, _, used2 = world.support_table_check(sid, rng.normal(size=(n, 3)).astype(np.float32), ident)
    assert used2.all()


def test_tables_do_not_change_a_world_step_or_pair_queries(npb, oracle):
    """the same world and pair list with NP_OPT_SUPPORT_TABLES on and off: identical bits (and both equal the oracle elsewhere in the suite)"""
    sc = scenes.lattice_scene(3000, spacing=1.2, seed=9, mix=(0.5, 0.3, 0.2))
    res = []
    for on in (1, 0):
        with npb.World(0) as w:
            w.set_option(w.OPT_SUPPORT_TABLES, on); w.set_option(w.OPT_COUNTERS, 1)
            w.load_scene(sc)
            for _ in range(5):
                w.step(sc["dt"])
            a, b, d = w.collisions_as_arrays()
            res.append((w.states().copy(), a.copy(), b.copy(), d.copy(), w.counters()))
    assert_biteq(res[0][0], res[1][0], "state"); assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert_biteq(res[0][3], res[1][3], "collision data")
    assert res[0][4] == res[1][4]                                                 # the counters count the reference's work either way
    specs, sa, pa, sb, pb = scenes.random_pairs(60000, seed=21, mix=(0.4, 0.4, 0.2))
    out = []
    for on in (1, 0):
        with npb.World(0) as w:
            w.set_option(w.OPT_SUPPORT_TABLES, on)
            ids = np.array([w.create_mesh(s) for s in scenes.resolve_specs(specs, oracle)], np.int32)
            out.append(w.detect_batch(ids[sa], pa, ids[sb], pb))
    assert np.array_equal(out[0][0], out[1][0]); assert_biteq(out[0][1], out[1][1], "pair queries")


@pytest.mark.parametrize("env", [{}, {"NP_GJK2": "0"}, {"NP_EPA2": "0"}, {"NP_EPA2_TAIL": "4"}, {"NP_TABLES": "0"}, {"NP_TAB_MIN": "17"}, {"NP_G_EPA": "1"}])
def test_thread_per_item_kernel_variants_equal_the_oracle(npb, oracle, env, monkeypatch):
    """The thread-per-item path forced onto a small workload (NP_G_GJK=1) with each of its kernels switched for the one it replaced, on shapes
    that have tables and one that cannot (300 vertices: scanned by the whole warp): pair list and world step == the oracle, counters included."""
    monkeypatch.setenv("NP_G_GJK", "1"); monkeypatch.setenv("NP_G_EPA", "4")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.RandomState(3)
    specs = [("box", 1.0, 1.0, 1.0), ("sphere", 0.5), ("mesh", scenes.random_hull(rng, nverts=16)), ("mesh", scenes.random_hull(rng, nverts=300))]
    _, sa, pa, sb, pb = scenes.random_pairs(60000, seed=77, extent=0.8, mix=(0.4, 0.3, 0.2, 0.1), shapes=specs)
    shapes = scenes.resolve_specs(specs, oracle)
    ho, oo, ctr = oracle.detect_batch(shapes, sa, pa, sb, pb, nthreads=8)
    with npb.World(0) as w:
        ids = np.array([w.create_mesh(s) for s in shapes], np.int32)
        w.set_option(w.OPT_COUNTERS, 1)
        hit, out = w.detect_batch(ids[sa], pa, ids[sb], pb)
        c = w.counters()
    assert np.array_equal(hit, ho); assert_biteq(out, oo, "pair list %s" % (env,))
    assert c["pair_tests"] == len(sa) and c["hits"] == int(ho.sum()) and c["gjk_steps"] == ctr["gjk_steps"] and c["epa_iters"] == ctr["epa_iters"]
    assert c["support_pairs"] == ctr["gjk_supports"] + ctr["epa_supports"] and c["epa_tri_tests"] == ctr["epa_tri_tests"]
    sc = scenes.lattice_scene(1500, spacing=1.1, seed=8, mix=(0.5, 0.3, 0.2))
    rs = scenes.resolve(sc, oracle)
    with npb.World(0) as w:
        w.load_scene(sc); ow = oracle.world(rs)
        for _ in range(4):
            w.step(sc["dt"]); ow.step(sc["dt"])
        assert_biteq(w.states(), ow.state(), "world state %s" % (env,))
