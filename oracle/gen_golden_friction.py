"""oracle/gen_golden_friction.py -- TEST INFRASTRUCTURE.  tests/golden/friction.npz: Physics::ApplyImpulseResponse with the friction
block of physics.cpp:49-72 in force, as computed by the REFERENCE ITSELF -- oracle/_ref/libnetphys_ref_friction.so is the reference's
engine compiled with the `//` markers of those 24 lines stripped on the fly (oracle/Makefile), nothing else changed.  The plain build's
answers for the same inputs are stored next to them (`out_plain`), and the few-boxes world stepped 300 times by both builds (the world
step passes zero normals, physics.cpp:99,109 with collision_detection.cpp:533-543, so the block must change nothing there).

    python oracle/gen_golden_friction.py          (dev container only: needs oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netphys_b200 import scenes  # noqa: E402
from oracle.bind import Oracle, Ref  # noqa: E402

F = np.float32
N = 512


def friction_cases(seed=7):
    """N bodies (one unit-box shape), each with its own StaticPhysicsData, state, contact and normal."""
    rng = np.random.RandomState(seed)
    rows = []
    for i in range(N):
        mass = F(rng.uniform(0.5, 4.0)); e = F(rng.uniform(0.0, 1.0))
        mu = F([0.0, 0.4, rng.uniform(0.05, 1.5), 0.4][i % 4])
        sd = scenes.static_data((0.0, -9.8, 0.0), rng.uniform(-5, 5, 3), rng.uniform(-3, 3, 3), mass, e, mu, 0.0, F(1.0) / F(6.0))
        a = rng.uniform(-0.3, 0.3, (3, 3)); inv = np.diag(rng.uniform(0.5, 8.0, 3)) + (a + a.T) * (i % 3 == 0)   # every third tensor is full
        sd[23:32] = inv.astype(F).ravel()
        rows.append(sd)
    sc = scenes._scene([("box", 1.0, 1.0, 1.0)], np.zeros(N, np.int32), rows, np.ones(N, np.int32), "friction_cases")
    st = np.zeros((N, 12), F)
    st[:, 0:3] = np.stack(rows)[:, 3:6]; st[:, 3:6] = np.stack(rows)[:, 6:9]
    st[:, 6:9] = rng.uniform(-6, 6, (N, 3)); st[:, 9:12] = rng.uniform(-2, 2, (N, 3))
    r = rng.uniform(-1, 1, (N, 3)).astype(F); depth = rng.uniform(0.0, 0.6, N).astype(F)
    n = rng.normal(size=(N, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True); n = n.astype(F)
    n[0::16] = 0.0                                                       # the world step's zero normal
    n[1::16] = [0.0, 1.0, 0.0]                                           # an axis
    k = np.arange(2, N, 16)                                              # velocity at the point (nearly) orthogonal to n: FloatEquals(relVelNormal, 0)
    st[k, 9:12] = 0.0; vv = st[k, 6:9].astype(np.float64); t = np.cross(vv, rng.normal(size=vv.shape)); n[k] = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(F)
    k = np.arange(3, N, 16)                                              # velocity along n: the tangential part is rounding noise
    st[k, 9:12] = 0.0; vv = st[k, 6:9].astype(np.float64); n[k] = (vv / np.linalg.norm(vv, axis=1, keepdims=True)).astype(F)
    k = np.arange(4, N, 32)                                              # a body at rest: 0 / x
    st[k, 6:12] = 0.0
    st[5, 6] = F(np.inf); st[21, 9] = F(np.nan); sc["sd"][37, 11] = F(np.nan); sc["sd"][53, 11] = F(-0.4)   # non-finite momenta, NaN and negative coefficient
    return sc, st, r, depth, n


def run_ref(friction, sc, st, r, depth, n):
    ref = Ref(friction=friction)
    ref.load_scene(sc); ref.set_state(st)
    for i in range(N):
        ref.apply_impulse_response(i, r[i], depth[i], n[i])
    return ref.state()


def run_steps(friction, steps=300):
    o = Oracle(); ref = Ref(friction=friction)
    sc = scenes.resolve(scenes.few_boxes_scene(), o)
    ref.load_scene(sc)
    for _ in range(steps):
        ref.step(sc["dt"])
    return ref.state()


def main():
    o = Oracle()
    sc, st, r, depth, n = friction_cases()
    sc = scenes.resolve(sc, o)
    out_f = run_ref(True, sc, st, r, depth, n)
    out_p = run_ref(False, sc, st, r, depth, n)
    changed = int(np.sum(np.any(out_f.view(np.uint32) != out_p.view(np.uint32), axis=1)))
    steps_f = run_steps(True); steps_p = run_steps(False)
    assert np.array_equal(steps_f.view(np.uint32), steps_p.view(np.uint32)), "the friction build changed a world step: normals are not zero?"
    path = os.path.join(ROOT, "tests", "golden", "friction.npz")
    np.savez_compressed(path, seed=7, sd=sc["sd"], state=st, r=r, depth=depth, n=n, out_friction=out_f, out_plain=out_p, few_boxes_300=steps_f)
    print("wrote %s: %d cases, friction changed %d of them; few-boxes 300 steps identical under both builds" % (path, N, changed))


if __name__ == "__main__":
    main()
