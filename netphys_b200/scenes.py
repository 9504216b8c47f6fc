"""Synthetic scenes for the netphys hot path (numpy only; no CUDA, no oracle).

A *scene* is plain data every consumer (CUDA path, C oracle, compiled reference) reads the same bits
from::

    shape_specs : list of ("sphere", r) | ("box", w, d, h) | ("mesh", (n,3) float32)
    body_shape  : int32[N]   index into shape_specs
    sd          : float32[N,32]  StaticPhysicsData, reference engine/physics.h:36-51, packed as
                  gravity[3] initPos[3] initRot[3] mass elasticity muStatic muDynamic momentOfInertia
                  inertia[9] invInertia[9]
    response    : int32[N]   COLLISION_RESPONSE (0 none, 1 impulse), physics.h:13-17
    dt          : float32    (1000.f/30.f)/1000.f, reference engine/platform.h:19 + windows.cpp:390

``resolve(scene, factory)`` turns shape_specs into vertex arrays with the factory's own CreateSphere /
CreateBox (so each implementation builds its shapes itself and tests compare them).
"""
import numpy as np

F = np.float32
DT = F(F(1000.0) / F(30.0)) / F(1000.0)          # 0x3d088888


def matrix3_inv_f32(m):
    """matrix3::inv (reference engine/matrix.h:83-103) in float32 scalar arithmetic, same op order."""
    x1, x2, x3, y1, y2, y3, z1, z2, z3 = [F(v) for v in np.asarray(m, dtype=np.float32).ravel()]
    with np.errstate(all="ignore"):
        d = x1 * y2 * z3 - x1 * y3 * z2 + x2 * y3 * z1 - x2 * y1 * z3 + x3 * y1 * z2 - x3 * y2 * z1
        det2 = lambda a, b, c, e: a * e - b * c
        out = [det2(y2, y3, z2, z3) / d, det2(x3, x2, z3, z2) / d, det2(x2, x3, y2, y3) / d,
               det2(y3, y1, z3, z1) / d, det2(x1, x3, z1, z3) / d, det2(x3, x1, y3, y1) / d,
               det2(y1, y2, z1, z2) / d, det2(x2, x1, z2, z1) / d, det2(x1, x2, y1, y2) / d]
    return np.array(out, dtype=np.float32)


def static_data(gravity=(0, 0, 0), pos=(0, 0, 0), rot=(0, 0, 0), mass=0.0, elasticity=0.0, mu_s=0.0, mu_d=0.0,
                moment=None):
    """One StaticPhysicsData row.  moment=None keeps the default identity tensors (physics.h:49-50)."""
    sd = np.zeros(32, np.float32)
    sd[0:3] = gravity; sd[3:6] = pos; sd[6:9] = rot
    sd[9] = mass; sd[10] = elasticity; sd[11] = mu_s; sd[12] = mu_d
    if moment is None:
        sd[13] = 0.0
        sd[14:23] = np.eye(3, dtype=np.float32).ravel()
        sd[23:32] = np.eye(3, dtype=np.float32).ravel()
    else:
        moment = F(moment)
        sd[13] = moment
        inertia = np.zeros(9, np.float32); inertia[[0, 4, 8]] = moment
        sd[14:23] = inertia
        sd[23:32] = matrix3_inv_f32(inertia)
    return sd


def _scene(specs, body_shape, rows, response, name):
    return {"name": name, "shape_specs": specs, "body_shape": np.asarray(body_shape, np.int32),
            "sd": np.ascontiguousarray(np.stack(rows), dtype=np.float32), "response": np.asarray(response, np.int32),
            "dt": DT}


def test_physics_scene():
    """reference engine/test_physics.cpp:71-106: sphere r=5 dropped on a 50x50x2 board."""
    size = F(5.0); mass = F(10.0)
    moment = F(0.4) * mass * size * size
    a = static_data((0.0, -9.8, 0.0), (0.0, 20.0, 0.0), (0.0, 20.0, 0.0), mass, 0.4, 0.4, 0.0, moment)
    b = static_data((0, 0, 0), (0.0, -1.0, 0.0), (0, 0, 0), 0.0, 0.0)
    return _scene([("sphere", 5.0), ("box", 50.0, 50.0, 2.0)], [0, 1], [a, b], [1, 0], "test_physics")


def test_3d_scene():
    """reference engine/test_3d.cpp:410-449: the same sphere, no gravity, at (0,3,0) over a 0.1-thick board."""
    size = F(5.0); mass = F(10.0)
    moment = F(0.4) * mass * size * size
    a = static_data((0, 0, 0), (0.0, 3.0, 0.0), (0, 0, 0), mass, 0.4, 0.4, 0.0, moment)
    b = static_data((0, 0, 0), (0, 0, 0), (0, 0, 0), 0.0, 0.0)
    return _scene([("sphere", 5.0), ("box", 50.0, 50.0, 0.1)], [0, 1], [a, b], [1, 0], "test_3d")


def few_boxes_scene(n=4):
    """BASELINE configs[0]: a few unit boxes falling onto a ground slab."""
    rows, shp, resp = [], [], []
    for i in range(n):
        pos = (F(i % 2) * F(3.0), F(5.0) + F(2.0) * F(i), F(i // 2) * F(3.0))
        rot = (F(0.1) * F(i), F(0.2) * F(i), F(0.3) * F(i))
        rows.append(static_data((0.0, -9.8, 0.0), pos, rot, 1.0, 0.4, 0.4, 0.0, F(1.0) / F(6.0)))
        shp.append(0); resp.append(1)
    rows.append(static_data((0, 0, 0), (0.0, -1.0, 0.0), (0, 0, 0), 0.0, 0.0))
    shp.append(1); resp.append(0)
    return _scene([("box", 1.0, 1.0, 1.0), ("box", 50.0, 50.0, 2.0)], shp, rows, resp, "few_boxes")


def random_hull(rng, nverts=16, radius=0.5):
    """nverts points on a sphere of `radius` (a convex vertex-list shape, like MeshPhysicsShape's m_vertexPos)."""
    p = rng.uniform(-0.5, 0.5, size=(nverts, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    return (p * radius).astype(np.float32)


def lattice_scene(n, spacing=2.5, jitter=0.3, seed=42, mix=(0.5, 0.5, 0.0), gravity=(0.0, -9.8, 0.0), name=None):
    """N bodies on a jittered cubic lattice.  mix = fractions of unit box / sphere r=.5 (162 verts) /
    random 16-vertex hull; mix=(.5,.5,0) alternates box/sphere by index parity (BASELINE configs[1]).
    Rotations U(-1,1)^3, mass 1, elasticity .4, impulse response."""
    rng = np.random.RandomState(seed)
    side = int(np.ceil(n ** (1.0 / 3.0)))
    idx = np.arange(n)
    grid = np.stack([idx % side, (idx // side) % side, idx // (side * side)], axis=1).astype(np.float64)
    pos = (grid * spacing + jitter * rng.uniform(-1, 1, size=(n, 3))).astype(np.float32)
    rot = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    specs = [("box", 1.0, 1.0, 1.0), ("sphere", 0.5)]
    if mix[2] > 0:
        specs.append(("mesh", random_hull(rng)))
    if tuple(mix) == (0.5, 0.5, 0.0):
        shp = (idx % 2).astype(np.int32)
    else:
        shp = rng.choice(len(mix), size=n, p=np.asarray(mix) / np.sum(mix)).astype(np.int32)
        shp = np.minimum(shp, len(specs) - 1)
    moments = {0: F(1.0) / F(6.0), 1: F(0.4) * F(1.0) * F(0.5) * F(0.5), 2: F(0.1)}
    base = {k: static_data(gravity, (0, 0, 0), (0, 0, 0), 1.0, 0.4, 0.4, 0.0, m) for k, m in moments.items()}
    sd = np.stack([base[int(s)] for s in shp]).astype(np.float32)
    sd[:, 3:6] = pos; sd[:, 6:9] = rot
    return {"name": name or f"lattice{n}", "shape_specs": specs, "body_shape": shp, "sd": np.ascontiguousarray(sd),
            "response": np.ones(n, np.int32), "dt": DT}


def random_pairs(n, seed=1234, extent=3.0, mix=(0.5, 0.5, 0.0), shapes=None):
    """n random (shapeA, poseA, shapeB, poseB) narrow-phase queries: positions U(-extent,extent)^3,
    rotations U(-pi,pi)^3 (the survey's probe distribution).  Returns specs, sa, posea, sb, poseb."""
    rng = np.random.RandomState(seed)
    specs = list(shapes) if shapes is not None else [("box", 1.0, 1.0, 1.0), ("sphere", 0.5), ("mesh", random_hull(rng))]
    p = np.asarray(mix[:len(specs)], dtype=np.float64); p = p / p.sum()
    sa = rng.choice(len(specs), size=n, p=p).astype(np.int32)
    sb = rng.choice(len(specs), size=n, p=p).astype(np.int32)

    def poses():
        return np.concatenate([rng.uniform(-extent, extent, size=(n, 3)), rng.uniform(-np.pi, np.pi, size=(n, 3))], axis=1).astype(np.float32)
    return specs, sa, poses(), sb, poses()


def resolve(scene, factory):
    """factory: object with make_sphere(r) -> (n,3) float32 and make_box(w,d,h) -> (n,3) float32."""
    out = dict(scene)
    out["shapes"] = resolve_specs(scene["shape_specs"], factory)
    return out


def resolve_specs(specs, factory):
    shapes = []
    for s in specs:
        if s[0] == "sphere":
            shapes.append(np.ascontiguousarray(factory.make_sphere(float(s[1])), dtype=np.float32))
        elif s[0] == "box":
            shapes.append(np.ascontiguousarray(factory.make_box(float(s[1]), float(s[2]), float(s[3])), dtype=np.float32))
        else:
            shapes.append(np.ascontiguousarray(s[1], dtype=np.float32))
    return shapes
