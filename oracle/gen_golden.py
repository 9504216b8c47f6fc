"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/*.npz from the REFERENCE ITSELF.

Runs only in the dev container: it needs oracle/_ref/libnetphys_ref.so, i.e. the reference's own engine
sources compiled by oracle/Makefile from /root/reference.  The fixtures are what travels to the GPU box,
where /root/reference does not exist.  Inputs are seeded (netphys_b200/scenes.py); shapes are passed to
the reference as raw vertex lists built by the reference's own CreateSphere/CreateBox.

    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netphys_b200 import scenes  # noqa: E402
from oracle.bind import Ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class RefFactory:
    """CreateSphere/CreateBox of the reference (engine/physics_shape.cpp:313-320)."""

    def __init__(self, r): self.r = r
    def make_sphere(self, rad): return self.r.shape_verts(self.r.sphere(rad))
    def make_box(self, w, d, h): return self.r.shape_verts(self.r.box(w, d, h))


def small_scenes():
    return [(scenes.test_physics_scene(), 1000), (scenes.test_3d_scene(), 100), (scenes.few_boxes_scene(), 1000),
            (scenes.lattice_scene(64, spacing=1.2, seed=11, mix=(0.4, 0.4, 0.2), name="lattice64_mixed"), 100),
            (scenes.lattice_scene(256, spacing=2.5, seed=42, name="lattice256_c2"), 30)]


def main():
    r = Ref(); fac = RefFactory(r)
    os.makedirs(GOLD, exist_ok=True)

    # shapes
    np.savez_compressed(os.path.join(GOLD, "shapes.npz"), sphere5=fac.make_sphere(5.0), sphere05=fac.make_sphere(0.5),
                        box_50_50_2=fac.make_box(50.0, 50.0, 2.0), box_1=fac.make_box(1.0, 1.0, 1.0), box_50_50_01=fac.make_box(50.0, 50.0, 0.1),
                        box_2_3_4=fac.make_box(2.0, 3.0, 4.0))

    # util.h solvers on random (and deliberately degenerate) inputs
    rng = np.random.RandomState(2024)
    n = 4000
    pts = rng.uniform(-3, 3, size=(n, 4, 3)).astype(np.float32)
    pts[::7] = np.round(pts[::7])                       # lattice points: exact zeros, ties, collinear cases
    pts[5::11, 2] = pts[5::11, 1]                       # repeated vertices
    zero = np.zeros(3, np.float32)
    line = [r.closest_line(zero, p[0], p[1]) for p in pts]
    tri = [r.closest_triangle(zero, p[0], p[1], p[2]) for p in pts]
    tet = [r.closest_tetra(zero, p[0], p[1], p[2], p[3]) for p in pts]
    d1 = np.stack([r.dir_to_origin(p[:1]) for p in pts]); d2 = np.stack([r.dir_to_origin(p[:2]) for p in pts]); d3 = np.stack([r.dir_to_origin(p[:3]) for p in pts])
    np.savez_compressed(os.path.join(GOLD, "solvers.npz"), pts=pts,
                        line_region=np.array([x[0] for x in line], np.int32), line_u=np.array([x[1] for x in line], np.float32),
                        tri_region=np.array([x[0] for x in tri], np.int32), tri_uv=np.stack([x[1] for x in tri]),
                        tet_region=np.array([x[0] for x in tet], np.int32), tet_uvw=np.stack([x[1] for x in tet]), dir1=d1, dir2=d2, dir3=d3)

    # sin/cos of the reference's libm on a spread of arguments (the exhaustive check is oracle/sincos_exhaustive.py)
    x = np.concatenate([rng.uniform(-130, 130, 20000), rng.uniform(-1e-2, 1e-2, 2000), rng.uniform(-1e6, 1e6, 5000), 10.0 ** rng.uniform(-40, 38, 3000),
                        [0.0, -0.0, 20.0, 0.78539816, 0.7853982, 120.0, 119.99999, 1e-45, 3.4e38, -3.4e38]]).astype(np.float32)
    s, c = r.sincos(x)
    np.savez_compressed(os.path.join(GOLD, "sincos_ref.npz"), x=x, sin=s, cos=c)

    # DetectCollision on random pairs (box / icosphere / random hull)
    specs, sa, pa, sb, pb = scenes.random_pairs(20000, seed=1234, mix=(0.4, 0.4, 0.2))
    shapes = scenes.resolve_specs(specs, fac)
    rid = np.array([r.mesh(s) for s in shapes], np.int32)
    hit, out = r.detect_batch(rid[sa], pa, rid[sb], pb)
    # transforms + support of a few
    xf = np.stack([r.transform(p[:3], p[3:]) for p in pa[:2000]])
    sup = np.stack([r.support(int(rid[sa[i]]), pb[i, :3], xf[i]) for i in range(2000)])
    np.savez_compressed(os.path.join(GOLD, "pairs_ref.npz"), hull=shapes[2], sa=sa, pa=pa, sb=sb, pb=pb, hit=hit, out=out, xf=xf, support=sup)

    # Physics_Update on the small scenes: per-step state + collision lists
    for sc, steps in small_scenes():
        s = scenes.resolve(sc, fac)
        r.load_scene(s)
        states = np.zeros((steps, len(sc["sd"]), 12), np.float32)
        ca, cb, cd, cn = [], [], [], []
        for k in range(steps):
            a, b, d = r.step(sc["dt"], True)
            states[k] = r.state(); ca.append(a); cb.append(b); cd.append(d); cn.append(len(a))
        keep = sorted(set(list(range(0, min(steps, 40))) + list(range(0, steps, 25)) + [steps - 1]))
        np.savez_compressed(os.path.join(GOLD, "scene_%s.npz" % sc["name"]), steps=np.int32(steps), kept_steps=np.array(keep, np.int32),
                            states=states[keep], final_state=states[-1], col_count=np.array(cn, np.int32),
                            col_a=np.concatenate(ca).astype(np.int32), col_b=np.concatenate(cb).astype(np.int32),
                            col_data=np.concatenate(cd).astype(np.float32) if sum(cn) else np.zeros((0, 11), np.float32))
        print(sc["name"], "steps", steps, "collisions", sum(cn))
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden fixtures: %.1f KiB" % (tot / 1024))


if __name__ == "__main__":
    main()
