"""oracle/gen_golden_steps.py -- TEST INFRASTRUCTURE.  Regenerates the single-step GJK/EPA fixture from the REFERENCE ITSELF.

tests/golden/step_trace_in.txt   seeded pair list (shapes + two 4x4 transforms, floats as hex words)
tests/golden/step_trace_ref.txt.gz  what the reference's DetectCollisionStep / FindIntersectionPoints / GetSearchDirection / Simplex
                                 (engine/physics.h:115-127, engine/simplex.h) produce when tests/cpp/step_trace_body.h drives them the way
                                 engine/test_3d.cpp:36-85 and test_2d.cpp:34-75 do: every intermediate simplex, face list, search
                                 direction and CollisionData.

Dev container only (needs oracle/_ref/libnetphys_ref.so).    python oracle/gen_golden_steps.py
"""
import ctypes as C
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netphys_b200 import scenes  # noqa: E402
from oracle.bind import Ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_CALLS = 40


def hexf(a):
    return " ".join("%08x" % x for x in np.ascontiguousarray(a, np.float32).ravel().view(np.uint32))


def shape_record(spec):
    if spec[0] == "box":
        return 0, np.array(spec[1:4], np.float32)
    if spec[0] == "sphere":
        return 1, np.array(spec[1:2], np.float32)
    return 2, np.asarray(spec[1], np.float32).ravel()


def write_pairs(path, r):
    lines = []
    # 3-D: the survey's probe distribution, tight enough that about half the pairs overlap, plus the test_3d.cpp scene
    specs, sa, pa, sb, pb = scenes.random_pairs(36, seed=77, extent=0.9, mix=(0.4, 0.4, 0.2))
    for i in range(len(sa)):
        ka, fa = shape_record(specs[sa[i]]); kb, fb = shape_record(specs[sb[i]])
        xa = r.transform(pa[i, 0:3], pa[i, 3:6]); xb = r.transform(pb[i, 0:3], pb[i, 3:6])
        lines.append("1 %d %d %d %d %s %s %s %s" % (ka, fa.size, kb, fb.size, hexf(fa), hexf(fb), hexf(xa), hexf(xb)))
    # engine/test_3d.cpp:410-449: a radius-5 sphere over a 50 x 50 x 0.1 board, moved up and down with the arrow keys (:86-112)
    zero = np.zeros(3, np.float32)
    for y in (3.0, 4.0, 5.0, 6.0, 9.0, -2.0):
        ka, fa = shape_record(("sphere", 5.0)); kb, fb = shape_record(("box", 50.0, 50.0, 0.1))
        xa = r.transform(np.array([0, y, 0], np.float32), zero); xb = r.transform(zero, zero)
        lines.append("1 %d %d %d %d %s %s %s %s" % (ka, fa.size, kb, fb.size, hexf(fa), hexf(fb), hexf(xa), hexf(xb)))
    # 2-D (test_2d.cpp): flat polygons in the z = 0 plane, rotations about z only
    rng = np.random.RandomState(78)
    tri = np.array([[0, 1, 0], [-1, -1, 0], [1, -1, 0]], np.float32)
    quad = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    hexa = np.stack([np.cos(np.arange(6) * np.pi / 3), np.sin(np.arange(6) * np.pi / 3), np.zeros(6)], axis=1).astype(np.float32)
    polys = [tri, quad, hexa]
    for i in range(16):
        a = polys[rng.randint(3)]; b = polys[rng.randint(3)]
        pos_a = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-1.2, 1.2), 0], np.float32); pos_b = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-1.2, 1.2), 0], np.float32)
        xa = r.transform(pos_a, np.array([0, 0, rng.uniform(-np.pi, np.pi)], np.float32)); xb = r.transform(pos_b, np.array([0, 0, rng.uniform(-np.pi, np.pi)], np.float32))
        lines.append("0 2 %d 2 %d %s %s %s %s" % (a.size, b.size, hexf(a), hexf(b), hexf(xa), hexf(xb)))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(lines)


def main():
    r = Ref()
    inp = os.path.join(GOLD, "step_trace_in.txt"); tmp = os.path.join(GOLD, "step_trace_ref.txt")
    n = write_pairs(inp, r)
    r.L.ref_step_trace.restype = C.c_int
    got = r.L.ref_step_trace(inp.encode(), tmp.encode(), MAX_CALLS)
    assert got == n, (got, n)
    with open(tmp, "rb") as f, gzip.GzipFile(tmp + ".gz", "wb", mtime=0) as g:
        g.write(f.read())
    os.remove(tmp)
    print("step trace: %d pairs -> %s.gz (%d bytes)" % (n, tmp, os.path.getsize(tmp + ".gz")))


if __name__ == "__main__":
    main()
