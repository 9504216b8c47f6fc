"""GPU suite: the ODE-shaped facade (include/netphys/ode_shim/ode/ode.h).  tests/shim_world.cpp drives the d* API the
way netphys_common/world.cpp, worldobject.cpp, player.cpp and netphys_server/world_s.cpp do; every frame must equal the oracle
stepping the same bodies with the same forces (the force / enable semantics are this project's extension, mirrored in the oracle)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32
N, BOX, DENSITY, GRAVITY, ACCEL, PLAYER = 5, F(0.5), F(5.0), F(9.8), F(20.0), 2.0


def _sd(gravity, pos, mass, idiag):
    """netphys_ode::static_data: float32 cofactor inverse of the (diagonal) inertia tensor, same operation order."""
    sd = np.zeros(32, np.float32)
    sd[3:6] = pos
    if mass is None:
        return sd, 0
    sd[0:3] = gravity; sd[9] = F(mass)
    I = np.zeros(9, np.float32); I[[0, 4, 8]] = [F(v) for v in idiag]
    det = F(F(I[0] * F(F(I[4] * I[8]) - F(I[5] * I[7]))) - F(I[1] * F(F(I[3] * I[8]) - F(I[5] * I[6])))) + F(I[2] * F(F(I[3] * I[7]) - F(I[4] * I[6])))
    idet = F(1.0) / F(det)
    inv = np.zeros(9, np.float32)
    inv[0] = F(F(I[4] * I[8]) - F(I[5] * I[7])) * idet; inv[4] = F(F(I[0] * I[8]) - F(I[2] * I[6])) * idet; inv[8] = F(F(I[0] * I[4]) - F(I[1] * I[3])) * idet
    sd[13] = I[0]; sd[14:23] = I; sd[23:32] = inv
    return sd, 1


class _Mirror:
    """The oracle world the shim is expected to have built: ground slab first, then bodies in dBodyCreate order."""

    def __init__(self, oracle):
        self.o = oracle; self.L = oracle.L
        self.w = self.L.npo_world_create(); self.n = 0
        slab = np.array([[(100.0 if k & 1 else -100.0), (100.0 if k & 2 else -100.0), (0.0 if k & 4 else -1.0)] for k in range(8)], np.float32)
        h = float(BOX) * 0.5
        box = np.array([[(h if k & 1 else -h), (h if k & 2 else -h), (h if k & 4 else -h)] for k in range(8)], np.float32)
        self.shape = {}
        for name, v in (("slab", slab), ("box", box), ("sphere", oracle.make_sphere(PLAYER))):
            v = np.ascontiguousarray(v, np.float32)
            self.shape[name] = self.L.npo_world_add_shape(self.w, v.ctypes.data_as(self.o_f32p()), len(v))
        self.add("slab", (0, 0, 0), None, None)

    def o_f32p(self):
        import ctypes as C
        return C.POINTER(C.c_float)

    def add(self, shape, pos, mass, idiag):
        sd, resp = _sd((0, 0, -GRAVITY), pos, mass, idiag)
        self.L.npo_world_add_body(self.w, self.shape[shape], sd.ctypes.data_as(self.o_f32p()), resp)
        self.n += 1
        return self.n - 1

    def force(self, body, f):
        f = np.ascontiguousarray(f, np.float32); self.L.npo_world_add_force(self.w, body, f.ctypes.data_as(self.o_f32p()))

    def step(self, dt):
        self.L.npo_world_step(self.w, float(F(dt)), None, None, None, 0)
        st = np.zeros((self.n, 12), np.float32); self.L.npo_world_get_state(self.w, st.ctypes.data_as(self.o_f32p()))
        return st


def _boxes(m):
    mass = float(DENSITY) * float(BOX) ** 3; i = mass / 12.0 * (float(BOX) ** 2 * 2)
    return [m.add("box", (float(a - N // 2), float(b - N // 2), float(BOX)), mass, (i, i, i)) for a in range(N) for b in range(N)]


def test_ode_call_sites_step_on_the_device_like_the_oracle(oracle, npb):
    exe = os.path.join(ROOT, "tests", "shim_world.bin"); lib = os.path.join(ROOT, "netphys_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include", "netphys", "ode_shim"), os.path.join(ROOT, "tests", "shim_world.cpp"),
                        "-o", exe, "-L", lib, "-lnetphys_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frames = 26
    out = subprocess.run([exe, str(frames)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = {}
    for line in out.stdout.strip().splitlines():
        f = line.split(); got[(int(f[0]), int(f[1]))] = (np.array([int(x, 16) for x in f[2:9]], np.uint32), int(f[9]))
    m = _Mirror(oracle); objs = _boxes(m); player = None; disabled = set()
    pm = (4.0 / 3.0) * 3.14159265358979323846 * PLAYER ** 3 * float(DENSITY)
    for f in range(frames):
        if f == 4:
            player = m.add("sphere", (0.0, 0.0, 5.0), pm, (0.4 * pm * PLAYER * PLAYER,) * 3); objs.append(player)
        if player is not None and f % 3 == 0:
            a = [float(ACCEL), float(ACCEL), float(ACCEL) if f % 2 else 0.0]
            m.force(player, [F(x * pm) for x in a])
        if f == 10: m.L.npo_world_set_enabled(m.w, objs[3], 0); disabled.add(objs[3])
        if f == 16: m.L.npo_world_set_enabled(m.w, objs[3], 1); disabled.discard(objs[3])
        if f == 20:
            m = _Mirror(oracle); objs = _boxes(m); player = None; disabled = set()
        st = m.step(0.01)
        assert len([k for k in got if k[0] == f]) == len(objs), "frame %d: object count" % f
        for k, body in enumerate(objs):
            bits, en = got[(f, k)]
            assert np.array_equal(bits[0:3], st[body, 0:3].view(np.uint32)), "frame %d object %d position" % (f, k)
            h = st[body, 3:6].astype(np.float64) * 0.5; (sx, sy, sz), (cx, cy, cz) = np.sin(h), np.cos(h)
            q = np.array([cx * cy * cz + sx * sy * sz, sx * cy * cz - cx * sy * sz, cx * sy * cz + sx * cy * sz, cx * cy * sz - sx * sy * cz])
            assert np.allclose(bits[3:7].view(np.float32), q, atol=2e-6), "frame %d object %d quaternion" % (f, k)
            assert en == (0 if body in disabled else 1)
    moved = got[(19, 25)][0][0:3].view(np.float32)
    assert abs(moved[0]) > 0.01 and abs(moved[1]) > 0.01, "the player's input forces had no effect"
