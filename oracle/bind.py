"""oracle/bind.py -- TEST INFRASTRUCTURE (ctypes bindings of the two CPU checkers).

* ``Oracle``  -> oracle/liboracle.so          (plain-C restatement, oracle/netphys_oracle.c)
* ``Ref``     -> oracle/_ref/libnetphys_ref.so (the reference's own sources, oracle/ref/ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(f32p)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(i32p)


def f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


class Counters(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "gjk_steps", "gjk_supports", "solve_line", "solve_tri", "solve_tet", "dir1", "dir2", "dir3",
        "epa_iters", "epa_supports", "epa_tri_tests", "epa_faces_added", "epa_max_faces", "epa_max_verts")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Contact(C.Structure):
    _fields_ = [("a_normal", C.c_float * 3), ("b_normal", C.c_float * 3), ("pen_dir", C.c_float * 3),
                ("depth", C.c_float), ("success", C.c_int)]


def build(force=False):
    """make -C oracle (liboracle.so always; _ref only where /root/reference exists)."""
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "netphys_oracle.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/engine"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class Oracle:
    def __init__(self):
        build()
        L = self.L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.npo_sinf.restype = C.c_float; L.npo_sinf.argtypes = [C.c_float]
        L.npo_cosf.restype = C.c_float; L.npo_cosf.argtypes = [C.c_float]
        L.npo_float_equals.argtypes = [C.c_float, C.c_float, C.c_float]
        L.npo_world_create.restype = C.c_void_p
        L.npo_world_destroy.argtypes = [C.c_void_p]
        L.npo_world_add_shape.argtypes = [C.c_void_p, f32p, C.c_int]
        L.npo_world_add_body.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        L.npo_world_count.argtypes = [C.c_void_p]
        L.npo_world_add_force.argtypes = [C.c_void_p, C.c_int, f32p]; L.npo_world_add_force.restype = None
        L.npo_world_set_enabled.argtypes = [C.c_void_p, C.c_int, C.c_int]; L.npo_world_set_enabled.restype = None
        L.npo_world_get_state.argtypes = [C.c_void_p, f32p]
        L.npo_world_set_friction.argtypes = [C.c_void_p, C.c_int]; L.npo_world_set_friction.restype = None
        L.npo_world_apply_impulse.argtypes = [C.c_void_p, C.c_int, f32p, C.c_float, f32p]; L.npo_world_apply_impulse.restype = None
        L.npo_world_set_state.argtypes = [C.c_void_p, f32p]
        L.npo_world_step.argtypes = [C.c_void_p, C.c_float, i32p, i32p, f32p, C.c_int]
        L.npo_world_counters.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(Counters)]
        L.npo_detect.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, f32p, C.c_int, C.POINTER(Contact), C.POINTER(Counters)]
        L.npo_detect_batch.argtypes = [C.c_int, f32p, i32p, i32p, i32p, f32p, i32p, f32p, i32p, f32p, C.POINTER(Counters), C.c_int]
        L.npo_make_box.argtypes = [C.c_float, C.c_float, C.c_float, f32p, C.c_int]
        L.npo_make_sphere.argtypes = [C.c_float, f32p, C.c_int]

    # -- scalars / solvers -------------------------------------------------------------------
    def sinf(self, x): return self.L.npo_sinf(float(np.float32(x)))
    def cosf(self, x): return self.L.npo_cosf(float(np.float32(x)))

    def closest_line(self, p, a, b):
        u = C.c_float()
        r = self.L.npo_closest_line(_f(f32(p)), _f(f32(a)), _f(f32(b)), C.byref(u))
        return r, np.float32(u.value)

    def closest_triangle(self, p, a, b, c):
        uv = np.zeros(2, np.float32)
        r = self.L.npo_closest_triangle(_f(f32(p)), _f(f32(a)), _f(f32(b)), _f(f32(c)), _f(uv))
        return r, uv

    def closest_tetra(self, p, a, b, c, d):
        uvw = np.zeros(3, np.float32)
        r = self.L.npo_closest_tetra(_f(f32(p)), _f(f32(a)), _f(f32(b)), _f(f32(c)), _f(f32(d)), _f(uvw))
        return r, uvw

    def dir_to_origin(self, pts):
        pts = f32(pts); out = np.zeros(3, np.float32)
        self.L.npo_dir_to_origin(len(pts), _f(pts), _f(out))
        return out

    # -- shapes / transforms -----------------------------------------------------------------
    def make_box(self, w, d, h):
        out = np.zeros((64, 3), np.float32)
        n = self.L.npo_make_box(w, d, h, _f(out), 64)
        return out[:n].copy()

    def make_sphere(self, r):
        out = np.zeros((1024, 3), np.float32)
        n = self.L.npo_make_sphere(r, _f(out), 1024)
        return out[:n].copy()

    def transform(self, pos, rot):
        out = np.zeros(16, np.float32)
        self.L.npo_transform(_f(f32(pos)), _f(f32(rot)), _f(out))
        return out

    def matrix3_inv(self, m):
        out = np.zeros(9, np.float32)
        self.L.npo_matrix3_inv(_f(f32(m).ravel()), _f(out))
        return out

    def support(self, verts, d, xf):
        out = np.zeros(3, np.float32); verts = f32(verts)
        self.L.npo_support(_f(verts), len(verts), _f(f32(d)), _f(f32(xf)), _f(out))
        return out

    def detect(self, va, xa, vb, xb, is3d=True, want_counters=False):
        va = f32(va); vb = f32(vb); c = Contact(); ctr = Counters()
        hit = self.L.npo_detect(_f(va), len(va), _f(f32(xa)), _f(vb), len(vb), _f(f32(xb)), int(is3d), C.byref(c), C.byref(ctr))
        out = np.array(list(c.pen_dir) + [c.depth, float(c.success)], np.float32)
        return (hit, out, ctr.as_dict()) if want_counters else (hit, out)

    def detect_batch(self, shapes, sa, posea, sb, poseb, nthreads=1):
        """shapes: list of (n,3) arrays. poses: (n,6) pos+rot. -> hit int32[n], out float32[n,5], counters dict"""
        xyz = f32(np.concatenate([f32(s).reshape(-1, 3) for s in shapes]))
        cnt = np.array([len(s) for s in shapes], np.int32)
        off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
        sa = np.ascontiguousarray(sa, np.int32); sb = np.ascontiguousarray(sb, np.int32)
        posea = f32(posea); poseb = f32(poseb)
        n = len(sa); hit = np.zeros(n, np.int32); out = np.zeros((n, 5), np.float32); ctr = Counters()
        self.L.npo_detect_batch(n, _f(xyz), _i(off), _i(cnt), _i(sa), _f(posea), _i(sb), _f(poseb), _i(hit), _f(out), C.byref(ctr), nthreads)
        return hit, out, ctr.as_dict()

    # -- world -------------------------------------------------------------------------------
    def world(self, scene):
        return OracleWorld(self, scene)


class OracleWorld:
    """scene: dict(shapes=[(n,3)...], body_shape int32[N], sd float32[N,32], response int32[N])"""

    def __init__(self, o, scene):
        self.o = o; L = o.L
        self.w = L.npo_world_create()
        for s in scene["shapes"]:
            s = f32(s); L.npo_world_add_shape(self.w, _f(s), len(s))
        sd = f32(scene["sd"])
        for i in range(len(sd)):
            L.npo_world_add_body(self.w, int(scene["body_shape"][i]), _f(sd[i]), int(scene["response"][i]))
        self.n = len(sd)

    def step(self, dt, want_collisions=False):
        L = self.o.L
        if not want_collisions:
            return L.npo_world_step(self.w, float(np.float32(dt)), None, None, None, 0)
        a = np.zeros(self.n + 1, np.int32); b = np.zeros(self.n + 1, np.int32); d = np.zeros((self.n + 1, 11), np.float32)
        nc = L.npo_world_step(self.w, float(np.float32(dt)), _i(a), _i(b), _f(d), self.n + 1)
        return a[:nc].copy(), b[:nc].copy(), d[:nc].copy()

    def state(self):
        out = np.zeros((self.n, 12), np.float32)
        self.o.L.npo_world_get_state(self.w, _f(out))
        return out

    def add_force(self, body, f):
        self.o.L.npo_world_add_force(self.w, int(body), _f(f32(f)))

    def set_enabled(self, body, enabled):
        self.o.L.npo_world_set_enabled(self.w, int(body), int(bool(enabled)))

    def set_state(self, st):
        self.o.L.npo_world_set_state(self.w, _f(f32(st)))

    def set_friction(self, on):
        self.o.L.npo_world_set_friction(self.w, int(bool(on)))

    def apply_impulse_response(self, body, pen_dir, depth, normal):
        self.o.L.npo_world_apply_impulse(self.w, int(body), _f(f32(pen_dir)), float(np.float32(depth)), _f(f32(normal)))

    def counters(self):
        pt = C.c_longlong(); h = C.c_longlong(); c = Counters()
        self.o.L.npo_world_counters(self.w, C.byref(pt), C.byref(h), C.byref(c))
        return pt.value, h.value, c.as_dict()

    def close(self):
        if self.w:
            self.o.L.npo_world_destroy(self.w); self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libnetphys_ref.so"))


class Ref:
    """The reference's own engine (process-global body list: one scene at a time)."""

    def __init__(self, friction=False):
        """friction=True: the variant compiled with the reference's friction block (physics.cpp:49-72) uncommented, see oracle/Makefile"""
        build()
        L = self.L = C.CDLL(os.path.join(HERE, "_ref", "libnetphys_ref_friction.so" if friction else "libnetphys_ref.so"))
        assert L.ref_has_friction() == int(bool(friction))
        L.ref_body_apply_impulse.argtypes = [C.c_int, f32p, C.c_float, f32p]; L.ref_body_apply_impulse.restype = None
        L.ref_shape_sphere.argtypes = [C.c_float]
        L.ref_shape_box.argtypes = [C.c_float] * 3
        L.ref_shape_mesh.argtypes = [f32p, C.c_int]
        L.ref_shape_verts.argtypes = [C.c_int, f32p]
        L.ref_body_create.argtypes = [C.c_int, f32p, C.c_int]
        L.ref_body_get_state.argtypes = [C.c_int, f32p]
        L.ref_body_set_state.argtypes = [C.c_int, f32p]
        L.ref_world_get_state.argtypes = [f32p]
        L.ref_body_transform.argtypes = [C.c_int, f32p]
        L.ref_matrix3_inv.argtypes = [f32p, f32p]
        L.ref_update.argtypes = [C.c_float]
        L.ref_update_n.argtypes = [C.c_float, C.c_int]
        L.ref_update_traced.argtypes = [C.c_float, i32p, i32p, f32p, C.c_int]
        L.ref_transform.argtypes = [f32p, f32p, f32p]
        L.ref_detect.argtypes = [C.c_int, f32p, C.c_int, f32p, C.c_int, f32p]
        L.ref_detect_batch.argtypes = [C.c_int, i32p, f32p, i32p, f32p, i32p, f32p]
        L.ref_support.argtypes = [C.c_int, f32p, f32p, f32p]
        L.ref_closest_line.argtypes = [f32p, f32p, f32p, f32p]
        L.ref_closest_triangle.argtypes = [f32p] * 5
        L.ref_closest_tetra.argtypes = [f32p] * 6
        L.ref_closest_line_point.argtypes = [f32p] * 4
        L.ref_closest_triangle_point.argtypes = [f32p] * 5
        L.ref_line_point_distance.argtypes = [f32p] * 3; L.ref_line_point_distance.restype = C.c_float
        L.ref_dir_to_origin.argtypes = [C.c_int, f32p, f32p]
        L.ref_float_equals.argtypes = [C.c_float] * 3
        L.ref_sinf.argtypes = [C.c_float]; L.ref_sinf.restype = C.c_float
        L.ref_cosf.argtypes = [C.c_float]; L.ref_cosf.restype = C.c_float
        L.ref_sincos_batch.argtypes = [f32p, C.c_long, f32p, f32p]
        L.ref_time_update.argtypes = [C.c_float, C.c_int]; L.ref_time_update.restype = C.c_double
        L.ref_time_detect.argtypes = [C.c_int, i32p, f32p, i32p, f32p, i32p, f32p, C.c_int]; L.ref_time_detect.restype = C.c_double
        self.n = 0

    def sphere(self, r): return self.L.ref_shape_sphere(r)
    def box(self, w, d, h): return self.L.ref_shape_box(w, d, h)

    def mesh(self, xyz):
        xyz = f32(xyz); return self.L.ref_shape_mesh(_f(xyz), len(xyz))

    def shape_verts(self, s):
        out = np.zeros((self.L.ref_shape_nverts(s), 3), np.float32)
        self.L.ref_shape_verts(s, _f(out)); return out

    def sincos(self, x):
        x = f32(x); s = np.zeros_like(x); c = np.zeros_like(x)
        self.L.ref_sincos_batch(_f(x), x.size, _f(s), _f(c)); return s, c

    def matrix3_inv(self, m):
        out = np.zeros(9, np.float32); self.L.ref_matrix3_inv(_f(f32(m).ravel()), _f(out)); return out

    def transform(self, pos, rot):
        out = np.zeros(16, np.float32); self.L.ref_transform(_f(f32(pos)), _f(f32(rot)), _f(out)); return out

    def support(self, s, d, xf):
        out = np.zeros(3, np.float32); self.L.ref_support(s, _f(f32(d)), _f(f32(xf)), _f(out)); return out

    def detect(self, sa, xa, sb, xb, is3d=True):
        out = np.zeros(5, np.float32)
        hit = self.L.ref_detect(sa, _f(f32(xa)), sb, _f(f32(xb)), int(is3d), _f(out))
        return hit, out

    def detect_batch(self, sa, posea, sb, poseb):
        sa = np.ascontiguousarray(sa, np.int32); sb = np.ascontiguousarray(sb, np.int32)
        n = len(sa); hit = np.zeros(n, np.int32); out = np.zeros((n, 5), np.float32)
        self.L.ref_detect_batch(n, _i(sa), _f(f32(posea)), _i(sb), _f(f32(poseb)), _i(hit), _f(out))
        return hit, out

    def time_detect(self, sa, posea, sb, poseb, nthreads):
        sa = np.ascontiguousarray(sa, np.int32); sb = np.ascontiguousarray(sb, np.int32)
        n = len(sa); hit = np.zeros(n, np.int32); out = np.zeros((n, 5), np.float32)
        t = self.L.ref_time_detect(n, _i(sa), _f(f32(posea)), _i(sb), _f(f32(poseb)), _i(hit), _f(out), nthreads)
        return t, hit, out

    def closest_line(self, p, a, b):
        u = np.zeros(1, np.float32); r = self.L.ref_closest_line(_f(f32(p)), _f(f32(a)), _f(f32(b)), _f(u)); return r, u[0]

    def closest_triangle(self, p, a, b, c):
        uv = np.zeros(2, np.float32); r = self.L.ref_closest_triangle(_f(f32(p)), _f(f32(a)), _f(f32(b)), _f(f32(c)), _f(uv)); return r, uv

    def closest_tetra(self, p, a, b, c, d):
        uvw = np.zeros(3, np.float32)
        r = self.L.ref_closest_tetra(_f(f32(p)), _f(f32(a)), _f(f32(b)), _f(f32(c)), _f(f32(d)), _f(uvw)); return r, uvw

    def dir_to_origin(self, pts):
        pts = f32(pts); out = np.zeros(3, np.float32); self.L.ref_dir_to_origin(len(pts), _f(pts), _f(out)); return out

    # -- world (process-global) --------------------------------------------------------------
    def load_scene(self, scene):
        """Clears the global body list and every shape, then loads `scene` (shapes as raw meshes)."""
        self.L.ref_world_clear(); self.L.ref_shapes_clear()
        ids = [self.mesh(s) for s in scene["shapes"]]
        sd = f32(scene["sd"])
        for i in range(len(sd)):
            self.L.ref_body_create(ids[int(scene["body_shape"][i])], _f(sd[i]), int(scene["response"][i]))
        self.n = len(sd)
        return ids

    def step(self, dt, want_collisions=False):
        if not want_collisions:
            self.L.ref_update(float(np.float32(dt))); return None
        a = np.zeros(self.n + 1, np.int32); b = np.zeros(self.n + 1, np.int32); d = np.zeros((self.n + 1, 11), np.float32)
        nc = self.L.ref_update_traced(float(np.float32(dt)), _i(a), _i(b), _f(d), self.n + 1)
        return a[:nc].copy(), b[:nc].copy(), d[:nc].copy()

    def apply_impulse_response(self, body, pen_dir, depth, normal):
        self.L.ref_body_apply_impulse(int(body), _f(f32(pen_dir)), float(np.float32(depth)), _f(f32(normal)))

    def time_steps(self, dt, steps):
        return self.L.ref_time_update(float(np.float32(dt)), steps)

    def state(self):
        out = np.zeros((self.n, 12), np.float32); self.L.ref_world_get_state(_f(out)); return out

    def set_state(self, st):
        st = f32(st)
        for i in range(self.n):
            self.L.ref_body_set_state(i, _f(st[i]))


class RefPacket:
    """The reference's own snapshot ring + wire serialiser (oracle/_ref/librefpacket.so = netphys_server/world_s.cpp compiled unchanged,
    see oracle/ref/ref_packet.cpp).  Process-global frame list, like the reference's."""

    def __init__(self):
        L = self.L = C.CDLL(os.path.join(HERE, "_ref", "librefpacket.so"))
        L.refpkt_set_objects.argtypes = [C.c_int, C.POINTER(C.c_uint32), f32p, f32p, i32p]
        L.refpkt_update.argtypes = [C.c_double]
        L.refpkt_fill.argtypes = [C.c_int, C.c_char_p, C.c_uint]
        assert L.refpkt_object_size() == 36

    def set_objects(self, uid, pos, quat, enabled):
        uid = np.ascontiguousarray(uid, np.uint32); en = np.ascontiguousarray(enabled, np.int32)
        self.L.refpkt_set_objects(len(uid), uid.ctypes.data_as(C.POINTER(C.c_uint32)), _f(f32(pos)), _f(f32(quat)), _i(en))

    def update(self, now_ms): self.L.refpkt_update(float(now_ms))
    def frame_count(self): return self.L.refpkt_frame_count()
    def reset_frames(self): self.L.refpkt_reset_frames()

    def fill(self, kind, cap=1 << 20):
        buf = C.create_string_buffer(cap)
        n = self.L.refpkt_fill(kind, buf, cap)
        return buf.raw[:n]


def refpacket_available():
    return os.path.exists(os.path.join(HERE, "_ref", "librefpacket.so"))
