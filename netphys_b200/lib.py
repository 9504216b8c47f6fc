"""ctypes binding of include/netphys_b200.h (the C ABI).  Fails loudly when the library is not built."""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class NetphysError(RuntimeError):
    pass


class StaticPhysicsData(C.Structure):        # np_static_physics_data == engine/physics.h:36-51
    _fields_ = [("gravity", C.c_float * 3), ("initial_position", C.c_float * 3), ("initial_rotation", C.c_float * 3),
                ("mass", C.c_float), ("elasticity", C.c_float), ("static_friction_coeff", C.c_float), ("dynamic_friction_coeff", C.c_float),
                ("collision_response", C.c_int32), ("moment_of_inertia", C.c_float),
                ("inertia_tensor", C.c_float * 9), ("inverse_inertia_tensor", C.c_float * 9)]


class CollisionData(C.Structure):            # np_collision_data == engine/physics.h:18-25
    _fields_ = [("a_normal", C.c_float * 3), ("b_normal", C.c_float * 3), ("penetration_direction", C.c_float * 3),
                ("depth", C.c_float), ("success", C.c_int32)]


class Collision(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("data", CollisionData)]


class BodyState(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 3), ("linear_momentum", C.c_float * 3), ("angular_momentum", C.c_float * 3)]


class SnapshotObject(C.Structure):
    _fields_ = [("guid", C.c_uint32), ("pos", C.c_float * 3), ("rot", C.c_float * 4), ("is_enabled", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("steps", "pair_tests", "hits", "gjk_steps", "support_pairs", "epa_iters", "epa_tri_tests", "epa_overflow", "candidates", "support_verts", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


SD_DTYPE = np.dtype([("gravity", "<f4", 3), ("initial_position", "<f4", 3), ("initial_rotation", "<f4", 3), ("mass", "<f4"), ("elasticity", "<f4"),
                     ("static_friction_coeff", "<f4"), ("dynamic_friction_coeff", "<f4"), ("collision_response", "<i4"), ("moment_of_inertia", "<f4"),
                     ("inertia_tensor", "<f4", 9), ("inverse_inertia_tensor", "<f4", 9)])
COLLISION_DATA_DTYPE = np.dtype([("a_normal", "<f4", 3), ("b_normal", "<f4", 3), ("penetration_direction", "<f4", 3), ("depth", "<f4"), ("success", "<i4")])
COLLISION_DTYPE = np.dtype([("a", "<i4"), ("b", "<i4"), ("data", COLLISION_DATA_DTYPE)])
SNAPSHOT_DTYPE = np.dtype([("guid", "<u4"), ("pos", "<f4", 3), ("rot", "<f4", 4), ("is_enabled", "<u4")])
assert SD_DTYPE.itemsize == C.sizeof(StaticPhysicsData) and COLLISION_DTYPE.itemsize == C.sizeof(Collision) and SNAPSHOT_DTYPE.itemsize == 36


def _declared_symbols():
    """Every function include/netphys_b200.h declares (the boundary the library must export)."""
    text = open(os.path.join(ROOT, "include", "netphys_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(np_[a-z0-9_]+)\s*\(", text)))


DECLARED_SYMBOLS = _declared_symbols()
_LIBS = {}


def library_path(variant="exact"):
    """exact: the parity build (-fmad=false, bit-identical to the reference).  fma: the same sources with FMA contraction
    (-fmad=true), whose poses stay within 1e-5 relative of the reference (tests/test_gpu_fma.py)."""
    if variant == "fma":
        return os.path.join(HERE, "libnetphys_b200_fma.so")
    return os.environ.get("NETPHYS_B200_LIB", os.path.join(HERE, "libnetphys_b200.so"))   # override = tuning builds only


def load_library(variant="exact"):
    path = library_path(variant)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise NetphysError("%s is not built -- run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                           "netphys_b200 has no CPU fallback." % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.np_last_error.restype = C.c_char_p
    L.np_world_create.restype = vp; L.np_world_create.argtypes = [C.c_int]
    L.np_world_destroy.restype = None; L.np_world_destroy.argtypes = [vp]
    L.np_world_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.np_world_begin_island.argtypes = [vp]
    L.np_world_body_count.argtypes = [vp]
    L.np_world_synchronize.argtypes = [vp]
    L.np_shape_create_mesh.argtypes = [vp, f32p, C.c_int]
    L.np_shape_create_sphere.argtypes = [vp, C.c_float]
    L.np_shape_create_box.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.np_shape_vertex_count.argtypes = [vp, C.c_int]
    L.np_shape_get_vertices.argtypes = [vp, C.c_int, f32p, C.c_int]
    L.np_shape_support.argtypes = [vp, C.c_int, f32p, f32p, f32p]
    L.np_body_create.argtypes = [vp, C.c_int, C.POINTER(StaticPhysicsData)]
    L.np_body_create_batch.argtypes = [vp, C.c_int, i32p, vp]
    L.np_body_reset.argtypes = [vp, C.c_int]
    L.np_body_destroy.argtypes = [vp, C.c_int]; L.np_shape_destroy.argtypes = [vp, C.c_int]
    L.np_body_get_state.argtypes = [vp, C.c_int, C.POINTER(BodyState)]
    L.np_body_set_state.argtypes = [vp, C.c_int, C.POINTER(BodyState)]
    L.np_body_get_transform.argtypes = [vp, C.c_int, f32p]
    L.np_body_apply_impulse_response.argtypes = [vp, C.c_int, C.POINTER(CollisionData), f32p]
    L.np_world_get_states.argtypes = [vp, vp, C.c_int]
    L.np_world_set_states.argtypes = [vp, vp, C.c_int]
    L.np_world_step.argtypes = [vp, C.c_float]
    L.np_world_step_n.argtypes = [vp, C.c_float, C.c_int]
    L.np_body_add_force.argtypes = [vp, C.c_int, vp]; L.np_body_set_enabled.argtypes = [vp, C.c_int, C.c_int]
    L.np_body_is_enabled.argtypes = [vp, C.c_int]; L.np_world_clear_bodies.argtypes = [vp]
    L.np_world_set_shard.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.np_world_step_begin.argtypes = [vp, C.c_float]
    L.np_world_step_end.argtypes = [vp]
    L.np_world_integrate.argtypes = [vp, C.c_float]
    L.np_world_get_collisions.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.np_world_step_host.argtypes = [vp, C.c_float, vp, vp, C.c_int]
    L.np_world_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.np_world_last_step_ms.argtypes = [vp, f32p]
    L.np_detect_collision.argtypes = [vp, C.c_int, f32p, C.c_int, f32p, C.c_int, C.POINTER(CollisionData), C.POINTER(C.c_int)]
    L.np_detect_collision_batch.argtypes = [vp, C.c_int, i32p, f32p, i32p, f32p, i32p, vp]
    L.np_detect_collision_batch_2d.argtypes = [vp, C.c_int, i32p, f32p, i32p, f32p, i32p, vp]
    L.np_detect_collision_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, f32p]
    L.np_world_snapshot.argtypes = [vp, vp, C.c_int]
    L.np_world_snapshot_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.np_device_alloc.restype = vp; L.np_device_alloc.argtypes = [vp, C.c_size_t]
    L.np_device_free.argtypes = [vp, vp]
    L.np_device_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.np_device_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.np_host_alloc.restype = vp; L.np_host_alloc.argtypes = [vp, C.c_size_t]
    L.np_host_free.argtypes = [vp, vp]
    L.np_world_stream.restype = vp; L.np_world_stream.argtypes = [vp]
    L.np_device_flush_l2.argtypes = [vp]
    L.np_world_record_frame.argtypes = [vp, C.c_double]
    L.np_world_frame_count.argtypes = [vp]
    L.np_world_fill_packet.argtypes = [vp, C.c_int, vp, C.c_int]
    L.np_world_set_snapshot_target.argtypes = [vp, vp, C.c_uint32]
    L.np_comm_get_unique_id.argtypes = [vp]
    L.np_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.np_comm_destroy.argtypes = [vp]
    L.np_world_set_snapshot_gather.argtypes = [vp, C.c_int, C.c_uint32]
    L.np_world_gathered_snapshot.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.np_world_last_gather_ms.argtypes = [vp, C.c_int, f32p]
    L.np_world_shard_over_comm.argtypes = [vp]
    L.np_world_step_sharded.argtypes = [vp, C.c_float]
    L.np_world_set_shard_halo.argtypes = [vp, C.c_int]
    L.np_world_step_integrate.argtypes = [vp, C.c_float]
    L.np_world_step_search.argtypes = [vp]
    L.np_world_transforms_device.argtypes = [vp, C.POINTER(vp)]
    L.np_selftest_fp32_peak.argtypes = [vp, f32p]
    L.np_selftest_fp32_peaks.argtypes = [vp, f32p]
    L.np_selftest_sincos_hash.argtypes = [vp, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.np_selftest_closest.argtypes = [vp, C.c_int, f32p, i32p, f32p]
    L.np_selftest_support_table.argtypes = [vp, C.c_int, C.c_int, f32p, f32p, f32p, f32p, i32p]
    L.np_shape_support_table_info.argtypes = [vp, C.c_int, f32p]
    _LIBS[path] = L
    return L


def device_count():
    return load_library().np_device_count()


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(f32p)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(i32p)


def sd_rows_to_struct(sd, response):
    """float32[N,32] scene rows (netphys_b200.scenes) -> np_static_physics_data[N] as a structured array."""
    sd = np.ascontiguousarray(sd, np.float32); n = len(sd)
    out = np.zeros(n, SD_DTYPE)
    out["gravity"] = sd[:, 0:3]; out["initial_position"] = sd[:, 3:6]; out["initial_rotation"] = sd[:, 6:9]
    out["mass"] = sd[:, 9]; out["elasticity"] = sd[:, 10]; out["static_friction_coeff"] = sd[:, 11]; out["dynamic_friction_coeff"] = sd[:, 12]
    out["collision_response"] = np.asarray(response, np.int32); out["moment_of_inertia"] = sd[:, 13]
    out["inertia_tensor"] = sd[:, 14:23]; out["inverse_inertia_tensor"] = sd[:, 23:32]
    return out


class World:
    """np_world.  Mirrors the reference's implicit global world (engine/physics.cpp:9) as an explicit object."""

    EXACT, GRID, SWEPT = 0, 1, 2
    OPT_BROADPHASE, OPT_GRID_CELL, OPT_COUNTERS, OPT_SNAPSHOT_EACH_STEP, OPT_PHASE_TIMING, OPT_SUPPORT_TABLES, OPT_FRICTION = 1, 2, 3, 4, 5, 6, 7

    def __init__(self, device=-1, variant="exact"):
        self.L = load_library(variant)
        self.w = self.L.np_world_create(device)
        if not self.w:
            raise NetphysError("np_world_create failed: %s" % self.L.np_last_error().decode())

    def _ck(self, r):
        if r < 0:
            raise NetphysError("netphys_b200 error %d: %s" % (r, self.L.np_last_error().decode()))
        return r

    def close(self):
        if getattr(self, "w", None):
            self.L.np_world_destroy(self.w); self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- shapes ----------------------------------------------------------------------------------
    def create_mesh(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        return self._ck(self.L.np_shape_create_mesh(self.w, _f(xyz), len(xyz)))

    def create_sphere(self, r): return self._ck(self.L.np_shape_create_sphere(self.w, r))
    def create_box(self, w, d, h): return self._ck(self.L.np_shape_create_box(self.w, w, d, h))

    def shape_vertices(self, s):
        n = self._ck(self.L.np_shape_vertex_count(self.w, s))
        out = np.zeros((n, 3), np.float32)
        self._ck(self.L.np_shape_get_vertices(self.w, s, _f(out), n))
        return out

    # factory protocol of netphys_b200.scenes.resolve
    def make_sphere(self, r): return self.shape_vertices(self.create_sphere(r))
    def make_box(self, w, d, h): return self.shape_vertices(self.create_box(w, d, h))

    def support(self, s, d, xf):
        out = np.zeros(3, np.float32)
        self._ck(self.L.np_shape_support(self.w, s, _f(np.ascontiguousarray(d, np.float32)), _f(np.ascontiguousarray(xf, np.float32)), _f(out)))
        return out

    # -- bodies ----------------------------------------------------------------------------------
    def create_bodies(self, shapes, sd_struct):
        shapes = np.ascontiguousarray(shapes, np.int32); sd_struct = np.ascontiguousarray(sd_struct)
        assert sd_struct.dtype == SD_DTYPE and len(shapes) == len(sd_struct)
        return self._ck(self.L.np_body_create_batch(self.w, len(shapes), _i(shapes), sd_struct.ctypes.data))

    def begin_island(self): return self._ck(self.L.np_world_begin_island(self.w))

    def load_scene(self, scene, builders=True):
        """scene from netphys_b200.scenes; shapes are built by THIS library's CreateSphere/CreateBox."""
        ids = []
        for s in scene["shape_specs"]:
            if s[0] == "sphere": ids.append(self.create_sphere(float(s[1])))
            elif s[0] == "box": ids.append(self.create_box(float(s[1]), float(s[2]), float(s[3])))
            else: ids.append(self.create_mesh(s[1]))
        ids = np.asarray(ids, np.int32)
        self.create_bodies(ids[scene["body_shape"]], sd_rows_to_struct(scene["sd"], scene["response"]))
        return ids

    @property
    def n(self): return self.L.np_world_body_count(self.w)

    def states(self):
        out = np.zeros((self.n, 12), np.float32)
        self._ck(self.L.np_world_get_states(self.w, out.ctypes.data, self.n))
        return out

    def set_states(self, st):
        st = np.ascontiguousarray(st, np.float32); assert st.shape == (self.n, 12)
        self._ck(self.L.np_world_set_states(self.w, st.ctypes.data, self.n))

    def body_state(self, b):
        s = BodyState(); self._ck(self.L.np_body_get_state(self.w, b, C.byref(s)))
        return np.frombuffer(bytes(s), np.float32).copy()

    def set_body_state(self, b, st):
        s = BodyState.from_buffer_copy(np.ascontiguousarray(st, np.float32).tobytes())
        self._ck(self.L.np_body_set_state(self.w, b, C.byref(s)))

    def reset_body(self, b): self._ck(self.L.np_body_reset(self.w, b))
    def destroy_body(self, b): self._ck(self.L.np_body_destroy(self.w, b))
    def destroy_shape(self, s): self._ck(self.L.np_shape_destroy(self.w, s))

    def transform(self, b):
        out = np.zeros(16, np.float32); self._ck(self.L.np_body_get_transform(self.w, b, _f(out))); return out

    def apply_impulse_response(self, b, pen_dir, depth, normal):
        cd = CollisionData(); cd.penetration_direction[:] = [float(x) for x in pen_dir]; cd.depth = float(depth)
        self._ck(self.L.np_body_apply_impulse_response(self.w, b, C.byref(cd), _f(np.ascontiguousarray(normal, np.float32))))

    # -- step ------------------------------------------------------------------------------------
    def set_option(self, opt, value): self._ck(self.L.np_world_set_option(self.w, opt, int(value)))
    def step(self, dt): self._ck(self.L.np_world_step(self.w, float(np.float32(dt))))
    def step_n(self, dt, steps): self._ck(self.L.np_world_step_n(self.w, float(np.float32(dt)), steps))

    def set_shard(self, first, count, d_hit_j=0, d_flags=0, d_contact=0):
        """Search only bodies [first, first + count) (count < 0: whole world); records go to the given full-length device arrays."""
        self._ck(self.L.np_world_set_shard(self.w, first, count, d_hit_j or None, d_flags or None, d_contact or None))

    def add_force(self, body, f):
        f = np.ascontiguousarray(f, np.float32); self._ck(self.L.np_body_add_force(self.w, body, f.ctypes.data))

    def set_enabled(self, body, enabled): self._ck(self.L.np_body_set_enabled(self.w, body, int(bool(enabled))))

    def is_enabled(self, body): return bool(self._ck(self.L.np_body_is_enabled(self.w, body)))

    def clear_bodies(self): self._ck(self.L.np_world_clear_bodies(self.w))

    def step_begin(self, dt): self._ck(self.L.np_world_step_begin(self.w, float(np.float32(dt))))

    def step_end(self): self._ck(self.L.np_world_step_end(self.w))
    def set_shard_halo(self, halo): self._ck(self.L.np_world_set_shard_halo(self.w, int(halo)))
    def step_integrate(self, dt): self._ck(self.L.np_world_step_integrate(self.w, float(np.float32(dt))))
    def step_search(self): self._ck(self.L.np_world_step_search(self.w))

    def transforms_device(self):
        p = C.c_void_p(); self._ck(self.L.np_world_transforms_device(self.w, C.byref(p))); return p.value
    def integrate(self, dt): self._ck(self.L.np_world_integrate(self.w, float(np.float32(dt))))
    def synchronize(self): self._ck(self.L.np_world_synchronize(self.w))

    def collisions(self):
        out = np.zeros(max(self.n, 1), COLLISION_DTYPE); cnt = C.c_int()
        self._ck(self.L.np_world_get_collisions(self.w, out.ctypes.data, len(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    def collisions_as_arrays(self):
        c = self.collisions(); d = np.zeros((len(c), 11), np.float32)
        d[:, 0:3] = c["data"]["a_normal"]; d[:, 3:6] = c["data"]["b_normal"]; d[:, 6:9] = c["data"]["penetration_direction"]
        d[:, 9] = c["data"]["depth"]; d[:, 10] = c["data"]["success"]
        return c["a"].copy(), c["b"].copy(), d

    def step_host(self, dt, st_in, st_out):
        n = self.n
        self._ck(self.L.np_world_step_host(self.w, float(np.float32(dt)), None if st_in is None else st_in.ctypes.data, st_out.ctypes.data, n))

    def counters(self):
        c = Counters(); self._ck(self.L.np_world_get_counters(self.w, C.byref(c))); return c.as_dict()

    def last_step_ms(self):
        out = np.zeros(6, np.float32); self._ck(self.L.np_world_last_step_ms(self.w, _f(out))); return out

    # -- narrow phase ----------------------------------------------------------------------------
    def detect(self, sa, xa, sb, xb, is3d=True):
        cd = CollisionData(); hit = C.c_int()
        self._ck(self.L.np_detect_collision(self.w, sa, _f(np.ascontiguousarray(xa, np.float32)), sb, _f(np.ascontiguousarray(xb, np.float32)), int(is3d), C.byref(cd), C.byref(hit)))
        return hit.value, np.array(list(cd.penetration_direction) + [cd.depth, float(cd.success)], np.float32)

    def detect_batch(self, sa, posea, sb, poseb, is3d=True):
        sa = np.ascontiguousarray(sa, np.int32); sb = np.ascontiguousarray(sb, np.int32)
        posea = np.ascontiguousarray(posea, np.float32); poseb = np.ascontiguousarray(poseb, np.float32)
        n = len(sa); hit = np.zeros(n, np.int32); cd = np.zeros(n, COLLISION_DATA_DTYPE)
        fn = self.L.np_detect_collision_batch if is3d else self.L.np_detect_collision_batch_2d
        self._ck(fn(self.w, n, _i(sa), _f(posea), _i(sb), _f(poseb), _i(hit), cd.ctypes.data))
        out = np.zeros((n, 5), np.float32)
        out[:, 0:3] = cd["penetration_direction"]; out[:, 3] = cd["depth"]; out[:, 4] = cd["success"]
        return hit, out

    # -- device buffers --------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = self.L.np_device_alloc(self.w, nbytes)
        if not p:
            raise NetphysError(self.L.np_last_error().decode())
        return p

    def dev_free(self, p): self._ck(self.L.np_device_free(self.w, p))
    def dev_upload(self, p, arr): arr = np.ascontiguousarray(arr); self._ck(self.L.np_device_upload(self.w, p, arr.ctypes.data, arr.nbytes))
    def dev_download(self, arr, p): self._ck(self.L.np_device_download(self.w, arr.ctypes.data, p, arr.nbytes))

    def host_array(self, shape, dtype=np.float32):
        """numpy array backed by page-locked host memory (np_host_alloc); free with host_free(arr) before the world closes."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.L.np_host_alloc(self.w, nbytes)
        if not p:
            raise NetphysError(self.L.np_last_error().decode())
        buf = (C.c_char * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        arr[...] = 0
        self._pinned = getattr(self, "_pinned", {}); self._pinned[arr.ctypes.data] = p
        return arr

    def host_free(self, arr):
        p = getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)
        if p: self._ck(self.L.np_host_free(self.w, p))

    def detect_batch_device(self, n, d_sa, d_pa, d_sb, d_pb, d_hit, d_out, timed=True):
        ms = C.c_float(0.0)
        self._ck(self.L.np_detect_collision_batch_device(self.w, n, d_sa, d_pa, d_sb, d_pb, d_hit, d_out, C.byref(ms) if timed else None))
        return ms.value

    def snapshot(self):
        out = np.zeros(self.n, SNAPSHOT_DTYPE)
        self._ck(self.L.np_world_snapshot(self.w, out.ctypes.data, self.n)); return out

    def snapshot_device(self):
        p = C.c_void_p(); nb = C.c_size_t()
        self._ck(self.L.np_world_snapshot_device(self.w, C.byref(p), C.byref(nb))); return p.value, nb.value

    def flush_l2(self): self._ck(self.L.np_device_flush_l2(self.w))
    def record_frame(self, time_ms): return self._ck(self.L.np_world_record_frame(self.w, float(time_ms)))
    def frame_count(self): return self.L.np_world_frame_count(self.w)

    def fill_packet(self, packet_id=2342144, cap=None):
        cap = cap if cap is not None else 24 + 36 * self.n
        buf = np.zeros(max(cap, 1), np.uint8)
        nb = self._ck(self.L.np_world_fill_packet(self.w, packet_id, buf.ctypes.data, cap))
        return buf[:nb].tobytes()

    def set_snapshot_target(self, d_ptr, guid_base=0): self._ck(self.L.np_world_set_snapshot_target(self.w, d_ptr, guid_base))

    # -- NCCL behind the C ABI (np_comm_*) -----------------------------------------------------------
    @staticmethod
    def _prefer_bundled_nccl():
        """inside a process that also runs torch, load the libnccl.so.2 torch ships (one NCCL per process)"""
        if "NP_NCCL_LIB" not in os.environ:
            try:
                import nvidia.nccl as _n
                for d in list(getattr(_n, "__path__", [])):
                    p = os.path.join(d, "lib", "libnccl.so.2")
                    if os.path.exists(p):
                        os.environ["NP_NCCL_LIB"] = p; break
            except Exception:
                pass

    def comm_unique_id(self):
        self._prefer_bundled_nccl()
        buf = (C.c_char * 128)(); self._ck(self.L.np_comm_get_unique_id(buf)); return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        self._prefer_bundled_nccl()
        assert len(unique_id) == 128
        self._ck(self.L.np_comm_init(self.w, nranks, rank, C.c_char_p(unique_id)))

    def comm_destroy(self): self._ck(self.L.np_comm_destroy(self.w))
    def set_snapshot_gather(self, slot_bodies, guid_base=0): self._ck(self.L.np_world_set_snapshot_gather(self.w, slot_bodies, guid_base))

    def gather_mode(self): return self.L.np_world_gather_mode(self.w)

    def gathered_snapshot(self):
        p = C.c_void_p(); nb = C.c_size_t()
        self._ck(self.L.np_world_gathered_snapshot(self.w, C.byref(p), C.byref(nb))); return p.value, nb.value

    def last_gather_ms(self, back=0):
        out = np.zeros(2, np.float32); self._ck(self.L.np_world_last_gather_ms(self.w, back, _f(out))); return out

    def shard_over_comm(self): self._ck(self.L.np_world_shard_over_comm(self.w))
    def step_sharded(self, dt): self._ck(self.L.np_world_step_sharded(self.w, float(np.float32(dt))))

    def fp32_peak_tflops(self):
        out = np.zeros(1, np.float32); self._ck(self.L.np_selftest_fp32_peak(self.w, _f(out))); return float(out[0])

    def fp32_peaks_tflops(self):
        """(scalar, packed f32x2) non-FMA FP32 issue peaks measured on this device"""
        out = np.zeros(2, np.float32); self._ck(self.L.np_selftest_fp32_peaks(self.w, _f(out))); return float(out[0]), float(out[1])

    # -- self tests ------------------------------------------------------------------------------
    def sincos_hash(self, start, count):
        hs = C.c_uint64(); hc = C.c_uint64()
        self._ck(self.L.np_selftest_sincos_hash(self.w, start, count, C.byref(hs), C.byref(hc))); return hs.value, hc.value

    def support_table_check(self, shape, dirs, xf16):
        """(scan results, table results, used) of n support queries: np_selftest_support_table."""
        dirs = np.ascontiguousarray(dirs, np.float32); xf16 = np.ascontiguousarray(xf16, np.float32); n = len(dirs)
        a = np.zeros((n, 3), np.float32); b = np.zeros((n, 3), np.float32); used = np.zeros(n, np.int32)
        self._ck(self.L.np_selftest_support_table(self.w, shape, n, _f(dirs), _f(xf16), _f(a), _f(b), _i(used)))
        return a, b, used

    def support_table_info(self, shape):
        out = np.zeros(4, np.float32)
        has = self.L.np_shape_support_table_info(self.w, shape, _f(out))
        return None if has <= 0 else {"n": int(out[0]), "scan_cells": int(out[1]), "mean_candidates": float(out[2]), "reach": float(out[3])}

    def closest(self, kind, pts):
        pts = np.ascontiguousarray(pts, np.float32).ravel(); reg = np.zeros(1, np.int32); uvw = np.zeros(3, np.float32)
        self._ck(self.L.np_selftest_closest(self.w, kind, _f(pts), _i(reg), _f(uvw))); return int(reg[0]), uvw
