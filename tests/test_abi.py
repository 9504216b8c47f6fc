"""CPU suite: the C-ABI library loads and exports exactly what include/netphys_b200.h declares; no compute calls."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("variant", ["exact", "fma"])
def test_library_exports_every_declared_symbol(npb, variant):
    L = npb.load_library(variant)
    assert len(npb.DECLARED_SYMBOLS) >= 40
    missing = [s for s in npb.DECLARED_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", npb.library_path(variant)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("np_")}
    assert exported == set(npb.DECLARED_SYMBOLS), (exported ^ set(npb.DECLARED_SYMBOLS))
    assert L.np_abi_version() == 1


def test_struct_layouts_match_header(npb):
    from netphys_b200 import lib
    assert C.sizeof(lib.StaticPhysicsData) == 4 * (3 + 3 + 3 + 4 + 1 + 1 + 9 + 9)      # physics.h:36-51, 33 words
    assert C.sizeof(lib.CollisionData) == 44 and C.sizeof(lib.Collision) == 52            # physics.h:18-31
    assert C.sizeof(lib.BodyState) == 48 and C.sizeof(lib.SnapshotObject) == 36           # common.h:120-127
    # the header compiles as plain C (no CUDA or C++ types on the boundary)
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "netphys_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback(npb):
    """Without a device the product refuses to run (it must never route through the oracle or any CPU path)."""
    if npb.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(npb.NetphysError, match="no CUDA device"):
        npb.World()
    needles = ("liboracle", "oracle/", "from oracle", "import oracle", "netphys_oracle", "npo_", "libnetphys_ref", "/root/reference")
    for root, _, files in os.walk(os.path.join(ROOT, "netphys_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, f)).read()
                found = [n for n in needles if n in text]
                assert not found, "%s references %s: the product path must not depend on the oracle or the reference tree" % (f, found)


def test_cpp_shim_compiles_against_reference_style_code():
    """include/netphys/physics.h keeps the reference's class names and signatures (engine/physics.h, physics_shape.h):
    a test_physics.cpp-style translation unit must compile and link against the C ABI."""
    src = os.path.join(ROOT, "tests", "shim_scene.cpp")
    out = os.path.join(ROOT, "tests", "shim_scene.bin")
    if not os.path.exists(src):
        pytest.skip("shim test program not present")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                        "-L", os.path.join(ROOT, "netphys_b200"), "-lnetphys_b200", "-Wl,-rpath," + os.path.join(ROOT, "netphys_b200")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_scene_generators_are_deterministic(npb):
    a = npb.scenes.lattice_scene(1000, spacing=1.2, seed=42); b = npb.scenes.lattice_scene(1000, spacing=1.2, seed=42)
    assert np.array_equal(a["sd"], b["sd"]) and np.array_equal(a["body_shape"], b["body_shape"])
    s, sa, pa, sb, pb = npb.scenes.random_pairs(100, seed=3); s2, sa2, pa2, sb2, pb2 = npb.scenes.random_pairs(100, seed=3)
    assert np.array_equal(pa, pa2) and np.array_equal(sb, sb2)
    sd = npb.lib.sd_rows_to_struct(a["sd"], a["response"])
    assert sd["mass"][0] == 1.0 and sd["collision_response"][5] == 1 and np.allclose(sd["inverse_inertia_tensor"][0][0], 6.0)


def test_error_codes_without_device(npb):
    """Argument validation happens before any device work: bad handles return NP_ERR_INVALID, never crash."""
    L = npb.load_library()
    assert L.np_world_body_count(None) == 0
    assert L.np_world_step(None, 0.0) == -1 and L.np_world_synchronize(None) == -1
    assert L.np_shape_create_box(None, 1.0, 1.0, 1.0) == -1
    assert L.np_detect_collision_batch(None, 0, None, None, None, None, None, None) == -1
    assert b"null" in L.np_last_error() or b"bad" in L.np_last_error()
    L.np_world_destroy(None)                       # no-op


def test_python_constants_match_the_header(npb):
    """The ctypes view mirrors the header's enums by hand: keep them in step."""
    import re
    text = open(os.path.join(ROOT, "include", "netphys_b200.h")).read()
    enum = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(NP_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    W = npb.World
    assert (W.OPT_BROADPHASE, W.OPT_GRID_CELL, W.OPT_COUNTERS, W.OPT_SNAPSHOT_EACH_STEP, W.OPT_PHASE_TIMING, W.OPT_SUPPORT_TABLES, W.OPT_FRICTION) == \
        tuple(enum[k] for k in ("NP_OPT_BROADPHASE", "NP_OPT_GRID_CELL", "NP_OPT_COUNTERS", "NP_OPT_SNAPSHOT_EACH_STEP", "NP_OPT_PHASE_TIMING", "NP_OPT_SUPPORT_TABLES", "NP_OPT_FRICTION"))
    assert (W.EXACT, W.GRID, W.SWEPT) == (enum["NP_BROADPHASE_EXACT"], enum["NP_BROADPHASE_GRID"], enum["NP_BROADPHASE_SWEPT"])
    assert enum["NP_OK"] == 0 and enum["NP_ERR_NO_DEVICE"] == -2
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(NP_PACKET_[A-Z_]+)\s+(\d+)", text)}
    assert defs == {"NP_PACKET_WORLD_STATE_UPDATE": 2342144, "NP_PACKET_NEW_CONNECTION": 2342341}
