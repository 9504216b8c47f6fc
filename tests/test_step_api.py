"""The reference's TEST_PROGRAM interface (engine/physics.h:115-127: DetectCollisionStep, FindIntersectionPoints, GetSearchDirection on a
caller-held Simplex, engine/simplex.h) through include/netphys/physics.h.

tests/cpp/step_trace_body.h is client code written against the reference's names only; oracle/gen_golden_steps.py compiled it against the
reference's own sources and committed every intermediate simplex (tests/golden/step_trace_ref.txt.gz).  Here the same text is compiled
against the shim; on the GPU its trace must equal the reference's bit for bit."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "tests", "shim_step.bin")


def build_exe():
    lib = os.path.join(ROOT, "netphys_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"),
                        os.path.join(ROOT, "tests", "shim_step.cpp"), "-o", EXE, "-L", lib, "-lnetphys_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def golden_lines():
    with gzip.open(os.path.join(GOLD, "step_trace_ref.txt.gz"), "rt") as f:
        return f.read().splitlines()


def test_step_client_compiles_against_the_shim(npb):
    build_exe()


def test_reference_test_3d_stepping_code_compiles_unchanged(tmp_path):
    """engine/test_3d.cpp:15-123 (its statics, ResetSimplex and the whole key handler: every call it makes into the engine) against the shim,
    with declarations for the platform layer's four input functions standing in for platform.h."""
    src = "/root/reference/engine/test_3d.cpp"
    if not os.path.exists(src):
        pytest.skip("/root/reference is not mounted here")
    lines = open(src).readlines()
    end = next(i for i, l in enumerate(lines) if l.startswith("static void Update("))       # the key handler ends where Update() begins
    body = "".join(lines[14:end])
    tu = tmp_path / "tu.cpp"
    tu.write_text("#define TEST_PROGRAM\n#include <cassert>\n#include <netphys/physics.h>\n#include <netphys/physics_shape.h>\n"
                  "enum { ARROW_KEY_UP = 256, ARROW_KEY_DOWN, ARROW_KEY_LEFT, ARROW_KEY_RIGHT };\n"
                  "void Platform_BasicCameraInput(float); bool Platform_InputChangedDown(int); void Platform_ResetCamera();\n" + body)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "include"), str(tu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_face_normals_match_the_reference(npb):
    """np_simplex_face_normal (host arithmetic) == the normals Simplex::AddFace stored in the reference's trace (the four SetupForEPA faces)."""
    L = npb.load_library()
    L.np_simplex_face_normal.restype = None
    checked = 0
    for line in golden_lines():
        head, verts, faces, _, _ = line.split("|")
        h = head.split()
        if h[2] != "gjk" or h[5] != "nf=4":
            continue
        v = np.array([int(x, 16) for x in verts.split()], np.uint32).view(np.float32).reshape(4, 9).copy()
        f = faces.split()
        for k in range(4):
            idx = [int(x) for x in f[6 * k:6 * k + 3]]
            want = np.array([int(x, 16) for x in f[6 * k + 3:6 * k + 6]], np.uint32)
            out = np.zeros(3, np.float32)
            L.np_simplex_face_normal(v.ctypes.data_as(C.c_void_p), idx[0], idx[1], idx[2], out.ctypes.data_as(C.c_void_p))
            got = out.view(np.uint32).copy(); got[(got & 0x7fffffff) > 0x7f800000] = 0x7fc00000
            assert np.array_equal(got, want), line
            checked += 1
    assert checked >= 40


@pytest.mark.gpu
def test_step_trace_equals_the_reference(npb, tmp_path):
    build_exe()
    out = tmp_path / "trace.txt"
    r = subprocess.run([EXE, os.path.join(GOLD, "step_trace_in.txt"), str(out), "40"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    got = out.read_text().splitlines(); want = golden_lines()
    assert int(r.stdout.strip()) == 58
    for g, w in zip(got, want):
        assert g == w, "first difference:\n got  %s\n want %s" % (g[:400], w[:400])
    assert len(got) == len(want)
