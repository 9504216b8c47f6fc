"""GPU suite: reference-style C++ client code through include/netphys/physics.h (the drop-in shim) == the oracle."""
import os
import subprocess

import numpy as np
import pytest

from netphys_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_runs_test_physics_scene(oracle, npb):
    exe = os.path.join(ROOT, "tests", "shim_scene.bin")
    lib = os.path.join(ROOT, "netphys_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "shim_scene.cpp"), "-o", exe,
                        "-L", lib, "-lnetphys_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    steps = 120
    out = subprocess.run([exe, str(steps)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    sc = scenes.resolve(scenes.test_physics_scene(), oracle); ow = oracle.world(sc)
    for k in range(steps):
        ow.step(sc["dt"]); want = ow.state()[0, 0:6].view(np.uint32)
        got = np.array([int(x, 16) for x in lines[k].split()[1:]], np.uint32)
        assert np.array_equal(got, want), "step %d" % k
    st = ow.state()
    hit, o = oracle.detect(sc["shapes"][0], oracle.transform(st[0, 0:3], st[0, 3:6]), sc["shapes"][1], oracle.transform(st[1, 0:3], st[1, 3:6]))
    f = lines[steps].split()
    assert int(f[1]) == hit and int(f[2]) == int(o[4])
    assert np.array_equal(np.array([int(x, 16) for x in f[3:7]], np.uint32), o[0:4].view(np.uint32))
