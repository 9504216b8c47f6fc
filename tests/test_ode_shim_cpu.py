"""CPU suite: the ODE-shaped facade compiles -- against this repo's call-site program and, where /root/reference is mounted, against
the reference's own ODE callers, unchanged (netphys_common/world.cpp, worldobject.cpp, player.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "include", "netphys", "ode_shim")


def test_call_site_program_compiles_and_links(npb):
    exe = os.path.join(ROOT, "tests", "shim_world.bin"); lib = os.path.join(ROOT, "netphys_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-I", SHIM, os.path.join(ROOT, "tests", "shim_world.cpp"), "-o", exe,
                        "-L", lib, "-lnetphys_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("src", ["netphys_common/world.cpp", "netphys_common/worldobject.cpp", "netphys_common/player.cpp", "netphys_server/world_s.cpp", "netphys_server/player_s.cpp"])
def test_reference_ode_callers_compile_unchanged(src):
    path = os.path.join("/root/reference", src)
    if not os.path.exists(path):
        pytest.skip("/root/reference is not mounted here")
    # -fpermissive / intrin.h + WinSock2.h stubs / <cstdint> / the LOG macros' empty __VA_ARGS__: MSVC-isms of netphys_common/lib.h, log.h and
    # netphys_server/network_s.h, unrelated to ODE
    extra = ["-include", "cstdint"]
    if src.startswith("netphys_server/"):
        extra = ["-I", "/root/reference/netphys_server", "-include", os.path.join(ROOT, "oracle", "ref", "pre_server.h")]
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-fpermissive"] + extra + ["-I", SHIM, "-I", os.path.join(ROOT, "oracle", "ref", "stub"), path],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
