"""GPU suite, >= 2 devices: the multi-GPU entry points of the C ABI driven by a pure C++ program (no Python, no torch in the
ranks) -- np_comm_init, np_world_shard_over_comm + np_world_step_sharded, np_world_set_snapshot_gather -- the way the
reference's server loop (netphys_server/server.cpp:38-56) would call them.  tests/cpp/comm_two_ranks.cpp forks one process per GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_program_runs_sharded_step_and_snapshot_gather(npb, tmp_path):
    if npb.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    exe = str(tmp_path / "comm_two_ranks")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "comm_two_ranks.cpp"), "-ldl"], check=True)
    r = subprocess.run([exe, npb.lib.library_path(), "2", "4000", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout + r.stderr
    assert "gather mode 2" in r.stdout, "one process per GPU: the snapshot gather should run on the copy engines (CUDA IPC): " + r.stdout
    # the same program with the NCCL all-gather: the same bytes on every rank
    r2 = subprocess.run([exe, npb.lib.library_path(), "2", "4000", "6"], capture_output=True, text=True, timeout=600, env=dict(os.environ, NP_GATHER_NCCL="1"))
    assert r2.returncode == 0 and r2.stdout.startswith("ok:") and "gather mode 1" in r2.stdout, r2.stdout + r2.stderr
    assert r.stdout.split("gather mode")[0] == r2.stdout.split("gather mode")[0], "copy-engine gather and ncclAllGather disagree:\n" + r.stdout + r2.stdout
