"""CPU suite: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line, and the refusal to run the product
arm without a CUDA device (there is no CPU path to fall back to)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, NP_BENCH_BODIES="20000")                 # the default (1 000 000 bodies, ~7 s per CPU step) is for the GPU box
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "body_steps_per_sec" and d["unit"] == "body-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 1e4 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "20000 mixed convex bodies" in d["config"]["workload"] and d["config"]["bodies_per_gpu"] == 20000
    # both arms print the SAME config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    os.environ["NP_BENCH_BODIES"] = "20000"
    try:
        import importlib, bench
        importlib.reload(bench)
        assert d["config"] == bench.CONFIG
    finally:
        del os.environ["NP_BENCH_BODIES"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", NP_BENCH_BODIES="20000")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu(npb):
    if npb.device_count() > 0:
        pytest.skip("a GPU is visible here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_clock_sampler_reports_the_samples_of_the_timed_region():
    """The sampler runs from before the warm-up; what it reports are the rows that arrived after mark() (or, if none did, the last
    rows before it, and it says so).  Throttle reasons come from the reported rows only."""
    import time
    import types
    sys.path.insert(0, ROOT)
    import bench
    row = lambda mhz, cap: ["0", str(mhz), "1965", "400", "x", "Not Active", "Not Active", "Not Active", cap]
    now = time.perf_counter()
    c = bench.ClockSampler(0)
    c.p = types.SimpleNamespace(terminate=lambda: None, wait=lambda timeout=None: None, kill=lambda: None)
    c.rows = [(now - 1.0, row(1800, "Active")), (now - 0.5, row(1900, "Not Active"))]
    c.t0 = now - 0.2
    out = c.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1850.0 and "before the timed region" in out["window"] and out["reasons"] == ["sw_power_cap"]
    c.rows.append((now + 0.01, row(1965, "Not Active")))
    out = c.stop()
    assert out["samples"] == 1 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == [] and "timed region to" in out["window"]
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]
