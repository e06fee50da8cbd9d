"""bench.py's reference arm on the host cores (no GPU needed): the JSON line the driver parses, and the N>1 launch where only rank 0
works.  The GPU arm's line is checked by the driver itself; its keys are asserted in tests/test_gpu_unet.py-style runs on the box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", *args],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    return r


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    assert abs(line["ms_per_step"] - 1e3 / line["value"]) < 1e-6 * line["ms_per_step"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    cfg = line["config"]
    assert cfg["workload"].startswith("cfg2") and cfg["per_gpu_batch"] == 1 and "model" not in cfg
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"


def test_reference_arm_other_ranks_exit_without_work():
    """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the other ranks print nothing and exit 0."""
    r = _run(dict(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29599"), args=())
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    """The GPU arm through the same command line the driver uses (short: no secondary records, no CPU baseline legs)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "5", "--warmup", "3", "--no-also", "--no-cpu-baseline"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert "impl" not in line or line["impl"] != "reference"
    assert line["unit"] == "steps/s" and line["n_gpus"] == 1 and line["steps"] == 5 and line["warmup"] == 3 and line["value"] > 100
    assert line["scaling"] == "weak" and line["higher_is_better"] is True and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("cfg2") and "model" not in line["config"]
    assert line["gpu_launches"] == 5 * line["launches_per_step"] and line["launches_per_step"] >= 10
    e = line["e2e"]
    assert e["value"] > 100 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != line["value"]
    rf = line["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    clk = line["clocks"]
    assert clk is None or {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clk)
    assert line["detail"]["graph_captures_in_timed_region"] == 0
