"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU port of the
reference's path on a bounded sample) prints one JSON line with the agreed keys; under a multi-rank launch only
rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--grid", "128", "--res", "100", "--beams", "32", "--cols", "256", *args],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("OS1-128 scans/sec") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2")) == []


def test_committed_traffic_profile_parses():
    """bench.py reads roofline.traffic from the committed ncu capture; template arguments put commas into kernel names."""
    import bench
    total, src = bench.traffic_from_profile()
    assert src and src.startswith("profiles/") and 1e8 < total < 1e10
