"""bench.py contract pieces that run without a GPU: the reference arm (--impl reference) prints exactly one JSON line
with the keys the driver reads, on the same metric / unit / config as our arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("port", [False, True])
def test_reference_arm_prints_one_json_line(port):
    """--impl reference: the unmodified reference from baseline/_ref when it is staged (kind 'reference'), the oracle port
    otherwise (kind 'port'; forced here with NPLANE_BENCH_PORT=1); same metric / unit / config / steps / warm-up as our arm."""
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "envs", "control_env.py"))
    if not port and not have_ref:
        pytest.skip("baseline/_ref is staged only where /root/reference exists")
    env = dict(os.environ, NPLANE_BENCH_PORT="1") if port else dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--n", "2000", "--cpu-n", "2000", "--no-side"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "aircraft-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("aircraft-steps/sec at N=10^6") and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    sys.path.insert(0, ROOT)
    import bench
    cfg = {k: v for k, v in d["config"].items() if not (port and k == "sample_aircraft")}   # the port notes its reduced population
    assert cfg == bench.workload_config(2000)                  # the very object our arm prints
    assert d["cpu_baseline"]["kind"] == ("port" if port else "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
