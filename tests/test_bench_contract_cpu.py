"""bench.py's reference arm (`--impl reference`: the oracle port of the path on the host cores) runs without a GPU and
prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--no-config2-sample"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["metric"] == "T2I images/s (1024 tok, 18 steps)" and r["unit"] == "images/s"
    assert r["higher_is_better"] is True and r["n_gpus"] == 1 and r["steps"] == 2 and r["warmup"] == 1 and r["value"] > 0
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == r["value"] and "sample" in cb
    assert r["e2e"] == {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and r["gpu_launches"] == 0
    # what is reported is what was timed: ms_per_step x steps is the measured wall time of the timed steps, and with fewer than 18
    # steps the per-image figure is flagged as extrapolated (with --steps >= 18 the first 18 steps are one complete generation)
    assert cb["extrapolated"] is True and cb["seconds_per_image"] is None
    assert abs(r["ms_per_step"] * 1e-3 - cb["seconds_per_maskgit_step"]) < 0.5 * cb["seconds_per_maskgit_step"]
    assert "256 image tokens" in r["config"]["workload"]


def test_ranks_other_than_zero_do_no_work_in_the_reference_arm():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
