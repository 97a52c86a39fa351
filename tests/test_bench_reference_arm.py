"""The reference arm of bench.py (`--impl reference`) runs on host cores only: its JSON contract and its wall-clock bound
are checked here without a GPU (the arm times oracle/_ref where it is built, else the oracle port)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_bounded_json_line():
    env = dict(os.environ, VGSIM_REF_BUDGET_S="24")
    t0 = time.time()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--cpu-seconds", "0.5"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=300)
    wall = time.time() - t0
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["steps"] == 3 and j["warmup"] == 1 and j["higher_is_better"] is True
    assert j["unit"] == "events/s" and j["value"] > 0 and j["e2e"]["value"] == j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["gpu_launches"] == 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "replicates" in cb["sample"]
    # 4 samples inside a 24 s budget (+ the calibration run and process start-up): far from the minutes an unbounded sample takes
    assert wall < 120, wall
