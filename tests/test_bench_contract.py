"""bench.py contract on CPU: the reference arm (the reference's own md.vv loop on the host cores)
prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "trajectory_steps_per_s" and d["unit"] == "trajectory-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_algorithmic_flop_accounting_matches_the_survey():
    """F_alg = 2 nd^2 + sum_nonlocal 2 (ml+2) nc^2 + sum_local 6 m nc^2 = 142.2 kflop at cfg2,
    and the per-kernel split adds up to it."""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.workload("cfg2", 1)
    fl = bench.algorithmic(c)
    assert fl["total"] == 142200 and fl["bytes"] == 8 * (90 + 3 + 1)
    assert abs(fl["fused"] + fl["tail"] - fl["total"]) < 1e-9
    assert fl["far_executed"] < 0.2 * fl["tail"]
    assert bench.default_piece(c) == 8192
    # config 4 stretch: 2 nd^2 + 2 x 2 (ml + 2) nc^2 without the electron bath
    s = bench.algorithmic(bench.workload("cfg4s", 1))
    assert s["total"] == 2 * 510 ** 2 + 2 * 2 * 2002 * 81 ** 2
    # config 3 is the strong-scaling layout: 8192 trajectories over the GPUs of the box
    assert bench.workload("cfg3", 8)["batch"] == 1024 and bench.workload("cfg3", 8)["scaling"] == "strong"
    w = bench.workload("sweep:256,1000,4096", 1)
    assert (w["nd"], w["ml"], w["batch"], w["ncl"]) == (255, 1000, 4096, 63)
