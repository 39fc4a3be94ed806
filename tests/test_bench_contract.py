"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with the
agreed keys; the helper functions behave."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ball_steps_per_sec" and d["unit"] == "ball-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_traffic_and_peaks_helpers():
    sys.path.insert(0, ROOT)
    import bench
    peak, src = bench.load_peaks()
    assert peak > 1000 and isinstance(src, str)
    t = bench.load_traffic("render")
    assert t is None or t > 1e9
    assert bench.BYTES_K3_PER_FRAME == 16384
    # one advance launch: header read + written, 28 B read + 12 B written per ball, 8 B per ball and snapshot
    assert bench._bytes_advance_per_launch(1, 8, 16, 16) == 32 + 8 * 40 + 16 * 8 * 8
    ref = bench.load_python_reference()
    assert "unavailable" in ref or ref["config4_sample_1_process"] > 100
