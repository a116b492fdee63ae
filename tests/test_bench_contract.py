"""bench.py's JSON contract, checked on the CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-seconds", "6"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "candidate allocations/sec" and d["unit"] == "candidates/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["cells_per_gpu"] == 100 * 4 * 256 * 64 and "workload" in d["config"]
    assert "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["processes"] >= 1
    assert cb["parallel_speedup"] > 0.5 and "sample" in cb and "cpu_model" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_sample_pairs_and_quota_helpers():
    sys.path.insert(0, ROOT)
    import bench
    p = bench.sample_pairs(16)
    assert len(p) == 16 and p == sorted(set(p)) and 0 <= p[0] and p[-1] < 400
    assert bench.sample_pairs(10 ** 6) == list(range(400))
    q = bench.cpu_quota_cores()
    assert q is None or q > 0
    cfg = bench.config_dict(8)
    assert "dp8" in cfg["parallelism"] and cfg["l2_flush_between_steps"] is True
