"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout with the
keys the driver reads, and the roofline byte model matches SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    # 3(2r+1)^2 + 21 + ceil(P^2/8)/A : configs[1] r=5, A=64, P=42
    assert abs(bench.algorithmic_bytes_per_agent_step(5, 64, 42) - 387.453125) < 1e-9
    assert abs(bench.algorithmic_bytes_per_agent_step(3, 64, 38) - (147 + 21 + 181 / 64)) < 1e-9


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "64",
                          "--warmup", "3"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 64 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"]
    # both arms print the same `config` object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, bench.INSTANCES_PER_GPU)


def test_timed_steps_are_split_into_multi_step_launches_only():
    # (the driver's --steps 20 goes out as one 20-step launch; longer runs are split by split_steps)
    sys.path.insert(0, ROOT)
    import bench
    assert bench.split_steps(20, 16) == [10, 10]           # the driver's --steps 20: no single-step remainder
    assert bench.split_steps(8192, 16) == [16] * 512
    assert bench.split_steps(5, 16) == [5] and bench.split_steps(17, 16) == [9, 8]
    for k in range(1, 100):
        assert sum(bench.split_steps(k, 16)) == k and max(bench.split_steps(k, 16)) <= 16


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "4", "--warmup", "3"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
