"""Space stand-ins (used when gymnasium is absent) and the bench.py JSON contract of the CPU
(`--impl reference`) arm, which runs without a GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from carl_b200 import spaces

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_discrete_dict():
    b = spaces.Box(low=np.array([-1.0, 0.0]), high=np.array([1.0, np.inf]), dtype=np.float32)
    assert b.shape == (2,) and b.dtype == np.float32
    assert b.contains(np.array([0.0, 5.0], dtype=np.float32)) and not b.contains(np.array([2.0, 5.0], dtype=np.float32))
    assert b.sample().shape == (2,)
    d = spaces.Discrete(3)
    assert d.contains(2) and not d.contains(3) and 0 <= d.sample() < 3
    dd = spaces.Dict({"obs": b, "k": d})
    assert set(dd.keys()) == {"obs", "k"} and dd["k"] is d and len(dd) == 2
    s = dd.sample()
    assert dd.contains(s)
    bb = spaces.batch_box(b, 5)
    assert bb.shape == (5, 2)


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "20",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 20 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_non_zero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "5",
                        "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "5", "--no-cpu-baseline",
                        "--no-ant"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "step_api"):
        assert k in d, k
    assert d["steps"] == 40 and d["n_gpus"] == 1 and d["gpu_launches"] >= 1 and d["dtype"] == "f32"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert 0 < d["roofline"]["frac"] < 1.5
    # the timed region is a train of passes of the K-step plan, not one launch on an idle stream
    assert d["config"]["passes"] >= 100 and d["gpu_launches"] == d["config"]["passes"] * d["config"]["launches_per_pass"]
    # synchronous gym-style call beside the split-batch headline; float64 leg; what parity rests on
    assert d["e2e_sync"]["value"] > 0 and d["e2e"]["parts"] == 2 and d["value_f64"]["value"] > 0
    assert "unpinned" in d["parity"] and "pinned_by_published_known_answers" in d["parity"]


def test_issue_roofline_helper_reads_the_committed_capture():
    """bench.issue_roof: the Brax kernels are issue bound, so `ant_8192` reports warp instructions per env-step (ncu
    capture committed under profiles/) against 148 SMs x 4 schedulers x the SM clock; it never raises."""
    import bench

    r = bench.issue_roof("prof_brax_fma", 8192 * 20, 1.6e8, 1965.0)
    assert r is not None and 2000 < r["warp_instructions_per_env_step"] < 10000
    assert r["peak_env_steps_per_s"] == pytest.approx(148 * 4 * 1965e6 / r["warp_instructions_per_env_step"])
    assert 0.3 < r["frac"] < 1.0 and "profiles/" in r["source"]
    assert bench.issue_roof("no_such_capture", 1, 1.0, None) is None
