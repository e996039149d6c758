"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints one JSON line
with the keys the measurement contract names, and times the reference's own run_trajectory (or,
where oracle/_ref is absent, the oracle port) on the bench's workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-rounds", "1", "--cpu-iters", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("collision events/sec")
    assert d["unit"] == "events/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 1e4 and d["e2e"] == {"value": d["value"], "unit": "events/s",
                                              "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == (os.cpu_count() or 1)
    assert cb["value"] == d["value"] and "run_trajectory" in cb["sample"]
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT,
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_is_the_twelve_transitions_from_the_committed_structures():
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.workload("test")
    assert len(wl) == 12 and all(pos.shape == (20, 3) for _, pos in wl)
    assert [len(d["permanent_bonds"]) for d, _ in wl] == [0] * 3 + [1] * 6 + [2] * 3
