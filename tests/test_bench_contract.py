"""bench.py's contract pieces that run without a GPU: the reference arm (the CPU oracle port timed on the host cores) and the
argument surface the driver uses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_prints_one_json_line_with_the_same_metric():
    lines = _run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--pairs", "20000")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "read_pairs_per_sec_identifyprimers" and d["unit"] == "read pairs/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "read pairs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "20000 read pairs" in d["config"]["workload"] and "jvm_probe" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    # under torchrun only rank 0 runs the arm; the other ranks print nothing and exit 0
    assert _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--pairs", "20000", env={"RANK": "1", "WORLD_SIZE": "2"}) == []
