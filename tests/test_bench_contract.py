"""CPU: the reference arm of bench.py (`--impl reference`) honours the JSON-line contract and the rank rules -- rank 0 alone
runs and prints, other ranks exit 0 without work -- and the host thread-count search returns a usable count."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env=None, *args):
    env = dict(os.environ, TORCHDYNAMO_DISABLE="1")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--heads", "2", *args], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = run_bench()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"].startswith("video-tokens/sec TTT-MLP layer-direction")
    cb = d["cpu_baseline"]
    ref_here = os.path.isdir("/root/reference/ttt/models/ssm")  # the real reference is timed where it exists (kind "reference")
    assert cb["kind"] == ("reference" if ref_here else "port") and cb["value"] == d["value"] and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) and 1 <= cb["threads"] <= cb["cores"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""
