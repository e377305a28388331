"""CPU: the committed measurement evidence is self-consistent -- the DRAM-traffic figures bench.py reports come out of the
committed ncu summaries, and the committed bench lines carry every key of the measurement contract."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _metric(summary, name):
    with open(os.path.join(PROF, summary)) as f:
        for row in csv.reader(f):
            if row and row[0] == name:
                return float(row[2])
    raise KeyError(name)


def test_traffic_json_matches_the_ncu_summaries():
    t = json.load(open(os.path.join(PROF, "r02_traffic.json")))
    # K-side backward: one launch = 16 steps x 48 CTAs
    rd, wr = _metric("r02_bwdK_ncu_summary.csv", "dram__bytes_read.sum"), _metric("r02_bwdK_ncu_summary.csv", "dram__bytes_write.sum")
    assert abs(t["fwdbwd"]["bytes_per_head_minibatch"] - (rd + wr) * 1e6 / (16 * 48)) < 2
    # forward: one launch = 48 steps x 48 CTAs
    rd, wr = _metric("r02_fwd_ncu_summary.csv", "dram__bytes_read.sum"), _metric("r02_fwd_ncu_summary.csv", "dram__bytes_write.sum")
    assert abs(t["fwd"]["bytes_per_head_minibatch"] - (rd + wr) * 1e6 / (48 * 48)) < 2


def test_committed_bench_lines_carry_the_contract_keys():
    for name, n in (("r02_bench_default_nc5487.json", 1), ("r02_bench_2gpu_sharded_nc5487_M16.json", 2), ("r02_bench_8gpu_sharded_nc5487_M64.json", 8)):
        line = open(os.path.join(PROF, name)).read().strip().splitlines()[-1]
        d = json.loads(line)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n
        if n == 1:
            assert d["warmup"] >= 3  # (the multi-GPU records of round 2 were taken with --steps 5 --warmup 2)
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        r = d["roofline"]
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
