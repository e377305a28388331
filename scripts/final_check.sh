#!/bin/bash
# Round-end verification in one gpurun call: GPU parity suite, smoke(), default bench line, secondary benches.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/final_smoke.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_default.json
cut -c1-260 gpurun_out/final_bench_default.json
timeout 600 python scripts/bench_extra.py 2>&1 | grep -v Warn | tail -8 > gpurun_out/final_bench_extra.log
cut -c1-300 gpurun_out/final_bench_extra.log
