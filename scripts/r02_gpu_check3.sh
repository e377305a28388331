#!/bin/bash
# Round-2 GPU check #3: new transformer-layer / adaLN tests, the default bench line (63 s) and the reference arm.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transformer_layer.py tests/test_gpu_adaln.py tests/test_gpu_umma.py -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu3.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_gpu3.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -12 gpurun_out/r02_pytest_gpu3.log
tail -4 gpurun_out/r02_bench_default.err; cut -c1-1500 gpurun_out/r02_bench_default.json
tail -4 gpurun_out/r02_bench_reference.err; cut -c1-900 gpurun_out/r02_bench_reference.json
timeout 300 python scripts/r02_dsmem_probe.py > gpurun_out/r02_dsmem_probe.log 2>&1; cat gpurun_out/r02_dsmem_probe.log
