#!/bin/bash
# Round-2 GPU check #6: full parity suite + smoke + default bench (with the nc804 and dit_layer secondary keys).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu6.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_gpu6.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke6.log 2>&1; echo "smoke exit: $?" >> gpurun_out/r02_smoke6.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default6.json 2> gpurun_out/r02_bench_default6.err
timeout 300 python bench.py --nc 282 --mode fwd --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_fwd282_6.json 2>/dev/null
tail -4 gpurun_out/r02_pytest_gpu6.log; tail -2 gpurun_out/r02_smoke6.log; tail -4 gpurun_out/r02_bench_default6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d.get('nc804'), d.get('dit_layer'), d.get('cpu_baseline',{}).get('value'))
d=json.loads(open('gpurun_out/r02_bench_fwd282_6.json').read().strip().splitlines()[-1]); print('fwd 282', d['ms_per_step'])
PY
