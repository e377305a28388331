#!/bin/bash
# Round-2 GPU check #7 (N GPUs, default 8): the driver's multi-GPU bench command on the sequence-sharded chain.
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02_${N}gpu_devices.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 5 --warmup 3 ) \
  > gpurun_out/r02_bench_${N}gpu_sharded.json 2> gpurun_out/r02_bench_${N}gpu_sharded.err
echo "rc=$?"; grep -v "ProcessGroupNCCL\|Warning\|^$" gpurun_out/r02_bench_${N}gpu_sharded.err | tail -8
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_${N}gpu_sharded.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
    print(d.get('single_sequence')); print(d.get('replicas')); print(d['config']['parallelism'])
except Exception as e: print('no line', e)
PY
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
