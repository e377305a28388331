#!/bin/bash
# Round-2 GPU check #5 (2 GPUs): sequence-sharded chain -- NCCL parity tests, then the sharded bench (small, then 63 s).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_2gpu_devices.txt
timeout 900 python -m pytest tests/test_gpu_seq_shard.py -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_seqshard_2gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_seqshard_2gpu.log
tail -5 gpurun_out/r02_pytest_seqshard_2gpu.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2: rc=$? $(tail -c 600 gpurun_out/$2.err | tail -2)"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/$2.json').read().strip().splitlines()[-1]); print({k: d.get(k) for k in ('value','ms_per_step','single_sequence','replicas')}, d['e2e']['value'], d['roofline']['frac'], d['config']['parallelism'][:60])
except Exception as e: print('no line', e)"; }
run 29611 r02_bench_2gpu_sharded_nc282 --nc 282 --seqs 8 --steps 5 --warmup 2
run 29612 r02_bench_2gpu_sharded_default --steps 5 --warmup 2
