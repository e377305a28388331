#!/bin/bash
# Round-2 GPU check #4: persistent K-side kernel -- backward parity suites, then A/B against the per-group launch mode.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mlp_backward.py tests/test_gpu_full_shape.py tests/test_gpu_tkmlp.py tests/test_gpu_ttt_layer.py -x -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu4.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_gpu4.log
tail -6 gpurun_out/r02_pytest_gpu4.log
for pm in 1 0; do
  for nc in 282 5487; do
    TTT_B200_PERSISTENT=$pm timeout 600 python bench.py --nc $nc --steps 10 --warmup 3 --no-cpu-baseline --no-secondary \
      > gpurun_out/r02_persist${pm}_nc${nc}.json 2> gpurun_out/r02_persist${pm}_nc${nc}.err
    echo "persistent=$pm nc=$nc: $(python -c "import json; d=json.loads(open('gpurun_out/r02_persist${pm}_nc${nc}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])" 2>&1 | tail -1)"
  done
done
