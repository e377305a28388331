#!/bin/bash
# Round-2 GPU check (one gpurun call): parity suite, smoke, then an A/B of the single-lane issue pattern
# (elect.sync vs the round-1 `if (tid == 0)`), bench lines into gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/r02_gpu_check.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/r02_smoke.log
for v in elect noelect; do
  lib="$PWD/ttt_video_dit_b200/lib/libttt_b200.so"
  [ "$v" = noelect ] && lib="$PWD/ttt_video_dit_b200/lib/libttt_b200_noelect.so"
  for mode in fwd fwdbwd; do
    TTT_B200_LIB="$lib" timeout 300 python bench.py --nc 282 --mode $mode --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_ab_${v}_${mode}.json 2> gpurun_out/r02_ab_${v}_${mode}.err
  done
done
tail -5 gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_smoke.log
for f in gpurun_out/r02_ab_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)"; done
