#!/bin/bash
# Round-2 GPU check #2: parity suite on the fp16-operand build, full-shape error report, bf16-vs-fp16 operand A/B.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu2.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02_pytest_gpu2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke2.log 2>&1; echo "smoke exit: $?" >> gpurun_out/r02_smoke2.log
for op in f16 bf16; do
  TTT_B200_OPERANDS=$op timeout 600 python scripts/r02_diag_bwd.py > gpurun_out/r02_diag_operands_${op}.log 2>&1
  for mode in fwd fwdbwd; do
    TTT_B200_OPERANDS=$op timeout 300 python bench.py --nc 282 --mode $mode --steps 10 --warmup 3 --no-cpu-baseline --no-secondary \
      > gpurun_out/r02_operands_${op}_${mode}.json 2> gpurun_out/r02_operands_${op}_${mode}.err
  done
done
tail -8 gpurun_out/r02_pytest_gpu2.log; tail -3 gpurun_out/r02_smoke2.log
for op in f16 bf16; do echo "== $op"; tail -2 gpurun_out/r02_diag_operands_${op}.log | cut -c1-420; done
for f in gpurun_out/r02_operands_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])" 2>&1)"; done
