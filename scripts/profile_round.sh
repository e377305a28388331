#!/bin/bash
# Round profile: launch list of one bench step + one `ncu --set full` capture per kernel family (B200_PROFILING.md recipe).
# Run under gpurun from the repo root; writes gpurun_out/*.csv and gpurun_out/*.ncu-rep.  Numbers printed under ncu are
# never bench values.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwdbwd.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/launches_bench.log 2>&1
cap() {  # name regex target skip [count]
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c "${5:-1}" -f -o "gpurun_out/prof_$1" \
      python scripts/ncu_target.py "$3" > "gpurun_out/prof_$1.log" 2>&1
  ncu -i "gpurun_out/prof_$1.ncu-rep" --page raw --csv > "gpurun_out/prof_$1_raw.csv" 2>/dev/null
}
cap bwdK 'ttt_mlp_bwd_kernel' mlp 3
cap traj 'ttt_mlp_traj_kernel' mlp 3
cap bwdQ 'ttt_mlp_bwd_q_kernel' mlp 3
cap fwd 'ttt_mlp_fwd_kernel' mlp 1
cap linbwd 'ttt_linear_bwd_kernel' linear 1
cap attnbwd 'attn_bwd_kernel' attention 2 2
cap attnfwd 'attn_fwd_kernel' attention 1
ls -la gpurun_out/ | head -40
