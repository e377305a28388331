#!/bin/bash
# Round profile: launch list of one bench step + one `ncu --set full` capture per kernel family (B200_PROFILING.md recipe).
# Run under gpurun from the repo root; writes gpurun_out/*.csv, *.ncu-rep and the reduced summaries.  Numbers printed under
# ncu are never bench values.  ncu serialises kernels, so the backward runs in its per-group launch mode here
# (TTT_B200_PERSISTENT=0: same kernel code; the persistent launch spins on flags of kernels ncu would never let run beside it).
set -u
export TTT_B200_PERSISTENT=0
R=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_fwdbwd_launches.csv \
    python bench.py --nc 282 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/launches_bench.log 2>&1
cap() {  # name regex target skip [count]
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c "${5:-1}" -f -o "gpurun_out/prof_$1" \
      python scripts/ncu_target.py "$3" > "gpurun_out/prof_$1.log" 2>&1
  ncu -i "gpurun_out/prof_$1.ncu-rep" --page raw --csv > "gpurun_out/prof_$1_raw.csv" 2>/dev/null
  python scripts/ncu_summarize.py "gpurun_out/prof_$1_raw.csv" "gpurun_out/${R}_$1_ncu_summary.csv"
}
cap bwdK 'ttt_mlp_bwd_kernel' mlp 3
cap traj 'ttt_mlp_traj_kernel' mlp 3
cap bwdQ 'ttt_mlp_bwd_q_kernel' mlp 3
cap fwd 'ttt_mlp_fwd_kernel' mlp 1
ls -la gpurun_out/ | grep "${R}_" | head -20
