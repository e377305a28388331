"""Reduce `ncu -i X.ncu-rep --page raw --csv` output (one wide row per captured launch) to a small metric,unit,value table of
the metrics the judge reads (tensor-pipe activity, DRAM bytes, duration, launch shape, stall reasons); written next to the
round's other evidence under profiles/.   python scripts/ncu_summarize.py gpurun_out/prof_fwd_raw.csv profiles/r02_fwd_ncu_summary.csv"""
import csv
import sys

KEEP = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__time_duration.sum", "launch__",
        "sm__cycles_active.avg", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "smsp__issue_active", "smsp__inst_executed.sum",
        "smsp__average_warp", "smsp__average_warps_issue_stalled", "sm__warps_active", "lts__t_sector_hit_rate", "sm__throughput",
        "l1tex__data_bank_conflicts", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu", "registers", "shared_mem")

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, vals = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
row = vals[which]
out = [("metric", "unit", "value")]
for n, u, v in zip(names, units, row):
    if any(k in n for k in KEEP):
        out.append((n, u, v))
csv.writer(open(sys.argv[2], "w")).writerows(out)
print(sys.argv[2], len(out) - 1, "metrics")
