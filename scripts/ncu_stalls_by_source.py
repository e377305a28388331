"""Reduce the source page of an `ncu --set full --import-source on` capture to a readable table: warp-stall samples per CUDA
source line (inlined helpers are attributed to their own lines: the mbarrier spin, the TMEM load wait, the GELU body ...),
and the totals per stall reason.

    ncu -i gpurun_out/prof_bwdK.ncu-rep --page source --csv --print-source cuda,sass | python scripts/ncu_stalls_by_source.py > profiles/r02_bwdK_stalls_by_source.txt
"""
import collections
import csv
import sys


def main():
    rows = csv.reader(sys.stdin)
    cur_file, header = None, None
    per_line = collections.Counter()
    text = {}
    reasons = collections.Counter()
    inst_class = collections.Counter()
    cur_line = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            i_samp = header.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(header) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if header is None:
            continue
        if r[0] != "":  # a CUDA source line
            cur_line = (cur_file, int(r[0]))
            text[cur_line] = r[1].strip()
            continue
        # a SASS instruction under the current source line
        try:
            n = int(r[i_samp])
        except (ValueError, IndexError):
            continue
        if n == 0:
            continue
        per_line[cur_line] += n
        for i, h in stall_cols:
            try:
                reasons[h] += int(r[i])
            except ValueError:
                pass
        op = r[3].split()[0] if r[3].split() else "?"
        if op.startswith("@"):
            op = r[3].split()[1]
        inst_class[op.split(".")[0]] += n
    total = sum(per_line.values())
    print(f"total warp-stall samples: {total}\n")
    print("by stall reason (all samples):")
    for h, n in reasons.most_common(12):
        print(f"  {h:28s} {n:8d}  {100.0 * n / max(1, sum(reasons.values())):5.1f} %")
    print("\nby SASS opcode the sampled warp was sitting at (top 16):")
    for op, n in inst_class.most_common(16):
        print(f"  {op:12s} {n:8d}  {100.0 * n / total:5.1f} %")
    print("\nby CUDA source line (top 40; inlined helpers appear under their own file):")
    for (f, ln), n in per_line.most_common(40):
        print(f"  {100.0 * n / total:5.1f} %  {f}:{ln:<5d} {text.get((f, ln), '')[:110]}")


if __name__ == "__main__":
    main()
