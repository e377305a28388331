"""Warp-stall samples of a scan kernel bucketed by the MMA batch that precedes them in the (straight-line) step loop: the
release-build counterpart of the TICK phase timing (no instrumentation in the kernel).  Input: the SASS source page,
    ncu -i gpurun_out/prof_bwdK.ncu-rep --page source --csv | python scripts/ncu_samples_by_mma_batch.py bwdK
A new batch starts when a UTCHMMA is more than 0x300 bytes after the previous one."""
import csv
import sys

LABELS = {
    "bwdK": ["kernel prologue + loop top (before the A1 issue)", "A1 wait + A2 (Z1 -> X2, gelu', gelu'')", "apply-Q wait + A0 (carry -> bf16)",
             "A3 wait + A4 (LN statistics, gradZ2)", "A5 wait + element-wise (tokens 0-31)", "A6 wait + element-wise (tokens 32-63)",
             "A7 wait + A8 (second-order LN backward, dZ2, dV, d eta)", "A9 wait + A10 (dZ1)", "A11 wait + A12 (dK) + loop end"],
    "fwd": ["kernel prologue (state staging)", "loop top + P1 wait + P2 (GELU of the K and Q halves)", "P3 wait + P4 (LayerNorm, gradZ2, output)",
            "P5 wait + P6 (gradZ1)", "P7 wait + P8a (W1 image)", "P8b (W2 image, under the next P1) + loop end"],
}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else ""
    header, seq = None, []
    for r in csv.reader(sys.stdin):
        if not r:
            continue
        if r[0] == "Address":
            header = r
            i_s = header.index("# Samples")
            continue
        if header is None or not r[0].startswith("0x"):
            continue
        try:
            n = int(r[i_s])
        except ValueError:
            n = 0
        seq.append((int(r[0], 16), r[1], n))
    seq.sort()
    buckets, last_mma = [0], None
    for a, src, n in seq:
        if "UTCHMMA" in src:
            if last_mma is None or a - last_mma > 0x300:
                buckets.append(0)
            last_mma = a
        buckets[-1] += n
    tot = sum(buckets)
    labels = LABELS.get(kind, [])
    print(f"total samples {tot}; {len(buckets) - 1} MMA batches in address order")
    for i, n in enumerate(buckets):
        print(f"  {100.0 * n / tot:5.1f} %  {n:6d}  {labels[i] if i < len(labels) else 'after batch %d' % i}")


if __name__ == "__main__":
    main()
