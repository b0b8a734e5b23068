"""Dynamic opcode histogram of one kernel from an ncu `--page source --csv` export (per-instruction executed counts).

usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv
       python scripts/ncu_opcode_hist.py src.csv [units]      # units: divide counts by this (e.g. the triplet count)
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot, smp, total = collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        parts = r[isrc].split()
        if parts and parts[0].startswith("@"):
            parts = parts[1:]
        if not parts:
            continue
        op = parts[0].split(".")[0].rstrip(";")
        n = int(r[ia])
        tot[op] += n
        smp[op] += int(r[ismp]) if r[ismp].isdigit() else 0
        total += n
    print(f"total warp instructions {total / 1e6:.1f} M" + (f"  ({total / units:.1f} per unit)" if units else ""))
    for op, n in tot.most_common(32):
        per = f"  per-unit {n / units:6.2f}" if units else ""
        print(f"{op:14s} {n / 1e6:8.2f} M {n / total * 100:5.1f} %{per}  samples {smp[op]}")


if __name__ == "__main__":
    main()
