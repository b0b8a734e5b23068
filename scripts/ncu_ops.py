"""Opcode histogram (instructions executed, stall samples) of one kernel of an .ncu-rep source page.
Usage: python scripts/ncu_ops.py report.ncu-rep kernel-regex [n_top_lines]"""
import csv, io, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 15
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if "# Samples" in r]
hi, end = his[0], (his[1] if len(his) > 1 else len(rows))   # first matching launch only
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"samples {tot}, warp instructions {ti/1e6:.1f} M, SASS lines {len(data)}")
c, cs = Counter(), Counter()
for r in data:
    t = r[ix["Source"]].strip().split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?")).split(".")[0]
    c[op] += int(r[ix["Instructions Executed"]]); cs[op] += int(r[ix["# Samples"]])
for op, n in c.most_common(22):
    print(f"{op:12s} inst {n/1e6:8.2f} M ({100*n/ti:4.1f}%)  samples {100*cs[op]/tot:5.1f}%")
print()
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
    st = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "(" not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
    top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[ix['# Samples']]):6d} ({100*int(r[ix['# Samples']])/tot:4.1f}%) x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:60]:60s} {top}")
