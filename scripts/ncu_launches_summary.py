"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the captured span).
Usage: python scripts/ncu_launches_summary.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "")
    name = re.sub(r"^void ", "", name)
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * 1000.0 if unit in ("ms", "msecond") else val
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"captured launches: {sum(v[0] for v in agg.values())}, total {tot/1000:.3f} ms (cold-cache, serialised: compare SHARES)\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k[:90]}` | {n} | {us:.1f} | {100*us/tot:.1f} % |")
