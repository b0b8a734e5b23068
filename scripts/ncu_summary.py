"""Summarise an .ncu-rep (raw page key metrics + top stalled SASS lines from the source page) per kernel.
Usage: python scripts/ncu_summary.py report.ncu-rep [n_top] [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]
name_ix = hdr.index("Kernel Name")
for k, vals in enumerate(rows[2:]):
    if want and want not in vals[name_ix]:
        continue
    print(f"\n## launch {k}: {vals[name_ix][:110]}\n\n| metric | unit | value |\n|---|---|---|")
    for h, u, v in zip(hdr, units, vals):
        if h in keep or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3:
            print(f"| {h} | {u} | {v} |")
if ntop <= 0:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + want] if want else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"\ntotal samples {tot}; top SASS lines by samples (first matching launch):\n")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
    st = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "(" not in k and int(r[ix[k]]) > 0}
    top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[ix['# Samples']]):6d} ({100*int(r[ix['# Samples']])/tot:4.1f}%) x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:64]:64s} {top}")
