"""Brief of one ncu report: duration, occupancy, pipe/issue utilisation, stall reasons per issued instruction, DRAM bytes.
Usage: python scripts/ncu_brief.py report.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    print("==", d.get("Kernel Name", "")[:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
            "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
    for k in keys:
        if k in d:
            print(f"  {k:75s} {d[k]}")
    st = [(float(x.replace(',', '')), k) for k, x in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and x not in ("", "n/a")]
    for x, k in sorted(st, reverse=True)[:10]:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {x:.3f}")
