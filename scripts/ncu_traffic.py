"""Regenerate profiles/ncu_traffic.json — the MEASURED DRAM traffic per launch of the edge kernels that bench.py quotes
as `roofline.traffic` — from an ncu metrics capture of one bench run:

    gpurun -- 'python bench.py --steps 2 --warmup 3 --ref-gpu-mols "" --cpu-reps 1 > gpurun_out/plain.log 2>&1 &&
      ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
          -k regex:^k_ --csv --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --ref-gpu-mols "" --cpu-reps 1'
    python scripts/ncu_traffic.py gpurun_out/traffic.csv

The file is stamped with the digest of the CUDA sources it was measured on (lcaonet_b200/csrc/build.py:_digest); bench.py
refuses it (traffic = null) when the sources have changed since.  Kernels are attributed to the C-ABI call that launches them."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lcaonet_b200.csrc.build import _digest  # noqa: E402

# kernel-name substring -> C-ABI call, and how many launches of that kernel one call makes
KERNELS = [("k_tb_fwd_mma", "lcao_threebody_fwd", 1), ("k_threebody_fwd", "lcao_threebody_fwd", 1),
           ("k_threebody_bwd", "lcao_threebody_bwd", 1), ("k_tb_bwd_staged", "lcao_threebody_bwd", 1), ("k_pair_contract_fwd", "lcao_pair_contract_fwd", 1),
           ("k_pair_reduce_partial", "lcao_pair_contract_bwd", 1), ("k_pair_reduce_final", "lcao_pair_contract_bwd", 1),
           ("k_chunk_ptr", "lcao_pair_contract_bwd", 1), ("k_pair_contract_drb", "lcao_pair_contract_bwd", 1),
           ("k_twobody_fwd", "lcao_twobody_fwd", 1), ("k_twobody_bwd", "lcao_twobody_bwd", 1),
           # dense layers: one C-ABI call = one (or, for 256 outputs, two) launches of the kernels below; the per-launch
           # figure of these calls is the step total divided by the calls per step (LINEAR_CALLS)
           ("k_tc_wgrad", "lcao_linear_wgrad", 1), ("k_wgrad_reduce", "lcao_linear_wgrad", 1), ("k_tiny_wgrad", "lcao_linear_wgrad", 1)]  # (k_wgrad_reduce matches k_wgrad_reduce_batch too)
LINEAR_CALLS = {"lcao_linear_wgrad": 27}  # calls per training step of the default model (bench.py `kernels`)


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    i_id = hdr.index("ID")
    per_launch = collections.defaultdict(dict)
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        per_launch[(r[i_id], r[i_name])][r[i_metric]] = v * scale
    agg = collections.defaultdict(lambda: {"bytes": 0.0, "us": 0.0, "launches": collections.Counter()})
    for (_, name), m in per_launch.items():
        for sub, call, _ in KERNELS:
            if sub in name:
                a = agg[call]
                a["bytes"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
                a["us"] += m.get("gpu__time_duration.sum", 0)
                a["launches"][sub] += 1
                break
    out = {"_source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of `{' '.join(sys.argv[1:])}` (bench.py step, config 2)",
           "csrc_digest": _digest()}
    steps = max(agg["lcao_threebody_bwd"]["launches"].values()) / 3.0 if "lcao_threebody_bwd" in agg else 1.0  # 3 layers per step
    for call, a in sorted(agg.items()):
        n_calls = max(a["launches"].values())  # every kernel of a call is launched once per call
        if call in LINEAR_CALLS:
            n_calls = int(round(steps * LINEAR_CALLS[call]))
        out[call] = {"dram_bytes_per_launch": int(a["bytes"] / n_calls), "ncu_us_per_launch": round(a["us"] / n_calls, 1),
                     "calls_profiled": n_calls, "kernels": dict(a["launches"])}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
