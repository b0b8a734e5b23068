"""lcao_pair_contract_bwd on a crystal-shaped batch (few species pairs, E = 201 k), with and without d_rb."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import _lib, ops  # noqa: E402
from lcaonet_b200.synth import crystal_like_batch  # noqa: E402
from scripts.bench_kernels import timeit  # noqa: E402

DEV = "cuda"
g = crystal_like_batch(64, seed=1000, cutoff=6.0).to(DEV)
E, C, NL, O, Zd = g["edge_index"].shape[1], 128, 3, 8, 37
z = g["z"]
print("species:", torch.unique(z).tolist(), "E", E)
pair = (z[g["edge_index"][0]] * Zd + z[g["edge_index"][1]]).contiguous()
tab = torch.randn(Zd * Zd, O, C, device=DEV)
rb = torch.randn(E, O, device=DEV)
lgrp = torch.tensor([0, 0, 1, 0, 1, 0, 2, 1], dtype=torch.int32, device=DEV)
kptr, kperm = ops.bucket_sort(pair, Zd * Zd, stable=False)
dB = torch.randn(E, NL, C, device=DEV)
d_tab, d_rb = torch.empty_like(tab), torch.empty(E, O, device=DEV)
nbytes = int(_lib.load().lcao_pair_contract_bwd_scratch(E, Zd * Zd, O, C, 0))
scratch = torch.empty(nbytes // 4 + 4, dtype=torch.int32, device=DEV)
P, st = ops.ptr, ops.stream_ptr
for name, drb in (("no d_rb", None), ("with d_rb", d_rb)):
    t = timeit(lambda: ops._call("lcao_pair_contract_bwd", P(tab), P(pair), P(kptr), P(kperm), P(rb), None, P(lgrp), P(dB), E,
                                 Zd * Zd, O, C, NL, 0, P(d_tab), P(drb), P(scratch), st()), reps=7)
    print(f"pair_contract_bwd {name}: {t*1e3:.0f} us", flush=True)
