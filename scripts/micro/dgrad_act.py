"""lcao_linear_dgrad_act (SiLU' folded into the GEMM epilogue) against lcao_linear_dgrad + lcao_act_bwd."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import ops  # noqa: E402
from scripts.bench_kernels import timeit  # noqa: E402

P, st = ops.ptr, ops.stream_ptr
m = ops.GEMM_MODES["tf32x3"]
for M in (10952, 252798):
    K = N = 128
    dy, w, pre = torch.randn(M, N, device="cuda"), torch.randn(N, K, device="cuda") / K**0.5, torch.randn(M, K, device="cuda")
    dx, out = torch.empty(M, K, device="cuda"), torch.empty(M, K, device="cuda")

    def two():
        ops._call("lcao_linear_dgrad", P(dy), N, None, 0, 0, P(w), P(dx), K, M, K, N, 0, m, None, st())
        ops._call("lcao_act_bwd", P(dx), K, P(pre), K, P(out), K, M, K, 1, st())

    t2 = timeit(two, reps=9)
    t1 = timeit(lambda: ops._call("lcao_linear_dgrad_act", P(dy), N, P(w), P(pre), K, 1, P(dx), K, M, K, N, m, st()), reps=9)
    sg = torch.sigmoid(pre)
    ref = (dy.double() @ w.double()) * (sg * (1 + pre * (1 - sg))).double()
    err = ((dx.double() - ref).norm() / ref.norm()).item()
    print(f"M={M}: dgrad + act_bwd {t2*1e3:.1f} us | dgrad_act {t1*1e3:.1f} us | rel err {err:.2e}", flush=True)
