"""One tcgen05 forward / dgrad / wgrad launch at the edge-sized shape of BASELINE configs[1] (M = E), for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcaonet_b200 import _lib, ops  # noqa: E402

M, K, N = 252_798, 128, 128
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") / K**0.5
y, pre = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
dy = torch.randn(M, N, device="cuda")
dx = torch.empty(M, K, device="cuda")
dw = torch.zeros(N, K, device="cuda")
P, st = ops.ptr, ops.stream_ptr
m = ops.GEMM_MODES[sys.argv[1] if len(sys.argv) > 1 else "tf32x3"]
n_scr = int(_lib.load().lcao_linear_bwd_scratch(P(dy), N, None, 0, 0, None, P(x), K, None, 0, M, K, N, m))
scr = torch.empty(max(n_scr, 1), device="cuda")
for _ in range(2):
    ops._call("lcao_linear_fwd", P(x), K, P(w), None, P(y), N, None, N, M, K, N, 0, m, st())
    ops._call("lcao_linear_fwd", P(x), K, P(w), None, P(y), N, P(pre), N, M, K, N, 1, m, st())
    ops._call("lcao_linear_dgrad", P(dy), N, None, 0, 0, P(w), P(dx), K, M, K, N, 0, m, None, st())
    ops._call("lcao_linear_wgrad", P(dy), N, None, 0, 0, P(x), K, P(dw), None, M, K, N, m, P(scr), st())
torch.cuda.synchronize()
print("ok")
