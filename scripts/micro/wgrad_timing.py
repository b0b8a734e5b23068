"""Where the tcgen05 weight-gradient kernel waits (LCAO_TC_DEBUG=64: CTA 0 prints its per-chunk cycle accounting)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import ops, _lib  # noqa: E402
M, K, N = 252798, 128, 128
x = torch.randn(M, K, device="cuda"); dy = torch.randn(M, N, device="cuda")
P, st = ops.ptr, ops.stream_ptr
m = ops.GEMM_MODES["tf32x3"]
n_scr = int(_lib.load().lcao_linear_bwd_scratch(P(dy), N, None, 0, 0, None, P(x), K, None, 0, M, K, N, m))
scr = torch.empty(max(n_scr, 1), device="cuda"); dw = torch.zeros(N, K, device="cuda")
for _ in range(2):
    ops._call("lcao_linear_wgrad", P(dy), N, None, 0, 0, P(x), K, P(dw), None, M, K, N, m, P(scr), st())
    torch.cuda.synchronize()
