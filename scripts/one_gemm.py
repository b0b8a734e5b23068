"""One tcgen05 forward GEMM launch sequence for ncu (M = E*O of BASELINE configs[1])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcaonet_b200 import ops  # noqa: E402

M, K, N = 2_022_384, 128, 128
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") / K**0.5
y = torch.empty(M, N, device="cuda")
dy = torch.randn(M, N, device="cuda")
dw = torch.zeros(N, K, device="cuda")
P, st = ops.ptr, ops.stream_ptr
m = ops.GEMM_MODES[sys.argv[1] if len(sys.argv) > 1 else "tf32x3"]
for _ in range(3):
    ops._call("lcao_linear_fwd", P(x), K, P(w), None, P(y), N, None, N, M, K, N, 0, m, st())
    ops._call("lcao_linear_wgrad", P(dy), N, None, 0, 0, P(x), K, P(dw), None, M, K, N, m, None, st())
torch.cuda.synchronize()
print("ok")
