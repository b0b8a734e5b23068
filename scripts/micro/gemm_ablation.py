import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import ops
M, K, N = 252_798, 128, 128
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K**0.5; y = torch.empty(M, N, device="cuda")
P, st = ops.ptr, ops.stream_ptr
m = ops.GEMM_MODES[os.environ.get("MODE", "tf32x3")]
fl = torch.empty(64 * 1024 * 1024, device="cuda")
ts = []
for i in range(8):
    fl.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops._call("lcao_linear_fwd", P(x), K, P(w), None, P(y), N, None, N, M, K, N, 0, m, st()); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(f"debug={os.environ.get('LCAO_TC_DEBUG', '0')} mode={os.environ.get('MODE', 'tf32x3')}: {sorted(ts)[3]:.1f} us")
