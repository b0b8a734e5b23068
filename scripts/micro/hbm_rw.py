import torch
x = torch.empty(1 << 28, device='cuda')  # 1 GiB
y = torch.empty(1 << 28, device='cuda')
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print(f"write-only (zero_ 1 GiB): {ms:.3f} ms, {x.numel()*4/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy 1 GiB: {ms:.3f} ms, {2*x.numel()*4/ms/1e6:.0f} GB/s (read+write)")
ms = t(lambda: x.sum()); print(f"read-only (sum 1 GiB): {ms:.3f} ms, {x.numel()*4/ms/1e6:.0f} GB/s")
