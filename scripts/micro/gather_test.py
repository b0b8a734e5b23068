import torch, sys, os
sys.path.insert(0, "/root/repo")
from lcaonet_b200 import ops
from lcaonet_b200.synth import qm9_like_batch
g = qm9_like_batch(1024, seed=1000).to("cuda")
N, E = g["z"].shape[0], g["edge_index"].shape[1]
gi = ops.GraphIndex(g["edge_index"], N)
B = torch.randn(E, 384, device="cuda")
out = torch.empty_like(B)
flush = torch.empty(128 * 1024 * 1024, device="cuda")
def timeit(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
idx = gi.in_edge.long()
rnd = torch.randperm(E, device="cuda")
t0 = timeit(lambda: out.copy_(B))
t1 = timeit(lambda: torch.index_select(B, 0, idx, out=out))
t2 = timeit(lambda: torch.index_select(B, 0, rnd, out=out))
nb = 2 * B.numel() * 4
print(f"copy {t0*1e3:.0f} us {nb/t0/1e6:.0f} GB/s | gather in_edge order {t1*1e3:.0f} us {nb/t1/1e6:.0f} GB/s | random {t2*1e3:.0f} us {nb/t2/1e6:.0f} GB/s")
# read-only reduction over gathered rows (sum) to isolate reads
t3 = timeit(lambda: B.sum())
print(f"read-only sum {t3*1e3:.0f} us {B.numel()*4/t3/1e6:.0f} GB/s")
