"""GPU time of the embedding block (species tables, weighted BatchNorm, gathers) forward and backward inside one step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import LCAONet  # noqa: E402
from lcaonet_b200.synth import qm9_like_batch  # noqa: E402

dev = "cuda"
g = qm9_like_batch(1024, seed=1000).to(dev)
model = LCAONet(cutoff=5.0, cutoff_net="polynomial").to(dev).train()
z, ei = g["z"], g["edge_index"]


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(6):
    torch.cuda.synchronize()
    big = torch.zeros(64 * 1024 * 1024, device=dev)  # keep the GPU busy so that the host runs ahead (as in a real step)
    for _ in range(20):
        big.add_(1.0)
    e0 = ev()
    x, cst = model.emb_layer(z, ei[0], ei[1])
    e1 = ev()
    loss = x.square().mean() + cst.table.square().mean()
    for _ in range(20):
        big.add_(1.0)
    e2 = ev()
    loss.backward()
    e3 = ev()
    torch.cuda.synchronize()
    print(f"embedding forward {e0.elapsed_time(e1)*1e3:.0f} us | backward {e2.elapsed_time(e3)*1e3:.0f} us", flush=True)
