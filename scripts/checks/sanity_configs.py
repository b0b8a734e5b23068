"""Ad-hoc robustness checks on the GPU box (not part of pytest): odd channel counts and max_z = 94 against the FP64
oracle through the training fast path (FlatGradBucket: in-place gradients + graphed embedding tables), and one
8192-molecule step for memory headroom."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcaonet_b200 import LCAONet  # noqa: E402
from lcaonet_b200.dist import FlatGradBucket  # noqa: E402
from lcaonet_b200.synth import qm9_like_batch  # noqa: E402
from oracle import lcao_oracle as O  # noqa: E402
from tests._util import full_cfg  # noqa: E402

dev = "cuda"
for kw in (dict(emb_size=40, emb_size_coeff=24, emb_size_conv=20, max_z=10), dict(max_z=94, emb_size=64, emb_size_coeff=64, emb_size_conv=64),
           dict(emb_size_conv=96, n_interaction=2, add_valence=True, activation="tanh")):
    kw = dict(cutoff=5.0, cutoff_net="polynomial", **kw)
    torch.manual_seed(0)
    model = LCAONet(**kw).to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    bucket = FlatGradBucket(model)
    g = qm9_like_batch(6, seed=11, cutoff=5.0, margin=0.05)
    for it in range(2):  # second step replays the graphs
        bucket.zero()
        out = model(g.to(dev))
        (out**2).mean().backward()
    p = O.cast_params(sd, torch.float64, requires_grad=True)
    g64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in g.items()}
    ref = O.forward(p, full_cfg(kw), g64, training=True)
    (ref**2).mean().backward()
    err = float((out.detach().cpu().double() - ref.detach()).norm() / ref.detach().norm())
    gerr = max(float((q.grad.cpu().double() - p[n].grad).norm() / (p[n].grad.norm() + 1e-30))
               for n, q in model.named_parameters() if p[n].grad is not None and float(p[n].grad.norm()) > 1e-8)
    print(f"{kw}: energy rel-L2 {err:.2e}, worst gradient rel-L2 {gerr:.2e}", flush=True)
    assert err < 1e-5 and gerr < 2e-4

torch.manual_seed(0)
model = LCAONet(cutoff=5.0, cutoff_net="polynomial").to(dev).train()
bucket = FlatGradBucket(model)
g = qm9_like_batch(8192, seed=3, cutoff=5.0).to(dev)
torch.cuda.reset_peak_memory_stats()
for _ in range(2):
    bucket.zero()
    torch.nn.functional.mse_loss(model(g), g["y"]).backward()
torch.cuda.synchronize()
print(f"8192 molecules: E={g['edge_index'].shape[1]}, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
