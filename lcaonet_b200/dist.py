"""Data-parallel plumbing: one process per GPU, parameters replicated, ONE flat FP32 gradient bucket
all-reduced per step over NCCL (NVLink 5 / NVSwitch).  The reference has no distributed code; molecule
batches are independent, so the only exchange step of training is this gradient sum.
BatchNorm statistics stay per replica, as they would under DDP with the reference."""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Makes every parameter's .grad a view into one contiguous buffer so that the gradient exchange
    is a single all-reduce with no packing copies."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()
        # the interaction blocks may now write their weight gradients straight into these views
        for m in module.modules():
            if hasattr(m, "grads_in_place"):
                m.grads_in_place = True
            # ... and the embedding block may replay its static-shape table arithmetic as CUDA graphs
            if hasattr(m, "graph_tables") and self.flat.is_cuda:
                m.graph_tables = True

    def attach(self) -> None:
        """(Re-)point every `.grad` at its view of the flat buffer.  `optimizer.zero_grad()` / `module.zero_grad()` set
        the grads to None by default (`set_to_none=True`) and autograd then allocates fresh ones: use `bucket.zero()`
        instead — but if it happened, the gradient found in a detached `.grad` is copied back so nothing is lost."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None or g.data_ptr() != v.data_ptr():
                if g is not None:
                    v.copy_(g)
                p.grad = v

    def zero(self) -> None:
        """Replaces `zero_grad()`: clears the flat buffer and keeps every `.grad` attached to it."""
        from . import ops
        ops._pending_wgrad.update(desc=[], keep=[], queued=False, stream=None)  # (a backward pass that raised may have left entries)
        self.flat.zero_()
        self.attach()

    def all_reduce_mean(self) -> None:
        from . import ops
        ops.flush_wgrad()  # (no-op unless something is pending: the backward pass flushes its deferred weight gradients itself)
        self.attach()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Rank `src`'s parameters and buffers to every replica (one-time, at start)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` independent molecules owned by `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
