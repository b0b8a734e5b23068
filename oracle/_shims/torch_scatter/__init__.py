"""Stand-in for torch_scatter.scatter (sum / mean over dim 0..k), pure torch. Test infrastructure."""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    res = src.new_zeros(shape) if out is None else out
    res = res.scatter_add(dim, idx, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = src.new_zeros(dim_size).scatter_add(0, index, torch.ones_like(index, dtype=src.dtype))
        return res / cnt.clamp(min=1).view(view)
    raise ValueError(reduce)
