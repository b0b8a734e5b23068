"""PyG-free sample handling for the drop-in model.

The reference stores one `torch_geometric.data.Data` per structure (`data/convert.py:63` -> `{save_dir}/{i}.pt`),
reads them back in `GraphDataset.get` (`data/dataset.py:34-44`) and lets PyG's `Batch.from_data_list` collate them:
tensors are concatenated on dim 0, keys containing "index" on the LAST dim and incremented by the running atom
count (`data/keys.py:15`), per-structure attributes of shape (1, ...) (`lattice`, `pbc`, `neighbors`) stack to (B, ...),
and a `batch` vector is added.  This module restates exactly that for plain dict samples, so that a training loop needs
neither torch_geometric nor the ASE-based converter: `sample_from_structure` + `neighbors.attach_neighbor_list` build
the graph keys on the GPU."""
from __future__ import annotations

import os
from typing import Mapping, Sequence

import torch
from torch import Tensor

from .keys import GraphKeys
from .synth import GraphBatch


def _n_atoms(sample: Mapping) -> int:
    return int(sample[GraphKeys.Z].shape[0])


def collate(samples: Sequence[Mapping], pin_memory: bool = False) -> GraphBatch:
    """`Batch.from_data_list` for dict-like samples (every sample must hold the same tensor keys)."""
    if len(samples) == 0:
        raise ValueError("collate: empty sample list")
    keys = [k for k in samples[0].keys() if torch.is_tensor(samples[0][k])]
    out = GraphBatch()
    offsets, run = [], 0
    for s in samples:
        offsets.append(run)
        run += _n_atoms(s)
    for k in keys:
        if "index" in k:  # (.., n_items) index tensors: concatenate on the last dim, shift by the atom offset
            out[k] = torch.cat([s[k] + off for s, off in zip(samples, offsets)], dim=-1)
        else:
            out[k] = torch.cat([s[k] for s in samples], dim=0)
    out[GraphKeys.Batch_idx] = torch.cat([torch.full((_n_atoms(s),), i, dtype=torch.long) for i, s in enumerate(samples)])
    for k, v in samples[0].items():  # non-tensor attributes: kept as lists
        if not torch.is_tensor(v):
            out[k] = [s[k] for s in samples]
    return out.pin_memory() if pin_memory else out


def split(batch: Mapping) -> list[GraphBatch]:
    """Inverse of `collate` (`Batch.to_data_list`) for the reference's key set."""
    b = batch[GraphKeys.Batch_idx]
    n_graph = int(batch[GraphKeys.Lattice].shape[0])
    counts = torch.bincount(b, minlength=n_graph)
    starts = torch.cumsum(counts, 0) - counts
    ei = batch.get(GraphKeys.Edge_idx)
    e_graph = b[ei[0]] if ei is not None else None
    n_total, e_total = int(b.shape[0]), (int(ei.shape[1]) if ei is not None else -1)
    out = []
    for g in range(n_graph):
        lo, hi = int(starts[g]), int(starts[g] + counts[g])
        s = GraphBatch()
        emask = (e_graph == g) if e_graph is not None else None
        for k, v in batch.items():
            if k == GraphKeys.Batch_idx or not torch.is_tensor(v):
                continue
            if "index" in k:
                s[k] = v[..., emask] - lo
            elif v.shape[0] == n_total:
                s[k] = v[lo:hi]
            elif v.shape[0] == e_total:
                s[k] = v[emask]
            elif v.shape[0] == n_graph:
                s[k] = v[g:g + 1]
            else:
                raise ValueError(f"split: cannot assign key {k!r} of shape {tuple(v.shape)} to structures")
        out.append(s)
    return out


def sample_from_structure(numbers: Tensor, positions: Tensor, cell: Tensor | None = None, pbc=None, **props) -> GraphBatch:
    """One structure as a sample without edges (`atoms2graphdata` minus the neighbour list, convert.py:159-170);
    add the edges with `neighbors.attach_neighbor_list` — per sample or, better, on the collated batch."""
    s = GraphBatch()
    s[GraphKeys.Z] = torch.as_tensor(numbers, dtype=torch.long)
    s[GraphKeys.Pos] = torch.as_tensor(positions, dtype=torch.float32)
    cell = torch.zeros(3, 3) if cell is None else torch.as_tensor(cell, dtype=torch.float32)
    s[GraphKeys.Lattice] = cell.reshape(1, 3, 3)
    s[GraphKeys.PBC] = torch.as_tensor([False] * 3 if pbc is None else pbc).to(torch.long).reshape(1, 3)
    for k, v in props.items():
        s[k] = torch.as_tensor(v).reshape(1, -1) if not torch.is_tensor(v) else v
    return s


class GraphDataset(torch.utils.data.Dataset):
    """Reads the reference's on-disk format: `{save_dir}/{i}.pt`, one mapping of tensors per structure
    (data/dataset.py:20-44).  Objects saved by the reference are PyG `Data`; anything dict-like with the keys of
    `GraphKeys` works (PyG objects need torch_geometric importable to unpickle)."""

    def __init__(self, save_dir: str, inmemory: bool = False):
        if not os.path.isdir(save_dir):
            raise FileNotFoundError(f"{save_dir} does not exist. Please convert the dataset first.")
        self.save_dir, self.inmemory = save_dir, inmemory
        self._n = len([f for f in os.listdir(save_dir) if f.endswith(".pt")])
        if self._n == 0:
            raise ValueError("The dataset is empty.")
        self._cache = [None] * self._n if inmemory else None

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, idx: int):
        if idx >= self._n or idx < -self._n:
            raise IndexError("index out of range")
        idx %= self._n
        if self._cache is not None and self._cache[idx] is not None:
            return self._cache[idx]
        obj = torch.load(os.path.join(self.save_dir, f"{idx}.pt"), weights_only=False)
        sample = GraphBatch({k: obj[k] for k in obj.keys()}) if not isinstance(obj, dict) else GraphBatch(obj)
        if self._cache is not None:
            self._cache[idx] = sample
        return sample


class SamplePool:
    """A set of structures stored FLAT (struct-of-arrays) so that a batch is collated without a Python loop over its
    samples — on the GPU when the pool lives there (the 1 M-molecule stream of BASELINE config 5 at
    4096 molecules per GPU and step would otherwise be bound by `Batch.from_data_list` on the host).

    Storage: every atom-sized key concatenated over the pool (`z`, `pos`, ...), every edge-sized key likewise, with
    `edge_index` kept LOCAL to its structure (the offsets are added at collation time, `data/keys.py:15`), per-structure
    keys stacked, plus `atom_ptr` / `edge_ptr` (n + 1 offsets).  `batch(ids)` gathers the structures `ids` in that
    order and returns exactly what `collate([pool[i] for i in ids])` returns."""

    def __init__(self, samples: Sequence[Mapping] | None = None, *, flat: Mapping | None = None):
        if flat is not None:
            self.flat = GraphBatch(flat)
            return
        if not samples:
            raise ValueError("SamplePool: empty sample list")
        keys = [k for k in samples[0].keys() if torch.is_tensor(samples[0][k])]
        n_at = torch.tensor([_n_atoms(s) for s in samples])
        n_ed = torch.tensor([int(s[GraphKeys.Edge_idx].shape[1]) for s in samples])
        f = GraphBatch()
        for k in keys:
            f[k] = torch.cat([s[k] for s in samples], dim=-1 if "index" in k else 0)
        f["atom_ptr"] = torch.cat([n_at.new_zeros(1), n_at.cumsum(0)])
        f["edge_ptr"] = torch.cat([n_ed.new_zeros(1), n_ed.cumsum(0)])
        self.flat = f

    @classmethod
    def from_batch(cls, batch: Mapping) -> "SamplePool":
        """Pool holding the structures of an already collated batch (index keys made local again)."""
        b, ei = batch[GraphKeys.Batch_idx], batch[GraphKeys.Edge_idx]
        n_graph = int(batch[GraphKeys.Lattice].shape[0])
        n_at = torch.bincount(b, minlength=n_graph)
        e_graph = b[ei[0]]
        if e_graph.numel() > 1 and bool((e_graph[1:] < e_graph[:-1]).any()):
            raise ValueError("SamplePool.from_batch: edges must be grouped by structure")
        n_ed = torch.bincount(e_graph, minlength=n_graph)
        f = GraphBatch({k: v for k, v in batch.items() if torch.is_tensor(v) and k != GraphKeys.Batch_idx})
        f["atom_ptr"] = torch.cat([n_at.new_zeros(1), n_at.cumsum(0)])
        f["edge_ptr"] = torch.cat([n_ed.new_zeros(1), n_ed.cumsum(0)])
        f[GraphKeys.Edge_idx] = ei - f["atom_ptr"][e_graph]
        return cls(flat=f)

    def __len__(self) -> int:
        return int(self.flat["atom_ptr"].shape[0]) - 1

    def to(self, *args, **kwargs) -> "SamplePool":
        return SamplePool(flat=self.flat.to(*args, **kwargs))

    def pin_memory(self) -> "SamplePool":
        return SamplePool(flat=self.flat.pin_memory())

    def batch(self, ids: Tensor) -> GraphBatch:
        """Collate the structures `ids` (int64, on the pool's device): a dozen vectorised gathers, no host loop and — for
        a pool on the GPU — no host synchronisation except the two sizes (atoms, edges) of the new batch."""
        f = self.flat
        ap, ep = f["atom_ptr"], f["edge_ptr"]
        n_at, n_ed = ap[ids + 1] - ap[ids], ep[ids + 1] - ep[ids]
        out_ap = torch.cat([n_at.new_zeros(1), n_at.cumsum(0)])
        out_ep = torch.cat([n_ed.new_zeros(1), n_ed.cumsum(0)])
        n_tot, e_tot = int(out_ap[-1]), int(out_ep[-1])  # (one host read of two integers)
        slot = torch.arange(ids.numel(), device=ids.device)
        a_b = torch.repeat_interleave(slot, n_at, output_size=n_tot)  # = the `batch` vector
        e_b = torch.repeat_interleave(slot, n_ed, output_size=e_tot)
        a_src = ap[ids][a_b] + (torch.arange(n_tot, device=ids.device) - out_ap[a_b])
        e_src = ep[ids][e_b] + (torch.arange(e_tot, device=ids.device) - out_ep[e_b])
        n_pool_at, n_pool_ed, n_pool = int(f[GraphKeys.Z].shape[0]), int(f[GraphKeys.Edge_idx].shape[1]), len(self)
        out = GraphBatch()
        for k, v in f.items():
            if k in ("atom_ptr", "edge_ptr"):
                continue
            if "index" in k:
                out[k] = v[..., e_src] + out_ap[e_b]
            elif v.shape[0] == n_pool_at:
                out[k] = v[a_src]
            elif v.shape[0] == n_pool_ed:
                out[k] = v[e_src]
            elif v.shape[0] == n_pool:
                out[k] = v[ids]
            else:
                raise ValueError(f"SamplePool: cannot assign key {k!r} of shape {tuple(v.shape)} to structures")
        out[GraphKeys.Batch_idx] = a_b
        return out

    def window(self, lo: int, hi: int) -> GraphBatch:
        """The structures [lo, hi) of the pool as one batch, from contiguous SLICES of the flat arrays (what a streaming
        loader copies host -> device); index keys still local: finish with `collate_window` on the device."""
        f = self.flat
        a0, a1, e0, e1 = (int(f["atom_ptr"][lo]), int(f["atom_ptr"][hi]), int(f["edge_ptr"][lo]), int(f["edge_ptr"][hi]))
        n_pool_at, n_pool_ed, n_pool = int(f[GraphKeys.Z].shape[0]), int(f[GraphKeys.Edge_idx].shape[1]), len(self)
        out = GraphBatch()
        for k, v in f.items():
            if k == "atom_ptr":
                out[k] = v[lo:hi + 1] - a0
            elif k == "edge_ptr":
                out[k] = v[lo:hi + 1] - e0
            elif "index" in k:
                out[k] = v[..., e0:e1]
            elif v.shape[0] == n_pool_at:
                out[k] = v[a0:a1]
            elif v.shape[0] == n_pool_ed:
                out[k] = v[e0:e1]
            elif v.shape[0] == n_pool:
                out[k] = v[lo:hi]
        return out


def collate_window(win: Mapping) -> GraphBatch:
    """Finish a `SamplePool.window` on the device it was copied to: add the atom offsets to the index keys and build the
    `batch` vector (the two steps of `Batch.from_data_list` that touch every item).  No host synchronisation."""
    ap, ep = win["atom_ptr"], win["edge_ptr"]
    n = ap.numel() - 1
    slot = torch.arange(n, device=ap.device)
    n_tot, e_tot = int(win[GraphKeys.Z].shape[0]), int(win[GraphKeys.Edge_idx].shape[1])  # known from the shapes
    a_b = torch.repeat_interleave(slot, ap[1:] - ap[:-1], output_size=n_tot)
    e_b = torch.repeat_interleave(slot, ep[1:] - ep[:-1], output_size=e_tot)
    out = GraphBatch({k: v for k, v in win.items() if k not in ("atom_ptr", "edge_ptr")})
    for k in list(out.keys()):
        if torch.is_tensor(out[k]) and "index" in k:
            out[k] = out[k] + ap[e_b]
    out[GraphKeys.Batch_idx] = a_b
    return out
