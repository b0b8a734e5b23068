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
