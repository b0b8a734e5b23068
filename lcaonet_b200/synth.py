"""Seeded synthetic graph batches with the shapes BASELINE.json names.

Host-side numpy; no dataset or network access is needed.  Two families:

* `qm9_like_batch`   – small organic molecules (H,C,N,O,F), ~18 atoms, all ordered pairs within the
  cutoff, edges sorted by (s, t); non-periodic (zero shifts, a dummy 50 Å lattice per molecule).
* `crystal_like_batch` – 64-atom jittered simple-cubic cells, edges over the 27 neighbouring
  images (duplicate (s, t) pairs with different shifts occur, as in real crystals).

`GraphBatch` is the minimal dict-like container `LCAONet.forward` needs (`[]`, `get`, item
assignment, `.to(device)`), playing the role PyG's `Batch` plays for the reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .keys import GraphKeys


class GraphBatch(dict):
    """dict of tensors with attribute access, `.to()` and `.pin_memory()`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to(self, *args, **kwargs) -> "GraphBatch":
        return GraphBatch({k: (v.to(*args, **kwargs) if torch.is_tensor(v) else v) for k, v in self.items()})

    def pin_memory(self) -> "GraphBatch":
        return GraphBatch({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in self.items()})

    def clone(self) -> "GraphBatch":
        return GraphBatch({k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.items()})

    @property
    def num_graphs(self) -> int:
        return int(self[GraphKeys.Lattice].shape[0])


_QM9_Z = np.array([1, 6, 7, 8, 9])
_QM9_P = np.array([0.51, 0.35, 0.06, 0.078, 0.002])


def _molecule_positions(rng: np.random.RandomState, n: int, min_dist: float = 0.95) -> np.ndarray:
    radius = (3.0 * n / (4.0 * math.pi * 0.09)) ** (1.0 / 3.0)
    pos = np.zeros((n, 3))
    k = 0
    while k < n:
        cand = rng.uniform(-radius, radius, size=(64, 3))
        cand = cand[(cand**2).sum(1) <= radius * radius]
        for c in cand:
            if k == 0 or ((pos[:k] - c) ** 2).sum(1).min() >= min_dist * min_dist:
                pos[k] = c
                k += 1
                if k == n:
                    break
    return pos


def qm9_like_batch(n_mol: int, seed: int = 0, cutoff: float = 5.0, margin: float = 0.0,
                   with_targets: bool = True) -> GraphBatch:
    """`n_mol` QM9-shaped molecules.  `margin` drops pairs with cutoff-margin < d <= cutoff (parity
    runs use 0.05 Å: the reference's FP32 polynomial cutoff is ill-conditioned on that shell,
    SURVEY.md Appendix B); throughput runs use 0."""
    rng = np.random.RandomState(seed)
    zs, ps, src, dst, bidx = [], [], [], [], []
    off = 0
    rc = cutoff - margin
    for m in range(n_mol):
        n = int(np.clip(np.rint(rng.normal(18.0, 3.0)), 4, 29))
        z = rng.choice(_QM9_Z, size=n, p=_QM9_P / _QM9_P.sum())
        p = _molecule_positions(rng, n)
        d = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
        s, t = np.nonzero((d <= rc) & ~np.eye(n, dtype=bool))  # row-major => sorted by (s, t)
        zs.append(z), ps.append(p), src.append(s + off), dst.append(t + off), bidx.append(np.full(n, m))
        off += n
    E = sum(len(s) for s in src)
    g = GraphBatch()
    g[GraphKeys.Z] = torch.from_numpy(np.concatenate(zs)).long()
    g[GraphKeys.Pos] = torch.from_numpy(np.concatenate(ps)).float()
    g[GraphKeys.Edge_idx] = torch.from_numpy(np.stack([np.concatenate(src), np.concatenate(dst)])).long()
    g[GraphKeys.Edge_shift] = torch.zeros(E, 3)
    g[GraphKeys.Lattice] = (50.0 * torch.eye(3)).repeat(n_mol, 1, 1)
    g[GraphKeys.Batch_idx] = torch.from_numpy(np.concatenate(bidx)).long()
    if with_targets:
        g["y"] = torch.from_numpy(rng.normal(size=(n_mol, 1))).float()
    return g


def crystal_like_batch(n_cell: int, seed: int = 0, cutoff: float = 6.0, margin: float = 0.0,
                       max_z: int = 36, with_targets: bool = True) -> GraphBatch:
    """`n_cell` periodic cells of 64 atoms (4x4x4 jittered simple cubic, ~50 neighbours/atom)."""
    rng = np.random.RandomState(seed)
    rho = 50.0 / (4.0 / 3.0 * math.pi * 6.0**3)
    L = (64.0 / rho) ** (1.0 / 3.0)
    a = L / 4.0
    grid = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    images = np.stack(np.meshgrid(*[np.arange(-1, 2)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    rc = cutoff - margin
    zs, ps, src, dst, shf, bidx = [], [], [], [], [], []
    for c in range(n_cell):
        p = (grid + 0.5 + rng.uniform(-0.15, 0.15, size=grid.shape)) * a
        z = rng.randint(1, max_z + 1, size=64)
        # d[s, t, img] = | p[t] + img*L - p[s] |
        diff = p[None, :, None, :] + images[None, None, :, :] * L - p[:, None, None, :]
        d = np.sqrt((diff**2).sum(-1))
        s, t, im = np.nonzero((d <= rc) & (d > 0))
        zs.append(z), ps.append(p), src.append(s + 64 * c), dst.append(t + 64 * c)
        shf.append(images[im]), bidx.append(np.full(64, c))
    g = GraphBatch()
    g[GraphKeys.Z] = torch.from_numpy(np.concatenate(zs)).long()
    g[GraphKeys.Pos] = torch.from_numpy(np.concatenate(ps)).float()
    g[GraphKeys.Edge_idx] = torch.from_numpy(np.stack([np.concatenate(src), np.concatenate(dst)])).long()
    g[GraphKeys.Edge_shift] = torch.from_numpy(np.concatenate(shf)).float()
    g[GraphKeys.Lattice] = (L * torch.eye(3)).repeat(n_cell, 1, 1).float()
    g[GraphKeys.Batch_idx] = torch.from_numpy(np.concatenate(bidx)).long()
    if with_targets:
        g["y"] = torch.from_numpy(rng.normal(size=(n_cell, 1))).float()
    return g


def reference_fixture_graph() -> GraphBatch:
    """The 3-atom periodic fixture the reference tests run on (tests/model/conftest.py:6-35): two
    periodic-image duplicates of edge (2->1) and edges far beyond any cutoff."""
    g = GraphBatch()
    g[GraphKeys.Lattice] = 15.0 * torch.eye(3).unsqueeze(0)
    g[GraphKeys.Pos] = torch.tensor([[5.187, 7.50, 7.50], [9.812, 7.50, 7.50], [5.187, 6.56, 7.50]])
    g[GraphKeys.Z] = torch.tensor([1, 3, 5])
    shift = torch.zeros(8, 3)
    shift[6] = torch.tensor([0.0, 1.0, -1.0])
    shift[7] = torch.tensor([1.0, 0.0, 0.0])
    g[GraphKeys.Edge_shift] = shift
    g[GraphKeys.Edge_idx] = torch.tensor([[0, 0, 1, 1, 2, 2, 2, 2], [1, 2, 0, 2, 0, 1, 1, 1]])
    return g


def graph_sizes(g) -> dict:
    """N, E, T (closed form: T = sum_e indeg(s_e) - #self-loop edges), B."""
    s, t = g[GraphKeys.Edge_idx]
    n = g[GraphKeys.Z].shape[0]
    indeg = torch.bincount(t.cpu(), minlength=n)
    T = int(indeg[s.cpu()].sum()) - int((s == t).sum())
    return {"N": n, "E": int(s.numel()), "T": T, "B": int(g[GraphKeys.Lattice].shape[0])}
