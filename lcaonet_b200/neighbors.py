"""GPU neighbour list under periodic boundary conditions: the `edge_index` / `edge_shift` / `neighbors` keys that the
reference's offline converter produces (lcaonet/data/convert.py:103-172, via ase.neighborlist), built on the B200 for a
whole batch of structures at once.  See csrc/neighbor.cu for the exact semantics
(canonical tie-break, float64 distances) and oracle/neighbor_oracle.py for the CPU restatement it is tested against."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import LcaoError, call, ptr, require_cuda, stream_ptr
from .keys import GraphKeys


def build_neighbor_list(pos: Tensor, batch: Tensor | None, lattice: Tensor, pbc: Tensor | None, cutoff: float,
                        max_neighbors: int = 32):
    """(edge_index (2,E) int64, edge_shift (E,3) float32, neighbors (B,) int64).

    pos (N,3) float32; batch (N,) int64 with the atoms of one structure contiguous (None = one structure); lattice
    (B,3,3) float32, rows = cell vectors; pbc (B,3) bool/int (None = non-periodic).  One host sync (E is data dependent)."""
    require_cuda(pos, lattice)
    dev = pos.device
    N, B = pos.shape[0], lattice.shape[0]
    pos_c, lat_c = pos.detach().contiguous().float(), lattice.detach().contiguous().float()
    if batch is None:
        batch = torch.zeros(N, dtype=torch.int64, device=dev)
    batch = batch.contiguous()
    if N > 1 and bool((batch[1:] < batch[:-1]).any()):
        raise ValueError("build_neighbor_list: atoms of one structure must be contiguous (batch must be non-decreasing)")
    pbc_i = (torch.zeros(B, 3, dtype=torch.int32, device=dev) if pbc is None else pbc.to(device=dev, dtype=torch.int32).contiguous())
    n_atoms = torch.bincount(batch, minlength=B)
    gptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    gptr[1:] = n_atoms.cumsum(0).to(torch.int32)
    cnt = torch.zeros(N, dtype=torch.int32, device=dev)
    st = stream_ptr()
    call("lcao_neighbor_count", ptr(pos_c), ptr(batch), ptr(gptr), ptr(lat_c), ptr(pbc_i), N, float(cutoff), ptr(cnt), st)
    per_graph = torch.zeros(B, dtype=torch.int64, device=dev).index_add_(0, batch, cnt.long())
    fallback = (per_graph == 0) & (n_atoms > 1)  # no neighbour at all: fully linked graph (convert.py:154-157)
    deg = torch.where(fallback[batch], n_atoms[batch] - 1, cnt.long().clamp(max=max_neighbors))
    out_ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    out_ptr[1:] = deg.cumsum(0)
    E = int(out_ptr[-1])  # host sync
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    edge_shift = torch.empty(E, 3, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    call("lcao_neighbor_fill", ptr(pos_c), ptr(batch), ptr(gptr), ptr(lat_c), ptr(pbc_i), ptr(fallback.to(torch.int32)),
         ptr(out_ptr), N, E, float(cutoff), int(max_neighbors), ptr(edge_index), ptr(edge_shift), ptr(status), st)
    if int(status) != 0:
        raise LcaoError("build_neighbor_list: an atom has more than 2048 neighbours within the cutoff (unsupported)")
    neighbors = torch.zeros(B, dtype=torch.int64, device=dev).index_add_(0, batch, deg)
    return edge_index, edge_shift, neighbors


def attach_neighbor_list(graph, cutoff: float, max_neighbors: int = 32):
    """Fill `edge_index`, `edge_shift`, `neighbors` of a batch (keys of lcaonet/data/keys.py) in place."""
    ei, sh, nb = build_neighbor_list(graph[GraphKeys.Pos], graph.get(GraphKeys.Batch_idx), graph[GraphKeys.Lattice],
                                     graph.get(GraphKeys.PBC), cutoff, max_neighbors)
    graph[GraphKeys.Edge_idx], graph[GraphKeys.Edge_shift], graph[GraphKeys.Neighbors] = ei, sh, nb
    return graph
