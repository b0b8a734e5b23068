"""Brute-force CPU oracle for the periodic neighbour list — TEST INFRASTRUCTURE, never the product path.

Restates `atoms2graphdata` (reference lcaonet/data/convert.py:103-172) for one structure in float64 numpy:
ase.neighborlist.neighbor_list("ijdS", cutoff, self_interaction=False) + per-centre distance sort +
`dist <= cutoff` filter + truncation to `max_neighbors` + the fully linked fallback (convert.py:154-157,
data/utils.py:10-20).

PARITY UNPINNED: `ase` (setup.py:17, `ase==3.*`, not vendored) is absent from this image and the reference has no
test for convert.py, so this restatement follows ASE's published behaviour — all (i, j, S) with
|pos_j + S.cell - pos_i| < cutoff, no (i, i, 0) — and is not checked against ASE itself.  The order among exactly
equidistant neighbours is implementation-defined in the reference (np.argsort, unstable); the canonical order here
and in the CUDA kernel is (distance, j, S0, S1, S2) ascending.  Arithmetic: every multiply and add separately rounded
in float64, in the order written below (the kernel uses __dmul_rn / __dadd_rn to match bit for bit).
"""
from __future__ import annotations

import math

import numpy as np


def image_ranges(cell: np.ndarray, pbc, cutoff: float):
    a = cell.astype(np.float64)
    cr = np.empty((3, 3))
    for d in range(3):
        e, f = (d + 1) % 3, (d + 2) % 3
        cr[d, 0] = a[e, 1] * a[f, 2] - a[e, 2] * a[f, 1]
        cr[d, 1] = a[e, 2] * a[f, 0] - a[e, 0] * a[f, 2]
        cr[d, 2] = a[e, 0] * a[f, 1] - a[e, 1] * a[f, 0]
    vol = abs((a[0, 0] * cr[0, 0] + a[0, 1] * cr[0, 1]) + a[0, 2] * cr[0, 2])
    n = []
    for d in range(3):
        area = math.sqrt((cr[d, 0] * cr[d, 0] + cr[d, 1] * cr[d, 1]) + cr[d, 2] * cr[d, 2])
        n.append(int(math.ceil(cutoff / (vol / area))) if (pbc[d] and vol > 0 and area > 0) else 0)
    return n


def _sorted_neighbours(pos: np.ndarray, cell: np.ndarray, pbc, cutoff: float):
    """per centre atom: list of (d, j, S0, S1, S2) within the cutoff, in the canonical order"""
    n_at = pos.shape[0]
    p = pos.astype(np.float64)
    a = cell.astype(np.float64)
    n0, n1, n2 = image_ranges(cell, pbc, cutoff)
    S = np.array([(s0, s1, s2) for s0 in range(-n0, n0 + 1) for s1 in range(-n1, n1 + 1) for s2 in range(-n2, n2 + 1)],
                 dtype=np.float64)
    # sh[img, k] = (S0*a0k + S1*a1k) + S2*a2k
    sh = (S[:, 0:1] * a[0][None, :] + S[:, 1:2] * a[1][None, :]) + S[:, 2:3] * a[2][None, :]
    zero_img = np.all(S == 0, axis=1)
    out = []
    for i in range(n_at):
        dx = p - p[i][None, :]                       # (n,3)  pos_j - pos_i
        v = dx[:, None, :] + sh[None, :, :]          # (n, img, 3)
        d = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])
        keep = d < cutoff
        keep[i, zero_img] = False
        jj, im = np.nonzero(keep)
        out.append(sorted(zip(d[jj, im].tolist(), jj.tolist(), S[im, 0].tolist(), S[im, 1].tolist(), S[im, 2].tolist())))
    return out


def raw_counts(pos: np.ndarray, cell: np.ndarray, pbc, cutoff: float) -> np.ndarray:
    """number of (j, image) within the cutoff of every atom, before truncation / fallback"""
    return np.asarray([len(x) for x in _sorted_neighbours(pos, cell, pbc, cutoff)], dtype=np.int64)


def neighbor_list(pos: np.ndarray, cell: np.ndarray, pbc, cutoff: float, max_neighbors: int):
    """(src, dst, shift) of one structure: pos (n,3) float32, cell (3,3) float32 rows = lattice vectors."""
    n_at = pos.shape[0]
    src, dst, shf = [], [], []
    for i, key in enumerate(_sorted_neighbours(pos, cell, pbc, cutoff)):
        key = key[:max_neighbors]
        src += [i] * len(key)
        dst += [k[1] for k in key]
        shf += [[k[2], k[3], k[4]] for k in key]
    if not src:  # convert.py:154-157 -> data/utils.py:10-20
        src = [i for i in range(n_at) for j in range(n_at) if j != i]
        dst = [j for i in range(n_at) for j in range(n_at) if j != i]
        shf = [[0.0, 0.0, 0.0]] * len(src)
    return (np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64),
            np.asarray(shf, dtype=np.float32).reshape(-1, 3))


def batch_neighbor_list(pos, graph_ptr, lattice, pbc, cutoff, max_neighbors):
    """concatenation over structures with global atom ids, plus the per-structure edge counts"""
    srcs, dsts, shfs, counts = [], [], [], []
    for g in range(len(graph_ptr) - 1):
        lo, hi = int(graph_ptr[g]), int(graph_ptr[g + 1])
        s, t, sh = neighbor_list(np.asarray(pos[lo:hi]), np.asarray(lattice[g]), [bool(x) for x in pbc[g]], cutoff, max_neighbors)
        srcs.append(s + lo), dsts.append(t + lo), shfs.append(sh), counts.append(len(s))
    return np.concatenate(srcs), np.concatenate(dsts), np.concatenate(shfs).reshape(-1, 3), np.asarray(counts)
