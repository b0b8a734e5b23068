"""torch.autograd Functions over the C ABI (include/lcao_b200.h).


PyTorch is used for device memory, the current stream and autograd bookkeeping only: every
edge-sized operation of the hot path (forward and backward) is a call into liblcao_b200.so.
All functions require CUDA tensors and raise `LcaoError` otherwise — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor
from torch.autograd.function import once_differentiable

from . import _lib, second_order
from ._lib import ACT_NONE, ACT_SILU, LcaoError, call, ptr, require_cuda, stream_ptr

# GEMM arithmetic mode for the dense layers: fp32 CUDA cores, 3xTF32 tcgen05 (fp32-equivalent), 1xTF32
GEMM_MODES = {"fp32": _lib.GEMM_FP32, "tf32x3": _lib.GEMM_TF32X3, "tf32": _lib.GEMM_TF32}
# The default is the tcgen05 path in its FP32-equivalent arithmetic (3xTF32 operand split, FP32 accumulation in TMEM):
# it is the mode the parity tests and the benchmark run.  "fp32" selects the CUDA-core kernels (gemm_simt.cu).
_gemm_mode = _lib.GEMM_TF32X3

# launch counter: every C-ABI compute call increments it (bench.py reports it as gpu_launches)
n_calls = 0


def set_gemm_mode(mode: str) -> None:
    global _gemm_mode
    _gemm_mode = GEMM_MODES[mode]


def get_gemm_mode() -> str:
    return {v: k for k, v in GEMM_MODES.items()}[_gemm_mode]


class positions_only:
    """Context for `torch.autograd.grad(energy, pos)` (autograd forces, reference lcaonet.py:310-317): inside it the
    backward kernels skip every PARAMETER gradient — the engine would discard them anyway — and never touch the
    in-place gradient sinks, which belong to the loss's own backward pass.  A plain module-level flag: autograd runs
    CUDA backward nodes on its device thread, so a thread-local would not be seen there."""
    active = False

    def __enter__(self):
        self._prev = positions_only.active
        positions_only.active = True

    def __exit__(self, *exc):
        positions_only.active = self._prev
        return False


def validate_graph(z: Tensor, max_z: int, batch: Tensor | None, n_graph: int, edge_index: Tensor) -> None:
    """Raise IndexError when z, batch or edge_index would index out of range (the reference raises the same from
    nn.Embedding / index_select: embed.py:41,91, base.py:38, lcaonet.py:462).  One small kernel + one host sync."""
    require_cuda(z, edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_index must be an int64 tensor of shape (2, E)")
    if z.dtype != torch.int64 or (batch is not None and batch.dtype != torch.int64):
        raise ValueError("z and batch must be int64 tensors")
    status = torch.empty(1, dtype=torch.int32, device=z.device)
    ei = edge_index.contiguous()
    _call("lcao_validate_graph", ptr(z.contiguous()), z.numel(), int(max_z), ptr(batch.contiguous() if batch is not None else None),
          int(n_graph), ptr(ei), ei.shape[1], ptr(status), stream_ptr())
    bad = int(status.item())
    if bad:
        what = [m for b, m in ((1, f"atomic numbers outside [1, {max_z}]"), (2, f"batch indices outside [0, {n_graph})"),
                               (4, f"edge_index entries outside [0, {z.numel()})")) if bad & b]
        raise IndexError("index out of range in the input batch: " + "; ".join(what))


def _call(name, *args):
    global n_calls
    n_calls += 1
    call(name, *args)


def _rows(x: Tensor) -> Tensor:
    """2-D view with unit inner stride (copy only if needed)."""
    if x.dim() != 2:
        x = x.reshape(-1, x.shape[-1])
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def _ld(x: Tensor) -> int:
    return x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1)


# ------------------------------------------------------------------------------------------------
# graph indices
# ------------------------------------------------------------------------------------------------
class GraphIndex:
    """int32 CSR views of `edge_index` built on the GPU (bit-exact, deterministic):
    in-CSR by target (ordered by source, then edge id), out-CSR by source, triplet offsets.
    Replaces torch_sparse.SparseTensor in `LCAONet.get_triplets` (reference lcaonet.py:439-486)."""

    def __init__(self, edge_index: Tensor, n_nodes: int):
        require_cuda(edge_index)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must be an int64 tensor of shape (2, E)")
        ei = edge_index.contiguous()
        E, N, dev = ei.shape[1], int(n_nodes), ei.device
        self.E, self.N, self.edge_index = E, N, ei
        i32 = dict(dtype=torch.int32, device=dev)
        self.src32, self.dst32 = torch.empty(E, **i32), torch.empty(E, **i32)
        self.in_ptr, self.out_ptr = torch.empty(N + 1, **i32), torch.empty(N + 1, **i32)
        self.in_edge, self.in_src, self.out_edge = torch.empty(E, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
        self._tri_ptr = None  # triplet offsets: only the reference's triplet LISTS need them -> built on demand
        scratch = torch.empty(2 * N + 2 * E + 8, **i32)
        _call("lcao_graph_index_build", ptr(ei), E, N, ptr(self.src32), ptr(self.dst32), ptr(self.in_ptr),
              ptr(self.in_edge), ptr(self.in_src), ptr(self.out_ptr), ptr(self.out_edge), None, ptr(scratch), stream_ptr())

    @property
    def tri_ptr(self) -> Tensor:
        if self._tri_ptr is None:
            i32 = dict(dtype=torch.int32, device=self.src32.device)
            self._tri_ptr = torch.empty(self.E + 1, **i32)
            scratch = torch.empty(self.E + self.E // 4096 + 2, **i32)
            _call("lcao_triplet_offsets", ptr(self.src32), ptr(self.dst32), ptr(self.in_ptr), self.E, ptr(self._tri_ptr),
                  ptr(scratch), stream_ptr())
        return self._tri_ptr

    def num_triplets(self) -> int:
        return int(self.tri_ptr[-1].item())  # host sync: T is data dependent

    def triplets(self, unit: Tensor | None = None):
        """(tri_idx_k, edge_idx_ks, edge_idx_st[, cos]) int64 lists in the reference's order."""
        T, dev = self.num_triplets(), self.src32.device
        k, e_ks, e_st = (torch.empty(T, dtype=torch.int64, device=dev) for _ in range(3))
        cos = torch.empty(T, dtype=torch.float32, device=dev) if unit is not None else None
        _call("lcao_triplets_fill", ptr(self.src32), ptr(self.in_ptr), ptr(self.in_edge), ptr(self.tri_ptr), self.E,
              ptr(k), ptr(e_ks), ptr(e_st), ptr(unit), ptr(cos), stream_ptr())
        return (k, e_ks, e_st) if unit is None else (k, e_ks, e_st, cos)


def bucket_sort(keys: Tensor, n_buckets: int, sec: Tensor | None = None, stable: bool | str = True):
    """(ptr int32 (nb+1), perm int32 (n)): items grouped by key; stable=True orders each bucket by
    (sec, id) (small buckets only), stable="ordered" lists the items of a bucket in ascending id (few huge buckets:
    species / species-pair keys; reproducible), stable=False just groups (order inside a bucket unspecified)."""
    require_cuda(keys)
    keys = keys.contiguous()
    n, dev = keys.numel(), keys.device
    p = torch.empty(n_buckets + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    mode = 2 if stable == "ordered" else (1 if stable else 0)
    scratch = torch.empty(n_buckets + n + 1 + (n_buckets * ((n + 1023) // 1024) if mode == 2 else 0), dtype=torch.int32, device=dev)
    _call("lcao_bucket_sort", ptr(keys), ptr(sec.contiguous() if sec is not None else None), n, n_buckets, ptr(p),
          ptr(perm), ptr(scratch), mode, stream_ptr())
    return p, perm


def histogram(keys: Tensor, n_buckets: int) -> Tensor:
    require_cuda(keys)
    out = torch.empty(n_buckets, dtype=torch.float32, device=keys.device)
    _call("lcao_histogram", ptr(keys.contiguous()), keys.numel(), n_buckets, ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------
# geometry + radial basis
# ------------------------------------------------------------------------------------------------
class _GeomBasis(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, shift, lattice, batch, gi: GraphIndex, spec, n_orb: int):
        require_cuda(pos, shift, lattice)
        pos_c, shift_c, lat_c = pos.contiguous().float(), shift.contiguous().float(), lattice.contiguous().float()
        E, dev = gi.E, pos.device
        dist = torch.empty(E, device=dev)
        unit = torch.empty(E, 3, device=dev)
        rb = torch.empty(E, n_orb, device=dev)
        need_grad = ctx.needs_input_grad[0]
        drb = torch.empty(E, n_orb, device=dev) if need_grad else None
        _call("lcao_geom_basis_fwd", ptr(pos_c), ptr(shift_c), ptr(lat_c), ptr(batch), ptr(gi.src32), ptr(gi.dst32), E,
              ctypes.byref(spec), ptr(dist), ptr(unit), ptr(rb), ptr(drb), stream_ptr())
        ctx.gi, ctx.n_orb, ctx.spec = gi, n_orb, spec
        ctx.save_for_backward(dist, unit, drb, pos, shift, lattice, batch)
        return dist, unit, rb

    @staticmethod
    def backward(ctx, d_dist, d_unit, d_rb):
        dist, unit, drb, pos, shift, lattice, batch = ctx.saved_tensors
        gi = ctx.gi
        if drb is None:
            return (None,) * 7
        if torch.is_grad_enabled():  # create_graph=True: this backward is itself differentiated (training on autograd forces)
            with torch.enable_grad():
                (pos_,) = second_order.aliases([pos])
                outs = second_order.geom_basis(pos_, shift, lattice, batch, gi.src32, gi.dst32, ctx.spec, ctx.n_orb)
                (d_pos,) = second_order.grads_with_graph(outs, [pos_], [d_dist, d_unit, d_rb], [True])
            return d_pos, None, None, None, None, None, None
        dev = dist.device
        dvec = torch.empty(gi.E, 3, device=dev)
        d_pos = torch.empty(gi.N, 3, device=dev)
        c = lambda t: t.contiguous() if t is not None else None  # noqa: E731
        d_dist, d_unit, d_rb = c(d_dist), c(d_unit), c(d_rb)
        _call("lcao_geom_basis_bwd", ptr(dist), ptr(unit), ptr(drb), ptr(d_dist), ptr(d_unit), ptr(d_rb), gi.E, gi.N,
              ctx.n_orb, ptr(gi.in_ptr), ptr(gi.in_edge), ptr(gi.out_ptr), ptr(gi.out_edge), ptr(dvec), ptr(d_pos),
              stream_ptr())
        return d_pos, None, None, None, None, None, None


def geom_basis(pos, shift, lattice, batch, gi, spec, n_orb):
    return _GeomBasis.apply(pos, shift, lattice, batch, gi, spec, n_orb)


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act: int):
        require_cuda(x, weight)
        shape = x.shape
        x2 = _rows(x)
        w = weight.contiguous()
        M, K, Nout = x2.shape[0], x2.shape[1], w.shape[0]
        y = torch.empty(M, Nout, device=x.device)
        need_pre = act != ACT_NONE and any(ctx.needs_input_grad)
        pre = torch.empty(M, Nout, device=x.device) if need_pre else None
        _call("lcao_linear_fwd", ptr(x2), _ld(x2), ptr(w), ptr(bias), ptr(y), Nout, ptr(pre), Nout, M, K, Nout, act,
              _gemm_mode, stream_ptr())
        ctx.act, ctx.shape, ctx.has_bias = act, shape, bias is not None
        ctx.save_for_backward(x2, w, pre, x, weight, bias)
        return y.reshape(*shape[:-1], Nout)

    @staticmethod
    def backward(ctx, dy):
        x2, w, pre, x_in, w_in, b_in = ctx.saved_tensors
        if torch.is_grad_enabled():  # create_graph=True (see second_order.py)
            with torch.enable_grad():
                ins = second_order.aliases([x_in, w_in, b_in])
                y = second_order.linear(*ins, ctx.act)
                wanted = list(ctx.needs_input_grad[:3])
                if positions_only.active:
                    wanted[1] = wanted[2] = False
                gx, gw, gb = second_order.grads_with_graph([y], ins, [dy], wanted)
            return gx, gw, gb, None
        M, K, Nout = x2.shape[0], x2.shape[1], w.shape[0]
        dy2 = _rows(dy)
        st = stream_ptr()
        need_dx = ctx.needs_input_grad[0]
        need_dw = (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])) and not positions_only.active
        if ctx.act != ACT_NONE:  # dH = dY * act'(pre), shared by the data and the weight gradient
            dh = torch.empty(M, Nout, device=dy.device)
            _call("lcao_act_bwd", ptr(dy2), _ld(dy2), ptr(pre), Nout, ptr(dh), Nout, M, Nout, ctx.act, st)
            dy2 = dh
        dx = dw = db = None
        if need_dx:
            dx = torch.empty(M, K, device=dy.device)
            _call("lcao_linear_dgrad", ptr(dy2), _ld(dy2), None, 0, ACT_NONE, ptr(w), ptr(dx), K, M, K, Nout, 0,
                  _gemm_mode, None, st)
            dx = dx.reshape(ctx.shape)
        if need_dw:
            dw = torch.zeros(Nout, K, device=dy.device)
            db = torch.zeros(Nout, device=dy.device) if ctx.has_bias else None
            n_scr = int(_lib.load().lcao_linear_bwd_scratch(ptr(dy2), _ld(dy2), None, 0, ACT_NONE, None, ptr(x2), _ld(x2),
                                                            None, 0, M, K, Nout, _gemm_mode))
            scr = torch.empty(n_scr, device=dy.device) if n_scr else None
            _call("lcao_linear_wgrad", ptr(dy2), _ld(dy2), None, 0, ACT_NONE, ptr(x2), _ld(x2), ptr(dw), ptr(db), M,
                  K, Nout, _gemm_mode, ptr(scr), st)
        return dx, dw, db, None


def _act_arg(act) -> int:
    """activation argument of the public ops: an LCAO_ACT_* code, or True / False for SiLU / none"""
    return ACT_SILU if act is True else ACT_NONE if act is False or act is None else int(act)


def linear(x: Tensor, weight: Tensor, bias: Tensor | None = None, silu=False) -> Tensor:
    """act(x W^T + b) — nn.Linear semantics (reference nn/base.py:72-81) with an optional fused SiLU.
    Requires the feature sizes to be multiples of 4 only for the vector paths; any size works."""
    return _Linear.apply(x, weight, bias, _act_arg(silu))


# ------------------------------------------------------------------------------------------------
# orbital contraction, three-body, two-body
# ------------------------------------------------------------------------------------------------
class _CoeffContract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cst1, rb, vmask, lgrp, NL: int, C: int):
        require_cuda(cst1, rb)
        cst1 = cst1.contiguous()
        E, O, Cp = cst1.shape
        valence = 1 if vmask is not None else 0
        assert Cp == C * (1 + valence)
        NG = NL + valence
        B = torch.empty(E, NG, C, device=cst1.device)
        _call("lcao_coeff_contract_fwd", ptr(cst1), ptr(rb), ptr(vmask), ptr(lgrp), E, O, C, NL, valence, ptr(B),
              stream_ptr())
        ctx.dims = (E, O, C, NL, valence)
        ctx.save_for_backward(cst1, rb, vmask, lgrp)
        return B

    @staticmethod
    @once_differentiable
    def backward(ctx, dB):
        cst1, rb, vmask, lgrp = ctx.saved_tensors
        E, O, C, NL, valence = ctx.dims
        dB = dB.contiguous()
        d_cst1 = torch.empty_like(cst1)
        d_rb = torch.empty_like(rb) if ctx.needs_input_grad[1] else None
        _call("lcao_coeff_contract_bwd", ptr(cst1), ptr(rb), ptr(vmask), ptr(lgrp), ptr(dB), E, O, C, NL, valence,
              ptr(d_cst1), ptr(d_rb), stream_ptr())
        return d_cst1, d_rb, None, None, None, None


def coeff_contract(cst1, rb, vmask, lgrp, NL, C):
    """B[e,l,:] = sum_{o in l} rb[e,o] (A[e,o,:] + m V[e,o,:]) (+ valence slot) — the orbital sums of
    lcaonet.py:180-183 and :200-203, grouped by angular momentum so both consumers share them."""
    return _CoeffContract.apply(cst1, rb.contiguous(), vmask, lgrp, NL, C)


class _PairContract(torch.autograd.Function):
    """B and its Gram matrices from the species-pair coefficient table (see lcao_pair_contract_fwd)."""

    @staticmethod
    def forward(ctx, tab, pair, grouping, rb, vmask, lgrp, NL: int, C: int):
        require_cuda(tab, pair, rb)
        tab = tab.contiguous()
        P, O, Cp = tab.shape
        E = pair.numel()
        valence = 1 if vmask is not None else 0
        assert Cp == C * (1 + valence) and pair.dtype == torch.int64
        NG = NL + valence
        B = torch.empty(E, NG, C, device=tab.device)
        gram = torch.empty(E, NL * (NL + 1) // 2, dtype=torch.float64, device=tab.device)
        _call("lcao_pair_contract_fwd", ptr(tab), ptr(pair), ptr(rb), ptr(vmask), ptr(lgrp), E, O, C, NL, valence, ptr(B),
              ptr(gram), None, stream_ptr())
        ctx.dims, ctx.grouping = (E, P, O, C, NL, valence), grouping
        ctx.save_for_backward(tab, pair, rb, vmask, lgrp)
        ctx.mark_non_differentiable(gram)
        return B, gram

    @staticmethod
    @once_differentiable
    def backward(ctx, dB, _dgram):
        tab, pair, rb, vmask, lgrp = ctx.saved_tensors
        E, P, O, C, NL, valence = ctx.dims
        kptr, kperm = _grouping(ctx.grouping, pair, P)
        dB = dB.contiguous()
        d_tab = torch.empty_like(tab)
        d_rb = torch.empty_like(rb) if ctx.needs_input_grad[3] else None
        nbytes = int(_lib.load().lcao_pair_contract_bwd_scratch(E, P, O, C, valence))
        scratch = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.int32, device=tab.device)
        _call("lcao_pair_contract_bwd", ptr(tab), ptr(pair), ptr(kptr), ptr(kperm), ptr(rb), ptr(vmask), ptr(lgrp), ptr(dB),
              E, P, O, C, NL, valence, ptr(d_tab), ptr(d_rb), ptr(scratch), stream_ptr())
        return d_tab, None, None, d_rb, None, None, None, None


def _grouping(grouping, pair, P):
    """(kptr, kperm) from whatever the caller supplied: a callable that builds them on demand (PairCoeffs.grouping), a
    ready (kptr, kperm) tuple, or None (built here)."""
    if callable(grouping):
        return grouping()
    if grouping is not None and grouping[0] is not None:
        return grouping
    with torch.no_grad():
        return bucket_sort(pair, P, stable="ordered")


def pair_contract(tab, pair, grouping, rb, vmask, lgrp, NL, C):
    """(B, gram): B[e,l,:] = sum_{o in l} rb[e,o] tab[pair[e],o,:] (+ valence slot) — the orbital sums of
    lcaonet.py:180-183 and :200-203 with f_coeffs (lcaonet.py:170) evaluated on the species-pair table."""
    return _PairContract.apply(tab, pair, grouping, rb.contiguous(), vmask, lgrp, NL, C)


def coeff_gram(B, NL):
    """(E, NL(NL+1)/2) FP64 Gram matrices B[e,l,:].B[e,l',:] of the first NL groups (no autograd: the
    three-body backward differentiates through them itself)."""
    require_cuda(B)
    E, NG, C = B.shape
    gram = torch.empty(E, NL * (NL + 1) // 2, dtype=torch.float64, device=B.device)
    _call("lcao_coeff_gram", ptr(B.detach().contiguous()), NG, E, C, NL, ptr(gram), stream_ptr())
    return gram


class BodyLink:
    """Couples the backward passes of `threebody` and `twobody` on the same B (lcaonet.py:173-204).

    Both consume B, so autograd would add two (E, NG, C) gradients in a separate pass.  The two-body
    gradient is the SAME row for every l < NL, so with a link the two-body backward leaves its compact
    (E, 1 + valence, C) gradient here and the three-body backward kernel folds it into the dB it writes.
    Works whatever order autograd runs the two nodes in: if the three-body backward has already run,
    the two-body one returns its gradient the ordinary way."""

    def __init__(self):
        self.dP = None
        self.consumed = False


class _ThreeBody(torch.autograd.Function):
    @staticmethod
    def forward(ctx, B, gram, unit, xk, gi: GraphIndex, NL: int, link):
        require_cuda(B, unit, xk)
        B, unit = B.contiguous(), unit.contiguous()
        E, NG, C = B.shape
        assert xk.stride(1) == 1 and gram.dtype == torch.float64 and gram.is_contiguous()
        st = stream_ptr()
        gate = torch.empty(gi.N, C, device=B.device)  # sigmoid(xk) once per node
        _call("lcao_sigmoid_rows", ptr(xk), xk.stride(0), ptr(gate), C, gi.N, C, st)
        tbw = torch.empty(E, C, device=B.device)
        _call("lcao_threebody_fwd", ptr(B), NG, ptr(gram), ptr(unit), ptr(gate), C, ptr(gi.in_ptr),
              ptr(gi.in_edge), ptr(gi.in_src), ptr(gi.out_ptr), ptr(gi.out_edge), gi.N, E, C, NL, ptr(tbw), st)
        ctx.gi, ctx.NL, ctx.link = gi, NL, link
        ctx.save_for_backward(B, gram, unit, gate)
        return tbw

    @staticmethod
    @once_differentiable
    def backward(ctx, d_tbw):
        B, gram, unit, gate = ctx.saved_tensors
        gi, NL, link = ctx.gi, ctx.NL, ctx.link
        E, NG, C = B.shape
        d_tbw = d_tbw.contiguous()
        dB = torch.empty_like(B)
        q = torch.empty(E, C, device=B.device)
        st = stream_ptr()
        forces = ctx.needs_input_grad[2]
        du_ks = torch.empty(E, 3, device=B.device) if forces else None
        du_st = torch.empty(E, 3, device=B.device) if forces else None
        dP = None
        if link is not None:
            dP, link.dP, link.consumed = link.dP, None, True
        _call("lcao_threebody_bwd", ptr(B), NG, ptr(gram), ptr(unit), ptr(gate), C, ptr(gi.in_ptr),
              ptr(gi.in_edge), ptr(gi.in_src), ptr(gi.out_ptr), ptr(gi.out_edge), gi.N, E, C, NL, ptr(d_tbw), ptr(dP),
              ptr(dB), ptr(q), ptr(du_ks), ptr(du_st), st)
        d_xk = torch.empty(gi.N, C, device=B.device)
        _call("lcao_segment_sum", ptr(q), C, None, 0, ptr(gi.out_ptr), ptr(gi.out_edge), gi.N, C, 0, ptr(d_xk), C, st)
        return dB, None, (du_ks + du_st) if forces else None, d_xk, None, None, None


def threebody(B, unit, xk, gi, NL, gram=None, link: BodyLink | None = None):
    """Fused gather / angular basis / normalise / gate / triplet->edge sum (lcaonet.py:173-189)."""
    if gram is None:
        gram = coeff_gram(B, NL)
    return _ThreeBody.apply(B, gram, unit, xk, gi, NL, link)


class _TwoBody(torch.autograd.Function):
    @staticmethod
    def forward(ctx, B, g, NL: int, valence: int, link):
        require_cuda(B, g)
        E, NG, C = B.shape
        g = g.contiguous()
        lw = torch.empty(E, C, device=B.device)
        _call("lcao_twobody_fwd", ptr(B), NG, ptr(g), E, C, NL, valence, ptr(lw), stream_ptr())
        ctx.dims, ctx.link = (NL, valence), link
        ctx.save_for_backward(B, g)
        return lw

    @staticmethod
    @once_differentiable
    def backward(ctx, d_lw):
        B, g = ctx.saved_tensors
        NL, valence = ctx.dims
        E, NG, C = B.shape
        link = ctx.link
        compact = link is not None and not link.consumed
        dB = torch.empty(E, 1 + valence, C, device=B.device) if compact else torch.empty_like(B)
        dg = torch.empty_like(g)
        _call("lcao_twobody_bwd", ptr(B), NG, ptr(g), ptr(d_lw.contiguous()), E, C, NL, valence, 1 if compact else 0,
              ptr(dB), ptr(dg), stream_ptr())
        if compact:  # handed to the three-body backward kernel, which adds it to the dB it writes
            link.dP = dB
            return None, dg, None, None, None
        return dB, dg, None, None, None


def twobody(B, g, NL, valence, link: BodyLink | None = None):
    """lw = normalize((1+g_A) P_A + (1+g_V) P_V) with P = sum_l B_l (lcaonet.py:192-204)."""
    return _TwoBody.apply(B, g, NL, valence, link)


# ------------------------------------------------------------------------------------------------
# one interaction block as ONE autograd node
# ------------------------------------------------------------------------------------------------
def _lin_fwd(x, ldx, M, w, bias, act, st, need_pre):
    """y (M, Nout) [, pre] = act(x W^T + b) on raw row views; returns (y, pre)."""
    Nout, K = w.shape
    y = torch.empty(M, Nout, device=w.device)
    pre = torch.empty(M, Nout, device=w.device) if (need_pre and act != ACT_NONE) else None
    _call("lcao_linear_fwd", ptr(x), ldx, ptr(w), ptr(bias), ptr(y), Nout, ptr(pre), Nout, M, K, Nout, act, _gemm_mode, st)
    return y, pre


def _lin_dgrad(dy, ldy, M, w, dx, ldx, accumulate, st):
    Nout, K = w.shape
    _call("lcao_linear_dgrad", ptr(dy), ldy, None, 0, ACT_NONE, ptr(w), ptr(dx), ldx, M, K, Nout, accumulate, _gemm_mode, None, st)


def _lin_dgrad_act(dy, ldy, M, w, pre_in, dx, ldx, st, act=ACT_SILU):
    """dx = (dy W) * act'(pre_in): data gradient chained through the activation that produced the layer's input."""
    Nout, K = w.shape
    _call("lcao_linear_dgrad_act", ptr(dy), ldy, ptr(w), ptr(pre_in), K, act, ptr(dx), ldx, M, K, Nout, _gemm_mode, st)


# Weight gradients that go straight into the parameters' .grad buffers (gradient sinks) are only read by the optimizer /
# the all-reduce, so the second stage of the tcgen05 weight-gradient kernel (the sum of its per-CTA partial tiles) is
# deferred: the descriptors pile up here and ONE batched launch at the end of the backward pass (an autograd engine
# callback) finishes them all.  `defer_wgrad = False` restores one reduction launch per layer.
defer_wgrad = True
_pending_wgrad = {"desc": [], "keep": [], "queued": False, "stream": None}


def flush_wgrad():
    """Finish the deferred weight gradients (runs by itself at the end of every backward pass that deferred any)."""
    p = _pending_wgrad
    p["queued"] = False
    if p["desc"]:
        import ctypes as C
        arr = (C.c_int64 * len(p["desc"]))(*p["desc"])
        _call("lcao_wgrad_reduce_batch", arr, len(p["desc"]) // 6, p["stream"])
    p["desc"], p["keep"], p["stream"] = [], [], None


def _lin_wgrad(dy, ldy, x, ldx, M, w, has_bias, st, dw_sink=None, db_sink=None):
    """dW (+ db) of one layer.  With sinks (the parameters' .grad buffers) the kernels accumulate straight into them
    — the C ABI's `dW +=` contract — and (None, None) is returned: no zero-fill, no separate accumulation pass."""
    Nout, K = w.shape
    dw = dw_sink if dw_sink is not None else torch.zeros(Nout, K, device=w.device)
    db = (db_sink if db_sink is not None else torch.zeros(Nout, device=w.device)) if has_bias else None
    n_scr = int(_lib.load().lcao_linear_bwd_scratch(ptr(dy), ldy, None, 0, ACT_NONE, None, ptr(x), ldx, None, 0, M, K, Nout,
                                                    _gemm_mode))
    if defer_wgrad and dw_sink is not None and (db_sink is not None or not has_bias) and n_scr:
        import ctypes as C
        passes = (Nout + 127) // 128
        scr = torch.empty(n_scr * passes, device=w.device)
        desc, nd = (C.c_int64 * (6 * passes))(), C.c_int32(0)
        _call("lcao_linear_wgrad_deferred", ptr(dy), ldy, ptr(x), ldx, ptr(dw), ptr(db), M, K, Nout, _gemm_mode, ptr(scr), desc,
              C.byref(nd), st)
        if nd.value:
            p = _pending_wgrad
            p["desc"].extend(desc[: 6 * nd.value])
            p["keep"].append((scr, dw, db))  # the partial tiles and their targets stay alive until the flush
            p["stream"] = st
            if not p["queued"]:
                torch.autograd.Variable._execution_engine.queue_callback(flush_wgrad)
                p["queued"] = True
        return None, None
    scr = torch.empty(n_scr, device=w.device) if n_scr else None
    _call("lcao_linear_wgrad", ptr(dy), ldy, None, 0, ACT_NONE, ptr(x), ldx, ptr(dw), ptr(db), M, K, Nout, _gemm_mode,
          ptr(scr), st)
    return (None if dw_sink is not None else dw), (None if (db_sink is not None or not has_bias) else db)


def _act_bwd(dy, pre, M, C, st, act=ACT_SILU):
    out = torch.empty(M, C, device=dy.device)
    _call("lcao_act_bwd", ptr(dy), C, ptr(pre), C, ptr(out), C, M, C, act, st)
    return out


class _InteractionLayer(torch.autograd.Function):
    """LCAOInteraction.forward (reference lcaonet.py:130-216) as a single autograd node.

    Same kernels, same order and same arithmetic as the op-by-op path (`coeff_contract/pair_contract`,
    `threebody`, `twobody`, `edge_pair`, `mul_segment_sum`, `linear`), but the ~15 forward and ~30
    backward C-ABI calls are issued from one Python frame each: no per-op autograd bookkeeping, no
    gradient-accumulation passes between consumers of the same tensor, strided halves written in place.
    The host cost of a layer drops from ~1.5 ms to ~0.3 ms, which matters once the GPU side of a
    1024-molecule step is ~10 ms."""

    @staticmethod
    def forward(ctx, x, table, rb, unit, w_n, b_n, w_c0, w_c2, w_3, w_b, w_1, b_1, w_2, b_2, w_o, aux):
        pair, grouping, vmask, lgrp, gi, NL, C, sinks, act = aux
        require_cuda(x, table, rb, unit)
        st = stream_ptr()
        dev = x.device
        N, H = x.shape
        E = gi.E
        P, O, K = table.shape
        valence = 1 if vmask is not None else 0
        Cp, NG = C * (1 + valence), NL + valence
        grad = any(ctx.needs_input_grad)  # (grad mode is always off inside Function.forward)
        x, table, rb, unit = x.contiguous(), table.contiguous(), rb.contiguous(), unit.contiguous()
        # node side: nw = [xc | xk]
        nw, _ = _lin_fwd(x, H, N, w_n, b_n, ACT_NONE, st, False)
        # f_coeffs on the species-pair table
        t1, pre1 = _lin_fwd(table, K, P * O, w_c0, None, act, st, grad)
        tab, pre2 = _lin_fwd(t1, C, P * O, w_c2, None, act, st, grad)
        B = torch.empty(E, NG, C, device=dev)
        gram = torch.empty(E, NL * (NL + 1) // 2, dtype=torch.float64, device=dev)
        psum = torch.empty(E, 1 + valence, C, device=dev)  # [sum_l B_l | valence slot]: the two-body weight's view of B
        _call("lcao_pair_contract_fwd", ptr(tab), ptr(pair), ptr(rb), ptr(vmask), ptr(lgrp), E, O, C, NL, valence, ptr(B),
              ptr(gram), ptr(psum), st)
        # three-body
        gate = torch.empty(N, C, device=dev)
        _call("lcao_sigmoid_rows", nw.data_ptr() + 4 * C, 2 * C, ptr(gate), C, N, C, st)
        tbw = torch.empty(E, C, device=dev)
        _call("lcao_threebody_fwd", ptr(B), NG, ptr(gram), ptr(unit), ptr(gate), C, ptr(gi.in_ptr), ptr(gi.in_edge),
              ptr(gi.in_src), ptr(gi.out_ptr), ptr(gi.out_edge), N, E, C, NL, ptr(tbw), st)
        g, _ = _lin_fwd(tbw, C, E, w_3, None, ACT_NONE, st, False)
        # two-body weight and message
        lw = torch.empty(E, C, device=dev)
        _call("lcao_twobody_fwd", ptr(psum), 1 + valence, ptr(g), E, C, 1, valence, ptr(lw), st)
        bw, _ = _lin_fwd(lw, C, E, w_b, None, ACT_NONE, st, False)
        w_1cat = torch.cat([w_1[:, :C], w_1[:, C:]], dim=0)  # (2C, C): [W1a ; W1b] acting on x_s / x_t per NODE
        u, _ = _lin_fwd(nw, 2 * C, N, w_1cat, None, ACT_NONE, st, False)  # reads xc = nw[:, :C]
        a1 = torch.empty(E, C, device=dev)
        pre_a = torch.empty(E, C, device=dev) if grad else None
        _call("lcao_edge_pair_fwd", ptr(u), 2 * C, u.data_ptr() + 4 * C, 2 * C, ptr(b_1), ptr(gi.src32), ptr(gi.dst32), E, C,
              act, ptr(a1), ptr(pre_a), st)
        # h = SiLU(pre_h) is never stored: the message sum and its backward apply the activation on the fly, which takes
        # one E x C write off the f_node GEMM and one E x C read off each consumer
        pre_h, _ = _lin_fwd(a1, C, E, w_2, b_2, ACT_NONE, st, False)
        agg = torch.empty(N, C, device=dev)
        _call("lcao_segment_sum", ptr(bw), C, ptr(pre_h), C, ptr(gi.out_ptr), ptr(gi.out_edge), N, C, 2 | (act << 4), ptr(agg), C,
              st)
        y, _ = _lin_fwd(agg, C, N, w_o, None, ACT_NONE, st, False)
        out = x + y
        if grad:
            ctx.aux = aux
            ctx.save_for_backward(x, table, rb, unit, w_n, w_c0, w_c2, w_3, w_b, w_1cat, w_2, w_o, nw, t1, pre1, tab, pre2,
                                  B, gram, gate, tbw, g, lw, bw, a1, pre_a, pre_h, agg, psum, b_n, w_1, b_1, b_2)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x, table, rb, unit, w_n, w_c0, w_c2, w_3, w_b, w_1cat, w_2, w_o, nw, t1, pre1, tab, pre2, B, gram, gate, tbw, g,
         lw, bw, a1, pre_a, pre_h, agg, psum, b_n, w_1, b_1, b_2) = ctx.saved_tensors
        pair, grouping, vmask, lgrp, gi, NL, C, sinks, act = ctx.aux
        if torch.is_grad_enabled():
            # create_graph=True: this backward pass is itself being differentiated (training on autograd forces,
            # lcaonet.py:310-317).  The operator is re-evaluated in its any-order differentiable form (second_order.py) on
            # the saved INPUTS, and autograd returns their gradients together with the graph behind them.
            inputs = second_order.aliases([x, table, rb, unit, w_n, b_n, w_c0, w_c2, w_3, w_b, w_1, b_1, w_2, b_2, w_o])
            wanted = list(ctx.needs_input_grad[:15])
            if positions_only.active:  # only the path to the positions: x, rb, unit
                wanted = [wanted[0], False, wanted[2], wanted[3]] + [False] * 11
            with torch.enable_grad():
                out = second_order.interaction(*inputs, pair, vmask, lgrp, gi, NL, C, act)
                grads = second_order.grads_with_graph([out], inputs, [d_out], wanted)
            return (*grads, None)
        # `positions_only` (autograd forces): no parameter gradient is wanted from this pass, and the sinks are not its to write
        wg = not positions_only.active
        if not wg:
            sinks = None
        # gradient sinks: the parameters' own .grad buffers (order: w_n, b_n, w_c0, w_c2, w_3, w_b, w_1, b_1, w_2, b_2, w_o)
        s_wn, s_bn, s_wc0, s_wc2, s_w3, s_wb, s_w1, s_b1, s_w2, s_b2, s_wo = sinks if sinks is not None else (None,) * 11
        st = stream_ptr()
        dev = x.device
        N, H = x.shape
        E = gi.E
        P, O, K = table.shape
        valence = 1 if vmask is not None else 0
        Cp, NG = C * (1 + valence), NL + valence
        need_rb, need_unit = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        d_out = d_out.contiguous()
        dw_n = db_n = dw_c0 = dw_c2 = dw_3 = dw_b = dw_1 = db_1 = dw_2 = db_2 = dw_o = d_table = None
        # x_out = x + out_weight(agg)
        if wg:
            dw_o, _ = _lin_wgrad(d_out, H, agg, C, N, w_o, False, st, s_wo)
        d_agg = torch.empty(N, C, device=dev)
        _lin_dgrad(d_out, H, N, w_o, d_agg, C, 0, st)
        # agg[s] = sum_{e in out(s)} bw[e] * h[e]
        d_bw, d_preh = torch.empty(E, C, device=dev), torch.empty(E, C, device=dev)
        _call("lcao_msg_bwd", ptr(d_agg), C, ptr(gi.src32), None, ptr(bw), ptr(pre_h), E, C, act, ptr(d_bw), ptr(d_preh), st)
        # h = silu(f_node.2(a1)) ; a1 = silu(u_a[s] + u_b[t] + b1)
        if wg:
            dw_2, db_2 = _lin_wgrad(d_preh, C, a1, C, E, w_2, True, st, s_w2, s_b2)
        # d_prea = d_a1 * SiLU'(pre_a) is never materialised: the two segment sums over it apply the factor on the fly
        d_a1 = torch.empty(E, C, device=dev)
        _lin_dgrad(d_preh, C, E, w_2, d_a1, C, 0, st)
        d_u = torch.empty(N, 2 * C, device=dev)
        _call("lcao_segment_sum", ptr(d_a1), C, ptr(pre_a), C, ptr(gi.out_ptr), ptr(gi.out_edge), N, C, 4 | (act << 4), ptr(d_u),
              2 * C, st)
        _call("lcao_segment_sum", ptr(d_a1), C, ptr(pre_a), C, ptr(gi.in_ptr), ptr(gi.in_edge), N, C, 4 | (act << 4),
              d_u.data_ptr() + 4 * C, 2 * C, st)
        d_nw = torch.empty(N, 2 * C, device=dev)
        _lin_dgrad(d_u, 2 * C, N, w_1cat, d_nw, 2 * C, 0, st)  # writes d_xc = d_nw[:, :C]
        if wg:
            # the bias b_1 enters once per edge through the u_a half: d b_1 = column sums of d_u[:, :C]
            dw_1cat, db_cat = _lin_wgrad(d_u, 2 * C, nw, 2 * C, N, w_1cat, True, st)
            dw_1, db_1 = torch.cat([dw_1cat[:C], dw_1cat[C:]], dim=1), db_cat[:C]
            if s_w1 is not None:
                s_w1.add_(dw_1)
                dw_1 = None
            if s_b1 is not None:
                s_b1.add_(db_1)
                db_1 = None
        # bw = basis_weight(lw) ; lw = twobody(B, g) ; g = f_three(tbw) ; tbw = threebody(B, ...)
        if wg:
            dw_b, _ = _lin_wgrad(d_bw, C, lw, C, E, w_b, False, st, s_wb)
        d_lw = d_preh  # reuse
        _lin_dgrad(d_bw, C, E, w_b, d_lw, C, 0, st)
        dP = torch.empty(E, 1 + valence, C, device=dev)
        d_g = torch.empty(E, Cp, device=dev)
        _call("lcao_twobody_bwd", ptr(psum), 1 + valence, ptr(g), ptr(d_lw), E, C, 1, valence, 1, ptr(dP), ptr(d_g), st)
        if wg:
            dw_3, _ = _lin_wgrad(d_g, Cp, tbw, C, E, w_3, False, st, s_w3)
        d_tbw = d_bw  # reuse
        _lin_dgrad(d_g, Cp, E, w_3, d_tbw, C, 0, st)
        dB = torch.empty(E, NG, C, device=dev)
        q = d_a1  # reuse
        du_ks = torch.empty(E, 3, device=dev) if need_unit else None
        du_st = torch.empty(E, 3, device=dev) if need_unit else None
        _call("lcao_threebody_bwd", ptr(B), NG, ptr(gram), ptr(unit), ptr(gate), C, ptr(gi.in_ptr), ptr(gi.in_edge),
              ptr(gi.in_src), ptr(gi.out_ptr), ptr(gi.out_edge), N, E, C, NL, ptr(d_tbw), ptr(dP), ptr(dB), ptr(q),
              ptr(du_ks), ptr(du_st), st)
        _call("lcao_segment_sum", ptr(q), C, None, 0, ptr(gi.out_ptr), ptr(gi.out_edge), N, C, 0, d_nw.data_ptr() + 4 * C,
              2 * C, st)  # d_xk = d_nw[:, C:]
        # B = pair_contract(tab, rb) ; tab = f_coeffs(table)
        need_tab = wg and (ctx.needs_input_grad[1] or any(ctx.needs_input_grad[6:8]))
        d_rb = torch.empty(E, O, device=dev) if need_rb else None
        if need_tab or need_rb:
            kptr, kperm = _grouping(grouping, pair, P)
            d_tab = torch.empty(P, O, Cp, device=dev) if need_tab else None
            scratch = None
            if need_tab:
                nbytes = int(_lib.load().lcao_pair_contract_bwd_scratch(E, P, O, C, valence))
                scratch = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.int32, device=dev)
            _call("lcao_pair_contract_bwd", ptr(tab), ptr(pair), ptr(kptr), ptr(kperm), ptr(rb), ptr(vmask), ptr(lgrp), ptr(dB), E,
                  P, O, C, NL, valence, ptr(d_tab), ptr(d_rb), ptr(scratch), st)
        if need_tab:
            d_pre2 = _act_bwd(d_tab, pre2, P * O, Cp, st, act)
            dw_c2, _ = _lin_wgrad(d_pre2, Cp, t1, C, P * O, w_c2, False, st, s_wc2)
            d_pre1 = torch.empty(P * O, C, device=dev)
            _lin_dgrad_act(d_pre2, Cp, P * O, w_c2, pre1, d_pre1, C, st, act)
            dw_c0, _ = _lin_wgrad(d_pre1, C, table, K, P * O, w_c0, False, st, s_wc0)
            if ctx.needs_input_grad[1]:
                d_table = torch.empty(P, O, K, device=dev)
                _lin_dgrad(d_pre1, C, P * O, w_c0, d_table, K, 0, st)
        # nw = node_weight(x) ; residual
        if wg:
            dw_n, db_n = _lin_wgrad(d_nw, 2 * C, x, H, N, w_n, True, st, s_wn, s_bn)
        dx = d_out.clone()
        _lin_dgrad(d_nw, 2 * C, N, w_n, dx, H, 1, st)
        d_unit = (du_ks + du_st) if need_unit else None
        return dx, d_table, d_rb, d_unit, dw_n, db_n, dw_c0, dw_c2, dw_3, dw_b, dw_1, db_1, dw_2, db_2, dw_o, None


def interaction_layer(x, table, rb, unit, w_n, b_n, w_c0, w_c2, w_3, w_b, w_1, b_1, w_2, b_2, w_o, pair, grouping, vmask,
                      lgrp, gi, NL, C, grad_sinks=None, act=ACT_SILU):
    """grad_sinks: optional 11-tuple of the weights' / biases' `.grad` buffers (contiguous, same shapes; None entries
    allowed).  The backward pass then ACCUMULATES those gradients in place and returns None for them to autograd —
    what AccumulateGrad would do, without the zero-fills and the per-parameter add kernels.  `grouping`: callable
    returning (kptr, kperm), a ready tuple, or None (see `_grouping`).  Opt-in
    (`LCAOInteraction.grads_in_place`, set by `dist.FlatGradBucket`): parameter hooks do not fire for them."""
    return _InteractionLayer.apply(x, table, rb, unit, w_n, b_n, w_c0, w_c2, w_3, w_b, w_1, b_1, w_2, b_2, w_o,
                                   (pair, grouping, vmask, lgrp, gi, NL, C, grad_sinks, _act_arg(act)))


# ------------------------------------------------------------------------------------------------
# count-weighted BatchNorm over table rows (embedding block)
# ------------------------------------------------------------------------------------------------
class _TableNorm(torch.autograd.Function):
    """nn.BatchNorm1d over a batch given as DISTINCT rows + multiplicities (embed.py:175,194,232,249)."""

    @staticmethod
    def forward(ctx, x, counts, weight, bias, running_mean, running_var, tracked, training: bool, momentum: float,
                eps: float):
        require_cuda(x)
        x = x.contiguous()
        R, F = x.shape
        y = torch.empty_like(x)
        mean, rstd = torch.empty(F, device=x.device), torch.empty(F, device=x.device)
        _call("lcao_table_norm_fwd", ptr(x), ptr(counts), ptr(weight), ptr(bias), R, F, eps, momentum, 1 if training else 0,
              ptr(running_mean), ptr(running_var), ptr(tracked), ptr(y), ptr(mean), ptr(rstd), stream_ptr())
        ctx.training = training
        ctx.save_for_backward(x, counts, weight, mean, rstd)
        ctx.mark_non_differentiable(*[t for t in (running_mean, running_var, tracked) if t is not None])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, counts, weight, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        R, F = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty(F, device=x.device) if weight is not None else None
        db = torch.empty(F, device=x.device) if weight is not None else None
        _call("lcao_table_norm_bwd", ptr(dy), ptr(x), ptr(counts), ptr(weight), ptr(mean), ptr(rstd), R, F,
              1 if ctx.training else 0, ptr(dx), ptr(dg), ptr(db), stream_ptr())
        return dx, None, dg, db, None, None, None, None, None, None


def table_norm(x, counts, weight, bias, running_mean, running_var, tracked, training, momentum, eps):
    return _TableNorm.apply(x, counts, weight, bias, running_mean, running_var, tracked, training, momentum, eps)


class _PairOuter(torch.autograd.Function):
    """pre[s,t,o,k] = fe[t,o,k] * (1 + za[s,k] + zb[t,k]): the species-pair coefficient table before its BatchNorm
    (embed.py:234-249 on the pair table) as one kernel forward and one backward."""

    @staticmethod
    def forward(ctx, fe, za, zb):
        require_cuda(fe, za, zb)
        fe, za, zb = fe.contiguous(), za.contiguous(), zb.contiguous()
        Zd, O, K = fe.shape
        pre = torch.empty(Zd, Zd, O, K, device=fe.device)
        _call("lcao_pair_outer_fwd", ptr(fe), ptr(za), ptr(zb), Zd, O, K, ptr(pre), stream_ptr())
        ctx.save_for_backward(fe, za, zb)
        return pre

    @staticmethod
    @once_differentiable
    def backward(ctx, dpre):
        fe, za, zb = ctx.saved_tensors
        Zd, O, K = fe.shape
        dpre = dpre.contiguous()
        d_fe, d_za, d_zb = torch.empty_like(fe), torch.empty_like(za), torch.empty_like(zb)
        _call("lcao_pair_outer_bwd", ptr(dpre), ptr(fe), ptr(za), ptr(zb), Zd, O, K, ptr(d_fe), ptr(d_za), ptr(d_zb), stream_ptr())
        return d_fe, d_za, d_zb


def pair_outer(fe, za, zb):
    return _PairOuter.apply(fe, za, zb)


# ------------------------------------------------------------------------------------------------
# gathers / segment sums
# ------------------------------------------------------------------------------------------------
class _EdgePair(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, bias, gi: GraphIndex, act: int):
        require_cuda(a, b)
        assert a.stride(1) == 1 and b.stride(1) == 1
        C = a.shape[1]
        out = torch.empty(gi.E, C, device=a.device)
        need_pre = act != ACT_NONE and any(ctx.needs_input_grad)
        pre = torch.empty(gi.E, C, device=a.device) if need_pre else None
        _call("lcao_edge_pair_fwd", ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(bias), ptr(gi.src32), ptr(gi.dst32),
              gi.E, C, act, ptr(out), ptr(pre), stream_ptr())
        ctx.gi, ctx.act, ctx.has_bias = gi, act, bias is not None
        ctx.save_for_backward(pre)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out):
        (pre,) = ctx.saved_tensors
        gi = ctx.gi
        d_out = d_out.contiguous()
        E, C = d_out.shape
        st = stream_ptr()
        if ctx.act != ACT_NONE:
            dh = torch.empty_like(d_out)
            _call("lcao_act_bwd", ptr(d_out), C, ptr(pre), C, ptr(dh), C, E, C, ctx.act, st)
            d_out = dh
        da = torch.empty(gi.N, C, device=d_out.device)
        db = torch.empty(gi.N, C, device=d_out.device)
        _call("lcao_segment_sum", ptr(d_out), C, None, 0, ptr(gi.out_ptr), ptr(gi.out_edge), gi.N, C, 0, ptr(da), C, st)
        _call("lcao_segment_sum", ptr(d_out), C, None, 0, ptr(gi.in_ptr), ptr(gi.in_edge), gi.N, C, 0, ptr(db), C, st)
        dbias = da.sum(0) if ctx.has_bias else None
        return da, db, dbias, None, None


def edge_pair(a, b, bias, gi, silu):
    """act(a[s_e] + b[t_e] + bias): `W [x_s ; x_t] + b` with W split per node (lcaonet.py:209, :304)."""
    return _EdgePair.apply(a, b, bias, gi, _act_arg(silu))


class _MulSegSum(torch.autograd.Function):
    """agg[s,:] = sum_{e in out(s)} x[e,:] * y[e,:]   (torch_scatter site lcaonet.py:207-214)."""

    @staticmethod
    def forward(ctx, x, y, gi: GraphIndex):
        require_cuda(x, y)
        x, y = x.contiguous(), y.contiguous()
        C = x.shape[1]
        out = torch.empty(gi.N, C, device=x.device)
        _call("lcao_segment_sum", ptr(x), C, ptr(y), C, ptr(gi.out_ptr), ptr(gi.out_edge), gi.N, C, 0, ptr(out), C,
              stream_ptr())
        ctx.gi = gi
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out):
        x, y = ctx.saved_tensors
        gi = ctx.gi
        d_out = d_out.contiguous()
        E, C = x.shape
        st = stream_ptr()
        dx, dy = torch.empty_like(x), torch.empty_like(y)
        _call("lcao_gather_rows", ptr(d_out), C, ptr(gi.src32), 0, ptr(y), C, E, C, ptr(dx), C, st)
        _call("lcao_gather_rows", ptr(d_out), C, ptr(gi.src32), 0, ptr(x), C, E, C, ptr(dy), C, st)
        return dx, dy, None


def mul_segment_sum(x, y, gi):
    return _MulSegSum.apply(x, y, gi)


class _SegmentReduce(torch.autograd.Function):
    """out[r] = sum (or mean) of x over the items of segment r; items listed by (ptr, perm), and
    `seg_of_item` (int64 or int32, n) maps items back to segments for the backward gather."""

    @staticmethod
    def forward(ctx, x, seg_ptr, seg_perm, seg_of_item, mean: bool):
        require_cuda(x)
        x2 = _rows(x)
        R, C = seg_ptr.numel() - 1, x2.shape[1]
        out = torch.empty(R, C, device=x.device)
        _call("lcao_segment_sum", ptr(x2), _ld(x2), None, 0, ptr(seg_ptr), ptr(seg_perm), R, C, 1 if mean else 0,
              ptr(out), C, stream_ptr())
        ctx.mean, ctx.n = mean, x2.shape[0]
        ctx.save_for_backward(seg_ptr, seg_of_item, x)
        return out

    @staticmethod
    def backward(ctx, d_out):
        seg_ptr, seg_of_item, x_in = ctx.saved_tensors
        if torch.is_grad_enabled():  # create_graph=True (see second_order.py)
            with torch.enable_grad():
                (x_,) = second_order.aliases([x_in])
                out = second_order.segment_reduce(x_, seg_of_item, seg_ptr.numel() - 1, ctx.mean)
                (gx,) = second_order.grads_with_graph([out], [x_], [d_out], [True])
            return gx, None, None, None, None
        d_out = d_out.contiguous()
        if ctx.mean:
            cnt = (seg_ptr[1:] - seg_ptr[:-1]).clamp(min=1).to(d_out.dtype)
            d_out = d_out / cnt.unsqueeze(-1)
        C = d_out.shape[1]
        dx = torch.empty(ctx.n, C, device=d_out.device)
        _call("lcao_gather_rows", ptr(d_out), C, ptr(seg_of_item), 1 if seg_of_item.dtype == torch.int64 else 0, None,
              0, ctx.n, C, ptr(dx), C, stream_ptr())
        return dx, None, None, None, None


def segment_reduce(x, seg_ptr, seg_perm, seg_of_item, mean=False):
    return _SegmentReduce.apply(x, seg_ptr, seg_perm, seg_of_item, mean)


class _GatherRows(torch.autograd.Function):
    """out[i,:] = table[idx[i],:]; backward is a keyed reduction over the (ptr, perm) grouping of idx."""

    @staticmethod
    def forward(ctx, table, idx, key_ptr, key_perm):
        require_cuda(table, idx)
        t2 = table.reshape(table.shape[0], -1)
        assert t2.stride(1) == 1
        n, W = idx.numel(), t2.shape[1]
        out = torch.empty(n, W, device=table.device)
        _call("lcao_gather_rows", ptr(t2), t2.stride(0), ptr(idx), 1 if idx.dtype == torch.int64 else 0, None, 0, n, W,
              ptr(out), W, stream_ptr())
        ctx.tshape = table.shape
        ctx.save_for_backward(key_ptr, key_perm)
        return out.reshape(n, *table.shape[1:])

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out):
        key_ptr, key_perm = ctx.saved_tensors
        d2 = d_out.reshape(d_out.shape[0], -1).contiguous()
        n, W = d2.shape
        nkeys = ctx.tshape[0]
        acc = torch.zeros(nkeys, W, device=d_out.device)
        _call("lcao_reduce_by_key", ptr(d2), W, ptr(key_ptr), ptr(key_perm), nkeys, n, W, ptr(acc), stream_ptr())
        return acc.reshape(ctx.tshape), None, None, None


def gather_rows(table, idx, key_ptr, key_perm):
    return _GatherRows.apply(table, idx, key_ptr, key_perm)
