"""CPU emulation of the C ABI (include/lcao_b200.h) — TEST INFRASTRUCTURE ONLY.

Each entry point is re-stated with plain torch ops acting on the raw pointers the host code passes,
using the SAME decomposition and the same hand-derived backward formulas as the CUDA kernels.
`install()` monkeypatches `lcaonet_b200._lib.call` so that the product's host logic (ops.py,
model.py: argument order, strides, autograd wiring, the embedding tables, weighted BatchNorm, ...)
can be exercised end-to-end against the oracle on a machine without a GPU.  It is never importable
from the product package and is slow by design.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

F32, I32, I64, F64 = torch.float32, torch.int32, torch.int64, torch.float64
_SZ = {F32: 4, I32: 4, I64: 8, F64: 8}


def view(ptr, rows, cols=None, dtype=F32, ld=None):
    """Tensor view of raw memory: (rows,) or (rows, cols) with row stride ld."""
    if ptr is None or ptr == 0:
        return None
    if cols is None:
        n = rows
        if n == 0:
            return torch.empty(0, dtype=dtype)
        buf = (C.c_char * (n * _SZ[dtype])).from_address(ptr)
        return torch.frombuffer(buf, dtype=dtype, count=n)
    ld = cols if ld is None else ld
    if rows == 0 or cols == 0:
        return torch.empty(rows, cols, dtype=dtype)
    n = (rows - 1) * ld + cols
    buf = (C.c_char * (n * _SZ[dtype])).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype, count=n).as_strided((rows, cols), (ld, 1))


def _bucket_sort(keys, sec, nb):
    n = keys.numel()
    comp = keys.to(I64)
    if sec is not None:
        order = torch.argsort(sec.to(I64), stable=True)
        order = order[torch.argsort(comp[order], stable=True)]
    else:
        order = torch.argsort(comp, stable=True)
    cnt = torch.bincount(comp, minlength=nb) if n else torch.zeros(nb, dtype=I64)
    ptr = torch.zeros(nb + 1, dtype=I64)
    ptr[1:] = cnt.cumsum(0)
    return ptr.to(I32), order.to(I32)


def lcao_bucket_sort(keys, sec, n, nb, ptr, perm, scratch, stable, stream):
    k = view(keys, n, dtype=I64)
    s = view(sec, n, dtype=I64)
    p, o = _bucket_sort(k if n else torch.empty(0, dtype=I64), s, nb)
    view(ptr, nb + 1, dtype=I32).copy_(p)
    if n:
        view(perm, n, dtype=I32).copy_(o)


def lcao_validate_graph(z, N, max_z, batch, n_graph, ei, E, status, stream):
    bad = 0
    zz, bb = view(z, N, dtype=I64), view(batch, N, dtype=I64)
    if zz is not None and N and bool(((zz < 1) | (zz > max_z)).any()):
        bad |= 1
    if bb is not None and N and bool(((bb < 0) | (bb >= n_graph)).any()):
        bad |= 2
    if E:
        e = view(ei, 2 * E, dtype=I64)
        if bool(((e < 0) | (e >= N)).any()):
            bad |= 4
    view(status, 1, dtype=I32).fill_(bad)


def lcao_graph_index_build(ei, E, N, src32, dst32, in_ptr, in_edge, in_src, out_ptr, out_edge, tri_ptr, scratch, stream):
    e = view(ei, 2, E, dtype=I64) if E else torch.empty(2, 0, dtype=I64)
    s, t = e[0], e[1]
    ip, ie = _bucket_sort(t, s, N)
    op, oe = _bucket_sort(s, None, N)
    view(in_ptr, N + 1, dtype=I32).copy_(ip)
    view(out_ptr, N + 1, dtype=I32).copy_(op)
    tp = torch.zeros(E + 1, dtype=I64)
    if E:
        view(src32, E, dtype=I32).copy_(s.to(I32))
        view(dst32, E, dtype=I32).copy_(t.to(I32))
        view(in_edge, E, dtype=I32).copy_(ie)
        view(in_src, E, dtype=I32).copy_(s[ie.long()].to(I32))
        view(out_edge, E, dtype=I32).copy_(oe)
        indeg = (ip[1:] - ip[:-1]).long()
        tp[1:] = (indeg[s] - (s == t).long()).cumsum(0)
    if tri_ptr:
        view(tri_ptr, E + 1, dtype=I32).copy_(tp.to(I32))


def lcao_triplet_offsets(src32, dst32, in_ptr, E, tri_ptr, scratch, stream):
    tp = torch.zeros(E + 1, dtype=I64)
    if E:
        s, t = view(src32, E, dtype=I32).long(), view(dst32, E, dtype=I32).long()
        N = int(max(s.max(), t.max())) + 1
        ip = view(in_ptr, N + 1, dtype=I32).long()
        tp[1:] = ((ip[1:] - ip[:-1])[s] - (s == t).long()).cumsum(0)
    view(tri_ptr, E + 1, dtype=I32).copy_(tp.to(I32))


def _pairs(in_ptr, in_edge, out_ptr, out_edge, N):
    """all (e, e') with e in out(s), e' in in(s), e' != e, ordered by out position then in position."""
    es, eps = [], []
    for s in range(N):
        o = out_edge[out_ptr[s]:out_ptr[s + 1]].long()
        i = in_edge[in_ptr[s]:in_ptr[s + 1]].long()
        if o.numel() and i.numel():
            es.append(o.repeat_interleave(i.numel()))
            eps.append(i.repeat(o.numel()))
    if not es:
        return torch.empty(0, dtype=I64), torch.empty(0, dtype=I64)
    e, ep = torch.cat(es), torch.cat(eps)
    keep = e != ep
    return e[keep], ep[keep]


def lcao_triplets_fill(src32, in_ptr, in_edge, tri_ptr, E, tri_k, e_ks, e_st, unit, cos_out, stream):
    tp = view(tri_ptr, E + 1, dtype=I32).long()
    T = int(tp[-1])
    src = view(src32, E, dtype=I32).long()
    N = int(src.max()) + 1 if E else 0
    ie = view(in_edge, E, dtype=I32).long()
    # N is not passed: recover the in_ptr length from the largest source index is unsafe -> read lazily
    ks, eks, est = [], [], []
    for e in range(E):
        s = int(src[e])
        lo, hi = int(view(in_ptr + 4 * s, 2, dtype=I32)[0]), int(view(in_ptr + 4 * s, 2, dtype=I32)[1])
        lst = ie[lo:hi]
        lst = lst[lst != e]
        eks.append(lst)
        est.append(torch.full_like(lst, e))
    eks = torch.cat(eks) if eks else torch.empty(0, dtype=I64)
    est = torch.cat(est) if est else torch.empty(0, dtype=I64)
    assert eks.numel() == T
    if T:
        view(tri_k, T, dtype=I64).copy_(src[eks])
        view(e_ks, T, dtype=I64).copy_(eks)
        view(e_st, T, dtype=I64).copy_(est)
        if cos_out:
            u = view(unit, E, 3)
            view(cos_out, T).copy_((u[est] * u[eks]).sum(-1))


def _nl_inputs(pos, batch, graph_ptr, lattice, pbc, N):
    b = view(batch, N, dtype=I64)
    B = int(b.max()) + 1
    return (view(pos, N, 3).numpy(), view(graph_ptr, B + 1, dtype=I32).numpy(), view(lattice, B * 3, 3).reshape(B, 3, 3).numpy(),
            view(pbc, B, 3, dtype=I32).numpy(), B)


def lcao_neighbor_count(pos, batch, graph_ptr, lattice, pbc, N, cutoff, count, stream):
    from oracle import neighbor_oracle as NO
    P, gp, lat, pb, B = _nl_inputs(pos, batch, graph_ptr, lattice, pbc, N)
    out = view(count, N, dtype=I32)
    for g in range(B):
        lo, hi = int(gp[g]), int(gp[g + 1])
        out[lo:hi] = torch.from_numpy(NO.raw_counts(P[lo:hi], lat[g], [bool(x) for x in pb[g]], cutoff)).to(I32)


def lcao_neighbor_fill(pos, batch, graph_ptr, lattice, pbc, fallback, out_ptr, N, E, cutoff, max_nb, edge_index, edge_shift,
                       status, stream):
    from oracle import neighbor_oracle as NO
    P, gp, lat, pb, B = _nl_inputs(pos, batch, graph_ptr, lattice, pbc, N)
    s, t, sh, _ = NO.batch_neighbor_list(P, gp, lat, pb, cutoff, max_nb)
    assert len(s) == E
    ei = view(edge_index, 2, E, dtype=I64)
    ei[0], ei[1] = torch.from_numpy(s), torch.from_numpy(t)
    view(edge_shift, E, 3).copy_(torch.from_numpy(sh))


def lcao_histogram(keys, n, nb, counts, stream):
    k = view(keys, n, dtype=I64)
    view(counts, nb).copy_(torch.bincount(k, minlength=nb).float() if n else torch.zeros(nb))


def _cutoff(kind, r, rc):
    q = r / rc
    u = 1 - q
    if kind == 0:
        f, df = u**3 * (1 + 3 * q + 6 * q * q), -30 * q * q * u * u / rc
    elif kind == 1:
        f, df = u**3 * (1 + 3 * q + 6 * q**2 + 10 * q**3 + 15 * q**4), -105 * q**4 * u * u / rc
    else:
        a = math.pi / rc
        f, df = 0.5 * (torch.cos(a * r) + 1), -0.5 * a * torch.sin(a * r)
    inside = r <= rc
    return torch.where(inside, f, torch.zeros_like(f)), torch.where(inside, df, torch.zeros_like(df))


def lcao_geom_basis_fwd(pos, shift, lattice, batch, src32, dst32, E, spec_ref, dist, unit, rb, drb, stream):
    sp = spec_ref._obj
    s, t = view(src32, E, dtype=I32).long(), view(dst32, E, dtype=I32).long()
    N = int(max(s.max(), t.max())) + 1
    P = view(pos, N, 3).double()
    b = view(batch, N, dtype=I64)[s] if batch else torch.zeros(E, dtype=I64)
    nb = int(b.max()) + 1
    L = view(lattice, nb * 3, 3).double().reshape(nb, 3, 3)[b]
    v = (P[t] - P[s]) + (view(shift, E, 3).double().unsqueeze(-1) * L).sum(1)
    r = v.norm(dim=1)
    view(dist, E).copy_(r.float())
    view(unit, E, 3).copy_((v / r.unsqueeze(-1)).float())
    fc, dfc = _cutoff(sp.cutoff_kind, r, sp.rc)
    O = sp.n_unique * sp.n_rep
    vals, dvals = [], []
    for u in range(sp.n_unique):
        n, l = sp.n[u], sp.l[u]
        if sp.rbf_kind == 0:
            zs = 2.0 / (n * sp.a0)
            zeta = zs * r
            p, dp = torch.zeros_like(r), torch.zeros_like(r)
            for i in range(sp.deg[u], -1, -1):
                dp = dp * zeta + p
                p = p * zeta + sp.poly[u][i]
            zl = zeta**l
            zlm1 = l * zeta ** (l - 1) if l > 0 else torch.zeros_like(r)
            ex = torch.exp(-0.5 * zeta)
            R = sp.norm[u] * p * zl * ex
            dR = sp.norm[u] * ex * (dp * zl + p * zlm1 - 0.5 * p * zl) * zs
        else:
            w = math.pi * n / sp.rc
            R = torch.sin(w * r) / r
            dR = w * torch.cos(w * r) / r - torch.sin(w * r) / (r * r)
        for _ in range(sp.n_rep):
            vals.append(fc * R)
            dvals.append(dfc * R + fc * dR)
    view(rb, E, O).copy_(torch.stack(vals, 1).float())
    if drb:
        view(drb, E, O).copy_(torch.stack(dvals, 1).float())


def lcao_geom_basis_bwd(dist, unit, drb, d_dist, d_unit, d_rb, E, N, O, in_ptr, in_edge, out_ptr, out_edge, dvec, d_pos, stream):
    u, r = view(unit, E, 3), view(dist, E)
    gr = view(d_dist, E).clone() if d_dist else torch.zeros(E)
    if d_rb:
        gr = gr + (view(d_rb, E, O) * view(drb, E, O)).sum(1)
    g = gr.unsqueeze(1) * u
    if d_unit:
        a = view(d_unit, E, 3)
        g = g + (a - u * (a * u).sum(1, keepdim=True)) / r.unsqueeze(1)
    view(dvec, E, 3).copy_(g)
    ip, ie = view(in_ptr, N + 1, dtype=I32).long(), view(in_edge, E, dtype=I32).long()
    op, oe = view(out_ptr, N + 1, dtype=I32).long(), view(out_edge, E, dtype=I32).long()
    out = torch.zeros(N, 3)
    for n in range(N):
        out[n] = g[ie[ip[n]:ip[n + 1]]].sum(0) - g[oe[op[n]:op[n + 1]]].sum(0)
    view(d_pos, N, 3).copy_(out)


def _groups(lgrp, O, NL):
    l = view(lgrp, O, dtype=I32).long()
    return torch.nn.functional.one_hot(l, NL).float()  # (O, NL)


def lcao_coeff_contract_fwd(cst1, rb, vmask, lgrp, E, O, C, NL, valence, B, stream):
    Cp = C * (1 + valence)
    c = view(cst1, E, O * Cp).reshape(E, O, Cp)
    r = view(rb, E, O)
    G = _groups(lgrp, O, NL)
    t = r.unsqueeze(-1) * c[..., :C]
    NG = NL + valence
    out = view(B, E, NG * C).reshape(E, NG, C)
    if valence:
        v = (r * view(vmask, E, O)).unsqueeze(-1) * c[..., C:]
        t = t + v
        out[:, NL] = v.sum(1)
    out[:, :NL] = torch.einsum("eoc,ol->elc", t, G)


def lcao_coeff_contract_bwd(cst1, rb, vmask, lgrp, dB, E, O, C, NL, valence, d_cst1, d_rb, stream):
    Cp = C * (1 + valence)
    NG = NL + valence
    r = view(rb, E, O)
    G = _groups(lgrp, O, NL)
    d = view(dB, E, NG * C).reshape(E, NG, C)
    dl = torch.einsum("elc,ol->eoc", d[:, :NL], G)  # dB_{l(o)}
    out = view(d_cst1, E, O * Cp).reshape(E, O, Cp)
    out[..., :C] = r.unsqueeze(-1) * dl
    if valence:
        m = view(vmask, E, O)
        dv = dl + d[:, NL].unsqueeze(1)
        out[..., C:] = (r * m).unsqueeze(-1) * dv
    if d_rb:
        c = view(cst1, E, O * Cp).reshape(E, O, Cp)
        g = (c[..., :C] * dl).sum(-1)
        if valence:
            g = g + m * (c[..., C:] * dv).sum(-1)
        view(d_rb, E, O).copy_(g)


def _sph(c, NL):
    y = [torch.full_like(c, 0.28209479177387814), 0.4886025119029199 * c,
         0.9461746957575601 * c * c - 0.31539156525252005, 0.3731763325901154 * c * (5 * c * c - 3)]
    return torch.stack(y[:NL], 1)  # (T, NL)


def _pair_rows(tab, pair, E, P_rows, O, Cp):
    pr = view(pair, E, dtype=I64)
    return pr, view(tab, P_rows, O * Cp).reshape(P_rows, O, Cp)


def _gram(Bt, NL):
    cols = [(Bt[:, a].double() * Bt[:, b].double()).sum(-1) for a in range(NL) for b in range(a, NL)]
    return torch.stack(cols, 1)


def lcao_pair_contract_fwd(tab, pair, rb, vmask, lgrp, E, O, C, NL, valence, B, gram, psum, stream):
    Cp = C * (1 + valence)
    pr = view(pair, E, dtype=I64)
    T = view(tab, int(pr.max()) + 1, O * Cp).reshape(-1, O, Cp)
    c = T[pr]
    r = view(rb, E, O)
    G = _groups(lgrp, O, NL)
    t = r.unsqueeze(-1) * c[..., :C]
    NG = NL + valence
    out = view(B, E, NG * C).reshape(E, NG, C)
    if valence:
        v = (r * view(vmask, E, O)).unsqueeze(-1) * c[..., C:]
        t = t + v
        out[:, NL] = v.sum(1)
    out[:, :NL] = torch.einsum("eoc,ol->elc", t, G)
    if gram:
        view(gram, E, NL * (NL + 1) // 2, dtype=torch.float64).copy_(_gram(out, NL))
    if psum:
        ps = view(psum, E, (1 + valence) * C).reshape(E, 1 + valence, C)
        ps[:, 0] = out[:, :NL].sum(1)
        if valence:
            ps[:, 1] = out[:, NL]


def lcao_coeff_gram(B, NG, E, C, NL, gram, stream):
    Bt = view(B, E, NG * C).reshape(E, NG, C)
    view(gram, E, NL * (NL + 1) // 2, dtype=torch.float64).copy_(_gram(Bt, NL))


def lcao_pair_contract_bwd(tab, pair, kptr, kperm, rb, vmask, lgrp, dB, E, P, O, C, NL, valence, d_tab, d_rb, scratch, stream):
    Cp = C * (1 + valence)
    NG = NL + valence
    kp = view(kptr, P + 1, dtype=I32).long()
    pm = view(kperm, E, dtype=I32).long() if E else torch.empty(0, dtype=I64)
    key_of = torch.empty(E, dtype=I64)
    key_of[pm] = torch.repeat_interleave(torch.arange(P), kp[1:] - kp[:-1])
    r = view(rb, E, O)
    G = _groups(lgrp, O, NL)
    d = view(dB, E, NG * C).reshape(E, NG, C)
    dl = torch.einsum("elc,ol->eoc", d[:, :NL], G)
    contrib = torch.zeros(E, O, Cp)
    contrib[..., :C] = r.unsqueeze(-1) * dl
    if valence:
        m = view(vmask, E, O)
        dv = dl + d[:, NL].unsqueeze(1)
        contrib[..., C:] = (r * m).unsqueeze(-1) * dv
    if d_tab:
        out = view(d_tab, P, O * Cp)
        out.zero_()
        out.index_add_(0, key_of, contrib.reshape(E, O * Cp))
    if d_rb:
        pr = view(pair, E, dtype=I64)
        c = view(tab, P, O * Cp).reshape(P, O, Cp)[pr]
        g = (c[..., :C] * dl).sum(-1)
        if valence:
            g = g + m * (c[..., C:] * dv).sum(-1)
        view(d_rb, E, O).copy_(g)


def lcao_sigmoid_rows(x, ldx, out, ldo, M, C, stream):
    view(out, M, C, ld=ldo).copy_(torch.sigmoid(view(x, M, C, ld=ldx)))


def _sph_grad(c, NL):
    y = [torch.zeros_like(c), torch.full_like(c, 0.4886025119029199), 2 * 0.9461746957575601 * c,
         0.3731763325901154 * (15 * c * c - 3)]
    return torch.stack(y[:NL], 1)


def _tb_common(B, NG, unit, xk, ldxk, in_ptr, in_edge, in_src, out_ptr, out_edge, N, E, C, NL):
    Bt = view(B, E, NG * C).reshape(E, NG, C)
    u = view(unit, E, 3)
    X = view(xk, N, C, ld=ldxk)
    ip, ie = view(in_ptr, N + 1, dtype=I32).long(), view(in_edge, E, dtype=I32).long()
    op, oe = view(out_ptr, N + 1, dtype=I32).long(), view(out_edge, E, dtype=I32).long()
    src = torch.empty(E, dtype=I64)
    src[ie] = view(in_src, E, dtype=I32).long()
    e, ep = _pairs(ip, ie, op, oe, N)
    cos = (u[e] * u[ep]).sum(-1)
    Y = _sph(cos, NL)
    v = torch.einsum("tl,tlc->tc", Y, Bt[ep][:, :NL])
    nrm = v.norm(dim=1, keepdim=True)
    sg = X[src[ep]]  # the gate rows are sigmoid(xk) already (lcao_sigmoid_rows)
    return Bt, e, ep, Y, v, nrm, sg, src, cos, u


def lcao_threebody_fwd(B, NG, gram, unit, xk, ldxk, in_ptr, in_edge, in_src, out_ptr, out_edge, N, E, C, NL, tbw, stream):
    Bt, e, ep, Y, v, nrm, sg, src, _, _ = _tb_common(B, NG, unit, xk, ldxk, in_ptr, in_edge, in_src, out_ptr, out_edge, N, E, C, NL)
    y = v / nrm.clamp(min=1e-12)
    view(tbw, E, C).copy_(torch.zeros(E, C).index_add(0, e, y * sg))


def lcao_threebody_bwd(B, NG, gram, unit, xk, ldxk, in_ptr, in_edge, in_src, out_ptr, out_edge, N, E, C, NL, d_tbw, dP, dB, q,
                       du_ks, du_st, stream):
    """reference-style chain (v, normalise, gate) differentiated directly — independent of the Gram formulation
    the CUDA kernel uses."""
    Bt, e, ep, Y, v, nrm, sg, src, cos, u = _tb_common(B, NG, unit, xk, ldxk, in_ptr, in_edge, in_src, out_ptr, out_edge, N, E, C, NL)
    G = view(d_tbw, E, C)[e]
    inv = 1.0 / nrm.clamp(min=1e-12)
    dy = G * sg
    coef = torch.where(nrm > 1e-12, (v * dy).sum(1, keepdim=True) * inv * inv, torch.zeros_like(nrm))
    dv = (dy - coef * v) * inv
    out = view(dB, E, NG * C).reshape(E, NG, C)
    out.zero_()
    out[:, :NL] = torch.zeros(E, NL, C).index_add(0, ep, Y.unsqueeze(-1) * dv.unsqueeze(1))
    if dP:
        two = view(dP, E, (NG - NL + 1) * C).reshape(E, NG - NL + 1, C)
        out[:, :NL] += two[:, :1]
        if NG > NL:
            out[:, NL] = two[:, 1]
    gy = torch.zeros(E, C).index_add(0, ep, G * v * inv)
    X = view(xk, N, C, ld=ldxk)
    s_all = X[src]
    view(q, E, C).copy_(gy * s_all * (1 - s_all))
    if du_ks:
        dYl = torch.einsum("tlc,tc->tl", Bt[ep][:, :NL], dv)  # dL/dY_l per triplet
        dc = (dYl * _sph_grad(cos, NL)).sum(1, keepdim=True)
        view(du_st, E, 3).copy_(torch.zeros(E, 3).index_add(0, e, dc * u[ep]))
        view(du_ks, E, 3).copy_(torch.zeros(E, 3).index_add(0, ep, dc * u[e]))


def _tw_load(B, NG, g, E, C, NL, valence):
    Bt = view(B, E, NG * C).reshape(E, NG, C)
    S = Bt[:, :NL].sum(1)
    if valence:
        gg = view(g, E, 2 * C)
        PV = Bt[:, NL]
        return S - PV, PV, gg[:, :C], gg[:, C:]
    return S, torch.zeros_like(S), view(g, E, C), torch.zeros_like(S)


def lcao_twobody_fwd(B, NG, g, E, C, NL, valence, lw, stream):
    PA, PV, gA, gV = _tw_load(B, NG, g, E, C, NL, valence)
    p = (1 + gA) * PA + (1 + gV) * PV
    view(lw, E, C).copy_(p / p.norm(dim=1, keepdim=True).clamp(min=1e-12))


def lcao_twobody_bwd(B, NG, g, d_lw, E, C, NL, valence, compact, dB, d_g, stream):
    PA, PV, gA, gV = _tw_load(B, NG, g, E, C, NL, valence)
    p = (1 + gA) * PA + (1 + gV) * PV
    nrm = p.norm(dim=1, keepdim=True)
    inv = 1.0 / nrm.clamp(min=1e-12)
    dl = view(d_lw, E, C)
    coef = torch.where(nrm > 1e-12, (p * dl).sum(1, keepdim=True) * inv * inv, torch.zeros_like(nrm))
    dp = (dl - coef * p) * inv
    NGo, NLo = ((1 + valence), 1) if compact else (NG, NL)
    out = view(dB, E, NGo * C).reshape(E, NGo, C)
    dPA = (1 + gA) * dp
    out[:, :NLo] = dPA.unsqueeze(1)
    if valence:
        out[:, NLo] = (1 + gV) * dp - dPA
        dg = view(d_g, E, 2 * C)
        dg[:, :C] = dp * PA
        dg[:, C:] = dp * PV
    else:
        view(d_g, E, C).copy_(dp * PA)


_F = torch.nn.functional
_ACTS = {0: lambda x: x, 1: _F.silu, 2: lambda x: _F.softplus(x) - math.log(2.0), 3: _F.softplus, 4: torch.relu, 5: torch.tanh,
         6: torch.sigmoid, 7: _F.gelu, 8: _F.elu, 9: _F.leaky_relu}


def _act(x, act):
    return _ACTS[act](x)


def _act_grad(h, act):
    """act'(h) of the activation codes of include/lcao_b200.h (autograd of the torch definition)."""
    if act == 0:
        return torch.ones_like(h)
    with torch.enable_grad():
        x = h.detach().clone().requires_grad_(True)
        (g,) = torch.autograd.grad(_ACTS[act](x).sum(), x)
    return g


def lcao_edge_pair_fwd(a, lda, b, ldb, bias, src32, dst32, E, C, act, out, pre, stream):
    s, t = view(src32, E, dtype=I32).long(), view(dst32, E, dtype=I32).long()
    N = int(max(s.max(), t.max())) + 1
    v = view(a, N, C, ld=lda)[s] + view(b, N, C, ld=ldb)[t]
    if bias:
        v = v + view(bias, C)
    if pre:
        view(pre, E, C).copy_(v)
    view(out, E, C).copy_(_act(v, act))


def lcao_segment_sum(x, ldx, y, ldy, ptr, perm, R, C, mean, out, ldo, stream):
    p = view(ptr, R + 1, dtype=I32).long()
    n = int(p[-1])
    pm = view(perm, n, dtype=I32).long() if perm else torch.arange(n)
    n_src = (int(pm.max()) + 1) if n else 0
    X = view(x, n_src, C, ld=ldx)
    if y:
        Y = view(y, n_src, C, ld=ldy)
        kind = ((mean >> 4) & 15) or 1
        if mean & 4:
            Y = _act_grad(Y, kind)
        X = X * (_act(Y, kind) if (mean & 2) else Y)
    O = view(out, R, C, ld=ldo)
    for r in range(R):
        seg = X[pm[p[r]:p[r + 1]]].sum(0)
        O[r] = seg / max(int(p[r + 1] - p[r]), 1) if (mean & 1) else seg


def lcao_msg_bwd(d_agg, lda, src32, h, bw, pre_h, E, C, act, d_bw, d_pre_h, stream):
    s = view(src32, E, dtype=I32).long()
    g = view(d_agg, int(s.max()) + 1, C, ld=lda)[s]
    p = view(pre_h, E, C)
    view(d_bw, E, C).copy_(g * (view(h, E, C) if h else _act(p, act)))
    view(d_pre_h, E, C).copy_(g * view(bw, E, C) * _act_grad(p, act))


def lcao_gather_rows(table, ldt, idx, is64, mul, ldm, n, W, out, ldo, stream):
    i = view(idx, n, dtype=I64 if is64 else I32).long()
    T = view(table, int(i.max()) + 1, W, ld=ldt)
    v = T[i]
    if mul:
        v = v * view(mul, n, W, ld=ldm)
    view(out, n, W, ld=ldo).copy_(v)


def lcao_reduce_by_key(x, ldx, kptr, kperm, nkeys, n, W, acc, stream):
    p = view(kptr, nkeys + 1, dtype=I32).long()
    pm = view(kperm, n, dtype=I32).long()
    X = view(x, n, W, ld=ldx)
    A = view(acc, nkeys, W)
    key_of = torch.repeat_interleave(torch.arange(nkeys), p[1:] - p[:-1])
    A.index_add_(0, key_of, X[pm])


def lcao_linear_fwd(X, ldx, W, bias, Y, ldy, pre, ldp, M, K, Nout, act, mode, stream):
    v = view(X, M, K, ld=ldx) @ view(W, Nout, K).t()
    if bias:
        v = v + view(bias, Nout)
    if pre:
        view(pre, M, Nout, ld=ldp).copy_(v)
    view(Y, M, Nout, ld=ldy).copy_(_act(v, act))


def _dy_eff(dY, ldy, H, ldh, act, M, Nout):
    d = view(dY, M, Nout, ld=ldy)
    if act != 0 and H:
        d = d * _act_grad(view(H, M, Nout, ld=ldh), act)
    return d


def lcao_linear_dgrad(dY, ldy, H, ldh, act, W, dX, ldx, M, K, Nout, accumulate, mode, scratch, stream):
    v = _dy_eff(dY, ldy, H, ldh, act, M, Nout) @ view(W, Nout, K)
    o = view(dX, M, K, ld=ldx)
    o.copy_(o + v if accumulate else v)


def lcao_linear_dgrad_act(dY, ldy, W, G, ldg, act, dX, ldx, M, K, Nout, mode, stream):
    v = view(dY, M, Nout, ld=ldy) @ view(W, Nout, K)
    view(dX, M, K, ld=ldx).copy_(v * _act_grad(view(G, M, K, ld=ldg), act))


def lcao_linear_wgrad(dY, ldy, H, ldh, act, X, ldx, dW, db, M, K, Nout, mode, scratch, stream):
    d = _dy_eff(dY, ldy, H, ldh, act, M, Nout)
    view(dW, Nout, K).add_(d.t() @ view(X, M, K, ld=ldx))
    if db:
        view(db, Nout).add_(d.sum(0))


def lcao_linear_wgrad_deferred(dY, ldy, X, ldx, dW, db, M, K, Nout, mode, scratch, desc, n_desc, stream):
    lcao_linear_wgrad(dY, ldy, None, 0, 0, X, ldx, dW, db, M, K, Nout, mode, scratch, stream)  # finished at once: no descriptors


def lcao_wgrad_reduce_batch(desc, n, stream):
    pass


def lcao_table_norm_fwd(x, counts, gamma, beta, R, Fd, eps, momentum, training, rmean, rvar, tracked, y, smean, srstd, stream):
    X = view(x, R, Fd)
    if training:
        c = view(counts, R)
        n = c.sum()
        w = (c / n).unsqueeze(1)
        mean = (w * X).sum(0)
        var = (w * (X - mean) ** 2).sum(0)
        if rmean:
            view(rmean, Fd).mul_(1 - momentum).add_(momentum * mean)
            view(rvar, Fd).mul_(1 - momentum).add_(momentum * var * (n / (n - 1)))
        if tracked:
            view(tracked, 1, dtype=I64).add_(1)
    else:
        mean, var = view(rmean, Fd).clone(), view(rvar, Fd).clone()
    rstd = torch.rsqrt(var + eps)
    out = (X - mean) * rstd
    if gamma:
        out = out * view(gamma, Fd) + view(beta, Fd)
    view(y, R, Fd).copy_(out)
    view(smean, Fd).copy_(mean)
    view(srstd, Fd).copy_(rstd)


def lcao_table_norm_bwd(dy, x, counts, gamma, smean, srstd, R, Fd, training, dx, dgamma, dbeta, stream):
    G, X = view(dy, R, Fd), view(x, R, Fd)
    mean, rstd = view(smean, Fd), view(srstd, Fd)
    xh = (X - mean) * rstd
    db, dg = G.sum(0), (G * xh).sum(0)
    sc = rstd * (view(gamma, Fd) if gamma else 1.0)
    v = G
    if training:
        c = view(counts, R)
        w = (c / c.sum()).unsqueeze(1)
        v = G - w * db - w * xh * dg
    view(dx, R, Fd).copy_(sc * v)
    if dgamma:
        view(dgamma, Fd).copy_(dg)
    if dbeta:
        view(dbeta, Fd).copy_(db)


def lcao_pair_outer_fwd(fe, za, zb, Zd, O, K, pre, stream):
    f, a, b = view(fe, Zd * O, K).reshape(Zd, O, K), view(za, Zd, K), view(zb, Zd, K)
    view(pre, Zd * Zd * O, K).copy_((f.unsqueeze(0) * (1.0 + a[:, None, None, :] + b[None, :, None, :])).reshape(-1, K))


def lcao_pair_outer_bwd(dpre, fe, za, zb, Zd, O, K, d_fe, d_za, d_zb, stream):
    g = view(dpre, Zd * Zd * O, K).reshape(Zd, Zd, O, K)
    f, a, b = view(fe, Zd * O, K).reshape(Zd, O, K), view(za, Zd, K), view(zb, Zd, K)
    view(d_fe, Zd * O, K).copy_((g * (1.0 + a[:, None, None, :] + b[None, :, None, :])).sum(0).reshape(-1, K))
    gf = g * f.unsqueeze(0)
    view(d_za, Zd, K).copy_(gf.sum((1, 2)))
    view(d_zb, Zd, K).copy_(gf.sum((0, 2)))


def lcao_act_fwd(X, ldx, Y, ldy, M, Cc, act, stream):
    view(Y, M, Cc, ld=ldy).copy_(_act(view(X, M, Cc, ld=ldx).clone(), act))


def lcao_act_bwd(dY, ldy, H, ldh, dH, ldd, M, Cc, act, stream):
    g = view(dY, M, Cc, ld=ldy)
    if act != 0:
        g = g * _act_grad(view(H, M, Cc, ld=ldh), act)
    view(dH, M, Cc, ld=ldd).copy_(g)


_TABLE = {k: v for k, v in list(globals().items()) if k.startswith("lcao_")}


class _FakeLib:
    """stands in for the CDLL in the few places ops.py calls the library object directly"""

    @staticmethod
    def lcao_linear_bwd_scratch(*a):
        return 0

    @staticmethod
    def lcao_pair_contract_bwd_scratch(*a):
        return 64


def install(monkeypatch):
    """Route lcaonet_b200's C-ABI calls to this emulator and lift the CUDA-only guards (pytest only)."""
    from lcaonet_b200 import _lib, ops

    def fake_call(name, *args):
        _TABLE[name](*args)

    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(_lib, "load", lambda: _FakeLib)
    monkeypatch.setattr(ops, "call", fake_call)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(ops, "stream_ptr", lambda: 0)
