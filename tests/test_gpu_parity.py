"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI,
against (a) the committed golden vectors of the unmodified reference, (b) the float64 oracle on seeded
inputs, (c) the per-kernel CPU restatement in tests/cpu_abi.py, and (d) size-independent properties
at the BASELINE.json batch size.  Integer outputs must be bit-exact; floating point tolerances are
stated at each assert (FP32 arithmetic, 1e-5 relative on energies per BASELINE.json:north_star)."""
import pytest
import torch

from lcaonet_b200 import LCAONet, ops
from lcaonet_b200.synth import GraphBatch, crystal_like_batch, graph_sizes, qm9_like_batch, reference_fixture_graph
from oracle import lcao_oracle as O
from tests import cpu_abi
from tests._util import full_cfg, graph_as, load_golden, rel_l2, same_triplets_up_to_duplicate_order

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN_CASES = ["qm9_default", "qm9_default_eval", "qm9_valence_ext_2perorb", "crystal_direct_forces_mean",
                "fixture_cosine_minmaxorb_atomref", "qm9_shiftedsoftplus", "qm9_gelu_valence", "qm9_sphericalbessel", "qm9_swish",
                "cfg1_qm9_32mol", "cfg3_valence_width128"]


# ---------------------------------------------------------------------------- index construction
@pytest.mark.parametrize("maker", [reference_fixture_graph, lambda: qm9_like_batch(16, 2), lambda: crystal_like_batch(2, 1),
                                   lambda: qm9_like_batch(1024, 0)])
def test_triplets_bit_exact_vs_oracle(maker):
    g = maker()
    ei, n = g["edge_index"], g["z"].shape[0]
    gi = ops.GraphIndex(ei.to(DEV), n)
    k, e_ks, e_st = gi.triplets()
    rk, reks, rest = O.triplets(ei, n)
    assert gi.num_triplets() == graph_sizes(g)["T"] == rk.numel()
    assert torch.equal(k.cpu(), rk) and torch.equal(e_ks.cpu(), reks) and torch.equal(e_st.cpu(), rest)


def test_triplets_edge_cases():
    # unsorted sources, self loops, duplicate pairs, isolated nodes, empty graph
    ei = torch.tensor([[3, 0, 0, 2, 0, 3, 3, 2], [0, 0, 3, 3, 3, 3, 2, 0]])
    gi = ops.GraphIndex(ei.to(DEV), 6)
    got = tuple(t.cpu() for t in gi.triplets())
    ref = O.triplets(ei, 6)
    assert all(torch.equal(a, b) for a, b in zip(got, ref))
    gi0 = ops.GraphIndex(torch.zeros(2, 0, dtype=torch.long, device=DEV), 5)
    assert gi0.num_triplets() == 0 and all(t.numel() == 0 for t in gi0.triplets())


def test_golden_triplets_fixture():
    gold = load_golden("fixture_small")
    gi = ops.GraphIndex(gold["graph"]["edge_index"].to(DEV), 3)
    k, e_ks, e_st = gi.triplets()
    t = gold["triplets"]
    assert k.tolist() == t["idx_k_3b"].tolist() and e_ks.tolist() == t["edge_idx_ks_3b"].tolist()
    assert e_st.tolist() == t["edge_idx_st_3b"].tolist()


def test_bucket_sort_is_stable_and_exact():
    torch.manual_seed(0)
    keys = torch.randint(0, 37, (5000,))
    sec = torch.randint(0, 11, (5000,))
    p, perm = ops.bucket_sort(keys.to(DEV), 37, sec.to(DEV))
    order = torch.argsort(sec, stable=True)
    order = order[torch.argsort(keys[order], stable=True)]
    assert torch.equal(perm.cpu().long(), order)
    assert torch.equal(p.cpu().long()[1:], torch.bincount(keys, minlength=37).cumsum(0))


# ---------------------------------------------------------------------------- basis
@pytest.mark.parametrize("cfgk", [dict(max_z=36, cutoff=5.0, cutoff_net="polynomial"), dict(max_z=36, n_per_orb=2, cutoff=6.0),
                                  dict(max_z=84, max_orb="6d", cutoff=3.0, cutoff_net="cosine"),
                                  dict(max_z=12, cutoff=2.0, rbf_type="sphericalbessel")])
def test_radial_basis_vs_oracle(cfgk):
    model = LCAONet(emb_size=16, emb_size_coeff=16, emb_size_conv=16, **cfgk)
    r = torch.linspace(0.05, 7.0, 1500)
    rb = model.rbf(r.to(DEV)).cpu().double()
    nl = O.orbital_quantum_numbers(cfgk["max_z"], cfgk.get("max_orb"), cfgk.get("n_per_orb", 1))
    ref = O.radial_basis(r.double(), nl, cfgk["cutoff"], cfgk.get("cutoff_net", "envelope"), cfgk.get("rbf_type", "hydrogen"))
    scale = ref.abs().max(0).values.clamp(min=1e-30)
    assert float(((rb - ref).abs() / scale).max()) < 2e-6  # fp32 rounding of an fp64 evaluation
    assert torch.all(rb[r > cfgk["cutoff"]] == 0)


# ---------------------------------------------------------------------------- per-kernel vs CPU restatement
def _both(fn, monkeypatch, *cpu_args):
    """run fn(*args) on the GPU through the real library and on the CPU through the emulator."""
    gpu_args = [a.to(DEV) if torch.is_tensor(a) else a for a in cpu_args]
    for a, b in zip(cpu_args, gpu_args):
        if torch.is_tensor(a) and a.requires_grad:
            b.requires_grad_(True)
    out_g = fn(*gpu_args)
    with monkeypatch.context() as m:
        cpu_abi.install(m)
        out_c = fn(*cpu_args)
    return out_g, out_c, gpu_args


@pytest.mark.parametrize("M,K,N,silu,bias", [(1000, 128, 128, True, False), (777, 256, 128, True, True), (64, 128, 256, False, True),
                                             (333, 64, 1, False, False), (5, 20, 12, True, True), (40000, 128, 128, True, False),
                                             (37, 256, 128, False, False), (296, 128, 128, True, False), (512, 128, 128, True, True),
                                             (513, 128, 128, True, True)])
def test_linear_fwd_bwd(M, K, N, silu, bias, monkeypatch):
    torch.manual_seed(M)
    x = torch.randn(M, K, requires_grad=True)
    w = (torch.randn(N, K) / K**0.5).requires_grad_(True)
    b = torch.randn(N, requires_grad=True) if bias else None
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    ref = torch.nn.functional.silu(ref) if silu else ref
    gx, gw = torch.autograd.grad((ref**2).sum(), [x, w])
    xg, wg = x.detach().to(DEV).requires_grad_(True), w.detach().to(DEV).requires_grad_(True)
    bg = b.detach().to(DEV).requires_grad_(True) if bias else None
    y = ops.linear(xg, wg, bg, silu)
    assert rel_l2(y, ref) < 2e-6
    (y**2).sum().backward()
    assert rel_l2(xg.grad, gx) < 5e-6 and rel_l2(wg.grad, gw) < 5e-6


def _edge_problem(seed=0, n_mol=6, C=32, NL=3, valence=0, crystal=False):
    g = crystal_like_batch(1, seed, margin=0.05) if crystal else qm9_like_batch(n_mol, seed, margin=0.05)
    torch.manual_seed(seed)
    E, N = g["edge_index"].shape[1], g["z"].shape[0]
    NG = NL + valence
    B = torch.randn(E, NG, C)
    unit = torch.nn.functional.normalize(torch.randn(E, 3), dim=1)
    xk = torch.randn(N, 2 * C)
    return g, B, unit, xk


@pytest.mark.parametrize("C,NL,valence,crystal,forces", [(32, 3, 0, False, False), (128, 3, 0, False, False), (128, 3, 0, False, True),
                                                         (128, 3, 1, True, True), (64, 4, 0, True, False), (12, 2, 1, False, True),
                                                         (256, 1, 0, False, False), (160, 4, 0, False, True)])
def test_threebody_fwd_bwd_vs_restatement(C, NL, valence, crystal, forces, monkeypatch):
    """Gram-matrix kernels vs the reference-style per-triplet chain (tests/cpu_abi.py) incl. dL/d unit."""
    g, B, unit, xk = _edge_problem(C=C, NL=NL, valence=valence, crystal=crystal)
    ei, N = g["edge_index"], g["z"].shape[0]
    B[::13] = 0.0  # edges beyond the cutoff: zero rows take the eps-clamped branch

    def fn(ei, B, unit, xk):
        B = B.clone().requires_grad_(True)
        xk = xk.clone().requires_grad_(True)
        unit = unit.clone().requires_grad_(forces)
        gi = ops.GraphIndex(ei, N)
        tbw = ops.threebody(B, unit, xk[:, C:], gi, NL)
        (tbw * torch.linspace(-1, 1, C, device=tbw.device)).sum().backward()
        return tbw.detach(), B.grad, xk.grad, (unit.grad if forces else tbw.detach())

    (t_g, dB_g, dx_g, du_g), (t_c, dB_c, dx_c, du_c), _ = _both(fn, monkeypatch, ei, B, unit, xk)
    assert torch.isfinite(dB_g).all() and torch.isfinite(du_g).all()
    assert rel_l2(t_g, t_c) < 2e-6 and rel_l2(dB_g, dB_c) < 1e-5 and rel_l2(dx_g, dx_c) < 1e-5
    assert rel_l2(du_g, du_c) < 1e-5


def test_threebody_chunked_degrees(monkeypatch):
    """a hub node with more in/out edges than one shared-memory chunk (32) plus directed leftovers"""
    torch.manual_seed(5)
    hub, n = 0, 80
    s = torch.cat([torch.zeros(n - 1, dtype=torch.long), torch.arange(1, n), torch.tensor([3, 4, 5])])
    t = torch.cat([torch.arange(1, n), torch.zeros(n - 1, dtype=torch.long), torch.tensor([4, 5, 5])])
    ei = torch.stack([s, t])
    E, C, NL = ei.shape[1], 128, 3
    B, unit, xk = torch.randn(E, NL, C), torch.nn.functional.normalize(torch.randn(E, 3), dim=1), torch.randn(n, C)

    def fn(ei, B, unit, xk):
        B, xk, unit = B.clone().requires_grad_(True), xk.clone().requires_grad_(True), unit.clone().requires_grad_(True)
        tbw = ops.threebody(B, unit, xk, ops.GraphIndex(ei, n), NL)
        (tbw * torch.linspace(-1, 1, C, device=tbw.device)).sum().backward()
        return tbw.detach(), B.grad, xk.grad, unit.grad

    out_g, out_c, _ = _both(fn, monkeypatch, ei, B, unit, xk)
    for a, b, tol in zip(out_g, out_c, (2e-6, 1e-5, 1e-5, 1e-5)):
        assert rel_l2(a, b) < tol


@pytest.mark.parametrize("C,NL,valence,O,P", [(128, 3, 0, 8, 37 * 37), (128, 3, 1, 16, 400), (32, 4, 1, 5, 9), (256, 2, 0, 8, 30)])
def test_pair_contract_vs_restatement(C, NL, valence, O, P, monkeypatch):
    torch.manual_seed(7)
    E = 20000
    Cp = C * (1 + valence)
    tab = torch.randn(P, O, Cp)
    pair = torch.randint(0, P, (E,))
    pair[: E // 2] = pair[0]  # one dominant key (many chunks) plus a long tail, some keys absent
    rb = torch.randn(E, O)
    vmask = (torch.rand(E, O) > 0.5).float() if valence else None
    lgrp = torch.randint(0, NL, (O,), dtype=torch.int32)
    lgrp[0] = NL - 1

    def fn(tab, pair, rb, vmask, lgrp):
        tab, rb = tab.clone().requires_grad_(True), rb.clone().requires_grad_(True)
        kptr, kperm = ops.bucket_sort(pair, P, stable=False)
        B, gram = ops.pair_contract(tab, pair, (kptr, kperm), rb, vmask, lgrp, NL, C)
        (B * torch.linspace(-1, 2, C, device=B.device)).sum().backward()
        return B.detach(), gram, tab.grad, rb.grad

    out_g, out_c, _ = _both(fn, monkeypatch, tab, pair, rb, vmask, lgrp)
    for a, b, tol in zip(out_g, out_c, (2e-6, 1e-6, 2e-5, 2e-5)):
        assert torch.isfinite(a).all() and rel_l2(a, b) < tol


@pytest.mark.parametrize("C,NL,valence", [(128, 3, 0), (128, 3, 1), (32, 4, 1), (256, 2, 0)])
def test_twobody_and_contract_vs_restatement(C, NL, valence, monkeypatch):
    torch.manual_seed(1)
    E, O = 3000, 8
    Cp = C * (1 + valence)
    cst1 = torch.randn(E, O, Cp)
    rb = torch.randn(E, O)
    rb[::17] = 0.0  # edges beyond the cutoff: zero rows must normalise to zero without NaN
    vmask = (torch.rand(E, O) > 0.5).float() if valence else None
    lgrp = torch.randint(0, NL, (O,), dtype=torch.int32)
    lgrp[0] = NL - 1
    g = torch.randn(E, Cp)

    def fn(cst1, rb, vmask, lgrp, g):
        cst1 = cst1.clone().requires_grad_(True)
        rb = rb.clone().requires_grad_(True)
        g = g.clone().requires_grad_(True)
        B = ops.coeff_contract(cst1, rb, vmask, lgrp, NL, C)
        lw = ops.twobody(B, g, NL, valence)
        (lw * torch.linspace(-1, 2, C, device=lw.device) + 0.01 * B.sum(1)).sum().backward()
        return B.detach(), lw.detach(), cst1.grad, rb.grad, g.grad

    out_g, out_c, _ = _both(fn, monkeypatch, cst1, rb, vmask, lgrp, g)
    for a, b, tol in zip(out_g, out_c, (2e-6, 2e-6, 2e-5, 2e-5, 2e-5)):
        assert torch.isfinite(a).all() and rel_l2(a, b) < tol


def test_edge_gathers_and_segment_sums_vs_restatement(monkeypatch):
    g = qm9_like_batch(8, 3)
    ei, N, C = g["edge_index"], g["z"].shape[0], 64
    torch.manual_seed(2)
    u, x, y, bias = torch.randn(N, 2 * C), torch.randn(ei.shape[1], C), torch.randn(ei.shape[1], C), torch.randn(C)

    def fn(ei, u, x, y, bias, batch):
        u, x, y = (t.clone().requires_grad_(True) for t in (u, x, y))
        bias = bias.clone().requires_grad_(True)
        gi = ops.GraphIndex(ei, N)
        a = ops.edge_pair(u[:, :C], u[:, C:], bias, gi, silu=True)
        agg = ops.mul_segment_sum(x * a, y, gi)
        seg = (*ops.bucket_sort(batch, 8), batch)
        pooled = ops.segment_reduce(agg, *seg, mean=True)
        (pooled**2).sum().backward()
        return a.detach(), agg.detach(), pooled.detach(), u.grad, x.grad, y.grad, bias.grad

    out_g, out_c, _ = _both(fn, monkeypatch, ei, u, x, y, bias, g["batch"])
    for a, b in zip(out_g, out_c):
        assert rel_l2(a, b) < 1e-5


def test_integral_norm_radial_basis_on_gpu():
    """HydrogenRadialBasis(integral_norm=True) (rbf.py:107-127) through the geometry kernel vs the reference's table"""
    from lcaonet_b200.model import RadialBasis
    from lcaonet_b200.orbitals import ElecInfo
    tab = load_golden("basis_tables")
    r = tab["r"][1:]  # (r = 0 is not a distance the geometry kernel sees)
    for (rc, cname), want in tab["rb_integral_norm"].items():
        rbf = RadialBasis(rc, ElecInfo(36, None, None, 1), cname, "hydrogen", integral_norm=True)
        got = rbf(r.float().to(DEV))
        assert rel_l2(got, want[1:]) < 2e-6, (rc, cname)


# ---------------------------------------------------------------------------- end to end vs the reference
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_model_matches_reference_golden(name):
    gold = load_golden(name)
    model = LCAONet(**gold["kwargs"])
    model.load_state_dict(gold["state_dict"], strict=True)
    model = model.to(DEV).train(gold["training"])
    g = GraphBatch(gold["graph"]).to(DEV)
    out = model(g)
    if isinstance(out, tuple):
        energy, forces = out
        assert rel_l2(forces, gold["forces_f64"]) < 1e-5
        loss = (energy**2).mean() + (forces**2).mean()
    else:
        energy, loss = out, (out**2).mean()
    # reference FP32's own distance to FP64 on these graphs is 1e-7..1e-6 (printed by make_golden.py)
    assert rel_l2(energy, gold["energy_f64"]) < 1e-5
    trip = gold["triplets"]
    assert same_triplets_up_to_duplicate_order(
        (g["idx_k_3b"], g["edge_idx_ks_3b"], g["edge_idx_st_3b"]),
        (trip["idx_k_3b"], trip["edge_idx_ks_3b"], trip["edge_idx_st_3b"]))
    assert torch.allclose(g["edge_dist"].cpu().double(), gold["edge_dist_f64"], rtol=1e-6, atol=1e-6)
    # cos(theta) per triplet (lcaonet.py:431-435): same multiset always, same positions when no (k, s) pair is duplicated
    ang = g["angles_3b"].cpu()
    assert torch.allclose(ang.sort().values, gold["angles_f32"].sort().values, atol=2e-6)
    if torch.equal(g["edge_idx_ks_3b"].cpu().to(torch.int32), trip["edge_idx_ks_3b"]):
        assert torch.allclose(ang, gold["angles_f32"], atol=2e-6)
    loss.backward()
    worst, worst_ref = 0.0, 0.0
    for n, p in model.named_parameters():
        ref = gold["grads_f64"][n]
        if ref is None or float(ref.norm()) == 0.0:
            continue
        worst = max(worst, rel_l2(p.grad, ref))
        worst_ref = max(worst_ref, gold["grads_f32_rel_to_f64"][n])
    # per-tensor gradient rel-L2 vs the FP64 reference; the FP32 reference itself sits at `worst_ref`
    assert worst < max(5e-5, 3 * worst_ref), (worst, worst_ref)
    if gold["training"]:
        sd = model.state_dict()
        for k, v in gold["bn_after_f64"].items():
            assert torch.allclose(sd[k].cpu().double(), v, rtol=1e-4, atol=1e-5), k


def test_autograd_forces_match_reference_golden():
    from tests.test_host_logic_cpu import _check_autograd_forces
    for name in ("crystal_autograd_forces", "cfg4_crystal_width128"):
        gold = load_golden(name)
        model = LCAONet(**gold["kwargs"])
        model.load_state_dict(gold["state_dict"], strict=True)
        model = model.to(DEV).train(gold["training"])
        out = model(GraphBatch(gold["graph"]).to(DEV))
        _check_autograd_forces(gold, model, out, 1e-5, 5e-5)


def test_degenerate_graphs_match_oracle():
    """isolated atoms (no edges at all for some nodes), a batch without any edge, and batch=None"""
    kwargs = dict(emb_size=32, emb_size_coeff=32, emb_size_conv=32, cutoff=5.0, cutoff_net="polynomial")
    torch.manual_seed(5)
    model = LCAONet(**kwargs)
    p = O.cast_params(model.state_dict(), torch.float64)
    model = model.to(DEV).train()
    g = qm9_like_batch(3, seed=12, margin=0.05)
    n0 = g["z"].shape[0]
    # append two isolated atoms as a 4th "molecule"
    g2 = GraphBatch(z=torch.cat([g["z"], torch.tensor([1, 8])]), pos=torch.cat([g["pos"], torch.tensor([[0.0, 0, 0], [30.0, 0, 0]])]),
                    edge_index=g["edge_index"], edge_shift=g["edge_shift"], lattice=torch.cat([g["lattice"], g["lattice"][:1]]),
                    batch=torch.cat([g["batch"], torch.tensor([3, 3])]))
    assert g2["z"].shape[0] == n0 + 2
    out = model(g2.to(DEV))
    ref = O.forward(p, full_cfg(kwargs), graph_as(g2, torch.float64), training=True)
    assert rel_l2(out, ref) < 1e-5
    (out**2).mean().backward()
    assert all(torch.isfinite(q.grad).all() for q in model.parameters() if q.grad is not None)
    # no edges at all
    g3 = GraphBatch(z=torch.tensor([1, 6, 8]), pos=torch.tensor([[0.0, 0, 0], [20.0, 0, 0], [0, 20.0, 0]]),
                    edge_index=torch.zeros(2, 0, dtype=torch.long), edge_shift=torch.zeros(0, 3), lattice=50 * torch.eye(3).unsqueeze(0),
                    batch=torch.zeros(3, dtype=torch.long))
    model.eval()
    with torch.no_grad():
        out3 = model(g3.to(DEV))  # (the reference itself cannot run this case: it reshapes a 0-row tensor with -1)
        assert out3.shape == (1, 1) and torch.isfinite(out3).all()
        singles = [GraphBatch(z=g3["z"][i:i + 1], pos=g3["pos"][i:i + 1], edge_index=g3["edge_index"], edge_shift=g3["edge_shift"],
                              lattice=g3["lattice"], batch=torch.zeros(1, dtype=torch.long)) for i in range(3)]
        assert rel_l2(sum(model(s.to(DEV)) for s in singles), out3) < 1e-6  # no edges: atoms contribute independently
        # batch = None: a single graph
        g1 = qm9_like_batch(1, seed=13, margin=0.05)
        p_now = O.cast_params({k: v.cpu() for k, v in model.state_dict().items()}, torch.float64)  # BN stats moved above
        ref1 = O.forward(p_now, full_cfg(kwargs), graph_as(g1, torch.float64), training=False)
        g1n = GraphBatch({k: v for k, v in g1.items() if k != "batch"})
        assert rel_l2(model(g1n.to(DEV)), ref1) < 1e-5


def test_node_features_match_oracle_layer_by_layer():
    kwargs = dict(cutoff=5.0, cutoff_net="polynomial")
    torch.manual_seed(3)
    model = LCAONet(**kwargs)
    g = qm9_like_batch(12, seed=9, margin=0.05)
    p = O.cast_params(model.state_dict(), torch.float64)
    trace = {}
    O.forward(p, full_cfg(kwargs), graph_as(g, torch.float64), training=True, trace=trace)
    model = model.to(DEV).train()
    feats = []
    hooks = [layer.register_forward_hook(lambda m, i, o: feats.append(o.detach())) for layer in model.int_layers]
    model(g.to(DEV))
    for h in hooks:
        h.remove()
    for i, f in enumerate(feats):
        assert rel_l2(f, trace[f"x{i + 1}"]) < 1e-5, i


# ---------------------------------------------------------------------------- full-size properties
def test_full_size_batch_properties():
    """BASELINE.json config 2 shape (1024 QM9-like molecules): determinism, per-molecule additivity
    (energies of a batch == energies of its halves, eval mode) and finite gradients."""
    torch.manual_seed(0)
    model = LCAONet(cutoff=5.0, cutoff_net="polynomial").to(DEV).eval()
    model.side_effect_keys = False
    g = qm9_like_batch(1024, seed=0)
    with torch.no_grad():
        e1 = model(g.to(DEV))
        e2 = model(g.to(DEV))
        assert torch.equal(e1, e2)  # no atomics on the forward path: bitwise reproducible
        z, b = g["z"], g["batch"]
        half = int((b < 512).sum())
        s, t = g["edge_index"]
        m = s < half
        lo = GraphBatch(z=z[:half], pos=g["pos"][:half], edge_index=g["edge_index"][:, m], edge_shift=g["edge_shift"][m],
                        lattice=g["lattice"][:512], batch=b[:half])
        e_lo = model(lo.to(DEV))
    assert rel_l2(e_lo, e1[:512]) < 1e-6
    model.train()
    out = model(g.to(DEV))
    (out**2).mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


# ---------------------------------------------------------------------------- tcgen05 GEMMs
@pytest.mark.parametrize("mode,tol", [("tf32x3", 3e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("M,K,N,silu,bias", [(5000, 128, 128, True, False), (4173, 128, 256, True, True), (3000, 64, 128, False, True),
                                             (2000, 32, 16, True, False), (1500, 128, 48, False, False), (600, 96, 128, True, True),
                                             (300000, 128, 128, True, False)])
def test_linear_tensor_core_modes(mode, tol, M, K, N, silu, bias):
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K**0.5
    b = torch.randn(N, device=DEV) if bias else None
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    ref = torch.nn.functional.linear(xd, wd, bd)
    ref = torch.nn.functional.silu(ref) if silu else ref
    probe = torch.randn(M, N, device=DEV, dtype=torch.float64)
    grads = torch.autograd.grad((ref * probe).sum(), [xd, wd] + ([bd] if bias else []))
    prev = ops.get_gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        bg = b.clone().requires_grad_(True) if bias else None
        y = ops.linear(xg, wg, bg, silu)
        (y * probe.float()).sum().backward()
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode(prev)
    assert rel_l2(y, ref) < tol
    # the weight gradient sums M terms in FP32: allow the sqrt(M) * eps growth any FP32 reduction shows
    tol_w = max(2 * tol, 2e-8 * M**0.5)
    assert rel_l2(xg.grad, grads[0]) < 2 * tol and rel_l2(wg.grad, grads[1]) < tol_w
    if bias:
        assert rel_l2(bg.grad, grads[2]) < tol_w


def test_model_golden_with_tf32x3_gemms():
    gold = load_golden("qm9_default")
    model = LCAONet(**gold["kwargs"])
    model.load_state_dict(gold["state_dict"], strict=True)
    model = model.to(DEV).train()
    prev = ops.get_gemm_mode()
    ops.set_gemm_mode("tf32x3")
    try:
        out = model(GraphBatch(gold["graph"]).to(DEV))
        (out**2).mean().backward()
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode(prev)
    assert rel_l2(out, gold["energy_f64"]) < 1e-5
    worst = max(rel_l2(p.grad, gold["grads_f64"][n]) for n, p in model.named_parameters()
                if gold["grads_f64"][n] is not None and float(gold["grads_f64"][n].norm()) > 0)
    assert worst < 5e-5, worst


def test_graphed_embedding_tables_equal_eager():
    """LCAOEmbedding.graph_tables replays the static-shape table arithmetic as CUDA graphs: over several training steps
    on DIFFERENT batches the outputs and the BatchNorm buffers must equal the eager path's bit for bit (same kernels, same
    order; building the graphs must not disturb the running statistics) and every gradient to FP32 rounding (the
    table-sized weight-gradient kernel and torch's embedding backward accumulate with atomics: run-to-run order)."""
    torch.manual_seed(0)
    kw = dict(cutoff=5.0, cutoff_net="polynomial")
    m_e, m_g = LCAONet(**kw).to(DEV).train(), LCAONet(**kw).to(DEV).train()
    m_g.load_state_dict(m_e.state_dict())
    m_g.emb_layer.graph_tables = True
    for step in range(3):
        g = qm9_like_batch(8 + 4 * step, seed=40 + step)
        outs = []
        for m in (m_e, m_g):
            m.zero_grad(set_to_none=True)
            out = m(GraphBatch(g).to(DEV))
            torch.nn.functional.mse_loss(out, g["y"].to(DEV)).backward()
            outs.append(out.detach())
        assert torch.equal(outs[0], outs[1]), step
        for (n, p), q in zip(m_e.named_parameters(), m_g.parameters()):
            assert (p.grad is None) == (q.grad is None), n
            if p.grad is not None:
                assert rel_l2(q.grad, p.grad) < 1e-5, (step, n, rel_l2(q.grad, p.grad))
        for (n, b), c in zip(m_e.named_buffers(), m_g.buffers()):
            assert torch.equal(b, c), (step, n)
    # moving the parameters invalidates the captured pointers: the graphs must be rebuilt, not replayed on freed storage
    m_g = m_g.cpu().to(DEV)
    g = qm9_like_batch(7, seed=91)
    outs = []
    for m in (m_e, m_g):
        m.zero_grad(set_to_none=True)
        out = m(GraphBatch(g).to(DEV))
        torch.nn.functional.mse_loss(out, g["y"].to(DEV)).backward()
        outs.append(out.detach())
    assert torch.equal(outs[0], outs[1])
    # inference falls back to the eager path and still agrees
    m_e.eval(), m_g.eval()
    g = qm9_like_batch(5, seed=77)
    with torch.no_grad():
        assert torch.equal(m_e(GraphBatch(g).to(DEV)), m_g(GraphBatch(g).to(DEV)))


def test_long_exclusive_scans_are_exact():
    """index arrays beyond one scan block take the two-launch scan: offsets must stay bit-exact (E = 10^5 .. 10^6)."""
    for n_mol, seed in ((400, 5), (2500, 6)):
        g = qm9_like_batch(n_mol, seed)
        ei, n = g["edge_index"], g["z"].shape[0]
        gi = ops.GraphIndex(ei.to(DEV), n)
        deg_in = torch.bincount(ei[1], minlength=n)
        ref = torch.zeros(ei.shape[1] + 1, dtype=torch.int64)
        ref[1:] = (deg_in[ei[0]] - (ei[0] == ei[1]).long()).cumsum(0)
        assert torch.equal(gi.tri_ptr.cpu().long(), ref)
        assert torch.equal(gi.in_ptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), deg_in.cumsum(0)]))


@pytest.mark.parametrize("name", ["silu", "shiftedsoftplus", "softplus", "relu", "tanh", "sigmoid", "gelu", "elu", "leakyrelu"])
@pytest.mark.parametrize("M", [300, 5000])  # CUDA-core and tcgen05 paths
def test_fused_activations_fwd_bwd(name, M):
    """every LCAO_ACT_* kind through the dense layer (epilogue + act' backward) and the edge gather-add, against torch."""
    import math

    from lcaonet_b200._lib import ACT
    F = torch.nn.functional
    ref_fn = {"silu": F.silu, "shiftedsoftplus": lambda v: F.softplus(v) - math.log(2.0), "softplus": F.softplus, "relu": F.relu,
              "tanh": torch.tanh, "sigmoid": torch.sigmoid, "gelu": F.gelu, "elu": F.elu, "leakyrelu": F.leaky_relu}[name]
    torch.manual_seed(1)
    K = N = 128
    x = (2.0 * torch.randn(M, K)).to(DEV).requires_grad_(True)
    w = (torch.randn(N, K) / K**0.5).to(DEV).requires_grad_(True)
    b = torch.randn(N).to(DEV).requires_grad_(True)
    dy = torch.randn(M, N).to(DEV)
    y = ops.linear(x, w, b, ACT[name])
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), dy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = ref_fn(F.linear(xd, wd, bd))
    rx, rw, rb = torch.autograd.grad(yr, (xd, wd, bd), dy.double())
    assert rel_l2(y, yr) < 3e-6
    assert rel_l2(gx, rx) < 1e-5 and rel_l2(gw, rw) < 1e-5 and rel_l2(gb, rb) < 1e-5
    # gather-add + activation per edge
    g = qm9_like_batch(4, seed=9)
    gi = ops.GraphIndex(g["edge_index"].to(DEV), g["z"].shape[0])
    a = torch.randn(gi.N, 64, device=DEV, requires_grad=True)
    c = torch.randn(gi.N, 64, device=DEV, requires_grad=True)
    out = ops.edge_pair(a, c, None, gi, ACT[name])
    do = torch.randn_like(out)
    ga, gc = torch.autograd.grad(out, (a, c), do)
    ad, cd = a.detach().double().requires_grad_(True), c.detach().double().requires_grad_(True)
    ei = g["edge_index"].to(DEV)
    outr = ref_fn(ad[ei[0]] + cd[ei[1]])
    ra, rc = torch.autograd.grad(outr, (ad, cd), do.double())
    assert rel_l2(out, outr) < 1e-6 and rel_l2(ga, ra) < 1e-5 and rel_l2(gc, rc) < 1e-5


def test_bucket_sort_many_buckets_multi_block_scan():
    """more than two scan blocks of buckets (the two-launch scan over the bucket counts) incl. exact block multiples."""
    for nb, n in ((4096 * 3, 50_000), (20_000, 70_001), (8193, 5)):
        torch.manual_seed(nb)
        keys = torch.randint(0, nb, (n,))
        p, perm = ops.bucket_sort(keys.to(DEV), nb, stable=True)
        ref_ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(keys, minlength=nb).cumsum(0)])
        assert torch.equal(p.cpu().long(), ref_ptr)
        assert torch.equal(perm.cpu().long(), torch.argsort(keys, stable=True))


def test_pair_outer_fwd_bwd_vs_torch():
    """lcao_pair_outer_*: pre = fe (1 + za + zb) on the species-pair table and its three gradients vs the eager expression (FP64)."""
    for Zd, O_, K in ((37, 8, 128), (5, 16, 32), (95, 3, 64)):
        torch.manual_seed(Zd)
        fe, za, zb = torch.randn(Zd, O_, K), torch.randn(Zd, K), torch.randn(Zd, K)
        w = torch.randn(Zd, Zd, O_, K)
        fd, ad, bd = (t.double().requires_grad_(True) for t in (fe, za, zb))
        ref = fd.unsqueeze(0) * (1.0 + ad[:, None, None, :] + bd[None, :, None, :])
        gr = torch.autograd.grad((ref * w.double()).sum(), [fd, ad, bd])
        fg, ag, bg = (t.to(DEV).requires_grad_(True) for t in (fe, za, zb))
        out = ops.pair_outer(fg, ag, bg)
        (out * w.to(DEV)).sum().backward()
        assert rel_l2(out, ref) < 1e-6
        for a, b in zip((fg.grad, ag.grad, bg.grad), gr):
            assert rel_l2(a, b) < 2e-6


def test_bucket_sort_ordered_few_huge_buckets():
    """stable="ordered": bucket order = ascending item id whatever order atomics land in (species / pair keys), incl. one
    dominant key, empty keys, partial last blocks, the (94 + 1)^2 pair table and an empty input."""
    for nb, n in ((37, 18_470), (1369, 252_798), (9025, 100_001), (5, 1), (1369, 1024), (7, 0)):
        torch.manual_seed(n + nb)
        keys = torch.randint(0, nb, (n,))
        if n > 10:
            keys[: n // 2] = keys[0]
        p, perm = ops.bucket_sort(keys.to(DEV), nb, stable="ordered")
        ref_ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(keys, minlength=nb).cumsum(0)])
        assert torch.equal(p.cpu().long(), ref_ptr)
        assert torch.equal(perm.cpu().long(), torch.argsort(keys, stable=True))


def test_interaction_gradients_are_bitwise_reproducible():
    """Two identical training steps on BASELINE config 2's shape (256 molecules) give bitwise equal gradients for every
    parameter of the interaction blocks and of the coefficient embedding (the species-pair table path): sorted-segment
    sums, the ordered pair grouping, the two-stage keyed reductions and the tcgen05 weight gradients do not depend on
    the order atomics land in.  Not covered (atomic flushes): the species-row reduction of the node
    embedding (lcao_reduce_by_key) and the CUDA-core fallback weight gradients of the 64- and 1-wide output layers."""
    g = qm9_like_batch(256, seed=5).to(DEV)
    grads = []
    for _ in range(2):
        torch.manual_seed(11)
        model = LCAONet(cutoff=5.0, cutoff_net="polynomial").to(DEV).train()
        model.side_effect_keys = False
        out = model(g)
        ((out - g["y"].to(DEV).reshape(out.shape)) ** 2).mean().backward()
        grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    checked = 0
    for n in grads[0]:
        if n.startswith("int_layers.") or n.startswith("emb_layer.coeff_embed.") or n.startswith("out_layer.out_lin.0."):
            assert torch.equal(grads[0][n], grads[1][n]), n
            checked += 1
        else:
            assert rel_l2(grads[0][n], grads[1][n]) < 1e-4, n
    assert checked >= 30


@pytest.mark.parametrize("name", ["silu", "shiftedsoftplus", "tanh"])
def test_segment_sum_with_activation_in_flight(name):
    """lcao_segment_sum flag bits 1 / 2 (factor = act(y) / act'(y)) and lcao_msg_bwd without a stored h, vs torch."""
    import math

    from lcaonet_b200._lib import ACT
    F = torch.nn.functional
    fn = {"silu": F.silu, "shiftedsoftplus": lambda v: F.softplus(v) - math.log(2.0), "tanh": torch.tanh}[name]
    g = qm9_like_batch(6, seed=21)
    gi = ops.GraphIndex(g["edge_index"].to(DEV), g["z"].shape[0])
    E, N, C, act = gi.E, gi.N, 128, ACT[name]
    torch.manual_seed(3)
    x, y = torch.randn(E, C, device=DEV), 2.0 * torch.randn(E, C, device=DEV)
    yd = y.double().requires_grad_(True)
    fy = fn(yd)
    (dfy,) = torch.autograd.grad(fy.sum(), yd)
    src = g["edge_index"][0].to(DEV)
    P, st = ops.ptr, ops.stream_ptr
    for flag, factor in ((2, fy.detach()), (4, dfy)):
        out = torch.empty(N, C, device=DEV)
        ops._call("lcao_segment_sum", P(x), C, P(y), C, P(gi.out_ptr), P(gi.out_edge), N, C, flag | (act << 4), P(out), C, st())
        ref = torch.zeros(N, C, dtype=torch.float64, device=DEV).index_add_(0, src, x.double() * factor)
        assert rel_l2(out, ref) < 2e-6, (name, flag)
    d_agg, bw = torch.randn(N, C, device=DEV), torch.randn(E, C, device=DEV)
    d_bw, d_pre = torch.empty(E, C, device=DEV), torch.empty(E, C, device=DEV)
    ops._call("lcao_msg_bwd", P(d_agg), C, P(gi.src32), None, P(bw), P(y), E, C, act, P(d_bw), P(d_pre), st())
    ga = d_agg.double()[src]
    assert rel_l2(d_bw, ga * fy.detach()) < 2e-6 and rel_l2(d_pre, ga * bw.double() * dfy) < 2e-6


@pytest.mark.parametrize("R,Fd,training", [(37, 128, True), (1369, 1024, True), (1369, 1000, False), (5, 33, True)])
def test_table_norm_vs_weighted_batchnorm(R, Fd, training):
    """lcao_table_norm_fwd/bwd against nn.BatchNorm1d applied to the EXPANDED batch (rows repeated by their counts) in
    FP64: outputs, running statistics, num_batches_tracked and all gradients; rows with a zero count included."""
    torch.manual_seed(R + Fd)
    x = torch.randn(R, Fd)
    counts = torch.randint(0, 7, (R,)).float()
    counts[0] = 3.0
    bn = torch.nn.BatchNorm1d(Fd).double()
    bn.weight.data.uniform_(0.5, 1.5), bn.bias.data.normal_()
    bn.running_mean.normal_(), bn.running_var.uniform_(0.5, 2.0)
    bn.train(training)
    rm, rv = bn.running_mean.clone().float().to(DEV), bn.running_var.clone().float().to(DEV)
    tracked = bn.num_batches_tracked.clone().to(DEV)
    # reference: the expanded batch, with one tagged copy of every table row appended to read y / dy at all rows
    xd = x.double().requires_grad_(True)
    rep = torch.repeat_interleave(torch.arange(R), counts.long())
    yb = bn(xd[rep])
    mean = xd[rep].mean(0) if training else bn.running_mean
    var = xd[rep].var(0, unbiased=False) if training else bn.running_var
    y_ref = (xd - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias
    dy = torch.randn(R, Fd)
    gx, gw, gb = torch.autograd.grad(y_ref, (xd, bn.weight, bn.bias), dy.double())
    xg = x.to(DEV).requires_grad_(True)
    w, b = bn.weight.detach().float().to(DEV).requires_grad_(True), bn.bias.detach().float().to(DEV).requires_grad_(True)
    y = ops.table_norm(xg, counts.to(DEV), w, b, rm, rv, tracked if training else None, training, 0.1, bn.eps)
    dx, dw, db = torch.autograd.grad(y, (xg, w, b), dy.to(DEV))
    assert rel_l2(y, y_ref) < 2e-6 and rel_l2(yb, y_ref[rep]) < 1e-12
    assert rel_l2(dx, gx) < 1e-5 and rel_l2(dw, gw) < 1e-5 and rel_l2(db, gb) < 1e-5
    assert rel_l2(rm, bn.running_mean) < 2e-6 and rel_l2(rv, bn.running_var) < 2e-6
    assert int(tracked) == int(bn.num_batches_tracked)


def test_graphed_embedding_two_forwards_before_backward():
    """the captured graphs serve one outstanding forward: a second forward before the first backward must not clobber
    the first one's tables (it takes the eager path) — gradients of the summed loss equal the eager model's."""
    torch.manual_seed(0)
    kw = dict(cutoff=5.0, cutoff_net="polynomial", emb_size=32, emb_size_coeff=32, emb_size_conv=32)
    m_e, m_g = LCAONet(**kw).to(DEV).train(), LCAONet(**kw).to(DEV).train()
    m_g.load_state_dict(m_e.state_dict())
    m_g.emb_layer.graph_tables = True
    ga, gb = qm9_like_batch(5, seed=61), qm9_like_batch(9, seed=62)
    for m in (m_e, m_g):
        oa = m(GraphBatch(ga).to(DEV))
        ob = m(GraphBatch(gb).to(DEV))
        ((oa**2).mean() + 3.0 * (ob**2).mean()).backward()
    for (n, p), q in zip(m_e.named_parameters(), m_g.parameters()):
        assert rel_l2(q.grad, p.grad) < 1e-5, n
    # ... and the next step replays the graphs again
    for m in (m_e, m_g):
        m.zero_grad(set_to_none=True)
        (m(GraphBatch(ga).to(DEV))**2).mean().backward()
    for (n, p), q in zip(m_e.named_parameters(), m_g.parameters()):
        assert rel_l2(q.grad, p.grad) < 1e-5, n
