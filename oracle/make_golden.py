"""Generate tests/golden/*.pt by running the UNMODIFIED reference (nmdl-mizo/lcaonet at
/root/reference) through the stand-ins in oracle/_shims.  TEST INFRASTRUCTURE.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden.py

Every case stores: constructor kwargs, the input graph (float32), the reference `state_dict`
(float32, torch.manual_seed(0) + the reference's own initialisers), and the reference's outputs
evaluated in float64 (`*_f64`) and float32 (`*_f32`): energies, forces, loss = mean(E^2)
[+ mean(F^2)], the gradient of the loss w.r.t. every parameter, and the BatchNorm running
statistics after the step (gradients are stored rounded to float32; for the float32 run only its
per-parameter rel-L2 distance to the float64 gradients is kept).  The float32 run is kept only to quote the reference's own distance to
float64 beside ours.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(HERE, "_shims"), "/root/reference"]

from lcaonet.atomistic import elec as ref_elec  # noqa: E402
from lcaonet.atomistic.info import ElecInfo as RefElecInfo  # noqa: E402
from lcaonet.model import LCAONet as RefLCAONet  # noqa: E402
from lcaonet.nn.cutoff import CosineCutoff, EnvelopeCutoff, PolynomialCutoff  # noqa: E402
from lcaonet.nn.rbf import HydrogenRadialBasis  # noqa: E402
from lcaonet.nn.shbf import SphericalHarmonicsBasis  # noqa: E402
from torch_geometric.data import Data  # noqa: E402

from lcaonet_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _run_case(name, kwargs, graph, training=True):
    res = {"kwargs": kwargs, "training": training,
           "graph": {k: v.detach().clone() for k, v in graph.items()}}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.manual_seed(0)
        model = RefLCAONet(**kwargs)
        if tag == "f64":
            res["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
        model = model.to(dtype)
        model.train(training)
        g = Data(**{k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in graph.items() if k != "y"})
        out = model(g)
        if isinstance(out, tuple):
            energy, forces = out
            loss = (energy**2).mean() + (forces**2).mean()
            res[f"forces_{tag}"] = forces.detach().clone()
        else:
            energy = out
            loss = (energy**2).mean()
        res[f"energy_{tag}"] = energy.detach().clone()
        res[f"loss_{tag}"] = loss.detach().clone()
        if isinstance(out, tuple) and tag == "f64":
            # gradient of the ENERGY term alone: what an implementation without double backward can be checked on
            ge = torch.autograd.grad((energy**2).mean(), list(model.parameters()), retain_graph=True, allow_unused=True)
            res["grads_energy_f64"] = {n: (v.to(torch.float32) if v is not None else None)
                                       for (n, _), v in zip(model.named_parameters(), ge)}
        loss.backward()
        grads = {n: (q.grad.detach().double() if q.grad is not None else None) for n, q in model.named_parameters()}
        if tag == "f64":
            g64 = grads
            res["grads_f64"] = {n: (v.to(torch.float32) if v is not None else None) for n, v in grads.items()}
        else:  # the reference's own float32 distance to float64, per parameter (rel-L2)
            res["grads_f32_rel_to_f64"] = {n: (float((v - g64[n]).norm() / (g64[n].norm() + 1e-300)) if v is not None
                                               else None) for n, v in grads.items()}
        res[f"grad_norms_{tag}"] = {n: (float(q.grad.double().norm()) if q.grad is not None else None)
                                    for n, q in model.named_parameters()}
        res[f"bn_after_{tag}"] = {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}
        if tag == "f64":
            res["triplets"] = {k: g[k].to(torch.int32) for k in ("idx_k_3b", "edge_idx_ks_3b", "edge_idx_st_3b")}
            res["edge_dist_f64"], res["edge_vec_f64"] = g["edge_dist"].detach().clone(), g["edge_vec"].detach().clone()
            res["angles_f32"] = g["angles_3b"].detach().to(torch.float32)
    torch.save(res, os.path.join(OUT, name + ".pt"))
    print(f"{name}: E={graph['edge_index'].shape[1]} T={res['triplets']['idx_k_3b'].numel()} "
          f"loss_f64={float(res['loss_f64']):.6g} |f32-f64|/|f64| energy="
          f"{float((res['energy_f32'].double() - res['energy_f64']).norm() / res['energy_f64'].norm()):.2e}")


def basis_tables():
    """Reference radial / angular / cutoff modules on fixed grids (float64 evaluation)."""
    out = {"r": torch.linspace(0.0, 10.0, 401, dtype=torch.float64),
           "c": torch.linspace(-1.0, 1.0, 201, dtype=torch.float64), "rb": {}, "shb": {}, "cut": {}}
    cuts = {"polynomial": PolynomialCutoff, "envelope": EnvelopeCutoff, "cosine": CosineCutoff}
    for rc in (2.0, 5.0, 6.0):
        for cname, ccls in cuts.items():
            out["cut"][(cname, rc)] = ccls(rc)(out["r"])
    for (max_z, max_orb, npo) in ((36, None, 1), (36, None, 2), (12, "3s", 1), (84, "6d", 1), (5, None, 1), (96, None, 1)):
        ei = RefElecInfo(max_z, max_orb, None, npo)
        for rc, cname in ((5.0, "polynomial"), (6.0, "envelope"), (3.0, "cosine")):
            rbf = HydrogenRadialBasis(rc, ei, cuts[cname](rc))
            out["rb"][(max_z, max_orb, npo, rc, cname)] = rbf(out["r"])
        out["shb"][(max_z, max_orb, npo)] = SphericalHarmonicsBasis(ei)(out["c"])
    # integral_norm=True (rbf.py:107-127): coefficients from scipy quad at constructor time
    out["rb_integral_norm"] = {}
    for rc, cname in ((5.0, "polynomial"), (6.0, "envelope"), (3.0, "cosine")):
        ei = RefElecInfo(36, None, None, 1)
        with torch.no_grad():
            out["rb_integral_norm"][(rc, cname)] = HydrogenRadialBasis(rc, ei, cuts[cname](rc), integral_norm=True)(out["r"])
    out["tables"] = {"ELEC_TABLE": ref_elec.ELEC_TABLE, "VALENCE_TABLE": ref_elec.VALENCE_TABLE,
                     "NL_LIST": ref_elec.NL_LIST, "MAX_ELEC_IDX": ref_elec.MAX_ELEC_IDX}
    torch.save(out, os.path.join(OUT, "basis_tables.pt"))
    print("basis_tables written")


def scheduler_table():
    """learning rates of the reference's WarmupCosineDecayAnnealingLR (train/scheduler.py) for two settings"""
    import torch.optim.lr_scheduler as L
    orig = L.LRScheduler.__init__
    L.LRScheduler.__init__ = lambda self, optimizer, last_epoch=-1, verbose=None: orig(self, optimizer, last_epoch)  # torch >= 2.7 dropped `verbose`
    from lcaonet.train.scheduler import WarmupCosineDecayAnnealingLR
    out = []
    for kw, lr0 in ((dict(num_epoch=60, num_warmup=5, T_max=7), 1e-5),
                    (dict(num_epoch=40, num_warmup=3, T_max=4, eta_min=1e-6, lr_max=5e-4, decay_coef=2.0), 1e-4)):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr0)
        sch = WarmupCosineDecayAnnealingLR(opt, **kw)
        lrs = []
        for _ in range(kw["num_epoch"]):
            opt.step()
            sch.step()
            lrs.append(opt.param_groups[0]["lr"])
        out.append({"kwargs": kw, "lr0": lr0, "lrs": lrs})
    L.LRScheduler.__init__ = orig
    torch.save(out, os.path.join(OUT, "scheduler_lrs.pt"))
    print("scheduler_lrs written")


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate

    def run_case(name, kwargs, graph, training=True):
        if not only or name in only:
            _run_case(name, kwargs, graph, training)

    if not only or "basis_tables" in only:
        basis_tables()
    if not only or "scheduler_lrs" in only:
        scheduler_table()
    qm9 = synth.qm9_like_batch(6, seed=3, cutoff=5.0, margin=0.05)
    xtl = synth.crystal_like_batch(1, seed=5, cutoff=6.0, margin=0.05)
    fix = synth.reference_fixture_graph()
    small = dict(emb_size=32, emb_size_coeff=32, emb_size_conv=32)
    run_case("qm9_default", dict(cutoff=5.0, cutoff_net="polynomial"), qm9)
    run_case("qm9_default_eval", dict(cutoff=5.0, cutoff_net="polynomial", **small), qm9, training=False)
    run_case("qm9_valence_ext_2perorb", dict(cutoff=5.0, cutoff_net="polynomial", add_valence=True, extend_orb=True,
                                             n_per_orb=2, **small), qm9)
    run_case("crystal_autograd_forces", dict(cutoff=6.0, cutoff_net="polynomial", regress_forces=True,
                                             direct_forces=False, **small), xtl)
    run_case("crystal_direct_forces_mean", dict(cutoff=6.0, cutoff_net="envelope", regress_forces=True,
                                                direct_forces=True, is_extensive=False, elec_to_node=False,
                                                emb_size=16, emb_size_coeff=32, emb_size_conv=16), xtl)
    run_case("fixture_small", dict(emb_size=16, emb_size_coeff=16, emb_size_conv=10, n_interaction=2, max_z=5,
                                   cutoff=6.0, cutoff_net="envelope"), fix)
    run_case("fixture_cosine_minmaxorb_atomref", dict(emb_size=16, emb_size_coeff=16, emb_size_conv=12, n_interaction=2,
                                                      max_z=5, cutoff=2.0, cutoff_net="cosine", max_orb="4p",
                                                      min_orb="2s", n_per_orb=2, add_valence=True,
                                                      atomref=torch.ones(6, 1), mean=torch.tensor([1.0])), fix)
    # activations other than the default (utils/resolve.py:65-76): the reference's own ShiftedSoftplus, and a torch.nn one
    qm9s = synth.qm9_like_batch(3, seed=7, cutoff=5.0, margin=0.05)
    run_case("qm9_shiftedsoftplus", dict(cutoff=5.0, cutoff_net="polynomial", activation="ShiftedSoftplus",
                                         regress_forces=True, direct_forces=True, **small), qm9s)
    run_case("qm9_gelu_valence", dict(cutoff=5.0, cutoff_net="polynomial", activation="gelu", add_valence=True,
                                      emb_size=16, emb_size_coeff=16, emb_size_conv=16), qm9s)
    # ---- the BASELINE.json configurations at their REAL widths (round-1 verdict: the benchmarked sizes had no
    #      reference-derived golden).  configs[0]: 32 QM9-like molecules, default model 128/128/128 x3.
    run_case("cfg1_qm9_32mol", dict(cutoff=5.0, cutoff_net="polynomial"), synth.qm9_like_batch(32, seed=0, cutoff=5.0, margin=0.05))
    # configs[2]: add_valence + extend_orb + n_per_orb=2, max_z=36 at width 128 (O = 16, C' = 256)
    run_case("cfg3_valence_width128", dict(cutoff=5.0, cutoff_net="polynomial", add_valence=True, extend_orb=True,
                                           n_per_orb=2, max_z=36), qm9)
    # configs[3]: periodic crystal cell (64 atoms, cutoff 6.0, ~49 neighbours/atom), energy + autograd forces at width 128
    run_case("cfg4_crystal_width128", dict(cutoff=6.0, cutoff_net="polynomial", regress_forces=True, direct_forces=False),
             xtl)
    # Swish with its trainable beta (nn/activation.py:7-33): beta's own gradient is part of the golden
    run_case("qm9_swish", dict(cutoff=5.0, cutoff_net="polynomial", activation="swish", regress_forces=True,
                               direct_forces=True, add_valence=True, **small), qm9s)
    # spherical-Bessel radial basis (rbf.py:145-182)
    run_case("qm9_sphericalbessel", dict(cutoff=5.0, cutoff_net="polynomial", rbf_type="sphericalbessel", **small), qm9s)


if __name__ == "__main__":
    main()
