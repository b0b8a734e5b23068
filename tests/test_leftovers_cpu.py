"""API-completeness pieces outside the BASELINE configs, pinned against vectors generated from the
unmodified reference (oracle/make_golden.py): the learning-rate schedule and the `integral_norm=True` radial basis."""
import math

import torch

from lcaonet_b200 import _lib
from lcaonet_b200.model import RadialBasis, _cutoff_value
from lcaonet_b200.orbitals import ElecInfo
from lcaonet_b200.scheduler import WarmupCosineDecayAnnealingLR
from tests._util import load_golden


def test_warmup_cosine_decay_annealing_matches_reference():
    for case in load_golden("scheduler_lrs"):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=case["lr0"])
        sch = WarmupCosineDecayAnnealingLR(opt, **case["kwargs"])
        lrs = []
        for _ in range(case["kwargs"]["num_epoch"]):
            opt.step()
            sch.step()
            lrs.append(opt.param_groups[0]["lr"])
        assert max(abs(a - b) / abs(b) for a, b in zip(lrs, case["lrs"])) < 1e-12
    import pytest
    with pytest.raises(ValueError):
        WarmupCosineDecayAnnealingLR(opt, num_epoch=3, num_warmup=3, T_max=2)
    with pytest.raises(ValueError):
        WarmupCosineDecayAnnealingLR(opt, num_epoch=5, num_warmup=1, T_max=2, decay_coef=0.0)


def test_integral_norm_coefficients_match_reference():
    """rbf.py:107-127: the constructor-time quadrature lands in lcao_basis_spec.norm; evaluated on the host in float64 it
    reproduces the reference module (whose own cutoff is evaluated in float32: agreement to ~1e-7)."""
    tab = load_golden("basis_tables")
    r = tab["r"]
    for (rc, cname), want in tab["rb_integral_norm"].items():
        sp = RadialBasis(rc, ElecInfo(36, None, None, 1), cname, "hydrogen", integral_norm=True).spec
        got = torch.zeros_like(want)
        for u in range(sp.n_unique):
            for k, rr in enumerate(r.tolist()):
                zeta = 2.0 / sp.n[u] / sp.a0 * rr
                poly = 0.0
                for i in range(sp.deg[u], -1, -1):
                    poly = poly * zeta + sp.poly[u][i]
                got[k, u] = _cutoff_value(sp.cutoff_kind, rr, rc) * sp.norm[u] * poly * zeta ** sp.l[u] * math.exp(-0.5 * zeta)
        assert float((got - want).abs().max() / want.abs().max()) < 5e-7, (rc, cname)
    assert _lib.RBF["hydrogen"] == 0
