"""Pin the CPU oracle (oracle/lcao_oracle.py) against the committed golden vectors produced by the
unmodified reference (oracle/make_golden.py) and against the known-answer vectors the reference's
own tests hold (tests/nn/test_rbf.py:37-48, test_shbf.py:46, test_cutoff.py, conftest.py:6-35)."""
import math

import numpy as np
import pytest
import scipy.special
import torch

from oracle import lcao_oracle as O
from tests._util import CASES, full_cfg, graph_as, load_golden, rel_l2, same_triplets_up_to_duplicate_order

A0 = 0.529
CLOSED_FORM = {  # textbook hydrogen R_nl (same table as the reference test)
    (1, 0): lambda r: (1 / A0) ** 1.5 * np.exp(-r / A0) * 2,
    (2, 0): lambda r: (1 / A0) ** 1.5 * (2 - r / A0) * np.exp(-r / 2 / A0) / 2 / np.sqrt(2),
    (2, 1): lambda r: (1 / A0) ** 1.5 * r / A0 * np.exp(-r / 2 / A0) / 2 / np.sqrt(6),
    (3, 0): lambda r: (1 / A0) ** 1.5 * (27 - 18 * r / A0 + 2 * r**2 / A0**2) * np.exp(-r / 3 / A0) * 2 / 81 / np.sqrt(3),
    (3, 1): lambda r: (1 / A0) ** 1.5 * (6 - r / A0) * r / A0 * np.exp(-r / 3 / A0) * 4 / 81 / np.sqrt(6),
    (3, 2): lambda r: (1 / A0) ** 1.5 * (r / A0) ** 2 * np.exp(-r / 3 / A0) * 4 / 81 / np.sqrt(30),
    (4, 0): lambda r: (1 / A0) ** 1.5 * (192 - 144 * r / A0 + 24 * r**2 / A0**2 - r**3 / A0**3) * np.exp(-r / 4 / A0) / 768,
    (4, 1): lambda r: (1 / A0) ** 1.5 * (80 - 20 * r / A0 + r**2 / A0**2) * r / A0 * np.exp(-r / 4 / A0) / 256 / np.sqrt(15),
    (4, 2): lambda r: (1 / A0) ** 1.5 * (12 - r / A0) * (r / A0) ** 2 * np.exp(-r / 4 / A0) / 768 / np.sqrt(5),
    (4, 3): lambda r: (1 / A0) ** 1.5 * (r / A0) ** 3 * np.exp(-r / 4 / A0) / 768 / np.sqrt(35),
}


@pytest.mark.parametrize("cutoff,max_z,max_orb,npo", [(1.0, 12, None, 1), (3.0, 12, "3s", 2), (1.0, 36, None, 1),
                                                      (3.0, 36, "6s", 1), (3.0, 84, "6d", 2), (3.0, 84, None, 1)])
def test_radial_basis_closed_form(cutoff, max_z, max_orb, npo):
    r = torch.linspace(0, 10, 200, dtype=torch.float64)
    nl = O.orbital_quantum_numbers(max_z, max_orb, npo)
    rb = O.radial_basis(r, nl, cutoff, "envelope")
    cw = O.cutoff_fn("envelope", r, cutoff).numpy()
    for i, key in enumerate(nl):
        if key in CLOSED_FORM:
            np.testing.assert_allclose(rb[:, i].numpy(), CLOSED_FORM[key](r.numpy()) * cw, rtol=1e-5, atol=1e-7)


def test_laguerre_polynomials_match_survey_listing():
    assert O.laguerre_poly(1, 0) == [-1]
    assert O.laguerre_poly(2, 0) == [-4, 2]
    assert O.laguerre_poly(3, 0) == [-18, 18, -3]
    assert O.laguerre_poly(4, 1) == [-1200, 600, -60]
    assert O.laguerre_poly(6, 2) == [-2257920, 1128960, -161280, 6720]
    assert O.laguerre_poly(7, 0) == [-35280, 105840, -88200, 29400, -4410, 294, -7]


@pytest.mark.parametrize("max_z,max_orb,npo", [(12, None, 1), (36, None, 2), (84, "6d", 1)])
def test_angular_basis_vs_scipy(max_z, max_orb, npo):
    c = torch.linspace(0, 2 * math.pi, 200, dtype=torch.float64).cos()
    nl = O.orbital_quantum_numbers(max_z, max_orb, npo)
    shb = O.angular_basis(c, nl)
    for i, (_, l) in enumerate(nl):
        ref = scipy.special.sph_harm_y(l, 0, np.arccos(np.clip(c.numpy(), -1, 1)), 0.0).real
        np.testing.assert_allclose(shb[:, i].numpy(), ref, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("kind", ["polynomial", "envelope", "cosine"])
def test_cutoff_properties(kind):
    r = torch.linspace(0, 6, 601, dtype=torch.float64)
    v = O.cutoff_fn(kind, r, 3.0)
    assert torch.all(v[r > 3.0] == 0) and torch.all(v >= -1e-12)
    inside = v[r <= 3.0]
    assert torch.all(inside[1:] <= inside[:-1] + 1e-12) and abs(float(inside[0]) - 1.0) < 1e-12


def test_basis_tables_match_reference_dump():
    t = load_golden("basis_tables")
    for (cname, rc), ref in t["cut"].items():
        assert torch.allclose(O.cutoff_fn(cname, t["r"], rc), ref, rtol=1e-12, atol=1e-14)
    for (max_z, max_orb, npo, rc, cname), ref in t["rb"].items():
        nl = O.orbital_quantum_numbers(max_z, max_orb, npo)
        got = O.radial_basis(t["r"], nl, rc, cname)
        scale = ref.abs().max(0).values.clamp(min=1e-300)
        assert float(((got - ref).abs() / scale).max()) < 1e-9, (max_z, max_orb, npo, rc, cname)
    for (max_z, max_orb, npo), ref in t["shb"].items():
        nl = O.orbital_quantum_numbers(max_z, max_orb, npo)
        assert torch.allclose(O.angular_basis(t["c"], nl), ref, rtol=1e-12, atol=1e-13)


def test_fixture_triplets_golden_vector():
    """SURVEY.md §8 a-1 golden vector for the reference fixture (conftest.py:6-35)."""
    ei = torch.tensor([[0, 0, 1, 1, 2, 2, 2, 2], [1, 2, 0, 2, 0, 1, 1, 1]])
    k, e_ks, e_st = O.triplets(ei, 3)
    assert k.tolist() == [1, 2, 1, 2, 0, 2, 2, 2, 0, 2, 2, 2, 0, 1, 0, 1, 0, 1, 0, 1]
    assert e_ks.tolist() == [2, 4, 2, 4, 0, 5, 6, 7, 0, 5, 6, 7, 1, 3, 1, 3, 1, 3, 1, 3]
    assert e_st.tolist() == [0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7]


def test_triplets_drop_only_self_loop_edges():
    ei = torch.tensor([[0, 0, 1, 1], [0, 1, 0, 1]])  # two self-loops
    k, e_ks, e_st = O.triplets(ei, 2)
    assert e_st.tolist() == [0, 1, 1, 2, 2, 3] and e_ks.tolist() == [2, 0, 2, 1, 3, 1]
    k, e_ks, e_st = O.triplets(torch.zeros(2, 0, dtype=torch.long), 4)
    assert k.numel() == e_ks.numel() == e_st.numel() == 0


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name):
    gold = load_golden(name)
    cfg = full_cfg(gold["kwargs"])
    g = graph_as(gold["graph"], torch.float64)
    params = O.cast_params(gold["state_dict"], torch.float64, requires_grad=True)
    trace, stats = {}, {}
    out = O.forward(params, cfg, g, training=gold["training"], trace=trace, new_stats=stats)
    trip = gold["triplets"]
    assert same_triplets_up_to_duplicate_order(
        (trace["tri_k"], trace["e_ks"], trace["e_st"]), (trip["idx_k_3b"], trip["edge_idx_ks_3b"], trip["edge_idx_st_3b"]))
    assert torch.allclose(trace["dist"], gold["edge_dist_f64"], rtol=1e-13, atol=1e-13)
    if isinstance(out, tuple):
        energy, forces = out
        assert rel_l2(forces, gold["forces_f64"]) < 1e-9
        loss = (energy**2).mean() + (forces**2).mean()
    else:
        energy, loss = out, (out**2).mean()
    assert rel_l2(energy, gold["energy_f64"]) < 1e-11
    loss.backward()
    for n, ref in gold["grads_f64"].items():
        if ref is None:
            continue
        assert rel_l2(params[n].grad, ref) < 2e-6, n  # golden grads are stored rounded to float32
    for k, v in gold["bn_after_f64"].items():
        if gold["training"]:
            assert torch.allclose(stats[k], v, rtol=1e-10, atol=1e-12), k
