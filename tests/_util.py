"""Shared helpers for the test-suite (golden loading, default constructor kwargs, error metrics)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CTOR_DEFAULTS = dict(emb_size=128, emb_size_coeff=128, emb_size_conv=128, out_size=1, n_interaction=3, n_per_orb=1,
                     cutoff=6.0, rbf_type="hydrogen", cutoff_net="envelope", max_z=36, min_orb=None, max_orb=None,
                     elec_to_node=True, add_valence=False, extend_orb=False, is_extensive=True, activation="SiLU",
                     weight_init="glorotorthogonal", atomref=None, mean=None, regress_forces=False, direct_forces=True)

CASES = ["qm9_default", "qm9_default_eval", "qm9_valence_ext_2perorb", "crystal_autograd_forces",
         "crystal_direct_forces_mean", "fixture_small", "fixture_cosine_minmaxorb_atomref", "qm9_shiftedsoftplus",
         "qm9_gelu_valence", "qm9_sphericalbessel", "qm9_swish", "cfg1_qm9_32mol", "cfg3_valence_width128", "cfg4_crystal_width128"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def full_cfg(kwargs):
    cfg = dict(CTOR_DEFAULTS)
    cfg.update(kwargs)
    return cfg


def graph_as(graph, dtype=None, device=None):
    out = {}
    for k, v in graph.items():
        if torch.is_tensor(v):
            if v.is_floating_point() and dtype is not None:
                v = v.to(dtype)
            if device is not None:
                v = v.to(device)
        out[k] = v
    return out


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def same_triplets_up_to_duplicate_order(a, b):
    """Triplet lists (k, e_ks, e_st) equal as multisets inside every (e_st, k) group — the reference's
    order among periodic-image duplicates of one (k, s) pair is implementation-defined."""
    def canon(k, e_ks, e_st):
        k, e_ks, e_st = k.long().cpu(), e_ks.long().cpu(), e_st.long().cpu()
        key = (e_st * (int(k.max()) + 1 if k.numel() else 1) + k)
        order = torch.argsort(key * (int(e_ks.max()) + 1 if e_ks.numel() else 1) + e_ks, stable=True)
        return k[order], e_ks[order], e_st[order], key
    ka, ea, sa, keya = canon(*a)
    kb, eb, sb, keyb = canon(*b)
    return (torch.equal(ka, kb) and torch.equal(ea, eb) and torch.equal(sa, sb)
            and torch.equal(a[2].long().cpu(), b[2].long().cpu()) and torch.equal(a[0].long().cpu(), b[0].long().cpu()))
