"""Run the oracle against the UNMODIFIED reference in-process (build container only: the reference
tree does not exist on the GPU box, where this module skips)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lcaonet")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_model_cls():
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_shims")
    sys.path[:0] = [shims, REF]
    try:
        from lcaonet.model import LCAONet
        yield LCAONet
    finally:
        sys.path.remove(shims)
        sys.path.remove(REF)


@pytest.mark.parametrize("kwargs,kind", [
    (dict(cutoff=5.0, cutoff_net="polynomial", emb_size=32, emb_size_coeff=32, emb_size_conv=32), "qm9"),
    (dict(cutoff=5.0, cutoff_net="polynomial", emb_size=16, emb_size_coeff=24, emb_size_conv=20, add_valence=True,
          max_orb="5s", min_orb="2p", out_size=2, is_extensive=False), "qm9"),
    (dict(cutoff=6.0, cutoff_net="envelope", emb_size=16, emb_size_coeff=16, emb_size_conv=16, regress_forces=True,
          direct_forces=True), "crystal"),
])
def test_fresh_random_case(ref_model_cls, kwargs, kind):
    from torch_geometric.data import Data

    from lcaonet_b200 import synth
    from oracle import lcao_oracle as O
    from tests._util import full_cfg, graph_as, rel_l2

    g = synth.qm9_like_batch(5, seed=11, margin=0.05) if kind == "qm9" else synth.crystal_like_batch(1, seed=7, margin=0.05)
    torch.manual_seed(1)
    ref = ref_model_cls(**kwargs).double()
    gd = Data(**{k: v for k, v in graph_as(g, torch.float64).items() if k != "y"})
    out_r = ref(gd)
    params = O.cast_params(ref.state_dict(), torch.float64)
    out_o = O.forward(params, full_cfg(kwargs), graph_as(g, torch.float64), training=True)
    if isinstance(out_r, tuple):
        assert rel_l2(out_o[1], out_r[1]) < 1e-10
        out_r, out_o = out_r[0], out_o[0]
    assert rel_l2(out_o, out_r) < 1e-11
