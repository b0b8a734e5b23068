"""The packed orbital tables must reproduce the reference tables digit for digit (golden dump)."""
import pytest
import torch

from lcaonet_b200 import orbitals
from tests._util import load_golden


def test_tables_equal_reference_dump():
    t = load_golden("basis_tables")["tables"]
    for name in ("ELEC_TABLE", "VALENCE_TABLE", "NL_LIST", "MAX_ELEC_IDX"):
        assert torch.equal(getattr(orbitals, name), t[name]), name


@pytest.mark.parametrize("args,n_orb", [((36, None, None, 1), 8), ((36, None, None, 2), 16), ((12, "3s", None, 2), 8),
                                        ((84, "6d", "2s", 1), 18), ((5, None, None, 1), 3)])
def test_elec_info_shapes(args, n_orb):
    ei = orbitals.ElecInfo(*args)
    assert ei.n_orb == n_orb
    assert ei.elec_table.shape == (args[0] + 1, n_orb) and ei.valence_table.shape == (args[0] + 1, n_orb)
    assert ei.nl_list.shape == (n_orb, 2) and ei.max_elec_idx.shape == (n_orb,)


def test_elec_info_errors():
    with pytest.raises(ValueError):
        orbitals.ElecInfo(0, None)
    with pytest.raises(ValueError):
        orbitals.ElecInfo(97, None)
    with pytest.raises(ValueError):
        orbitals.ElecInfo(10, "9z")
