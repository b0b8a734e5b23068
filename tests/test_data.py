"""Collation / on-disk samples: `collate` follows PyG's `Batch.from_data_list` rules for the
reference's key set (index keys shifted by the atom offset, per-structure attributes stacked)."""
import os

import pytest
import torch

from lcaonet_b200.data import GraphDataset, collate, sample_from_structure, split
from lcaonet_b200.synth import crystal_like_batch, qm9_like_batch


@pytest.mark.parametrize("maker", [lambda: qm9_like_batch(7, seed=2), lambda: crystal_like_batch(3, seed=1)])
def test_split_then_collate_is_identity(maker):
    g = maker()
    samples = split(g)
    assert len(samples) == g["lattice"].shape[0]
    assert all(int(s["edge_index"].min()) >= 0 and int(s["edge_index"].max()) < s["z"].shape[0] for s in samples)
    back = collate(samples)
    assert set(back.keys()) == set(g.keys())
    for k, v in g.items():
        assert torch.equal(back[k], v), k


def test_collate_offsets_and_stacking():
    a = sample_from_structure([1, 8, 1], torch.zeros(3, 3), cell=torch.eye(3), pbc=[True, False, True], y=[1.5])
    b = sample_from_structure([6, 6], torch.ones(2, 3), y=[2.5])
    a["edge_index"], b["edge_index"] = torch.tensor([[0, 1, 2], [1, 0, 1]]), torch.tensor([[0, 1], [1, 0]])
    a["edge_shift"], b["edge_shift"] = torch.zeros(3, 3), torch.zeros(2, 3)
    g = collate([a, b])
    assert g["edge_index"].tolist() == [[0, 1, 2, 3, 4], [1, 0, 1, 4, 3]]
    assert g["batch"].tolist() == [0, 0, 0, 1, 1] and g["lattice"].shape == (2, 3, 3) and g["pbc"].tolist() == [[1, 0, 1], [0, 0, 0]]
    assert g["y"].shape == (2, 1) and g["z"].tolist() == [1, 8, 1, 6, 6]
    with pytest.raises(ValueError):
        collate([])


def test_dataset_reads_the_reference_layout(tmp_path):
    samples = split(qm9_like_batch(4, seed=5))
    for i, s in enumerate(samples):
        torch.save(dict(s), os.path.join(tmp_path, f"{i}.pt"))
    for inmemory in (False, True):
        ds = GraphDataset(str(tmp_path), inmemory=inmemory)
        assert len(ds) == 4
        g = collate([ds[i] for i in range(4)])
        assert torch.equal(g["edge_index"], qm9_like_batch(4, seed=5)["edge_index"])
        with pytest.raises(IndexError):
            ds[4]
    with pytest.raises(FileNotFoundError):
        GraphDataset(os.path.join(tmp_path, "missing"))


@pytest.mark.parametrize("maker", [lambda: qm9_like_batch(9, seed=4), lambda: crystal_like_batch(3, seed=2)])
def test_sample_pool_collates_like_the_host_loop(maker):
    from lcaonet_b200.data import SamplePool, collate_window
    g = maker()
    samples = split(g)
    for pool in (SamplePool(samples), SamplePool.from_batch(g)):
        assert len(pool) == len(samples)
        ids = torch.tensor([len(samples) - 1, 0, 2, 2])
        got, want = pool.batch(ids), collate([samples[i] for i in ids.tolist()])
        assert set(got.keys()) == set(want.keys())
        for k, v in want.items():
            assert torch.equal(got[k], v), k
        win = collate_window(pool.window(1, 3))
        want = collate(samples[1:3])
        for k, v in want.items():
            assert torch.equal(win[k], v), k
