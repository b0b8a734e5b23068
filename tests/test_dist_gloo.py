"""Data-parallel plumbing on CPU: world_size-2 `gloo` processes.  The molecule batch is
sharded by `shard_range`, every rank runs the drop-in model through the C-ABI emulator on its shard, and
the flat gradient bucket is all-reduced once.  The averaged gradients must equal the gradients of the
mean-of-shard-losses computed in a single process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lcaonet_b200.dist import FlatGradBucket, broadcast_module, shard_range

KW = dict(emb_size=16, emb_size_coeff=16, emb_size_conv=16, n_interaction=2, cutoff=5.0, cutoff_net="polynomial")
N_MOL = 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_loss(model, rank, world):
    from lcaonet_b200.synth import qm9_like_batch
    lo, hi = shard_range(N_MOL, rank, world)
    g = qm9_like_batch(hi - lo, seed=100 + rank, margin=0.05)  # each rank owns its own molecules
    out = model(g)
    return (out**2).mean()


def _install_emulator():
    from tests import cpu_abi

    class _MP:  # minimal stand-in for pytest's monkeypatch inside the worker processes
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)

    cpu_abi.install(_MP)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lcaonet_b200 import LCAONet
        _install_emulator()
        torch.manual_seed(rank)  # different initial weights per rank: the broadcast must make them equal
        model = LCAONet(**KW).train()
        broadcast_module(model)
        bucket = FlatGradBucket(model)
        bucket.zero()
        _shard_loss(model, rank, world).backward()
        bucket.all_reduce_mean()
        q.put((rank, bucket.flat.clone(), torch.cat([p.detach().reshape(-1) for p in model.parameters()])))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process(monkeypatch):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, g0, w0), (_, g1, w1) = res
    assert torch.equal(w0, w1)          # parameters were broadcast from rank 0
    assert torch.equal(g0, g1)          # every rank holds the same averaged gradient
    # single-process reference: mean over ranks of the per-shard losses, rank 0's initial weights
    from lcaonet_b200 import LCAONet
    from tests import cpu_abi
    cpu_abi.install(monkeypatch)  # (undone at test exit; the workers patched their own processes)
    torch.manual_seed(0)
    model = LCAONet(**KW).train()
    loss = sum(_shard_loss(model, r, world) for r in range(world)) / world
    loss.backward()
    ref = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in model.parameters()])
    assert float((g0 - ref).norm() / ref.norm()) < 1e-5
