"""Host-side logic of the data-parallel step (neurovit_b200/trainer.py) on CPU with the gloo backend,
world_size 2: batch sharding, flat-gradient bucketing, hook-driven bucket all-reduce, and equality of the
averaged shard gradients with the full-batch gradients. The model here is a plain torch module — the
trainer is model-agnostic; the CUDA model itself has no CPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurovit_b200.trainer import DataParallelTrainer, FlatGradBuckets, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.GELU(), torch.nn.LayerNorm(32),
                               torch.nn.Linear(32, 16), torch.nn.GELU(), torch.nn.Linear(16, 2))


def _worker(rank, world, port, q, defer=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(1)
        X = torch.randn(8, 12)
        Y = torch.randint(0, 2, (8,))
        m = _model()
        if rank != 0:   # replicas that start from DIFFERENT weights: the trainer must broadcast rank 0's
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(torch.randn_like(p))
        tr = DataParallelTrainer(m, optimizer=torch.optim.SGD(m.parameters(), lr=0.0), bucket_mb=0)
        assert len(tr.buckets.buckets) == len(list(m.parameters()))  # bucket_mb=0: one bucket per parameter
        tr.buckets.defer = defer  # True: what the graph-captured multi-rank step does — ONE all-reduce after backward
        loss = tr.step(shard_batch(X, rank, world), shard_batch(Y, rank, world))
        grads = [p.grad.clone() for p in m.parameters()]
        # every grad is still a view into the flat buffer, laid out in reverse parameter order, each slot
        # 128-byte aligned (the backward kernels red.add vectors into it)
        flat = tr.buckets.flat
        A = FlatGradBuckets.ALIGN
        off = 0
        for p in reversed(list(m.parameters())):
            assert p.grad.data_ptr() == flat.data_ptr() + 4 * off
            assert (p.grad.data_ptr() - flat.data_ptr()) % 128 == 0  # (CUDA allocations themselves are 512 B aligned)
            off += (p.numel() + A - 1) // A * A
        q.put((rank, float(loss), [g.numpy() for g in grads]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("defer", [False, True])
def test_dp_step_world2_gloo_matches_full_batch(defer):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, defer)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process full batch
    torch.manual_seed(1)
    X = torch.randn(8, 12)
    Y = torch.randint(0, 2, (8,))
    m = _model()
    torch.nn.functional.cross_entropy(m(X), Y).backward()
    full = [p.grad.numpy() for p in m.parameters()]
    for rank, loss, grads in res:
        for g, f in zip(grads, full):
            assert abs(g - f).max() < 1e-6, "averaged shard gradients must equal the full-batch gradients"
    assert abs(sum(r[1] for r in res) / world - float(torch.nn.functional.cross_entropy(_model()(X), Y))) < 1e-6


def test_bucket_countdown_rejects_a_second_report():
    """A parameter that reports twice in one step (shared weights / two forwards before one backward) must not
    release its bucket early: the set-based countdown raises once the bucket has already been reduced."""
    m = _model()
    b = FlatGradBuckets(list(m.parameters()), bucket_bytes=0)
    p0 = b.params[0]
    b.world = 2                      # pretend: the countdown is only live with several ranks
    fired = []
    b._reduce_bucket = lambda i: (fired.append(i), b.fired.__setitem__(i, True))
    b._on_grad(p0)
    assert fired == [b.bucket_of[p0]]
    with pytest.raises(RuntimeError):
        b._on_grad(p0)
    b.zero()
    b._on_grad(p0)                   # a new step starts clean
    assert fired == [b.bucket_of[p0]] * 2


def test_shard_batch_and_buckets_single_process():
    x = torch.arange(12).view(6, 2)
    assert torch.equal(shard_batch(x, 1, 3), x[2:4])
    with pytest.raises(ValueError):
        shard_batch(x, 0, 4)
    m = _model()
    b = FlatGradBuckets(list(m.parameters()), bucket_bytes=1 << 30)
    A = FlatGradBuckets.ALIGN
    assert len(b.buckets) == 1 and b.buckets[0][1] == sum((p.numel() + A - 1) // A * A for p in m.parameters())
    m(torch.randn(3, 12)).sum().backward()
    assert b.flat.abs().sum() > 0
    b.zero()
    assert b.flat.abs().sum() == 0 and all(p.grad.abs().sum() == 0 for p in m.parameters())
