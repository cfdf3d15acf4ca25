"""world_size-2 gloo tests (CPU) of the multi-rank host logic: utterance sharding and the flat gradient
all-reduce of the training step.  The GPU kernels are not involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dl4ss_b200.training import allreduce_gradients, shard_range
        torch.manual_seed(0)                                   # same replica on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
        xs = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0          # global batch of 7 "utterances"
        ys = torch.ones(7, 3)
        lo, hi = shard_range(7, rank, world)
        # loss normalised by the GLOBAL element count, as TrainStep does
        loss = ((model(xs[lo:hi]) - ys[lo:hi]) ** 2).sum() / (7 * 3)
        loss.backward()
        nbytes = allreduce_gradients(list(model.parameters()))
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        # the same through the persistent segmented bucket TrainStep.step uses: gradients are views, segments are
        # reduced asynchronously in completion order, nothing is copied
        from dl4ss_b200.training import GradBucket, global_batch_size
        assert global_batch_size(hi - lo, torch.device('cpu')) == 7
        bucket = GradBucket([list(model[1].parameters()), list(model[0].parameters())])
        ptr = bucket.flat.data_ptr()
        for _ in range(2):                                     # two steps: the buffer and the views persist
            bucket.zero()
            loss = ((model(xs[lo:hi]) - ys[lo:hi]) ** 2).sum() / (7 * 3)
            loss.backward()
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
            n2 = bucket.reduce_async(0) + bucket.reduce_async(1)
            bucket.wait()
        assert bucket.flat.data_ptr() == ptr and n2 == nbytes == bucket.nbytes
        flat2 = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        assert torch.allclose(flat2, flat, atol=1e-7)
        q.put((rank, lo, hi, nbytes, flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_equal_global_batch():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # reference: the un-sharded batch on one process
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    xs = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
    loss = ((model(xs) - torch.ones(7, 3)) ** 2).mean()
    loss.backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 4, 4, 7)          # contiguous shards, extras first
    for rank, lo, hi, nbytes, flat in res:
        assert nbytes == want.numel() * 4
        assert torch.allclose(torch.tensor(flat), want, atol=1e-6)


def test_shard_range_covers_batch():
    from dl4ss_b200.training import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_allreduce_is_noop_without_process_group():
    from dl4ss_b200.training import allreduce_gradients
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    g = lin.weight.grad.clone()
    assert allreduce_gradients(list(lin.parameters())) == 0
    assert torch.equal(lin.weight.grad, g)
