"""Data-parallel plumbing on CPU: world_size 2, gloo backend.  One flat gradient bucket, one
all-reduce per step, rank-strided pair shards, weights broadcast once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        from deepatlas_b200.dist import FlatGradBucket, broadcast_parameters, init_from_env, shard_pairs
        r, l, w = init_from_env("gloo")
        assert (r, w) == (rank, world)
        torch.manual_seed(100 + rank)                      # different initial weights per rank on purpose
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
        broadcast_parameters(net, 0)
        flat_w = torch.cat([p.detach().flatten() for p in net.parameters()])
        gathered = [torch.empty_like(flat_w) for _ in range(world)]
        dist.all_gather(gathered, flat_w)
        assert all(torch.equal(gathered[0], g) for g in gathered), "weights differ after the broadcast"
        bucket = FlatGradBucket(net.parameters())
        assert bucket.nbytes == 4 * sum(p.numel() for p in net.parameters())
        for p in net.parameters():                         # p.grad are views into ONE contiguous buffer
            assert p.grad.data_ptr() >= bucket.flat.data_ptr() and p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.nbytes
        calls = []
        orig = dist.all_reduce
        dist.all_reduce = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        bucket.zero()
        x = torch.randn(4, 5, generator=torch.Generator().manual_seed(7 + rank))
        net(x).square().sum().backward()                   # autograd accumulates straight into the bucket
        local = bucket.flat.clone()
        bucket.allreduce(world)
        dist.all_reduce = orig
        assert len(calls) == 1, "exactly one collective per step"
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        assert torch.allclose(bucket.flat, sum(both) / world, rtol=1e-6, atol=1e-7)
        shard = shard_pairs(6, rank, world, seed=230, epoch=3)
        q.put((rank, "ok", shard))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), None))


@pytest.mark.timeout(180)
def test_flat_bucket_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    for rank, status, _ in res:
        assert status == "ok", f"rank {rank}: {status}"
    shards = {rank: s for rank, _, s in res}
    allp = shards[0] + shards[1]
    assert len(allp) == 6 * 5 and len(set(allp)) == 30          # disjoint cover of the N(N-1) ordered pairs
    assert all(m != f and 0 <= m < 6 and 0 <= f < 6 for m, f in allp)


def test_shard_pairs_follows_reference_enumeration():
    """Pair id -> (fixed, moving) as lib/datasets.py:344-359: fixed = id // (N-1); moving = id % (N-1), +1 if >= fixed."""
    from deepatlas_b200.dist import shard_pairs
    N = 5
    one = shard_pairs(N, 0, 1, seed=1)
    expect = set()
    for pid in range(N * (N - 1)):
        f, m = pid // (N - 1), pid % (N - 1)
        expect.add((m + 1 if m >= f else m, f))
    assert set(one) == expect
    assert shard_pairs(N, 1, 4, seed=1) == one[1::4]
    assert shard_pairs(N, 0, 1, seed=1, epoch=1) != one


def test_single_process_is_a_no_op():
    from deepatlas_b200.dist import FlatGradBucket
    lin = torch.nn.Linear(3, 2)
    b = FlatGradBucket(lin.parameters())
    lin(torch.ones(1, 3)).sum().backward()
    before = b.flat.clone()
    b.allreduce()
    assert torch.equal(before, b.flat) and float(before.abs().sum()) > 0
    b.zero()
    assert float(lin.weight.grad.abs().sum()) == 0.0


def test_shard_pairs_equal_lengths_on_every_rank():
    """5 volumes = 20 ordered pairs over 8 ranks: every rank must run the same number of steps (one all-reduce each),
    otherwise the job hangs in NCCL at the end of the epoch.  Padding wraps around like DistributedSampler."""
    from deepatlas_b200.dist import shard_pairs
    N, world = 5, 8
    shards = [shard_pairs(N, r, world, seed=230) for r in range(world)]
    assert {len(s) for s in shards} == {3}                      # ceil(20 / 8)
    allp = [p for s in shards for p in s]
    assert len(set(allp)) == 20                                  # every pair is still visited
    dropped = [shard_pairs(N, r, world, seed=230, drop_last=True) for r in range(world)]
    assert {len(s) for s in dropped} == {2} and len({p for s in dropped for p in s}) == 16
    with pytest.raises(ValueError):
        shard_pairs(N, 8, world)


def test_bucket_survives_optimizer_zero_grad_set_to_none():
    """The reference trainer calls optimizer.zero_grad() (models/segmentation.py:142), default set_to_none=True: the
    next backward allocates gradients outside the bucket.  allreduce() must find them, copy them in and re-bind."""
    from deepatlas_b200.dist import FlatGradBucket
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    b = FlatGradBucket(lin.parameters())
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    opt.zero_grad()                                              # set_to_none=True: p.grad is None now
    assert lin.weight.grad is None
    lin(torch.ones(1, 3)).sum().backward()                       # fresh gradient tensors, not views of the bucket
    expect = torch.cat([p.grad.flatten().clone() for p in lin.parameters()])
    assert float(b.flat.abs().sum()) == 0.0
    b.allreduce()
    assert b.rebound == 2
    assert torch.equal(b.flat, expect)
    for p in lin.parameters():
        assert b.flat.data_ptr() <= p.grad.data_ptr() < b.flat.data_ptr() + b.nbytes
    b.zero()                                                     # the bucket's own zero keeps the views
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.flat, expect) and b.rebind() == 0
