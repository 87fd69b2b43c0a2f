"""world_size-2 gloo tests of the multi-rank plumbing (scene sharding, max-over-ranks timing,
whole-job throughput aggregation, result gather) -- SURVEY.md section 8(e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msmdfusion_b200 import dist_utils


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r, lr, w = dist_utils.init(backend='gloo')
    assert (r, w) == (rank, world)
    scenes = dist_utils.scene_shard(8, rank, world)
    # pretend each scene took (rank+1) ms on this rank
    elapsed = float(len(scenes) * (rank + 1))
    thr, total, ms = dist_utils.aggregate_throughput(len(scenes), elapsed)
    mx = dist_utils.max_over_ranks([elapsed, 1.0 + rank])
    gathered = dist_utils.gather_objects({'rank': rank, 'scenes': scenes})
    dist_utils.barrier()
    if rank == 0:
        out.put(dict(thr=thr, total=total, ms=ms, mx=mx, gathered=gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res['total'] == 8.0                       # every scene processed exactly once
    assert res['ms'] == 8.0 and res['mx'] == [8.0, 2.0]  # max over ranks, never the local value
    assert abs(res['thr'] - 8.0 / 8e-3) < 1e-6       # whole-job units / max time
    got = sorted(s for g in res['gathered'] for s in g['scenes'])
    assert got == list(range(8))
    assert res['gathered'][1]['scenes'] == [1, 3, 5, 7]


def test_single_process_is_a_noop():
    assert dist_utils.scene_shard(3, 0, 1) == [0, 1, 2]
    assert dist_utils.max_over_ranks([3.0]) == [3.0]
    thr, total, ms = dist_utils.aggregate_throughput(4, 2.0)
    assert (thr, total, ms) == (2000.0, 4.0, 2.0)


# --------------------------------------------------------------------------------------
# train step (config 5): the gradient exchange -- ONE all-reduce over a flat buffer
# --------------------------------------------------------------------------------------
def _train_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from msmdfusion_b200 import train
    dist_utils.init(backend='gloo')
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    unused = torch.nn.Linear(4, 4)                         # never called: find_unused_parameters semantics
    params = list(net.parameters())
    flat = train.FlatGradients(params)
    x = torch.randn(7, 6, generator=torch.Generator().manual_seed(10 + rank))   # different data per rank
    flat.zero()
    net(x).pow(2).mean().backward()
    assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in params), 'grads must stay views'
    local = flat.flat.clone()
    flat.all_reduce_mean()
    gathered = dist_utils.gather_objects(local)
    norm = flat.clip_(0.05)
    if rank == 0:
        out.put(dict(mean=flat.flat.clone(), locals=gathered, norm=float(norm), unused_grad=unused.weight.grad))
    dist_utils.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_all_reduce():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    avg = (res['locals'][0] + res['locals'][1]) / 2
    assert not torch.equal(res['locals'][0], res['locals'][1])
    assert abs(res['norm'] - float(avg.norm())) < 1e-6
    expect = avg * min(1.0, 0.05 / (float(avg.norm()) + 1e-6))
    assert torch.allclose(res['mean'], expect, atol=1e-7)
    assert res['unused_grad'] is None
