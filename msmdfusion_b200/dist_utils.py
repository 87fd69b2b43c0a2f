"""Multi-GPU plumbing of the hot path: scenes are independent, so the
forward path shards scenes across ranks with NO data-path collective; the only communication
is the timing reduction (max over ranks) and an optional result gather for evaluation, as in
``tools/test.py:186-195`` of the reference.  One process per GPU; ``nccl`` on GPUs, ``gloo``
in the CPU tests."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for 1 rank)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl' and device is not None:
            kwargs['device_id'] = device
        dist.init_process_group(backend, **kwargs)
    return rank, local_rank, world


def scene_shard(num_scenes, rank, world):
    """Scene ids handled by ``rank``: scene i -> rank i mod world (BASELINE: one scene per GPU)."""
    return list(range(rank, num_scenes, world))


def max_over_ranks(values, device='cpu'):
    """Element-wise max of a list of floats over all ranks (device-timed milliseconds)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device='cpu'):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def aggregate_throughput(units_this_rank, elapsed_ms_this_rank, device='cpu'):
    """Whole-job throughput = units all ranks processed / max-over-ranks time (units per second)."""
    total = sum_over_ranks([units_this_rank], device)[0]
    ms = max_over_ranks([elapsed_ms_this_rank], device)[0]
    return total / (ms * 1e-3), total, ms


def gather_objects(obj, dst=0):
    """Collect per-rank python objects on ``dst`` (evaluation only; not on the timed path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
