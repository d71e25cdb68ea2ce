"""Scene sharding for multi-GPU forward throughput.

The hot path shards by scene with no data-path collective (forward has no cross-sample
dependency; the reference does the same with one process per GPU and a DistributedSampler,
det3d/datasets/loader/build_loader.py:35-36).  The only exchanges are control-plane: a barrier around the timed
region and a MAX / SUM reduction of (elapsed time, scenes processed) so that rank 0 can report whole-job throughput.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Contiguous-strided partition used for scenes: item i goes to rank i % world_size (every item exactly once)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_items, world_size))


def scene_seed(rank, slot, batch_index, batch):
    """Distinct synthetic-scene seed per (rank, pool slot, batch element) so that no two replicas see the same cloud."""
    return rank * 1000 + slot * batch + batch_index


def reduce_throughput(local_ms, local_units, device=None):
    """Whole-job throughput: sum of units over ranks / max elapsed time over ranks.  Works on gloo and nccl."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_units) / (float(local_ms) / 1e3), float(local_ms), float(local_units)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / (float(t.item()) / 1e3), float(t.item()), float(u.item())
