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


class GradSync:
    """Data-parallel gradient exchange of the training step: ONE all-reduce(sum)/world over the parameter gradients
    per step, as torch DDP does for the reference (det3d/torchie/apis/train.py:311-317); the reference's redundant
    second all-reduce (det3d/torchie/apis/dist_utils.py:51-57) and apex SyncBN are not reproduced.

    Gradients already live in flat buckets (train.GradBuckets), in the order backward finalises them, so there is no
    pack/unpack copy: as soon as the last gradient of a bucket is written its all-reduce is launched asynchronously
    (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests) and overlaps the rest of backward;
    `finish()` waits for the outstanding reductions.  On NCCL the averaging is part of the collective
    (ReduceOp.AVG: no extra elementwise launch per bucket); gloo has no AVG, so the CPU tests scale after the wait.

    `inline=True` (CUDA-graph capture of a multi-rank step): every bucket's all-reduce is enqueued synchronously on
    the current stream the moment the bucket is final, so the NCCL kernels become nodes of the captured graph and a
    replayed step involves no Python at all; nothing is left for finish() to wait for.
    """

    def __init__(self, buckets, group=None, inline=False):
        self.buckets, self.group, self.inline = buckets, group, inline
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.avg = self.world > 1 and dist.get_backend(group) == "nccl"
        self.handles = []
        self.bytes_reduced = 0
        buckets.on_bucket_ready = self._launch

    def _launch(self, index, flat):
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        if self.inline:
            dist.all_reduce(flat, op=op, group=self.group)
            if not self.avg:
                flat.mul_(1.0 / self.world)
        else:
            self.handles.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat))
        self.bytes_reduced += flat.numel() * 4

    def finish(self):
        for h, flat in self.handles:
            h.wait()
            if not self.avg:
                flat.mul_(1.0 / self.world)
        self.handles = []


def pin_rank_to_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores (8 launch-heavy Python processes plus NCCL proxy
    threads sharing all cores migrate and contend; measured on the 8-GPU box).  Returns the cores or None."""
    import os
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(local_world, 1)
        if per < 1:
            return None
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return None


def broadcast_parameters(module, src=0, group=None):
    """Rank `src`'s parameters and buffers to every rank (what DDP does at construction)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
