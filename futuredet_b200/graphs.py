"""CUDA-graph replay of the launch-bound paths.

A forward pass is 81 native launches and a training step 764; at one scene per step the Python / ctypes / allocator time
around them (~40 us per launch) is comparable to the kernels themselves.  Both paths are synchronisation-free (row counts
stay on the device, every buffer has a static capacity), so they capture into a CUDA graph as they are: the ctypes calls
launch on torch's current stream, which is the capturing stream inside `torch.cuda.graph`.

  GraphedForward(model, max_points, batch_size)   replays VoxelNet.forward_points for a fixed (capacity, batch) shape
  GraphedTrainStep(trainer, ...)                   replays NativeTrainer.step (train-mode forward + backward)

Inputs are copied into static buffers; point rows past the live count are NaN, which the voxelizer rejects.
"""
import torch

from . import ops


def clear_weight_caches(model):
    """Drop every derived-weight cache (layout copies, folded BatchNorm, tensor-core packs) so that the next pass
    recomputes them -- inside a capture this records the recomputation into the graph."""
    for m in model.modules():
        for key in ("_kio_cache", "_kio_dmajor_cache", "_fold_cache", "_fused_cache", "_fused_last_cache", "_wpad_cache"):
            m.__dict__.pop(key, None)
    ops._PACK_CACHE.clear()


class _StaticPoints:
    def __init__(self, device, max_points, batch_size, point_dim=5):
        self.points = torch.full((max_points, point_dim), float("nan"), dtype=torch.float32, device=device)
        self.offsets = torch.zeros((batch_size + 1,), dtype=torch.int32, device=device)
        self._live = 0

    def load(self, points, offsets):
        n = points.shape[0]
        if n > self.points.shape[0] or offsets.numel() != self.offsets.numel():
            raise RuntimeError("graph was captured for at most %d points / batch %d" %
                               (self.points.shape[0], self.offsets.numel() - 1))
        self.points[:n].copy_(points, non_blocking=True)
        if n < self._live:
            self.points[n:self._live].fill_(float("nan"))
        self._live = n
        self.offsets.copy_(offsets, non_blocking=True)


class GraphedForward:
    """model.forward_points as one graph launch.  `model` must be in eval mode with a configured voxelizer."""

    def __init__(self, model, max_points, batch_size):
        self.model = model
        dev = next(model.parameters()).device
        self.inputs = _StaticPoints(dev, max_points, batch_size)
        self.graph = None
        self.preds = None

    def capture(self, points, offsets, warmup=3):
        self.inputs.load(points, offsets)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                      # fills the derived-weight caches and the allocator pools
                self.model.forward_points(self.inputs.points, self.inputs.offsets)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.preds = self.model.forward_points(self.inputs.points, self.inputs.offsets)
        return self

    def __call__(self, points, offsets):
        """Head tensors of the batch (views of static graph memory: consume or copy them before the next call)."""
        if self.graph is None:
            self.capture(points, offsets)                # records only: nothing has run yet
        else:
            self.inputs.load(points, offsets)
        self.graph.replay()
        return self.preds


class GraphedTrainStep:
    """NativeTrainer.forward + backward as one graph launch (single process; with several ranks use the eager step so
    that the NCCL bucket all-reduces can be interleaved).  Targets are copied into static buffers as well."""

    def __init__(self, trainer, max_points, batch_size):
        self.trainer = trainer
        dev = next(trainer.model.parameters()).device
        self.inputs = _StaticPoints(dev, max_points, batch_size)
        self.example = None
        self.graph = None
        self.losses = None

    def _load_example(self, example):
        if self.example is None:
            self.example = {k: [[t.clone() for t in row] for row in v] for k, v in example.items()}
            return
        for k, v in example.items():
            for row_s, row in zip(self.example[k], v):
                for dst, src in zip(row_s, row):
                    dst.copy_(src, non_blocking=True)

    def capture(self, example, points, offsets, warmup=3):
        self.inputs.load(points, offsets)
        self._load_example(example)
        tr = self.trainer
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                tr.step(self.example, points=self.inputs.points, batch_offsets=self.inputs.offsets)
        torch.cuda.current_stream().wait_stream(side)
        clear_weight_caches(tr.model)                    # layout copies / packs must be part of the graph: weights change
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = tr.step(self.example, points=self.inputs.points, batch_offsets=self.inputs.offsets)
        self._keepalive = tr._last_loss_ctx              # pinned pointer tables read by the graph's memcpy nodes
        return self

    def __call__(self, example, points, offsets):
        if self.graph is None:
            self.capture(example, points, offsets)       # records only; the warm-up steps ran without optimizer updates
        else:                                            # (BatchNorm running statistics advanced by those steps)
            self.inputs.load(points, offsets)
            self._load_example(example)
        self.graph.replay()
        return self.losses
