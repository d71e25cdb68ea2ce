"""Profiling helper (run under ncu on the GPU box): a few forwards of the bench workload, nothing else."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

prec = os.environ.get("FD_PRECISION", "bf16x3")
dev = torch.device("cuda:0")
model = bench.build_model().set_precision(prec).to(dev).configure_voxelizer(bench.VOXEL_CFG)
nb = int(os.environ.get("FD_BATCH", "1"))
scenes = [synth_scene(bench.N_TARGET, seed=i) for i in range(nb)]
pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
off = torch.tensor(np.concatenate([[0], np.cumsum([len(s) for s in scenes])]), dtype=torch.int32, device=dev)
n = int(os.environ.get("FD_ITERS", "3"))
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        model.forward_points(pts, off)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
