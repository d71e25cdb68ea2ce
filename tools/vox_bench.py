"""Voxelize+VFE family timing at several batch sizes (perf triage; run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scenes = [synth_scene(bench.N_TARGET, seed=5000 + i) for i in range(4)]
for nb in [int(a) for a in sys.argv[1:]] or [1, 4, 16, 32]:
    pts = torch.from_numpy(np.concatenate([scenes[i % 4] for i in range(nb)])).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(scenes[i % 4]) for i in range(nb)])], dtype=torch.int32, device=dev)
    for _ in range(3):
        vox = model.voxelize(pts, off)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        vox = model.voxelize(pts, off)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    m = int(vox["total"].item())
    nbytes = 20.0 * pts.shape[0] + 40.0 * m
    print("batch %2d: %.3f ms  %.1f GB/s algorithmic (%d pts, %d voxels) -> frac %.3f of 6530" %
          (nb, ms, nbytes / ms / 1e6, pts.shape[0], m, nbytes / ms / 1e6 / 6530.3), flush=True)
