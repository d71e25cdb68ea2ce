"""Weight-gradient kernel timing on the bench workload's sparse levels (perf triage; run on the GPU box).
usage: python tools/wgrad_bench.py -- the four SubM rulebooks of one 305k-point scene (levels chained through real
strided rulebooks, channels 16 / 32 / 64 / 128); 5 fd_conv_wgrad_det calls are captured in a CUDA graph (the Python
call costs more than the kernels) and the replay is timed with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200 import ops, train_ops as T  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scene = synth_scene(bench.N_TARGET, seed=3)
pts = torch.from_numpy(scene).to(dev)
off = torch.tensor([0, len(scene)], dtype=torch.int32, device=dev)
with torch.no_grad():
    _, vox = model.forward_points(pts, off, return_voxels=True)
coords, nd, cap = vox["coords"], vox["total"], int(vox["coords"].shape[0])
shape = [41, 1440, 1440]
REP = 5
ONLY = int(os.environ.get("FD_WG_LEVEL", "0"))      # run one channel width only (ncu captures)
for lvl, C in enumerate((16, 32, 64, 128)):
    if lvl > 0:
        rbs, _ = ops.rulebook_conv(coords, nd, cap, 1, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1] if lvl < 3 else [0, 1, 1])
        coords, nd, cap, shape = rbs.out_coords, rbs.n_out_dev, rbs.n_out_cap, rbs.out_shape
    if ONLY and C != ONLY:
        continue
    rb, _ = ops.rulebook_subm(coords, nd, cap, shape, [3, 3, 3], batch_size=1)
    m = int(nd.item())
    x = torch.randn((cap, C), device=dev)
    gy = torch.randn((cap, C), device=dev)
    dw = torch.zeros((27, C, C), device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            T.sparse_conv_wgrad(x, gy, rb, dw, precision="bf16x3")
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                T.sparse_conv_wgrad(x, gy, rb, dw, precision="bf16x3")
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    pairs = int((rb.nbr[:, :m] >= 0).sum().item())
    us = e0.elapsed_time(e1) / REP * 1e3
    print("C%-3d rows %7d (cap %7d) pairs/row %5.2f : %7.1f us  (%.1f TFLOP/s algorithmic)" % (C, m, cap, pairs / m, us, 2.0 * pairs * C * C / us / 1e6), flush=True)
