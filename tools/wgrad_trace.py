"""clock64 timeline of block (0,0) of the output-stationary weight-gradient kernel (FD_WG_DBG |= 32; perf triage).
usage: FD_WG_DBG=32 python tools/wgrad_trace.py [C=64]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["FD_WG_DBG"] = str(int(os.environ.get("FD_WG_DBG", "0")) | 32)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200 import lib, ops, train_ops as T  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

Cw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lvl_of = {16: 0, 32: 1, 64: 2, 128: 3}[Cw]
dev = torch.device("cuda:0")
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scene = synth_scene(bench.N_TARGET, seed=3)
pts = torch.from_numpy(scene).to(dev)
off = torch.tensor([0, len(scene)], dtype=torch.int32, device=dev)
with torch.no_grad():
    _, vox = model.forward_points(pts, off, return_voxels=True)
coords, nd, cap = vox["coords"], vox["total"], int(vox["coords"].shape[0])
shape = [41, 1440, 1440]
for lvl in range(1, lvl_of + 1):
    rbs, _ = ops.rulebook_conv(coords, nd, cap, 1, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1] if lvl < 3 else [0, 1, 1])
    coords, nd, cap, shape = rbs.out_coords, rbs.n_out_dev, rbs.n_out_cap, rbs.out_shape
rb, _ = ops.rulebook_subm(coords, nd, cap, shape, [3, 3, 3], batch_size=1)
x = torch.randn((cap, Cw), device=dev)
gy = torch.randn((cap, Cw), device=dev)
dw = torch.zeros((27, Cw, Cw), device=dev)
for _ in range(3):
    T.sparse_conv_wgrad(x, gy, rb, dw, precision="bf16x3")
torch.cuda.synchronize()
L = lib.load()
tr = []
for role in range(4):
    buf = np.zeros(4096, np.int64)
    L.fd_debug_read_wgrad_trace(buf.ctypes.data_as(C.c_void_p), role)
    tr.append(buf)
n = int((tr[3] > 0).sum())
t0 = tr[0][0]
print("items traced:", n, " total cycles:", tr[3][n - 1] - t0, " per item:", (tr[3][n - 1] - t0) / max(n, 1))
print("item: prod_wait_begin prod_wait_end(+wait) mma_ready commit  [deltas vs previous item's commit]")
for i in list(range(0, 16)) + list(range(140, 160)):
    if i >= n:
        break
    print("%4d: P0 %8d  P1 %8d (wait %5d)  M2 %8d  M3 %8d (issue %5d)  d_commit %5d" % (
        i, tr[0][i] - t0, tr[1][i] - t0, tr[1][i] - tr[0][i], tr[2][i] - t0, tr[3][i] - t0, tr[3][i] - tr[2][i],
        tr[3][i] - tr[3][i - 1] if i else 0))
