"""Timing of the strided-rulebook builders (output-side search vs input-side scatter) on the bench scenes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200 import ops  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scenes = [synth_scene(bench.N_TARGET, seed=i) for i in range(nb)]
pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
vox = model.voxelize(pts, off)
coords, n_dev, cap = vox["coords"], vox["total"], vox["coords"].shape[0]
shape = [41, 1440, 1440]
levels = [([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [0, 1, 1]),
          ([3, 1, 1], [2, 1, 1], [0, 0, 0])]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


c, n, ncap, shp = coords, n_dev, cap, shape
for li, (k, s, p) in enumerate(levels):
    res = {}
    for mode in (False, True):
        ops.SCATTER_STRIDED = mode
        ms, (rb, _) = timed(lambda: ops.rulebook_conv(c, n, ncap, nb, shp, k, s, p))
        res[mode] = ms
    ms_sub, _ = timed(lambda: ops.rulebook_subm(rb.out_coords, rb.n_out_dev, rb.n_out_cap, rb.out_shape, [3, 3, 3], index=rb.out_index))
    print("level %d: in rows %d -> out rows %d (cap %d): search %.3f ms, scatter %.3f ms | subm(bitmap) %.3f ms" %
          (li + 1, int(n.item()), int(rb.n_out_dev.item()), rb.n_out_cap, res[False], res[True], ms_sub), flush=True)
    c, n, ncap, shp = rb.out_coords, rb.n_out_dev, rb.n_out_cap, rb.out_shape
