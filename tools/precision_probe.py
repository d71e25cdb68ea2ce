"""Max |error| of the head tensors per arithmetic arm against the exact-fp32 arm, bench model, full scenes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

dev = torch.device("cuda:0")
for seed in (0, 1):
    model = bench.build_model(seed).to(dev).configure_voxelizer(bench.VOXEL_CFG)
    g = torch.Generator().manual_seed(100 + seed)
    for m in model.modules():                               # randomised BN affine too (as the parity tests do)
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data.copy_((torch.rand(m.weight.shape, generator=g) + 0.5).to(dev))
            m.bias.data.copy_((torch.randn(m.bias.shape, generator=g) * 0.1).to(dev))
    scenes = [synth_scene(bench.N_TARGET, seed=10 * seed + i) for i in range(2)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
    outs = {}
    with torch.no_grad():
        for prec in ("fp32", "bf16x3", "bf16"):
            model.set_precision(prec)
            outs[prec] = {k: v.clone() for k, v in model.forward_points(pts, off)[0].items()}
    for prec in ("bf16x3", "bf16"):
        errs = {k: float((outs[prec][k] - outs["fp32"][k]).abs().max()) for k in outs["fp32"]}
        mags = {k: float(outs["fp32"][k].abs().max()) for k in outs["fp32"]}
        print("seed %d %-7s max|err| %s   (max |value| %s)" % (seed, prec, {k: "%.2e" % v for k, v in errs.items()},
                                                            {k: "%.1f" % v for k, v in mags.items()}), flush=True)
