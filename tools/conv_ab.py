"""A/B of the tensor-core conv kernel's run-time knobs on the bench workload (perf triage; run on the GPU box).
usage: python tools/conv_ab.py [batch] -- prints per-class CUDA-event ms of one forward for every knob setting."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from futuredet_b200 import lib  # noqa: E402
from futuredet_b200.synth import synth_scene  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scenes = [synth_scene(bench.N_TARGET, seed=i) for i in range(nb)]
pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
L = lib.load()
L.fd_debug_set_tc.argtypes = [C.c_int, C.c_int]
CONFIGS = [("default", {})]
if len(sys.argv) > 2:
    CONFIGS = [(a, eval(a)) for a in sys.argv[2:]]
with torch.no_grad():
    for name, knobs in CONFIGS:
        L.fd_debug_set_tc(0, 0)
        L.fd_debug_set_tc(5, 0)
        L.fd_debug_set_tc(6, 1)
        for k, v in knobs.items():
            L.fd_debug_set_tc(k, v)
        for _ in range(2):
            model.forward_points(pts, off)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            model.forward_points(pts, off)
        e1.record()
        torch.cuda.synchronize()
        p = bench.conv_profile(model, pts, off)
        print("%-32s fwd %.2f ms | conv %.2f ms | %s" % (name, e0.elapsed_time(e1) / 3, p["ms"],
              "  ".join("%s %.2f" % (k.replace("sparse3d_", ""), v["ms"]) for k, v in p["by_kind"].items())), flush=True)
