"""Two training steps of the forecast_n3 model on one 305k-point scene (target of the ncu launch-list capture)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import futuredet_b200 as fb
from futuredet_b200 import train
from futuredet_b200.synth import synth_scene, synth_targets
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = fb.build_detector(bench.model_cfg(timesteps=7)).to(dev).train()
m.configure_voxelizer(bench.VOXEL_CFG, training=True)
scene = synth_scene(bench.N_TARGET, seed=7)
pts = torch.from_numpy(scene).to(dev); off = torch.tensor([0, len(scene)], dtype=torch.int32, device=dev)
ex = synth_targets(1, 180, 180, 7, seed=0); ex = {k: [[t.to(dev) for t in ts] for ts in v] for k, v in ex.items()}
tr = train.NativeTrainer(m, precision=os.environ.get("FD_TRAIN_PRECISION", "bf16x3"))
for _ in range(int(os.environ.get("FD_TRAIN_STEPS", "2"))):
    tr.step(ex, points=pts, batch_offsets=off)
torch.cuda.synchronize()
