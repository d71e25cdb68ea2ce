"""GPU: the widened path end to end on one batch -- raw sweeps -> fd_assemble_sweeps -> fused voxelizer -> train-mode
VoxelNet with targets from fd_assign_center_targets -> native backward -> optimizer step; then inference on the same
batch through CenterHead.predict.  No host synchronisation between the loader and the loss."""
import numpy as np
import torch

import pytest

import futuredet_b200 as fb
from futuredet_b200 import assign, loader, train
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene
from oracle import assign_ref as AR

pytestmark = pytest.mark.gpu


def raw_sweeps_from_scene(seed):
    """Split a synthetic 10-sweep cloud back into per-sweep nuScenes records (x,y,z,intensity,ring) + ego transforms."""
    pts = synth_scene(40000, seed=seed)
    key = pts[pts[:, 4] == 0.0]
    sweeps = []
    for s in range(1, 10):
        p = pts[np.isclose(pts[:, 4], 0.05 * s)]
        rec = np.zeros((len(p), 5), np.float32)
        rec[:, :4] = p[:, :4]
        T = np.eye(4)
        T[0, 3] = 0.01 * s
        sweeps.append((rec, T, 0.05 * s))
    krec = np.zeros((len(key), 5), np.float32)
    krec[:, :4] = key[:, :4]
    return krec, sweeps


def test_train_iteration_and_inference_from_raw_sweeps(cuda):
    from test_gpu_train import build_model
    model = build_model(3, cuda).to(cuda).train()
    vox_cfg = dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10, max_voxel_num=[120000, 160000])
    model.configure_voxelizer(vox_cfg, training=True)
    sb = loader.SweepBatch()
    annos = []
    for seed in (0, 1):
        sb.add_scene(*raw_sweeps_from_scene(seed))
        boxes = AR.synth_annotations(seed, n_obj=25, timesteps=3)
        annos.append(dict(gt_boxes=boxes, gt_classes=[np.ones(25, np.int32)] * 3))
    pts, boff, count = loader.assemble_sweeps(sb, cuda)
    example = assign.assign_targets(annos, [dict(num_class=1, class_names=["car"])],
                                    dict(out_size_factor=8, gaussian_overlap=0.1, max_objs=500, min_radius=2),
                                    [1440, 1440, 40], NUSC_RANGE, NUSC_VOXEL, cuda)
    tr = train.NativeTrainer(model, precision="bf16x3")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    first = None
    for _ in range(4):
        losses = tr.step(example, points=pts.contiguous(), batch_offsets=boff)
        opt.step()
        val = float(sum(losses["loss"]))
        assert np.isfinite(val)
        first = val if first is None else first
    assert val < first                                            # four AdamW steps on one batch reduce its loss
    # inference on the same batch
    model.eval()
    model.configure_voxelizer(vox_cfg, training=False)
    model.test_cfg = fb.ConfigDict(dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                                        nms=dict(nms_pre_max_size=1000, nms_post_max_size=83, nms_iou_threshold=0.2),
                                        score_threshold=0.1, pc_range=[-54, -54], out_size_factor=8,
                                        voxel_size=[0.075, 0.075]))
    with torch.no_grad():
        preds = model.forward_points(pts.contiguous(), boff)
        dets = model.bbox_head.predict({}, preds, model.test_cfg)
    assert len(dets) == 2 and all(d["box3d_lidar"].shape[1] == 9 for d in dets)
    assert all(len(d["scores"]) == len(d["label_preds"]) for d in dets)
