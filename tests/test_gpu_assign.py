"""GPU parity of the batched target assignment against the reference AssignLabel output (golden) and
its numpy restatement: heat map, ind, mask, cat bit-exact; anno_box to float32 transcendental precision."""
import os

import numpy as np
import pytest
import torch

from futuredet_b200 import assign
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL
from oracle import assign_ref as AR

pytestmark = pytest.mark.gpu
CFG = dict(out_size_factor=8, gaussian_overlap=0.1, max_objs=500, min_radius=2, radius_mult=False)


@pytest.mark.parametrize("radius_mult", [False, True])
def test_assign_matches_reference_golden(cuda, golden_dir, radius_mult):
    g = np.load(os.path.join(golden_dir, "assign.npz"))
    annos = []
    for seed in (0, 1):
        boxes = AR.synth_annotations(seed)
        annos.append(dict(gt_boxes=boxes, gt_classes=[np.ones(len(b), np.int32) for b in boxes]))
    ex = assign.assign_targets(annos, [dict(num_class=1, class_names=["car"])], dict(CFG, radius_mult=radius_mult),
                               [1440, 1440, 40], NUSC_RANGE, NUSC_VOXEL, cuda)
    rm = int(radius_mult)
    for t in range(3):
        for key in ("hm", "ind", "mask", "cat"):
            got = ex[key][t][0].cpu().numpy()
            want = np.stack([g["%s_%d_%d" % (key, rm, s)][t] for s in (0, 1)])
            assert got.dtype == want.dtype and np.array_equal(got, want), (key, t)
        got = ex["anno_box"][t][0].cpu().numpy()
        want = np.stack([g["anno_box_%d_%d" % (rm, s)][t] for s in (0, 1)])
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-7)
        assert np.array_equal(got[..., [0, 1, 2, 6, 7, 8, 9]], want[..., [0, 1, 2, 6, 7, 8, 9]])   # no transcendental: exact


@pytest.mark.parametrize("radius_mult", [False, True])
def test_trajectory_sampler_matches_reference_golden(cuda, golden_dir, radius_mult):
    """sampler_type "trajectory" (n3dtf / n3dtfm configs): standard + 3-class trajectory + 7-class forecast targets."""
    g = np.load(os.path.join(golden_dir, "assign.npz"))
    boxes = AR.synth_annotations(2)
    traj = AR.synth_trajectories(2, len(boxes[0]))
    anno = dict(gt_boxes=boxes, gt_classes=[np.ones(len(b), np.int32) for b in boxes], gt_trajectory=[traj] * 3)
    ex = assign.assign_targets([anno], [dict(num_class=1, class_names=["car"])],
                               dict(CFG, radius_mult=radius_mult, sampler_type="trajectory"), [1440, 1440, 40], NUSC_RANGE,
                               NUSC_VOXEL, cuda)
    tag = "traj%d" % int(radius_mult)
    for suffix in ("", "_trajectory", "_forecast"):
        for t in range(3):
            for key in ("hm", "ind", "mask", "cat"):
                got = ex[key + suffix][t][0][0].cpu().numpy()
                want = g[key + suffix + "_" + tag][t]
                assert got.dtype == want.dtype and np.array_equal(got, want), (key + suffix, t)
            np.testing.assert_allclose(ex["anno_box" + suffix][t][0][0].cpu().numpy(), g["anno_box" + suffix + "_" + tag][t],
                                       rtol=2e-6, atol=2e-7)
    assert ex["hm_forecast"][0][0].shape[1] == 7 and ex["hm_trajectory"][0][0].shape[1] == 3


def test_assign_two_tasks_many_objects(cuda):
    """car + pedestrian tasks (BASELINE configs[4] shape): class grouping and per-task local class ids, max_objs cap."""
    rng = np.random.default_rng(3)
    boxes = AR.synth_annotations(7, n_obj=700, timesteps=1)[0]
    boxes[350:, 3:5] = np.abs(rng.normal([0.7, 0.7], 0.1, (350, 2))).astype(np.float32)
    classes = np.where(np.arange(700) < 350, 1, 2).astype(np.int32)
    perm = rng.permutation(700)
    boxes, classes = boxes[perm], classes[perm]
    tasks = [dict(num_class=1, class_names=["car"]), dict(num_class=1, class_names=["pedestrian"])]
    ex = assign.assign_targets([dict(gt_boxes=[boxes], gt_classes=[classes])], tasks, dict(CFG, max_objs=300),
                               [1440, 1440, 40], NUSC_RANGE, NUSC_VOXEL, cuda)
    for task_id in (0, 1):
        sel = np.where(classes == task_id + 1)[0]
        hm, anno, ind, mask, cat = AR.assign_ref(boxes[sel], np.ones(len(sel), np.int32), 1, (180, 180), NUSC_RANGE[:2],
                                                 NUSC_VOXEL[:2], 8, 0.1, 2, 300)
        assert np.array_equal(ex["hm"][0][task_id][0].cpu().numpy(), hm)
        assert np.array_equal(ex["ind"][0][task_id][0].cpu().numpy(), ind)
        assert np.array_equal(ex["mask"][0][task_id][0].cpu().numpy(), mask) and mask.sum() > 200
        np.testing.assert_allclose(ex["anno_box"][0][task_id][0].cpu().numpy(), anno, rtol=2e-6, atol=2e-7)


def test_assigned_targets_feed_the_loss(cuda):
    """The produced dict is consumed as-is by CenterHead.loss (native) and matches the loss on the oracle's targets."""
    from futuredet_b200.loss import center_head_loss
    from oracle.loss_ref import center_head_loss_ref
    boxes = AR.synth_annotations(11, n_obj=30, timesteps=3)
    ex = assign.assign_targets([dict(gt_boxes=boxes, gt_classes=[np.ones(30, np.int32)] * 3)],
                               [dict(num_class=1, class_names=["car"])], CFG, [1440, 1440, 40], NUSC_RANGE, NUSC_VOXEL, cuda)
    g = torch.Generator().manual_seed(0)
    chans = dict(reg=2, height=1, dim=3, rot=2, vel=6, hm=1)
    preds = {k: torch.randn((1, c, 180, 180), generator=g) for k, c in chans.items()}
    head = type("H", (), dict(timesteps=3, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0], weight=0.25))()
    head.code_weights_forecast = list(np.array(head.code_weights) * np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 0]))
    buf = torch.cat([preds[k].permute(0, 2, 3, 1) for k in chans], -1).contiguous().to(cuda)
    views, col = {}, 0
    for k, c in chans.items():
        views[k] = buf[..., col:col + c].permute(0, 3, 1, 2)
        col += c
    got = center_head_loss(head, ex, [views])
    ex_cpu = {k: [[x.cpu() for x in row] for row in v] for k, v in ex.items()}
    want = center_head_loss_ref([preds], ex_cpu, 3, head.code_weights, head.weight)
    torch.testing.assert_close(got["loss"][0].cpu(), want["loss"][0], rtol=1e-4, atol=1e-5)
