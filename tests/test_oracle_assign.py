"""CPU: the target-assignment oracle reproduces the reference `AssignLabel.__call__` (tests/golden/assign.npz, written by
oracle/gen_golden.py from /root/reference): heat maps, indices, masks bit for bit, anno_box exactly."""
import os

import numpy as np

from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL
from oracle import assign_ref as AR


def test_assign_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "assign.npz"))
    for rm in (0, 1):
        for sample, seed in enumerate((0, 1)):
            boxes = AR.synth_annotations(seed)
            tag = "%d_%d" % (rm, sample)
            for t in range(3):
                hm, anno, ind, mask, cat = AR.assign_ref(boxes[t], np.ones(len(boxes[t]), np.int32), 1, (180, 180),
                                                         NUSC_RANGE[:2], NUSC_VOXEL[:2], 8, 0.1, 2, 500, bool(rm), t)
                assert np.array_equal(hm, g["hm_" + tag][t]), (tag, t)
                assert np.array_equal(ind, g["ind_" + tag][t]) and np.array_equal(mask, g["mask_" + tag][t])
                assert np.array_equal(cat, g["cat_" + tag][t])
                assert np.array_equal(anno, g["anno_box_" + tag][t])
            assert g["mask_" + tag][0].sum() < len(boxes[0])            # out-of-range and degenerate objects were dropped
    assert not np.array_equal(g["hm_0_0"], g["hm_1_0"])                # radius_mult changes the splats


def test_trajectory_sampler_oracle_matches_reference(golden_dir):
    """sampler_type = "trajectory" (the n3dtf / n3dtfm configs, preprocess.py:573-897): standard + `*_trajectory`
    (3 classes) + `*_forecast` (7 classes, boxes of all timesteps) targets."""
    g = np.load(os.path.join(golden_dir, "assign.npz"))
    boxes = AR.synth_annotations(2)
    traj = AR.synth_trajectories(2, len(boxes[0]))
    args = ((180, 180), NUSC_RANGE[:2], NUSC_VOXEL[:2], 8, 0.1, 2, 500)
    for rm in (0, 1):
        tag = "traj%d" % rm
        fb_, fc_ = AR.forecast_task(boxes)
        for t in range(3):
            cases = {"": (boxes[t], np.ones(len(boxes[t]), np.int32), 1), "_trajectory": (*AR.trajectory_task(boxes[t], traj), 3),
                     "_forecast": (fb_, fc_, 7)}
            for suffix, (b, c, ncls) in cases.items():
                hm, anno, ind, mask, cat = AR.assign_ref(b, c, ncls, *args, bool(rm), t)
                for key, val in (("hm", hm), ("anno_box", anno), ("ind", ind), ("mask", mask), ("cat", cat)):
                    assert np.array_equal(val, g[key + suffix + "_" + tag][t]), (key + suffix, rm, t)
    assert g["hm_forecast_traj0"].shape[1] == 7 and g["hm_trajectory_traj0"].shape[1] == 3
    assert len(set(g["cat_trajectory_traj0"][0].tolist())) == 3 and g["cat_forecast_traj0"][0].max() == 2
