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
