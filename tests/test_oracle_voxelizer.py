"""CPU: the voxelizer oracles (C and numpy restatements) against golden vectors produced by the reference
numba implementation (oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import pytest

from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, random_points
from oracle import voxelizer as V

CASES = ["random", "boundary", "pile", "cap", "scene"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("impl", ["c", "np"])
def test_oracle_matches_reference_golden(golden_dir, case, impl):
    g = np.load(os.path.join(golden_dir, "voxel_%s.npz" % case))
    fn = V.points_to_voxel_c if impl == "c" else V.points_to_voxel_np
    r = fn(g["points"], NUSC_VOXEL, NUSC_RANGE, 10, int(g["max_voxels"]))
    assert np.array_equal(r["coors"], g["coors"])               # bit-exact indices and voxel order
    assert np.array_equal(r["num_points"], g["num_points"])
    np.testing.assert_allclose(r["mean"], g["mean"], rtol=1e-6, atol=1e-6)   # fp32 VFE mean


def test_golden_cases_cover_edge_conditions(golden_dir):
    g = np.load(os.path.join(golden_dir, "voxel_cap.npz"))
    assert len(g["coors"]) == int(g["max_voxels"])               # cap hit
    g = np.load(os.path.join(golden_dir, "voxel_pile.npz"))
    assert (g["num_points"] == 10).sum() > 10                    # max_points hit
    g = np.load(os.path.join(golden_dir, "voxel_random.npz"))
    assert len(g["coors"]) < len(g["points"]) * 0.8              # out-of-range rejection happened
    assert len(glob.glob(os.path.join(golden_dir, "voxel_*.npz"))) == len(CASES)


def test_c_and_numpy_oracles_agree_on_config1():
    """BASELINE config 1: 50k random points, CPU path, index bit-exactness."""
    pts = random_points(50000, seed=0, snap_frac=0.1, pile=2000)
    a = V.points_to_voxel_c(pts, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    b = V.points_to_voxel_np(pts, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    for k in ("coors", "num_points", "voxels", "mean"):
        assert np.array_equal(a[k], b[k]), k


def test_empty_and_all_rejected():
    empty = np.zeros((0, 5), np.float32)
    r = V.points_to_voxel_c(empty, NUSC_VOXEL, NUSC_RANGE, 10, 100)
    assert r["coors"].shape == (0, 3) and r["num_points"].shape == (0,)
    far = np.full((7, 5), 1000.0, np.float32)
    r = V.points_to_voxel_c(far, NUSC_VOXEL, NUSC_RANGE, 10, 100)
    assert r["coors"].shape == (0, 3)
    r = V.points_to_voxel_np(far, NUSC_VOXEL, NUSC_RANGE, 10, 100)
    assert r["coors"].shape == (0, 3)
