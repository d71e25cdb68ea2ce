"""GPU parity of the multi-sweep assembly against the reference loader's output (golden) and its
restatement, and of the sync-free chain loader -> voxelizer against the voxelizer oracle."""
import os

import numpy as np
import pytest
import torch

from futuredet_b200 import loader, ops
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL
from oracle import loader_ref as LR
from oracle import voxelizer as V

pytestmark = pytest.mark.gpu


def build_batch(seeds, orders):
    sb = loader.SweepBatch()
    for seed, order in zip(seeds, orders):
        key, sweeps = LR.synth_sweeps(seed)
        sb.add_scene(key, [sweeps[i] for i in order])
    return sb


def test_assembly_matches_reference_loader(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "loader.npz"))
    seeds = [int(g["seed_a"]), int(g["seed_b"])]
    orders = [g["order_a"], g["order_b"]]
    want = [g["combined_a"], g["combined_b"]]
    pts, boff, count = loader.assemble_sweeps(build_batch(seeds, orders), cuda)
    n = int(count.item())
    assert n == len(want[0]) + len(want[1])
    assert boff.cpu().tolist() == [0, len(want[0]), n]
    got = pts[:n].cpu().numpy()
    ref = np.concatenate(want, 0)
    # bit-exact: the kernel pins the float64 accumulation order of numpy's dgemm (sequential FMA over k, from zero)
    assert np.array_equal(got, ref)
    assert torch.isnan(pts[n:]).all().item() and pts.shape[0] > n                   # tail is NaN (capacity rows)


def test_single_scene_and_empty_sweep(cuda):
    key, sweeps = LR.synth_sweeps(5, n_sweeps=3, n_pts=500)
    sweeps[1] = (np.zeros((0, 5), np.float32), sweeps[1][1], sweeps[1][2])         # a sweep file with no points
    sb = loader.SweepBatch().add_scene(key, sweeps)
    pts, boff, count = loader.assemble_sweeps(sb, cuda)
    want = LR.assemble_ref(key, sweeps)
    n = int(count.item())
    assert n == len(want) and boff.cpu().tolist() == [0, n]
    assert np.array_equal(pts[:n].cpu().numpy(), want)


def test_loader_feeds_voxelizer_without_host_sync(cuda, golden_dir):
    """points (capacity rows, NaN tail) + device batch offsets go straight into fd_voxelize_vfe: same voxels as the
    oracle voxelizer on the reference loader's output."""
    g = np.load(os.path.join(golden_dir, "loader.npz"))
    seeds, orders = [int(g["seed_a"]), int(g["seed_b"])], [g["order_a"], g["order_b"]]
    pts, boff, count = loader.assemble_sweeps(build_batch(seeds, orders), cuda)
    vox = ops.voxelize_vfe(pts.contiguous(), boff, NUSC_VOXEL, NUSC_RANGE, 10, 20000)
    n = int(count.item())
    got_pts = pts[:n].cpu().numpy()
    o = V.voxelize_batch_c([got_pts[:int(boff[1])], got_pts[int(boff[1]):]], NUSC_VOXEL, NUSC_RANGE, 10, 20000)
    m = int(vox["total"].item())
    assert m == len(o["coords"])
    assert np.array_equal(vox["coords"][:m].cpu().numpy(), o["coords"])
    assert np.array_equal(vox["num_points"][:m].cpu().numpy(), o["num_points"])
    np.testing.assert_allclose(vox["features"][:m, :5].cpu().numpy(), o["features"], rtol=1e-6, atol=1e-6)
