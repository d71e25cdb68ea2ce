"""GPU parity: rulebook kernels vs the CPU restatement (bit-exact integer contract, canonical order)."""
import numpy as np
import pytest
import torch

from futuredet_b200 import ops
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene
from oracle import spconv_ref as S
from oracle import voxelizer as V

pytestmark = pytest.mark.gpu


def random_sites(rng, B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = rng.choice(cells, size=min(n, cells), replace=False)
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


def dev_coords(c, dev, extra=0):
    t = torch.zeros((len(c) + extra, 4), dtype=torch.int32, device=dev)
    t[:len(c)] = torch.from_numpy(c).to(dev)
    n = torch.tensor([len(c)], dtype=torch.int32, device=dev)
    return t, n


@pytest.mark.parametrize("n,extra", [(500, 0), (3000, 777), (1, 0), (0, 5)])
def test_subm_rulebook(cuda, n, extra):
    rng = np.random.default_rng(n)
    shape, B = [9, 20, 24], 2
    c = random_sites(rng, B, shape, n)
    ct, nd = dev_coords(c, cuda, extra)
    rb, _ = ops.rulebook_subm(ct, nd, len(c) + extra, shape, [3, 3, 3], batch_size=B)
    want = S.subm_rulebook(c, shape, [3, 3, 3])
    got = rb.nbr[:, :len(c)].cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(rb.pair_num.cpu().numpy(), (want >= 0).sum(1))
    if len(c):                                                   # per-128-row-tile activity masks
        tm = rb.tile_mask.cpu().numpy().astype(np.uint32)
        for t in range((len(c) + 127) // 128):
            bits = 0
            for k in range(27):
                if (want[k, t * 128:(t + 1) * 128] >= 0).any():
                    bits |= 1 << k
            assert int(tm[t]) == bits


@pytest.mark.parametrize("ksize,stride,pad", [([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [0, 1, 1]),
                                              ([3, 1, 1], [2, 1, 1], [0, 0, 0]), ([2, 2, 2], [2, 2, 2], [0, 0, 0])])
def test_strided_rulebook(cuda, ksize, stride, pad):
    rng = np.random.default_rng(3)
    shape, B = [11, 30, 26], 3
    c = random_sites(rng, B, shape, 4000)
    ct, nd = dev_coords(c, cuda, 100)
    rb, _ = ops.rulebook_conv(ct, nd, len(c) + 100, B, shape, ksize, stride, pad)
    oc, oshape, nbr = S.conv_rulebook(c, B, shape, ksize, stride, pad)
    n_out = int(rb.n_out_dev.item())
    assert rb.out_shape == oshape and n_out == len(oc)
    assert np.array_equal(rb.out_coords[:n_out].cpu().numpy(), oc)            # ascending linear order
    assert np.array_equal(rb.nbr[:, :n_out].cpu().numpy(), nbr)
    assert np.array_equal(rb.pair_num.cpu().numpy(), (nbr >= 0).sum(1))
    if rb.tile_mask is not None:                                 # per-128-row-tile activity masks
        tm = rb.tile_mask.cpu().numpy().astype(np.uint32)
        for t in range((n_out + 127) // 128):
            bits = 0
            for k in range(nbr.shape[0]):
                if (nbr[k, t * 128:(t + 1) * 128] >= 0).any():
                    bits |= 1 << k
            assert int(tm[t]) == bits
    # the input-stationary (scatter) build and the output-stationary search give the same table, bit for bit
    assert ops.SCATTER_STRIDED
    ops.SCATTER_STRIDED = False
    try:
        rb2, _ = ops.rulebook_conv(ct, nd, len(c) + 100, B, shape, ksize, stride, pad)
    finally:
        ops.SCATTER_STRIDED = True
    assert rb2.nbr.stride(0) != rb.nbr.stride(0) or True
    assert np.array_equal(rb2.nbr[:, :n_out].cpu().numpy(), rb.nbr[:, :n_out].cpu().numpy())
    assert np.array_equal(rb2.tile_mask.cpu().numpy(), rb.tile_mask.cpu().numpy())
    # spconv-layout export
    pairs, pair_num = rb.to_pairs()
    pairs = pairs.cpu().numpy()
    for k, (i, o) in enumerate(S.nbr_to_pairs(nbr)):
        assert np.array_equal(pairs[k, 0, :len(i)], i) and np.array_equal(pairs[k, 1, :len(o)], o)
        assert (pairs[k, :, len(i):] == -1).all()


def test_backbone_rulebook_chain_on_real_scene(cuda):
    """All 8 rulebooks of SpMiddleResNetFHD on a voxelized synthetic scene at the nuScenes grid."""
    pts = synth_scene(120000, seed=0)
    vox = V.points_to_voxel_c(pts, NUSC_VOXEL, NUSC_RANGE, 10, 160000, want_voxels=False)
    c = np.pad(vox["coors"], ((0, 0), (1, 0))).astype(np.int32)
    shape = [41, 1440, 1440]
    ct, nd = dev_coords(c, cuda, 1000)
    cap = len(c) + 1000
    geoms = [([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [0, 1, 1]),
             ([3, 1, 1], [2, 1, 1], [0, 0, 0])]
    cur_c, cur_t, cur_n, cur_cap, cur_shape = c, ct, nd, cap, shape
    cur_index = None                                   # level 0: hash index; levels >= 1: bitmap index of the out set
    for li, (k, s, p) in enumerate(geoms):
        rb, idx = ops.rulebook_subm(cur_t, cur_n, cur_cap, cur_shape, [3, 3, 3], index=cur_index, batch_size=1)
        assert (li == 0) == isinstance(idx, ops.CoordIndex)
        want = S.subm_rulebook(cur_c, cur_shape, [3, 3, 3])
        assert np.array_equal(rb.nbr[:, :len(cur_c)].cpu().numpy(), want), "subm level %d" % li
        rbc, _ = ops.rulebook_conv(cur_t, cur_n, cur_cap, 1, cur_shape, k, s, p, index=idx)
        oc, oshape, nbr = S.conv_rulebook(cur_c, 1, cur_shape, k, s, p)
        n_out = int(rbc.n_out_dev.item())
        assert n_out == len(oc) and rbc.out_shape == oshape
        assert np.array_equal(rbc.out_coords[:n_out].cpu().numpy(), oc), "out coords level %d" % li
        assert np.array_equal(rbc.nbr[:, :n_out].cpu().numpy(), nbr), "conv nbr level %d" % li
        cur_c, cur_t, cur_n, cur_cap, cur_shape = oc, rbc.out_coords, rbc.n_out_dev, rbc.n_out_cap, oshape
        cur_index = rbc.out_index
    assert cur_shape == [2, 180, 180]
