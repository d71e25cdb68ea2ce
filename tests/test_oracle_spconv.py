"""CPU: the spconv-1.x restatement (parity unpinned against real spconv -- see oracle/spconv_ref.py) is
checked against an independent definition: densify + torch conv3d (cross-correlation) sampled at the active set."""
import numpy as np
import pytest
import torch

from oracle import spconv_ref as S


def random_sites(rng, B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = rng.choice(cells, size=min(n, cells), replace=False)
    rng.shuffle(lin)
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


@pytest.mark.parametrize("seed", [0, 1])
def test_subm_matches_dense_conv3d(seed):
    rng = np.random.default_rng(seed)
    shape, B = [7, 10, 9], 2
    coords = random_sites(rng, B, shape, 300)
    x = torch.from_numpy(rng.standard_normal((len(coords), 6)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((3, 3, 3, 6, 5)).astype(np.float32))
    nbr = S.subm_rulebook(coords, shape, [3, 3, 3])
    y = S.indice_conv(x, w, nbr, len(coords))
    ref = S.dense_conv3d_reference(x, coords, B, shape, w, [3, 3, 3], [1, 1, 1], [1, 1, 1], coords)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    assert (nbr[13] == np.arange(len(coords))).all()          # centre offset maps every site to itself


@pytest.mark.parametrize("ksize,stride,pad", [([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [0, 1, 1]),
                                              ([3, 1, 1], [2, 1, 1], [0, 0, 0])])
def test_strided_matches_dense_conv3d(ksize, stride, pad):
    rng = np.random.default_rng(7)
    shape, B = [11, 12, 10], 2
    coords = random_sites(rng, B, shape, 250)
    x = torch.from_numpy(rng.standard_normal((len(coords), 4)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((*ksize, 4, 8)).astype(np.float32))
    oc, oshape, nbr = S.conv_rulebook(coords, B, shape, ksize, stride, pad)
    assert oshape == [(s + 2 * p - k) // st + 1 for s, k, st, p in zip(shape, ksize, stride, pad)]
    key = ((oc[:, 0].astype(np.int64) * oshape[0] + oc[:, 1]) * oshape[1] + oc[:, 2]) * oshape[2] + oc[:, 3]
    assert (np.diff(key) > 0).all()                           # canonical ascending order, unique
    y = S.indice_conv(x, w, nbr, len(oc))
    ref = S.dense_conv3d_reference(x, coords, B, shape, w, ksize, stride, pad, oc)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    # the active set is exactly the support of the dense result for positive inputs/weights
    dense = S.dense_conv3d_reference(torch.ones(len(coords), 1), coords, B, shape, torch.ones(*ksize, 1, 1), ksize,
                                     stride, pad, oc)
    assert (dense > 0).all()
    full = torch.zeros((B, 1, *shape), dtype=torch.float64)
    ci = torch.from_numpy(coords.astype(np.int64))
    full[ci[:, 0], 0, ci[:, 1], ci[:, 2], ci[:, 3]] = 1
    cnt = torch.nn.functional.conv3d(full, torch.ones(1, 1, *ksize, dtype=torch.float64), stride=stride, padding=pad)
    assert int((cnt > 0).sum()) == len(oc)


def test_backbone_oracle_shapes():
    from futuredet_b200.backbone import SpMiddleResNetFHD
    torch.manual_seed(0)
    m = SpMiddleResNetFHD(num_input_features=5).eval()
    rng = np.random.default_rng(0)
    grid = [32, 32, 40]                       # x, y, z  -> sparse shape [41, 32, 32] -> 21 -> 11 -> 5 -> 2
    coords = random_sites(rng, 2, [40, 32, 32], 600)
    x = torch.from_numpy(rng.standard_normal((len(coords), 5)).astype(np.float32))
    dense, st = S.backbone_forward(m.state_dict(), x, coords, 2, grid, return_stages=True)
    assert dense.shape == (2, 256, 4, 4)
    assert st["conv2"][2] == [21, 16, 16] and st["conv3"][2] == [11, 8, 8]
    assert st["conv4"][2] == [5, 4, 4] and st["extra_conv"][2] == [2, 4, 4]
    # dense().view(N, C*D, H, W): channel index = c*D + d  (scn.py:165-168)
    x4, c4, _ = st["extra_conv"]
    r = 0
    b, d, yy, xx = c4[r]
    torch.testing.assert_close(dense[b, d::2, yy, xx], x4[r])
