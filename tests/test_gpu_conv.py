"""GPU parity: gather->implicit-GEMM convolution (fd_conv_forward) vs CPU references.
fp32 arm: 1e-4 (accumulation order only).  Tensor-core arm (bf16x3): 1e-3 abs as BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from futuredet_b200 import ops
from oracle import spconv_ref as S

pytestmark = pytest.mark.gpu
PRECS = [("fp32", 1e-4), ("bf16x3", 3e-4), ("bf16", 6e-2)]


def random_sites(rng, B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = rng.choice(cells, size=min(n, cells), replace=False)
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


@pytest.mark.parametrize("prec,tol", PRECS)
@pytest.mark.parametrize("cin,cout", [(5, 16), (16, 16), (16, 32), (32, 64), (64, 64), (128, 128)])
def test_subm_conv_fused_epilogue(cuda, prec, tol, cin, cout):
    rng = np.random.default_rng(cin * 1000 + cout)
    shape, B = [9, 24, 24], 2
    c = random_sites(rng, B, shape, 3000)
    n = len(c)
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32))
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    res = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32))
    nbr = S.subm_rulebook(c, shape, [3, 3, 3])
    want = F.relu(S.indice_conv(x, w, nbr, n) * scale + shift + res)
    cap = n + 300
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    xg = torch.zeros((cap, cin), device=cuda); xg[:n] = x.to(cuda)
    rg = torch.zeros((cap, cout), device=cuda); rg[:n] = res.to(cuda)
    if prec != "fp32" and not ops.tc_supported(cin, 27):
        with pytest.raises(RuntimeError, match="tensor-core arm needs Cin"):      # no silent fallback in the ABI
            ops.sparse_conv(xg, w.to(cuda), rb, precision=prec)
        return
    y = ops.sparse_conv(xg, w.to(cuda), rb, scale.to(cuda), shift.to(cuda), rg, True, precision=prec)
    torch.testing.assert_close(y[:n].cpu(), want, rtol=tol, atol=tol)
    # plain conv (no epilogue), rows beyond n untouched by contract
    y2 = ops.sparse_conv(xg, w.to(cuda), rb, precision=prec)
    torch.testing.assert_close(y2[:n].cpu(), S.indice_conv(x, w, nbr, n), rtol=tol, atol=tol)


@pytest.mark.parametrize("prec,tol", PRECS)
def test_strided_conv_and_bev_epilogue(cuda, prec, tol):
    rng = np.random.default_rng(11)
    shape, B = [5, 16, 16], 2
    c = random_sites(rng, B, shape, 1500)
    n = len(c)
    cin, cout = 32, 32
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((3, cin, cout)) / np.sqrt(3 * cin)).astype(np.float32))
    oc, oshape, nbr = S.conv_rulebook(c, B, shape, [3, 1, 1], [2, 1, 1], [0, 0, 0])
    feat = F.relu(S.indice_conv(x, w, nbr, len(oc)))
    dense = torch.zeros((B, cout, *oshape))
    ci = torch.from_numpy(oc.astype(np.int64))
    dense[ci[:, 0], :, ci[:, 1], ci[:, 2], ci[:, 3]] = feat
    want = dense.view(B, cout * oshape[0], oshape[1], oshape[2])                       # scn.py:165-168
    ct = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    rb, _ = ops.rulebook_conv(ct, nd, n, B, shape, [3, 1, 1], [2, 1, 1], [0, 0, 0])
    bev = ops.sparse_conv(x.to(cuda), w.to(cuda), rb, relu=True, precision=prec, bev=(B, *oshape))
    torch.testing.assert_close(bev.permute(0, 3, 1, 2).cpu(), want, rtol=tol, atol=tol)
    # SparseConvTensor.dense() kernel agrees too
    y = ops.sparse_conv(x.to(cuda), w.to(cuda), rb, relu=True, precision=prec)
    d2 = ops.sparse_to_dense(y, rb.out_coords, rb.n_out_dev, rb.n_out_cap, B, oshape)
    torch.testing.assert_close(d2.cpu(), dense, rtol=tol, atol=tol)


@pytest.mark.parametrize("prec,tol", PRECS)
@pytest.mark.parametrize("cin,cout,k,s,p,hw", [(32, 64, 3, 1, 1, 20), (64, 32, 3, 2, 1, 21), (48, 64, 1, 1, 0, 9),
                                              (256, 128, 3, 1, 1, 12), (64, 3, 3, 1, 1, 10), (16, 16, 2, 2, 0, 12)])
def test_conv2d_nhwc(cuda, prec, tol, cin, cout, k, s, p, hw):
    g = torch.Generator().manual_seed(cin + cout)
    B = 2
    x = torch.randn((B, cin, hw, hw + 3), generator=g)
    w = torch.randn((cout, cin, k, k), generator=g) / np.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g)
    want = F.relu(F.conv2d(x, w, None, stride=s, padding=p) * scale[None, :, None, None] + shift[None, :, None, None])
    wk = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout).contiguous().to(cuda)
    xg = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    if prec != "fp32" and not ops.tc_supported(cin, k * k):
        with pytest.raises(RuntimeError, match="tensor-core arm needs Cin"):
            ops.conv2d_nhwc(xg, wk, (k, k), (s, s), (p, p), precision=prec)
        return
    y = ops.conv2d_nhwc(xg, wk, (k, k), (s, s), (p, p), scale.to(cuda), shift.to(cuda), True, precision=prec)
    torch.testing.assert_close(y.permute(0, 3, 1, 2).cpu(), want, rtol=tol, atol=tol)


@pytest.mark.parametrize("prec,tol", PRECS)
def test_channel_slices_and_conv_transpose(cuda, prec, tol):
    """Reading a channel slice / writing into a slice of a wider buffer (fused torch.cat), ConvTranspose2d(2, s2)."""
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 9, 11
    wide = torch.randn((B, H, W, 96), generator=g)
    x = wide[..., 32:64]
    wt = torch.randn((32, 48, 2, 2), generator=g) / 8                                   # ConvTranspose2d layout [Cin,Cout,kh,kw]
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, None, stride=2)
    out = torch.full((B, 2 * H, 2 * W, 80), -7.0, device=cuda)
    wk = wt.permute(2, 3, 0, 1).reshape(4, 32, 48).contiguous().to(cuda)
    ops.conv2d_nhwc(wide.to(cuda)[..., 32:64], wk, (2, 2), (2, 2), (0, 0), out=out[..., 16:64], precision=prec,
                    transposed=True)
    torch.testing.assert_close(out[..., 16:64].permute(0, 3, 1, 2).cpu(), want, rtol=tol, atol=tol)
    assert (out[..., :16] == -7).all() and (out[..., 64:] == -7).all()                  # neighbours untouched


def test_split_row_format_round_trip(cuda):
    """FD_FMT_SPLIT_BF16 rows: hi + lo reproduces fp32 to ~2^-17 relative; channel slices address both planes."""
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((1000, 48), generator=g) * 10 ** torch.randint(-3, 4, (1000, 48), generator=g).float()).to(cuda)
    s = ops.to_split(x)
    back = s.to_fp32()
    rel = ((back - x).abs() / x.abs().clamp_min(1e-30)).max().item()
    assert rel < 2 ** -16
    torch.testing.assert_close(s.slice(16, 8).to_fp32(), back[:, 16:24], rtol=0, atol=0)
    assert ops.to_split(torch.zeros((4, 8), device=cuda)).to_fp32().abs().max().item() == 0.0


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (64, 64), (64, 128), (128, 128)])
def test_subm_conv_split_format_pipeline(cuda, cin, cout):
    """Tensor-core arm with split bf16 hi/lo rows for input, residual and output (the inter-layer format),
    chained twice, against the fp32 oracle."""
    rng = np.random.default_rng(cin + cout)
    shape, B = [9, 24, 24], 2
    c = random_sites(rng, B, shape, 3500)
    n = len(c)
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w1 = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32))
    w2 = torch.from_numpy((rng.standard_normal((27, cout, cout)) / np.sqrt(27 * cout)).astype(np.float32))
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    nbr = S.subm_rulebook(c, shape, [3, 3, 3])
    y1 = F.relu(S.indice_conv(x, w1, nbr, n) * scale + shift)
    want = F.relu(S.indice_conv(y1, w2, nbr, n) * scale + shift + y1)
    cap = n + 77
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    xg = torch.zeros((cap, cin), device=cuda); xg[:n] = x.to(cuda)
    xs = ops.to_split(xg)
    g1 = ops.sparse_conv(xs, w1.to(cuda), rb, scale.to(cuda), shift.to(cuda), None, True, precision="bf16x3", out_fmt="split")
    assert isinstance(g1, ops.Feat) and g1.fmt == "split"
    torch.testing.assert_close(g1.to_fp32()[:n].cpu(), y1, rtol=3e-4, atol=3e-4)
    g2 = ops.sparse_conv(g1, w2.to(cuda), rb, scale.to(cuda), shift.to(cuda), g1, True, precision="bf16x3", out_fmt="fp32")
    torch.testing.assert_close(g2[:n].cpu(), want, rtol=5e-4, atol=5e-4)
    # the fp32 CUDA-core arm reads the same split rows (mixed pipelines stay well-defined)
    g3 = ops.sparse_conv(g1, w2.to(cuda), rb, scale.to(cuda), shift.to(cuda), g1, True, precision="fp32", out_fmt="split")
    torch.testing.assert_close(g3.to_fp32()[:n].cpu(), want, rtol=5e-4, atol=5e-4)


def test_dense_conv_split_format_and_slices(cuda):
    g = torch.Generator().manual_seed(3)
    B, H, W, cin, cout = 2, 14, 13, 64, 32
    x = torch.randn((B, cin, H, W), generator=g)
    w = torch.randn((cout, cin, 3, 3), generator=g) / np.sqrt(cin * 9)
    want = F.relu(F.conv2d(x, w, None, padding=1))
    wk = w.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous().to(cuda)
    xs = ops.to_split(x.permute(0, 2, 3, 1).contiguous().to(cuda))
    wide = ops.Feat(torch.zeros((B, H, W, 96), device=cuda), "split")
    ops.conv2d_nhwc(xs, wk, (3, 3), (1, 1), (1, 1), relu=True, out=wide.slice(32, cout), precision="bf16x3")
    torch.testing.assert_close(wide.slice(32, cout).to_fp32().permute(0, 3, 1, 2).cpu(), want, rtol=3e-4, atol=3e-4)
    assert wide.slice(0, 32).to_fp32().abs().max().item() == 0 and wide.slice(64, 32).to_fp32().abs().max().item() == 0


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (64, 64), (128, 128)])
def test_sorted_tiles_bit_identical(cuda, cin, cout, monkeypatch):
    """fd_rulebook_sort_rows: the sorted table is the source table with its rows permuted inside the windows, its masks
    are those of the sorted tiles, and a convolution over it (row_perm epilogue) is bit-identical to the unsorted call
    on both arms, residual + ReLU included; the sparse, clustered sites make most tiles skip most kernel offsets."""
    monkeypatch.setattr(ops, "SORT_WINDOW", 2048)
    monkeypatch.setattr(ops, "SORT_MIN_ROWS", 0)
    rng = np.random.default_rng(cin + cout)
    shape, B = [12, 96, 96], 2
    c = random_sites(rng, B, shape, 9000)                 # ~4 % occupancy: few neighbours per row
    n = len(c)
    cap = n + 517
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    assert rb.row_key is not None                         # keys handed over by the neighbour search
    perm, nbr_s, tmask_s = rb.sorted_tiles()
    keys = rb.row_key[:n].cpu().numpy().astype(np.int64)
    rb2, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    rb2.row_key = None                                    # ... or computed from the table: same buckets
    perm2 = rb2.sorted_tiles()[0][:n].cpu().numpy()
    p = perm[:n].cpu().numpy()
    assert np.array_equal(keys[p], keys[perm2])
    assert all(np.all(np.diff(keys[p][w:w + 2048]) >= 0) for w in range(0, n, 2048))   # ascending keys inside a window
    assert np.array_equal(np.sort(p), np.arange(n))                                     # a permutation of the live rows
    win = 2048
    assert np.array_equal(p // win, np.arange(n) // win)                                # ... inside the windows
    nbr, ns = rb.nbr.cpu().numpy()[:, :n], nbr_s.cpu().numpy()[:, :n]
    assert np.array_equal(ns, nbr[:, p])
    bits = ((ns >= 0).astype(np.uint32) << np.arange(27, dtype=np.uint32)[:, None]).sum(0).astype(np.uint32)
    want_mask = np.array([np.bitwise_or.reduce(bits[t:t + 128]) for t in range(0, n, 128)], np.uint32)
    assert np.array_equal(tmask_s.cpu().numpy().view(np.uint32)[:len(want_mask)], want_mask)
    unsorted_mask = rb.tile_mask.cpu().numpy().view(np.uint32)[:len(want_mask)]
    popc = lambda m: sum(int(v).bit_count() for v in m)
    assert popc(want_mask) < 0.8 * popc(unsorted_mask)                                  # the point of sorting
    x = torch.zeros((cap, cin), device=cuda); x[:n] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)).to(cuda)
    res = torch.zeros((cap, cout), device=cuda); res[:n] = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32)).to(cuda)
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)).to(cuda)
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)).to(cuda)
    shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).to(cuda)
    for prec in ("fp32", "bf16x3"):
        for fmt in (("fp32",) if prec == "fp32" else ("fp32", "split")):
            xin, rin = (ops.to_split(x, nd), ops.to_split(res, nd)) if fmt == "split" else (x, res)
            ya = ops.sparse_conv(xin, w, rb, scale, shift, rin, True, precision=prec, out_fmt=fmt, sort_tiles=False)
            yb = ops.sparse_conv(xin, w, rb, scale, shift, rin, True, precision=prec, out_fmt=fmt, sort_tiles=True)
            ya, yb = (ya.t, yb.t) if fmt == "split" else (ya, yb)
            assert torch.equal(ya[:n].view(torch.int32), yb[:n].view(torch.int32)), (prec, fmt)


@pytest.mark.parametrize("cin,cout,H,W", [(64, 64, 37, 29), (128, 128, 16, 8), (256, 256, 45, 45), (128, 11, 33, 40),
                                          (512, 64, 20, 19), (64, 384, 17, 9)])
def test_dense_3x3_tall_stages(cuda, cin, cout, H, W):
    """Dense 3x3 stride-1 convolutions on split-bf16 rows take the tall-stage TMA path (8 x 16-pixel tiles, one
    {64 ch, 8 px, 18 lines} box per horizontal tap feeding the three vertical taps): maps that are not multiples of the
    tile, image borders (zero padding = out-of-bounds fill), several channel chunks, N tiles and tile widths, residual."""
    g = torch.Generator().manual_seed(cin + cout + H)
    B = 3
    x = torch.randn((B, cin, H, W), generator=g)
    w = torch.randn((cout, cin, 3, 3), generator=g) / np.sqrt(cin * 9)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g)
    res = torch.randn((B, cout, H, W), generator=g)
    want = F.relu(F.conv2d(x, w, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None] + res)
    wk = w.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous().to(cuda)
    xs = ops.to_split(x.permute(0, 2, 3, 1).contiguous().to(cuda))
    rs = res.permute(0, 2, 3, 1).contiguous().to(cuda)
    y = ops.conv2d_nhwc(xs, wk, (3, 3), (1, 1), (1, 1), scale.to(cuda), shift.to(cuda), True, residual=rs,
                        precision="bf16x3", out_fmt="split")
    torch.testing.assert_close(y.to_fp32().permute(0, 3, 1, 2).cpu(), want, rtol=3e-4, atol=3e-4)
