"""GPU parity of the training path (BASELINE configs[2]: forecast_n3, fwd+bwd single sample): every backward kernel
against torch autograd on the CPU oracle, RPN + CenterHead gradients against the golden fixture written from the
REFERENCE classes in training mode, and the whole VoxelNet train step (loss, every parameter gradient, BatchNorm
running statistics) against autograd over the chained oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn

import futuredet_b200 as fb
from futuredet_b200 import ops, train
from futuredet_b200 import train_ops as T
from futuredet_b200.loss import center_head_loss, center_head_loss_backward
from oracle import dense_ref as D
from oracle import spconv_ref as S
from oracle.loss_ref import center_head_loss_ref

pytestmark = pytest.mark.gpu
HEADS = ["reg", "height", "dim", "rot", "vel", "hm"]


def random_sites(rng, B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = np.sort(rng.choice(cells, size=min(n, cells), replace=False))
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


def close(got, want, rel=1e-4, what="", atol=1e-7):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = float((got - want).abs().max()) if got.numel() else 0.0
    ref = float(want.abs().max()) if want.numel() else 0.0
    assert err <= rel * ref + atol, "%s: max |err| %.3e vs max |ref| %.3e (rel tol %g)" % (what, err, ref, rel)


def close_l2(got, want, rel, what=""):
    """Relative L2 check.  Used where a comparison is ReLU-flip sensitive: with training-mode BatchNorm over tiny
    batches, two fp32 evaluations that differ by rounding can disagree on the sign of a pre-activation that sits within
    ~1e-6 of zero, which changes individual gradient entries by percents (the CPU oracle evaluated in fp32 and in fp64
    shows the same effect) while leaving the tensor as a whole intact."""
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err, ref = float((got - want).norm()), float(want.norm())
    # + 1e-4: conv biases in front of a BatchNorm have a mathematically zero gradient (pure rounding noise)
    assert err <= rel * ref + 1e-4, "%s: |err|_2 %.3e vs |ref|_2 %.3e (rel tol %g)" % (what, err, ref, rel)


# ------------------------------------------------------------------------------------------------ unit: BN
@pytest.mark.parametrize("C,relu,res", [(16, True, True), (64, True, False), (384, False, False), (5, True, True)])
def test_batchnorm_train_forward_backward(cuda, C, relu, res):
    rng = np.random.default_rng(C)
    n, cap = 1234, 1300
    x = torch.from_numpy(rng.standard_normal((n, C)).astype(np.float32) * 2 + 0.5)
    r = torch.from_numpy(rng.standard_normal((n, C)).astype(np.float32))
    gy = torch.from_numpy(rng.standard_normal((n, C)).astype(np.float32))
    bn = nn.BatchNorm1d(C, eps=1e-3, momentum=0.01)
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)))
        bn.bias.copy_(torch.from_numpy(rng.standard_normal(C).astype(np.float32)))
        bn.running_mean.copy_(torch.from_numpy(rng.standard_normal(C).astype(np.float32)))
    ref = nn.BatchNorm1d(C, eps=1e-3, momentum=0.01)
    ref.load_state_dict(bn.state_dict())
    xr, rr = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    y = ref(xr) + (rr if res else 0)
    y = F.relu(y) if relu else y
    y.backward(gy)
    bn.to(cuda).train()
    xg = torch.zeros((cap, C), device=cuda); xg[:n] = x.to(cuda)
    rg = torch.zeros((cap, C), device=cuda); rg[:n] = r.to(cuda)
    gg = torch.zeros((cap, C), device=cuda); gg[:n] = gy.to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    saved = T.bn_train_stats(xg, bn, nd, cap)
    yg = T.affine_act(xg, saved.scale, saved.shift, rg if res else None, relu, None, nd, cap)
    close(yg[:n], y, 1e-5, "bn forward")
    close(bn.running_mean, ref.running_mean, 1e-5, "running_mean")
    close(bn.running_var, ref.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == 1
    dgamma, dbeta = torch.empty(C, device=cuda), torch.empty(C, device=cuda)
    dx, dres = T.bn_backward(gg, yg, relu, xg, saved, bn.weight.detach(), dgamma, dbeta, res, nd, cap)
    close(dx[:n], xr.grad, 2e-4, "bn dx")
    close(dgamma, ref.weight.grad, 1e-4, "dgamma")
    close(dbeta, ref.bias.grad, 1e-4, "dbeta")
    if res:
        close(dres[:n], rr.grad, 1e-6, "dres")
    out = torch.empty(C, device=cuda)
    close(T.col_sum(gg, out, nd, cap), gy.sum(0), 1e-5, "col_sum")
    a = xg.clone()
    T.add_rows_(a, gg, nd, cap)
    close(a[:n], x + gy, 1e-6, "add_rows")


# ------------------------------------------------------------------------------------------------ unit: sparse conv
@pytest.mark.parametrize("cin,cout", [(5, 16), (16, 32), (64, 64), (128, 128)])
@pytest.mark.parametrize("strided", [False, True])
def test_sparse_conv_backward(cuda, cin, cout, strided):
    rng = np.random.default_rng(cin * 7 + cout + int(strided))
    shape, B = [9, 20, 20], 2
    c = random_sites(rng, B, shape, 2500)
    n = len(c)
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32))
    cap = n + 77
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    if strided:
        oc, oshape, nbr = S.conv_rulebook(c, B, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1])
        rb, _ = ops.rulebook_conv(ct, nd, cap, B, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1])
        n_out = len(oc)
    else:
        nbr = S.subm_rulebook(c, shape, [3, 3, 3])
        rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
        n_out = n
    gy = torch.from_numpy(rng.standard_normal((n_out, cout)).astype(np.float32))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    S.indice_conv(xr, wr, nbr, n_out).backward(gy)
    xg = torch.zeros((cap, cin), device=cuda); xg[:n] = x.to(cuda)
    gg = torch.zeros((rb.n_out_cap, cout), device=cuda); gg[:n_out] = gy.to(cuda)
    dw = torch.zeros((27, cin, cout), device=cuda)
    T.sparse_conv_wgrad(xg, gg, rb, dw)
    close(dw, wr.grad, 2e-4, "wgrad")
    dw3 = torch.zeros((27, cin, cout), device=cuda)            # tcgen05 arm (CUDA-core kernel when Cin % 8 != 0)
    T.sparse_conv_wgrad(xg, gg, rb, dw3, precision="bf16x3")
    close(dw3, wr.grad, 3e-4, "wgrad bf16x3")
    # deterministic accumulation (per-chunk partial tiles + ordered reduce): bit-identical run to run, on both arms,
    # and equal (to rounding) to the atomics path
    for prec, first in (("fp32", dw), ("bf16x3", dw3)):
        again = torch.zeros((27, cin, cout), device=cuda)
        T.sparse_conv_wgrad(xg, gg, rb, again, precision=prec)
        assert torch.equal(again, first), "weight gradient not bit-reproducible (%s)" % prec
    assert T.DETERMINISTIC_WGRAD
    T.DETERMINISTIC_WGRAD = False
    try:
        dwa = torch.zeros((27, cin, cout), device=cuda)
        T.sparse_conv_wgrad(xg, gg, rb, dwa, precision="bf16x3")
    finally:
        T.DETERMINISTIC_WGRAD = True
    close(dwa, dw3, 1e-5, "atomics vs ordered reduce")
    wg = w.to(cuda)
    if strided:
        nbr_t = T.rulebook_transpose(rb, cap)
        want_t = np.full((27, n), -1, np.int32)
        for k in range(27):
            o = np.nonzero(nbr[k] >= 0)[0]
            want_t[k, nbr[k][o]] = o
        assert np.array_equal(nbr_t[:, :n].cpu().numpy(), want_t)          # transposed rulebook is bit-exact
        table, wt = T.TableView(nbr_t, 27, nd, cap), wg.transpose(1, 2).contiguous()
    else:
        table, wt = rb, wg.flip(0).transpose(1, 2).contiguous()
    dx = ops.sparse_conv(gg, wt, table, precision="fp32")
    close(dx[:n], xr.grad, 2e-4, "dgrad")
    if ops.tc_supported(cout, 27):
        dx3 = ops.sparse_conv(gg, wt, table, precision="bf16x3")
        close(dx3[:n], xr.grad, 1e-3, "dgrad bf16x3")


# ------------------------------------------------------------------------------------------------ unit: dense conv
@pytest.mark.parametrize("kind", ["3x3", "3x3s2", "1x1", "convT", "sliced"])
def test_conv2d_backward(cuda, kind):
    rng = np.random.default_rng(len(kind))
    B, H, W, cin, cout = 2, 12, 10, 24, 40
    x = torch.from_numpy(rng.standard_normal((B, cin, H, W)).astype(np.float32)).requires_grad_(True)
    if kind == "convT":
        conv = nn.ConvTranspose2d(cin, cout, 2, stride=2, bias=False)
        pad, y = (0, 0), None
        y = conv(x)
    elif kind == "3x3s2":
        conv = nn.Conv2d(cin, cout, 3, stride=2, bias=False)
        pad = (1, 1)
        y = conv(F.pad(x, (1, 1, 1, 1)))
    elif kind == "1x1":
        conv = nn.Conv2d(cin, cout, 1, bias=True)
        pad = (0, 0)
        y = conv(x)
    else:
        conv = nn.Conv2d(cin, cout, 3, padding=1, bias=True)
        pad = (1, 1)
        y = conv(x)
    gy = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    transposed = kind == "convT"
    w = train.conv_weight_kio(conv).to(cuda)
    xg = x.detach().permute(0, 2, 3, 1).contiguous().to(cuda)
    gyg = gy.permute(0, 2, 3, 1).contiguous().to(cuda)
    if kind == "sliced":        # dL/dy handed over as a channel slice of a wider buffer (multi-head output tensor)
        wide = torch.zeros(gyg.shape[:3] + (cout + 9,), device=cuda)
        wide[..., 5:5 + cout] = gyg
        gyg = wide[..., 5:5 + cout]
    K = w.shape[0]
    dw = torch.zeros((K, cin, cout), device=cuda)
    T.conv2d_wgrad(xg, gyg, dw, tuple(conv.kernel_size), tuple(conv.stride), pad, transposed)
    g4 = dw.view(conv.kernel_size[0], conv.kernel_size[1], cin, cout)
    got_w = g4.permute(2, 3, 0, 1) if transposed else g4.permute(3, 2, 0, 1)
    close(got_w, conv.weight.grad, 2e-4, "wgrad " + kind)
    dw3 = torch.zeros((K, cin, cout), device=cuda)
    T.conv2d_wgrad(xg, gyg, dw3, tuple(conv.kernel_size), tuple(conv.stride), pad, transposed, precision="bf16x3")
    close(dw3, dw, 3e-4, "wgrad bf16x3 " + kind)
    wt = w.transpose(1, 2).contiguous()
    if transposed:
        dx = ops.conv2d_nhwc(gyg, wt, (2, 2), (2, 2), (0, 0), precision="fp32")
    else:
        dx = T.conv2d_dgrad(gyg, wt, (H, W), tuple(conv.kernel_size), tuple(conv.stride), pad)
    close(dx.permute(0, 3, 1, 2), x.grad, 2e-4, "dgrad " + kind)
    if conv.bias is not None:
        db = torch.empty(cout, device=cuda)
        close(T.col_sum(gyg, db), conv.bias.grad, 1e-4, "bias grad")


@pytest.mark.parametrize("cin,cout", [(256, 256), (128, 256), (512, 64), (64, 8)])
def test_conv2d_wgrad_tensor_core_wide(cuda, cin, cout):
    """Neck / head sized weight gradients (several 128-wide ci / co tiles, partial N) on the tcgen05 arm vs the exact
    CUDA-core arm."""
    rng = np.random.default_rng(cin + cout)
    B, H, W = 2, 20, 18
    x = torch.from_numpy(rng.standard_normal((B, H, W, cin)).astype(np.float32)).to(cuda)
    gy = torch.from_numpy(rng.standard_normal((B, H, W, cout)).astype(np.float32)).to(cuda)
    dw32 = torch.zeros((9, cin, cout), device=cuda)
    dw3 = torch.zeros((9, cin, cout), device=cuda)
    T.conv2d_wgrad(x, gy, dw32, (3, 3), (1, 1), (1, 1))
    T.conv2d_wgrad(x, gy, dw3, (3, 3), (1, 1), (1, 1), precision="bf16x3")
    close(dw3, dw32, 2e-4, "wide wgrad")


@pytest.mark.parametrize("cin,cout", [(128, 128), (256, 256)])
def test_conv_transpose_wgrad_tensor_core(cuda, cin, cout):
    """ConvTranspose2d(k == s == 2) weight gradient (the RPN deblocks): every phase runs the output-stationary tcgen05
    kernel over the phase's pixels of dL/dy; against the exact CUDA-core arm."""
    rng = np.random.default_rng(cin)
    B, H, W = 2, 11, 9
    x = torch.from_numpy(rng.standard_normal((B, H, W, cin)).astype(np.float32)).to(cuda)
    gy = torch.from_numpy(rng.standard_normal((B, 2 * H, 2 * W, cout)).astype(np.float32)).to(cuda)
    dw32 = torch.zeros((4, cin, cout), device=cuda)
    dw3 = torch.zeros((4, cin, cout), device=cuda)
    T.conv2d_wgrad(x, gy, dw32, (2, 2), (2, 2), (0, 0), transposed=True)
    T.conv2d_wgrad(x, gy, dw3, (2, 2), (2, 2), (0, 0), transposed=True, precision="bf16x3")
    close(dw3, dw32, 2e-4, "convT wgrad")
    again = torch.zeros((4, cin, cout), device=cuda)
    T.conv2d_wgrad(x, gy, again, (2, 2), (2, 2), (0, 0), transposed=True, precision="bf16x3")
    assert torch.equal(again, dw3)


@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 128), (128, 64), (16, 16)])
def test_sparse_wgrad_output_stationary_shapes(cuda, cin, cout):
    """The output-stationary tcgen05 weight gradient (2 / 4 / 8 kernel offsets stacked in the accumulator lanes, several
    passes, a chunk tail that is not a multiple of 64 rows) against the exact CUDA-core arm; the caller-provided split
    copies (fd_affine_act / fd_bn_backward outputs in training) must give the same bits as the internal pre-pass."""
    from futuredet_b200 import lib as L
    rng = np.random.default_rng(cin * 3 + cout)
    shape, B = [9, 24, 24], 2
    c = random_sites(rng, B, shape, 3300)
    n = len(c)
    cap = n + 41
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    x = torch.zeros((cap, cin), device=cuda); x[:n] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)).to(cuda)
    gy = torch.zeros((cap, cout), device=cuda); gy[:n] = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32)).to(cuda)
    dw32 = torch.zeros((27, cin, cout), device=cuda)
    dw3 = torch.zeros((27, cin, cout), device=cuda)
    T.sparse_conv_wgrad(x, gy, rb, dw32)
    T.sparse_conv_wgrad(x, gy, rb, dw3, precision="bf16x3")
    close(dw3, dw32, 3e-4, "output-stationary wgrad")
    # split copies made by the conversion entry point == what the kernel's own pre-pass makes
    lib = L.load()
    xs, gs = torch.empty_like(x), torch.empty_like(gy)
    for src, dst, C_ in ((x, xs, cin), (gy, gs, cout)):
        rc = lib.fd_convert_rows(ops._ptr(src), 0, C_, C_, ops._ptr(dst), 1, C_, C_, C_, None, cap, ops._stream())
        L.check(rc, "fd_convert_rows")
    dws = torch.zeros((27, cin, cout), device=cuda)
    T.sparse_conv_wgrad(x, gy, rb, dws, precision="bf16x3", x_split=xs, dy_split=gs)
    assert torch.equal(dws, dw3), "caller-provided split copies change the weight gradient"


@pytest.mark.parametrize("case", ["empty", "tiny", "strided", "wide_cin", "k1"])
def test_sparse_wgrad_output_stationary_edges(cuda, case):
    """Edge cases of the output-stationary weight gradient: no active rows (device count 0 under a large capacity), fewer
    rows than one 64-row stage, a strided rulebook (input and output row sets differ), Cin = 256 (two 128-channel slices
    per offset) and a single kernel offset."""
    rng = np.random.default_rng(len(case))
    shape, B = [9, 24, 24], 2
    n_sites = {"empty": 200, "tiny": 37}.get(case, 2100)
    cin, cout = {"wide_cin": (256, 64), "k1": (64, 32)}.get(case, (64, 64))
    c = random_sites(rng, B, shape, n_sites)
    n = len(c)
    cap = n + 300
    ct = torch.zeros((cap, 4), dtype=torch.int32, device=cuda); ct[:n] = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([0 if case == "empty" else n], dtype=torch.int32, device=cuda)
    if case == "strided":
        rb, _ = ops.rulebook_conv(ct, nd, cap, B, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1])
    elif case == "k1":
        rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [1, 1, 1], batch_size=B)
    else:
        rb, _ = ops.rulebook_subm(ct, nd, cap, shape, [3, 3, 3], batch_size=B)
    K = rb.K
    x = torch.from_numpy(rng.standard_normal((cap, cin)).astype(np.float32)).to(cuda)
    gy = torch.from_numpy(rng.standard_normal((rb.n_out_cap, cout)).astype(np.float32)).to(cuda)
    dw32 = torch.zeros((K, cin, cout), device=cuda)
    dw3 = torch.zeros((K, cin, cout), device=cuda)
    T.sparse_conv_wgrad(x, gy, rb, dw32)
    T.sparse_conv_wgrad(x, gy, rb, dw3, precision="bf16x3")
    if case == "empty":
        assert float(dw3.abs().max()) == 0.0 and float(dw32.abs().max()) == 0.0
    else:
        assert float(dw32.abs().max()) > 0
        close(dw3, dw32, 3e-4, "wgrad " + case)
    again = torch.zeros((K, cin, cout), device=cuda)
    T.sparse_conv_wgrad(x, gy, rb, again, precision="bf16x3")
    assert torch.equal(again, dw3)


def test_split_copies_leave_training_bit_identical(cuda, golden_dir):
    """train.SPLIT_COPIES (activations / conv-output gradients also written as split-bf16 rows by their producers and
    gathered by the tensor-core convolutions) is a data-movement change only: same losses, same gradients, bit for bit."""
    g = torch.load(os.path.join(golden_dir, "neck_head_train.pt"), weights_only=False)
    assert train.SPLIT_COPIES
    m_on, x_on, l_on = _neck_head_step(g, cuda, "bf16x3", smooth=True)
    train.SPLIT_COPIES = False
    try:
        m_off, x_off, l_off = _neck_head_step(g, cuda, "bf16x3", smooth=True)
    finally:
        train.SPLIT_COPIES = True
    assert torch.equal(sum(l_on["loss"]), sum(l_off["loss"]))
    assert torch.equal(x_on.grad, x_off.grad)
    for (k, p_on), (_, p_off) in zip(m_on.named_parameters(), m_off.named_parameters()):
        assert torch.equal(p_on.grad, p_off.grad), k


def test_bev_scatter_gather(cuda):
    rng = np.random.default_rng(3)
    B, D, H, W, Cc = 2, 2, 7, 9, 12
    c = random_sites(rng, B, [D, H, W], 100)
    n = len(c)
    rows = torch.from_numpy(rng.standard_normal((n, Cc)).astype(np.float32))
    dense = torch.zeros((B, Cc, D, H, W))
    ci = torch.from_numpy(c.astype(np.int64))
    dense[ci[:, 0], :, ci[:, 1], ci[:, 2], ci[:, 3]] = rows
    want = dense.view(B, Cc * D, H, W)                                    # scn.py:165-168
    ct = torch.from_numpy(c).to(cuda)
    nd = torch.tensor([n], dtype=torch.int32, device=cuda)
    bev = T.rows_to_bev(rows.to(cuda), ct, nd, n, B, D, H, W)
    assert torch.equal(bev.permute(0, 3, 1, 2).cpu(), want)
    back = T.bev_to_rows(bev, Cc, ct, nd, n, B, D, H, W)
    assert torch.equal(back.cpu(), rows)


# ------------------------------------------------------------------------------------------------ loss backward
@pytest.mark.parametrize("timesteps", [1, 3])
def test_center_head_loss_backward(cuda, timesteps):
    from oracle.gen_golden import make_targets
    gen = torch.Generator().manual_seed(5 + timesteps)
    B, H, W = 2, 10, 12
    chans = dict(reg=2, height=1, dim=3, rot=2, vel=2 * timesteps, hm=1)
    total = sum(chans.values())
    out = torch.randn((B, H, W, total), generator=gen)
    out[..., -1] = out[..., -1] * 3 - 2                       # hm logits, some beyond the sigmoid clamp
    out[0, 0, 0, -1], out[0, 0, 1, -1] = 15.0, -15.0
    example = make_targets(B, H, W, timesteps, gen, max_objs=15)
    example["ind"][0][0][0, 1] = example["ind"][0][0][0, 0]    # duplicate centre: gradients must accumulate
    example["mask"][0][0][0, :2] = 1
    head = type("H", (), dict(timesteps=timesteps, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0], weight=0.25))()
    head.code_weights_forecast = list(np.array(head.code_weights) * np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 0]))

    def views(base):
        ret, col = {}, 0
        for k, c in chans.items():
            ret[k] = base[..., col:col + c].permute(0, 3, 1, 2)
            col += c
        return ret

    ref_out = out.clone().requires_grad_(True)
    ref_loss = center_head_loss_ref([views(ref_out)], example, timesteps, head.code_weights, head.weight)
    (sum(ref_loss["loss"]) * 1.7).backward()
    og = out.to(cuda)
    losses, ctxs = center_head_loss(head, example, [views(og)], return_ctx=True)
    close(losses["loss"][0], ref_loss["loss"][0], 1e-5, "loss")
    gout = torch.zeros_like(og)
    center_head_loss_backward(ctxs[0], og, gout, torch.tensor([1.7], device=cuda))
    close(gout, ref_out.grad, 1e-4, "dL/dpreds")
    assert float(gout[0, 0, 0, -1]) == 0.0 and float(gout[0, 0, 1, -1]) == 0.0     # clamped logits carry no gradient


# ------------------------------------------------------------------------------------------------ neck + head golden
class _NeckHead(nn.Module):
    def __init__(self, neck, head):
        super().__init__()
        self.neck, self.bbox_head = neck, head


def _neck_head_step(g, cuda, precision, smooth=False):
    neck = fb.build_neck(dict(g["neck_cfg"]))
    head = fb.build_head(dict(g["head_cfg"]))
    neck.load_state_dict(g["neck_state"]); head.load_state_dict(g["head_state"])
    if smooth:
        g0 = torch.Generator().manual_seed(5)
        for mod in list(neck.modules()) + list(head.modules()):
            if isinstance(mod, nn.modules.batchnorm._BatchNorm):
                mod.weight.data.copy_(0.4 + 0.2 * torch.rand(mod.weight.shape, generator=g0))
                mod.bias.data.copy_(1.5 + 0.3 * torch.rand(mod.bias.shape, generator=g0))
    m = _NeckHead(neck, head).to(cuda).train()
    tr = train.NativeTrainer(m, precision=precision)
    tape = train.Tape()
    x = train.Var(g["x"].permute(0, 2, 3, 1).contiguous().to(cuda))
    feat = tr._neck(tape, m.neck, x)
    preds, outs = tr._head(tape, m.bbox_head, feat)
    losses, ctxs = center_head_loss(m.bbox_head, g["example"], preds, return_ctx=True)
    tr.tape, tr._loss_ctx = tape, (ctxs, outs)
    tr.backward()
    return m, x, losses


def test_tensor_core_training_matches_fp32_training(cuda, golden_dir):
    """precision="bf16x3" (forward / data-gradient convolutions on tcgen05) against the exact-fp32 arm of the same
    native training step, on a well-conditioned instance (ReLU thresholds at -3 sigma): element-wise 1e-3."""
    g = torch.load(os.path.join(golden_dir, "neck_head_train.pt"), weights_only=False)
    m32, x32, l32 = _neck_head_step(g, cuda, "fp32", smooth=True)
    m3, x3, l3 = _neck_head_step(g, cuda, "bf16x3", smooth=True)
    close(sum(l3["loss"]), sum(l32["loss"]), 1e-4, "loss")
    close(x3.grad, x32.grad, 1e-3, "dL/dx")
    for (k, p3), (_, p32) in zip(m3.named_parameters(), m32.named_parameters()):
        close(p3.grad, p32.grad, 1e-3, k, atol=2e-6)


def test_rpn_center_head_training_matches_reference_golden(cuda, golden_dir):
    """Exact-fp32 arm against the gradients of the REFERENCE RPN + CenterHead classes (training mode), element-wise.
    The tensor-core arm is pinned to this arm by test_tensor_core_training_matches_fp32_training on a well-conditioned
    instance: this fixture (randomised BatchNorm, ~24 objects) is ReLU-flip sensitive at the 5e-6 level of bf16x3."""
    precision = "fp32"
    g = torch.load(os.path.join(golden_dir, "neck_head_train.pt"), weights_only=False)
    m, x, losses = _neck_head_step(g, cuda, precision)
    # fp32: element-wise against the reference's gradients; bf16x3 (forward / data-gradient convolutions on the tensor
    # cores, ~5e-6 relative): relative L2, the fixture's randomised BatchNorm makes it ReLU-flip sensitive
    if precision == "fp32":
        check = lambda got, want, what: close(got, want, 2e-4, what, atol=2e-6)   # biases in front of a BN: ~0 gradient
    else:
        check = lambda got, want, what: close_l2(got, want, 5e-2, what)
    close(sum(losses["loss"]), g["total"], 1e-5 if precision == "fp32" else 1e-4, "total loss")
    for k in ("hm_loss", "num_positive"):
        close(torch.as_tensor(losses[k][0]), torch.as_tensor(g["loss"][k][0]), 1e-4, k)
    check(x.grad.permute(0, 3, 1, 2), g["x_grad"], "dL/dx")
    named = {"neck." + k: p for k, p in m.neck.named_parameters()}
    named.update({"head." + k: p for k, p in m.bbox_head.named_parameters()})
    assert set(named) == set(g["grads"])
    for k, want in g["grads"].items():
        check(named[k].grad, want, k)
    for mod, after in ((m.neck, g["neck_state_after"]), (m.bbox_head, g["head_state_after"])):
        sd = mod.state_dict()
        for k, v in after.items():
            if "running" in k or "num_batches" in k:
                close(sd[k].float(), v.float(), 1e-4, k)


# ------------------------------------------------------------------------------------------------ whole model
def build_model(timesteps, dev, tasks=None):
    torch.manual_seed(0)
    tasks = tasks or [dict(num_class=1, class_names=["car"])]
    cfg = dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(type="CenterHead", in_channels=512, tasks=tasks,
                       dataset="nuscenes", weight=0.25, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                       share_conv_channel=64, dcn_head=False, timesteps=timesteps, classify=False))
    return fb.build_detector(cfg)


def test_two_task_head_train_step_matches_oracle(cuda):
    """BASELINE configs[4] shape: mixed car + pedestrian heads (two SepHeads on the shared feature, per-task targets and
    losses, trainer.py:85 sums them): loss per task and every gradient vs autograd over the oracle."""
    from oracle.gen_golden import make_targets
    rng = np.random.default_rng(8)
    tasks = [dict(num_class=1, class_names=["car"]), dict(num_class=1, class_names=["pedestrian"])]
    model = build_model(3, cuda, tasks=tasks)
    g0 = torch.Generator().manual_seed(21)
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.weight.data.copy_(0.4 + 0.2 * torch.rand(m.weight.shape, generator=g0))
            m.bias.data.copy_(1.5 + 0.3 * torch.rand(m.bias.shape, generator=g0))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    B, grid = 1, [64, 64, 40]
    c = random_sites(rng, B, [40, 64, 64], 4000)
    n = len(c)
    feats = rng.standard_normal((n, 5)).astype(np.float32)
    voxels = np.zeros((n, 10, 5), np.float32); voxels[:, 0] = feats
    ta = make_targets(B, 8, 8, 3, torch.Generator().manual_seed(1), max_objs=10)
    tb = make_targets(B, 8, 8, 3, torch.Generator().manual_seed(2), max_objs=10)
    example = {k: [[ta[k][t][0], tb[k][t][0]] for t in range(3)] for k in ta}            # [timestep][task]
    example.update(voxels=torch.from_numpy(voxels).to(cuda), num_points=torch.ones(n, dtype=torch.int32, device=cuda),
                   coordinates=torch.from_numpy(c).to(cuda), num_voxels=torch.tensor([0] * B), shape=[np.array(grid)] * B)
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    bev = S.backbone_forward(sub("backbone."), torch.from_numpy(feats), c, B, grid, bn_eval=S.bn_train(0.01))
    feat = D.rpn_forward(sub("neck."), bev, [5, 5], [1, 2], [1, 2], train=0.01)
    preds = D.center_head_forward(sub("bbox_head."), feat, [HEADS, HEADS], train=0.1)
    ref = center_head_loss_ref(preds, example, 3, [1.0] * 6 + [0.2, 0.2, 1.0, 1.0], 0.25)
    sum(ref["loss"]).backward()
    model.to(cuda).train()
    tr = train.NativeTrainer(model, precision="bf16x3")
    losses = tr.step(example)
    assert len(losses["loss"]) == 2
    for t_id in range(2):
        close(losses["loss"][t_id], ref["loss"][t_id], 1e-3, "loss task %d" % t_id)
    bad = []
    for k, p in model.named_parameters():
        want = sd[k].grad
        err, refmax = float((p.grad.cpu() - want).abs().max()), float(want.abs().max())
        if err > 5e-3 * refmax + 2e-6:
            bad.append((k, "%.1e" % (err / max(refmax, 1e-30))))
    assert not bad, bad


@pytest.mark.parametrize("timesteps,precision,init", [(3, "fp32", "smooth"), (7, "bf16x3", "smooth"),
                                                      (3, "fp32", "default")])
def test_voxelnet_train_step_matches_oracle_autograd(cuda, timesteps, precision, init):
    """forecast_n3-shaped model (multi-timestep head), forward + backward of one batch: loss, every parameter
    gradient and every BatchNorm running statistic vs torch autograd over the chained CPU oracle.
    init="smooth": BatchNorm affine parameters put the ReLU thresholds at about -3 sigma, which makes the comparison
    well conditioned (fp32 and fp64 runs of the oracle then agree to 2e-5) -> element-wise tolerance;
    init="default": the modules' own initialisation (thresholds at the mode of the pre-activations, flip sensitive)
    -> relative L2 per parameter."""
    from oracle.gen_golden import make_targets
    rng = np.random.default_rng(1)
    model = build_model(timesteps, cuda)
    if init == "smooth":
        g0 = torch.Generator().manual_seed(11)
        for m in model.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.weight.data.copy_(0.4 + 0.2 * torch.rand(m.weight.shape, generator=g0))
                m.bias.data.copy_(1.5 + 0.3 * torch.rand(m.bias.shape, generator=g0))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    B, grid = 2, [64, 64, 40]                                   # (x, y, z) -> BEV 8 x 8
    c = random_sites(rng, B, [40, 64, 64], 6000)                # z < 40: voxel coordinates never use the extra plane
    n = len(c)
    feats = rng.standard_normal((n, 5)).astype(np.float32)
    voxels = np.zeros((n, 10, 5), np.float32); voxels[:, 0] = feats
    gen = torch.Generator().manual_seed(3)
    example = make_targets(B, 8, 8, timesteps, gen, max_objs=12)
    example.update(voxels=torch.from_numpy(voxels).to(cuda), num_points=torch.ones(n, dtype=torch.int32, device=cuda),
                   coordinates=torch.from_numpy(c).to(cuda), num_voxels=torch.tensor([0] * B),
                   shape=[np.array(grid)] * B)
    # ---- oracle: autograd over the functional restatements, training-mode BatchNorm
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    bev = S.backbone_forward(sub("backbone."), torch.from_numpy(feats), c, B, grid, bn_eval=S.bn_train(0.01))
    feat = D.rpn_forward(sub("neck."), bev, [5, 5], [1, 2], [1, 2], train=0.01)
    preds = D.center_head_forward(sub("bbox_head."), feat, [HEADS], train=0.1)
    ref = center_head_loss_ref(preds, example, timesteps, [1.0] * 6 + [0.2, 0.2, 1.0, 1.0], 0.25)
    sum(ref["loss"]).backward()
    # ---- native
    model.to(cuda).train()
    tr = train.NativeTrainer(model, precision=precision)
    n0 = fb.lib.launch_count()
    losses = tr.step(example)
    assert fb.lib.launch_count() - n0 > 300                    # forward + backward really ran in the library
    ltol, gtol = (1e-4, 2e-3) if precision == "fp32" else (1e-3, 5e-3)
    close(sum(losses["loss"]), sum(ref["loss"]), ltol, "loss")
    bad = []
    for k, p in model.named_parameters():
        want = sd[k].grad
        assert want is not None and p.grad is not None, k
        if init == "smooth":
            err, refmax, tol = float((p.grad.cpu() - want).abs().max()), float(want.abs().max()), gtol
        else:
            err, refmax, tol = float((p.grad.cpu() - want).norm()), float(want.norm()), 5e-2
        if err > tol * refmax + (2e-6 if init == "smooth" else 1e-4):
            bad.append((k, err, refmax))
    assert not bad, "%d/%d gradients off: %s" % (len(bad), len(list(model.parameters())), [(k, "%.1e" % (e / max(r, 1e-30))) for k, e, r in bad])
    after = model.state_dict()
    for k, v in sd.items():
        if "running" in k:
            close(after[k], v, 1e-3 if precision == "fp32" else 1e-2, k)
    # second step reuses the buckets: gradients are rebuilt, not accumulated
    g_first = model.neck.blocks[0][1].weight.grad.clone()
    for mod in model.modules():
        if isinstance(mod, nn.modules.batchnorm._BatchNorm):
            mod.momentum = 0.0                                  # keep statistics fixed for the repeat
    tr.step(example)
    close(model.neck.blocks[0][1].weight.grad, g_first, 1e-3, "repeatability")


def test_train_step_on_fused_points_path(cuda):
    """Raw points -> fused voxelizer -> train step: runs end to end, finite loss and gradients, grads live in buckets."""
    from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene
    from oracle.gen_golden import make_targets
    model = build_model(3, cuda).to(cuda).train()
    model.configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                                   max_voxel_num=[120000, 160000]), training=True)
    scene = synth_scene(20000, seed=2)
    pts = torch.from_numpy(scene).to(cuda)
    off = torch.tensor([0, len(scene)], dtype=torch.int32, device=cuda)
    example = make_targets(1, 180, 180, 3, torch.Generator().manual_seed(0), max_objs=50)
    tr = train.NativeTrainer(model, precision="bf16x3")
    losses = tr.step(example, points=pts, batch_offsets=off)
    assert torch.isfinite(losses["loss"][0]).item()
    flat = torch.cat([f for f, _ in tr.grads.buckets])
    assert torch.isfinite(flat).all().item() and float(flat.abs().sum()) > 0
    assert all(p.grad is not None and p.grad.data_ptr() == tr.grads.grad(p).data_ptr() for p in model.parameters())


def test_sorted_tiles_leave_training_bit_identical(cuda, monkeypatch):
    """Pattern-sorted SubM tiles (forward and data-gradient convolutions of a multi-sample step) only change which tile a
    row is computed in: same losses and gradients, bit for bit, as the unsorted step."""
    from futuredet_b200 import ops
    from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene
    from oracle.gen_golden import make_targets
    monkeypatch.setattr(ops, "SORT_MIN_ROWS", 0)
    scenes = [synth_scene(20000, seed=s) for s in (2, 3)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(cuda)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=cuda)
    example = make_targets(2, 180, 180, 3, torch.Generator().manual_seed(0), max_objs=50)
    out = []
    for min_batch in (1, 100):
        monkeypatch.setattr(ops, "SORT_MIN_BATCH", min_batch)
        model = build_model(3, cuda).to(cuda).train()
        model.configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                                       max_voxel_num=[120000, 160000]), training=True)
        tr = train.NativeTrainer(model, precision="bf16x3")
        losses = tr.step(example, points=pts, batch_offsets=off)
        out.append((sum(losses["loss"]).clone(), {k: p.grad.clone() for k, p in model.named_parameters()}))
    assert torch.equal(out[0][0], out[1][0])
    for k in out[0][1]:
        assert torch.equal(out[0][1][k], out[1][1][k]), k


def test_reference_trainer_idiom_loss_backward(cuda):
    """`losses = model(example, return_loss=True); sum(losses["loss"]).backward()` (trainer.py:85,317-344) drives the
    native backward through the autograd bridge and fills param.grad with the same gradients as NativeTrainer.step."""
    from oracle.gen_golden import make_targets
    rng = np.random.default_rng(4)
    model = build_model(3, cuda).to(cuda).train().set_precision("fp32")      # compared with the exact-fp32 trainer below
    B, grid = 1, [64, 64, 40]
    c = random_sites(rng, B, [40, 64, 64], 3000)
    n = len(c)
    voxels = np.zeros((n, 10, 5), np.float32); voxels[:, 0] = rng.standard_normal((n, 5)).astype(np.float32)
    example = make_targets(B, 8, 8, 3, torch.Generator().manual_seed(9), max_objs=12)
    example.update(voxels=torch.from_numpy(voxels).to(cuda), num_points=torch.ones(n, dtype=torch.int32, device=cuda),
                   coordinates=torch.from_numpy(c).to(cuda), num_voxels=torch.tensor([0] * B), shape=[np.array(grid)] * B)
    losses = model(example, return_loss=True)
    loss = sum(losses["loss"])
    assert loss.requires_grad
    (2.0 * loss).backward()
    got = {k: p.grad.clone() for k, p in model.named_parameters()}
    assert all(g is not None for g in got.values())
    tr = train.NativeTrainer(model, precision="fp32")
    ref_losses = tr.step(example)
    close(loss, sum(ref_losses["loss"]), 1e-5, "bridged loss")
    for k, p in model.named_parameters():
        close(got[k], 2.0 * p.grad, 2e-3, k, atol=2e-6)      # wgrad accumulates with fp32 atomics: order-level noise
