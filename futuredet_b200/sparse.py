"""spconv-1.x-shaped API over the native rulebook + implicit-GEMM kernels.

Replaces the external `spconv` package the reference backbone imports (det3d/models/backbones/scn.py:2-3):
`SparseConvTensor`, `SubMConv3d`, `SparseConv3d`, `SparseSequential`, `SparseModule`, with the semantics
summarised in SURVEY.md appendix B.  Parameters keep spconv 1.x's layout (`weight [kD,kH,kW,Cin,Cout]`,
`bias [Cout]`) so reference checkpoints load unchanged.

Differences that matter to callers:
  * row counts live on the device (`n_dev`), buffers are allocated at a static capacity (`n_cap`) and rows
    >= n are undefined -- nothing in a forward pass synchronises with the host;
  * `SparseSequential` fuses conv -> BatchNorm1d(eval) -> ReLU triples into the conv epilogue;
  * only inference (module.eval()) is implemented natively in this round; train-mode BatchNorm raises.
"""
import math

import torch
from torch import nn

from . import ops
from . import precision as _precision


def _triple(v):
    return [int(v)] * 3 if isinstance(v, int) else [int(x) for x in v]


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, n_dev=None, n_cap=None):
        if indices.dtype != torch.int32:
            indices = indices.int()
        self._feat = ops.as_feat(features)     # ops.Feat: fp32 rows, or split bf16 hi/lo rows between tensor-core layers
        self._fp32 = None
        self.indices = indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.n_cap = int(indices.shape[0] if n_cap is None else n_cap)
        if n_dev is None:
            n_dev = torch.full((1,), self.n_cap, dtype=torch.int32, device=indices.device)
        self.n_dev = n_dev
        self.indice_dict = {}     # indice_key -> (Rulebook, out spatial shape)
        self._index = None        # CoordIndex of self.indices

    @property
    def features(self):
        """[n_cap, C] fp32 features (converted on demand when the tensor-core pipeline holds them split)."""
        if self._feat.fmt == "fp32":
            return self._feat.to_fp32()
        if self._fp32 is None:
            self._fp32 = self._feat.to_fp32(self.n_dev)
        return self._fp32

    @features.setter
    def features(self, value):
        self._feat = ops.as_feat(value)
        self._fp32 = None

    def coord_index(self):
        if self._index is None:
            self._index = ops.CoordIndex(self.indices, self.n_dev, self.n_cap, self.spatial_shape, self.batch_size)
        return self._index

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def num_active(self):
        """Host-side row count (synchronises)."""
        return int(self.n_dev.item())

    def trimmed(self):
        n = self.num_active()
        return self.features[:n], self.indices[:n]

    def dense(self, channels_first=True):
        """[B, C, D, H, W] (zeros at inactive sites), as spconv's `.dense()` (scn.py:165)."""
        d = ops.sparse_to_dense(self._feat, self.indices, self.n_dev, self.n_cap, self.batch_size,
                                self.spatial_shape)
        return d if channels_first else d.permute(0, 2, 3, 4, 1).contiguous()

    def _like(self, features, indices=None, spatial_shape=None, n_dev=None, n_cap=None):
        same_sites = indices is None
        t = SparseConvTensor(features, self.indices if same_sites else indices,
                             self.spatial_shape if spatial_shape is None else spatial_shape, self.batch_size,
                             self.n_dev if same_sites else n_dev, self.n_cap if same_sites else n_cap)
        t.indice_dict = self.indice_dict
        if same_sites:
            t._index = self._index
        return t


class SparseModule(nn.Module):
    """Marker: SparseSequential passes the SparseConvTensor itself to these (scn.py:37)."""


class _SparseConvBase(SparseModule):
    subm = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None):
        super().__init__()
        if _triple(dilation) != [1, 1, 1] or groups != 1:
            raise NotImplementedError("dilation/groups are not used by the reference backbone")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.precision = None     # None -> precision.default_precision()
        self.reset_parameters()

    def reset_parameters(self):
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = math.sqrt(6.0 / ((1 + 5.0) * fan_in))          # kaiming_uniform_(a=sqrt(5)) on fan_in
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                b = 1.0 / math.sqrt(fan_in)
                self.bias.uniform_(-b, b)

    def weight_kio(self, cin_pad=None):
        w = self.weight.detach().reshape(-1, self.in_channels, self.out_channels)
        if cin_pad is None or cin_pad == self.in_channels:
            return w
        key = (self.weight.data_ptr(), self.weight._version, cin_pad)      # zero rows for the padded input channels
        cache = self.__dict__.setdefault("_wpad_cache", {})
        if cache.get("k") != key:
            wp = torch.zeros((w.shape[0], cin_pad, self.out_channels), dtype=w.dtype, device=w.device)
            wp[:, :self.in_channels] = w
            cache["k"], cache["v"] = key, wp
        return cache["v"]

    def rulebook(self, x):
        cached = x.find_indice_pair(self.indice_key)
        if cached is not None:
            return cached
        if self.subm:
            rb, idx = ops.rulebook_subm(x.indices, x.n_dev, x.n_cap, x.spatial_shape, self.kernel_size,
                                        index=x.coord_index())
        else:
            rb, idx = ops.rulebook_conv(x.indices, x.n_dev, x.n_cap, x.batch_size, x.spatial_shape,
                                        self.kernel_size, self.stride, self.padding, index=x.coord_index())
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = rb
        return rb

    def forward(self, x, bn=None, residual=None, relu=False, bev=False, out_fmt=None, bev_dmajor=False):
        """out = act(bn(conv(x) [+bias]) [+ residual]); `bn` is an eval-mode BatchNorm1d folded into the epilogue.
        Tensor-core precisions keep activations in the split bf16 hi/lo row format between layers."""
        rb = self.rulebook(x)
        scale, shift = folded_epilogue(self, bn)
        prec = _precision.resolve(self.precision)
        xin = x._feat
        w = self.weight_kio()
        if prec != "fp32" and self.in_channels % 8 != 0:
            cin_pad = (self.in_channels + 7) // 8 * 8
            if xin.fmt == "fp32" and xin.ctot >= cin_pad and xin.c0 == 0:
                # e.g. the 5-channel stem fed by the fused voxelizer (rows zero-padded to 8 channels)
                xin, w = ops.Feat(xin.t, "fp32", 0, cin_pad), self.weight_kio(cin_pad)
                if ops.STEM_SPLIT and x.batch_size >= ops.SORT_MIN_BATCH:     # (one more launch: not at one scene per forward)
                    # one small conversion pass, then the stem gathers ready-made bf16 hi/lo rows with cp.async like every
                    # other layer instead of splitting fp32 rows in its producer warps
                    xin = ops.to_split(xin, x.n_dev)
            else:
                prec = "fp32"       # no tensor-core tile shape for this Cin: exact fp32 CUDA-core arm, explicitly
        if out_fmt is None:
            out_fmt = "fp32" if prec == "fp32" else "split"
        if bev:
            D, H, W = rb.out_shape
            return ops.sparse_conv(xin, w, rb, scale, shift, None, relu, precision=prec,
                                   bev=(x.batch_size, D, H, W), out_fmt=out_fmt, bev_dmajor=bev_dmajor)
        y = ops.sparse_conv(xin, w, rb, scale, shift, residual, relu, precision=prec, out_fmt=out_fmt,
                            sort_tiles=prec != "fp32" and x.batch_size >= ops.SORT_MIN_BATCH and
                            (self.subm or ops.SORT_STRIDED or rb.row_key is not None))
        if self.subm:
            return x._like(y)
        out = x._like(y, rb.out_coords, rb.out_shape, rb.n_out_dev, rb.n_out_cap)
        out._index = getattr(rb, "out_index", None)       # bitmap index of the new active set (no hash build needed)
        return out


class SubMConv3d(_SparseConvBase):
    subm = True


class SparseConv3d(_SparseConvBase):
    subm = False


def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


def folded_epilogue(conv, bn):
    """(scale, shift) with shift absorbing the conv bias; cached until a parameter/buffer changes."""
    bias = getattr(conv, "bias", None)
    if bn is None and bias is None:
        return None, None
    key = _versions(bias, *( (bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else ()))
    # `_fd_stats_version`: bumped by the native train-mode BatchNorm, which updates the running statistics through raw
    # pointers (tensor version counters do not see it)
    key = (id(bn), bn.training if bn is not None else False, getattr(bn, "_fd_stats_version", 0)) + key
    cache = conv.__dict__.setdefault("_fold_cache", {})
    hit = cache.get("k")
    if hit == key:
        return cache["v"]
    with torch.no_grad():
        if bn is None:
            scale, shift = None, bias.detach().float().contiguous()
        else:
            scale, shift = fold_bn(bn)
            if bias is not None:
                shift = (shift + bias.detach() * scale).contiguous()
    cache["k"], cache["v"] = key, (scale, shift)
    return scale, shift


def fold_bn(bn):
    """Eval-mode BatchNorm -> per-channel (scale, shift)."""
    if bn.training:
        raise NotImplementedError("futuredet_b200: train-mode BatchNorm is not implemented natively yet; "
                                  "call model.eval() (no PyTorch fallback is provided)")
    inv = torch.rsqrt(bn.running_var.detach() + bn.eps)
    scale = bn.weight.detach() * inv if bn.affine else inv
    shift = -bn.running_mean.detach() * scale
    if bn.affine:
        shift = shift + bn.bias.detach()
    return scale.contiguous(), shift.contiguous()


class SparseSequential(SparseModule):
    """Runs sparse modules on the tensor; (conv, BatchNorm1d, ReLU) runs are fused into one kernel."""

    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def add(self, module, name=None):
        self.add_module(name or str(len(self._modules)), module)

    def forward(self, x, bev_last=False, out_fmt=None, bev_dmajor=False):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, _SparseConvBase):
                bn = None
                relu = False
                j = i + 1
                if j < len(mods) and isinstance(mods[j], nn.BatchNorm1d):
                    bn = mods[j]
                    j += 1
                if j < len(mods) and isinstance(mods[j], nn.ReLU):
                    relu = True
                    j += 1
                last = j == len(mods)
                x = m(x, bn=bn, relu=relu, bev=(bev_last and last), out_fmt=out_fmt if last else None,
                      bev_dmajor=bev_dmajor and bev_last and last)
                i = j
            elif isinstance(m, SparseModule):
                x = m(x)
                i += 1
            else:
                raise NotImplementedError(
                    "SparseSequential: %s outside a conv->BatchNorm1d->ReLU run has no native kernel" % type(m).__name__)
        return x
