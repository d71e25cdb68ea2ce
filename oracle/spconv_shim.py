"""ORACLE (test infrastructure only): a CPU `spconv`-1.x-shaped module set over oracle/spconv_ref.py.

Purpose: let the REFERENCE's own backbone source (det3d/models/backbones/scn.py:37-176 -- SparseBasicBlock and
SpMiddleResNetFHD, unmodified, executed from /root/reference by oracle/gen_golden.py) run on the CPU with `spconv`
bound to these classes.  That pins everything scn.py itself decides -- topology, `indice_key` sharing, bias / BatchNorm /
residual / ReLU order, strides and paddings, `dense()` + `view(N, C*D, H, W)` -- to the reference source; only the per-op
arithmetic of spconv (`get_indice_pairs` + `indice_conv`, restated in spconv_ref.py and cross-checked against
torch.conv3d) stays unpinned against the real third-party library (not installable here).

API surface = what scn.py touches: SparseConvTensor(features, indices, spatial_shape,
batch_size) with .features / .indices / .spatial_shape / .batch_size / .indice_dict / .dense(); SubMConv3d / SparseConv3d
(in, out, kernel_size, stride, padding, bias, indice_key) with weight [kD,kH,kW,Cin,Cout]; SparseSequential applying
plain nn modules to `.features`; SparseModule marker.
"""
import math

import numpy as np
import torch
from torch import nn

from . import spconv_ref as S


def _triple(v):
    return [int(v)] * 3 if isinstance(v, int) else [int(x) for x in v]


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def dense(self, channels_first=True):
        C = self.features.shape[1]
        d = torch.zeros((self.batch_size, *self.spatial_shape, C), dtype=self.features.dtype)
        ci = torch.as_tensor(np.asarray(self.indices), dtype=torch.int64)
        d[ci[:, 0], ci[:, 1], ci[:, 2], ci[:, 3]] = self.features
        return d.permute(0, 4, 1, 2, 3).contiguous() if channels_first else d


class SparseModule(nn.Module):
    pass


class _Conv(SparseModule):
    subm = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        fan_in = in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x):
        coords = np.asarray(x.indices, np.int32)
        cached = x.find_indice_pair(self.indice_key)
        if cached is None:
            if self.subm:
                cached = (coords, x.spatial_shape, S.subm_rulebook(coords, x.spatial_shape, self.kernel_size))
            else:
                cached = S.conv_rulebook(coords, x.batch_size, x.spatial_shape, self.kernel_size, self.stride, self.padding)
            if self.indice_key is not None:
                x.indice_dict[self.indice_key] = cached
        oc, oshape, nbr = cached
        y = S.indice_conv(x.features, self.weight, nbr, len(oc))
        if self.bias is not None:
            y = y + self.bias
        out = SparseConvTensor(y, torch.from_numpy(np.ascontiguousarray(oc)), oshape, x.batch_size)
        out.indice_dict = x.indice_dict
        return out


class SubMConv3d(_Conv):
    subm = True


class SparseConv3d(_Conv):
    subm = False


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x.features = m(x.features)
            else:
                x = m(x)
        return x
