"""ORACLE (test infrastructure only): CPU restatement of the spconv-1.x sparse backbone.

*** PARITY UNPINNED against the real spconv library. ***
The arithmetic of this segment lives in third-party `spconv` 1.x (fork github.com/neeharperi/spconv,
"spconv: 1.0" per the reference README.md:26,31,57-59; built from an unpinned local checkout by
setup.sh:27-34; not vendored, no version pin, not installable here -- no network).  The reference
holds no tests or golden vectors for it.  This file restates spconv 1.x's published algorithm
(`get_indice_pairs` + `indice_conv`: per kernel offset gather -> fp32 GEMM -> scatter-add) and anchors
on the reference's own call sites:
  det3d/models/backbones/scn.py:11-34   conv3x3/conv1x1 -> SubMConv3d(k, padding=1, bias, indice_key)
  det3d/models/backbones/scn.py:37-80   SparseBasicBlock (conv-bn-relu-conv-bn-add-relu, biased convs)
  det3d/models/backbones/scn.py:83-176  SpMiddleResNetFHD topology, dense().view(N, C*D, H, W)
It is additionally cross-checked against torch.nn.functional.conv3d on densified inputs
(tests/test_oracle_spconv.py), which pins the convolution semantics (cross-correlation,
weight layout [kD,kH,kW,Cin,Cout], out = (in + 2p - k)//s + 1) independently of spconv.

Canonical conventions: kernel offsets row-major (kz,ky,kx); SubM keeps
input order; strided conv outputs in ascending linear (b,z,y,x) order (spconv-1.x GPU behaviour);
pairs of one offset listed in ascending output row.
"""
import numpy as np
import torch
import torch.nn.functional as F


def seeded_state(module_or_shapes, seed):
    """Deterministic state_dict for golden fixtures that store a seed instead of megabytes of weights: every tensor is
    drawn from one seeded generator in sorted key order (weights ~ U(-b, b) with b = 1/sqrt(fan_in); BatchNorm
    weight in [0.5, 1.5), bias / running_mean ~ 0.1 N(0,1), running_var in [0.5, 1.5))."""
    shapes = module_or_shapes if isinstance(module_or_shapes, dict) else \
        {k: tuple(v.shape) for k, v in module_or_shapes.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[k] = torch.zeros(shp, dtype=torch.int64)
        elif leaf == "running_var" or (leaf == "weight" and len(shp) == 1):
            sd[k] = torch.rand(shp, generator=g) + 0.5
        elif leaf in ("running_mean", "bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        else:
            n = 1
            for d in shp:
                n *= d
            # conv weights: [kD,kH,kW,Cin,Cout] (spconv) or [Cout,Cin,kh,kw] (torch) -> fan_in = numel / Cout
            cout = shp[-1] if len(shp) == 5 else shp[0]
            b = 1.0 / max(n // max(cout, 1), 1) ** 0.5
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * b
    return sd


def _lin(coords, shape):
    c = coords.astype(np.int64)
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def _lookup(keys_sorted, perm, query):
    if len(keys_sorted) == 0:
        return np.full(query.shape, -1, np.int64)
    pos = np.minimum(np.searchsorted(keys_sorted, query), len(keys_sorted) - 1)
    return np.where(keys_sorted[pos] == query, perm[pos], -1)


def conv_out_shape(shape, ksize, stride, padding):
    return [(int(s) + 2 * p - k) // st + 1 for s, k, st, p in zip(shape, ksize, stride, padding)]


def neighbor_table(out_coords, in_coords, in_shape, ksize, stride, padding):
    """nbr [K, N_out]: input row feeding output row o through offset k (in = out*s - p + k), else -1."""
    in_keys = _lin(in_coords, in_shape)
    perm = np.argsort(in_keys, kind="stable")
    keys_sorted = in_keys[perm]
    K = ksize[0] * ksize[1] * ksize[2]
    nbr = np.full((K, len(out_coords)), -1, np.int32)
    oc = out_coords.astype(np.int64)
    k = 0
    for kz in range(ksize[0]):
        for ky in range(ksize[1]):
            for kx in range(ksize[2]):
                z = oc[:, 1] * stride[0] - padding[0] + kz
                y = oc[:, 2] * stride[1] - padding[1] + ky
                x = oc[:, 3] * stride[2] - padding[2] + kx
                ok = (z >= 0) & (z < in_shape[0]) & (y >= 0) & (y < in_shape[1]) & (x >= 0) & (x < in_shape[2])
                q = ((oc[:, 0] * in_shape[0] + z) * in_shape[1] + y) * in_shape[2] + x
                r = _lookup(keys_sorted, perm, np.where(ok, q, -1))
                nbr[k] = np.where(ok, r, -1)
                k += 1
    return nbr


def subm_rulebook(coords, shape, ksize):
    pad = [k // 2 for k in ksize]
    return neighbor_table(coords, coords, shape, ksize, [1, 1, 1], pad)


def conv_rulebook(coords, batch_size, shape, ksize, stride, padding):
    """-> (out_coords [N_out,4] ascending linear order, out_shape, nbr [K,N_out])."""
    out_shape = conv_out_shape(shape, ksize, stride, padding)
    c = coords.astype(np.int64)
    outs = []
    for kz in range(ksize[0]):
        for ky in range(ksize[1]):
            for kx in range(ksize[2]):
                tz, ty, tx = c[:, 1] + padding[0] - kz, c[:, 2] + padding[1] - ky, c[:, 3] + padding[2] - kx
                ok = (tz >= 0) & (ty >= 0) & (tx >= 0) & (tz % stride[0] == 0) & (ty % stride[1] == 0) & (tx % stride[2] == 0)
                oz, oy, ox = tz // stride[0], ty // stride[1], tx // stride[2]
                ok &= (oz < out_shape[0]) & (oy < out_shape[1]) & (ox < out_shape[2])
                outs.append((((c[:, 0] * out_shape[0] + oz) * out_shape[1] + oy) * out_shape[2] + ox)[ok])
    keys = np.unique(np.concatenate(outs)) if outs else np.zeros((0,), np.int64)
    oc = np.empty((len(keys), 4), np.int32)
    r = keys.copy()
    oc[:, 3] = r % out_shape[2]; r //= out_shape[2]
    oc[:, 2] = r % out_shape[1]; r //= out_shape[1]
    oc[:, 1] = r % out_shape[0]; r //= out_shape[0]
    oc[:, 0] = r
    return oc, out_shape, neighbor_table(oc, coords, shape, ksize, stride, padding)


def nbr_to_pairs(nbr):
    """spconv layout: list over k of (in_rows, out_rows), ascending output row."""
    pairs = []
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        pairs.append((nbr[k][o].astype(np.int64), o.astype(np.int64)))
    return pairs


def indice_conv(features, weight, nbr, n_out):
    """spconv-1.x indice_conv forward: per offset gather -> mm -> index_add_, fp32.
    features [N_in,Cin] torch fp32; weight [kD,kH,kW,Cin,Cout] or [K,Cin,Cout]."""
    w = weight.reshape(-1, weight.shape[-2], weight.shape[-1])
    out = torch.zeros((n_out, w.shape[-1]), dtype=features.dtype)
    for k, (i, o) in enumerate(nbr_to_pairs(nbr)):
        if len(i) == 0:
            continue
        out.index_add_(0, torch.from_numpy(o), features[torch.from_numpy(i)] @ w[k])
    return out


def bn_eval(x, sd, prefix, eps):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                        sd[prefix + "bias"], False, 0.0, eps)


def bn_train(momentum):
    """Training-mode BatchNorm1d (batch statistics; updates sd's running statistics in place, as nn.BatchNorm1d)."""
    def fn(x, sd, prefix, eps):
        return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                            sd[prefix + "bias"], True, momentum, eps)
    return fn


def backbone_forward(sd, voxel_features, coors, batch_size, input_shape, prefix="", eps=1e-3, return_stages=False,
                     bn_eval=bn_eval):
    """SpMiddleResNetFHD.forward (scn.py:148-176) in eval mode, state_dict `sd` in the reference key layout.
    voxel_features [M,5] fp32 torch, coors [M,4] int (b,z,y,x), input_shape = grid (x,y,z).
    Returns dense [B, C*D, H, W] (and per-stage dicts when return_stages).
    bn_eval=bn_train(momentum) switches every BatchNorm1d to training mode (torch-autograd differentiable: the
    oracle of the native backward pass).  Differentiable w.r.t. every tensor of `sd` that requires grad."""
    shape = [int(input_shape[2]) + 1, int(input_shape[1]), int(input_shape[0])]          # scn.py:151
    coords = np.asarray(coors, np.int32)
    x = voxel_features if voxel_features.dtype == torch.float64 else voxel_features.float()
    stages = {}
    rb_cache = {}

    def subm(x, coords, shape, wkey, key, bias=True):
        if key not in rb_cache:
            rb_cache[key] = subm_rulebook(coords, shape, [3, 3, 3])
        y = indice_conv(x, sd[prefix + wkey + ".weight"], rb_cache[key], len(coords))
        if bias and (prefix + wkey + ".bias") in sd:
            y = y + sd[prefix + wkey + ".bias"]
        return y

    def block(x, coords, shape, name, key):                                               # scn.py:64-80
        out = subm(x, coords, shape, name + ".conv1", key)
        out = F.relu(bn_eval(out, sd, prefix + name + ".bn1.", eps))
        out = subm(out, coords, shape, name + ".conv2", key)
        out = bn_eval(out, sd, prefix + name + ".bn2.", eps)
        return F.relu(out + x)

    def down(x, coords, shape, name, ksize, stride, padding):
        oc, oshape, nbr = conv_rulebook(coords, batch_size, shape, ksize, stride, padding)
        y = indice_conv(x, sd[prefix + name + ".0.weight"], nbr, len(oc))
        y = F.relu(bn_eval(y, sd, prefix + name + ".1.", eps))
        stages[name + "_rulebook"] = (oc, oshape, nbr)
        return y, oc, oshape

    x = subm(x, coords, shape, "conv_input.0", "res0", bias=False)
    x = F.relu(bn_eval(x, sd, prefix + "conv_input.1.", eps))
    x = block(x, coords, shape, "conv1.0", "res0")
    x = block(x, coords, shape, "conv1.1", "res0")
    stages["conv1"] = (x, coords, shape)
    for name, key, pad in (("conv2", "res1", [1, 1, 1]), ("conv3", "res2", [1, 1, 1]), ("conv4", "res3", [0, 1, 1])):
        x, coords, shape = down(x, coords, shape, name, [3, 3, 3], [2, 2, 2], pad)
        x = block(x, coords, shape, name + ".3", key)
        x = block(x, coords, shape, name + ".4", key)
        stages[name] = (x, coords, shape)
    x, coords, shape = down(x, coords, shape, "extra_conv", [3, 1, 1], [2, 1, 1], [0, 0, 0])
    stages["extra_conv"] = (x, coords, shape)
    stages["subm_rulebooks"] = rb_cache
    C = x.shape[1]
    dense = torch.zeros((batch_size, C, shape[0], shape[1], shape[2]), dtype=x.dtype)             # .dense()
    ci = torch.from_numpy(coords.astype(np.int64))
    dense[ci[:, 0], :, ci[:, 1], ci[:, 2], ci[:, 3]] = x
    dense = dense.view(batch_size, C * shape[0], shape[1], shape[2])                              # scn.py:167-168
    return (dense, stages) if return_stages else dense


def dense_conv3d_reference(features, coords, batch_size, shape, weight, ksize, stride, padding, out_coords):
    """Independent check: densify, F.conv3d (cross-correlation), sample at the active outputs."""
    Cin = features.shape[1]
    vol = torch.zeros((batch_size, Cin, *shape), dtype=torch.float64)
    ci = torch.from_numpy(np.asarray(coords, np.int64))
    vol[ci[:, 0], :, ci[:, 1], ci[:, 2], ci[:, 3]] = features.double()
    w = weight.reshape(*ksize, Cin, -1).permute(4, 3, 0, 1, 2).double()       # [Cout,Cin,kD,kH,kW]
    y = F.conv3d(vol, w, stride=stride, padding=padding)
    oc = torch.from_numpy(np.asarray(out_coords, np.int64))
    return y[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]].float()
