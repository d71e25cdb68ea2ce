"""CenterHead / SepHead (multi-timestep CenterPoint heatmap + regression heads) on the native kernels.

Constructor arguments, module tree and state_dict keys follow det3d/models/bbox_heads/center_head.py:81-152
(SepHead) and :231-390 (CenterHead): `shared_conv.{0,1}`, `tasks.{t}.{reg,height,dim,rot,vel,hm}.{0,1,3}`.
`vel` carries 2*timesteps channels (:354-356) -- this is the multi-timestep head of the forecast_n3 configs.

Variants (center_head.py:99-124,268-372): `dense` (one single-timestep SepHead per forecast timestep), `forecast_feature`
(per-task `forecast_conv` 2x(conv3x3+BN+ReLU) whose output `feats` is concatenated to the shared feature of the next
task), `bev_map` (`bev_conv` 6->16->32->C added to the shared feature), `two_stage` / `wide_head` / `reverse` /
`sparse` / `classify` (module tree + forward; their loss / predict branches are not on the hot path) -- so every
VoxelNet config under configs/centerpoint builds and runs forward.  `dcn_head` (DCN op) is out of scope.

forward() in the standard mode (all variant flags False, as in every BASELINE config) runs
  1 kernel : shared 3x3 conv 512->64 + bias + BN + ReLU
  1 kernel : the first 3x3 conv of all heads of a task fused into one 64 -> 64*n_heads GEMM (+bias+BN+ReLU)
  n kernels: per head 3x3 conv 64 -> c (+bias), each reading its channel slice of the fused activation and
             writing its channel slice of one [B,H,W,sum c] output tensor.
The torch modules are parameter containers; no torch arithmetic runs in forward().
"""
import copy
import logging

import numpy as np
import torch
from torch import nn

from . import ops
from . import precision as _precision
from .neck import act_fmt, as_nhwc_feat, conv_weight_kio, run_conv, to_nhwc
from .registry import HEADS
from .sparse import folded_epilogue


class SepHead(nn.Module):
    def __init__(self, in_channels, heads, head_conv=64, final_kernel=1, bn=False, init_bias=-2.19,
                 two_stage=False, forecast_feature=False, wide_head=False, **kwargs):
        super().__init__(**kwargs)
        self.heads = heads
        self.two_stage, self.forecast_feature, self.wide_head = two_stage, forecast_feature, wide_head
        self.precision = None     # None -> precision.default_precision()

        def cbr(cin, cout):
            return [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=True), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        if two_stage:                                         # center_head.py:101-116 (parameters only: the reference's
            if "vel" in heads and "rot" in heads:             # forward overwrites what these branches compute, :168)
                self.forecast_conv = nn.Sequential(*cbr(in_channels, head_conv))
            if "rvel" in heads and "rrot" in heads:
                self.reverse_conv = nn.Sequential(*cbr(in_channels, head_conv))
        if forecast_feature:                                  # :118-123
            self.forecast_conv = nn.Sequential(*(cbr(in_channels, head_conv) + cbr(head_conv, head_conv)))
        if wide_head:                                         # :125-126
            head_conv = in_channels
        for head in self.heads:
            classes, num_conv = self.heads[head]
            layers = []
            for _ in range(num_conv - 1):
                layers.append(nn.Conv2d(head_conv, head_conv, kernel_size=final_kernel, stride=1,
                                        padding=final_kernel // 2, bias=True))
                if bn:
                    layers.append(nn.BatchNorm2d(head_conv))
                layers.append(nn.ReLU())
            layers.append(nn.Conv2d(head_conv, classes, kernel_size=final_kernel, stride=1,
                                    padding=final_kernel // 2, bias=True))
            fc = nn.Sequential(*layers)
            if "hm" in head:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out", nonlinearity="relu")
                        nn.init.constant_(m.bias, 0)
            self.__setattr__(head, fc)

    # ---- fused execution plan -------------------------------------------------------------------------
    def _stage_groups(self):
        """[(conv, bn|None, relu)] per head, split into layers."""
        per_head = {}
        for head in self.heads:
            mods = list(getattr(self, head))
            steps, i = [], 0
            while i < len(mods):
                conv = mods[i]
                bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
                j = i + 1 + (bn is not None)
                relu = j < len(mods) and isinstance(mods[j], nn.ReLU)
                steps.append((conv, bn, relu))
                i = j + int(relu)
            per_head[head] = steps
        return per_head

    def _fused_first(self, groups):
        """Concatenate the first conv (+folded BN) of every head along Cout; cached on parameter versions."""
        firsts = [groups[h][0] for h in self.heads]
        key = tuple((c.weight.data_ptr(), c.weight._version, c.bias._version,
                     None if b is None else (b.weight._version, b.bias._version, b.running_mean._version,
                                             b.running_var._version, b.training, getattr(b, "_fd_stats_version", 0)))
                    for c, b, _ in firsts)
        cache = self.__dict__.setdefault("_fused_cache", {})
        if cache.get("k") != key:
            ws, scs, shs = [], [], []
            for conv, bnm, _ in firsts:
                ws.append(conv_weight_kio(conv))
                sc, sh = folded_epilogue(conv, bnm)
                scs.append(sc if sc is not None else torch.ones_like(sh))
                shs.append(sh)
            cache["v"] = (torch.cat(ws, dim=2).contiguous(), torch.cat(scs).contiguous(), torch.cat(shs).contiguous())
            cache["k"] = key
        return cache["v"]

    def _fused_last(self, groups):
        """Block-diagonal [K, n_heads*head_conv, sum c] weight + bias of the final conv of every head (one launch
        instead of n_heads tiny ones); cached on parameter versions."""
        lasts = [groups[h][1][0] for h in self.heads]
        key = tuple((c.weight.data_ptr(), c.weight._version, c.bias._version) for c in lasts)
        cache = self.__dict__.setdefault("_fused_last_cache", {})
        if cache.get("k") != key:
            ws = [conv_weight_kio(c) for c in lasts]
            K, hc = ws[0].shape[0], ws[0].shape[1]
            total = sum(w.shape[2] for w in ws)
            wbd = torch.zeros((K, hc * len(ws), total), dtype=torch.float32, device=ws[0].device)
            col = 0
            for i, w in enumerate(ws):
                wbd[:, i * hc:(i + 1) * hc, col:col + w.shape[2]] = w
                col += w.shape[2]
            bias = torch.cat([c.bias.detach().float() for c in lasts]).contiguous()
            cache["k"], cache["v"] = key, (wbd.contiguous(), bias)
        return cache["v"]

    def forward(self, x, precision=None, feats_out=None):
        """x: channels-last [B,H,W,C] tensor / ops.Feat (or logical NCHW tensor).  Returns {head: logical [B,c,H,W]}
        fp32 views of one channels-last [B,H,W,sum c] result tensor.  forecast_feature: `feats` (the output of
        forecast_conv, center_head.py:157-159) is added to the dict as an ops.Feat; `feats_out` lets the caller place
        it (CenterHead writes it straight into the next task's concatenated input)."""
        prec = _precision.resolve(precision or self.precision)
        fmt = act_fmt(prec)
        if isinstance(x, torch.Tensor) and x.dim() == 4 and x.stride(3) != 1:
            x = to_nhwc(x)
        x = ops.as_feat(x)
        if x.fmt != fmt and x.fmt != "fp32":
            x = as_nhwc_feat(x, fmt)
        feats = None
        if self.forecast_feature:
            fc = self.forecast_conv
            x = ops.as_feat(run_conv(x, fc[0], fc[1], True, precision=prec))
            x = feats = ops.as_feat(run_conv(x, fc[3], fc[4], True, precision=prec, out=feats_out, out_fmt=fmt))
        groups = self._stage_groups()
        names = list(self.heads)
        B, H, W = x.t.shape[0], x.t.shape[1], x.t.shape[2]
        total_c = sum(self.heads[h][0] for h in names)
        out = ops.Feat(torch.empty((B, H, W, total_c), dtype=torch.float32, device=x.t.device), "fp32")
        fuse = all(len(groups[h]) == 2 for h in names) and len({groups[h][0][0].kernel_size for h in names}) == 1
        ret, col = {}, 0
        if fuse:
            w, sc, sh = self._fused_first(groups)
            c0 = groups[names[0]][0][0]
            mid = ops.as_feat(ops.conv2d_nhwc(x, w, c0.kernel_size, c0.stride, c0.padding, sc, sh, True, precision=prec,
                                              out_fmt=fmt))
            lasts = [groups[h][1] for h in names]
            same = (len({(c.kernel_size, c.stride, c.padding) for c, _, _ in lasts}) == 1
                    and all(b is None and not r for _, b, r in lasts))
            if same:
                # final convs of all heads as ONE block-diagonal conv over the fused activation (n_heads*64 -> sum c)
                wbd, bias = self._fused_last(groups)
                cl = lasts[0][0]
                ops.conv2d_nhwc(mid, wbd, cl.kernel_size, cl.stride, cl.padding, None, bias, False, out=out, precision=prec)
                for h in names:
                    c = self.heads[h][0]
                    ret[h] = out.t[..., col:col + c].permute(0, 3, 1, 2)
                    col += c
            else:
                hc = c0.out_channels
                for i, h in enumerate(names):
                    conv, bnm, relu = groups[h][1]
                    c = conv.out_channels
                    run_conv(mid.slice(i * hc, hc), conv, bnm, relu, out=out.slice(col, c), precision=prec)
                    ret[h] = out.t[..., col:col + c].permute(0, 3, 1, 2)
                    col += c
        else:
            for h in names:
                y = x
                steps = groups[h]
                for si, (conv, bnm, relu) in enumerate(steps):
                    last = si == len(steps) - 1
                    c = conv.out_channels
                    y = ops.as_feat(run_conv(y, conv, bnm, relu, out=out.slice(col, c) if last else None, precision=prec))
                ret[h] = out.t[..., col:col + c].permute(0, 3, 1, 2)
                col += c
        if feats is not None:
            ret["feats"] = feats
        return ret


@HEADS.register_module
class CenterHead(nn.Module):
    def __init__(self, in_channels=[128, ], tasks=[], dataset="nuscenes", weight=0.25, code_weights=[],
                 common_heads=dict(), logger=None, init_bias=-2.19, share_conv_channel=64, num_hm_conv=2,
                 dcn_head=False, timesteps=1, two_stage=False, reverse=False, sparse=False, dense=False,
                 bev_map=False, forecast_feature=False, classify=True, wide_head=False):
        super().__init__()
        self.two_stage, self.reverse, self.sparse, self.dense = two_stage, reverse, sparse, dense
        self.bev_map, self.forecast_feature, self.classify, self.wide_head = bev_map, forecast_feature, classify, wide_head
        self.target_timesteps = 7
        self.standard = not (reverse or sparse or dense or classify or wide_head)
        if dcn_head:
            raise NotImplementedError("CenterHead: dcn_head needs the deformable-convolution op (det3d/ops/dcn), which is "
                                      "outside the hot-path scope (SURVEY.md section 2, OUT)")
        self.precision = None     # None -> precision.default_precision()
        num_classes = [len(t["class_names"]) for t in tasks]
        self.class_names = [t["class_names"] for t in tasks]
        self.code_weights = code_weights
        self.box_n_dim = 7
        if all(k in common_heads for k in ("vel", "rvel", "rot", "rrot")):
            self.box_n_dim = 13
            self.code_weights_forecast = list(np.array(code_weights) * np.array([0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0]))
        elif "vel" in common_heads and "rot" in common_heads:
            self.box_n_dim = 9
            self.code_weights_forecast = list(np.array(code_weights) * np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 0]))
        self.weight = weight
        self.dataset = dataset
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.use_direction_classifier = False
        self.timesteps = timesteps
        self.logger = logger or logging.getLogger("CenterHead")
        self.logger.info("num_classes: %s", num_classes)

        if sparse:                                            # center_head.py:320-334
            self.num_classes = 2 * [1]
        if dense:
            self.num_classes = self.timesteps * [1]
        if classify:
            self.num_classes = self.timesteps * [3]
        if wide_head:
            self.num_classes = [7]
            share_conv_channel = 512
        if bev_map:                                           # :336-341
            def cbr(cin, cout):
                return [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=True), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
            self.bev_conv = nn.Sequential(*(cbr(6, 16) + cbr(16, 32) + cbr(32, share_conv_channel)))
        self.shared_conv = nn.Sequential(
            nn.Conv2d(in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            nn.BatchNorm2d(share_conv_channel), nn.ReLU(inplace=True))
        self.share_conv_channel = share_conv_channel
        self.tasks = nn.ModuleList()
        for i, num_cls in enumerate(self.num_classes):
            heads = copy.deepcopy(dict(common_heads))
            for head in heads:
                if not dense and not classify and not wide_head and head in ("vel", "rvel"):
                    heads[head] = (self.timesteps * heads[head][0], heads[head][1])      # center_head.py:354-356
            heads.update(dict(hm=(num_cls, num_hm_conv)))
            cin = 2 * share_conv_channel if (i != 0 and forecast_feature) else share_conv_channel      # :361-368
            self.tasks.append(SepHead(cin, heads, bn=True, init_bias=init_bias, final_kernel=3, two_stage=two_stage,
                                      forecast_feature=forecast_feature, wide_head=wide_head))
        self.logger.info("Finish CenterHead Initialization")

    def forward(self, x, bev_map=None, *kwargs):
        """x logical [B,512,H,W] -> list over tasks of {head: logical [B,c,H,W]} (center_head.py:375-390)."""
        prec = _precision.resolve(self.precision)
        fmt = act_fmt(prec)
        x = as_nhwc_feat(x, fmt)
        if self.bev_map:
            # x = shared_conv(x) + bev_conv(bev_map) (:380-381): both branches end in a ReLU, so the sum is a separate
            # row add over fp32 rows
            from .train_ops import add_rows_
            if bev_map is None:
                raise RuntimeError("CenterHead(bev_map=True).forward needs the bev_map tensor [B,6,H,W]")
            x = ops.as_feat(run_conv(x, self.shared_conv[0], self.shared_conv[1], True, precision=prec, out_fmt="fp32"))
            b = ops.Feat(to_nhwc(bev_map.float()).contiguous())
            bc = self.bev_conv
            b = ops.as_feat(run_conv(b, bc[0], bc[1], True, precision=prec))
            b = ops.as_feat(run_conv(b, bc[3], bc[4], True, precision=prec))
            b = ops.as_feat(run_conv(b, bc[6], bc[7], True, precision=prec, out_fmt="fp32"))
            add_rows_(x.t, b.t)
        else:
            x = ops.as_feat(run_conv(x, self.shared_conv[0], self.shared_conv[1], True, precision=prec))
        if not self.forecast_feature:
            return [task(x, precision=prec) for task in self.tasks]
        # forecast_feature (:383-388): task i > 0 reads cat([x, feats of task i-1]).  Every task's forecast_conv writes
        # its `feats` straight into the upper channel half of the next task's input buffer; x is copied into the lower.
        rets = []
        Cs = self.share_conv_channel
        B, H, W = x.t.shape[0], x.t.shape[1], x.t.shape[2]
        cur = x
        for i, task in enumerate(self.tasks):
            nxt = None
            if i + 1 < len(self.tasks):
                nxt = ops.Feat(torch.empty((B, H, W, 2 * Cs), dtype=torch.float32, device=x.t.device), fmt)
                ops.copy_rows(x, nxt.slice(0, Cs))
            r = task(cur, precision=prec, feats_out=None if nxt is None else nxt.slice(Cs, Cs))
            r["feats"] = r["feats"].to_fp32().permute(0, 3, 1, 2)          # the reference returns it as a tensor
            rets.append(r)
            cur = nxt
        return rets

    def loss(self, example, preds_dicts, **kwargs):
        from .loss import center_head_loss
        return center_head_loss(self, example, preds_dicts)

    @torch.no_grad()
    def predict(self, example, preds_dicts, test_cfg, **kwargs):
        """decode + score/range masks + rotated NMS per forecast timestep (center_head.py:541-747) -> list over samples
        of {box3d_lidar [n,9], scores [n], label_preds [n], metadata}."""
        from .predict import center_head_predict
        return center_head_predict(self, example, preds_dicts, test_cfg)
