"""VoxelNet single-stage detector: reader -> sparse backbone -> BEV neck -> CenterHead.

Keeps det3d's detector API (det3d/models/detectors/base.py:10-70, single_stage.py:11-61, voxelnet.py:8-56):
`VoxelNet(reader, backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)`, `extract_feat(data)`,
`forward(example, return_loss=True)` over the collated `example` dict of
det3d/torchie/parallel/collate.py:163-245.

Additive GPU-resident entry: `forward_points(points, batch_offsets)` takes the raw concatenated
multi-sweep points [sum N, 5] and runs the fused voxelize+VFE kernel in place of the reference's CPU
voxelizer + H2D copy of padded voxels + VFE.
"""
import logging

import numpy as np
import torch
from torch import nn

from . import ops
from . import precision as _precision
from .registry import DETECTORS, build_backbone, build_head, build_neck, build_reader


class BaseDetector(nn.Module):
    """Detector base (base.py:10-70): feature flags + train/test dispatch."""

    def __init__(self):
        super().__init__()
        self.fp16_enabled = False

    @property
    def with_reader(self):
        return getattr(self, "reader", None) is not None

    @property
    def with_neck(self):
        return getattr(self, "neck", None) is not None

    @property
    def with_bbox(self):
        return getattr(self, "bbox_head", None) is not None

    def init_weights(self, pretrained=None):
        if pretrained is not None:
            logging.getLogger().info("load model from: %s", pretrained)


class SingleStageDetector(BaseDetector):
    def __init__(self, reader, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.reader = build_reader(reader)
        self.backbone = build_backbone(backbone)
        if neck is not None:
            self.neck = build_neck(neck)
        self.bbox_head = build_head(bbox_head)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.init_weights(pretrained=pretrained)

    def init_weights(self, pretrained=None):
        if pretrained is None:
            return
        sd = torch.load(pretrained, map_location="cpu")
        sd = sd.get("state_dict", sd)
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        self.load_state_dict(sd, strict=False)


@DETECTORS.register_module
class VoxelNet(SingleStageDetector):
    def __init__(self, reader, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__(reader, backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)
        self.voxel_cfg = None      # set by configure_voxelizer() for forward_points()
        self.precision = None      # None -> precision.default_precision() ("bf16x3": the tcgen05 arm)
        for cfg in (train_cfg, test_cfg):      # optional `precision` key in either config dict
            p = cfg.get("precision") if hasattr(cfg, "get") else None
            if p is not None:
                self.set_precision(p)

    def train(self, mode=True):
        # derived-weight caches (folded BatchNorm, layout copies, tensor-core packs) are keyed on tensor versions, which
        # a replayed CUDA-graph training step does not advance: drop them whenever the mode flips
        if mode != self.training:
            from .graphs import clear_weight_caches
            clear_weight_caches(self)
        return super().train(mode)

    def set_precision(self, p):
        """Arithmetic arm of every convolution of this detector: "bf16x3" (tensor cores, fp32-class results; the
        default), "fp32" (exact fp32 on CUDA cores) or "bf16" (single pass); None follows the process default."""
        _precision.set_module_precision(self, p)
        self.precision = p
        self.__dict__.pop("_trainer", None)
        return self

    def extract_feat(self, data):
        if "mean_features" in data:          # fused path: VFE already done by the voxelizer
            feats = data["mean_features"]
        else:
            feats = self.reader(data["features"], data["num_voxels"])
        # the fused pipeline hands activations between backbone, neck and head in the kernels' inter-layer format
        # (split bf16 hi/lo rows on the tensor-core arm); `fused=False` keeps every boundary a plain fp32 tensor
        fmt = "fp32"
        if data.get("fused", False):
            fmt = _precision.act_fmt(self.precision)
        x, voxel_feature = self.backbone(feats, data["coors"], data["batch_size"], data["input_shape"],
                                         n_dev=data.get("n_dev"), n_cap=data.get("n_cap"), out_fmt=fmt)
        if self.with_neck:
            x = self.neck(x, out_fmt=fmt)
        return x, voxel_feature

    def native_trainer(self, precision=None, attach_grads=True):
        """The NativeTrainer of this model (train-mode forward + hand-written backward), created on first use."""
        from . import train
        key = (_precision.resolve(precision or self.precision), attach_grads)
        tr = self.__dict__.get("_trainer")
        if tr is None or tr[0] != key:
            tr = (key, train.NativeTrainer(self, precision=key[0], attach_grads=attach_grads))
            self.__dict__["_trainer"] = tr
        return tr[1]

    def forward(self, example, return_loss=True, **kwargs):
        if self.training and return_loss:
            # training mode (trainer.py:317-344): native train-mode forward now, native backward when the caller runs
            # `sum(losses["loss"]).backward()`
            from . import train
            tr = self.native_trainer(attach_grads=False)
            losses = tr.forward(example)
            return train.bridged_losses(tr, losses) if torch.is_grad_enabled() else losses
        num_voxels = example["num_voxels"]
        data = dict(features=example["voxels"], num_voxels=example["num_points"], coors=example["coordinates"],
                    batch_size=len(num_voxels), input_shape=example["shape"][0])
        x, _ = self.extract_feat(data)
        bev_map = None
        if getattr(self.bbox_head, "bev_map", False):          # voxelnet.py:49
            bev_map = torch.stack(list(example["bev_map"]), dim=1).float()
        preds = self.bbox_head(x, bev_map)
        if return_loss:
            return self.bbox_head.loss(example, preds)
        return self.bbox_head.predict(example, preds, self.test_cfg)

    # ---- fused GPU-resident path ----------------------------------------------------------------------
    def configure_voxelizer(self, voxel_generator_cfg, training=False):
        """voxel_generator_cfg: the config's `voxel_generator` dict (range, voxel_size, max_points_in_voxel,
        max_voxel_num=[train, test]); picks the cap as preprocess.py:249-258 does."""
        mv = voxel_generator_cfg["max_voxel_num"]
        max_voxels = (mv[0] if training else mv[1]) if isinstance(mv, (list, tuple)) else mv
        self.voxel_cfg = dict(range=[float(v) for v in voxel_generator_cfg["range"]],
                              voxel_size=[float(v) for v in voxel_generator_cfg["voxel_size"]],
                              max_points=int(voxel_generator_cfg["max_points_in_voxel"]), max_voxels=int(max_voxels))
        return self

    def voxelize(self, points, batch_offsets):
        c = self.voxel_cfg
        if c is None:
            raise RuntimeError("call configure_voxelizer(cfg.voxel_generator) before forward_points()")
        nf = self.reader.num_input_features
        # rows zero-padded to a multiple of 8 channels so the stem conv can run on the tensor-core arm
        return ops.voxelize_vfe(points, batch_offsets, c["voxel_size"], c["range"], c["max_points"], c["max_voxels"],
                                num_feat=nf, feat_stride=(nf + 7) // 8 * 8)

    def forward_host(self, points_host, batch_offsets_host, out_host=None):
        """End-to-end entry for host-resident inputs: pinned `points_host [sum N, >=5]` fp32 and `batch_offsets_host
        [B+1]` int32 are copied to the module's device, the fused forward runs, and every head tensor of every task is
        copied back into one pinned host tensor `[B, H, W, sum c]` (returned with the per-head channel ranges).
        The H2D copy, the kernels and the D2H copy run on three streams chained by events, so consecutive calls
        pipeline (the upload of batch i+1 overlaps the kernels of batch i, the download of batch i overlaps the kernels
        of batch i+1).  Nothing synchronises with the host: wait on `self.host_result_ready` (a CUDA event) or
        `torch.cuda.synchronize()` before reading the returned tensors."""
        dev = next(self.parameters()).device
        cur = torch.cuda.current_stream(dev)
        streams = self.__dict__.setdefault("_io_streams", {})
        if dev not in streams:
            streams[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = streams[dev]
        with torch.cuda.stream(s_in):
            pts = points_host.to(dev, non_blocking=True)
            off = batch_offsets_host.to(dev, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(s_in)
        cur.wait_event(uploaded)
        pts.record_stream(cur)
        off.record_stream(cur)
        preds = self.forward_points(pts, off)
        outs, layout = [], []
        for t_id, p in enumerate(preds):
            for name, v in p.items():
                base = v.permute(0, 2, 3, 1)                    # channels-last view of the head's result buffer
                layout.append((t_id, name, v.shape[1]))
                outs.append(base)
        # all heads of a task are channel slices of one buffer: copy that buffer once per task
        bufs = []
        seen = set()
        for v in outs:
            key = v.untyped_storage().data_ptr()
            if key not in seen:
                seen.add(key)
                bufs.append(v.as_strided((v.shape[0], v.shape[1], v.shape[2], v.stride(2)),
                                         (v.stride(0), v.stride(1), v.stride(2), 1),
                                         v.storage_offset() - v.storage_offset() % v.stride(2)))
        if out_host is None:
            out_host = [torch.empty(b.shape, dtype=b.dtype).pin_memory() for b in bufs]
        computed = torch.cuda.Event()
        computed.record(cur)
        s_out.wait_event(computed)
        with torch.cuda.stream(s_out):
            for h, b in zip(out_host, bufs):
                h.copy_(b, non_blocking=True)
                b.record_stream(s_out)
            self.host_result_ready = torch.cuda.Event()
            self.host_result_ready.record(s_out)
        return out_host, layout

    def forward_points(self, points, batch_offsets, return_voxels=False):
        """points [sum N, >=5] fp32 CUDA, batch_offsets [B+1] int32 CUDA -> CenterHead predictions."""
        c = self.voxel_cfg
        vox = self.voxelize(points, batch_offsets)
        B = batch_offsets.numel() - 1
        grid = ops.grid_size_of(c["range"], c["voxel_size"])
        data = dict(mean_features=vox["features"], coors=vox["coords"], batch_size=B, input_shape=grid,
                    n_dev=vox["total"], n_cap=vox["coords"].shape[0], fused=True)
        x, _ = self.extract_feat(data)
        preds = self.bbox_head(x, None)
        return (preds, vox) if return_voxels else preds
