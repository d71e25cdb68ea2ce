#!/usr/bin/env python
"""bench.py -- forward scenes/sec of the LiDAR hot path on synthetic 10-sweep nuScenes-shaped scenes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (fused voxelize+VFE -> SpMiddleResNetFHD -> RPN -> CenterHead) over one
batch of synthetic scenes (BASELINE.json configs[1]: forecast_n0 CenterPoint-VoxelNet car, ~300k-point
10-sweep scene, forward only).  One process per GPU; under torchrun each rank runs independent scenes
(the path shards by scene, no data-path collective) and rank 0 prints ONE JSON line.

`--impl reference` times the reference's CPU path restated in oracle/ (the reference is Python: numba
voxelizer, external spconv, torch neck/head -- nothing compiles into oracle/_ref) with all host threads, one FULL
scene per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene  # noqa: E402

METRIC = "scenes/sec fwd (300k-pt, 10-sweep synth)"
N_TARGET = 360_000           # synth_scene(360000) -> 305,677 points, the 160k-voxel val cap is hit
VOXEL_CFG = dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10, max_voxel_num=[120000, 160000])
HEADS = ["reg", "height", "dim", "rot", "vel", "hm"]
TEST_CFG = dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_per_img=500,
                nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=83,
                         nms_iou_threshold=0.2),
                score_threshold=0.1, pc_range=[-54, -54], out_size_factor=8, voxel_size=[0.075, 0.075])   # configs/...n0...:88-103


def model_cfg(timesteps=1):
    return dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(type="CenterHead", in_channels=512, tasks=[dict(num_class=1, class_names=["car"])],
                       dataset="nuscenes", weight=0.25, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                       share_conv_channel=64, dcn_head=False, timesteps=timesteps, two_stage=False, reverse=False,
                       sparse=False, dense=False, bev_map=False, forecast_feature=False, classify=False,
                       wide_head=False))


def build_model(seed=0):
    import futuredet_b200 as fb
    torch.manual_seed(seed)
    m = fb.build_detector(model_cfg()).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():                       # non-trivial eval-mode BN
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    return m


def peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json when present (HBM copy GB/s, dense bf16 TFLOP/s burst
    and sustained), else the fallback of /opt/skills/guides/B200_PROFILING.md; `src` says which."""
    fb = dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        return fb
    try:
        d = json.load(open(p))

        def pick(*names):
            for n in names:
                v = d.get(n)
                if isinstance(v, dict):
                    v = v.get("value")
                if isinstance(v, (int, float)) and v > 0:
                    return float(v)
            return None
        hbm = pick("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs", "hbm")
        burst = pick("bf16_tflops", "bf16_tflops_burst", "bf16_dense_tflops")
        sust = pick("bf16_tflops_sustained", "bf16_sustained_tflops") or burst
        if hbm and burst:
            return dict(hbm=hbm, tf_burst=burst, tf_sustained=sust, src="measured")
    except (OSError, ValueError):
        pass
    return fb


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), power_w_max=max(float(r[3]) for r in rows),
                    samples=len(rows), reasons=sorted(reasons))


def dist_setup(n_gpus, init):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and init:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    return rank, world, local


def barrier_max(value, world, dev):
    from futuredet_b200 import shard
    return shard.reduce_throughput(value, 0, device=dev)[1] if world > 1 else value


# ----------------------------------------------------------------------------------------- CPU arms
def oracle_forward(sd, scenes):
    """Reference CPU path (restated): voxelize+VFE (C) -> spconv backbone -> RPN -> CenterHead, all host threads."""
    from oracle import dense_ref as D, spconv_ref as S, voxelizer as V
    vox = V.voxelize_batch_c(scenes, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    bev = S.backbone_forward({k[9:]: v for k, v in sd.items() if k.startswith("backbone.")},
                             torch.from_numpy(vox["features"]), vox["coords"], len(scenes), [1440, 1440, 40])
    feat = D.rpn_forward({k[5:]: v for k, v in sd.items() if k.startswith("neck.")}, bev, [5, 5], [1, 2], [1, 2])
    return D.center_head_forward({k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}, feat, [HEADS])


def cpu_time_scene(sd, n_target, seed=0):
    scene = synth_scene(n_target, seed=seed)
    t = time.perf_counter()
    with torch.no_grad():
        oracle_forward(sd, [scene])
    return time.perf_counter() - t, len(scene)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms must use the box's cores regardless."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return n


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the box's host cores, rank 0 only.  Every step is ONE FULL
    305k-point scene (never sub-sampled or scaled); if the requested K steps would not fit the time budget, fewer
    steps are timed and `steps` says how many.  The reference itself cannot travel to the GPU box (/root/reference
    is absent there, its sources may not be copied) and it is Python over numba / external spconv / torch: the arm is
    the oracle port (kind "port"): C restatement of the numba voxelizer, restated spconv-CPU indice_conv, and the same
    torch.nn.functional conv2d / batch_norm calls the reference RPN / CenterHead modules make."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    model = build_model()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    n_target = int(os.environ.get("FD_REF_N_TARGET", N_TARGET))          # test hook: tiny scene
    budget_s = float(os.environ.get("FD_REF_BUDGET_S", 200.0))
    t_first, n_pts = cpu_time_scene(sd, n_target, seed=1000)             # warm-up step 0 doubles as the probe
    warm = max(0, min(args.warmup - 1, int(budget_s * 0.25 / max(t_first, 1e-6))))
    for w in range(warm):
        cpu_time_scene(sd, n_target, seed=1001 + w)
    steps = max(1, min(args.steps, int(budget_s * 0.75 / max(t_first, 1e-6))))
    t_total = 0.0
    for s_ in range(steps):
        t, n_pts = cpu_time_scene(sd, n_target, seed=s_)
        t_total += t
    value = steps / t_total
    sample = ("%d timed step(s) (of %d requested; %d warm-up) of one FULL synth_scene(%d) (%d pts) each through the oracle "
              "port: C voxelizer (1 thread) + restated spconv-CPU + torch RPN/CenterHead (%d threads)"
              % (steps, args.steps, warm + 1, n_target, n_pts, cores))
    line = dict(metric=METRIC, value=value, unit="scenes/s", n_gpus=args.gpus, steps=steps, warmup=warm + 1,
                ms_per_step=1e3 * t_total / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="forecast_n0 CenterPoint-VoxelNet car, fwd-only (BASELINE configs[1])", batch_per_gpu=1,
                            points_per_scene=n_pts, sweeps=10, max_voxels=160000, host_cpus=os.cpu_count(),
                            note="CPU reference path; one full scene per step"),
                cpu_baseline=dict(value=value, unit="scenes/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="scenes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- GPU arm
def conv_profile(model, pts, off):
    """Per-launch CUDA-event timing of the dominant kernel family (implicit-GEMM conv) on the launching stream."""
    from futuredet_b200 import ops
    ops.PROFILE = []
    model.forward_points(pts, off)
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    tot_flop, tot_ms, n = 0.0, 0.0, 0
    by_kind = {}
    for r in recs:
        ms = r["start"].elapsed_time(r["end"])
        flop = r["flops"]() if callable(r["flops"]) else r["flops"]
        tot_flop += flop
        tot_ms += ms
        n += r["launches"]
        k = by_kind.setdefault(r["kind"], [0.0, 0.0])
        k[0] += flop
        k[1] += ms
    return dict(flop=tot_flop, ms=tot_ms, launches=n,
                by_kind={k: dict(gflop=v[0] / 1e9, ms=v[1], tflops=v[0] / max(v[1], 1e-9) / 1e9) for k, v in by_kind.items()})


def voxelize_spconv_ms(model, pts, off, reps=5):
    """The second half of BASELINE's metric: voxelize + VFE + sparse backbone (rulebooks included), ms per scene, CUDA
    events around exactly that prefix of the forward at the benched batch size."""
    from futuredet_b200 import ops, precision as P
    c = model.voxel_cfg
    B = off.numel() - 1
    grid = ops.grid_size_of(c["range"], c["voxel_size"])
    fmt = P.act_fmt(model.precision)

    def run():
        vox = model.voxelize(pts, off)
        return model.backbone(vox["features"], vox["coords"], B, grid, n_dev=vox["total"], n_cap=vox["coords"].shape[0],
                              out_fmt=fmt)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / B


def voxelize_roofline(model, dev, n_scenes=32):
    """HBM roofline of the fused voxelize+VFE family on one batched launch of `n_scenes` scenes (a single 12.5 MB
    scene is below launch latency, SURVEY.md hard part 1): algorithmic bytes 20*N + 40*M over the CUDA-event time."""
    scenes = [synth_scene(N_TARGET, seed=5000 + i) for i in range(4)]
    pts = torch.from_numpy(np.concatenate([scenes[i % 4] for i in range(n_scenes)])).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(scenes[i % 4]) for i in range(n_scenes)])], dtype=torch.int32, device=dev)
    for _ in range(3):
        vox = model.voxelize(pts, off)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        vox = model.voxelize(pts, off)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    m = int(vox["total"].item())
    nbytes = 20.0 * pts.shape[0] + 40.0 * m
    pk = peaks()
    gbs = nbytes / (ms * 1e-3) / 1e9
    return dict(bound="hbm", kernel="fused voxelize+VFE family (7 launches), %d scenes per launch" % n_scenes,
                achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], traffic=None,
                algorithmic_bytes=nbytes, ms=ms, points=int(pts.shape[0]), voxels=m, peak_source=pk["src"] + " HBM copy")


def cudnn_bar(model, dev, B):
    """the bar for the dense neck + head is cuDNN-via-torch on the same box.  The model's own torch modules
    (nn.Conv2d / BatchNorm2d / ReLU parameter containers of RPN and CenterHead, identical weights) are run through torch's
    module forward (rpn.py:150-159, center_head.py:375-390) on a [B,256,180,180] BEV map, channels_last, TF32 off / on and
    bf16 autocast, against the native dense kernels on the same input (CUDA events, 3 warm-up + 5 timed)."""
    import torch.nn.functional as F  # noqa: F401
    neck, head = model.neck, model.bbox_head
    x = torch.randn((B, 256, 180, 180), device=dev).contiguous(memory_format=torch.channels_last)

    def torch_fwd(inp):
        ups, h = [], inp
        for i, blk in enumerate(neck.blocks):
            h = blk(h)
            j = i - neck._upsample_start_idx
            if j >= 0:
                ups.append(neck.deblocks[j](h))
        f = torch.cat(ups, dim=1) if ups else h
        s_ = head.shared_conv(f)
        return [{hn: getattr(t, hn)(s_) for hn in t.heads} for t in head.tasks]

    def native_fwd(inp):
        return head(neck(inp, out_fmt="split" if model.precision != "fp32" else "fp32"))

    def time_it(fn, inp):
        for _ in range(3):
            fn(inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(inp)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5

    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("cudnn_fp32_ms", False), ("cudnn_tf32_ms", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            out[name] = time_it(torch_fwd, x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out["cudnn_bf16_autocast_ms"] = time_it(torch_fwd, x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    out["native_ms"] = time_it(native_fwd, x)
    out["batch"] = B
    out["what"] = ("RPN + CenterHead (n0) on a [B,256,180,180] map: torch module forward (cuDNN, channels_last, "
                   "cudnn.benchmark) vs the native dense kernels (%s); fp32-class accuracy: native and cudnn_fp32; "
                   "TF32 / bf16 autocast do not hold the 1e-3 contract over the whole network" % model.precision)
    return out


def stress_bench(args, dev):
    """BASELINE configs[4]: mixed car + pedestrian heads (two SepHeads, center_head.py:351-372), 7-timestep forecast_n3
    heads, 500k-point dense scenes, forward only; 2 scenes per step per GPU, same timing rules as the main metric."""
    import futuredet_b200 as fb
    torch.manual_seed(3)
    cfg = model_cfg(timesteps=7)
    cfg["bbox_head"]["tasks"] = [dict(num_class=1, class_names=["car"]), dict(num_class=1, class_names=["pedestrian"])]
    m = fb.build_detector(cfg).eval().set_precision(args.precision).to(dev).configure_voxelizer(VOXEL_CFG)
    scenes = [synth_scene(590000, seed=70 + i) for i in range(2)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(sc) for sc in scenes])], dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m.forward_points(pts, off)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        e0.record()
        for _ in range(reps):
            flush.zero_()
            m.forward_points(pts, off)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return dict(workload="pedestrian_forecast_n3-style mixed car+ped heads, 2 x %d-pt scenes per step, fwd-only "
                         "(BASELINE configs[4])" % (pts.shape[0] // 2), ms_per_step=ms, scenes_per_s_per_gpu=2 / (ms / 1e3))


def train_bench(args, rank, world, dev):
    """BASELINE configs[2]/[3]: forecast_n3 (7-timestep heads) forward + backward (+ AdamW step).  configs[2] (one GPU):
    a single 305k-point sample per step; configs[3] (N GPUs): 4 samples per GPU per step (batch 32 at 8 GPUs,
    det3d/torchie/apis/train.py:311-317, build_loader.py:35-36) with the parameter gradients all-reduced over NCCL.
    The whole step (forward, backward and the bucket all-reduces) is captured into ONE CUDA graph per rank -- the NCCL
    kernels are graph nodes, so a replayed step runs no Python (the eager multi-rank step was host-bound: 8 launch-heavy
    ranks on 32 vCPUs took 70 ms at N=8 in round 1).  Auxiliary measurement, reported under "train" in the JSON line."""
    import futuredet_b200 as fb
    from futuredet_b200 import lib, shard, train
    from futuredet_b200.synth import synth_targets
    torch.manual_seed(0)
    old_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    cores = shard.pin_rank_to_cores(dev.index, world) if world > 1 else None
    m = fb.build_detector(model_cfg(timesteps=7)).to(dev).train()
    m.configure_voxelizer(VOXEL_CFG, training=True)
    shard.broadcast_parameters(m)
    tr = train.NativeTrainer(m, precision=args.train_precision)
    use_graph = not args.train_eager
    sync = shard.GradSync(tr.grads, inline=use_graph)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    B = args.train_batch if args.train_batch > 0 else (4 if world > 1 else 1)
    scenes = [synth_scene(N_TARGET, seed=shard.scene_seed(rank, 7, b, B)) for b in range(B)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(sc) for sc in scenes])], dtype=torch.int32, device=dev)
    ex = synth_targets(B, 180, 180, 7, seed=rank)
    ex = {k: [[t.to(dev) for t in ts] for ts in v] for k, v in ex.items()}
    losses = None
    graphed, graph_note = None, ""
    if use_graph:
        from futuredet_b200 import graphs
        try:
            graphed = graphs.GraphedTrainStep(tr, max_points=pts.shape[0], batch_size=B).capture(ex, pts, off)
        except Exception as e:          # e.g. an NCCL build that refuses capture: keep measuring, say so
            graphed, graph_note = None, "graph capture failed (%s: %s); eager step" % (type(e).__name__, str(e)[:120])
            sync.inline = False

    def step():
        nonlocal losses
        if graphed is not None:
            losses = graphed(ex, pts, off)
        else:
            losses = tr.step(ex, points=pts, batch_offsets=off)
            sync.finish()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    n0 = lib.launch_count()
    sync.bytes_reduced = 0
    sampler = ClockSampler(dev.index) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.train_steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = barrier_max(e0.elapsed_time(e1), world, dev) / args.train_steps
    loss = float(sum(losses["loss"]))
    grad_bytes = sum(b[0].numel() * 4 for b in tr.grads.buckets)
    if cores is not None and old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)
    return dict(workload="forecast_n3 (7-timestep heads) car, fwd+bwd+AdamW, %d x 305k-pt scene(s) per GPU per step "
                         "(BASELINE configs[%d]: global batch %d)" % (B, 2 if world == 1 else 3, B * world),
                ms_per_step=ms, samples_per_s=B * world / (ms / 1e3), batch_per_gpu=B, global_batch=B * world,
                precision=args.train_precision, steps=args.train_steps,
                launch_mode=("cuda graph replay (NCCL all-reduce nodes inside the graph)" if world > 1 else
                             "cuda graph replay") if graphed is not None else "eager", note=graph_note,
                gpu_launches_per_step=(lib.launch_count() - n0) // args.train_steps, loss=loss,
                params=sum(p.numel() for p in m.parameters()), clocks=clocks, cpu_affinity=cores,
                allreduce_bytes_per_step=grad_bytes if world > 1 else 0,
                exchange="bucketed all-reduce(avg) of the parameter gradients, %s" %
                         ("NCCL over NVLink" if world > 1 else "single rank: none"))


def run_gpu(args, rank, world, local):
    from futuredet_b200 import lib, shard
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib.load()
    model = build_model().set_precision(args.precision)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()} if rank == 0 else None
    model.to(dev).configure_voxelizer(VOXEL_CFG)
    B = args.batch
    n_pool = 4
    pool_host = []
    for i in range(n_pool):                     # distinct scenes per rank and per pool slot
        scenes = [synth_scene(N_TARGET, seed=shard.scene_seed(rank, i, b, B)) for b in range(B)]
        pts = torch.from_numpy(np.concatenate(scenes)).pin_memory()
        off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32).pin_memory()
        pool_host.append((pts, off))
    pool_dev = [(p.to(dev), o.to(dev)) for p, o in pool_host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_resident(i):
        flush.zero_()
        p, o = pool_dev[i % n_pool]
        return model.forward_points(p, o)

    out_host = None

    def step_e2e(i):
        # the public host-buffer API: pinned points -> H2D -> fused forward -> every head tensor D2H into pinned memory
        nonlocal out_host
        flush.zero_()
        p, o = pool_host[i % n_pool]
        out_host, _ = model.forward_host(p, o, out_host)
        return out_host

    det_bytes = [0]

    def step_detect(i):
        # inference as the reference runs it (voxelnet.py:51-56): forward + CenterHead.predict; only the detections
        # (boxes, scores, labels of the kept objects) travel back to the host
        flush.zero_()
        p, o = pool_host[i % n_pool]
        preds = model.forward_points(p.to(dev, non_blocking=True), o.to(dev, non_blocking=True))
        dets = model.bbox_head.predict({}, preds, TEST_CFG)
        host = [(d["box3d_lidar"].cpu(), d["scores"].cpu(), d["label_preds"].cpu()) for d in dets]
        det_bytes[0] = sum(t.numel() * t.element_size() for h in host for t in h)
        return host

    def timed(step_fn):
        for i in range(args.warmup):
            step_fn(i)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        n0 = lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_fn(args.warmup + i)
        e1.record()
        torch.cuda.synchronize()
        launches = lib.launch_count() - n0
        ms = e0.elapsed_time(e1)
        if world > 1:
            torch.distributed.barrier()
        return barrier_max(ms, world, dev), launches

    with torch.no_grad():
        sampler = ClockSampler(local) if rank == 0 else None
        ms_res, launches = timed(step_resident)
        clocks = sampler.stop() if sampler else None
        ms_e2e, _ = timed(step_e2e)
        ms_det, _ = timed(step_detect)
        prof = conv_profile(model, *pool_dev[0]) if rank == 0 else None
        vox_roof = voxelize_roofline(model, dev) if rank == 0 else None
        vs_ms = voxelize_spconv_ms(model, *pool_dev[0]) if rank == 0 else None
        # the reference's own samples_per_gpu = 1: latency-bound single-scene steps, same timing rules
        one = [(p[:int(o[1])].contiguous(), o[:2].contiguous()) for p, o in pool_dev]
        ms_b1, _ = timed(lambda i: (flush.zero_(), model.forward_points(*one[i % n_pool]))[1])
        # the same as ONE CUDA-graph launch per scene (graphs.GraphedForward: static NaN-padded point buffer)
        from futuredet_b200 import graphs
        gf = graphs.GraphedForward(model, max_points=max(int(p.shape[0]) for p, _ in one) + 1024, batch_size=1)
        gf(*one[0])
        ms_b1g, _ = timed(lambda i: (flush.zero_(), gf(*one[i % n_pool]))[1])
        del gf
        bar = cudnn_bar(model, dev, min(B, 4)) if rank == 0 and not args.no_cudnn_bar else None
        stress = stress_bench(args, dev) if rank == 0 else None
    train_info = None
    if args.train_steps > 0:
        pool_dev.clear()
        flush = None
        torch.cuda.empty_cache()
        train_info = train_bench(args, rank, world, dev)
    if rank != 0:
        return
    scenes = B * world * args.steps
    value = scenes / (ms_res / 1e3)
    e2e_value = scenes / (ms_e2e / 1e3)
    pk = peaks()
    h2d = int(pool_host[0][0].numel() * 4 + pool_host[0][1].numel() * 4)
    d2h = int(sum(h.numel() for h in out_host) * 4)
    achieved = prof["flop"] / max(prof["ms"], 1e-9) / 1e9      # TFLOP/s
    # dram bytes of the conv family from an ncu capture: only reported when a capture of THIS batch size is committed
    traffic = None
    tpath = os.path.join(REPO, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("batch_per_gpu") == B and tj.get("precision") == args.precision:
            traffic = tj.get("conv_family_dram_bytes_per_step")
    # L2 -> SM side of the same family: the gather roofline is measured live (a kernel that only issues the producers'
    # access pattern, csrc/probe.cu); the family's bytes through L2 come from an ncu capture of THIS batch size
    l2 = None
    if rank == 0:
        import ctypes as _C
        from futuredet_b200 import lib as _lib
        _L = _lib.load()
        _L.fd_debug_l2_gather_probe.argtypes = [_C.c_int, _C.POINTER(_C.c_double), _C.c_void_p]
        bps = _C.c_double(0.0)
        if _L.fd_debug_l2_gather_probe(212740, _C.byref(bps), None) == 0:
            l2_bytes = None
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("batch_per_gpu") == B and tj.get("precision") == args.precision:
                    l2_bytes = tj.get("conv_family_l2_bytes_per_step")
            l2 = dict(gather_peak_tbs=bps.value / 1e12,
                      peak_what="measured in this run: 16-byte cp.async gathers of 2 x 128-byte row pieces at neighbour-like "
                                "indices of a 54 MB buffer into a 4 x 32 KB shared-memory ring, nothing else running",
                      bytes_per_step=l2_bytes,
                      bytes_what="lts__t_bytes of the family per step, ncu capture of this batch size (null otherwise)",
                      achieved_tbs=(l2_bytes / (prof["ms"] * 1e-3) / 1e12) if l2_bytes else None,
                      frac=(l2_bytes / (prof["ms"] * 1e-3) / bps.value) if l2_bytes else None)
    roofline = dict(bound="tensor", kernel="gather->implicit-GEMM conv family (%s), %d launches/step" %
                    (args.precision, prof["launches"]), achieved=achieved, peak=pk["tf_sustained"], unit="TFLOP/s",
                    frac=achieved / pk["tf_sustained"], traffic=traffic,
                    traffic_what="dram bytes of the family per step (all its launches), ncu capture of this batch size",
                    mma_tflops=3.0 * achieved if args.precision == "bf16x3" else achieved,
                    note="bf16x3 executes 3 tensor-core MMAs per algorithmic product: frac <= 1/3 by construction",
                    peak_source=pk["src"] + " bf16 dense, sustained",
                    algorithmic_gflop_per_step=prof["flop"] / 1e9, kernel_ms_per_step=prof["ms"], by_kind=prof["by_kind"],
                    l2=l2)
    cpu_baseline = None
    if world == 1:                                                       # reported at N = 1 only (the contract)
        cores = use_all_host_threads()
        cpu_time_scene(sd_cpu, N_TARGET // 8, seed=1)                   # warm the thread pool / allocator
        cpu_t, cpu_n = cpu_time_scene(sd_cpu, N_TARGET, seed=0)
        cpu_baseline = dict(value=1.0 / cpu_t, unit="scenes/s", cores=cores, kind="port",
                            sample="1 FULL synth_scene(%d) (%d pts) through the oracle port (C voxelizer + restated "
                                   "spconv-CPU + torch RPN/CenterHead), %.1f s" % (N_TARGET, cpu_n, cpu_t),
                            host_cpus=os.cpu_count())
    line = dict(metric=METRIC, value=value, unit="scenes/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"fp32": "f32", "bf16x3": "bf16x3 (3-term split, fp32 accumulate)", "bf16": "bf16"}[args.precision],
                data="synthetic",
                config=dict(workload="forecast_n0 CenterPoint-VoxelNet car, fwd-only (BASELINE configs[1])",
                            batch_per_gpu=B, points_per_scene=int(pool_host[0][0].shape[0] // B), sweeps=10,
                            max_voxels=160000, l2="256 MiB memset between steps (inside the timed region)",
                            precision=args.precision, parallelism="scene replicas x%d, no collective" % world),
                e2e=dict(value=e2e_value, unit="scenes/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps),
                detect=dict(value=scenes / (ms_det / 1e3), unit="scenes/s", ms_per_step=ms_det / args.steps,
                            what="host points -> H2D -> forward -> CenterHead.predict (decode + rotated NMS on device) -> "
                                 "detections D2H", h2d_bytes_per_step=h2d, d2h_bytes_per_step=int(det_bytes[0])),
                batch1=dict(value=world * args.steps / (ms_b1 / 1e3), unit="scenes/s", ms_per_scene=ms_b1 / args.steps,
                            graph_value=world * args.steps / (ms_b1g / 1e3), graph_ms_per_scene=ms_b1g / args.steps,
                            what="one scene per step per GPU (the reference's samples_per_gpu), points resident; "
                                 "graph_*: the same forward replayed as one CUDA graph (graphs.GraphedForward)"),
                voxelize_spconv_ms_per_scene=vs_ms,
                gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_voxelize=vox_roof,
                cpu_baseline=cpu_baseline)
    if bar is not None:
        line["dense_bar"] = bar
    if stress is not None:
        line["stress_500k_two_task"] = stress
    if train_info is not None:
        line["train"] = train_info
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16,
                    help="scenes per step per GPU (throughput setting: 280 / 298 / 308 scenes/s at 4 / 8 / 16 on one B200; "
                         "the reference's samples_per_gpu is 1: 200 scenes/s)")
    ap.add_argument("--precision", default=os.environ.get("FD_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"],
                    help="bf16x3 (default): tcgen05 tensor cores with a 3-term bf16 split, holds the 1e-3 parity contract; "
                         "fp32: CUDA-core exact arm; bf16: single pass, outside the parity contract")
    ap.add_argument("--train-steps", type=int, default=8, help="timed forecast_n3 fwd+bwd steps reported under 'train' (0: skip)")
    ap.add_argument("--no-cudnn-bar", action="store_true", help="skip the cuDNN-via-torch neck+head comparison")
    ap.add_argument("--train-eager", action="store_true", help="do not replay the training step as a CUDA graph")
    ap.add_argument("--train-batch", type=int, default=0,
                    help="samples per GPU per training step (0: 1 on one GPU = configs[2], 4 on several = configs[3])")
    ap.add_argument("--train-precision", default="bf16x3", choices=["fp32", "bf16x3"],
                    help="forward / data-gradient convolutions of the training step (weight gradients are always fp32)")
    args = ap.parse_args()
    rank, world, local = dist_setup(args.gpus, init=args.impl != "reference")
    if args.impl == "reference":
        run_reference(args, rank, world)        # rank 0 alone; no process group needed
        return
    args.warmup = max(args.warmup, 3)
    run_gpu(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
