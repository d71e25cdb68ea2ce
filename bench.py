#!/usr/bin/env python
"""bench.py -- forward scenes/sec of the LiDAR hot path on synthetic 10-sweep nuScenes-shaped scenes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (fused voxelize+VFE -> SpMiddleResNetFHD -> RPN -> CenterHead) over one
batch of synthetic scenes (BASELINE.json configs[1]: forecast_n0 CenterPoint-VoxelNet car, ~300k-point
10-sweep scene, forward only).  One process per GPU; under torchrun each rank runs independent scenes
(the path shards by scene, no data-path collective) and rank 0 prints ONE JSON line.

`--impl reference` times the reference's CPU path restated in oracle/ (the reference is Python: numba
voxelizer, external spconv, torch neck/head -- nothing compiles into oracle/_ref) with all host threads.
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


def run_reference(args, rank, world):
    """--impl reference: rank 0 only; bounded sample per step so the run ends within a few minutes."""
    if rank != 0:
        return
    model = build_model()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    full_pts = len(synth_scene(N_TARGET, seed=0))
    budget_s = 170.0
    if os.environ.get("FD_REF_N_TARGET"):          # test hook: tiny scene
        n_target = int(os.environ["FD_REF_N_TARGET"])
        frac = max(N_TARGET // n_target, 1)
    else:
        # probe with a 1/8 scene to size the per-step sample
        t_probe, n_probe = cpu_time_scene(sd, N_TARGET // 8, seed=99)
        est_full = t_probe * 8
        frac = 1
        while frac < 8 and est_full / frac * (args.steps + args.warmup) > budget_s:
            frac *= 2
        n_target = N_TARGET // frac
    for w in range(args.warmup):
        cpu_time_scene(sd, n_target, seed=1000 + w)
    t_total, pts_total = 0.0, 0
    for s in range(args.steps):
        t, n = cpu_time_scene(sd, n_target, seed=s)
        t_total += t
        pts_total += n
    scenes_equiv = pts_total / full_pts
    value = scenes_equiv / t_total
    cores = torch.get_num_threads()
    sample = ("%d step(s) of one synth_scene(%d) (%d pts; 1/%d of the 305,677-pt scene, scaled by points) through the "
              "oracle port: C voxelizer (1 thread) + restated spconv-CPU + torch RPN/CenterHead"
              % (args.steps, n_target, pts_total // max(args.steps, 1), frac))
    line = dict(metric=METRIC, value=value, unit="scenes/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_total / max(args.steps, 1), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="forecast_n0 CenterPoint-VoxelNet car, fwd-only, CPU reference path", batch=1,
                            points_per_scene=full_pts, host_cpus=os.cpu_count()),
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


def train_bench(args, rank, world, dev):
    """BASELINE configs[2]/[3]: forecast_n3 (7-timestep heads) forward + backward (+ AdamW step) per sample, one
    305k-point scene per GPU per step; for world > 1 the parameter gradients are all-reduced over NCCL (bucketed,
    overlapped with backward: shard.GradSync).  Auxiliary measurement, reported under "train" in the JSON line."""
    import futuredet_b200 as fb
    from futuredet_b200 import lib, shard, train
    from futuredet_b200.synth import synth_targets
    torch.manual_seed(0)
    m = fb.build_detector(model_cfg(timesteps=7)).to(dev).train()
    m.configure_voxelizer(VOXEL_CFG, training=True)
    shard.broadcast_parameters(m)
    tr = train.NativeTrainer(m, precision=args.train_precision)
    sync = shard.GradSync(tr.grads)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    B = 1
    scene = synth_scene(N_TARGET, seed=shard.scene_seed(rank, 7, 0, B))
    pts = torch.from_numpy(scene).to(dev)
    off = torch.tensor([0, len(scene)], dtype=torch.int32, device=dev)
    ex = synth_targets(B, 180, 180, 7, seed=rank)
    ex = {k: [[t.to(dev) for t in ts] for ts in v] for k, v in ex.items()}
    losses = None

    # single rank: the sync-free step (764 launches) is replayed as ONE CUDA graph; several ranks keep the eager step so
    # that the bucket all-reduces interleave with backward
    graphed = None
    if world == 1 and not args.train_eager:
        from futuredet_b200 import graphs
        graphed = graphs.GraphedTrainStep(tr, max_points=pts.shape[0], batch_size=B)

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
    return dict(workload="forecast_n3 (7-timestep heads) car, fwd+bwd+AdamW, 1 x 305k-pt scene per GPU per step "
                         "(BASELINE configs[2]; configs[3] for n_gpus > 1)", ms_per_step=ms,
                samples_per_s=B * world / (ms / 1e3), precision=args.train_precision, steps=args.train_steps,
                launch_mode="cuda graph replay" if graphed is not None else "eager",
                gpu_launches_per_step=(lib.launch_count() - n0) // args.train_steps, loss=loss,
                params=sum(p.numel() for p in m.parameters()), clocks=clocks,
                allreduce_bytes_per_step=sync.bytes_reduced // max(args.train_steps, 1),
                exchange="bucketed all-reduce(sum)/world of the parameter gradients, %s" %
                         ("NCCL over NVLink, overlapped with backward" if world > 1 else "single rank: none"))


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
    traffic = None
    tpath = os.path.join(REPO, "profiles", "r1_traffic.json")      # dram bytes of the same kernels from one ncu capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("conv_family_dram_bytes_per_step")
    roofline = dict(bound="tensor", kernel="gather->implicit-GEMM conv family (%s), %d launches/step" %
                    (args.precision, prof["launches"]), achieved=achieved, peak=pk["tf_sustained"], unit="TFLOP/s",
                    frac=achieved / pk["tf_sustained"], traffic=traffic, peak_source=pk["src"] + " bf16 dense, sustained",
                    algorithmic_gflop_per_step=prof["flop"] / 1e9, kernel_ms_per_step=prof["ms"], by_kind=prof["by_kind"])
    cpu_t, cpu_n = cpu_time_scene(sd_cpu, N_TARGET, seed=0)
    cpu_baseline = dict(value=1.0 / cpu_t, unit="scenes/s", cores=torch.get_num_threads(), kind="port",
                        sample="1 synth_scene(%d) (%d pts) through the oracle port (C voxelizer + restated spconv-CPU + "
                               "torch RPN/CenterHead), %.1f s" % (N_TARGET, cpu_n, cpu_t), host_cpus=os.cpu_count())
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
                gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_voxelize=vox_roof,
                cpu_baseline=cpu_baseline)
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
    ap.add_argument("--train-eager", action="store_true", help="do not replay the training step as a CUDA graph")
    ap.add_argument("--train-precision", default="bf16x3", choices=["fp32", "bf16x3"],
                    help="forward / data-gradient convolutions of the training step (weight gradients are always fp32)")
    args = ap.parse_args()
    rank, world, local = dist_setup(args.gpus, init=args.impl != "reference")
    if args.impl == "reference":
        if args.steps > 3 and "--steps" not in sys.argv:
            args.steps, args.warmup = 3, 1
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
