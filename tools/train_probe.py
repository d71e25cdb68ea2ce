"""Phase timing of the data-parallel training step (perf triage, not a bench): backward+exchange vs optimizer, with and
without the gradient all-reduce, and the raw NCCL all-reduce time of the same bytes."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import futuredet_b200 as fb
from futuredet_b200 import shard, train
from futuredet_b200.synth import synth_scene, synth_targets
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    x = torch.ones(31211164 // 4, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("PROBE raw allreduce 31MB: %.3f ms" % (e0.elapsed_time(e1) / 5), flush=True)
torch.manual_seed(0)
m = fb.build_detector(bench.model_cfg(timesteps=7)).to(dev).train()
m.configure_voxelizer(bench.VOXEL_CFG, training=True)
shard.broadcast_parameters(m)
scene = synth_scene(bench.N_TARGET, seed=rank + 7)
pts = torch.from_numpy(scene).to(dev)
off = torch.tensor([0, len(scene)], dtype=torch.int32, device=dev)
ex = synth_targets(1, 180, 180, 7, seed=rank)
ex = {k: [[t.to(dev) for t in ts] for ts in v] for k, v in ex.items()}
for prec in ("bf16x3", "fp32"):
    for use_sync in ((True, False) if world > 1 else (False,)):
        tr = train.NativeTrainer(m, precision=prec)
        sync = shard.GradSync(tr.grads) if use_sync else None
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True)
        def step(ev=None):
            if ev: ev[0].record()
            tr.forward(ex, points=pts, batch_offsets=off)
            if ev: ev[1].record()
            tr.backward()
            if ev: ev[2].record()
            if sync: sync.finish()
            if ev: ev[3].record()
            opt.step()
            if ev: ev[4].record()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        acc = [0.0] * 4
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            step(ev)
            torch.cuda.synchronize()
            for i in range(4):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) / n
        wall = (time.perf_counter() - t0) / n * 1e3
        if rank == 0:
            print("PROBE prec=%s allreduce=%s: forward %.1f ms, backward %.1f ms, exchange-wait %.2f ms, adamw %.2f ms, wall %.1f ms"
                  % (prec, use_sync, acc[0], acc[1], acc[2], acc[3], wall), flush=True)
        tr.grads.on_bucket_ready = None
if world > 1:
    dist.destroy_process_group()
