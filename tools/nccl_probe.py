"""Diagnose the gradient exchange on this box: NCCL transport (NCCL_DEBUG=INFO) and all-reduce time of a 31 MB bucket."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(31211164 // 4, device="cuda")
for _ in range(3):
    dist.all_reduce(x)
torch.cuda.synchronize()
for n in (1, 5):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("allreduce 31MB x%d: %.3f ms each" % (n, e0.elapsed_time(e1) / n), flush=True)
# async + wait pattern used by GradSync
hs = []
dist.barrier(); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3):
    h = dist.all_reduce(x, async_op=True); h.wait(); x.mul_(0.5)
torch.cuda.synchronize()
if rank == 0:
    print("async pattern: %.3f ms each (wall)" % ((time.perf_counter() - t) / 3 * 1e3), flush=True)
    print("can_device_access_peer", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
