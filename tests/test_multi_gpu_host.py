"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (scene sharding + throughput reduction) and the
bench.py --impl reference contract under a multi-rank launch (rank 0 prints, other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from futuredet_b200 import shard

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_indices(11, rank, world)
    seeds = [shard.scene_seed(rank, s, b, 2) for s in range(4) for b in range(2)]
    # rank 1 is "slower": whole-job throughput must use the max time and the summed scenes
    value, t_max, units = shard.reduce_throughput(100.0 if rank == 0 else 250.0, len(mine))
    q.put((rank, mine, seeds, value, t_max, units))
    dist.barrier()
    dist.destroy_process_group()


def test_scene_sharding_and_throughput_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, s0, v0, t0, u0), (r1, m1, s1, v1, t1, u1) = res
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)          # every scene exactly once
    assert not set(s0) & set(s1)                                                 # replicas never share a scene
    assert t0 == t1 == 250.0 and u0 == u1 == 11.0
    assert abs(v0 - 11 / 0.25) < 1e-9 and v0 == v1


def test_single_process_reduce_is_identity():
    v, t, u = shard.reduce_throughput(50.0, 4)
    assert (v, t, u) == (80.0, 50.0, 4.0)


def test_bench_reference_arm_contract_two_ranks():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints one JSON line, rank 1 exits 0 without work."""
    env = dict(os.environ, FD_REF_N_TARGET="6000", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    outs = []
    for rank in (1, 0):
        r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0"], env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == ""
    line = json.loads(outs[1].splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "scenes/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["higher_is_better"] is True and line["n_gpus"] == 2


def _grad_worker(rank, world, port, q, inline=False):
    from torch import nn
    from futuredet_b200 import train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                                   # ranks start with DIFFERENT parameters ...
    model = nn.Sequential(nn.Linear(5, 7), nn.BatchNorm1d(7), nn.Linear(7, 3))
    shard.broadcast_parameters(model)                         # ... and leave with rank 0's
    buckets = train.GradBuckets(list(model.parameters()), bucket_bytes=64)      # tiny buckets -> several all-reduces
    sync = shard.GradSync(buckets, inline=inline)     # inline: the mode captured into the multi-rank CUDA graph
    launched = []
    orig = sync._launch
    sync._launch = lambda i, flat: (launched.append(i), orig(i, flat))[1]
    buckets.on_bucket_ready = sync._launch
    buckets.zero()
    params = [p for p in model.parameters()]
    for j, p in enumerate(reversed(params)):                  # backward order
        buckets.grad(p).fill_(float(rank + 1) * (j + 1))
        buckets.done(p)
    buckets.flush()
    if inline:
        assert sync.handles == []                              # nothing left to wait for: already reduced in place
    sync.finish()
    got = [float(p.grad.flatten()[0]) for p in reversed(params)]
    psum = float(sum(p.detach().sum() for p in model.parameters()))
    q.put((rank, got, launched, len(buckets.buckets), psum, sync.bytes_reduced))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("inline", [False, True])
def test_gradient_buckets_allreduce_world2(inline):
    """The training exchange: bucketed all-reduce(avg) of the parameter gradients, launched as buckets
    complete in backward order (asynchronously, or inline on the launching stream = the CUDA-graph mode); parameters
    broadcast from rank 0 first."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (1000 if inline else 0)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q, inline)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, l0, nb0, ps0, by0), (_, g1, l1, nb1, ps1, by1) = res
    want = [1.5 * (j + 1) for j in range(len(g0))]            # mean of (1, 2) * (j + 1)
    assert g0 == want and g1 == want
    assert nb0 == nb1 and nb0 >= 2 and l0 == list(range(nb0)) and l1 == l0      # every bucket once, in backward order
    assert abs(ps0 - ps1) < 1e-6                                                 # same parameters after the broadcast
    n_params = 5 * 7 + 7 + 7 + 7 + 7 * 3 + 3
    assert by0 == by1 == 4 * n_params                                            # one exchange of every gradient byte
