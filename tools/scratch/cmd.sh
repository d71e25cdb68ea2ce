cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 8 --warmup 3 --train-steps 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 500 gpurun_out/bench_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print('BENCH N4', d['n_gpus'], d['value'], d['e2e']['value'], d['train'])"
