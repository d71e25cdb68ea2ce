cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -k "forward_host or two_task" 2>&1 | tail -6 | cut -c1-1500
timeout 900 python bench.py --train-steps 0 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -c 300 gpurun_out/bench_e2e.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1]); print('BENCH', d['value'], d['e2e'], d['detect']['value'])"
