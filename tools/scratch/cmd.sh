cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1]); print('BENCH', d['value'], d['e2e']['value'], d['detect']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --batch 1 --train-steps 0 > gpurun_out/bench_final_n1_batch1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_final_n1_batch1.json').read().strip().splitlines()[-1]); print('BENCH b1', d['value'], d['e2e']['value'], d['detect']['value'])"
