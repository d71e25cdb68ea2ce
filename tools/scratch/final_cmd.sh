cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-1200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 400 gpurun_out/bench_final_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1]); print('BENCH', d['value'], d['e2e']['value'], d['detect']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['roofline_voxelize']['frac'], d['clocks'])"
timeout 300 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2>/dev/null; tail -c 600 gpurun_out/bench_final_ref.json
