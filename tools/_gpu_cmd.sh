cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3 | cut -c1-800
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH fwd', d['value'], 'train', d['train']['ms_per_step'], d['train']['clocks'])"; done
