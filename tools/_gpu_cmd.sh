cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3 | cut -c1-1500
for p in bf16x3 fp32; do timeout 600 python bench.py --steps 5 --warmup 3 --train-precision $p 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH fwd', d['value'], 'train', d['train']['precision'], d['train']['ms_per_step'])"; done
