cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
for i in 1 2 3; do timeout 600 python bench.py --batch 1 --steps 2 --warmup 3 --train-steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TRAIN', d['train']['ms_per_step'], d['train']['clocks'])"; done
timeout 300 python tools/train_probe.py 2>&1 | grep PROBE
