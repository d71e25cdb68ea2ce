cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_predict.py -q -x 2>&1 | tail -40 | cut -c1-1200
