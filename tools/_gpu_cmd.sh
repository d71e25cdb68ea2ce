cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_assign.py -q 2>&1 | tail -40 | cut -c1-1500
