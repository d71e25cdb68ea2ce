cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_predict.py tests/test_gpu_pipeline.py -q 2>&1 | tail -25 | cut -c1-1200
cat > /tmp/pred_once.py <<'PY'
import torch, sys
sys.path.insert(0, "tests")
from futuredet_b200 import predict as P
from oracle import predict_ref as PR
from test_gpu_predict import _Head, to_head_views, TEST_CFG
dev = torch.device("cuda:0")
for bg, n_obj in ((-2.0, 80), (-4.0, 80)):
    preds = PR.synth_preds(4, 180, 180, 7, seed=1, n_obj=n_obj, background=bg)
    v = to_head_views(preds, dev)
    for _ in range(3):
        P.center_head_predict(_Head(7), {}, [v], TEST_CFG)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        r = P.center_head_predict(_Head(7), {}, [v], TEST_CFG)
    e1.record(); torch.cuda.synchronize()
    print("PREDICT 4 scenes x 7 timesteps, background %.0f: %.3f ms per call (incl. count readback), kept %s" % (bg, e0.elapsed_time(e1) / 20, [len(x["cells"]) for x in r]))
PY
python /tmp/pred_once.py 2>&1 | tail -3
