"""Graphed training-step timing only (bench.train_bench), for A/B of training kernels: python tools/train_bench.py [batch]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

a = argparse.Namespace(train_precision="bf16x3", train_eager=False, train_batch=int(sys.argv[1]) if len(sys.argv) > 1 else 1,
                       train_steps=10)
torch.cuda.set_device(0)
r = bench.train_bench(a, 0, 1, torch.device("cuda", 0))
print("train: %.2f ms/step, %.1f samples/s, batch %d, %s" % (r["ms_per_step"], r["samples_per_s"], r["batch_per_gpu"], r["launch_mode"]))
