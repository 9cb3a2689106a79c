"""Scratch timing of the fused rollout at a few shapes (CUDA events).  Not the judged bench (bench.py)."""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
from oracle import restatement as O  # noqa: E402  (instance generator + weight init only)

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="vrptw20:4096,vrptw100:18944,vrptw100:131072")
ap.add_argument("--rb", type=int, default=0)
ap.add_argument("--mode", default="greedy")
ap.add_argument("--full", type=int, default=0)
ap.add_argument("--gemm", default="tc3")
a = ap.parse_args()
for case in a.cases.split(","):
    task, B = case.split(":")
    B = int(B)
    args = O.default_args(task)
    params = O.init_params(args, seed=1)
    base = O.create_dataset(task, min(B, 4096), seed=2).astype(np.float32)
    data = np.tile(base, [(B + len(base) - 1) // len(base), 1, 1])[:B]
    # de-duplicate tiled instances a little so rows are not identical
    x = torch.from_numpy(data).cuda()
    eng = pkg.Engine(args, rows_per_cta=a.rb, gemm=a.gemm)
    eng.set_weights(params)
    for _ in range(2):
        out = eng.rollout(x, a.mode, trace=False, want_steps=True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 3
    for _ in range(reps):
        out = eng.rollout(x, a.mode, want_steps=True)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    steps = out.steps.float().mean().item()
    print(f"{task} B={B} mode={a.mode}: {ms:.3f} ms  {B / ms * 1e3:,.0f} inst/s  mean steps/block {steps:.1f} of {args['decode_len']}"
          f"  cost {out.R.mean().item():.4f} launch {eng.last_launch()}", flush=True)
