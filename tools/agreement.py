"""Greedy-tour agreement of the fused kernel with the CPU oracle (numpy restatement of the reference graph), per task.
Prints the fraction of instances whose whole tour is identical and the relative cost difference of the others."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
from oracle import restatement as O
for task, B, ups in (("vrp10", 2000, False), ("vrptw20", 2000, False), ("vrptw50", 1000, False), ("vrptw100", 400, False), ("vrptw100", 400, True)):
    args = O.default_args(task)
    params = pkg.data.glorot_params(args, seed=7)
    data = pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 3, ups=ups)
    eng = pkg.Engine(args, gemm=os.environ.get("GEMM", "tc4x"))
    eng.set_weights(params)
    out = eng.rollout(torch.from_numpy(data.astype(np.float32)).cuda(), "greedy")
    ref = O.rollout(params, args, data, "greedy")
    idx = out.idxs.cpu().numpy().astype(np.int64)
    same = (idx == ref.idxs).all(0)
    R = out.R.cpu().numpy()
    rel = np.abs(R - ref.R) / ref.R
    print(f"{task:9s} ups={int(ups)} B={B}: identical tours {same.mean():.4%}; cost rel. diff: identical {rel[same].max() if same.any() else 0:.2e}, "
          f"others mean {rel[~same].mean() if (~same).any() else 0:.2e}", flush=True)
