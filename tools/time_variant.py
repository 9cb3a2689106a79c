"""Scratch: time the fused rollout of one library variant.
usage: time_variant.py <lib.so> <gemm> <B> [mode] [ups|uniform] [task] [rows_per_cta]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
L = pkg._lib
L.LIB_PATH = os.path.join(L.LIB_DIR, sys.argv[1]); L._lib = None
gemm, B = sys.argv[2], int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "greedy"
ups = (sys.argv[5] == "ups") if len(sys.argv) > 5 else True
task = sys.argv[6] if len(sys.argv) > 6 else "vrptw100"
rb = int(sys.argv[7]) if len(sys.argv) > 7 else 0
args = pkg.task_specific_params.default_args(task)
x = torch.from_numpy(pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 2, ups=ups).astype(np.float32)).cuda()
eng = pkg.Engine(args, gemm=gemm, rows_per_cta=rb)
eng.set_weights(pkg.data.glorot_params(args, 1234))
eng.rollout(x, mode, seed=1); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.rollout(x, mode, seed=2); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ll = eng.last_launch()
print(sys.argv[1], gemm, task, B, mode, "ups" if ups else "uniform", "rb", ll["rows_per_cta"], "fam", ll["family"], "grid", ll["grid"], "ms per launch:",
      " ".join(f"{t:.3f}" for t in ts), "cost", float(out.R.mean()), flush=True)
