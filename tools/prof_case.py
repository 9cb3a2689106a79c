"""One profiling case for ncu: a few fused rollouts of UPS VRPTW-100 instances (not a benchmark).
    ncu --set full --import-source on --clock-control none -k regex:rollout --launch-skip 1 -c 1 -o gpurun_out/prof python tools/prof_case.py tc4 14208"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
gemm = sys.argv[1] if len(sys.argv) > 1 else "tc4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 14208
task = sys.argv[3] if len(sys.argv) > 3 else "vrptw100"
args = pkg.task_specific_params.default_args(task)
eng = pkg.Engine(args, gemm=gemm)
eng.set_weights(pkg.data.glorot_params(args, 1234))
x = torch.from_numpy(pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 7, ups=not os.environ.get("UNIFORM")).astype(np.float32)).cuda()
for _ in range(3):
    out = eng.rollout(x, "greedy")
torch.cuda.synchronize()
print(gemm, B, eng.last_launch(), float(out.R.mean()))
