"""Scratch: time the fused rollout of one library variant (path in argv[1]) on UPS VRPTW-100."""
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
args = pkg.task_specific_params.default_args("vrptw100")
x = torch.from_numpy(pkg.data.create_dataset("vrptw", B, 100, 2, ups=ups).astype(np.float32)).cuda()
eng = pkg.Engine(args, gemm=gemm)
eng.set_weights(pkg.data.glorot_params(args, 1234))
eng.rollout(x, mode, seed=1); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.rollout(x, mode, seed=2); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(sys.argv[1], gemm, B, mode, "ups" if ups else "uniform", eng.last_launch()["rows_per_cta"], "ms per launch:", " ".join(f"{t:.2f}" for t in ts), "cost", float(out.R.mean()), flush=True)
