import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
import subprocess
if len(sys.argv) > 1:
    L = pkg._lib
    so = os.path.join(ROOT, "gpurun_out", "libvrpb200_dbg.so")
    subprocess.run(["/usr/local/cuda/bin/nvcc"] + L.NVCC_FLAGS + sys.argv[1:] + ["-o", so, os.path.join(L.CSRC, "vrpb200_api.cu")], check=True)
    L.LIB_PATH = so
from oracle import restatement as O
task, B = "vrptw10", 16
args = O.default_args(task)
params = O.init_params(args, seed=3, bias_scale=0.2)
data = O.create_dataset(task, B, seed=5)
x = torch.from_numpy(data.astype(np.float32)).cuda()
outs = {}
for gemm in ("ffma", "tc3"):
    eng = pkg.Engine(args, gemm=gemm, rows_per_cta=16)
    eng.set_weights(params)
    outs[gemm] = eng.rollout(x, "greedy", trace=True)
    torch.cuda.synchronize()
a, b = outs["ffma"], outs["tc3"]
for t in range(4):
    um = (a.masks[t] == 0) & (b.masks[t] == 0)
    d = (a.logits[t] - b.logits[t]).abs() * um
    print("t", t, "idx same", (a.idxs[t] == b.idxs[t]).float().mean().item(), "max logit diff", d.max().item(), "masks same", torch.equal(a.masks[t], b.masks[t]))
    if False:
        print(" ffma row0", a.logits[t, 0].cpu().numpy().round(4))
        print(" tc3  row0", b.logits[t, 0].cpu().numpy().round(4))
        print(" ffma row5", a.logits[t, 5].cpu().numpy().round(4))
        print(" tc3  row5", b.logits[t, 5].cpu().numpy().round(4))
