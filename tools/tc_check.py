"""Scratch: first-light check of the tcgen05 path against the FFMA path and the oracle."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
from oracle import restatement as O
task, B = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("vrptw10", 64)
args = O.default_args(task)
params = O.init_params(args, seed=3, bias_scale=0.2)
data = O.create_dataset(task, B, seed=5)
x = torch.from_numpy(data.astype(np.float32)).cuda()
outs = {}
for gemm in ("ffma", "tc2", "tc3"):
    eng = pkg.Engine(args, gemm=gemm)
    eng.set_weights(params)
    o = eng.rollout(x, "greedy", want_logprobs=True, trace=True)
    torch.cuda.synchronize()
    outs[gemm] = o
    print(gemm, "launch", eng.last_launch(), "mean cost", o.R.mean().item(), flush=True)
a = outs["ffma"]
for nm in ("tc2", "tc3"):
    b = outs[nm]
    same = (a.idxs == b.idxs).all(0)
    print(nm, "vs ffma identical tours:", same.float().mean().item())
    um = (a.masks == 0) & (b.masks == 0)
    um[:, ~same] = False
    d = (a.logits - b.logits).abs()[um]
    print("  logit diff (unmasked, agreeing rows): max", d.max().item(), "mean", d.mean().item())
if B <= 512:
    ref = O.rollout(params, args, data, "greedy", keep_probs=True)
    for k, o in outs.items():
        idx = o.idxs.cpu().numpy()
        agree = (idx == ref.idxs).all(0)
        dl = np.abs(o.logits.cpu().numpy() - ref.logits)[(ref.masks == 0) & (o.masks.cpu().numpy() == 0)]
        print(k, "vs oracle: identical tours", agree.mean(), "logit max diff", dl.max(), "mean", dl.mean())
