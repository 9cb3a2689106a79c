"""Scratch: rollout5 (tc5) against rollout4 (tc4x) - identical tours expected - and a first timing."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
cases = [("vrptw10", 64, False, 0), ("vrp10", 100, False, 0), ("vrptw20", 1000, False, 16), ("vrptw50", 3000, False, 32), ("vrptw100", 14208, True, 48), ("vrptw100", 125000, True, 0)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for task, B, ups, rb in cases:
    args = pkg.task_specific_params.default_args(task)
    params = pkg.data.glorot_params(args, 1234)
    x = torch.from_numpy(pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 2, ups=ups).astype(np.float32)).cuda()
    res = {}
    for gemm in ("tc4x", "tc5"):
        eng = pkg.Engine(args, gemm=gemm, rows_per_cta=rb)
        eng.set_weights(params)
        out = eng.rollout(x, "greedy")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = eng.rollout(x, "greedy")
        e1.record()
        torch.cuda.synchronize()
        full = eng.rollout(x, "greedy", want_logprobs=True, final_state=True)
        st = eng.rollout(x, "stochastic", seed=5)
        torch.cuda.synchronize()
        res[gemm] = (out, full, st, e0.elapsed_time(e1) / 3, eng.last_launch())
    a, b = res["tc4x"], res["tc5"]
    same = (a[0].idxs == b[0].idxs).all(0).float().mean().item()
    same_full = (a[1].idxs == b[1].idxs).all(0).float().mean().item()
    same_st = (a[2].idxs == b[2].idxs).all(0).float().mean().item()
    fast_eq_exact = (b[0].idxs == b[1].idxs).all(0).float().mean().item()
    lp = (a[1].logprobs - b[1].logprobs).abs().max().item()
    print(f"{task} B={B} ups={ups}: tc4x {a[3]:.3f} ms, tc5 {b[3]:.3f} ms ({B / b[3] * 1e3:,.0f} inst/s); identical tours fast {same:.4f} exact {same_full:.4f} "
          f"sampled {same_st:.4f}; tc5 fast=exact {fast_eq_exact:.4f}; max |dlogprob| {lp:.2e}; cost equal {torch.equal(a[0].R, b[0].R)}; launch {b[4]}", flush=True)
