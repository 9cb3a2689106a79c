"""Scratch profiling: build a -DVRPB200_PHASE_CLOCKS variant of the library and print SM cycles per phase
(read by warp 2 lane 0 of every CTA with clock64, summed over CTAs)."""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
L = pkg._lib
so = os.path.join(L.LIB_DIR, "libvrpb200_clk.so")   # prebuilt on the CPU box (travels with the snapshot): python tools/phase_clocks.py --build
if "--build" in sys.argv or not os.path.exists(so):
    L.build_library(force=True, extra_flags=["-DVRPB200_PHASE_CLOCKS"], out_path=so)
    if "--build" in sys.argv:
        sys.exit(0)
L.LIB_PATH = so
L._lib = None
_ = 0
task, B, gemm, rb = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0
W = int(os.environ.get("BEAM", "0"))          # BEAM=10: beam search with 10 beams (rollout3)
args = pkg.task_specific_params.default_args(task)
eng = pkg.Engine(args, gemm=gemm, rows_per_cta=rb)
eng.set_weights(pkg.data.glorot_params(args, 1))
x = torch.from_numpy(pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 2, ups=bool(os.environ.get("UPS"))).astype(np.float32)).cuda()
eng.rollout(x, "beam_search", beam_width=W) if W else eng.rollout(x)
stats = torch.zeros(64, dtype=torch.int64, device="cuda")
tr = L.Trace(); tr.d_stats = stats.data_ptr()
RW = B * max(W, 1)
idx = torch.empty((eng.T, RW), dtype=torch.int32, device="cuda"); cost = torch.empty(RW, device="cuda")
par = torch.empty((eng.T, RW), dtype=torch.int32, device="cuda")
MODE = 1 if os.environ.get("STOCHASTIC") else (2 if W else 0)     # STOCHASTIC=1: sampling with the in-kernel stream
rc = eng.lib.vrpb200_rollout(eng.handle, C.c_void_p(x.data_ptr()), B, MODE, max(W, 1), None, None, C.c_uint64(0), C.c_void_p(idx.data_ptr()),
                             C.c_void_p(cost.data_ptr()), None, C.c_void_p(par.data_ptr()) if W else None, C.byref(tr), None)
torch.cuda.synchronize()
s = stats.cpu().numpy()
blocksteps = s[2] / eng.last_launch()["rows_per_cta"]
names = (["wait gates MMA", "LSTM epilogue + masks", "C0 items + wait q", "q epilogue", "C1 scores", "C2 select/env"] if gemm == "tc3" else
         ["A: wait MMA (TMA+tcgen05)", "A: LSTM epilogue + sync", "B: wait MMA", "B: q epilogue + sync", "C: rows", "C: step barrier wait"])
print(task, B, gemm, eng.last_launch(), "block-steps", blocksteps)
if gemm == "tc5":
    sn = {0: "T0' barrier (wait for the aux warps)", 1: "LSTM epilogue + barrier", 15: "offsets + barrier", 11: "first tiles staged", 2: "wait q", 3: "q epilogue + barrier",
          4: "C1: (loop top)", 5: "C1: wait MMA", 6: "C1: TMEM load, free", 7: "C1: arithmetic", 8: "C1: psum / logits / next F", 10: "after the last tile"}
    an = {0: "wait T0 + item lists", 1: "prefetch + wait tiles", 2: "choice", 3: "Env.step", 5: "next mask + item list", 4: "row end"}
    print("  score warp 2")
    for i in (0, 1, 15, 11, 2, 3, 4, 5, 6, 7, 8, 10):
        print(f"    {sn[i]:40s} {s[4 + i] / blocksteps:10.0f} clk/step")
    print(f"    total {s[4:20].sum() / blocksteps:.0f} clk/step")
    print("  aux warp 16")
    for i in (0, 1, 2, 3, 5, 4):
        print(f"    {an[i]:40s} {s[20 + i] / blocksteps:10.0f} clk/step")
    print(f"    total {s[20:27].sum() / blocksteps:.0f} clk/step;  near-tie softmax visits (row-steps of this warp): {s[27]}")
    sys.exit(0)
if gemm == "tc6":
    nm = {0: "wait gates (W_h pass of both teams)", 1: "offsets + LSTM epilogue + team barrier", 2: "C1 set-up (first tiles)", 3: "wait q", 4: "q epilogue + team barrier",
          18: "wait for the score token", 11: "C1: loop top", 7: "C1: wait MMA", 9: "C1: TMEM load, free", 10: "C1: arithmetic", 8: "C1: psum / logits / next F",
          5: "after scores / after rows", 12: "rows: wait for logits", 14: "rows fast: prefetch issue", 15: "rows fast: arg-max + merge", 16: "rows fast: Env.step",
          17: "rows fast: feasibility + lists", 13: "rows: rest", 6: "consts + end-of-step team barrier"}
    for who, o in (("team 0, warp 2", 4), ("team 1, warp 2", 36)):
        print(" ", who)
        for i in (0, 1, 2, 3, 4, 18, 11, 7, 9, 10, 8, 5, 12, 14, 15, 16, 17, 13, 6):
            print(f"    {nm[i]:42s} {s[o + i] / blocksteps:10.0f} clk/step")
        print(f"    total {s[o:o + 25].sum() / blocksteps:.0f} clk/step")
    print(f"  near-tie softmax visits (warp-steps): {s[29]} = {s[29] / blocksteps:.3f} per block-step")
    sys.exit(0)
if gemm in ("tc4", "tc4s", "tc4x"):
    names = ["wait gates MMA", "LSTM epilogue", "C1 set-up (first tiles)", "wait q", "q epilogue + barrier", "C1 tiles | rows (C2, Env, mask, items)",
             "end-of-step barrier", "C1: wait MMA", "C1: next F tile + group barrier", "C1: TMEM load, free, issue", "C1: arithmetic", "C1: tile end / loop", "rows: wait for logits", "rows: processing (rest)", "rows fast: prefetch issue", "rows fast: arg-max + merge", "rows fast: Env.step", "rows fast: feasibility + lists"]
    for who, o in (("warp 2", 4),):
        print(" ", who)
        for n, v in zip(names, s[o:o + len(names)]):
            print(f"    {n:40s} {v / blocksteps:10.0f} clk/step")
        print(f"    {'row offsets of the step':40s} {s[o + 19] / blocksteps:10.0f} clk/step")
        print(f"    total {s[o:o + 25].sum() / blocksteps:.0f} clk/step")
    print(f"  near-tie softmax visits (warp-steps): {s[29]} = {s[29] / blocksteps:.3f} per block-step")
    inames = ["wait for h' (bar_x)", "q: wait for weights", "q: issue + commit", "gates: wait for weights", "gates: throttle (previous stage's MMAs)", "gates: issue + commit"]
    print("  MMA issuer (warp 17)")
    for n, v in zip(inames, s[40:46]):
        print(f"    {n:40s} {v / blocksteps:10.0f} clk/step")
    print(f"  exact-tail visits (warp-steps, all warps): {s[30]} = {s[30] / blocksteps:.3f} per block-step; rows without un-masked node: {s[31]}")
    sys.exit(0)
for n, v in zip(names, s[4:4 + len(names)]):
    print(f"  {n:40s} {v / blocksteps:10.0f} clk/step")
print(f"  total {s[4:16].sum() / blocksteps:.0f} clk/step")
