#!/usr/bin/env python
"""bench.py - decoded instances/s of the batched policy rollout (BASELINE.json metric), on this repo's CUDA path and,
beside it, on the reference algorithm's CPU execution.

    python bench.py --gpus N --steps K --warmup W                     # headline: BASELINE config 5
    python bench.py --workload vrptw50_beam10 ...                     # any other BASELINE config (see WORKLOADS)
    python bench.py --all-workloads --steps 5 --warmup 3              # one JSON line per BASELINE config
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference algorithm on the host CPU

One "step" = one pass of the hot path (RLAgent.build_model(decode_type) + evaluate_batch of the reference,
model/attention_agent.py:84-240, 475-518) over one batch of synthetic instances.  Default workload = BASELINE.json
configs[4]: UPS VRPTW, 100 customers, greedy, 1 M instances over 8 GPUs = 125 000 instances per GPU per step (weak
scaling; instances are sharded, no collective on the decode path, one NCCL gather of the tour costs per step).
VRPTW-100 is an extrapolated task (the reference's table stops at vrptw50): n_nodes 101, capacity 50, decode_len 180.

value   : whole-job instances/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e     : the same through the host-buffer entry point (greedy / sampling: vrpb200_rollout_host, pinned host input,
          chunked H2D / decode / D2H of costs and tours inside the timed region; beam search: pinned H2D copy + rollout +
          D2H of costs, tours and beam parents).
roofline: the path is neither HBM- nor tensor-bound (2.7 KB of HBM traffic per instance against ~50 MFLOP and ~2 M
          transcendentals): the binding unit is the MUFU pipe (the tanh grid).  Pipe peaks are measured live with the
          library's register-only micro-benchmarks; the HBM line uses MEASURED_PEAKS.json.  Work is counted by the kernel
          (executed node evaluations and live row-steps), not taken from the nominal formula.
cpu_baseline / --impl reference: TensorFlow is not installable in this image and the reference is TF-1 graph code, so the
          CPU arm is the as-written restatement of the reference graph on torch-CPU fp32 (oracle/restatement_torch.py,
          vectorised multi-threaded kernels like TF's; torch.set_num_threads(all cores); warm-up 1, best of 3 for the
          cpu_baseline leg) on a bounded sample of the same workload; the sample size is what its `config` reports.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADLINE_METRIC = "decoded instances/sec (VRPTW-100 greedy)"
WORKLOADS = {
    # name: (task, ups instances, decode type, beam width, instances per GPU per step, BASELINE config, CPU sample)
    "ups_vrptw100_greedy": ("vrptw100", True, "greedy", 1, 125000, 5, 256),
    "vrptw100_stochastic": ("vrptw100", False, "stochastic", 1, 8192, 4, 256),     # 65 536 over 8 GPUs
    "vrptw50_beam10": ("vrptw50", False, "beam_search", 10, 16384, 3, 64),
    "vrptw20_greedy": ("vrptw20", False, "greedy", 1, 4096, 2, 1024),
    "vrp10_greedy": ("vrp10", False, "greedy", 1, 128, 1, 128),
    # not BASELINE configs: the headline task on the reference's own (uniform) generator, and VRPTW-50 greedy
    "vrptw100_greedy": ("vrptw100", False, "greedy", 1, 125000, 0, 256),
    "vrptw50_greedy": ("vrptw50", False, "greedy", 1, 131072, 0, 512),
}
BASELINE_ORDER = ["vrp10_greedy", "vrptw20_greedy", "vrptw50_beam10", "vrptw100_stochastic", "ups_vrptw100_greedy"]
KERNEL_NAMES = {1: "rollout_ffma_kernel", 3: "rollout3_kernel", 6: "rollout4_kernel"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ups_vrptw100_greedy", choices=sorted(WORKLOADS))
    ap.add_argument("--all-workloads", action="store_true", help="one JSON line per BASELINE config (the headline last)")
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU baseline sample (0: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--kernel", default="auto", help="fused kernel family (Engine gemm=...; default: the library's choice)")
    return ap.parse_args()


def make_args(pkg, task, W):
    return pkg.task_specific_params.default_args(task, beam_width=W)


def make_instances(pkg, task, ups, n, seed):
    a = make_args(pkg, task, 1)
    return pkg.data.create_dataset(a["task_name"], n, a["n_cust"], seed, ups=ups).astype(np.float32)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_times(pkg, task, ups, decode_type, W, sample, steps, warmup):
    """The reference algorithm on the host cores: torch-CPU fp32 restatement of the TF-1 graph, every core."""
    import torch
    from oracle import restatement_torch as OT   # the one place bench.py executes oracle/: the CPU arm
    torch.set_num_threads(os.cpu_count())
    args = make_args(pkg, task, W)
    params = OT.to_torch_params(pkg.data.glorot_params(args, seed=1234))
    data = torch.from_numpy(make_instances(pkg, task, ups, sample, seed=4242))
    T = args["decode_len"]
    rnd = np.random.RandomState(7)
    u1 = rnd.uniform(size=(T, sample)).astype(np.float32) if decode_type == "stochastic" else None
    u2 = rnd.uniform(size=(T, sample)).astype(np.float32) if decode_type == "stochastic" else None
    for _ in range(warmup):
        OT.rollout(params, args, data[: max(8, sample // 8)], decode_type, u1[:, : max(8, sample // 8)] if u1 is not None else None,
                   u2[:, : max(8, sample // 8)] if u2 is not None else None)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        OT.rollout(params, args, data, decode_type, u1, u2)
        times.append(time.perf_counter() - t0)
    return times


def cpu_label(sample, cores, what):
    return (f"{sample} instances of the workload per step, {what}; as-written restatement of the reference graph on torch-CPU fp32 "
            f"(TensorFlow is not installable in this image), torch.set_num_threads({cores}) = all host cores")


def run_workload(a, pkg, name, rank, world, local_rank, dist):
    task, ups, decode_type, W, default_b, cfg_no, cpu_sample = WORKLOADS[name]
    Bg = a.batch_per_gpu or default_b
    sample = a.cpu_sample or cpu_sample
    args = make_args(pkg, task, W)
    T, N, D = args["decode_len"], args["n_nodes"], args["input_dim"]
    cores = os.cpu_count()
    metric = HEADLINE_METRIC if name == "ups_vrptw100_greedy" else f"decoded instances/sec ({name})"
    config = {"workload": name, "baseline_config": cfg_no, "task": task, "n_nodes": N, "decode_len": T,
              "instances_per_gpu_per_step": Bg, "instances_per_step": Bg * max(world, a.gpus if a.impl == "ours" else 1),
              "decode": decode_type, "beam_width": W, "embedding_dim": args["embedding_dim"],
              "hidden_dim": args["hidden_dim"], "weights": "random-init (Glorot uniform, zero biases), seed 1234",
              "instances": "UPS Gaussian-mixture generator" if ups else "uniform generator",
              "task_note": "vrptw100 is extrapolated: the reference's task table stops at vrptw50" if task == "vrptw100" else "",
              "parallelism": f"instances sharded over {world} GPU(s), no decode-path collective, one NCCL gather of costs per step",
              "l2": f"input of one step is {Bg * N * D * 4 / 2**20:.0f} MiB per GPU (> 126 MB L2: no flush needed)" if Bg * N * D * 4 > 126e6
                    else "input smaller than L2: a 256 MiB buffer is written between timed steps"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return 0
        times = cpu_reference_times(pkg, task, ups, decode_type, W, sample, a.steps, min(a.warmup, 1))
        sec = float(np.mean(times))
        rate = sample / sec
        config.update(instances_per_gpu_per_step=sample, instances_per_step=sample, sample_of_instances_per_step=Bg * max(1, a.gpus),
                      parallelism="host CPU, all cores, one process", l2="n/a (CPU arm)")
        cb = {"value": rate, "unit": "instances/s", "cores": cores, "kind": "port",
              "sample": cpu_label(sample, cores, f"mean of {a.steps} timed passes")}
        line = {"impl": "reference", "metric": metric, "value": rate, "unit": "instances/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": rate, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "us_per_decode_step": sec * 1e6 / T}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- our arm (CUDA)
    import torch
    torch.cuda.set_device(local_rank)
    eng = pkg.Engine(args, device=local_rank, rows_per_cta=a.rows_per_cta, gemm=a.kernel)
    eng.set_weights(pkg.data.glorot_params(args, seed=1234))
    x_host = torch.from_numpy(make_instances(pkg, task, ups, Bg, seed=1000 + rank)).pin_memory()
    x_dev = x_host.cuda(non_blocking=False)
    R = Bg * W
    gathered = [torch.empty(R, device="cuda") for _ in range(world)] if (world > 1 and rank == 0) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if Bg * N * D * 4 <= 126e6 else None
    kw = dict(beam_width=W)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_no = [0]

    def step_device():
        step_no[0] += 1
        out = eng.rollout(x_dev, decode_type, seed=step_no[0], **kw)      # sampling: in-kernel counter-based stream
        if world > 1:   # the only collective of the path: tour costs to rank 0
            dist.gather(out.R, gathered, dst=0)
        return out

    for _ in range(a.warmup):
        out = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for k in range(a.steps):
        if flush is not None:
            flush.fill_(k)
        evs[k][0].record()
        out = step_device()
        evs[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    mean_cost = float(out.R.mean().item())
    launches_per_step = 2 if eng.last_launch()["family"] in (6,) else 1   # rollout4/5: feature re-packing kernel + fused rollout kernel

    # ---- end to end: host buffers in, host buffers out, every copy inside the timed region
    idx_host = torch.empty((T, R), dtype=torch.int32).pin_memory()
    cost_host = torch.empty((R,), dtype=torch.float32).pin_memory()
    par_host = torch.empty((T, R), dtype=torch.int32).pin_memory() if decode_type == "beam_search" else None
    xh, ih, ch = x_host.numpy(), idx_host.numpy(), cost_host.numpy()

    def step_e2e(k):
        if decode_type == "beam_search":
            xd = x_host.cuda(non_blocking=True)
            o = eng.rollout(xd, decode_type, **kw)
            cost_host.copy_(o.R, non_blocking=True); idx_host.copy_(o.idxs, non_blocking=True); par_host.copy_(o.beam_parents, non_blocking=True)
            torch.cuda.synchronize()
        else:
            eng.rollout_host(xh, decode_type, seed=k, idx_out=ih, cost_out=ch)

    for k in range(max(1, min(a.warmup, 2))):
        step_e2e(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        step_e2e(100 + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = eng.last_launch()["kernels"] * launches_per_step if decode_type != "beam_search" else launches_per_step
    t2 = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_s = float(t2.item())
    if decode_type == "greedy":
        assert np.array_equal(ch, out.R.cpu().numpy()), "host entry point and device entry point disagree"

    # ---- work accounting for the roofline (one extra, untimed, counted pass; the rollout is deterministic)
    st = eng.rollout(x_dev, decode_type, seed=step_no[0], want_stats=True, want_steps=True, **kw)
    torch.cuda.synchronize()
    node_evals, live_rowsteps, block_rowsteps, blocks = [int(v) for v in st.stats.cpu().tolist()]
    launch = eng.last_launch()
    if rank == 0:
        H = args["hidden_dim"]
        ms_step = ms_total / a.steps
        sec_step = ms_step * 1e-3
        fam = launch["family"]
        # executed work of one launch, from the kernel's own counters:
        #   node_evals     (row, node) pairs whose 128 tanh terms were evaluated
        #   live_rowsteps  live rows x executed steps: LSTM cell + query of a row that is still decoding (the GEMMs themselves
        #                  run over every row slot of a block, block_rowsteps; only live valid rows are counted as work)
        softmax = 0.0 if (decode_type == "greedy") else 1.0 * node_evals + 2.0 * live_rowsteps     # exp per item, log per row
        if fam in (6,):
            # 4 ex2 + ONE shared rcp per 4 tanh terms; LSTM gates 5 x (ex2 + rcp); greedy fast tail: no softmax
            mufu_ops = 1.25 * H * node_evals + 10.0 * H * live_rowsteps + softmax
        else:
            mufu_ops = 2.0 * H * node_evals + 10.0 * H * live_rowsteps + 3.0 * node_evals            # tanh grid, LSTM gates, softmax exp
        gemm_macs = (128 * 512 + 128 * 128 + 4 * 512) * live_rowsteps                              # h.W_h, h'.W_q, x.W_x
        score_macs = 5.0 * H * node_evals                                                          # feat.A + demand.g  (K = 5)
        if fam == 1:
            fp32_ops = 8.0 * H * node_evals + gemm_macs                                            # everything on the FP32 pipe
            tensor_flops = 0.0
        elif fam in (6,):
            fp32_ops = 7.25 * H * node_evals + 46.0 * H * live_rowsteps                            # add, clamp, +1, shared-reciprocal algebra (lane-ops; issued as packed FP32x2)
            tensor_flops = 3 * 2.0 * (gemm_macs + score_macs * 8.0 / 5.0)
            f16_flops = 3 * 2.0 * gemm_macs                                                        # LSTM / query GEMMs: kind::f16, two-term FP16 split (3 products)
        else:
            fp32_ops = 4.0 * H * node_evals + 46.0 * H * live_rowsteps
            tensor_flops = 3 * 2.0 * (gemm_macs + score_macs * 8.0 / 5.0)                          # K padded to 8 for the scores
        hbm_bytes = Bg * N * D * 4 + R * (T * 4 + 4) * (2 if decode_type == "beam_search" else 1)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src, tc_peak = 6650.0, "fallback (B200_PROFILING.md)", 1590.0
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                mp = json.load(f)
                hbm_peak, hbm_src, tc_peak = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json", float(mp.get("bf16_tflops", 1590.0))
        mufu_peak = eng.pipe_peak("mufu")
        fma_peak = eng.pipe_peak("fma")
        pipes = {
            "mufu": {"achieved": mufu_ops / sec_step / 1e9, "peak": mufu_peak / 1e9, "unit": "Gop/s"},
            "fp32": {"achieved": fp32_ops / sec_step / 1e9, "peak": fma_peak / 1e9, "unit": "Gop/s (FMA = 1 op)"},
            "hbm": {"achieved": hbm_bytes / sec_step / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src},
            "tensor": {"achieved": tensor_flops / sec_step / 1e12, "peak": tc_peak / 2.0, "unit": "TFLOP/s",
                       "note": "issued flops incl. the 3-product splits; peak = half the measured bf16 peak (TF32 runs at half the bf16 rate)"},
        }
        for p in pipes.values():
            p["frac"] = (p["achieved"] / p["peak"]) if p["peak"] else 0.0
        if fam in (6,):   # the GEMM part runs as kind::f16 (full bf16/fp16 rate), the score tiles as kind::tf32 (half rate)
            pipes["tensor"]["frac"] = (f16_flops / tc_peak + (tensor_flops - f16_flops) / (tc_peak / 2.0)) / sec_step / 1e12
            pipes["tensor"]["note"] = ("issued flops incl. the 3-product splits: LSTM / query GEMMs as kind::f16 (FP16 hi/lo) at the measured bf16 peak, "
                                       "score tiles as kind::tf32 (3xTF32) at half of it; frac = busy fraction of the two combined")
        bound = max(("mufu", "fp32", "hbm", "tensor"), key=lambda k: pipes[k]["frac"])
        roofline = {"bound": bound, "achieved": pipes[bound]["achieved"], "peak": pipes[bound]["peak"], "unit": pipes[bound]["unit"],
                    "frac": pipes[bound]["frac"], "traffic": None,
                    "peak_source": "pipe peaks measured live by vrpb200_pipe_peak (register-only micro-benchmark on all SMs; MUFU = chains of rcp(1 + ex2(x)), 16 lanes/clk/SM); "
                                   "MEASURED_PEAKS.json holds only HBM and bf16-tensor peaks and neither binds this path",
                    "work_per_launch": {"node_evaluations": node_evals, "live_row_steps": live_rowsteps, "block_row_steps": block_rowsteps,
                                        "blocks": blocks, "mean_steps_per_block": float(st.steps.float().mean().item()),
                                        "mufu_ops": mufu_ops, "fp32_lane_ops": fp32_ops, "tensor_flops": tensor_flops, "hbm_bytes": hbm_bytes},
                    "pipes": pipes,
                    "kernel": f"{KERNEL_NAMES.get(fam, '?')}<{launch['rows_per_cta']},{'TW' if D == 5 else 'VRP'}> grid {launch['grid']} x {launch['block']} thr, {launch['smem']} B smem, "
                              f"1 launch per step, {ms_step:.3f} ms per launch (CUDA events)"}
        # profiles/ holds the ncu capture of the same kernel; dram bytes per launch from it if present
        tr_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_path):
            with open(tr_path) as f:
                tr = json.load(f).get(name)
                if tr:   # DRAM bytes of one launch: measured per instance in the (smaller) ncu capture, scaled to this launch
                    roofline["traffic"] = tr["dram_bytes_per_instance"] * Bg
                    roofline["traffic_source"] = tr
        value = Bg * world * a.steps / (ms_total * 1e-3)
        d2h = R * 4 + T * R * 4 * (2 if decode_type == "beam_search" else 1)
        line = {"metric": metric, "value": value, "unit": "instances/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "us_per_decode_step": ms_step * 1e3 / T,
                "ns_per_row_step": ms_step * 1e6 / max(1, live_rowsteps), "mean_tour_length": mean_cost,
                "roofline": roofline, "clocks": clocks,
                "e2e": {"value": Bg * world * a.steps / e2e_s, "unit": "instances/s", "h2d_bytes_per_step": Bg * N * D * 4,
                        "d2h_bytes_per_step": d2h,
                        "api": "vrpb200_rollout_host (pinned host buffers, chunked copy/decode overlap)" if decode_type != "beam_search"
                               else "pinned H2D copy + vrpb200_rollout + D2H of costs, tours and beam parents",
                        "kernel_launches_per_step": e2e_launches},
                "gpu_launches": a.steps * launches_per_step}
        if not a.no_cpu_baseline and world == 1:
            times = cpu_reference_times(pkg, task, ups, decode_type, W, sample, 3, 1)
            line["cpu_baseline"] = {"value": sample / min(times), "unit": "instances/s", "cores": cores, "kind": "port",
                                    "sample": cpu_label(sample, cores, f"warm-up 1, best of 3 passes ({min(times):.1f} s)")}
        print(json.dumps(line), flush=True)
    del eng
    return 0


def main():
    a = parse()
    pkg = importlib.import_module("mit_rl-vrptw_b200")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and "RANK" not in os.environ and a.impl == "ours":   # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    names = BASELINE_ORDER if a.all_workloads else [a.workload]
    dist = None
    if a.impl == "ours":
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            print(json.dumps({"error": "no CUDA device; this repo has no CPU path for the rollout"}))
            return 1
        torch.cuda.set_device(local_rank)
        if world > 1:
            # NCCL's own banner ("NCCL version ...", printed on stdout with NCCL_DEBUG=VERSION) must not share stdout with the
            # JSON line: keep warnings only unless the caller asked for more, and send whatever NCCL prints to stderr
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rc = 0
    for name in names:
        rc |= run_workload(a, pkg, name, rank, world, local_rank, dist)
    if a.impl == "ours" and world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
