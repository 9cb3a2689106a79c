"""Scratch: time the pieces of one REINFORCE training step (stochastic rollout, critic value, losses, actor gradient, critic
gradient, clip + Adam) with CUDA events on the launching stream.
usage: time_train_step.py [task] [B]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mit_rl-vrptw_b200")
task = sys.argv[1] if len(sys.argv) > 1 else "vrptw100"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
args = pkg.task_specific_params.default_args(task)
x = torch.from_numpy(pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 2, ups=False).astype(np.float32)).cuda()
params = pkg.data.glorot_params(args, 1234)
from oracle import restatement as O      # scratch tool only: Glorot critic variables
params.update(O.init_critic_params(args, seed=4321))
eng = pkg.Engine(args)
eng.set_weights(params)
an = {k: v.shape for k, v in params.items() if k.startswith("Actor/")}
cn = {k: v.shape for k, v in params.items() if k.startswith("Critic/")}
names = sorted(an) + sorted(cn)
var = [torch.from_numpy(np.ascontiguousarray(params[k], np.float32)).cuda() for k in names]
m, v2 = [torch.zeros_like(a) for a in var], [torch.zeros_like(a) for a in var]


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


for it in range(4):
    out, t_roll = timed(lambda: eng.rollout(x, "stochastic", seed=it + 1, want_logprobs=True))
    v, t_crit = timed(lambda: eng.critic(x))
    _, t_loss = timed(lambda: eng.reinforce_loss(out.R, v, out.logprobs))
    w = (out.R - v) / B
    ga, t_ag = timed(lambda: eng.actor_grad(x, out.idxs, w, an))
    (gc, _), t_cg = timed(lambda: eng.critic_grad(x, out.R, cn))
    grads = [ga[k] for k in sorted(an)] + [gc[k] for k in sorted(cn)]
    _, t_adam = timed(lambda: eng.clip_adam(var, grads, m, v2, 2.0, 1e-4, it + 1))
    tot = t_roll + t_crit + t_loss + t_ag + t_cg + t_adam
    print(f"{task} B={B} it {it}: rollout {t_roll:.3f} critic {t_crit:.3f} losses {t_loss:.3f} actor_grad {t_ag:.3f} critic_grad {t_cg:.3f} "
          f"clip+adam {t_adam:.3f}  total {tot:.3f} ms = {B / tot * 1e3:.0f} instances/s", flush=True)

# the same step on the host cores: stochastic rollout + critic + both gradient sets by torch autograd through the as-written graph
# (oracle/train_torch.py; test infrastructure, timed here as the CPU baseline of the training step), then the update in numpy
if os.environ.get("VRPB200_TRAIN_CPU", "1") == "1":
    import json, time
    from oracle import restatement_torch as RT, train_torch as TT
    torch.set_num_threads(os.cpu_count())
    xd = x.cpu().numpy()
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        u = np.random.RandomState(1).uniform(size=(args["decode_len"], B)).astype(np.float32)
        Rc, idxc = RT.rollout(params, args, xd, "stochastic", uniforms=u, uniforms_retry=u)
        _, gcr, vc = TT.critic_grads(params, args, xd, Rc)
        _, gac, _, _ = TT.actor_grads(params, args, xd, idxc, ((Rc - vc) / B).astype(np.float32))
        for k, g in list(gac.items()) + list(gcr.items()):
            TT.clip_adam(params[k], g, np.zeros_like(g), np.zeros_like(g), 1, 1e-4, 2.0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(json.dumps({"metric": "REINFORCE training step (rollout + critic + both gradients + clip + Adam), instances/s", "task": task,
                      "batch": B, "gpu_ms_per_step": round(tot, 3), "gpu_instances_per_s": round(B / tot * 1e3, 1),
                      "cpu_baseline": {"value": round(B / best, 2), "unit": "instances/s", "ms_per_step": round(best * 1e3, 1),
                                       "cores": os.cpu_count(), "kind": "port",
                                       "sample": "the same batch, best of 2; torch-CPU autograd through the as-written graph"},
                      "pieces_ms": {"rollout": round(t_roll, 3), "critic": round(t_crit, 3), "losses": round(t_loss, 3),
                                    "actor_grad": round(t_ag, 3), "critic_grad": round(t_cg, 3), "clip_adam": round(t_adam, 3)}}), flush=True)
