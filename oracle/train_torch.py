"""torch-CPU oracle of the backward / update side of RLAgent.build_train_step (model/attention_agent.py:272-322).
TEST INFRASTRUCTURE ONLY (tests/ import it; nothing under mit_rl-vrptw_b200/ does).

The forward graph is the as-written restatement of ``oracle/restatement_torch.py`` (pinned through ``oracle/restatement.py``
against fixtures produced by the reference's own code), run teacher-forced along a given tour with autograd switched on;
``torch.autograd`` then plays the part of ``optimizer.compute_gradients``.  Parity with real TensorFlow gradients is
unpinned (TensorFlow is not installable here): what is pinned is the forward graph that is differentiated.
"""
from __future__ import annotations

import numpy as np
import torch

from . import restatement_torch as RT


def teacher_forced_logprobs(p, args, x, idxs, input_scale=None):
    """Per-step log prob of the given decisions, [T, B] (attention_agent.py:139-220 with idx fixed), and the Env state seen by
    every step (demand [T,B,N], mask [T,B,N], load [T,B], time [T,B]).  input_scale [T, B, E]: the LSTM-input dropout of
    DropoutWrapper(input_keep_prob) (shared/decode_step.py:177-179) as keep-mask / keep-probability."""
    env = RT.Env(args, x)
    B, N, T = env.batch_size, env.n_nodes, args["decode_len"]
    emb = RT.conv1d_k1(env.input_pnt, p["Actor/Embedding/conv1d/kernel"], p["Actor/Embedding/conv1d/bias"])
    env.reset(1)
    rows = torch.arange(B)
    c, h = torch.zeros([B, args["hidden_dim"]]), torch.zeros([B, args["hidden_dim"]])
    decoder_input = emb[:, N - 1][:, None, :]
    lps, st = [], dict(demand=[], mask=[], load=[], time=[])
    for t in range(T):
        st["demand"].append(env.demand.clone()); st["mask"].append(env.mask.clone()); st["load"].append(env.load.clone())
        st["time"].append(env.time.clone() if env.tw else torch.zeros([B]))
        if input_scale is not None:
            decoder_input = decoder_input * torch.as_tensor(input_scale[t], dtype=torch.float32)[:, None, :]
        logit, prob, logprob, (c, h) = RT.decode_step(p, args, decoder_input, emb, env, c, h)
        idx = torch.as_tensor(idxs[t]).long()
        lps.append(torch.log(prob[rows, idx]))          # attention_agent.py:214
        env.step(idx)
        decoder_input = emb[rows, idx][:, None, :]
    return torch.stack(lps), {k: torch.stack(v).numpy() for k, v in st.items()}


def actor_grads(params, args, data, idxs, weight, input_scale=None):
    """Gradients of  sum_r weight[r] sum_t logprob[t, r]  w.r.t. every Actor variable (weight = (R - v) / B gives
    compute_gradients(actor_loss), attention_agent.py:289-297).  Returns (loss, {name: grad}, env states)."""
    p = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).requires_grad_(k.startswith("Actor/")) for k, v in params.items()}
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    lps, st = teacher_forced_logprobs(p, args, x, idxs, input_scale)
    loss = (torch.as_tensor(weight, dtype=torch.float32) * lps.sum(0)).sum()
    names = [k for k in p if k.startswith("Actor/")]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return float(loss.detach()), {k: (g.numpy() if g is not None else None) for k, g in zip(names, grads)}, st, lps.detach().numpy()


def critic_value(p, args, x):
    """v of build_model('stochastic') (attention_agent.py:243-270) in torch, as oracle/restatement.critic_value."""
    tw = args["task_name"] == "vrptw"
    H = args["hidden_dim"]
    emb = RT.conv1d_k1(x[:, :, : args["input_dim"] - 1], p["Actor/Embedding/conv1d/kernel"], p["Actor/Embedding/conv1d/bias"])
    B, N = x.shape[0], x.shape[1]
    hy = torch.zeros([B, H])
    demand = x[:, :, -1]
    for i in range(args["n_process_blocks"]):
        g = lambda n: p[f"Critic/Process/P{i}/{n}"]
        d = RT.conv1d_k1(RT.conv1d_k1(demand[:, :, None], g("emb_d/kernel"), g("emb_d/bias")), g("proj_d/kernel"), g("proj_d/bias"))
        e = RT.conv1d_k1(emb, g("proj_e/kernel"), g("proj_e/bias"))
        q = hy @ g("proj_q/kernel") + g("proj_q/bias")
        arg = q[:, None, :] + e + d
        if tw:
            for col in (2, 3):   # emb_begin_tw / proj_end_tw serve both windows (vrptw_attention.py:147-150)
                arg = arg + RT.conv1d_k1(RT.conv1d_k1(x[:, :, col][:, :, None], g("emb_begin_tw/kernel"), g("emb_begin_tw/bias")),
                                         g("proj_end_tw/kernel"), g("proj_end_tw/bias"))
        u = (torch.tanh(arg) @ g("v").reshape(H, 1))[:, :, 0]
        hy = (torch.softmax(u, 1)[:, None, :] @ e)[:, 0, :]
    l1 = torch.relu(hy @ p["Critic/Linear/L1/kernel"] + p["Critic/Linear/L1/bias"])
    return (l1 @ p["Critic/Linear/L2/kernel"] + p["Critic/Linear/L2/bias"])[:, 0]


def critic_grads(params, args, data, R):
    """Gradients of tf.losses.mean_squared_error(R, v) w.r.t. every Critic variable (attention_agent.py:290, 298-299)."""
    p = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).requires_grad_(k.startswith("Critic/")) for k, v in params.items()}
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    v = critic_value(p, args, x)
    loss = torch.mean(torch.square(torch.as_tensor(R, dtype=torch.float32) - v))
    names = [k for k in p if k.startswith("Critic/")]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return float(loss.detach()), {k: (g.numpy() if g is not None else None) for k, g in zip(names, grads)}, v.detach().numpy()


def clip_adam(var, grad, m, v, step, lr, max_norm, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.clip_by_norm(grad, max_norm) (attention_agent.py:303-307) + one tf.train.AdamOptimizer update (:310-311), float32.
    step = 1 for the first update.  Returns (var, m, v)."""
    f = np.float32
    g = grad.astype(f)
    norm = np.sqrt(np.sum(g * g, dtype=f))
    if norm > max_norm:
        g = g * f(max_norm / norm)
    m = f(beta1) * m + (f(1) - f(beta1)) * g       # 1 - beta in float32, as TF's ApplyAdam kernel forms it
    v = f(beta2) * v + (f(1) - f(beta2)) * g * g
    lr_t = f(lr * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step))
    return var - lr_t * m / (np.sqrt(v) + f(eps)), m, v
