"""torch-CPU restatement of the reference's batched rollout - the CPU TIMING BASELINE.  TEST INFRASTRUCTURE ONLY.

Same graph, op for op, as ``oracle/restatement.py`` (which is the parity oracle and is pinned against fixtures produced by
the reference's own code); this twin exists because numpy's ``tanh``/``exp`` are single-threaded, which made the numpy port
a weak stand-in for "the reference's TensorFlow CPU path on all host cores" (TensorFlow itself is not installable in this
image).  torch's CPU kernels are vectorised and multi-threaded like TF's Eigen kernels; with
``torch.set_num_threads(os.cpu_count())`` this is the strongest CPU execution of the reference graph available here.
It is checked against the numpy oracle in tests/test_oracle_torch.py (identical tours up to fp32 near-ties).

Only ``bench.py`` (cpu_baseline leg, ``--impl reference``) and ``tests/`` import it.  Citations relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

BIGNUMBER = 100000.0  # shared/decode_step.py:36


def _t(p, name):
    return p[name]


def to_torch_params(params):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}


def conv1d_k1(x, kernel, bias):
    """tf.layers.Conv1D(filters, 1)."""
    return x @ kernel[0] + bias


class Env:
    """Env.reset / Env.step of VRPTW/vrptw_utils.py:199-358 and VRP/vrp_utils.py:159-255 (beam_parent gather included)."""

    def __init__(self, args, input_data):
        self.capacity = float(args["capacity"])
        self.n_nodes, self.n_cust, self.input_dim = args["n_nodes"], args["n_cust"], args["input_dim"]
        self.tw = args["task_name"] == "vrptw"
        self.input_data = input_data
        self.batch_size = input_data.shape[0]
        self.input_pnt = input_data[:, :, : self.input_dim - 1]

    def reset(self, beam_width=1):
        W = self.beam_width = beam_width
        R = self.batch_beam = self.batch_size * W
        x = self.input_data
        self.demand = x[:, :, -1].repeat(W, 1)
        self.load = torch.full([R], self.capacity)
        if self.tw:
            self.all_x, self.all_y = x[:, :, 0].repeat(W, 1), x[:, :, 1].repeat(W, 1)
            self.all_b_tw, self.all_e_tw = x[:, :, 2].repeat(W, 1), x[:, :, 3].repeat(W, 1)
            self.previous_x = x[:, self.n_nodes - 1, 0:1].repeat(W, 1)
            self.previous_y = x[:, self.n_nodes - 1, 1:2].repeat(W, 1)
            self.time = torch.zeros([R])
        self.mask = torch.cat([(self.demand == 0).float()[:, :-1], torch.ones([R, 1])], 1)

    def step(self, idx, beam_parent=None):
        R = self.batch_beam
        rows = torch.arange(R)
        if beam_parent is not None:
            src = torch.arange(self.batch_size).repeat(self.beam_width) + self.batch_size * beam_parent
            self.demand, self.load, self.mask = self.demand[src], self.load[src], self.mask[src]
            if self.tw:
                self.time, self.previous_x, self.previous_y = self.time[src], self.previous_x[src], self.previous_y[src]
                self.all_x, self.all_y = self.all_x[src], self.all_y[src]
                self.all_b_tw, self.all_e_tw = self.all_b_tw[src], self.all_e_tw[src]
        d_sat = torch.minimum(self.demand[rows, idx], self.load)
        d_scatter = torch.zeros_like(self.demand)
        d_scatter[rows, idx] = d_sat
        self.demand = self.demand - d_scatter
        self.load = self.load - d_sat
        if self.tw:
            vx, vy = self.all_x[rows, idx][:, None], self.all_y[rows, idx][:, None]
            t_spent = torch.sqrt(torch.square(self.previous_x - vx) + torch.square(self.previous_y - vy))[:, 0]
            self.previous_x, self.previous_y = vx, vy
            self.time = torch.maximum(self.time + t_spent, self.all_b_tw[rows, idx])
        dep = (idx == self.n_cust).float()
        self.load = self.load * (1 - dep) + dep * self.capacity
        if self.tw:
            self.time = self.time * (1 - dep)
        mask = torch.cat([(self.demand == 0).float()[:, :-1], torch.zeros([R, 1])], 1)
        mask = mask + torch.cat([(self.load == 0).float()[:, None].repeat(1, self.n_cust),
                                 ((self.demand.sum(1) > 0).float() * dep)[:, None]], 1)
        if self.tw:
            travel = torch.sqrt(torch.square(self.previous_x - self.all_x) + torch.square(self.previous_y - self.all_y))
            arrival = self.time[:, None] + travel
            mask = mask + torch.cat([(arrival > self.all_e_tw).float()[:, :-1], torch.zeros([R, 1])], 1)
        self.mask = mask


def attention_actor(p, prefix, query, ref, env, tw, use_tanh=False, C=10.0):
    """AttentionVRPTWActor.__call__ / AttentionVRPActor.__call__ as written (vrptw_attention.py:30-84, vrp_attention.py:27-76)."""
    g = lambda n: p[f"{prefix}/{n}"]
    demand, load = env.demand, env.load
    N = demand.shape[1]
    d = conv1d_k1(conv1d_k1(demand[:, :, None], g("emb_d/kernel"), g("emb_d/bias")), g("proj_d/kernel"), g("proj_d/bias"))
    ld = conv1d_k1(conv1d_k1((load[:, None].repeat(1, N) - demand)[:, :, None], g("emb_ld/kernel"), g("emb_ld/bias")),
                   g("proj_ld/kernel"), g("proj_ld/bias"))
    e = conv1d_k1(ref, g("proj_ref/kernel"), g("proj_ref/bias"))
    q = query @ g("proj_q/kernel") + g("proj_q/bias")
    arg = q[:, None, :].repeat(1, N, 1) + e + d + ld
    if tw:
        arg = arg + conv1d_k1(conv1d_k1(env.time[:, None, None], g("emb_time/kernel"), g("emb_time/bias")), g("proj_t/kernel"), g("proj_t/bias"))
    u = (torch.tanh(arg) @ g("v").reshape(-1, 1))[:, :, 0]
    return e, (C * torch.tanh(u) if use_tanh else u)


def lstm_cell(p, x, c, h, forget_bias):
    name = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell"
    g = torch.cat([x, h], 1) @ p[name + "/kernel"] + p[name + "/bias"]
    i, j, f, o = torch.split(g, g.shape[1] // 4, dim=1)
    new_c = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
    return new_c, torch.tanh(new_c) * torch.sigmoid(o)


def decode_step(p, args, decoder_inp, context, env, c, h):
    """RNNDecodeStep.step (shared/decode_step.py:97-128, 182-234)."""
    tw = args["task_name"] == "vrptw"
    c, h = lstm_cell(p, decoder_inp[:, 0, :], c, h, args["forget_bias"])
    hy = h
    for i in range(args["n_glimpses"]):
        ref, logit = attention_actor(p, f"Actor/Glimpse{i}", hy, context, env, tw)
        if args["mask_glimpses"]:
            logit = logit - BIGNUMBER * env.mask
        hy = (torch.softmax(logit, 1)[:, None, :] @ ref)[:, 0, :]
    _, logit = attention_actor(p, "Actor/Decoder/Attention", hy, context, env, tw, args["use_tanh"], args["tanh_exploration"])
    if args["mask_pointer"]:
        logit = logit - BIGNUMBER * env.mask
    logprob = torch.log_softmax(logit, 1)
    return logit, torch.exp(logprob), logprob, (c, h)


@torch.no_grad()
def rollout(params, args, input_data, decode_type="greedy", uniforms=None, uniforms_retry=None):
    """RLAgent.build_model(decode_type) (model/attention_agent.py:84-240) executed eagerly on the CPU.
    Returns (R [rows], idxs [T, rows]); fetches every per-step probability tensor like evaluate_batch does (they are
    materialised, not kept)."""
    p = params if isinstance(next(iter(params.values())), torch.Tensor) else to_torch_params(params)
    x = torch.from_numpy(np.ascontiguousarray(input_data, dtype=np.float32)) if not isinstance(input_data, torch.Tensor) else input_data
    env = Env(args, x)
    B, N, T = env.batch_size, env.n_nodes, args["decode_len"]
    W = args["beam_width"] if decode_type == "beam_search" else 1
    R_ = B * W
    emb = conv1d_k1(env.input_pnt, p["Actor/Embedding/conv1d/kernel"], p["Actor/Embedding/conv1d/bias"])   # embeddings.py:40-46
    env.reset(W)
    rows = torch.arange(R_)
    c, h = torch.zeros([R_, args["hidden_dim"]]), torch.zeros([R_, args["hidden_dim"]])
    decoder_input = emb[:, N - 1][:, None, :].repeat(W, 1, 1)
    context = emb.repeat(W, 1, 1)
    pnt_t = env.input_pnt.repeat(W, 1, 1)
    bbs = torch.arange(B).repeat(W)
    idxs, actions_tmp, beam_path = [], [], []
    log_beam_prev = None
    for i in range(T):
        logit, prob, logprob, (c, h) = decode_step(p, args, decoder_input, context, env, c, h)
        beam_parent = None
        if decode_type == "greedy":
            idx = torch.argmax(prob, 1)
        elif decode_type == "stochastic":
            def multinomial(u):
                cum = torch.cumsum(prob, 1)
                rand = u[:, None].repeat(1, N)
                tmp = (cum > rand).long() * torch.arange(N)[None, :] + 10000 * (rand >= cum).long()
                return tmp, torch.argmin(tmp, 1)
            u1 = torch.as_tensor(uniforms[i]) if uniforms is not None else torch.rand(R_)
            tmp, idx = multinomial(u1)
            if bool((tmp.sum(1) > 10000 * N - 1).any()):
                u2 = torch.as_tensor(uniforms_retry[i]) if uniforms_retry is not None else torch.rand(R_)
                tmp, idx = multinomial(u2)
        else:
            if i == 0:
                lbp = torch.log(prob[:B])
            else:
                lbp = torch.cat(torch.split(torch.log(prob) + log_beam_prev[:, None], B, 0), 1)
            order = torch.sort(lbp, dim=1, descending=True, stable=True)[1][:, :W]
            topv = torch.gather(lbp, 1, order).t().reshape(-1)
            topi = order.t().reshape(-1)
            idx, beam_parent = topi % N, topi // N
            prob = prob[bbs + B * beam_parent]
            beam_path.append(beam_parent)
            log_beam_prev = topv
        env.step(idx, beam_parent)
        decoder_input = context[rows, idx][:, None, :]
        _ = torch.log(prob[rows, idx])
        idxs.append(idx)
        actions_tmp.append(pnt_t[rows, idx])
    if decode_type == "beam_search":
        actions = [None] * T
        ind = rows.clone()
        for k in reversed(range(T)):
            actions[k] = actions_tmp[k][ind]
            ind = (bbs + B * beam_path[k])[ind]
    else:
        actions = actions_tmp
    sol = torch.stack(actions, 0)[:, :, :2]
    tilted = torch.cat([sol[-1:], sol[:-1]], 0)
    Rw = torch.sum(torch.pow(torch.sum(torch.pow(tilted - sol, 2), 2), 0.5), 0)        # vrptw_utils.py:397-414
    return Rw.numpy(), torch.stack(idxs).numpy()
