"""CPU restatement (numpy) of the reference's batched policy rollout.  TEST INFRASTRUCTURE ONLY.

This module is the parity oracle for the CUDA path.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg;
nothing under ``mit_rl-vrptw_b200/`` may import it.

PINNING STATUS.  The reference (jpoulletXaccount/MIT_rl-vrptw) is TensorFlow-1
graph code and ships no tests or golden vectors; TensorFlow is not installable
in this image.  The restatement is therefore pinned one level down: the
reference's own, unmodified Python files (``VRPTW/vrptw_utils.py``,
``VRP/vrp_utils.py``, ``*/…_attention.py``, ``shared/decode_step.py``,
``model/attention_agent.py``) are executed under a lazy numpy implementation
of the TF-1 ops they call (``oracle/tf1_shim``) and the resulting fixtures are
committed under ``tests/golden/`` (``tests/golden/make_golden.py``).  The
TF *kernels* themselves (Eigen exp/tanh, MatMul summation order, Glorot
initialiser, BasicLSTMCell gate order) are restated from their documented
semantics, never executed: with respect to real TensorFlow **parity is
unpinned**.

Everything is written op-for-op "as the reference writes it" (un-collapsed
Conv1D chains, float masks subtracted as ``1e5*mask``, ``log_softmax -> exp ->
argmax``) so that it can double as the reference-CPU timing baseline.  All
arithmetic is done in ``dtype`` (float32 = the reference's precision;
float64 = "truth" used to calibrate how much fp32 rounding alone changes
tours).

Citations are relative to /root/reference.
"""
from __future__ import annotations

import collections
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BIGNUMBER = 100000.0  # shared/decode_step.py:36


# --------------------------------------------------------------------------- #
# task table  (task_specific_params.py:22-115)  + the VRPTW-100 extrapolation
# --------------------------------------------------------------------------- #
Task = collections.namedtuple(
    "Task", ["task_name", "input_dim", "n_nodes", "n_cust", "decode_len", "capacity", "demand_max"])

TASKS: Dict[str, Task] = {
    "vrp10": Task("vrp", 3, 11, 10, 20, 20, 9),        # task_specific_params.py:27-34
    "vrp10_rl": Task("vrp", 3, 11, 10, 20, 30, 9),     # :37-44
    "vrp20": Task("vrp", 3, 21, 20, 30, 30, 9),        # :47-54
    "vrp20_rl": Task("vrp", 3, 21, 20, 45, 30, 9),     # :57-64
    "vrp50": Task("vrp", 3, 51, 50, 70, 40, 9),        # :68-75
    "vrp100": Task("vrp", 3, 101, 100, 140, 50, 9),    # :78-85
    "vrptw10": Task("vrptw", 5, 11, 10, 20, 20, 9),    # :90-97
    "vrptw20": Task("vrptw", 5, 21, 20, 40, 25, 9),    # :100-107
    "vrptw50": Task("vrptw", 5, 51, 50, 90, 40, 9),    # :110-117
    # NOT in the reference: capacity from vrp100, decode_len 1.8*n_cust like vrptw50.
    "vrptw100": Task("vrptw", 5, 101, 100, 180, 50, 9),
}


def default_args(task: str, **over) -> dict:
    """The subset of configs.py:28-111 the hot path reads, with embedding_graph=0."""
    t = TASKS[task]
    a = dict(task=task, task_name=t.task_name, input_dim=t.input_dim, n_nodes=t.n_nodes,
             n_cust=t.n_cust, decode_len=t.decode_len, capacity=t.capacity, demand_max=t.demand_max,
             embedding_dim=30, hidden_dim=128, rnn_layers=1, n_glimpses=0, use_tanh=False,
             tanh_exploration=10.0, mask_glimpses=True, mask_pointer=True, forget_bias=1.0,
             beam_width=5, embedding_graph=0, min_trucks=False, ups=False, random_seed=24601,
             n_process_blocks=3)
    a.update(over)
    return a


# --------------------------------------------------------------------------- #
# synthetic instances (VRP/vrp_utils.py:41-52, VRPTW/vrptw_utils.py:46-67)
# --------------------------------------------------------------------------- #
def create_dataset(task: str, n_problems: int, seed: int) -> np.ndarray:
    """float64 [n_problems, n_nodes, input_dim], same draw order as the reference generators."""
    t = TASKS[task]
    rnd = np.random.RandomState(seed)
    n = t.n_nodes
    if t.task_name == "vrp":
        x = rnd.uniform(0, 1, size=(n_problems, n, 2))
        d = rnd.randint(1, 10, [n_problems, n, 1])
        x = np.around(x, decimals=5)
        d[:, -1] = 0
        return np.concatenate([x, d], 2)
    x = rnd.uniform(0, 1, size=(n_problems, n, 2))
    x = np.around(x, decimals=5)
    d = rnd.randint(1, 10, [n_problems, n, 1])
    b_tw, e_tw, dur_min, dur_max = 0, 8, 0.5, 1
    tw_begin = rnd.uniform(b_tw + dur_max, e_tw - dur_min * 2, size=(n_problems, n, 1))
    tw_end = tw_begin + rnd.uniform(dur_min, dur_max, size=(n_problems, n, 1))
    d[:, -1] = 0
    tw_begin[:, -1] = b_tw
    tw_end[:, -1] = e_tw
    return np.concatenate([x, tw_begin, tw_end, d], axis=2)


def create_dataset_ups_old(n_cust: int, n_problems: int, seed: int) -> np.ndarray:
    """Synthetic instances in the UNITS of UPS_old/vrptw_ups_utils.py:44-71 (its lat/long Gaussian mixture is not shipped with
    the reference): latitude, longitude in radians around the depot of :44 (42.3775 N, 71.0796 W), time windows in "clicks"
    after the 892 shift (b in [0, 600), e = b + [100, 300) capped at 950), integer demand 1..9; depot [lat, lon, 0, 1000, 0]."""
    rnd = np.random.RandomState(seed)
    dep = np.array([42.3775 * np.pi / 180, -71.0796 * np.pi / 180, 0.0, 1000.0, 0.0])
    lat = dep[0] + rnd.uniform(-0.12, 0.12, size=(n_problems, n_cust)) * np.pi / 180
    lon = dep[1] + rnd.uniform(-0.16, 0.16, size=(n_problems, n_cust)) * np.pi / 180
    b = rnd.uniform(0, 600, size=(n_problems, n_cust))
    e = np.minimum(b + rnd.uniform(100, 300, size=(n_problems, n_cust)), 950.0)
    d = rnd.randint(1, 10, size=(n_problems, n_cust)).astype(np.float64)
    cust = np.stack([lat, lon, b, e, d], 2)
    return np.concatenate([cust, np.tile(dep[None, None, :], [n_problems, 1, 1])], 1)


# the constants of UPS_old/vrptw_ups_utils.py:321-333 as the float32 graph constants TF makes of the Python scalars
UPS_OLD_RADIUS = 6371                  # km
UPS_OLD_CLICKS_PER_KM = (100 / 13) * 1.2
UPS_OLD_SERVICE = 1.5                  # clicks per package


def _ups_old_km(ax, ay, bx, by):
    """6371 * sqrt(square((by - ay) * cos(0.5 * (bx + ax))) + square(bx - ax)): the equirectangular distance of
    UPS_old/vrptw_ups_utils.py:321-322, 359-360, 421-423 (x = latitude, y = longitude, radians), every op rounded separately."""
    dt = np.asarray(ax).dtype.type
    interm = (by - ay) * np.cos(dt(0.5) * (bx + ax))
    return dt(UPS_OLD_RADIUS) * np.sqrt(np.square(interm) + np.square(bx - ax))


# --------------------------------------------------------------------------- #
# parameters.  Names are the TF-1 variable names.
# --------------------------------------------------------------------------- #
def _glorot(rnd, fan_in, fan_out, shape):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rnd.uniform(-lim, lim, size=shape).astype(np.float32)


def attention_param_names(prefix: str, tw: bool) -> List[str]:
    names = [prefix + "/v"]
    for blk in ["emb_d", "emb_ld"] + (["emb_time"] if tw else []) + ["proj_d", "proj_ld"] + \
               (["proj_t"] if tw else []) + ["proj_q", "proj_ref"]:
        names += [f"{prefix}/{blk}/kernel", f"{prefix}/{blk}/bias"]
    return names


def init_params(args: dict, seed: int = 1234, bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Glorot-uniform kernels, zero biases (tf.layers defaults), Glorot ``v``.

    ``bias_scale > 0`` draws U(-s, s) biases instead, so that tests exercise every bias
    term (reference-init biases are all zero).
    """
    rnd = np.random.RandomState(seed)
    tw = args["task_name"] == "vrptw"
    D1, E, H = args["input_dim"] - 1, args["embedding_dim"], args["hidden_dim"]
    p: Dict[str, np.ndarray] = {}

    def bias(n):
        if bias_scale == 0.0:
            return np.zeros([n], np.float32)
        return rnd.uniform(-bias_scale, bias_scale, size=[n]).astype(np.float32)

    p["Actor/Embedding/conv1d/kernel"] = _glorot(rnd, D1, E, [1, D1, E])      # shared/embeddings.py:37-38
    p["Actor/Embedding/conv1d/bias"] = bias(E)
    lstm = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell"     # shared/decode_step.py:175-180
    p[lstm + "/kernel"] = _glorot(rnd, E + H, 4 * H, [E + H, 4 * H])
    p[lstm + "/bias"] = bias(4 * H)
    prefixes = [f"Actor/Glimpse{i}" for i in range(args["n_glimpses"])] + ["Actor/Decoder/Attention"]
    for pre in prefixes:                                                      # VRPTW/vrptw_attention.py:10-24
        p[pre + "/v"] = _glorot(rnd, 1, H, [1, H])
        for blk in ["emb_d", "emb_ld"] + (["emb_time"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = _glorot(rnd, 1, H, [1, 1, H])
            p[f"{pre}/{blk}/bias"] = bias(H)
        for blk in ["proj_d", "proj_ld"] + (["proj_t"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = _glorot(rnd, H, H, [1, H, H])
            p[f"{pre}/{blk}/bias"] = bias(H)
        p[pre + "/proj_q/kernel"] = _glorot(rnd, H, H, [H, H])
        p[pre + "/proj_q/bias"] = bias(H)
        p[pre + "/proj_ref/kernel"] = _glorot(rnd, E, H, [1, E, H])
        p[pre + "/proj_ref/bias"] = bias(H)
    return p


# --------------------------------------------------------------------------- #
# elementwise helpers in a fixed dtype
# --------------------------------------------------------------------------- #
def _sigmoid(x):
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


def log_softmax(logit):
    """tf.nn.log_softmax: logits - max - log(sum(exp(logits - max)))  (shared/decode_step.py:125)."""
    m = logit.max(axis=1, keepdims=True)
    sh = logit - m
    return sh - np.log(np.exp(sh).sum(axis=1, keepdims=True, dtype=logit.dtype))


def softmax(logit):
    return np.exp(log_softmax(logit))


def conv1d_k1(x, kernel, bias):
    """tf.layers.Conv1D(filters, 1): x[..., in] @ kernel[0] + bias."""
    return (x @ kernel[0] + bias).astype(x.dtype, copy=False)


# --------------------------------------------------------------------------- #
# environment  (VRPTW/vrptw_utils.py:161-358, VRP/vrp_utils.py:133-255)
# --------------------------------------------------------------------------- #
State = collections.namedtuple("State", ["load", "time", "demand", "d_sat", "mask"])


class EnvOracle:
    """Vectorised restatement of ``Env.reset`` / ``Env.step`` (VRP and VRPTW in one class)."""

    def __init__(self, args: dict, input_data: np.ndarray, dtype=np.float32):
        self.dtype = dtype
        self.capacity = dtype(args["capacity"])
        self.n_nodes = args["n_nodes"]
        self.n_cust = args["n_cust"]
        self.input_dim = args["input_dim"]
        self.tw = args["task_name"] == "vrptw"
        self.ups_old = bool(args.get("ups_old", False))            # UPS_old/vrptw_ups_utils.py physics (VRPTW only)
        self.input_data = np.asarray(input_data).astype(dtype)     # tf.float32 placeholder, vrptw_utils.py:176
        self.batch_size = self.input_data.shape[0]
        self.input_pnt = self.input_data[:, :, : self.input_dim - 1]

    def reset(self, beam_width: int = 1) -> State:                  # vrptw_utils.py:199-252
        dt = self.dtype
        W = self.beam_width = beam_width
        R = self.batch_beam = self.batch_size * W
        self.demand = np.tile(self.input_data[:, :, -1], [W, 1])
        self.load = np.ones([R], dt) * self.capacity
        if self.tw:
            self.all_x = np.tile(self.input_data[:, :, 0], [W, 1])
            self.all_y = np.tile(self.input_data[:, :, 1], [W, 1])
            self.all_b_tw = np.tile(self.input_data[:, :, 2], [W, 1])
            self.all_e_tw = np.tile(self.input_data[:, :, 3], [W, 1])
            self.previous_x = np.tile(self.input_data[:, self.n_nodes - 1, 0][:, None], [W, 1])
            self.previous_y = np.tile(self.input_data[:, self.n_nodes - 1, 1][:, None], [W, 1])
            self.time = np.zeros([R], dt)
        else:
            self.time = None
        self.mask = np.concatenate([(self.demand == 0).astype(dt)[:, :-1], np.ones([R, 1], dt)], 1)
        return State(self.load, self.time, self.demand, np.zeros([R, self.n_nodes], dt), self.mask)

    def step(self, idx: np.ndarray, beam_parent: Optional[np.ndarray] = None) -> State:
        """idx, beam_parent: int64 [R, 1]  (vrptw_utils.py:254-358)."""
        dt = self.dtype
        R = self.batch_beam
        idx = np.asarray(idx).reshape(R).astype(np.int64)
        if beam_parent is not None:                                 # :262-287
            src = np.tile(np.arange(self.batch_size, dtype=np.int64), self.beam_width) \
                + self.batch_size * np.asarray(beam_parent).reshape(R).astype(np.int64)
            self.demand = self.demand[src]
            self.load = self.load[src]
            self.mask = self.mask[src]
            if self.tw:
                self.time = self.time[src]
                self.previous_x = self.previous_x[src]
                self.previous_y = self.previous_y[src]
                self.all_x = self.all_x[src]
                self.all_y = self.all_y[src]
                self.all_b_tw = self.all_b_tw[src]
                self.all_e_tw = self.all_e_tw[src]
        rows = np.arange(R)
        d_sat = np.minimum(self.demand[rows, idx], self.load)       # :294
        d_scatter = np.zeros_like(self.demand)
        d_scatter[rows, idx] = d_sat
        self.demand = self.demand - d_scatter                       # :297-298
        self.load = self.load - d_sat                               # :301
        if self.tw:
            vx = self.all_x[rows, idx][:, None]
            vy = self.all_y[rows, idx][:, None]
            if self.ups_old:                                        # UPS_old/vrptw_ups_utils.py:321-333
                t_spent = (dt(UPS_OLD_CLICKS_PER_KM) * _ups_old_km(self.previous_x, self.previous_y, vx, vy))[:, 0]
            else:
                t_spent = np.sqrt(np.square(self.previous_x - vx) + np.square(self.previous_y - vy))[:, 0]  # :304-309
            self.previous_x, self.previous_y = vx, vy
            self.time = np.maximum(self.time + t_spent, self.all_b_tw[rows, idx])   # :316
            if self.ups_old:
                self.time = self.time + dt(UPS_OLD_SERVICE) * d_sat  # service time, UPS_old :332-333
        dep = (idx == self.n_cust).astype(dt)                       # :319
        self.load = self.load * (1 - dep) + dep * self.capacity     # :322
        if self.tw:
            self.time = self.time * (1 - dep)                       # :325
        self.mask = np.concatenate([(self.demand == 0).astype(dt)[:, :-1], np.zeros([R, 1], dt)], 1)  # :328-329
        self.mask = self.mask + np.concatenate(                     # :334-337
            [np.tile((self.load == 0).astype(dt)[:, None], [1, self.n_cust]),
             ((self.demand.sum(1, dtype=dt) > 0).astype(dt) * dep)[:, None]], 1)
        if self.tw:                                                 # :340-348
            if self.ups_old:                                        # UPS_old :359-361 (previous - all; squares and cos are even)
                travel = dt(UPS_OLD_CLICKS_PER_KM) * _ups_old_km(self.all_x, self.all_y, self.previous_x, self.previous_y)
            else:
                travel = np.sqrt(np.square(self.previous_x - self.all_x) + np.square(self.previous_y - self.all_y))
            arrival = self.time[:, None] + travel
            self.mask = self.mask + np.concatenate(
                [(arrival > self.all_e_tw).astype(dt)[:, :-1], np.zeros([R, 1], dt)], 1)
        return State(self.load, self.time, self.demand, d_sat, self.mask)


def reward_func(actions: Sequence[np.ndarray], tw: bool, ups_old: bool = False) -> np.ndarray:
    """Closed tour length, ``depot=None`` branch (vrptw_utils.py:397-414, vrp_utils.py:293-307); ups_old: in km over the
    equirectangular distance (UPS_old/vrptw_ups_utils.py:416-424)."""
    sol = np.stack(actions, 0)
    if tw:
        sol = sol[:, :, :2]
    tilted = np.concatenate([sol[-1:], sol[:-1]], 0)
    dt = sol.dtype.type
    if ups_old:
        return np.sum(_ups_old_km(sol[:, :, 0], sol[:, :, 1], tilted[:, :, 0], tilted[:, :, 1]), 0)
    return np.sum(np.power(np.sum(np.power(tilted - sol, dt(2)), 2), dt(0.5)), 0)


def reward_func_min_trucks(actions: Sequence[np.ndarray], depot: np.ndarray, decode_len: int, n_nodes: float, tw: bool) -> np.ndarray:
    """``reward_func(sample_solution, decode_len, n_nodes, depot)`` with a depot: 70/n_nodes x (returns to the depot,
    consecutive stays counted once) + 30 x route length / sum of the distances to the depot
    (VRPTW/vrptw_utils.py:384-395, 410-412; VRP/vrp_utils.py:275-285, 302-304).

    As written: the depot test compares COLUMN 0 only (``tf.equal(...)[:, 0]``); the normalising length sums over ALL
    action columns (for VRPTW: x, y, b_tw, e_tw - it is computed before the slice to x, y)."""
    dt = actions[0].dtype.type
    counter = np.zeros_like(depot)[:, 0]
    visits = (actions[0] == depot).astype(dt)[:, 0]
    for i in range(1, len(actions)):
        interm = (actions[i] == depot).astype(dt)[:, 0]
        counter = counter * interm + interm
        visits = visits + interm * (counter < 1.5).astype(dt)
    max_length = np.stack([depot for _ in range(decode_len)], 0)
    sol = np.stack(actions, 0)
    max_lens = np.sum(np.power(np.sum(np.power(max_length - sol, dt(2)), 2), dt(0.5)), 0)
    if tw:
        sol = sol[:, :, :2]
    tilted = np.concatenate([sol[-1:], sol[:-1]], 0)
    route = np.sum(np.power(np.sum(np.power(tilted - sol, dt(2)), 2), dt(0.5)), 0)
    return dt(70.0) * (dt(1.0 / n_nodes) * visits) + dt(30.0) * (route / max_lens)


def reward_func_min_trucks_ups(actions: Sequence[np.ndarray], depot: np.ndarray) -> np.ndarray:
    """``reward_func`` with a depot of the UPS VRPTW twin (UPS/vrptw_ups_utils.py:387-419): 100 x (returns to the depot,
    consecutive stays counted once, column 0 compared) + the closed route length over x, y.  The great-circle distance to the
    depot (:399-402) is computed by the reference but never used in the returned value."""
    dt = actions[0].dtype.type
    counter = np.zeros_like(depot)[:, 0]
    visits = (actions[0] == depot).astype(dt)[:, 0]
    for i in range(1, len(actions)):
        interm = (actions[i] == depot).astype(dt)[:, 0]
        counter = counter * interm + interm
        visits = visits + interm * (counter < 1.5).astype(dt)
    sol = np.stack(actions, 0)[:, :, :2]
    tilted = np.concatenate([sol[-1:], sol[:-1]], 0)
    route = np.sum(np.power(np.sum(np.power(tilted - sol, dt(2)), 2), dt(0.5)), 0)
    return dt(100.0) * visits + route


def reward_func_min_trucks_ups_old(actions: Sequence[np.ndarray], depot: np.ndarray, decode_len: int, n_nodes: float) -> np.ndarray:
    """``reward_func`` with a depot of the UPS_old package (UPS_old/vrptw_ups_utils.py:400-428): 70/n_nodes x (number of
    steps at the depot: column 0 compared, EVERY step counts - this version has no "stay counts once" counter) + 30 x the
    closed route length in equirectangular km (:421-424) / the sum of the EUCLIDEAN distances to the depot over all action
    columns (:406-408; a km numerator over a coordinate-unit denominator, as written)."""
    dt = actions[0].dtype.type
    visits = (actions[0] == depot).astype(dt)[:, 0]
    for i in range(1, len(actions)):
        visits = visits + (actions[i] == depot).astype(dt)[:, 0]
    max_length = np.stack([depot for _ in range(decode_len)], 0)
    sol = np.stack(actions, 0)
    max_lens = np.sum(np.power(np.sum(np.power(max_length - sol, dt(2)), 2), dt(0.5)), 0)
    sol = sol[:, :, :2]
    tilted = np.concatenate([sol[-1:], sol[:-1]], 0)
    route = np.sum(_ups_old_km(sol[:, :, 0], sol[:, :, 1], tilted[:, :, 0], tilted[:, :, 1]), 0)
    return dt(70.0) * (dt(1.0 / n_nodes) * visits) + dt(30.0) * (route / max_lens)


# --------------------------------------------------------------------------- #
# model pieces
# --------------------------------------------------------------------------- #
def linear_embedding(p, input_pnt):
    """shared/embeddings.py:40-46."""
    dt = input_pnt.dtype
    return conv1d_k1(input_pnt, p["Actor/Embedding/conv1d/kernel"].astype(dt),
                     p["Actor/Embedding/conv1d/bias"].astype(dt))


def attention_actor(p, prefix, query, ref, env, tw, use_tanh=False, C=10.0):
    """``AttentionVRPTWActor.__call__`` / ``AttentionVRPActor.__call__`` as written.

    VRPTW/vrptw_attention.py:30-84, VRP/vrp_attention.py:27-76.  Returns (e, logits).
    """
    dt = query.dtype
    g = lambda n: p[f"{prefix}/{n}"].astype(dt)
    demand, load = env.demand, env.load
    N = demand.shape[1]
    emb_d = conv1d_k1(demand[:, :, None], g("emb_d/kernel"), g("emb_d/bias"))
    d = conv1d_k1(emb_d, g("proj_d/kernel"), g("proj_d/bias"))
    emb_ld = conv1d_k1((np.tile(load[:, None], [1, N]) - demand)[:, :, None], g("emb_ld/kernel"), g("emb_ld/bias"))
    ld = conv1d_k1(emb_ld, g("proj_ld/kernel"), g("proj_ld/bias"))
    e = conv1d_k1(ref, g("proj_ref/kernel"), g("proj_ref/bias"))
    q = (query @ g("proj_q/kernel") + g("proj_q/bias")).astype(dt, copy=False)
    expanded_q = np.tile(q[:, None, :], [1, N, 1])
    if tw:
        emb_time = conv1d_k1(env.time[:, None, None], g("emb_time/kernel"), g("emb_time/bias"))
        t = conv1d_k1(emb_time, g("proj_t/kernel"), g("proj_t/bias"))
        arg = expanded_q + e + d + ld + t                           # vrptw_attention.py:77
    else:
        arg = expanded_q + e + d + ld                               # vrp_attention.py:69
    v = g("v").reshape(-1, 1)
    u = (np.tanh(arg) @ v)[:, :, 0].astype(dt, copy=False)
    logits = dt.type(C) * np.tanh(u) if use_tanh else u
    return e, logits


def generic_attention(kernel_q, bias_q, kernel_ref, bias_ref, v, query, ref, use_tanh=False, C=10.0):
    """shared/attention.py:19-49 (not wired into main.py, kept for the drop-in surface)."""
    dt = query.dtype
    e = conv1d_k1(ref, kernel_ref.astype(dt), bias_ref.astype(dt))
    q = query @ kernel_q.astype(dt) + bias_q.astype(dt)
    u = (np.tanh(q[:, None, :] + e) @ v.astype(dt).reshape(-1, 1))[:, :, 0]
    return e, (dt.type(C) * np.tanh(u) if use_tanh else u)


def lstm_cell(p, x, c, h, forget_bias):
    """One BasicLSTMCell step (TF-1 rnn_cell_impl): gates i, j, f, o."""
    dt = x.dtype
    name = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell"
    g = np.concatenate([x, h], 1) @ p[name + "/kernel"].astype(dt) + p[name + "/bias"].astype(dt)
    i, j, f, o = np.split(g.astype(dt, copy=False), 4, axis=1)
    new_c = c * _sigmoid(f + dt.type(forget_bias)) + _sigmoid(i) * np.tanh(j)
    new_h = np.tanh(new_c) * _sigmoid(o)
    return new_c, new_h


def decode_step(p, args, decoder_inp, context, env, c, h):
    """``RNNDecodeStep.step`` (shared/decode_step.py:97-128, 182-234).

    decoder_inp [R,1,E]; returns logit, prob, logprob, (c, h).
    """
    tw = args["task_name"] == "vrptw"
    c, h = lstm_cell(p, decoder_inp[:, 0, :], c, h, args["forget_bias"])
    hy = h
    dt = hy.dtype
    for i in range(args["n_glimpses"]):                             # :218-227
        ref, logit = attention_actor(p, f"Actor/Glimpse{i}", hy, context, env, tw, use_tanh=False)
        if args["mask_glimpses"]:
            logit = logit - dt.type(BIGNUMBER) * env.mask
        prob = softmax(logit)
        hy = (prob[:, None, :] @ ref)[:, 0, :]
    _, logit = attention_actor(p, "Actor/Decoder/Attention", hy, context, env, tw,
                               use_tanh=args["use_tanh"], C=args["tanh_exploration"])
    if args["mask_pointer"]:
        logit = logit - dt.type(BIGNUMBER) * env.mask               # :231-232
    logprob = log_softmax(logit)
    prob = np.exp(logprob)
    return logit, prob, logprob, (c, h)


# --------------------------------------------------------------------------- #
# critic + REINFORCE losses (model/attention_agent.py:243-270, 286-294; critics
# VRPTW/vrptw_attention.py:87-169, VRP/vrp_attention.py:79-140)
# --------------------------------------------------------------------------- #
def init_critic_params(args: dict, seed: int = 4321, bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Variables of the critic as the reference creates them: ``Critic/Process/P<i>/{v, emb_d, proj_d, proj_q, proj_e}`` (+
    ``emb_begin_tw``, ``proj_end_tw`` for VRPTW - the only two time-window layers the reference ever CALLS, each for both
    windows, vrptw_attention.py:147-150; ``emb_end_tw`` / ``proj_begin_tw`` are constructed but never built), then
    ``Critic/Linear/{L1, L2}``.  Glorot-uniform kernels, zero (or U(-s, s)) biases."""
    rnd = np.random.RandomState(seed)
    tw = args["task_name"] == "vrptw"
    E, H = args["embedding_dim"], args["hidden_dim"]
    p: Dict[str, np.ndarray] = {}
    bias = (lambda n: np.zeros([n], np.float32)) if bias_scale == 0.0 else \
        (lambda n: rnd.uniform(-bias_scale, bias_scale, size=[n]).astype(np.float32))
    for i in range(args["n_process_blocks"]):
        pre = f"Critic/Process/P{i}"
        p[pre + "/v"] = _glorot(rnd, 1, H, [1, H])
        for blk in ["emb_d"] + (["emb_begin_tw"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = _glorot(rnd, 1, H, [1, 1, H])
            p[f"{pre}/{blk}/bias"] = bias(H)
        for blk in ["proj_d"] + (["proj_end_tw"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = _glorot(rnd, H, H, [1, H, H])
            p[f"{pre}/{blk}/bias"] = bias(H)
        p[pre + "/proj_q/kernel"] = _glorot(rnd, H, H, [H, H])
        p[pre + "/proj_q/bias"] = bias(H)
        p[pre + "/proj_e/kernel"] = _glorot(rnd, E, H, [1, E, H])
        p[pre + "/proj_e/bias"] = bias(H)
    p["Critic/Linear/L1/kernel"] = _glorot(rnd, H, H, [H, H])
    p["Critic/Linear/L1/bias"] = bias(H)
    p["Critic/Linear/L2/kernel"] = _glorot(rnd, H, 1, [H, 1])
    p["Critic/Linear/L2/bias"] = bias(1)
    return p


def critic_value(p, args, input_data, dtype=np.float32) -> np.ndarray:
    """``v`` of ``RLAgent.build_model('stochastic')`` (attention_agent.py:243-270): hy = 0; n_process_blocks times
    ``(e, logit) = AttentionCritic(hy, emb, env)``, ``hy = softmax(logit) . e`` (no mask); ``v = L2(relu(L1(hy)))``.
    The critic reads the INITIAL demand and the time windows of ``env.input_data`` (vrptw_attention.py:135-137) and the
    actor's embedding.  p holds the actor AND the critic variables."""
    x = np.asarray(input_data).astype(dtype)
    tw = args["task_name"] == "vrptw"
    H = args["hidden_dim"]
    emb = linear_embedding(p, x[:, :, : args["input_dim"] - 1])
    B, N = x.shape[0], x.shape[1]
    hy = np.zeros([B, H], dtype)
    demand = x[:, :, -1]
    for i in range(args["n_process_blocks"]):
        g = lambda n: p[f"Critic/Process/P{i}/{n}"].astype(dtype)
        d = conv1d_k1(conv1d_k1(demand[:, :, None], g("emb_d/kernel"), g("emb_d/bias")), g("proj_d/kernel"), g("proj_d/bias"))
        e = conv1d_k1(emb, g("proj_e/kernel"), g("proj_e/bias"))
        q = (hy @ g("proj_q/kernel") + g("proj_q/bias")).astype(dtype, copy=False)
        arg = np.tile(q[:, None, :], [1, N, 1]) + e + d
        if tw:   # emb_begin_tw embeds BOTH windows, proj_end_tw projects BOTH (vrptw_attention.py:147-150)
            emb_end = conv1d_k1(x[:, :, 3][:, :, None], g("emb_begin_tw/kernel"), g("emb_begin_tw/bias"))
            emb_begin = conv1d_k1(x[:, :, 2][:, :, None], g("emb_begin_tw/kernel"), g("emb_begin_tw/bias"))
            p_end = conv1d_k1(emb_end, g("proj_end_tw/kernel"), g("proj_end_tw/bias"))
            p_begin = conv1d_k1(emb_begin, g("proj_end_tw/kernel"), g("proj_end_tw/bias"))
            arg = arg + p_begin + p_end                             # :163
        u = (np.tanh(arg) @ g("v").reshape(H, 1))[:, :, 0]
        prob = softmax(u)
        hy = (prob[:, None, :] @ e)[:, 0, :]
    l1 = np.maximum(hy @ p["Critic/Linear/L1/kernel"].astype(dtype) + p["Critic/Linear/L1/bias"].astype(dtype), dtype(0))
    return (l1 @ p["Critic/Linear/L2/kernel"].astype(dtype) + p["Critic/Linear/L2/bias"].astype(dtype))[:, 0].astype(dtype)


def reinforce_losses(R, v, logprobs):
    """``build_train_step`` (attention_agent.py:286-294): actor_loss = mean((R - stop_grad(v)) * add_n(logprobs)),
    critic_loss = tf.losses.mean_squared_error(R, v).  logprobs [T, B]."""
    dt = R.dtype
    slp = np.zeros_like(R)
    for lp in logprobs:                                             # tf.add_n: in list order
        slp = slp + lp.astype(dt)
    actor = np.mean((R - v) * slp, 0, dtype=dt)
    critic = np.mean(np.square(R - v), dtype=dt)                    # mean_squared_error: SUM_BY_NONZERO_WEIGHTS = mean
    return actor, critic


# --------------------------------------------------------------------------- #
# the rollout  (model/attention_agent.py:84-240)
# --------------------------------------------------------------------------- #
@dataclass
class RolloutResult:
    R: np.ndarray                  # [rows] tour length (per beam row)
    idxs: np.ndarray               # [T, rows] int64
    logprobs: np.ndarray           # [T, rows]
    actions: np.ndarray            # [T, rows, D-1]  (back-tracked for beam search)
    beam_parents: Optional[np.ndarray] = None   # [T, rows] int64 (beam only)
    probs: Optional[np.ndarray] = None          # [T, rows, N]
    logits: Optional[np.ndarray] = None         # [T, rows, N]
    masks: Optional[np.ndarray] = None          # [T, rows, N]  mask seen by step t
    final_state: Optional[State] = None
    retried: List[int] = field(default_factory=list)   # steps at which the whole batch was redrawn


def multinomial_as_written(prob, u, n_nodes):
    """``my_multinomial`` (attention_agent.py:151-161) for one set of uniforms u [R]."""
    cum = np.cumsum(prob, 1, dtype=prob.dtype)
    rand = np.tile(u[:, None].astype(prob.dtype), [1, n_nodes])
    sorted_ind = np.tile(np.arange(n_nodes, dtype=np.int64)[None, :], [prob.shape[0], 1])
    tmp = (cum > rand).astype(np.int64) * sorted_ind + 10000 * (rand >= cum).astype(np.int64)
    return tmp, np.argmin(tmp, 1)


def rollout(p, args, input_data, decode_type="greedy", uniforms=None, uniforms_retry=None,
            dtype=np.float32, keep_probs=False, per_row_retry=False) -> RolloutResult:
    """``RLAgent.build_model(decode_type)`` executed eagerly on ``input_data`` [B,N,D].

    ``uniforms`` / ``uniforms_retry``: [T, R] numbers in [0,1) standing in for the two
    ``tf.random_uniform`` draws of ``my_multinomial`` and its single retry
    (attention_agent.py:163-167).  ``per_row_retry=True`` applies the retry only to the
    failing rows (what the CUDA kernel does by default).
    """
    env = EnvOracle(args, input_data, dtype)
    tw = env.tw
    B, N, T = env.batch_size, env.n_nodes, args["decode_len"]
    W = args["beam_width"] if decode_type == "beam_search" else 1
    R_ = B * W
    input_pnt = env.input_pnt
    emb = linear_embedding(p, input_pnt)                            # attention_agent.py:96
    env.reset(W)                                                    # :110
    rows = np.arange(R_)
    c = np.zeros([R_, args["hidden_dim"]], dtype)
    h = np.zeros([R_, args["hidden_dim"]], dtype)
    decoder_input = np.tile(emb[:, N - 1][:, None, :], [W, 1, 1])   # :133-134
    context = np.tile(emb, [W, 1, 1])                               # :137
    emb_t = context
    pnt_t = np.tile(input_pnt, [W, 1, 1])
    batchBeamSeq = np.tile(np.arange(B, dtype=np.int64), W)
    idxs, logprobs, actions_tmp, probs, logits, masks, beam_path, retried = [], [], [], [], [], [], [], []
    log_beam_prev = None
    for i in range(T):
        if keep_probs:
            masks.append(env.mask.copy())
        logit, prob, logprob, (c, h) = decode_step(p, args, decoder_input, context, env, c, h)
        beam_parent = None
        if decode_type == "greedy":
            idx = np.argmax(prob, 1)                                # :147
        elif decode_type == "stochastic":                           # :148-167
            tmp, idx = multinomial_as_written(prob, uniforms[i], N)
            fail = tmp.sum(1) > (10000 * N) - 1
            if fail.any():
                retried.append(i)
                tmp2, idx2 = multinomial_as_written(prob, uniforms_retry[i], N)
                idx = np.where(fail, idx2, idx) if per_row_retry else idx2
        elif decode_type == "beam_search":                          # :169-205
            with np.errstate(divide="ignore"):
                if i == 0:
                    log_beam_prob = np.log(prob[:B])
                else:
                    lbp = np.log(prob) + log_beam_prev[:, None]
                    log_beam_prob = np.concatenate(np.split(lbp, W, axis=0), 1)
            order = np.argsort(-log_beam_prob, axis=1, kind="stable")[:, :W]   # tf.nn.top_k: ties -> lower index
            topv = np.take_along_axis(log_beam_prob, order, 1)
            topv = topv.T.reshape(-1)                               # row = w*B + b
            topi = order.T.reshape(-1).astype(np.int64)
            idx = topi % N
            beam_parent = topi // N
            prob = prob[batchBeamSeq + B * beam_parent]
            beam_path.append(beam_parent)
            log_beam_prev = topv
        else:
            raise ValueError(decode_type)
        env.step(idx[:, None], None if beam_parent is None else beam_parent[:, None])   # :207
        decoder_input = emb_t[rows, idx][:, None, :]                # :211-212
        with np.errstate(divide="ignore"):
            logprobs.append(np.log(prob[rows, idx]))                # :214
        idxs.append(idx.astype(np.int64))
        actions_tmp.append(pnt_t[rows, idx])                        # :219
        if keep_probs:
            probs.append(prob)
            logits.append(logit)
    if decode_type == "beam_search":                                # :222-231
        actions = [None] * T
        ind = rows.copy()
        for k in reversed(range(T)):
            actions[k] = actions_tmp[k][ind]
            ind = (batchBeamSeq + B * beam_path[k])[ind]
    else:
        actions = actions_tmp
    Rw = reward_func(actions, tw, bool(args.get("ups_old", False)))   # :236-240
    return RolloutResult(
        R=Rw, idxs=np.stack(idxs), logprobs=np.stack(logprobs), actions=np.stack(actions),
        beam_parents=np.stack(beam_path) if beam_path else None,
        probs=np.stack(probs) if keep_probs else None,
        logits=np.stack(logits) if keep_probs else None,
        masks=np.stack(masks) if keep_probs else None,
        final_state=State(env.load, env.time, env.demand, None, env.mask), retried=retried)


def best_of_beams(R, B, W):
    """evaluate_batch's min over beams (attention_agent.py:512-513)."""
    return np.amin(np.concatenate(np.split(R[:, None], W, axis=0), 1), 1)
