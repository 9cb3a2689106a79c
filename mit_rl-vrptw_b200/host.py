"""Host-side mirror of the reference's Python protocol for the rollout path (drop-in surface).

Same class names, constructor arguments, attributes and return values as the reference, so that the
code in main.py:17-53 / model/attention_agent.py that wires them together keeps working:

    reference                                             here
    ---------------------------------------------------   -------------------------------------------
    shared/embeddings.py        LinearEmbedding           LinearEmbedding
    VRP/vrp_attention.py        AttentionVRPActor         AttentionVRPActor
    VRPTW/vrptw_attention.py    AttentionVRPTWActor       AttentionVRPTWActor
    UPS/*_attention.py          Attention*_UPS_Actor      AttentionVRP_UPS_Actor, AttentionVRPTW_UPS_Actor
    shared/decode_step.py       RNNDecodeStep             RNNDecodeStep
    VRP|VRPTW|UPS/*_utils.py    Env, State, reward_func,  VRPEnv / VRPTWEnv (``Env`` via the loader), State*,
                                DataGenerator             reward_func, DataGenerator
    model/attention_agent.py    RLAgent                   RLAgent (build_model, evaluate_batch, inference, ...)
    main.py                     load_task_specific_components  load_task_specific_components

TensorFlow placeholders + feed_dict become eager CUDA tensors: assign ``env.input_data = tensor``.
TF variables become entries of a ``VariableStore`` keyed by the TF-1 variable names (what
tf.train.Saver writes), so a converted checkpoint loads by name.  All arithmetic is done by
libvrpb200.so through ``Engine``; torch provides device memory, indexing and the beam bookkeeping of the
step-wise path.  Nothing here computes the model on the CPU.
"""
from __future__ import annotations

import collections
import os
import time
import warnings
from typing import Dict, Optional

import numpy as np

from . import data as _data
from .task_specific_params import task_lst  # noqa: F401  (re-export, reference: task_specific_params.py)

# --------------------------------------------------------------------------------------------------
# variables
# --------------------------------------------------------------------------------------------------


class VariableStore:
    """name -> float32 array; the stand-in for TF's global variable collection."""

    _uids = iter(range(1, 1 << 62))

    def __init__(self, seed: int = 1234):
        self.uid = next(VariableStore._uids)    # id() values are recycled after garbage collection; this is not
        self.vars: Dict[str, np.ndarray] = collections.OrderedDict()
        self.rnd = np.random.RandomState(seed)
        self.version = 0

    def get(self, name, shape, init="glorot", fans=None):
        """tf.get_variable / tf.layers semantics: create on first request (Glorot-uniform kernels, zero biases)."""
        if name in self.vars:
            if tuple(self.vars[name].shape) != tuple(shape):
                raise ValueError(f"variable {name}: shape {self.vars[name].shape} != requested {tuple(shape)}")
            return self.vars[name]
        if init == "zeros":
            v = np.zeros(shape, np.float32)
        else:
            fi, fo = fans
            lim = np.sqrt(6.0 / (fi + fo))
            v = self.rnd.uniform(-lim, lim, size=shape).astype(np.float32)
        self.vars[name] = v
        self.version += 1
        return v

    def assign(self, params: Dict[str, np.ndarray], strict: bool = True):
        """saver.restore: overwrite by name; unknown names are an error when strict."""
        for k, v in params.items():
            v = np.asarray(v, np.float32)
            if k in self.vars and self.vars[k].shape != v.shape:
                raise ValueError(f"variable {k}: checkpoint shape {v.shape} != {self.vars[k].shape}")
            if k not in self.vars and strict:
                raise KeyError(f"checkpoint variable {k} is not part of the model")
            self.vars[k] = v.copy()
        self.version += 1

    def save(self, path):
        np.savez(path, **self.vars)

    def load(self, path, strict=True):
        with np.load(path) as z:
            self.assign({k: z[k] for k in z.files}, strict=strict)


_default_store = VariableStore()


def default_store() -> VariableStore:
    return _default_store


def reset_default_store(seed: int = 1234) -> VariableStore:
    """tf.reset_default_graph() for the variable collection."""
    global _default_store
    _default_store = VariableStore(seed)
    return _default_store


# --------------------------------------------------------------------------------------------------
# engines: one C-ABI handle per (configuration, weights version)
# --------------------------------------------------------------------------------------------------
_engine_cache: Dict[tuple, object] = {}


def _get_engine(args: dict, params, version_key, device: int = 0):
    """params: the variable dict, or a callable that builds it (only called when the engine is not cached yet)."""
    from .engine import Engine
    key = (tuple(sorted((k, str(v)) for k, v in args.items() if k in (
        "task_name", "n_nodes", "n_cust", "input_dim", "embedding_dim", "hidden_dim", "decode_len", "n_glimpses",
        "use_tanh", "mask_glimpses", "mask_pointer", "capacity", "tanh_exploration", "forget_bias", "ups_old"))), version_key, device)
    eng = _engine_cache.get(key)
    if eng is None:
        if len(_engine_cache) > 8:
            _engine_cache.clear()
        eng = Engine(args, device=device)
        eng.set_weights(params() if callable(params) else params)
        _engine_cache[key] = eng
    return eng


def _dummy_model_params(args: dict) -> Dict[str, np.ndarray]:
    """Zero tensors for every variable the C ABI insists on; callers overwrite the ones they own."""
    p = _data.glorot_params(args, seed=0)
    return {k: np.zeros_like(v) for k, v in p.items()}


# --------------------------------------------------------------------------------------------------
# model pieces
# --------------------------------------------------------------------------------------------------
class Embedding(object):
    """shared/embeddings.py:9-22"""

    def __init__(self, emb_type, embedding_dim):
        self.emb_type = emb_type
        self.embedding_dim = embedding_dim
        self.total_time = 0

    def __call__(self, input_pnt):
        pass


class LinearEmbedding(Embedding):
    """shared/embeddings.py:25-46: Conv1D(embedding_dim, 1) under scope ``_scope + 'Embedding/conv1d'``."""

    def __init__(self, embedding_dim, _scope='', store: Optional[VariableStore] = None):
        super(LinearEmbedding, self).__init__('linear', embedding_dim)
        self.store = store or default_store()
        self.scope = _scope + 'Embedding/conv1d'
        self._built = None

    def _build(self, d1):
        if self._built != d1:
            self.store.get(self.scope + '/kernel', [1, d1, self.embedding_dim], fans=(d1, self.embedding_dim))
            self.store.get(self.scope + '/bias', [self.embedding_dim], init="zeros")
            self._built = d1

    def __call__(self, input_pnt):
        import torch
        t0 = time.time()
        x = input_pnt if isinstance(input_pnt, torch.Tensor) else torch.as_tensor(np.asarray(input_pnt, np.float32))
        d1 = x.shape[-1]
        self._build(d1)
        task_name = "vrptw" if d1 == 4 else "vrp"
        args = dict(task_name=task_name, n_nodes=x.shape[-2], n_cust=x.shape[-2] - 1, input_dim=d1 + 1,
                    embedding_dim=self.embedding_dim, hidden_dim=128, decode_len=1, capacity=1.0)
        def params():
            p = _dummy_model_params(args)
            p['Actor/Embedding/conv1d/kernel'] = self.store.vars[self.scope + '/kernel']
            p['Actor/Embedding/conv1d/bias'] = self.store.vars[self.scope + '/bias']
            return p
        eng = _get_engine(args, params, ("emb", self.store.uid, self.store.version, self.scope))
        out = eng.embed(x)
        self.total_time += time.time() - t0
        return out


class _AttentionActorBase(object):
    """Attention*Actor of VRP/vrp_attention.py:3-76 and VRPTW/vrptw_attention.py:4-84 (and the UPS twins)."""
    TW = False

    def __init__(self, dim, use_tanh=False, C=10, _name='Attention', _scope='', store: Optional[VariableStore] = None):
        self.use_tanh = use_tanh
        self._scope = _scope
        self._name = _name
        self.dim = dim
        self.C = C
        self.store = store or default_store()
        self.prefix = _scope + _name
        st, p, H = self.store, self.prefix, dim
        st.get(p + '/v', [1, H], fans=(1, H))
        for blk in ['emb_d', 'emb_ld'] + (['emb_time'] if self.TW else []):
            st.get(f'{p}/{blk}/kernel', [1, 1, H], fans=(1, H))
            st.get(f'{p}/{blk}/bias', [H], init="zeros")
        for blk in ['proj_d', 'proj_ld'] + (['proj_t'] if self.TW else []):
            st.get(f'{p}/{blk}/kernel', [1, H, H], fans=(H, H))
            st.get(f'{p}/{blk}/bias', [H], init="zeros")
        st.get(p + '/proj_q/kernel', [H, H], fans=(H, H))
        st.get(p + '/proj_q/bias', [H], init="zeros")
        self._ref_built = None

    def _build_ref(self, E):
        if self._ref_built != E:
            self.store.get(self.prefix + '/proj_ref/kernel', [1, E, self.dim], fans=(E, self.dim))
            self.store.get(self.prefix + '/proj_ref/bias', [self.dim], init="zeros")
            self._ref_built = E

    def variable_names(self):
        return [k for k in self.store.vars if k.startswith(self.prefix + '/')]

    def __call__(self, query, ref, env):
        """query [R, dim], ref [R, N, E], env with .demand/.load(/.time) -> (e [R, N, dim], logits [R, N])."""
        E, N = ref.shape[-1], ref.shape[-2]
        self._build_ref(E)
        args = dict(task_name="vrptw" if self.TW else "vrp", n_nodes=N, n_cust=N - 1, input_dim=5 if self.TW else 3,
                    embedding_dim=E, hidden_dim=self.dim, decode_len=1, capacity=1.0, use_tanh=self.use_tanh,
                    tanh_exploration=float(self.C))
        def params():
            p = _dummy_model_params(args)
            for k in self.variable_names():
                p['Actor/Decoder/Attention' + k[len(self.prefix):]] = self.store.vars[k]
            return p
        eng = _get_engine(args, params, ("att", self.store.uid, self.store.version, self.prefix))
        return eng.attention(-1, query, ref, env.demand, env.load, env.time if self.TW else None)


class AttentionVRPActor(_AttentionActorBase):
    TW = False


class AttentionVRPTWActor(_AttentionActorBase):
    TW = True


class AttentionVRP_UPS_Actor(AttentionVRPActor):      # UPS/vrp_ups_attention.py: identical arithmetic
    pass


class AttentionVRPTW_UPS_Actor(AttentionVRPTWActor):  # UPS/vrptw_ups_attention.py: identical arithmetic
    pass


class _AttentionCriticBase(object):
    """Attention*Critic of VRP/vrp_attention.py:79-140 and VRPTW/vrptw_attention.py:87-169 (and the UPS twins): one process
    block of the critic.  It owns the variables the reference creates - v, emb_d, proj_d, proj_q, proj_e and, for VRPTW, the
    two time-window layers the reference actually CALLS (emb_begin_tw and proj_end_tw, each for both windows, :147-150).
    The blocks are evaluated together by vrpb200_critic (RLAgent.build_model('stochastic'))."""
    TW = False

    def __init__(self, dim, use_tanh=False, C=10, _name='Attention', _scope='', store: Optional[VariableStore] = None):
        self.use_tanh = use_tanh
        self.dim = dim
        self.C = C
        self.store = store or default_store()
        self.prefix = _scope + _name
        st, p, H = self.store, self.prefix, dim
        st.get(p + '/v', [1, H], fans=(1, H))
        for blk in ['emb_d'] + (['emb_begin_tw'] if self.TW else []):
            st.get(f'{p}/{blk}/kernel', [1, 1, H], fans=(1, H))
            st.get(f'{p}/{blk}/bias', [H], init="zeros")
        for blk in ['proj_d'] + (['proj_end_tw'] if self.TW else []):
            st.get(f'{p}/{blk}/kernel', [1, H, H], fans=(H, H))
            st.get(f'{p}/{blk}/bias', [H], init="zeros")
        st.get(p + '/proj_q/kernel', [H, H], fans=(H, H))
        st.get(p + '/proj_q/bias', [H], init="zeros")

    def _build_ref(self, E):
        self.store.get(self.prefix + '/proj_e/kernel', [1, E, self.dim], fans=(E, self.dim))
        self.store.get(self.prefix + '/proj_e/bias', [self.dim], init="zeros")

    def __call__(self, query, ref, env):
        raise NotImplementedError("a single critic block is not exposed; RLAgent.build_model('stochastic') evaluates the "
                                  "whole critic (vrpb200_critic)")


class AttentionVRPCritic(_AttentionCriticBase):
    TW = False


class AttentionVRPTWCritic(_AttentionCriticBase):
    TW = True


class AttentionVRP_UPS_Critic(AttentionVRPCritic):      # UPS/vrp_ups_attention.py: identical arithmetic
    pass


class AttentionVRPTW_UPS_Critic(AttentionVRPTWCritic):  # UPS/vrptw_ups_attention.py: identical arithmetic
    pass


class DecodeStep(object):
    """shared/decode_step.py:3-128 (base class: owns the glimpse and pointer attention modules)."""

    def __init__(self, ClAttention, hidden_dim, use_tanh=False, tanh_exploration=10., n_glimpses=0,
                 mask_glimpses=True, mask_pointer=True, _scope='', store=None):
        self.hidden_dim = hidden_dim
        self.use_tanh = use_tanh
        self.tanh_exploration = tanh_exploration
        self.n_glimpses = n_glimpses
        self.mask_glimpses = mask_glimpses
        self.mask_pointer = mask_pointer
        self._scope = _scope
        self.BIGNUMBER = 100000.
        self.store = store or default_store()
        self.glimpses = [None for _ in range(self.n_glimpses)]
        for i in range(self.n_glimpses):
            self.glimpses[i] = ClAttention(hidden_dim, use_tanh=False, _scope=self._scope, _name="Glimpse" + str(i), store=self.store)
        self.pointer = ClAttention(hidden_dim, use_tanh=use_tanh, C=tanh_exploration, _scope=self._scope,
                                   _name="Decoder/Attention", store=self.store)


class RNNDecodeStep(DecodeStep):
    """shared/decode_step.py:131-234: one BasicLSTMCell layer + glimpses + pointer."""
    LSTM = 'Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell'

    def __init__(self, ClAttention, hidden_dim, use_tanh=False, tanh_exploration=10., n_glimpses=0, mask_glimpses=True,
                 mask_pointer=True, forget_bias=1.0, rnn_layers=1, _scope='', store=None):
        super(RNNDecodeStep, self).__init__(ClAttention, hidden_dim, use_tanh=use_tanh, tanh_exploration=tanh_exploration,
                                            n_glimpses=n_glimpses, mask_glimpses=mask_glimpses, mask_pointer=mask_pointer,
                                            _scope=_scope, store=store)
        if rnn_layers != 1:
            raise ValueError("rnn_layers must be 1 (the reference shares one cell across layers, decode_step.py:180)")
        self.forget_bias = forget_bias
        self.rnn_layers = rnn_layers
        self.dropout = 0.0     # the reference's placeholder; the rollout path always feeds 0.0 (attention_agent.py:495)
        self._lstm_built = None

    def _build_lstm(self, E):
        if self._lstm_built != E:
            H = self.hidden_dim
            self.store.get(self._scope + self.LSTM + '/kernel', [E + H, 4 * H], fans=(E + H, 4 * H))
            self.store.get(self._scope + self.LSTM + '/bias', [4 * H], init="zeros")
            self.pointer._build_ref(E)
            for g in self.glimpses:
                g._build_ref(E)
            self._lstm_built = E

    def model_params(self, embedder: Optional[LinearEmbedding] = None, args=None) -> Dict[str, np.ndarray]:
        """The store entries under the C ABI's canonical names (scope 'Actor/')."""
        out = {}
        sc = self._scope
        for k, v in self.store.vars.items():
            if k.startswith(sc) and (k.startswith(sc + 'Decoder/') or k.startswith(sc + 'Glimpse')):
                out['Actor/' + k[len(sc):]] = v
        if embedder is not None:
            out['Actor/Embedding/conv1d/kernel'] = self.store.vars[embedder.scope + '/kernel']
            out['Actor/Embedding/conv1d/bias'] = self.store.vars[embedder.scope + '/bias']
        elif args is not None:
            d1 = args['input_dim'] - 1
            out['Actor/Embedding/conv1d/kernel'] = np.zeros([1, d1, args['embedding_dim']], np.float32)
            out['Actor/Embedding/conv1d/bias'] = np.zeros([args['embedding_dim']], np.float32)
        return out

    def _engine(self, Env, E, N):
        self._build_lstm(E)
        tw = self.pointer.TW
        args = dict(task_name="vrptw" if tw else "vrp", n_nodes=N, n_cust=N - 1, input_dim=5 if tw else 3, embedding_dim=E,
                    hidden_dim=self.hidden_dim, decode_len=1, capacity=float(getattr(Env, "capacity", 1.0)),
                    n_glimpses=self.n_glimpses, use_tanh=self.use_tanh, tanh_exploration=float(self.tanh_exploration),
                    mask_glimpses=self.mask_glimpses, mask_pointer=self.mask_pointer, forget_bias=float(self.forget_bias))
        return _get_engine(args, lambda: self.model_params(args=args), ("dec", self.store.uid, self.store.version, self._scope))

    def step(self, decoder_inp, context, Env, decoder_state=None, *args, **kwargs):
        """decoder_inp [R,1,E], context [R,N,E], decoder_state = (c, h) each [R,H] (LSTMStateTuple order)
        -> logit, prob, logprob [R,N], decoder_state."""
        if self.dropout not in (0, 0.0):
            raise NotImplementedError("dropout > 0 is the training step's setting; the rollout path uses 0.0")
        import torch
        E, N = context.shape[-1], context.shape[-2]
        eng = self._engine(Env, E, N)
        c, h = decoder_state
        c, h = c.clone().contiguous(), h.clone().contiguous()
        logit, prob, logprob = eng.decode_step(decoder_inp, context, Env.demand, Env.load,
                                               Env.time if self.pointer.TW else None, Env.mask, c, h)
        return logit, prob, logprob, (c, h)


# --------------------------------------------------------------------------------------------------
# environment
# --------------------------------------------------------------------------------------------------
StateVRP = collections.namedtuple("State", ("load", "demand", "d_sat", "mask"))            # VRP/vrp_utils.py:126-131
StateVRPTW = collections.namedtuple("State", ("load", "time", "demand", "d_sat", "mask"))  # VRPTW/vrptw_utils.py:153-159


class _EnvBase(object):
    TW = False
    UPS_OLD = False

    def __init__(self, args):
        self.args = args
        self.capacity = args['capacity']
        self.n_nodes = args['n_nodes']
        self.n_cust = args['n_cust']
        self.input_dim = args['input_dim']
        self._input = None
        self.time = self.load = self.mask = self.demand = None
        self.beam_width = 1
        eargs = dict(task_name="vrptw" if self.TW else "vrp", n_nodes=self.n_nodes, n_cust=self.n_cust,
                     input_dim=self.input_dim, embedding_dim=args.get('embedding_dim', 30), hidden_dim=128, decode_len=1,
                     capacity=float(self.capacity), ups_old=bool(self.UPS_OLD))
        self._eargs = eargs

    # placeholder semantics: assign the batch, everything derived follows
    @property
    def input_data(self):
        return self._input

    @input_data.setter
    def input_data(self, value):
        import torch
        x = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value, np.float32))
        self._input = x.to(device="cuda", dtype=torch.float32).contiguous()
        x = self._input
        self.input_pnt = x[:, :, :self.input_dim - 1]
        self.demand = x[:, :, -1]
        self.batch_size = x.shape[0]
        if self.TW:
            self.all_x, self.all_y, self.all_b_tw, self.all_e_tw = x[:, :, 0], x[:, :, 1], x[:, :, 2], x[:, :, 3]
            self.previous_x, self.previous_y = x[:, self.n_nodes - 1, 0:1], x[:, self.n_nodes - 1, 1:2]

    def _eng(self):
        return _get_engine(self._eargs, lambda: _dummy_model_params(self._eargs), ("env",))

    def _publish(self, st, d_sat):
        import torch
        self._state = st
        self.demand, self.load, self.mask = st["demand"], st["load"], st["mask"]
        W = self.beam_width
        if self.TW:
            self.time = st["time"]
            self.previous_x, self.previous_y = st["prev_xy"][:, 0:1], st["prev_xy"][:, 1:2]
            x = self._input
            self.all_x, self.all_y = x[:, :, 0].repeat(W, 1), x[:, :, 1].repeat(W, 1)
            self.all_b_tw, self.all_e_tw = x[:, :, 2].repeat(W, 1), x[:, :, 3].repeat(W, 1)
            return StateVRPTW(load=self.load, time=self.time, demand=self.demand, d_sat=d_sat, mask=self.mask)
        return StateVRP(load=self.load, demand=self.demand, d_sat=d_sat, mask=self.mask)

    def reset(self, beam_width=1):
        """Env.reset (VRPTW/vrptw_utils.py:199-252, VRP/vrp_utils.py:159-196)."""
        import torch
        if self._input is None:
            raise RuntimeError("assign env.input_data before reset()")
        self.beam_width = beam_width
        self.batch_beam = self.batch_size * beam_width
        st = self._eng().env_reset(self._input, beam_width)
        return self._publish(st, torch.zeros((self.batch_beam, self.n_nodes), device=self._input.device))

    def step(self, idx, beam_parent=None):
        """Env.step (VRPTW/vrptw_utils.py:254-358, VRP/vrp_utils.py:198-255): idx, beam_parent [R,1] int64."""
        st, d_sat = self._eng().env_step(self._input, self.beam_width, self._state, idx, beam_parent)
        return self._publish(st, d_sat)


class VRPEnv(_EnvBase):
    TW = False


class VRPTWEnv(_EnvBase):
    TW = True


class VRPTWEnvUPSOld(VRPTWEnv):
    """Env of UPS_old/vrptw_ups_utils.py:170-370: the VRPTW Env with x, y = latitude, longitude in radians, travel time
    (100/13 * 1.2) x equirectangular km (:321-323, :359-361) and 1.5 clicks of service per package (:332-333)."""
    UPS_OLD = True


def reward_func(sample_solution, decode_len=0.0, n_nodes=0.0, depot=None, _ups=False):
    """reward_func of VRPTW/vrptw_utils.py:361-414 and VRP/vrp_utils.py:257-307.
    sample_solution: list of T tensors [R, A] or a tensor [T, R, A].
    depot=None: closed tour length over x, y (:397-414).  depot [R, A]: the min_trucks objective (:384-395, 410-412),
    70/n_nodes x returns to the depot + 30 x route length / sum of the distances to the depot."""
    import torch
    sol = sample_solution if isinstance(sample_solution, torch.Tensor) else torch.stack(list(sample_solution), 0)
    sol = sol.to(device="cuda", dtype=torch.float32).contiguous()
    args = dict(task_name="vrp", n_nodes=2, n_cust=1, input_dim=3, embedding_dim=1, hidden_dim=128, decode_len=1, capacity=1.0)
    eng = _get_engine(args, lambda: _dummy_model_params(args), ("reward",))
    if depot is None:
        return eng.reward(sol)
    dep = depot.to(device="cuda", dtype=torch.float32).contiguous()
    return eng.reward_min_trucks(sol, dep, float(n_nodes), ups=_ups)


def reward_func_vrptw_ups_old(sample_solution, decode_len=0.0, n_nodes=0.0, depot=None):
    """reward_func of UPS_old/vrptw_ups_utils.py:380-430.  depot=None: the closed tour length in km over the
    equirectangular distance (:416-424).  depot [R, A]: 70/n_nodes x steps at the depot (every one counts, :401-405) + 30 x
    that km length / sum of the Euclidean distances to the depot (:406-408, 426-427)."""
    import torch
    sol = sample_solution if isinstance(sample_solution, torch.Tensor) else torch.stack(list(sample_solution), 0)
    sol = sol.to(device="cuda", dtype=torch.float32).contiguous()
    args = dict(task_name="vrptw", n_nodes=2, n_cust=1, input_dim=5, embedding_dim=1, hidden_dim=128, decode_len=1, capacity=1.0,
                ups_old=True)
    eng = _get_engine(args, lambda: _dummy_model_params(args), ("reward",))
    if depot is None:
        return eng.reward(sol)
    dep = depot.to(device="cuda", dtype=torch.float32).contiguous()
    return eng.reward_min_trucks(sol, dep, float(n_nodes), ups_old=True)


def reward_func_vrptw_ups(sample_solution, decode_len=0.0, n_nodes=0.0, depot=None):
    """reward_func of UPS/vrptw_ups_utils.py:362-422: the tour length, or with a depot 100 x returns to the depot + the
    tour length (the great-circle normaliser of :399-402 is computed by the reference but does not enter the reward)."""
    return reward_func(sample_solution, decode_len, n_nodes, depot, _ups=True)


class DataGenerator(object):
    """DataGenerator of VRP/vrp_utils.py:55-123, VRPTW/vrptw_utils.py:70-150 and the UPS twins: host-side numpy."""

    def __init__(self, args):
        self.args = args
        self.rnd = np.random.RandomState(seed=args['random_seed'])
        self.n_problems = args['test_size']
        if args.get('data_dir'):   # the reference's on-disk test set: loaded if the file exists, else created and saved
            self.test_data = _data.load_or_create_dataset(args['task_name'], self.n_problems, args['n_cust'], args['data_dir'],
                                                          args['random_seed'] + 1, 'test', ups=bool(args.get('ups', False)))
        else:
            self.test_data = _data.create_dataset(args['task_name'], self.n_problems, args['n_cust'], args['random_seed'] + 1,
                                                  ups=bool(args.get('ups', False)))
        self.reset()

    def reset(self):
        self.count = 0

    def get_train_next(self):
        a = self.args
        seed = int(self.rnd.randint(0, 2**31 - 1))
        return _data.create_dataset(a['task_name'], a['batch_size'], a['n_cust'], seed, ups=bool(a.get('ups', False)))

    def get_test_next(self):
        if self.count >= self.args['test_size']:
            warnings.warn("The test iterator reset.")
            self.count = 0
        out = self.test_data[self.count:self.count + 1]
        self.count += 1
        return out

    def get_test_all(self):
        return self.test_data


def load_ups_old_components():
    """The UPS_old package (UPS_old/vrptw_ups_utils.py, UPS_old/vrptw_ups_attention.py): nothing in main.py selects it; its
    Env physics and reward with the VRPTW actor (the actor file is identical to VRPTW/vrptw_attention.py).  Run the agent
    with args['ups_old'] = True so that the fused rollout uses the same physics."""
    return (DataGenerator, VRPTWEnvUPSOld, reward_func_vrptw_ups_old, AttentionVRPTWActor, AttentionVRPTWCritic)


def load_task_specific_components(task, ups=False):
    """main.py:17-53 -> (DataGenerator, Env, reward_func, AttentionActor, AttentionCritic)."""
    if task == 'vrp':
        return (DataGenerator, VRPEnv, reward_func, AttentionVRP_UPS_Actor if ups else AttentionVRPActor, AttentionVRPCritic)
    if task == 'vrptw':
        return (DataGenerator, VRPTWEnv, reward_func_vrptw_ups if ups else reward_func,
                AttentionVRPTW_UPS_Actor if ups else AttentionVRPTWActor, AttentionVRPTWCritic)
    raise Exception('Task is not implemented')


# --------------------------------------------------------------------------------------------------
# the agent
# --------------------------------------------------------------------------------------------------
class _Printer:
    def print_out(self, s, *a, **k):
        print(s)


class RLAgent(object):
    """model/attention_agent.py:11-556, inference side.  ``build_model(decode_type)`` EXECUTES the rollout on
    ``env.input_data`` (the reference builds a graph that sess.run executes later) and returns the same tuple
    ``(R, v, logprobs, actions, idxs, input_pnt, probs)``."""

    def __init__(self, args, prt, env, dataGen, reward_func, clAttentionActor, clAttentionCritic, is_train=False, _scope='',
                 store: Optional[VariableStore] = None, device: int = 0):
        self.is_train = is_train
        if args.get('embedding_graph', 0) != 0:
            raise NotImplementedError("only embedding_graph=0 (LinearEmbedding) is on the rollout path")
        self.args = args
        self.prt = prt or _Printer()
        self.env = env
        self.dataGen = dataGen
        self.reward_func = reward_func
        self.clAttentionCritic = clAttentionCritic
        self.store = store or default_store()
        self.device = device
        self.embedder_model = LinearEmbedding(args['embedding_dim'], _scope=_scope + 'Actor/', store=self.store)
        self.decodeStep = RNNDecodeStep(clAttentionActor, args['hidden_dim'], use_tanh=args['use_tanh'],
                                        tanh_exploration=args['tanh_exploration'], n_glimpses=args['n_glimpses'],
                                        mask_glimpses=args['mask_glimpses'], mask_pointer=args['mask_pointer'],
                                        forget_bias=args['forget_bias'], rnn_layers=args['rnn_layers'], _scope=_scope + 'Actor/',
                                        store=self.store)
        # created, saved, never used by the reference (attention_agent.py:65-66, shadowed at :133)
        self.store.get(_scope + 'decoder_input', [1, 1, args['embedding_dim']], fans=(args['embedding_dim'], args['embedding_dim']))
        self.embedder_model._build(args['input_dim'] - 1)
        self.decodeStep._build_lstm(args['embedding_dim'])
        self.sess = None
        self.last_time = None
        self._scope = _scope
        self._critic_built = False
        self._train_calls = 0
        self._adam = {}                  # 'actor' / 'critic' -> device copies of the variables + Adam slots (build_train_step)

    # -- checkpoint ------------------------------------------------------------------------------
    def Initialize(self, sess=None):
        self.sess = sess
        self.load_model()

    def load_model(self):
        """args['load_path']: directory holding model.npz, or an .npz file, keyed by TF variable name."""
        path = self.args.get('load_path') or ''
        if os.path.isdir(path):
            path = os.path.join(path, 'model.npz')
        if path and os.path.exists(path):
            self.store.load(path, strict=False)
            print("have load model")

    def model_params(self):
        p = self.decodeStep.model_params(self.embedder_model)
        sc = self._scope + 'Critic/'
        for k, v in self.store.vars.items():
            if k.startswith(sc):
                p['Critic/' + k[len(sc):]] = v
        return p

    def _build_critic(self):
        """Variables of the critic, created where the reference creates them: in build_model('stochastic')
        (attention_agent.py:246-268) - n_process_blocks Attention*Critic blocks under Critic/Process, Critic/Linear/L1, L2."""
        if self._critic_built or self.clAttentionCritic is None:
            return
        H, st, sc = self.args['hidden_dim'], self.store, self._scope + 'Critic/'
        for i in range(self.args.get('n_process_blocks', 3)):
            blk = self.clAttentionCritic(H, _name="P" + str(i), _scope=sc + 'Process/', store=st)
            blk._build_ref(self.args['embedding_dim'])
        st.get(sc + 'Linear/L1/kernel', [H, H], fans=(H, H))
        st.get(sc + 'Linear/L1/bias', [H], init="zeros")
        st.get(sc + 'Linear/L2/kernel', [H, 1], fans=(H, 1))
        st.get(sc + 'Linear/L2/bias', [1], init="zeros")
        self._critic_built = True

    def build_train_step(self, uniforms=None, uniforms_retry=None, seed=0, apply=True, input_scale=None):
        """attention_agent.py:272-322: ONE training step - the stochastic rollout, the critic, the two losses,
        compute_gradients(actor_loss) over the Actor variables (vrpb200_actor_grad: back-propagation through the teacher-forced
        rollout) and compute_gradients(critic_loss) over the Critic variables (vrpb200_critic_grad), per-variable
        clip_by_norm(max_grad_norm) and the two Adam updates with actor_net_lr / critic_net_lr (vrpb200_clip_adam; apply=False
        only computes).  Returns the reference's list [actor_train_step, critic_train_step, actor_loss, critic_loss,
        actor_gra_and_var, critic_gra_and_var, R, v, logprobs, probs, actions, idxs]: the *_train_step slots hold the number of
        Adam steps taken so far, the *_gra_and_var slots lists of (gradient tensor, variable name).  input_scale [T, B, E]: the
        LSTM-input dropout of the reference's decodeStep.dropout feed as keep-mask / keep-probability (None: keep-probability 1);
        the tour is then decoded step by step with the scaled inputs and the gradient is taken through the same scale.
        Glimpses are not differentiated: with n_glimpses > 0 the actor slots are None."""
        import torch
        if input_scale is not None:
            input_scale = torch.as_tensor(input_scale, dtype=torch.float32).to("cuda").contiguous()
            if uniforms is None:                    # the step-by-step route draws on the host side of the ABI
                gen = torch.Generator(device="cuda").manual_seed(int(seed))
                shape = (self.args['decode_len'], self.env.input_data.shape[0])
                uniforms, uniforms_retry = (torch.rand(shape, device="cuda", generator=gen) for _ in range(2))
        R, v, logprobs, actions, idxs, batch, probs = self.build_model('stochastic', uniforms=uniforms, uniforms_retry=uniforms_retry,
                                                                       seed=seed, input_scale=input_scale)
        eng = self.engine()
        actor_loss, critic_loss = eng.reinforce_loss(R, v, torch.stack(logprobs))
        rows = R.shape[0]
        out = [None, None, actor_loss, critic_loss, None, None, R, v, logprobs, probs, actions, idxs]
        todo = []
        if self.args['n_glimpses'] == 0:
            sc = self._scope + 'Actor/'
            names = [k for k in self.store.vars if k.startswith(sc)]
            weight = (R - v) / float(rows)             # d actor_loss / d sum_t logprob[t, r]; v enters as a constant (:286)
            grads = eng.actor_grad(self.env.input_data, torch.stack(idxs)[:, :, 0], weight,
                                   {'Actor/' + k[len(sc):]: self.store.vars[k].shape for k in names}, input_scale=input_scale)
            out[4] = [(grads['Actor/' + k[len(sc):]], k) for k in names]
            todo.append((0, 'actor', names, out[4], self.args.get('actor_net_lr', 1e-4)))
        if self.clAttentionCritic is not None:
            sc = self._scope + 'Critic/'
            names = [k for k in self.store.vars if k.startswith(sc)]
            grads, _ = eng.critic_grad(self.env.input_data, R, {'Critic/' + k[len(sc):]: self.store.vars[k].shape for k in names})
            out[5] = [(grads['Critic/' + k[len(sc):]], k) for k in names]
            todo.append((1, 'critic', names, out[5], self.args.get('critic_net_lr', 1e-4)))
        for slot, kind, names, gv, lr in todo:
            out[slot] = self._adam[kind]['step'] if kind in self._adam else 0
            if apply:
                out[slot] = self._apply_adam(eng, kind, names, [g for g, _ in gv], lr)
        return out

    def run_train_step(self, uniforms=None, uniforms_retry=None):
        """attention_agent.py:530-556: draw the next training batch from the data generator, feed it to the Env and run the train
        step on it with decodeStep.dropout = args['dropout'] (:543): a fresh keep-mask of the LSTM inputs per step, row and
        embedding component, drawn on the device from a generator seeded by (random_seed, step count).  (The feature-normalised
        copy of the batch the reference also feeds is read by the graph embeddings only, which are not on this path.)  With
        dropout 0 the rollout is the fused kernel's, its sampling stream seeded the same way."""
        import torch
        data = self.dataGen.get_train_next()
        self.env.input_data = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
        self._train_calls += 1
        seed = int(self.args.get('random_seed', 0)) * 1000003 + self._train_calls
        scale, rate = None, float(self.args.get('dropout', 0.0) or 0.0)
        if rate > 0.0 and self.args['n_glimpses'] == 0:
            gen = torch.Generator(device="cuda").manual_seed(seed ^ 0x5DEECE66D)
            shape = (self.args['decode_len'], data.shape[0], self.args['embedding_dim'])
            scale = (torch.rand(shape, device="cuda", generator=gen) < (1.0 - rate)).float() / (1.0 - rate)     # tf.nn.dropout
        return self.build_train_step(uniforms=uniforms, uniforms_retry=uniforms_retry, seed=seed, input_scale=scale)

    def save_model(self, path):
        """saver.save (main.py:131): every variable of the collection under its TF name, as load_model reads it back."""
        self.store.save(path)

    def _apply_adam(self, eng, kind, names, grads, lr):
        """clip_by_norm + Adam on device copies of the variables (attention_agent.py:303-311); the variable collection stays the
        single source of truth and is re-read whenever something else assigned to it."""
        import torch
        st = self._adam.get(kind)
        if st is None or st['names'] != names or any(self.store.vars[k] is not a for k, a in zip(names, st['host'])):
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
            keep = st is not None and st['names'] == names           # same variables, new values (a restore): the slots stay
            st = self._adam[kind] = dict(names=names, var=[dev(self.store.vars[k]) for k in names],
                                         m=st['m'] if keep else [torch.zeros_like(g) for g in grads],
                                         v=st['v'] if keep else [torch.zeros_like(g) for g in grads], step=st['step'] if keep else 0)
        st['step'] += 1
        eng.clip_adam(st['var'], grads, st['m'], st['v'], max_norm=self.args.get('max_grad_norm', 2.0), lr=lr, step=st['step'])
        st['host'] = []
        for k, t in zip(names, st['var']):
            self.store.vars[k] = t.cpu().numpy().reshape(self.store.vars[k].shape)
            st['host'].append(self.store.vars[k])
        self.store.version += 1
        return st['step']

    def engine(self):
        eargs = dict(self.args)
        eargs.setdefault('task_name', 'vrptw' if self.env.TW else 'vrp')
        return _get_engine(eargs, self.model_params, ("agent", self.store.uid, self.store.version), self.device)

    # -- the rollout -----------------------------------------------------------------------------
    def build_model(self, decode_type="greedy", uniforms=None, uniforms_retry=None, seed=0, want_probs=False,
                    backend="auto", want_logprobs=True, input_scale=None):
        """attention_agent.py:84-272.  backend 'fused' = the persistent rollout kernel, 'stepwise' = the reference's
        loop over RNNDecodeStep.step / Env.step on device state (per-step probabilities, and beam search combined with
        glimpses or UPS_old physics, take this route), 'auto' picks."""
        import torch
        special = self.args['n_glimpses'] != 0 or bool(self.args.get('ups_old'))   # fused in the FP32-pipe kernel, no beams
        if input_scale is not None:                 # [T, rows, E] LSTM-input dropout scale (training): the step-by-step route applies it
            input_scale = torch.as_tensor(input_scale, dtype=torch.float32).to("cuda")
        fused_ok = decode_type in ("greedy", "stochastic", "beam_search") and not want_probs and \
            not (special and decode_type == "beam_search") and input_scale is None
        if backend == "auto":
            backend = "fused" if fused_ok else "stepwise"
        if backend == "fused":
            if not fused_ok:
                raise NotImplementedError("fused kernel: greedy / stochastic (beam search without glimpses), no per-step probabilities")
            W = self.args['beam_width'] if decode_type == 'beam_search' else 1
            self.beam_width = W
            # the per-step log-probabilities are only computed on request: without them plain greedy decoding runs the
            # kernel's fast per-row tail (evaluate_batch / evaluate_single never read them)
            # args['whole_batch_retry']: the reference's sampling retry (one redraw of EVERY row when any row fails) instead
            # of the per-row retry; single-GPU batches only
            self.engine().set_sampling_retry(bool(self.args.get('whole_batch_retry', False)))
            out = self.engine().rollout(self.env.input_data, decode_type, beam_width=W, uniforms=uniforms,
                                        uniforms_retry=uniforms_retry, seed=seed, want_logprobs=want_logprobs)
            idx = out.idxs.long()                                      # [T, R]
            T, R = idx.shape
            pnt = self.env.input_pnt                                    # [B, N, D-1]
            B = pnt.shape[0]
            inst = torch.arange(R, device=idx.device) % B               # instance of beam-major row w*B + b
            actions = pnt[inst[None, :].expand(T, R), idx]              # [T, R, D-1]
            if decode_type == 'beam_search':
                par = out.beam_parents.long()
                self.last_beam_path = [p[:, None] for p in par]
                actions = torch.stack(self._backtrack(list(actions), list(par), B))
            R_out = self._reward(list(actions), pnt, W) if self.args.get('min_trucks') else out.R
            v = self._critic_value(decode_type)
            return (R_out, v, list(out.logprobs) if want_logprobs else None, list(actions), [i[:, None] for i in idx], pnt, None)
        return self._build_model_stepwise(decode_type, uniforms, uniforms_retry, want_probs, input_scale)

    def _critic_value(self, decode_type):
        """attention_agent.py:243-270: v = 0 unless decoding stochastically; then the critic's value of every instance."""
        import torch
        if decode_type != 'stochastic' or self.clAttentionCritic is None:
            return torch.zeros((), device="cuda")
        self._build_critic()
        return self.engine().critic(self.env.input_data)

    def _reward(self, actions, input_pnt, beam_width):
        """attention_agent.py:236-240: the tour length, or with args['min_trucks'] the depot form of reward_func on the
        (back-tracked) actions with the depot coordinates tiled over the beams."""
        if self.args.get('min_trucks'):
            depot = input_pnt[:, self.env.n_nodes - 1, :].repeat(beam_width, 1)
            return self.reward_func(actions, self.args['decode_len'], self.args['n_nodes'] - 1, depot)
        return self.reward_func(actions)

    # -- the step-by-step route: one decode_step / env_step launch per step ---------------------------
    # Serves what the persistent kernels do not return or combine: per-step probability tables (want_probs) and beam search
    # together with glimpses or UPS_old physics.  Semantics of attention_agent.py:89-240; rows are beam-major (row = w*B + b).
    @staticmethod
    def _pick_greedy(prob):
        """argmax over the probabilities, first maximal index (:146-147)."""
        import torch
        return torch.argmax(prob, 1, keepdim=True), None

    def _pick_sampled(self, prob, step, uniforms, uniforms_retry):
        """Inverse-CDF draw (:148-167): the first node whose running sum exceeds the row's uniform; if ANY row finds none,
        every row redraws once with the second set of uniforms; a row failing again takes node 0."""
        import torch
        rows, n_nodes = prob.shape
        cdf = torch.cumsum(prob, 1)

        def draw(table):
            u = (torch.as_tensor(table[step]) if table is not None else torch.rand(rows)).to(prob.device)[:, None]
            beyond = ~(cdf > u)                                   # nodes whose running sum has not passed u yet
            return beyond.all(1), beyond.long().sum(1, keepdim=True).clamp_(max=n_nodes - 1) * (~beyond.all(1, keepdim=True)).long()

        failed, idx = draw(uniforms)
        if bool(failed.any()):
            failed, idx = draw(uniforms_retry)
        return idx, None

    def _pick_beam(self, prob, step, batch, width, carried):
        """Per instance the `width` best of (log-probability so far + log prob) over width x n_nodes candidates (:169-205);
        ties go to the lower flat index (tf.nn.top_k), step 0 looks at beam 0 only; returns node, parent beam, the parents'
        probability rows and the new running log-probabilities."""
        import torch
        n_nodes = prob.shape[1]
        if step == 0:
            cand = torch.log(prob[:batch])
        else:
            cand = (torch.log(prob) + carried).view(width, batch, n_nodes).permute(1, 0, 2).reshape(batch, width * n_nodes)
        order = torch.sort(cand, dim=1, descending=True, stable=True)[1][:, :width]         # [B, W], stable = lower index first
        flat = order.t().reshape(-1, 1)                                                     # beam-major rows
        score = torch.gather(cand, 1, order).t().reshape(-1, 1)
        node, parent = flat % n_nodes, flat // n_nodes
        src = (torch.arange(batch, device=prob.device).repeat(width)[:, None] + batch * parent)[:, 0]
        return node, parent, prob[src], score

    @staticmethod
    def _backtrack(per_step, parents, batch):
        """attention_agent.py:222-231: follow the parent pointers from the last step to the first, so that row r of every
        step belongs to the beam that ENDS in row r.  per_step: list of [R, ...] tensors, parents: list of [R, 1] or [R]."""
        import torch
        rows = per_step[0].shape[0]
        inst = torch.arange(rows, device=per_step[0].device) % batch
        at = torch.arange(rows, device=per_step[0].device)
        out = [None] * len(per_step)
        for k in reversed(range(len(per_step))):
            out[k] = per_step[k][at]
            at = (inst + batch * parents[k].reshape(-1))[at]
        return out

    def _build_model_stepwise(self, decode_type, uniforms, uniforms_retry, want_probs, input_scale=None):
        import torch
        if decode_type not in ('greedy', 'stochastic', 'beam_search'):
            raise ValueError(decode_type)
        args, env = self.args, self.env
        points = env.input_pnt
        batch, dev = points.shape[0], points.device
        width = args['beam_width'] if decode_type == 'beam_search' else 1
        self.beam_width = width
        rows = batch * width
        emb = self.embedder_model(points.contiguous())
        env.reset(width)
        context, tiled_points = emb.repeat(width, 1, 1), points.repeat(width, 1, 1)
        state = (torch.zeros(rows, args['hidden_dim'], device=dev), torch.zeros(rows, args['hidden_dim'], device=dev))
        dec_in = context[:, env.n_nodes - 1:env.n_nodes]                  # the depot's embedding is the first decoder input
        every = torch.arange(rows, device=dev)
        visited, logprobs, probs, idxs, parents, carried = [], [], [], [], [], None
        for step in range(args['decode_len']):
            if input_scale is not None:               # DropoutWrapper(input_keep_prob): the cell sees the scaled input (decode_step.py:177-179)
                dec_in = dec_in * input_scale[step][:, None, :]
            _, prob, _, state = self.decodeStep.step(dec_in, context, env, state)
            if decode_type == 'greedy':
                idx, parent = self._pick_greedy(prob)
            elif decode_type == 'stochastic':
                idx, parent = self._pick_sampled(prob, step, uniforms, uniforms_retry)
            else:
                idx, parent, prob, carried = self._pick_beam(prob, step, batch, width, carried)
                parents.append(parent)
            env.step(idx, parent)
            chosen = idx[:, 0]
            dec_in = context[every, chosen][:, None, :]
            logprobs.append(torch.log(prob[every, chosen]))
            probs.append(prob)
            idxs.append(idx)
            visited.append(tiled_points[every, chosen])
        actions = self._backtrack(visited, parents, batch) if decode_type == 'beam_search' else visited
        self.last_beam_path = parents
        R = self._reward(actions, points, width)
        v = self._critic_value(decode_type)
        return (R, v, logprobs, actions, idxs, points, probs if want_probs else None)

    # -- evaluation ------------------------------------------------------------------------------
    def evaluate_batch(self, eval_type='greedy', backend="auto"):
        """attention_agent.py:475-518: decode dataGen.get_test_all(); min over beams; print mean/std/time."""
        import torch
        beam_width = self.args['beam_width'] if eval_type == 'beam_search' else 1
        data = self.dataGen.get_test_all()
        self.env.input_data = torch.as_tensor(np.asarray(data, np.float32))
        torch.cuda.synchronize()
        start_time = time.time()
        R, v, logprobs, actions, idxs, batch, _ = self.build_model(eval_type, backend=backend, want_logprobs=False)
        R = R.cpu().numpy()
        R = np.concatenate(np.split(np.expand_dims(R, 1), beam_width, axis=0), 1)
        R = np.amin(R, 1, keepdims=False)
        end_time = time.time() - start_time
        self.last_time = end_time
        self.prt.print_out('Average of {} in batch-mode: {} -- std {} -- time {} s'.format(
            eval_type, np.mean(R), np.sqrt(np.var(R)), end_time))
        return R

    def evaluate_single(self, eval_type='greedy'):
        """attention_agent.py:336-413 (B=1 latency mode): one problem at a time through the same path; the decoded routes
        are written in the reference's result-file format when args['log_dir'] is set (_output_results, :416-471)."""
        import torch
        beam_width = self.args['beam_width'] if eval_type == 'beam_search' else 1
        self.dataGen.reset()
        rewards, all_output = [], []
        start_time = time.time()
        for _ in range(self.dataGen.n_problems):
            batch = np.asarray(self.dataGen.get_test_next(), np.float32)
            self.env.input_data = torch.as_tensor(batch)
            out = self.build_model(eval_type, want_logprobs=False)
            R = out[0].cpu().numpy().reshape(beam_width, -1)
            best = int(np.argmin(R[:, 0]))
            rewards.append(float(R[best, 0]))
            actions = out[3]                                              # T x [R, D-1]
            route = [list(batch[0, self.env.n_nodes - 1, :])]             # "we begin by the depot"
            for a_t in actions:
                route.append([float(v) for v in a_t[best * batch.shape[0]].cpu().numpy()])
            all_output.append(route)
        end_time = time.time() - start_time
        self.prt.print_out('Average of {} in single-mode: {} -- std {} -- time {} s'.format(
            eval_type, np.mean(rewards), np.sqrt(np.var(rewards)), end_time))
        if self.args.get('log_dir'):
            self._output_results(all_output, eval_type)
        return np.asarray(rewards)

    def _output_results(self, all_output, eval_type):
        """attention_agent.py:416-471: '<log_dir>/results/<task>[-ups]-size-B-len-N-results-<eval_type>.txt' (+ a copy of
        the test set file next to it when args['data_dir'] holds one)."""
        import shutil
        a = self.args
        dir_name = os.path.join(a['log_dir'], 'results')
        os.makedirs(dir_name, exist_ok=True)
        ups = bool(a.get('ups', False))
        stem = '{}{}-size-{}-len-{}'.format(a['task_name'], '-ups' if ups else '', a['test_size'], self.env.n_nodes)
        _data.write_results(os.path.join(dir_name, stem + '-results-{}.txt'.format(eval_type)), all_output, a['task_name'],
                            self.env.n_nodes)
        if a.get('data_dir'):
            src = os.path.join(a['data_dir'], stem + '-test.txt')
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(dir_name, stem + '-test.txt'))

    def inference(self, infer_type='batch'):
        if infer_type == 'batch':
            self.evaluate_batch('greedy')
            self.evaluate_batch('beam_search')
        elif infer_type == 'single':
            self.evaluate_single('greedy')
            self.evaluate_single('beam_search')
        self.prt.print_out("##################################################################")
