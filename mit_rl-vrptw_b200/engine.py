"""Thin host wrapper over the C ABI: one ``Engine`` = one ``vrpb200_handle`` on one device.

torch is used for device memory and streams only (plumbing); every computation is a call into
libvrpb200.so.  There is no fallback: constructing an Engine without the library or without a
B200 raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Config, Trace, VrpB200Error, DECODE_MODES


def config_from_args(args: dict, rows_per_cta: int = 0, gemm_path: int = 0) -> Config:
    """args = the reference's merged flag/task dict (configs.py:28-111 + task_specific_params.py)."""
    task_name = args.get("task_name") or ("vrptw" if args["input_dim"] == 5 else "vrp")
    return Config(
        task=_lib.TASK_VRPTW if task_name.startswith("vrptw") else _lib.TASK_VRP,
        n_nodes=args["n_nodes"], n_cust=args["n_cust"], input_dim=args["input_dim"],
        embedding_dim=args.get("embedding_dim", 30), hidden_dim=args.get("hidden_dim", 128),
        decode_len=args["decode_len"], n_glimpses=args.get("n_glimpses", 0),
        use_tanh=int(bool(args.get("use_tanh", False))), mask_glimpses=int(bool(args.get("mask_glimpses", True))),
        mask_pointer=int(bool(args.get("mask_pointer", True))), rows_per_cta=rows_per_cta, gemm_path=gemm_path,
        capacity=float(args["capacity"]), tanh_exploration=float(args.get("tanh_exploration", 10.0)),
        forget_bias=float(args.get("forget_bias", 1.0)), physics=int(bool(args.get("ups_old", False))))


@dataclass
class RolloutOutput:
    idxs: torch.Tensor                      # [T, R] int32
    R: torch.Tensor                         # [R] fp32 tour length
    logprobs: Optional[torch.Tensor] = None  # [T, R]
    logits: Optional[torch.Tensor] = None    # [T, R, N]
    probs: Optional[torch.Tensor] = None
    masks: Optional[torch.Tensor] = None
    final_demand: Optional[torch.Tensor] = None
    final_load: Optional[torch.Tensor] = None
    final_time: Optional[torch.Tensor] = None
    final_mask: Optional[torch.Tensor] = None
    steps: Optional[torch.Tensor] = None     # [grid] int32 decode steps executed per instance block
    beam_parents: Optional[torch.Tensor] = None   # [T, R] int32 (beam search)
    stats: Optional[torch.Tensor] = None     # [4] int64 work counters: node evals, live row-steps, block-steps*rows, blocks


# fused-kernel families (vrpb200_config.gemm_path): the default, the FP32-pipe cross-check, rollout3, rollout4 (split)
GEMM_PATHS = {"auto": 0, "ffma": 1, "tc3": 3, "tc4x": 6}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, args: dict, device: int = 0, rows_per_cta: int = 0, gemm: str = "auto"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise VrpB200Error("no CUDA device: the rollout path has no CPU implementation")
        self.args = dict(args)
        self.cfg = config_from_args(args, rows_per_cta, GEMM_PATHS[gemm])
        self.device = torch.device("cuda", device)
        self.handle = C.c_void_p()
        rc = self.lib.vrpb200_create(C.byref(self.cfg), device, C.byref(self.handle))
        if rc != 0:
            raise VrpB200Error(f"vrpb200_create: {self.lib.vrpb200_last_error(None).decode()}")
        self.tw = self.cfg.task == _lib.TASK_VRPTW
        self.N, self.T, self.D = self.cfg.n_nodes, self.cfg.decode_len, self.cfg.input_dim
        self.E, self.H = self.cfg.embedding_dim, self.cfg.hidden_dim

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise VrpB200Error(f"{what}: {self.lib.vrpb200_last_error(self.handle).decode()} (code {rc})")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, name):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32)
        return t.contiguous()

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.vrpb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ---------------------------------------------------------------------------------
    def set_weights(self, params: Dict[str, np.ndarray]):
        """params: TF-1 variable name -> array (what tf.train.Saver holds, attention_agent.py:79-80)."""
        names = sorted(params)
        arrs = [np.ascontiguousarray(np.asarray(params[k], dtype=np.float32)) for k in names]
        c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        c_ptrs = (C.c_void_p * len(names))(*[a.ctypes.data for a in arrs])
        c_sizes = (C.c_int64 * len(names))(*[a.size for a in arrs])
        self._check(self.lib.vrpb200_set_weights(self.handle, len(names), c_names, c_ptrs, c_sizes), "set_weights")

    # -- the hot path ----------------------------------------------------------------------------
    def rollout(self, input_data, decode_type: str = "greedy", beam_width: int = 1, uniforms=None, uniforms_retry=None,
                seed: int = 0, want_logprobs: bool = False, trace: bool = False, final_state: bool = False,
                want_steps: bool = False, want_stats: bool = False) -> RolloutOutput:
        x = self._f32(input_data, "input_data")
        if x.dim() != 3 or x.shape[1] != self.N or x.shape[2] != self.D:
            raise VrpB200Error(f"input_data must be [B,{self.N},{self.D}], got {tuple(x.shape)}")
        B = x.shape[0]
        mode = DECODE_MODES[decode_type]
        W = beam_width if mode == _lib.BEAM else 1
        R = B * W
        dev = self.device
        idx = torch.empty((self.T, R), dtype=torch.int32, device=dev)
        cost = torch.empty((R,), dtype=torch.float32, device=dev)
        out = RolloutOutput(idxs=idx, R=cost)
        if want_logprobs:
            out.logprobs = torch.empty((self.T, R), dtype=torch.float32, device=dev)
        tr = Trace()
        if trace:
            out.logits = torch.empty((self.T, R, self.N), dtype=torch.float32, device=dev)
            out.probs = torch.empty_like(out.logits)
            out.masks = torch.empty_like(out.logits)
            tr.d_logits, tr.d_probs, tr.d_masks = out.logits.data_ptr(), out.probs.data_ptr(), out.masks.data_ptr()
        if final_state or trace:
            out.final_demand = torch.empty((R, self.N), dtype=torch.float32, device=dev)
            out.final_mask = torch.empty((R, self.N), dtype=torch.float32, device=dev)
            out.final_load = torch.empty((R,), dtype=torch.float32, device=dev)
            out.final_time = torch.zeros((R,), dtype=torch.float32, device=dev)
            tr.d_final_demand, tr.d_final_mask = out.final_demand.data_ptr(), out.final_mask.data_ptr()
            tr.d_final_load, tr.d_final_time = out.final_load.data_ptr(), out.final_time.data_ptr()
        if want_steps:
            out.steps = torch.zeros((B + 1,), dtype=torch.int32, device=dev)   # upper bound on the grid
            tr.d_steps = out.steps.data_ptr()
        if want_stats:
            out.stats = torch.zeros((4,), dtype=torch.int64, device=dev)
            tr.d_stats = out.stats.data_ptr()
        u1 = u2 = None
        if uniforms is not None:
            u1, u2 = self._f32(uniforms, "uniforms"), self._f32(uniforms_retry, "uniforms_retry")
            if tuple(u1.shape) != (self.T, R) or tuple(u2.shape) != (self.T, R):
                raise VrpB200Error(f"uniforms must be [{self.T},{R}]")
        if mode == _lib.BEAM:
            out.beam_parents = torch.empty((self.T, R), dtype=torch.int32, device=dev)
        self._check(self.lib.vrpb200_rollout(
            self.handle, _ptr(x), B, mode, W, _ptr(u1), _ptr(u2), C.c_uint64(seed), _ptr(idx), _ptr(cost),
            _ptr(out.logprobs), _ptr(out.beam_parents), C.byref(tr), self._stream()), "rollout")
        if want_steps:
            out.steps = out.steps[: self.last_launch()["grid"]]
        return out

    def rollout_host(self, input_np: np.ndarray, decode_type: str = "greedy", beam_width: int = 1, seed: int = 0,
                     idx_out: Optional[np.ndarray] = None, cost_out: Optional[np.ndarray] = None, want_idx: bool = True,
                     want_stats: bool = False):
        """Host buffers in, host buffers out (chunked H2D + rollout + D2H inside the call).
        Pass pinned arrays (e.g. views of torch pinned tensors) for full PCIe speed."""
        x = input_np if (isinstance(input_np, np.ndarray) and input_np.dtype == np.float32 and input_np.flags.c_contiguous) \
            else np.ascontiguousarray(input_np, dtype=np.float32)
        B = x.shape[0]
        mode = DECODE_MODES[decode_type]
        W = beam_width if mode == _lib.BEAM else 1
        cost = cost_out if cost_out is not None else np.empty((B * W,), np.float32)
        idx = idx_out if idx_out is not None else (np.empty((self.T, B * W), np.int32) if want_idx else None)
        stats = np.zeros((4,), np.uint64) if want_stats else None
        self._check(self.lib.vrpb200_rollout_host(
            self.handle, C.c_void_p(x.ctypes.data), B, mode, W, C.c_uint64(seed),
            C.c_void_p(idx.ctypes.data) if idx is not None else None, C.c_void_p(cost.ctypes.data),
            C.c_void_p(stats.ctypes.data) if want_stats else None), "rollout_host")
        return (idx, cost, stats) if want_stats else (idx, cost)

    def set_sampling_retry(self, whole_batch: bool) -> None:
        """False (default): per-row retry.  True: the reference's whole-batch single redraw (attention_agent.py:148-167);
        stochastic rollout() calls then synchronise their stream."""
        self._check(self.lib.vrpb200_set_sampling_retry(self.handle, int(bool(whole_batch))), "set_sampling_retry")

    def last_redraws(self) -> int:
        """Steps the last whole-batch stochastic rollout redrew for every row."""
        return int(self.lib.vrpb200_last_redraws(self.handle))

    def pipe_peak(self, kind: str, iters: int = 256) -> float:
        """Measured throughput (ops/s) of the MUFU ('mufu') or FP32 FMA ('fma') pipe on this device."""
        out = C.c_double()
        self._check(self.lib.vrpb200_pipe_peak(self.handle, {"mufu": 0, "fma": 1}[kind], iters, C.byref(out)), "pipe_peak")
        return out.value

    def last_launch(self) -> dict:
        arr = (C.c_int32 * 8)()
        self._check(self.lib.vrpb200_last_launch(self.handle, arr), "last_launch")
        return dict(grid=arr[0], block=arr[1], smem=arr[2], rows_per_cta=arr[3], kernels=arr[4], family=arr[5])

    # -- Env -------------------------------------------------------------------------------------
    def env_reset(self, input_data: torch.Tensor, beam_width: int = 1):
        x = self._f32(input_data, "input_data")
        B = x.shape[0]
        R = B * beam_width
        dev = self.device
        st = dict(demand=torch.empty((R, self.N), device=dev), load=torch.empty((R,), device=dev),
                  time=torch.zeros((R,), device=dev), mask=torch.empty((R, self.N), device=dev),
                  prev_xy=torch.zeros((R, 2), device=dev))
        self._check(self.lib.vrpb200_env_reset(self.handle, _ptr(x), B, beam_width, _ptr(st["demand"]), _ptr(st["load"]),
                                               _ptr(st["time"]), _ptr(st["mask"]), _ptr(st["prev_xy"]), self._stream()),
                    "env_reset")
        return st

    def env_step(self, input_data: torch.Tensor, beam_width: int, state: dict, idx: torch.Tensor,
                 beam_parent: Optional[torch.Tensor] = None):
        x = self._f32(input_data, "input_data")
        B = x.shape[0]
        R = B * beam_width
        dev = self.device
        idx = idx.to(device=dev, dtype=torch.int64).reshape(R).contiguous()
        par = None if beam_parent is None else beam_parent.to(device=dev, dtype=torch.int64).reshape(R).contiguous()
        new = dict(demand=torch.empty((R, self.N), device=dev), load=torch.empty((R,), device=dev),
                   time=torch.zeros((R,), device=dev), mask=torch.empty((R, self.N), device=dev),
                   prev_xy=torch.zeros((R, 2), device=dev))
        d_sat = torch.empty((R,), device=dev)
        self._check(self.lib.vrpb200_env_step(
            self.handle, _ptr(x), B, beam_width, _ptr(idx), _ptr(par), _ptr(state["demand"]), _ptr(state["load"]),
            _ptr(state["time"]), _ptr(state["prev_xy"]), _ptr(new["demand"]), _ptr(new["load"]), _ptr(new["time"]),
            _ptr(new["mask"]), _ptr(new["prev_xy"]), _ptr(d_sat), self._stream()), "env_step")
        return new, d_sat

    # -- stand-alone model pieces ------------------------------------------------------------------
    def embed(self, input_pnt: torch.Tensor) -> torch.Tensor:
        x = self._f32(input_pnt, "input_pnt")
        lead = x.shape[:-1]
        M = int(np.prod(lead))
        out = torch.empty((*lead, self.E), device=self.device)
        self._check(self.lib.vrpb200_embed(self.handle, _ptr(x), M, _ptr(out), self._stream()), "embed")
        return out

    def attention(self, att: int, query, ref, demand, load, time=None):
        q, r, d, l = self._f32(query, "q"), self._f32(ref, "ref"), self._f32(demand, "demand"), self._f32(load, "load")
        t = None if time is None else self._f32(time, "time")
        R = q.shape[0]
        e = torch.empty((R, self.N, self.H), device=self.device)
        logits = torch.empty((R, self.N), device=self.device)
        self._check(self.lib.vrpb200_attention(self.handle, att, R, _ptr(q), _ptr(r), _ptr(d), _ptr(l), _ptr(t), _ptr(e),
                                               _ptr(logits), self._stream()), "attention")
        return e, logits

    def decode_step(self, decoder_inp, context, demand, load, time, mask, c, h):
        """c, h are updated in place; returns logit, prob, logprob [R, N]."""
        x, ctx = self._f32(decoder_inp, "decoder_inp"), self._f32(context, "context")
        R = ctx.shape[0]
        x = x.reshape(R, self.E)
        d, l, m = self._f32(demand, "demand"), self._f32(load, "load"), self._f32(mask, "mask")
        t = None if time is None else self._f32(time, "time")
        logit = torch.empty((R, self.N), device=self.device)
        prob, logprob = torch.empty_like(logit), torch.empty_like(logit)
        self._check(self.lib.vrpb200_decode_step(self.handle, R, _ptr(x), _ptr(ctx), _ptr(d), _ptr(l), _ptr(t), _ptr(m),
                                                 _ptr(c), _ptr(h), _ptr(logit), _ptr(prob), _ptr(logprob), self._stream()),
                    "decode_step")
        return logit, prob, logprob

    def critic(self, input_data: torch.Tensor) -> torch.Tensor:
        """v of build_model('stochastic') (attention_agent.py:243-270): input [B, N, D] -> [B].  The weights given to
        set_weights must include the Critic/* variables."""
        x = self._f32(input_data, "input_data")
        out = torch.empty((x.shape[0],), device=self.device)
        self._check(self.lib.vrpb200_critic(self.handle, _ptr(x), x.shape[0], _ptr(out), self._stream()), "critic")
        return out

    def reinforce_loss(self, R: torch.Tensor, v: torch.Tensor, logprobs: torch.Tensor):
        """build_train_step's (actor_loss, critic_loss) (attention_agent.py:286-294): R, v [rows], logprobs [T, rows]."""
        r, vv, lp = self._f32(R, "R"), self._f32(v, "v"), self._f32(logprobs, "logprobs")
        T, rows = lp.shape
        if r.shape != (rows,) or vv.shape != (rows,):
            raise VrpB200Error(f"R and v must be [{rows}]")
        out = torch.empty((2,), device=self.device)
        self._check(self.lib.vrpb200_reinforce_loss(self.handle, _ptr(r), _ptr(vv), _ptr(lp), T, rows, _ptr(out), self._stream()),
                    "reinforce_loss")
        return out[0], out[1]

    def actor_grad(self, input_data, idxs, weight, shapes: Dict[str, tuple], input_scale=None) -> Dict[str, torch.Tensor]:
        """compute_gradients(actor_loss, Actor variables) (attention_agent.py:289-297): gradients of
        sum_r weight[r] sum_t logprob[t, r] along the tour idxs [T, B] (int32, the rollout's).  shapes: variable name -> shape of
        every variable whose gradient is wanted; returns name -> device tensor of that shape.  input_scale [T, B, E]: the LSTM-input
        dropout scale (keep-mask / keep-probability) the tour was decoded with, None = no dropout."""
        x = self._f32(input_data, "input_data")
        idx = idxs.to(device=self.device, dtype=torch.int32).contiguous()
        w = self._f32(weight, "weight")
        B = x.shape[0]
        if tuple(idx.shape) != (self.args["decode_len"], B) or tuple(w.shape) != (B,):
            raise VrpB200Error(f"idxs must be [{self.args['decode_len']}, {B}] and weight [{B}]")
        xs = None if input_scale is None else self._f32(input_scale, "input_scale")
        if xs is not None and tuple(xs.shape) != (self.args["decode_len"], B, self.args["embedding_dim"]):
            raise VrpB200Error(f"input_scale must be [{self.args['decode_len']}, {B}, {self.args['embedding_dim']}]")
        names = sorted(shapes)
        grads = {k: torch.zeros(tuple(shapes[k]), device=self.device, dtype=torch.float32) for k in names}
        c_names = (C.c_char_p * len(names))(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * len(names))(*[grads[k].data_ptr() for k in names])
        c_sizes = (C.c_int64 * len(names))(*[grads[k].numel() for k in names])
        self._check(self.lib.vrpb200_actor_grad(self.handle, _ptr(x), B, _ptr(idx), _ptr(w), _ptr(xs), len(names), c_names, c_ptrs, c_sizes,
                                                self._stream()), "actor_grad")
        return grads

    def critic_grad(self, input_data, R, shapes: Dict[str, tuple]):
        """compute_gradients(critic_loss, Critic variables) (attention_agent.py:290, 298-299): gradients of mean((R - v)^2);
        returns (name -> device tensor, v [B])."""
        x = self._f32(input_data, "input_data")
        r = self._f32(R, "R")
        B = x.shape[0]
        if tuple(r.shape) != (B,):
            raise VrpB200Error(f"R must be [{B}]")
        names = sorted(shapes)
        grads = {k: torch.zeros(tuple(shapes[k]), device=self.device, dtype=torch.float32) for k in names}
        v = torch.empty((B,), device=self.device, dtype=torch.float32)
        c_names = (C.c_char_p * len(names))(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * len(names))(*[grads[k].data_ptr() for k in names])
        c_sizes = (C.c_int64 * len(names))(*[grads[k].numel() for k in names])
        self._check(self.lib.vrpb200_critic_grad(self.handle, _ptr(x), B, _ptr(r), len(names), c_names, c_ptrs, c_sizes, _ptr(v),
                                                 self._stream()), "critic_grad")
        return grads, v

    def clip_adam(self, variables, grads, m, v, max_norm: float, lr: float, step: int, beta1: float = 0.9, beta2: float = 0.999,
                  eps: float = 1e-8):
        """tf.clip_by_norm per variable + one AdamOptimizer update (attention_agent.py:303-311), in place on the device tensors
        variables[i], m[i], v[i] (lists of equal length; grads[i] is read)."""
        n = len(variables)
        for lst in (grads, m, v):
            if len(lst) != n:
                raise VrpB200Error("clip_adam: lists of different length")
        for a in list(variables) + list(grads) + list(m) + list(v):
            if not (a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()):
                raise VrpB200Error("clip_adam: contiguous float32 device tensors only")
        arr = lambda lst: (C.c_void_p * n)(*[a.data_ptr() for a in lst])
        sizes = (C.c_int64 * n)(*[a.numel() for a in variables])
        self._check(self.lib.vrpb200_clip_adam(self.handle, n, arr(variables), arr(grads), arr(m), arr(v), sizes, float(max_norm), float(lr),
                                               int(step), float(beta1), float(beta2), float(eps), self._stream()), "clip_adam")

    def reward(self, actions: torch.Tensor) -> torch.Tensor:
        a = self._f32(actions, "actions")
        T, R, A = a.shape
        out = torch.empty((R,), device=self.device)
        self._check(self.lib.vrpb200_reward(self.handle, _ptr(a), T, R, A, _ptr(out), self._stream()), "reward")
        return out

    def reward_min_trucks(self, actions: torch.Tensor, depot: torch.Tensor, n_nodes: float, ups: bool = False,
                          ups_old: bool = False) -> torch.Tensor:
        """reward_func with a depot (VRPTW/vrptw_utils.py:384-395, 410-412; ups: UPS/vrptw_ups_utils.py:387-419; ups_old:
        UPS_old/vrptw_ups_utils.py:400-428): actions [T, R, A], depot [R, A] -> [R]."""
        a = self._f32(actions, "actions")
        d = self._f32(depot, "depot")
        T, R, A = a.shape
        if tuple(d.shape) != (R, A):
            raise VrpB200Error(f"depot must be [{R}, {A}]")
        out = torch.empty((R,), device=self.device)
        if ups_old:
            self._check(self.lib.vrpb200_reward_min_trucks_ups_old(self.handle, _ptr(a), _ptr(d), T, R, A, C.c_float(float(n_nodes)),
                                                                   _ptr(out), self._stream()), "reward_min_trucks_ups_old")
            return out
        if ups:
            self._check(self.lib.vrpb200_reward_min_trucks_ups(self.handle, _ptr(a), _ptr(d), T, R, A, _ptr(out), self._stream()),
                        "reward_min_trucks_ups")
            return out
        self._check(self.lib.vrpb200_reward_min_trucks(self.handle, _ptr(a), _ptr(d), T, R, A, C.c_float(float(n_nodes)), _ptr(out),
                                                       self._stream()), "reward_min_trucks")
        return out
