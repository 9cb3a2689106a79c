"""ctypes binding of libvrpb200.so (include/vrpb200.h) and its in-tree build recipe.

The library is the product: there is no Python/CPU fallback.  ``load()`` raises if the shared
object is missing (run ``python __graft_entry__.py build`` or ``build_library()``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libvrpb200.so")
HEADER = os.path.join(ROOT, "include", "vrpb200.h")

TASK_VRP, TASK_VRPTW = 0, 1
GREEDY, STOCHASTIC, BEAM = 0, 1, 2
DECODE_MODES = {"greedy": GREEDY, "stochastic": STOCHASTIC, "beam_search": BEAM}


class Config(C.Structure):
    """struct vrpb200_config"""
    _fields_ = [(n, C.c_int32) for n in (
        "task", "n_nodes", "n_cust", "input_dim", "embedding_dim", "hidden_dim", "decode_len", "n_glimpses",
        "use_tanh", "mask_glimpses", "mask_pointer", "rows_per_cta", "gemm_path")] + [
        ("capacity", C.c_float), ("tanh_exploration", C.c_float), ("forget_bias", C.c_float), ("physics", C.c_int32)]


class Trace(C.Structure):
    """struct vrpb200_trace (device pointers)"""
    _fields_ = [(n, C.c_void_p) for n in (
        "d_logits", "d_probs", "d_masks", "d_final_demand", "d_final_load", "d_final_time", "d_final_mask", "d_steps", "d_stats")]


EXPORTS = (
    "vrpb200_create", "vrpb200_destroy", "vrpb200_last_error", "vrpb200_version", "vrpb200_set_weights",
    "vrpb200_rollout", "vrpb200_rollout_host", "vrpb200_env_reset", "vrpb200_env_step", "vrpb200_embed",
    "vrpb200_attention", "vrpb200_decode_step", "vrpb200_reward", "vrpb200_reward_min_trucks", "vrpb200_reward_min_trucks_ups", "vrpb200_reward_min_trucks_ups_old",
    "vrpb200_last_launch",
    "vrpb200_pipe_peak", "vrpb200_critic", "vrpb200_reinforce_loss", "vrpb200_set_sampling_retry", "vrpb200_last_redraws",
    "vrpb200_actor_grad", "vrpb200_critic_grad", "vrpb200_clip_adam",
)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
BUILD_DIR = os.path.join(HERE, "build")


def _translation_units():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = None) -> str:
    """nvcc -> mit_rl-vrptw_b200/lib/libvrpb200.so (sm_100a only; cross-compiles without a GPU).
    One object per translation unit of csrc/ (the kernel families compile in parallel), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    out_path = out_path or LIB_PATH
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER]
    if not force and not extra_flags and os.path.exists(out_path) and all(os.path.getmtime(out_path) >= os.path.getmtime(s) for s in srcs):
        return out_path
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    tag = "" if out_path == LIB_PATH else "_" + os.path.splitext(os.path.basename(out_path))[0]
    bdir = BUILD_DIR + tag
    os.makedirs(bdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers_mtime = max(os.path.getmtime(s) for s in srcs if not s.endswith(".cu"))

    def compile_one(tu):
        src, obj = os.path.join(CSRC, tu), os.path.join(bdir, tu[:-3] + ".o")
        if not force and not extra_flags and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), headers_mtime):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {tu}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, _translation_units()))
    if verbose:
        for _, log in results:
            print(log)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path] + [o for o, _ in results],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return out_path


_lib = None


def load() -> C.CDLL:
    """dlopen the library and declare every prototype of include/vrpb200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                           "(there is no CPU fallback for the rollout path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p
    lib.vrpb200_create.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    lib.vrpb200_destroy.argtypes = [vp]
    lib.vrpb200_last_error.argtypes = [vp]
    lib.vrpb200_last_error.restype = C.c_char_p
    lib.vrpb200_version.argtypes = []
    lib.vrpb200_set_weights.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(i64)]
    lib.vrpb200_rollout.argtypes = [vp, f32p, i64, i32, i32, f32p, f32p, u64, vp, f32p, f32p, vp, C.POINTER(Trace), vp]
    lib.vrpb200_rollout_host.argtypes = [vp, f32p, i64, i32, i32, u64, vp, f32p, vp]
    lib.vrpb200_env_reset.argtypes = [vp, f32p, i64, i32, f32p, f32p, f32p, f32p, f32p, vp]
    lib.vrpb200_env_step.argtypes = [vp, f32p, i64, i32, vp, vp, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, vp]
    lib.vrpb200_embed.argtypes = [vp, f32p, i64, f32p, vp]
    lib.vrpb200_attention.argtypes = [vp, i32, i64, f32p, f32p, f32p, f32p, f32p, f32p, f32p, vp]
    lib.vrpb200_decode_step.argtypes = [vp, i64, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, vp]
    lib.vrpb200_reward.argtypes = [vp, f32p, i64, i64, i32, f32p, vp]
    lib.vrpb200_reward_min_trucks.argtypes = [vp, f32p, f32p, i64, i64, i32, C.c_float, f32p, vp]
    lib.vrpb200_reward_min_trucks_ups.argtypes = [vp, f32p, f32p, i64, i64, i32, f32p, vp]
    lib.vrpb200_reward_min_trucks_ups_old.argtypes = [vp, f32p, f32p, i64, i64, i32, C.c_float, f32p, vp]
    lib.vrpb200_set_sampling_retry.argtypes = [vp, i32]
    lib.vrpb200_last_redraws.argtypes = [vp]
    lib.vrpb200_critic.argtypes = [vp, f32p, i64, f32p, vp]
    lib.vrpb200_reinforce_loss.argtypes = [vp, f32p, f32p, f32p, i64, i64, f32p, vp]
    lib.vrpb200_actor_grad.argtypes = [vp, f32p, i64, vp, f32p, f32p, i32, vp, vp, vp, vp]
    lib.vrpb200_critic_grad.argtypes = [vp, f32p, i64, f32p, i32, vp, vp, vp, f32p, vp]
    lib.vrpb200_clip_adam.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.c_float, C.c_float, i64, C.c_float, C.c_float, C.c_float, vp]
    lib.vrpb200_last_launch.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.vrpb200_pipe_peak.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name != "vrpb200_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


class VrpB200Error(RuntimeError):
    pass
