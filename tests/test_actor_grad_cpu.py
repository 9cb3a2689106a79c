"""The algebra of the CUDA back-propagation kernel, checked WITHOUT a GPU: mit_rl-vrptw_b200/csrc/actor_grad_row.h is written
so that the same source compiles as plain C++ (tests/native/actor_grad_host.cpp, g++); its per-row results, reduced here with
numpy exactly as train_kernels.cuh reduces them on the device, are compared with torch autograd through the as-written
forward graph (oracle/train_torch.py).  The GPU twin of this test is tests/test_gpu_train.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import restatement as O
from oracle import train_torch as TT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ATT = "Actor/Decoder/Attention"
LSTM = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell"


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libactor_grad_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "native", "actor_grad_host.cpp")], check=True)
    return C.CDLL(out)


def weight_arrays(params, tw):
    g = lambda n: np.ascontiguousarray(params[n], dtype=np.float32).reshape(-1)
    z = np.zeros(1, np.float32)
    order = ["Actor/Embedding/conv1d/kernel", "Actor/Embedding/conv1d/bias", LSTM + "/kernel", LSTM + "/bias", ATT + "/v"]
    for blk in ("d", "ld", "t"):
        e, p = ("emb_time", "proj_t") if blk == "t" else ("emb_" + blk, "proj_" + blk)
        order += [f"{ATT}/{e}/kernel", f"{ATT}/{e}/bias", f"{ATT}/{p}/kernel", f"{ATT}/{p}/bias"]
    order += [f"{ATT}/proj_q/kernel", f"{ATT}/proj_q/bias", f"{ATT}/proj_ref/kernel", f"{ATT}/proj_ref/bias"]
    arrs = [g(n) if n in params else z for n in order]
    E = params["Actor/Embedding/conv1d/kernel"].shape[-1]
    arrs.append(np.ascontiguousarray(np.asarray(params[LSTM + "/kernel"], np.float32).reshape(E + 128, 512).T).reshape(-1))   # lstm_kT
    arrs.append(np.ascontiguousarray(np.asarray(params[ATT + "/proj_q/kernel"], np.float32).reshape(128, 128).T).reshape(-1))  # proj_q_kT
    return order, arrs


def run_host(lib, params, args, data, idxs, weight, st, input_scale=None):
    tw = args["task_name"] == "vrptw"
    N, E, D, T, H = args["n_nodes"], args["embedding_dim"], args["input_dim"], args["decode_len"], 128
    R = data.shape[0]
    order, warr = weight_arrays(params, tw)
    fp = C.POINTER(C.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    wp = (fp * len(warr))(*[ptr(a) for a in warr])
    x = np.ascontiguousarray(data, np.float32)
    idx = np.ascontiguousarray(idxs, np.int32)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    demand, mask, load, time = f32(st["demand"]), f32(st["mask"]), f32(st["load"]), f32(st["time"])
    wt = f32(weight)
    out = dict(dgates=np.zeros((R, T, 4 * H), np.float32), dq=np.zeros((R, T, H), np.float32), Dsum=np.zeros((R, N, H), np.float32),
               demb=np.zeros((R, N, E), np.float32), sums=np.zeros((R, 5, H), np.float32), hs=np.zeros((R, T + 1, H), np.float32),
               xin=np.zeros((R, T, E), np.float32), emb=np.zeros((R, N, E), np.float32))
    small_names = [ATT + "/v", ATT + "/proj_q/bias", ATT + "/proj_ref/bias"]
    blocks = [("emb_d", "proj_d"), ("emb_ld", "proj_ld"), ("emb_time", "proj_t")]
    small_names += [f"{ATT}/{e}/kernel" for e, _ in blocks] + [f"{ATT}/{e}/bias" for e, _ in blocks]
    small_names += [f"{ATT}/{p}/kernel" for _, p in blocks] + [f"{ATT}/{p}/bias" for _, p in blocks]
    small = {n: np.zeros(np.asarray(params[n]).size if n in params else 1, np.float32) for n in small_names}
    sp = (fp * len(small_names))(*[ptr(small[n]) if n in params else None for n in small_names])
    lib.actor_grad_host.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_float, C.c_float, C.c_longlong] + [C.c_void_p] * 17
    xs = None if input_scale is None else f32(input_scale)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.actor_grad_host(C.cast(wp, C.c_void_p), N, E, D, T, int(tw), int(args["use_tanh"]), int(args["mask_pointer"]),
                             float(args["tanh_exploration"]), float(args["forget_bias"]), R, vp(x), vp(idx), vp(wt), vp(demand),
                             vp(mask), vp(load), vp(time) if tw else None, vp(out["dgates"]), vp(out["dq"]), vp(out["Dsum"]),
                             vp(out["demb"]), vp(out["sums"]), vp(out["hs"]), vp(out["xin"]), vp(out["emb"]), C.cast(sp, C.c_void_p),
                             None if xs is None else vp(xs))
    assert rc == 0
    # the device-wide reductions of train_kernels.cuh (outer_reduce / col_reduce), in numpy
    g = {n: small[n].reshape(np.asarray(params[n]).shape) for n in small_names if n in params}
    lin = np.concatenate([out["xin"], out["hs"][:, :T]], 2).reshape(R * T, E + H)
    dg = out["dgates"].reshape(R * T, 4 * H)
    g[LSTM + "/kernel"] = lin.T @ dg
    g[LSTM + "/bias"] = dg.sum(0)
    g[ATT + "/proj_q/kernel"] = out["hs"][:, 1:].reshape(R * T, H).T @ out["dq"].reshape(R * T, H)
    g[ATT + "/proj_ref/kernel"] = (out["emb"].reshape(R * N, E).T @ out["Dsum"].reshape(R * N, H))[None]
    g["Actor/Embedding/conv1d/kernel"] = (x[:, :, : D - 1].reshape(R * N, D - 1).T @ out["demb"].reshape(R * N, E))[None]
    g["Actor/Embedding/conv1d/bias"] = out["demb"].reshape(R * N, E).sum(0)
    return g


def make_case(task, B, seed, **over):
    args = O.default_args(task, **over)
    params = O.init_params(args, seed=seed, bias_scale=0.2)
    data = O.create_dataset(task, B, seed=seed + 1).astype(np.float32)
    rnd = np.random.RandomState(seed + 2)
    u = rnd.uniform(size=(args["decode_len"], B)).astype(np.float32)
    ref = O.rollout(params, args, data, "stochastic", uniforms=u, uniforms_retry=u[::-1].copy())
    weight = ((ref.R - ref.R.mean()) / B).astype(np.float32)      # (R - v) / B with a constant baseline
    return args, params, data, ref.idxs.astype(np.int32), weight


def check(got, want, names):
    for n in names:
        assert want[n] is not None, n
        scale = np.abs(want[n]).max()
        assert scale > 0, n
        np.testing.assert_allclose(got[n], want[n], rtol=1e-4, atol=2e-5 * scale, err_msg=n)


@pytest.mark.parametrize("task,over", [("vrptw10", {}), ("vrp10", {}), ("vrptw20", {"use_tanh": True}),
                                       ("vrptw10", {"mask_pointer": False})])
def test_actor_gradient_algebra_matches_autograd(hostlib, task, over):
    args, params, data, idxs, weight = make_case(task, 6, 31, **over)
    loss, want, st, _ = TT.actor_grads(params, args, data, idxs, weight)
    got = run_host(hostlib, params, args, data, idxs, weight, st)
    names = [k for k in params if k.startswith("Actor/")]
    assert set(names) == set(got), set(names) ^ set(got)
    check(got, want, names)


def dropout_scale(args, B, seed, rate=0.3):
    """keep-mask / keep-probability of tf.nn.dropout on the LSTM input, [T, B, E]."""
    keep = 1.0 - rate
    rnd = np.random.RandomState(seed)
    return ((rnd.uniform(size=(args["decode_len"], B, args["embedding_dim"])) < keep) / keep).astype(np.float32)


@pytest.mark.parametrize("task", ["vrptw10", "vrp10"])
def test_actor_gradient_with_lstm_input_dropout(hostlib, task):
    """DropoutWrapper(input_keep_prob = 1 - dropout) (shared/decode_step.py:177-179) given the mask: the scaled LSTM inputs in the
    forward sweep and the same scale on the way back into the embedding."""
    args, params, data, idxs, weight = make_case(task, 6, 33)
    scale = dropout_scale(args, 6, 5)
    loss, want, st, _ = TT.actor_grads(params, args, data, idxs, weight, scale)
    _, plain, _, _ = TT.actor_grads(params, args, data, idxs, weight)
    got = run_host(hostlib, params, args, data, idxs, weight, st, scale)
    names = [k for k in params if k.startswith("Actor/")]
    check(got, want, names)
    assert np.abs(want[LSTM + "/kernel"] - plain[LSTM + "/kernel"]).max() > 1e-3 * np.abs(plain[LSTM + "/kernel"]).max()   # the mask matters


def test_actor_gradient_is_additive_over_rows_and_linear_in_the_weights(hostlib):
    """Size-independent properties of the gradient kernel's source: the gradient of a batch is the sum of the gradients of its
    halves (rows are independent; only the device-wide reductions join them) and it is linear in the per-row weights."""
    args, params, data, idxs, weight = make_case("vrptw10", 8, 35)
    _, _, st, _ = TT.actor_grads(params, args, data, idxs, weight)
    whole = run_host(hostlib, params, args, data, idxs, weight, st)
    half = lambda sl: run_host(hostlib, params, args, data[sl], idxs[:, sl], weight[sl], {k: v[:, sl] for k, v in st.items()})
    a, b = half(slice(0, 3)), half(slice(3, 8))
    twice = run_host(hostlib, params, args, data, idxs, 2.0 * weight, st)
    for n in whole:
        scale = np.abs(whole[n]).max()
        np.testing.assert_allclose(a[n] + b[n], whole[n], rtol=1e-5, atol=1e-6 * scale, err_msg=n)
        np.testing.assert_allclose(twice[n], 2.0 * whole[n], rtol=1e-6, atol=1e-7 * scale, err_msg=n)
    zero = run_host(hostlib, params, args, data, idxs, 0.0 * weight, st)
    assert all(np.abs(g).max() == 0 for g in zero.values())
