"""The algebra of the critic's back-propagation kernel (mit_rl-vrptw_b200/csrc/critic_grad_row.h) checked WITHOUT a GPU: the
same source built as plain C++ (tests/native/actor_grad_host.cpp), reduced with numpy as train_kernels.cuh reduces it on
the device, against torch autograd through the as-written critic graph (oracle/train_torch.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import restatement as O
from oracle import train_torch as TT
from tests.test_actor_grad_cpu import hostlib  # noqa: F401  (fixture)

BLOCK_ORDER = ["v", "emb_d/kernel", "emb_d/bias", "proj_d/kernel", "proj_d/bias", "proj_q/kernel", "proj_q/bias", "proj_e/kernel",
               "proj_e/bias", "emb_begin_tw/kernel", "emb_begin_tw/bias", "proj_end_tw/kernel", "proj_end_tw/bias"]
SMALL_ORDER = ["v", "proj_q/bias", "emb_d/kernel", "emb_d/bias", "proj_d/kernel", "proj_d/bias", "emb_begin_tw/kernel",
               "emb_begin_tw/bias", "proj_end_tw/kernel", "proj_end_tw/bias"]


def run_host(lib, params, args, data, Rw):
    tw = args["task_name"] == "vrptw"
    N, E, D, P, H = args["n_nodes"], args["embedding_dim"], args["input_dim"], args["n_process_blocks"], 128
    R = data.shape[0]
    fp = C.POINTER(C.c_float)
    z = np.zeros(1, np.float32)
    keep = []

    def ptr(name):
        a = np.ascontiguousarray(params[name], np.float32).reshape(-1) if name in params else z
        keep.append(a)
        return a.ctypes.data_as(fp)
    wl = [ptr("Actor/Embedding/conv1d/kernel"), ptr("Actor/Embedding/conv1d/bias")]
    for b in range(8):
        wl += [ptr(f"Critic/Process/P{b}/{n}") for n in BLOCK_ORDER]
    wl += [ptr("Critic/Linear/L1/kernel"), ptr("Critic/Linear/L1/bias"), ptr("Critic/Linear/L2/kernel"), ptr("Critic/Linear/L2/bias")]
    assert len(wl) == 110
    wp = (fp * 110)(*wl)
    x = np.ascontiguousarray(data, np.float32)
    Rw = np.ascontiguousarray(Rw, np.float32)
    out = dict(hyin=np.zeros((R, P + 1, H), np.float32), De=np.zeros((R, P, N, H), np.float32), sums=np.zeros((R, P, 4, H), np.float32),
               dq=np.zeros((R, P, H), np.float32), l1=np.zeros((R, H), np.float32), dl1=np.zeros((R, H), np.float32),
               vdv=np.zeros((R, 2), np.float32), emb=np.zeros((R, N, E), np.float32))
    small = {}
    sl = []
    for b in range(P):
        for n in SMALL_ORDER:
            name = f"Critic/Process/P{b}/{n}"
            if name in params:
                small[name] = np.zeros(np.asarray(params[name]).size, np.float32)
                sl.append(small[name].ctypes.data_as(fp))
            else:
                sl.append(None)
    sp = (fp * len(sl))(*sl)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.critic_grad_host.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_longlong] + [C.c_void_p] * 11
    rc = lib.critic_grad_host(C.cast(wp, C.c_void_p), N, E, D, P, int(tw), R, vp(x), vp(Rw), vp(out["hyin"]), vp(out["De"]), vp(out["sums"]),
                              vp(out["dq"]), vp(out["l1"]), vp(out["dl1"]), vp(out["vdv"]), vp(out["emb"]), C.cast(sp, C.c_void_p))
    assert rc == 0
    g = {n: a.reshape(np.asarray(params[n]).shape) for n, a in small.items()}
    emb = out["emb"].reshape(R * N, E)
    for b in range(P):
        pre = f"Critic/Process/P{b}"
        De = out["De"][:, b].reshape(R * N, H)
        g[pre + "/proj_e/kernel"] = (emb.T @ De)[None]
        g[pre + "/proj_e/bias"] = De.sum(0)
        g[pre + "/proj_q/kernel"] = out["hyin"][:, b].T @ out["dq"][:, b]
    g["Critic/Linear/L1/kernel"] = out["hyin"][:, P].T @ out["dl1"]
    g["Critic/Linear/L1/bias"] = out["dl1"].sum(0)
    g["Critic/Linear/L2/kernel"] = (out["l1"].T @ out["vdv"][:, 1])[:, None]
    g["Critic/Linear/L2/bias"] = out["vdv"][:, 1].sum(keepdims=True)
    return g, out["vdv"][:, 0]


def make_case(task, B, seed):
    args = O.default_args(task)
    params = O.init_params(args, seed=seed, bias_scale=0.2)
    params.update(O.init_critic_params(args, seed=seed + 100, bias_scale=0.2))
    data = O.create_dataset(task, B, seed=seed + 1).astype(np.float32)
    Rw = np.random.RandomState(seed + 2).uniform(3.0, 9.0, size=B).astype(np.float32)
    return args, params, data, Rw


def check(got, want, names, ref_scale):
    """1e-4 relative + 2e-5 of the variable's largest gradient entry; a variable whose gradient is identically zero (block 0's
    proj_q kernel: its input is the constant hy = 0) must be zero here as well."""
    for n in names:
        assert want[n] is not None, n
        scale = np.abs(want[n]).max()
        if scale == 0:
            assert np.abs(got[n]).max() <= 1e-7 * ref_scale, n
        else:
            np.testing.assert_allclose(got[n], want[n], rtol=1e-4, atol=2e-5 * scale, err_msg=n)


@pytest.mark.parametrize("task", ["vrptw10", "vrp10", "vrptw20"])
def test_critic_gradient_algebra_matches_autograd(hostlib, task):  # noqa: F811
    args, params, data, Rw = make_case(task, 7, 41)
    loss, want, v = TT.critic_grads(params, args, data, Rw)
    got, v_got = run_host(hostlib, params, args, data, Rw)
    np.testing.assert_allclose(v_got, v, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v, O.critic_value(params, args, data), rtol=1e-5, atol=1e-6)     # the torch twin = the pinned numpy oracle
    names = [k for k in params if k.startswith("Critic/")]
    assert set(names) == set(got), set(names) ^ set(got)
    check(got, want, names, max(np.abs(w).max() for w in want.values()))
    assert np.abs(want["Critic/Process/P0/proj_q/kernel"]).max() == 0 and np.abs(want["Critic/Process/P1/proj_q/kernel"]).max() > 0


def test_critic_gradient_is_the_slope_of_the_loss(hostlib):  # noqa: F811
    """Independent of autograd: the kernel source's own value v gives the loss mean((R - v)^2); its central finite difference
    along a random direction of a variable equals the directional derivative formed from the kernel source's gradient."""
    args, params, data, Rw = make_case("vrptw10", 5, 43)
    got, v0 = run_host(hostlib, params, args, data, Rw)
    rnd = np.random.RandomState(0)
    for name in ("Critic/Process/P1/proj_q/kernel", "Critic/Process/P0/emb_begin_tw/kernel", "Critic/Process/P2/proj_e/kernel",
                 "Critic/Linear/L1/kernel", "Critic/Process/P1/v"):
        direction = rnd.normal(size=params[name].shape)
        direction /= np.linalg.norm(direction)
        eps = 2e-2
        loss = []
        for sgn in (+1.0, -1.0):
            p = dict(params)
            p[name] = (params[name].astype(np.float64) + sgn * eps * direction).astype(np.float32)
            _, v = run_host(hostlib, p, args, data, Rw)
            loss.append(np.mean((Rw.astype(np.float64) - v.astype(np.float64)) ** 2))
        slope = (loss[0] - loss[1]) / (2 * eps)
        want = float((got[name].astype(np.float64) * direction).sum())
        assert abs(slope - want) <= 2e-2 * abs(want) + 1e-4, (name, slope, want)


def test_critic_gradient_is_additive_over_rows(hostlib):  # noqa: F811
    """Rows are independent up to the device-wide reductions and the 2 / B factor of the mean."""
    args, params, data, Rw = make_case("vrp10", 8, 45)
    whole, _ = run_host(hostlib, params, args, data, Rw)
    a, _ = run_host(hostlib, params, args, data[:3], Rw[:3])
    b, _ = run_host(hostlib, params, args, data[3:], Rw[3:])
    ref = max(np.abs(g).max() for g in whole.values())      # block 0's gradients are tiny sums of cancelling terms: absolute floor
    for n in whole:
        np.testing.assert_allclose((3 * a[n] + 5 * b[n]) / 8, whole[n], rtol=1e-5, atol=1e-6 * ref, err_msg=n)
