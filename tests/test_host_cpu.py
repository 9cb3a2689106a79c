"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol of include/vrpb200.h, the host
mirror creates exactly the reference's variable names, generators match the oracle's, sharding logic (gloo, 2 ranks)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from oracle import restatement as O
from tests.helpers import load_pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    pkg.build_library()
    hdr = open(os.path.join(ROOT, "include", "vrpb200.h")).read()
    declared = set(re.findall(r"\b(vrpb200_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vrpb200.h but not exported"
    assert set(pkg._lib.EXPORTS) == declared
    assert pkg._lib.load().vrpb200_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    pkg = load_pkg()
    with pytest.raises(pkg.VrpB200Error):
        pkg.Engine(O.default_args("vrptw10"))
    lib = pkg._lib.load()
    cfg = pkg.config_from_args(O.default_args("vrptw10"))
    h = ctypes.c_void_p()
    assert lib.vrpb200_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -2      # VRPB200_E_CUDA
    assert b"no CUDA device" in lib.vrpb200_last_error(None)


def test_create_rejects_bad_configs_before_touching_cuda():
    pkg = load_pkg()
    lib = pkg._lib.load()
    for over in (dict(hidden_dim=64), dict(n_nodes=300, n_cust=299), dict(input_dim=4), dict(n_cust=3), dict(decode_len=0)):
        cfg = pkg.config_from_args(dict(O.default_args("vrptw10"), **over))
        h = ctypes.c_void_p()
        assert lib.vrpb200_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1, over   # VRPB200_E_ARG
        assert lib.vrpb200_last_error(None)


@pytest.mark.parametrize("task,ng", [("vrptw10", 0), ("vrp10", 0), ("vrptw20", 2)])
def test_host_mirror_creates_the_reference_variable_names(task, ng):
    pkg = load_pkg()
    H = pkg.host
    args = pkg.task_specific_params.default_args(task, n_glimpses=ng)
    store = H.VariableStore(seed=1)
    _, Env, reward_func, Actor, Critic = H.load_task_specific_components(args["task_name"], False)
    agent = H.RLAgent(args, None, Env(args), None, reward_func, Actor, Critic, is_train=False, store=store)
    want = O.init_params(O.default_args(task, n_glimpses=ng))
    got = {k: v.shape for k, v in store.vars.items() if k != "decoder_input"}
    assert got == {k: v.shape for k, v in want.items()}
    assert set(agent.model_params()) == set(want)
    # assigning a checkpoint by name replaces the values; an unknown name is rejected
    store.assign(want)
    np.testing.assert_array_equal(agent.model_params()["Actor/Decoder/Attention/v"], want["Actor/Decoder/Attention/v"])
    with pytest.raises(KeyError):
        store.assign({"Actor/Nope": np.zeros(3)})


def test_generators_match_the_reference_restatement():
    pkg = load_pkg()
    for task in ("vrp10", "vrptw20", "vrp100"):
        t = O.TASKS[task]
        np.testing.assert_array_equal(pkg.data.create_dataset(t.task_name, 50, t.n_cust, 7), O.create_dataset(task, 50, 7))
    d = pkg.data.create_vrptw_ups_dataset(200, 100, 3)
    assert d.shape == (200, 101, 5)
    np.testing.assert_allclose(d[:, -1], np.tile([0, 0, 0, 1000 * 0.1567, 0], [200, 1]), rtol=0, atol=0); assert True                      # depot row, UPS/vrptw_ups_utils.py:52
    assert (d[:, :-1, 4] >= 1).all() and (d[:, :-1, 4] == np.round(d[:, :-1, 4])).all()
    assert (d[:, :-1, 2] >= 0).all() and (d[:, :-1, 3] <= 950 * 0.1567 + 1e-9).all() and (d[:, :-1, 2] < d[:, :-1, 3]).all()
    np.testing.assert_array_equal(d, pkg.data.create_vrptw_ups_dataset(200, 100, 3))


def test_glorot_params_have_reference_names_and_bounds():
    pkg = load_pkg()
    args = pkg.task_specific_params.default_args("vrptw50")
    p = pkg.data.glorot_params(args, seed=5)
    want = O.init_params(O.default_args("vrptw50"))
    assert {k: v.shape for k, v in p.items()} == {k: v.shape for k, v in want.items()}
    k = p["Actor/Decoder/Attention/proj_q/kernel"]
    assert np.abs(k).max() <= np.sqrt(6.0 / 256) and all(np.all(v == 0) for n, v in p.items() if n.endswith("bias"))


def test_task_table_matches_reference_sizes():
    pkg = load_pkg()
    for name, t in O.TASKS.items():
        assert tuple(pkg.task_specific_params.task_lst[name]) == tuple(t)


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = load_pkg()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = pkg.distributed.shard_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 2.0          # stand-in for the shard's tour costs
    out = pkg.distributed.gather_costs(local, n, dst=0)
    if rank == 0:
        q.put(out.tolist())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_sharding_and_cost_gather_two_ranks_gloo(n):
    import torch.multiprocessing as mp
    pkg = load_pkg()
    assert [pkg.distributed.shard_range(11, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 11)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [2.0 * i for i in range(n)]


def test_dataset_files_round_trip_in_the_reference_format(tmp_path):
    """create_VRP(TW)_dataset's on-disk format (VRP/vrp_utils.py:36-53, VRPTW/vrptw_utils.py:38-67): name
    '<task>-size-B-len-N-<type>.txt', one instance per line; a second call loads the file instead of generating."""
    pkg = load_pkg()
    d = pkg.data
    for task, D in (("vrp", 3), ("vrptw", 5)):
        a = d.load_or_create_dataset(task, 7, 10, str(tmp_path), seed=5, data_type="test")
        f = tmp_path / f"{task}-size-7-len-11-test.txt"
        assert f.exists()
        rows = f.read_text().strip().split("\n")
        assert len(rows) == 7 and len(rows[0].split(" ")) == 11 * D
        b = d.load_or_create_dataset(task, 7, 10, str(tmp_path), seed=999, data_type="test")   # other seed: the FILE wins
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, d.create_dataset(task, 7, 10, 5))
    assert d.dataset_file_name("vrptw", 3, 101, "test", ups=True) == "vrptw-ups-size-3-len-101-test.txt"


def test_result_file_format(tmp_path):
    """_output_results (model/attention_agent.py:416-459): 'x y b e ' per visited node from the depot on, cut when the
    last customer has been seen, closed with the depot."""
    pkg = load_pkg()
    depot = [0.5, 0.5, 0.0, 8.0]
    c1, c2 = [0.1, 0.2, 1.0, 2.0], [0.9, 0.8, 3.0, 4.0]
    route = [depot, c1, depot, c2, depot, depot]          # n_nodes = 3: two customers
    f = tmp_path / "res.txt"
    pkg.data.write_results(str(f), [route], "vrptw", 3)
    toks = f.read_text().strip().split(" ")
    vals = [float(t) for t in toks]
    assert vals == depot + c1 + depot + c2 + depot        # stops after the second customer, then the depot once
    pkg.data.write_results(str(f), [[d[:2] for d in route]], "vrp", 3)
    assert [float(t) for t in f.read_text().strip().split(" ")] == depot[:2] + c1[:2] + depot[:2] + c2[:2] + depot[:2]


def test_host_selection_rules_match_the_oracle_on_cpu_tensors():
    """The selection rules of the step-by-step route (host.RLAgent._pick_*; pure torch, so they run without a GPU) against the
    oracle's as-written restatement: inverse-CDF draw incl. the whole-batch retry and the 'no node found -> node 0' rule,
    beam top-k with ties to the lower flat index, and the back-tracking of beam parents."""
    H = load_pkg().host
    rnd = np.random.RandomState(0)
    R, N = 64, 11
    prob = rnd.dirichlet(np.ones(N) * 0.3, size=R).astype(np.float32)
    prob[:, 3] = 0.0
    prob /= prob.sum(1, keepdims=True)
    u1, u2 = rnd.uniform(size=(1, R)).astype(np.float32), rnd.uniform(size=(1, R)).astype(np.float32)
    agent = object.__new__(H.RLAgent)                      # the rules use no agent state
    for force_fail in (False, True):
        a1 = u1.copy()
        if force_fail:
            a1[0, 7] = 2.0                                  # beyond any CDF: every row must redraw from u2
        idx, par = agent._pick_sampled(torch.from_numpy(prob), 0, a1, u2)
        tmp, want = O.multinomial_as_written(prob, a1[0], N)
        if (tmp.sum(1) > 10000 * N - 1).any():
            tmp, want = O.multinomial_as_written(prob, u2[0], N)
        assert par is None and np.array_equal(idx[:, 0].numpy(), want)
    both = np.full((1, R), 2.0, np.float32)
    idx, _ = agent._pick_sampled(torch.from_numpy(prob), 0, both, both)
    assert (idx == 0).all()                                # argmin of an all-10000 row
    # beam: B = 8 instances, W = 4 beams, duplicate candidates force the tie rule
    B, W = 8, 4
    p = rnd.dirichlet(np.ones(N), size=B * W).astype(np.float32)
    p[:, 5] = p[:, 2]                                       # exact ties inside every row
    carried = torch.from_numpy(np.log(rnd.dirichlet(np.ones(W), size=B).T.reshape(-1, 1)).astype(np.float32))
    node, parent, pp, score = agent._pick_beam(torch.from_numpy(p), 1, B, W, carried)
    lbp = np.log(p) + carried.numpy()
    cand = np.concatenate(np.split(lbp, W, axis=0), 1)
    order = np.argsort(-cand, axis=1, kind="stable")[:, :W]
    assert np.array_equal(node[:, 0].numpy(), (order.T.reshape(-1) % N))
    assert np.array_equal(parent[:, 0].numpy(), (order.T.reshape(-1) // N))
    np.testing.assert_array_equal(score[:, 0].numpy(), np.take_along_axis(cand, order, 1).T.reshape(-1))
    src = np.tile(np.arange(B), W) + B * parent[:, 0].numpy()
    np.testing.assert_array_equal(pp.numpy(), p[src])
    node0, parent0, _, _ = agent._pick_beam(torch.from_numpy(p), 0, B, W, None)
    assert (parent0 == 0).all()                             # step 0 looks at beam 0 only
    # back-tracking: row r of every step must belong to the beam that ends in row r
    T = 6
    parents = [torch.from_numpy(rnd.randint(0, W, size=(B * W, 1))) for _ in range(T)]
    steps = [torch.arange(B * W) + 1000 * t for t in range(T)]
    bt = H.RLAgent._backtrack(steps, parents, B)
    for r in range(B * W):
        at = r
        for k in reversed(range(T)):
            assert int(bt[k][r]) == 1000 * k + at
            at = (r % B) + B * int(parents[k][at])
