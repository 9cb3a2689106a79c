"""The reference-named host classes (Env, Attention*Actor, RNNDecodeStep, LinearEmbedding, reward_func, RLAgent)
on the GPU against the golden fixtures and the oracle.  These tests read like the checks one would write
against the reference's own classes."""
import numpy as np
import pytest

from oracle import restatement as O
from tests.helpers import GOLDEN_INDEX, load_golden, load_pkg

pytestmark = pytest.mark.gpu


def _agent(args, params, task_name=None):
    import torch
    pkg = load_pkg()
    H = pkg.host
    store = H.VariableStore(seed=0)
    _, Env, reward_func, Actor, Critic = H.load_ups_old_components() if args.get("ups_old") else \
        H.load_task_specific_components(task_name or args["task_name"], False)
    env = Env(args)
    agent = H.RLAgent(args, None, env, None, reward_func, Actor, Critic, is_train=False, store=store)
    store.assign(params)
    return agent, env, torch, H


@pytest.mark.parametrize("name,task,W", [("env_vrptw10", "vrptw10", 1), ("env_vrp10", "vrp10", 1),
                                         ("env_vrptw10_beam", "vrptw10", 3)])
def test_env_reset_step_bit_exact(name, task, W):
    import torch
    fx, _, _ = load_golden(name)
    pkg = load_pkg()
    args = pkg.task_specific_params.default_args(task)
    _, Env, _, _, _ = pkg.host.load_task_specific_components(args["task_name"], False)
    env = Env(args)
    env.input_data = torch.from_numpy(fx["data"].astype(np.float32))
    st = env.reset(W)
    np.testing.assert_array_equal(st.mask.cpu().numpy(), fx["mask"][0])
    np.testing.assert_array_equal(st.load.cpu().numpy(), fx["load"][0])
    assert env.batch_beam == fx["data"].shape[0] * W
    for t in range(fx["idx_seq"].shape[0]):
        par = torch.from_numpy(fx["parent_seq"][t][:, None]) if "parent_seq" in fx else None
        st = env.step(torch.from_numpy(fx["idx_seq"][t][:, None]), par)
        np.testing.assert_array_equal(st.mask.cpu().numpy(), fx["mask"][t + 1], err_msg=f"mask, step {t}")
        np.testing.assert_array_equal(st.demand.cpu().numpy(), fx["demand"][t + 1])
        np.testing.assert_array_equal(st.load.cpu().numpy(), fx["load"][t + 1])
        np.testing.assert_array_equal(st.d_sat.cpu().numpy(), fx["d_sat"][t + 1])
        if args["task_name"] == "vrptw":
            np.testing.assert_array_equal(st.time.cpu().numpy(), fx["time"][t + 1])


@pytest.mark.parametrize("name,W", [("env_upsold_vrptw10", 1), ("env_upsold_vrptw10_beam", 2)])
def test_ups_old_env_matches_reference_fixture(name, W):
    """host.VRPTWEnvUPSOld -> vrpb200_env_step with VRPB200_PHYSICS_UPS_OLD against UPS_old/vrptw_ups_utils.Env driven
    directly (tests/golden/make_golden.py): km travel time through cos(), 1.5 clicks of service per package.  Integer state
    and masks identical; the clock within 2 ulp (cosf of CUDA and of numpy may differ in the last bit)."""
    import torch
    fx, _, _ = load_golden(name)
    pkg = load_pkg()
    args = pkg.task_specific_params.default_args("vrptw10")
    _, Env, reward, _, _ = pkg.host.load_ups_old_components()
    env = Env(args)
    env.input_data = torch.from_numpy(fx["data"].astype(np.float32))
    st = env.reset(W)
    np.testing.assert_array_equal(st.mask.cpu().numpy(), fx["mask"][0])
    for t in range(fx["idx_seq"].shape[0]):
        par = torch.from_numpy(fx["parent_seq"][t][:, None]) if "parent_seq" in fx else None
        st = env.step(torch.from_numpy(fx["idx_seq"][t][:, None]), par)
        np.testing.assert_array_equal(st.mask.cpu().numpy(), fx["mask"][t + 1], err_msg=f"mask, step {t}")
        np.testing.assert_array_equal(st.demand.cpu().numpy(), fx["demand"][t + 1])
        np.testing.assert_array_equal(st.load.cpu().numpy(), fx["load"][t + 1])
        np.testing.assert_array_equal(st.d_sat.cpu().numpy(), fx["d_sat"][t + 1])
        np.testing.assert_allclose(st.time.cpu().numpy(), fx["time"][t + 1], rtol=3e-7, atol=1e-4)
    assert fx["time"].max() > 50.0     # clicks, not unit-square distances


def test_ups_old_reward_is_km_tour_length():
    """reward_func of UPS_old (depot=None): closed tour in km, against the R the reference's own graph produced."""
    import torch
    fx, args, params = load_golden("upsold_vrptw20_greedy")
    pkg = load_pkg()
    rf = pkg.host.load_ups_old_components()[2]
    r = rf([torch.from_numpy(a) for a in fx["actions"]])
    np.testing.assert_allclose(r.cpu().numpy(), fx["R"], rtol=2e-6)
    acts = [a for a in fx["actions"]]                   # with a depot: the UPS_old min_trucks form, against the oracle
    dep = np.ascontiguousarray(fx["data"][:, -1, :-1]).astype(np.float32)
    rd = rf([torch.from_numpy(a) for a in acts], len(acts), float(args["n_nodes"]), torch.from_numpy(dep))
    np.testing.assert_allclose(rd.cpu().numpy(), O.reward_func_min_trucks_ups_old(acts, dep, len(acts), float(args["n_nodes"])), rtol=1e-5)


@pytest.mark.parametrize("case", ["vrptw20_glimpse_tanh", "upsold_vrptw20_greedy"])
def test_rlagent_fuses_glimpses_and_ups_old(case):
    """RLAgent.build_model picks the fused kernel for glimpses / UPS_old physics (one launch) and returns the reference's
    tours, rewards and log-probabilities."""
    import torch
    fx, args, params = load_golden(case)
    pkg = load_pkg()
    H = pkg.host
    comps = H.load_ups_old_components() if args.get("ups_old") else H.load_task_specific_components("vrptw", False)
    store = H.VariableStore(seed=1)
    env = comps[1](args)
    agent = H.RLAgent(args, None, env, None, comps[2], comps[3], comps[4], is_train=False, store=store)
    store.assign(params)
    env.input_data = torch.from_numpy(fx["data"].astype(np.float32))
    out = agent.build_model("greedy")
    assert agent.engine().last_launch()["kernels"] == 1
    idx = torch.stack(out[4])[:, :, 0].cpu().numpy()
    np.testing.assert_array_equal(idx, fx["idxs"])
    np.testing.assert_allclose(out[0].cpu().numpy(), fx["R"], rtol=1e-5)
    np.testing.assert_allclose(torch.stack(out[2]).cpu().numpy(), fx["logprobs"], rtol=2e-4, atol=2e-5)
    np.testing.assert_array_equal(torch.stack(out[3]).cpu().numpy(), fx["actions"])


@pytest.mark.parametrize("task", ["vrp10", "vrptw20"])
@pytest.mark.parametrize("use_tanh", [False, True])
def test_actor_call_matches_oracle(task, use_tanh):
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = O.default_args(task, use_tanh=use_tanh)
    params = O.init_params(args, seed=21, bias_scale=0.3)
    tw = args["task_name"] == "vrptw"
    store = H.VariableStore()
    cls = H.AttentionVRPTWActor if tw else H.AttentionVRPActor
    actor = cls(128, use_tanh=use_tanh, C=10, _name="Decoder/Attention", _scope="Actor/", store=store)
    actor._build_ref(args["embedding_dim"])
    store.assign({k: v for k, v in params.items() if k.startswith("Actor/Decoder/Attention/")})
    rnd = np.random.RandomState(0)
    R, N, E = 37, args["n_nodes"], args["embedding_dim"]
    q = rnd.randn(R, 128).astype(np.float32)
    ref = rnd.randn(R, N, E).astype(np.float32)

    class EnvStub:
        pass
    env = EnvStub()
    env.demand = rnd.randint(0, 10, size=(R, N)).astype(np.float32)
    env.load = rnd.randint(0, 20, size=(R,)).astype(np.float32)
    env.time = rnd.uniform(0, 8, size=(R,)).astype(np.float32)
    e_ref, l_ref = O.attention_actor(params, "Actor/Decoder/Attention", q, ref, env, tw, use_tanh=use_tanh, C=10.0)
    tenv = EnvStub()
    tenv.demand, tenv.load, tenv.time = (torch.from_numpy(x).cuda() for x in (env.demand, env.load, env.time))
    e, logits = actor(torch.from_numpy(q).cuda(), torch.from_numpy(ref).cuda(), tenv)
    np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), l_ref, rtol=1e-4, atol=2e-5)


def test_linear_embedding_and_reward_match_oracle():
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=4, bias_scale=0.3)
    store = H.VariableStore()
    emb = H.LinearEmbedding(30, _scope="Actor/", store=store)
    emb._build(4)
    store.assign({k: v for k, v in params.items() if "Embedding" in k})
    data = O.create_dataset("vrptw20", 33, seed=2).astype(np.float32)
    out = emb(torch.from_numpy(data[:, :, :4]).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), O.linear_embedding(params, data[:, :, :4]), rtol=1e-5, atol=1e-6)
    rnd = np.random.RandomState(1)
    acts = [rnd.uniform(size=(50, 4)).astype(np.float32) for _ in range(12)]
    R = H.reward_func([torch.from_numpy(a).cuda() for a in acts])
    np.testing.assert_allclose(R.cpu().numpy(), O.reward_func(acts, True), rtol=1e-5)


@pytest.mark.parametrize("name", sorted(GOLDEN_INDEX))
def test_rlagent_stepwise_matches_reference_fixture(name):
    """RLAgent.build_model through RNNDecodeStep.step / Env.step on device state: every decode type, incl.
    beam search (parents, back-tracked actions) and glimpses + tanh exploration."""
    fx, args, params = load_golden(name)
    meta = GOLDEN_INDEX[name]
    agent, env, torch, H = _agent(args, params)
    env.input_data = torch.from_numpy(fx["data"].astype(np.float32))
    R, v, logprobs, actions, idxs, pnt, probs = agent.build_model(
        meta["decode_type"], uniforms=fx.get("uniforms"), uniforms_retry=fx.get("uniforms_retry"), want_probs=True,
        backend="stepwise")
    idx = torch.stack(idxs)[:, :, 0].cpu().numpy()
    np.testing.assert_array_equal(idx, fx["idxs"])
    ok = np.ones(idx.shape, bool)
    if meta["decode_type"] == "beam_search":
        # duplicate beams tie exactly in the reference and to ~1e-7 here: the same node is then reached from the
        # twin parent, whose un-reordered LSTM slot differs.  Compare probabilities where the parents agree.
        ref = O.rollout(params, args, fx["data"], "beam_search")
        ours = torch.stack(agent.last_beam_path)[:, :, 0].cpu().numpy()
        ok = ours == ref.beam_parents
        assert ok.mean() > 0.9
    np.testing.assert_allclose(torch.stack(probs).cpu().numpy()[ok], fx["probs"][ok], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(torch.stack(logprobs).cpu().numpy()[ok], fx["logprobs"][ok], rtol=2e-4, atol=2e-5)
    if meta["decode_type"] == "beam_search" and not ok.all():
        # twin beams may swap slots: per instance the multiset of beam costs (and hence the best beam) is unchanged
        W, B = args["beam_width"], fx["data"].shape[0]
        np.testing.assert_allclose(np.sort(R.cpu().numpy().reshape(W, B), 0), np.sort(fx["R"].reshape(W, B), 0), rtol=1e-4)
    else:
        np.testing.assert_array_equal(torch.stack(actions).cpu().numpy(), fx["actions"])
        np.testing.assert_allclose(R.cpu().numpy(), fx["R"], rtol=1e-4)
    np.testing.assert_array_equal(env.demand.cpu().numpy(), fx["final_demand"])
    np.testing.assert_array_equal(env.mask.cpu().numpy(), fx["final_mask"])


@pytest.mark.parametrize("decode_type", ["greedy", "stochastic"])
def test_rlagent_fused_equals_stepwise(decode_type):
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=31, bias_scale=0.2)
    agent, env, torch, H = _agent(args, params)
    B, T = 200, args["decode_len"]
    env.input_data = torch.from_numpy(O.create_dataset("vrptw20", B, seed=8).astype(np.float32))
    rnd = np.random.RandomState(2)
    u1, u2 = rnd.uniform(size=(T, B)).astype(np.float32), rnd.uniform(size=(T, B)).astype(np.float32)
    a = agent.build_model(decode_type, uniforms=u1, uniforms_retry=u2, backend="fused")
    b = agent.build_model(decode_type, uniforms=u1, uniforms_retry=u2, backend="stepwise")
    ia, ib = torch.stack(a[4])[:, :, 0], torch.stack(b[4])[:, :, 0]
    same = (ia == ib).all(0)
    assert same.float().mean().item() >= 0.97
    np.testing.assert_allclose(a[0][same].cpu().numpy(), b[0][same].cpu().numpy(), rtol=1e-5)
    np.testing.assert_array_equal(torch.stack(a[3])[:, same].cpu().numpy(), torch.stack(b[3])[:, same].cpu().numpy())


def test_evaluate_batch_and_checkpoint_by_name(tmp_path):
    pkg = load_pkg()
    H = pkg.host
    args = pkg.task_specific_params.default_args("vrptw10", test_size=64, beam_width=3)
    params = O.init_params(O.default_args("vrptw10"), seed=3)
    np.savez(tmp_path / "model.npz", **params)
    args["load_path"] = str(tmp_path)
    DataGenerator, Env, reward_func, Actor, Critic = H.load_task_specific_components("vrptw", False)
    store = H.VariableStore(seed=99)
    agent = H.RLAgent(args, None, Env(args), DataGenerator(args), reward_func, Actor, Critic, is_train=False, store=store)
    agent.Initialize(None)
    Rg = agent.evaluate_batch("greedy")
    Rb = agent.evaluate_batch("beam_search")
    data = agent.dataGen.get_test_all()
    ref = O.rollout(params, O.default_args("vrptw10"), data, "greedy")
    refb = O.rollout(params, O.default_args("vrptw10", beam_width=3), data, "beam_search")
    np.testing.assert_allclose(Rg, ref.R, rtol=1e-4)
    np.testing.assert_allclose(Rb, O.best_of_beams(refb.R, 64, 3), rtol=1e-4)
    assert Rb.shape == (64,)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["vrp", "vrptw"])
def test_min_trucks_reward_kernel_matches_reference_fixture(task):
    """host.reward_func(..., depot) -> vrpb200_reward_min_trucks against the output of the reference's own reward_func
    (tests/golden/reward_min_trucks.npz); fp32 sums in a different order: 1e-5 relative."""
    import os
    import torch
    from tests.helpers import GOLDEN
    pkg = load_pkg()
    fx = np.load(os.path.join(GOLDEN, "reward_min_trucks.npz"))
    acts = [torch.from_numpy(a).cuda() for a in fx[f"{task}_actions"]]
    dep = torch.from_numpy(fx[f"{task}_depot"]).cuda()
    r = pkg.host.reward_func(acts, int(fx[f"{task}_T"]), float(fx[f"{task}_n_nodes"]), dep)
    np.testing.assert_allclose(r.cpu().numpy(), fx[f"{task}_reward"], rtol=1e-5)


@pytest.mark.gpu
def test_min_trucks_reward_ups_kernel_matches_reference_fixture():
    """load_task_specific_components('vrptw', ups=True) hands out the UPS reward (UPS/vrptw_ups_utils.py:387-419:
    100 x returns to the depot + route length) -> vrpb200_reward_min_trucks_ups against the reference's own output."""
    import os
    import torch
    from tests.helpers import GOLDEN
    pkg = load_pkg()
    fx = np.load(os.path.join(GOLDEN, "reward_min_trucks.npz"))
    rf = pkg.host.load_task_specific_components("vrptw", True)[2]
    acts = [torch.from_numpy(a).cuda() for a in fx["vrptwups_actions"]]
    dep = torch.from_numpy(fx["vrptwups_depot"]).cuda()
    r = rf(acts, int(fx["vrptwups_T"]), float(fx["vrptwups_n_nodes"]), dep)
    np.testing.assert_allclose(r.cpu().numpy(), fx["vrptwups_reward"], rtol=1e-5)
    np.testing.assert_allclose(rf(acts).cpu().numpy(), pkg.host.reward_func(acts).cpu().numpy(), rtol=0)   # no depot: tour length


@pytest.mark.gpu
def test_min_trucks_reward_ups_old_kernel_matches_reference_fixture():
    """load_ups_old_components() hands out the UPS_old reward (UPS_old/vrptw_ups_utils.py:380-430); with a depot: 70/n x steps
    at the depot + 30 x km route length / Euclidean normaliser -> vrpb200_reward_min_trucks_ups_old against the reference's own
    output (cosf against numpy's float32 cos: 1e-5 relative)."""
    import os
    import torch
    from tests.helpers import GOLDEN
    pkg = load_pkg()
    fx = np.load(os.path.join(GOLDEN, "reward_min_trucks.npz"))
    rf = pkg.host.load_ups_old_components()[2]
    acts = [torch.from_numpy(a).cuda() for a in fx["vrptwupsold_actions"]]
    dep = torch.from_numpy(fx["vrptwupsold_depot"]).cuda()
    r = rf(acts, int(fx["vrptwupsold_T"]), float(fx["vrptwupsold_n_nodes"]), dep)
    np.testing.assert_allclose(r.cpu().numpy(), fx["vrptwupsold_reward"], rtol=1e-5)
    km = O.reward_func([a for a in fx["vrptwupsold_actions"]], tw=True, ups_old=True)
    np.testing.assert_allclose(rf(acts).cpu().numpy(), km, rtol=1e-5)           # no depot: the km tour length


@pytest.mark.gpu
def test_evaluate_single_writes_reference_result_files(tmp_path):
    """evaluate_single (attention_agent.py:336-413) + _output_results (:416-471): one route per problem in the result file,
    each route a valid VRPTW route over the instance's nodes starting and ending at the depot."""
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = pkg.task_specific_params.default_args("vrptw10", test_size=5, random_seed=11, beam_width=3)
    args.update(log_dir=str(tmp_path / "logs"), data_dir=str(tmp_path / "data"), batch_size=1, ups=False)
    (tmp_path / "logs").mkdir()
    DataGenerator, Env, reward_func, Actor, Critic = H.load_task_specific_components("vrptw", False)
    dataGen = DataGenerator(args)
    assert (tmp_path / "data" / "vrptw-size-5-len-11-test.txt").exists()
    agent = H.RLAgent(args, None, Env(args), dataGen, reward_func, Actor, Critic, is_train=False, store=H.VariableStore(seed=7))
    agent.Initialize(None)
    R = agent.evaluate_single("greedy")
    assert R.shape == (5,) and np.isfinite(R).all()
    res = tmp_path / "logs" / "results" / "vrptw-size-5-len-11-results-greedy.txt"
    assert res.exists() and (tmp_path / "logs" / "results" / "vrptw-size-5-len-11-test.txt").exists()
    lines = res.read_text().strip().split("\n")
    assert len(lines) == 5
    data = dataGen.get_test_all()
    for b, ln in enumerate(lines):
        v = np.array([float(t) for t in ln.split(" ")]).reshape(-1, 4)
        np.testing.assert_allclose(v[0], data[b, -1, :4], rtol=1e-6)       # starts at the depot
        np.testing.assert_allclose(v[-1], data[b, -1, :4], rtol=1e-6)      # closed with the depot
        nodes = {tuple(np.round(r, 5)) for r in data[b, :, :4]}
        assert all(tuple(np.round(r, 5)) in nodes for r in v)


@pytest.mark.parametrize("decode_type,W", [("greedy", 1), ("beam_search", 3)])
def test_build_model_min_trucks_reward(decode_type, W):
    """args['min_trucks'] switches build_model's R to reward_func(actions, decode_len, n_nodes-1, depot tiled over the
    beams) (attention_agent.py:236-238) - on the fused and on the step-wise route, with back-tracked beam actions."""
    args = O.default_args("vrptw10", beam_width=W, min_trucks=True)
    params = O.init_params(args, seed=41, bias_scale=0.2)
    agent, env, torch, H = _agent(args, params)
    data = O.create_dataset("vrptw10", 24, seed=6)
    env.input_data = torch.from_numpy(data.astype(np.float32))
    ref = O.rollout(params, args, data, decode_type)
    depot = np.tile(data[:, -1, :4].astype(np.float32), [W, 1])
    want = O.reward_func_min_trucks(list(ref.actions), depot, args["decode_len"], args["n_nodes"] - 1, True)
    for backend in ("fused", "stepwise"):
        out = agent.build_model(decode_type, backend=backend)
        idx = torch.stack(out[4])[:, :, 0].cpu().numpy()
        same = (idx == ref.idxs).all(0)
        if decode_type == "beam_search":      # twin beams may swap slots: compare per-instance multisets
            np.testing.assert_allclose(np.sort(out[0].cpu().numpy().reshape(W, -1), 0), np.sort(want.reshape(W, -1), 0), rtol=2e-5)
        else:
            assert same.mean() >= 0.95
            np.testing.assert_allclose(out[0].cpu().numpy()[same], want[same], rtol=2e-5)
    plain = dict(args, min_trucks=False)
    agent2, env2, _, _ = _agent(plain, params)
    env2.input_data = env.input_data
    assert not np.allclose(agent2.build_model(decode_type)[0].cpu().numpy(), want)


@pytest.mark.parametrize("task,W,B", [("vrptw20", 20, 1), ("vrptw20", 20, 100), ("vrptw50", 40, 1), ("vrptw50", 48, 24), ("vrptw20", 17, 1500)])
def test_wide_beams_pick_a_block_size(task, W, B):
    """beam_width in [17, 64] on small and medium batches: the automatic rows-per-CTA choice falls back to the smallest
    block that holds every beam of an instance (it used to find none; evaluate_single with beam_width 20 is B = 1).
    With more beams than finite candidates the reference's top_k picks -inf entries too (masked nodes, probability 0),
    so the comparison is with the oracle, not with a feasibility replay."""
    import torch
    pkg = load_pkg()
    args = O.default_args(task, beam_width=W)
    params = O.init_params(args, seed=5, bias_scale=0.1)
    data = O.create_dataset(task, B, seed=9)
    eng = pkg.Engine(args)
    eng.set_weights(params)
    x = torch.from_numpy(data.astype(np.float32)).cuda()
    with pytest.raises(pkg.VrpB200Error, match="exceeds n_nodes"):     # top_k(k > candidates) is an error in the reference too
        eng.rollout(x, "beam_search", beam_width=args["n_nodes"] + 1)
    out = eng.rollout(x, "beam_search", beam_width=W)
    torch.cuda.synchronize()
    assert eng.last_launch()["rows_per_cta"] >= W
    ref = O.rollout(params, args, data, "beam_search")
    best, ref_best = out.R.cpu().numpy().reshape(W, B).min(0), O.best_of_beams(ref.R, B, W)
    assert np.isclose(best, ref_best, rtol=1e-4).mean() >= 0.9
    same = (out.idxs.cpu().numpy() == ref.idxs).all(0) & (out.beam_parents.cpu().numpy() == ref.beam_parents).all(0)
    assert same.reshape(W, B).all(0).mean() >= 0.8


@pytest.mark.parametrize("task", ["vrptw20", "vrp20"])
@pytest.mark.parametrize("gemm", ["auto", "tc3", "ffma"])
def test_mask_pointer_false_matches_oracle(task, gemm):
    """mask_pointer=False (decode_step.py:231-232 skipped): logits are not masked, a served row keeps moving for all T
    steps instead of being frozen at the depot, every node is evaluated at every step."""
    import torch
    pkg = load_pkg()
    args = O.default_args(task, mask_pointer=False)
    params = O.init_params(args, seed=8, bias_scale=0.3)
    data = O.create_dataset(task, 96, seed=4)
    eng = pkg.Engine(args, gemm=gemm)
    eng.set_weights(params)
    out = eng.rollout(torch.from_numpy(data.astype(np.float32)).cuda(), "greedy", want_logprobs=True, final_state=True, want_steps=True)
    torch.cuda.synchronize()
    ref = O.rollout(params, args, data, "greedy")
    idx = out.idxs.cpu().numpy().astype(np.int64)
    same = (idx == ref.idxs).all(0)
    assert same.mean() >= 0.95
    assert int(out.steps.min()) == args["decode_len"]          # no early exit without the pointer mask
    np.testing.assert_allclose(out.R.cpu().numpy()[same], ref.R[same], rtol=1e-4)
    np.testing.assert_allclose(out.logprobs.cpu().numpy()[:, same], ref.logprobs[:, same], rtol=2e-4, atol=2e-5)
    np.testing.assert_array_equal(out.final_demand.cpu().numpy()[same], ref.final_state.demand[same])


@pytest.mark.parametrize("name,task", [("vrptw10", "vrptw10"), ("vrp10", "vrp10"), ("vrptw20", "vrptw20")])
def test_critic_and_reinforce_losses_match_reference_fixture(name, task):
    """SURVEY 8f N1, forward side: vrpb200_critic against the v the reference's own AttentionVRP(TW)Critic produced
    (tests/golden/critic.npz), and vrpb200_reinforce_loss on the kernel's own stochastic rollout (same uniforms as the
    fixture) against build_train_step's formulas (attention_agent.py:286-294)."""
    import os
    import torch
    from tests.helpers import GOLDEN
    pkg = load_pkg()
    fx = np.load(os.path.join(GOLDEN, "critic.npz"))
    args = O.default_args(task)
    seed, bs = int(fx[f"{name}_seed"]), float(fx[f"{name}_bias_scale"])
    p = O.init_params(args, seed=seed, bias_scale=bs)
    p.update(O.init_critic_params(args, seed=seed + 100, bias_scale=bs))
    eng = pkg.Engine(args, device=0)
    eng.set_weights(p)
    x = torch.from_numpy(fx[f"{name}_data"].astype(np.float32)).cuda()
    v = eng.critic(x)
    np.testing.assert_allclose(v.cpu().numpy(), fx[f"{name}_v"], rtol=1e-4, atol=2e-6)
    out = eng.rollout(x, "stochastic", uniforms=fx[f"{name}_uniforms"], uniforms_retry=fx[f"{name}_uniforms_retry"], want_logprobs=True)
    np.testing.assert_allclose(out.R.cpu().numpy(), fx[f"{name}_R"], rtol=1e-4)
    np.testing.assert_allclose(out.logprobs.cpu().numpy(), fx[f"{name}_logprobs"], rtol=2e-4, atol=2e-5)
    actor, critic = eng.reinforce_loss(out.R, v, out.logprobs)
    want_a, want_c = O.reinforce_losses(fx[f"{name}_R"], fx[f"{name}_v"], fx[f"{name}_logprobs"])
    np.testing.assert_allclose(actor.item(), want_a, rtol=2e-4)
    np.testing.assert_allclose(critic.item(), want_c, rtol=2e-4)
    # without the Critic/* variables the entry point refuses (no silent zeros)
    eng2 = pkg.Engine(args, device=0)
    eng2.set_weights(O.init_params(args, seed=seed, bias_scale=bs))
    with pytest.raises(pkg.VrpB200Error):
        eng2.critic(x)


def test_rlagent_build_train_step_forward_side():
    """RLAgent(is_train=True).build_train_step(): stochastic rollout + critic + losses through the reference's protocol;
    the critic variables appear in the store under the reference's names when the stochastic model is first built."""
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = O.default_args("vrptw10")
    comps = H.load_task_specific_components("vrptw", False)
    store = H.VariableStore(seed=5)
    env = comps[1](args)
    agent = H.RLAgent(args, None, env, None, comps[2], comps[3], comps[4], is_train=True, store=store)
    data = O.create_dataset("vrptw10", 32, seed=9)
    env.input_data = torch.from_numpy(data.astype(np.float32))
    assert not any(k.startswith("Critic/") for k in store.vars)
    T = args["decode_len"]
    rnd = np.random.RandomState(3)
    u1, u2 = rnd.uniform(size=(T, 32)).astype(np.float32), rnd.uniform(size=(T, 32)).astype(np.float32)
    step = agent.build_train_step(uniforms=u1, uniforms_retry=u2, apply=False)      # compute only: the variables stay as they are
    assert len(step) == 12 and step[0] == 0 and step[1] == 0
    names = {k for k in store.vars if k.startswith("Critic/")}
    assert "Critic/Process/P2/proj_end_tw/kernel" in names and "Critic/Linear/L2/bias" in names and len(names) == 3 * 13 + 4
    params = {k: v for k, v in store.vars.items() if k != "decoder_input"}
    ref = O.rollout(params, args, data, "stochastic", uniforms=u1, uniforms_retry=u2)
    v_ref = O.critic_value(params, args, data)
    R, v = step[6].cpu().numpy(), step[7].cpu().numpy()
    np.testing.assert_allclose(v, v_ref, rtol=1e-4, atol=2e-6)
    same = (torch.stack(step[11])[:, :, 0].cpu().numpy() == ref.idxs).all(0)
    assert same.mean() >= 0.9
    np.testing.assert_allclose(R[same], ref.R[same], rtol=1e-4)
    if same.all():
        a_ref, c_ref = O.reinforce_losses(ref.R, v_ref, ref.logprobs)
        np.testing.assert_allclose(step[2].item(), a_ref, rtol=5e-4)
        np.testing.assert_allclose(step[3].item(), c_ref, rtol=5e-4)
    # greedy decoding keeps v = 0 (attention_agent.py:244)
    assert float(agent.build_model("greedy")[1]) == 0.0
