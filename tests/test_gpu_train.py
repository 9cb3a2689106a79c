"""The backward / update side of build_train_step on the GPU (vrpb200_actor_grad, vrpb200_clip_adam) against torch autograd
through the as-written forward graph (oracle/train_torch.py).  The algebra of the kernel is also checked on the CPU from the
same source (tests/test_actor_grad_cpu.py)."""
import numpy as np
import pytest

from oracle import restatement as O
from oracle import train_torch as TT
from tests.helpers import load_pkg
from tests.test_actor_grad_cpu import check, dropout_scale, make_case
from tests.test_critic_grad_cpu import check as check_c
from tests.test_critic_grad_cpu import make_case as make_case_c

pytestmark = pytest.mark.gpu


def _engine(args, params):
    import torch
    pkg = load_pkg()
    eng = pkg.Engine(args, device=0)
    eng.set_weights(params)
    return eng, torch, pkg


@pytest.mark.parametrize("task,B,over", [("vrptw10", 6, {}), ("vrp10", 6, {}), ("vrptw20", 6, {"use_tanh": True}),
                                         ("vrptw10", 6, {"mask_pointer": False}), ("vrptw20", 128, {}), ("vrp20", 100, {})])
def test_actor_gradient_matches_autograd(task, B, over):
    """compute_gradients(actor_loss) (attention_agent.py:289-297) along a sampled tour: every Actor variable's gradient against
    autograd through the as-written graph, 1e-4 relative (+ 2e-5 of the variable's largest gradient entry)."""
    args, params, data, idxs, weight = make_case(task, B, 31, **over)
    eng, torch, pkg = _engine(args, params)
    names = [k for k in params if k.startswith("Actor/")]
    got = eng.actor_grad(torch.from_numpy(data).cuda(), torch.from_numpy(idxs).cuda(), torch.from_numpy(weight).cuda(),
                         {k: params[k].shape for k in names})
    torch.cuda.synchronize()
    loss, want, st, _ = TT.actor_grads(params, args, data, idxs, weight)
    check({k: v.cpu().numpy() for k, v in got.items()}, want, names)
    # deterministic: a second call gives the same bits
    again = eng.actor_grad(torch.from_numpy(data).cuda(), torch.from_numpy(idxs).cuda(), torch.from_numpy(weight).cuda(),
                           {k: params[k].shape for k in names})
    for k in names:
        assert torch.equal(got[k], again[k]), k


def test_actor_gradient_follows_the_rollout_kernel():
    """The tour and the log-probabilities of vrpb200_rollout(STOCHASTIC) fed back: the teacher-forced graph reproduces the
    kernel's log-probabilities, and the gradient along the kernel's own tour matches autograd; names and sizes that are not
    the graph's are refused."""
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=77, bias_scale=0.1)
    data = O.create_dataset("vrptw20", 64, seed=78).astype(np.float32)
    eng, torch, pkg = _engine(args, params)
    x = torch.from_numpy(data).cuda()
    out = eng.rollout(x, "stochastic", seed=5, want_logprobs=True)
    torch.cuda.synchronize()
    idxs = out.idxs.cpu().numpy().astype(np.int32)
    R = out.R.cpu().numpy()
    weight = ((R - R.mean()) / len(R)).astype(np.float32)
    loss, want, st, lps = TT.actor_grads(params, args, data, idxs, weight)
    np.testing.assert_allclose(out.logprobs.cpu().numpy(), lps, rtol=2e-4, atol=2e-5)
    names = [k for k in params if k.startswith("Actor/")]
    got = eng.actor_grad(x, out.idxs, torch.from_numpy(weight).cuda(), {k: params[k].shape for k in names})
    check({k: v.cpu().numpy() for k, v in got.items()}, want, names)
    with pytest.raises(pkg.VrpB200Error):
        eng.actor_grad(x, out.idxs, torch.from_numpy(weight).cuda(), {"Actor/Decoder/Attention/nope": (3,)})
    with pytest.raises(pkg.VrpB200Error):
        eng.actor_grad(x, out.idxs, torch.from_numpy(weight).cuda(), {"Actor/Decoder/Attention/v": (1, 7)})


@pytest.mark.parametrize("task,B", [("vrptw10", 7), ("vrp10", 7), ("vrptw20", 128), ("vrptw50", 64)])
def test_critic_gradient_matches_autograd(task, B):
    """compute_gradients(critic_loss) (attention_agent.py:290, 298-299): every Critic variable's gradient of mean((R - v)^2) against
    autograd through the as-written critic, and the value it was taken at against the pinned oracle."""
    args, params, data, Rw = make_case_c(task, B, 41)
    eng, torch, pkg = _engine(args, params)
    names = [k for k in params if k.startswith("Critic/")]
    got, v = eng.critic_grad(torch.from_numpy(data).cuda(), torch.from_numpy(Rw).cuda(), {k: params[k].shape for k in names})
    torch.cuda.synchronize()
    loss, want, v_ref = TT.critic_grads(params, args, data, Rw)
    np.testing.assert_allclose(v.cpu().numpy(), v_ref, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(v.cpu().numpy(), eng.critic(torch.from_numpy(data).cuda()).cpu().numpy(), rtol=1e-4, atol=2e-6)
    check_c({k: g.cpu().numpy() for k, g in got.items()}, want, names, max(np.abs(w).max() for w in want.values()))
    again, _ = eng.critic_grad(torch.from_numpy(data).cuda(), torch.from_numpy(Rw).cuda(), {k: params[k].shape for k in names})
    for k in names:
        assert torch.equal(got[k], again[k]), k
    eng2, _, _ = _engine(args, {k: a for k, a in params.items() if k.startswith("Actor/")})
    with pytest.raises(pkg.VrpB200Error):
        eng2.critic_grad(torch.from_numpy(data).cuda(), torch.from_numpy(Rw).cuda(), {})


def test_clip_by_norm_and_adam_match_the_tf1_formulas():
    """attention_agent.py:303-311: per-variable clip_by_norm, then Adam with TF-1's bias-corrected step size; three steps on a
    clipped and an un-clipped variable against the numpy statement of the same formulas."""
    args = O.default_args("vrp10")
    eng, torch, pkg = _engine(args, O.init_params(args, seed=1))
    rnd = np.random.RandomState(3)
    var = [rnd.normal(size=s).astype(np.float32) for s in ((128, 128), (1, 30))]
    m = [np.zeros_like(a) for a in var]
    v = [np.zeros_like(a) for a in var]
    dvar, dm, dv = ([torch.from_numpy(a.copy()).cuda() for a in lst] for lst in (var, m, v))
    for step in (1, 2, 3):
        grads = [rnd.normal(size=var[0].shape).astype(np.float32) * 0.5, rnd.normal(size=var[1].shape).astype(np.float32) * 0.01]
        assert np.linalg.norm(grads[0]) > 2.0 > np.linalg.norm(grads[1])
        eng.clip_adam(dvar, [torch.from_numpy(g).cuda() for g in grads], dm, dv, max_norm=2.0, lr=1e-3, step=step)
        for i in range(2):
            var[i], m[i], v[i] = TT.clip_adam(var[i], grads[i], m[i], v[i], step, 1e-3, 2.0)
        torch.cuda.synchronize()
        for i in range(2):
            np.testing.assert_allclose(dm[i].cpu().numpy(), m[i], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(dv[i].cpu().numpy(), v[i], rtol=1e-5, atol=1e-12)
            np.testing.assert_allclose(dvar[i].cpu().numpy(), var[i], rtol=1e-5, atol=2e-7)


def test_rlagent_train_step_updates_actor_and_critic():
    """RLAgent.build_train_step() end to end (attention_agent.py:272-322): the gradients of the agent's own Actor and Critic
    variables against autograd, then two Adam steps of each optimizer against the numpy statement of clip_by_norm + Adam applied
    to those gradients; the store is the updated variable collection afterwards and the next rollout uses it."""
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = O.default_args("vrptw10")
    args.update(actor_net_lr=1e-3, critic_net_lr=2e-3, max_grad_norm=2.0)
    comps = H.load_task_specific_components("vrptw", False)
    store = H.VariableStore(seed=5)
    env = comps[1](args)
    agent = H.RLAgent(args, None, env, None, comps[2], comps[3], comps[4], is_train=True, store=store)
    data = O.create_dataset("vrptw10", 32, seed=9).astype(np.float32)
    env.input_data = torch.from_numpy(data)
    T = args["decode_len"]
    rnd = np.random.RandomState(3)
    agent.build_train_step(uniforms=rnd.uniform(size=(T, 32)).astype(np.float32), uniforms_retry=rnd.uniform(size=(T, 32)).astype(np.float32),
                           apply=False)                      # creates the Critic variables, changes nothing
    m, v = {}, {}
    for step_no in (1, 2):
        u1, u2 = rnd.uniform(size=(T, 32)).astype(np.float32), rnd.uniform(size=(T, 32)).astype(np.float32)
        before = {k: a.copy() for k, a in store.vars.items()}
        step = agent.build_train_step(uniforms=u1, uniforms_retry=u2)
        assert step[0] == step_no and step[1] == step_no
        R, vv = step[6].cpu().numpy(), step[7].cpu().numpy()
        idxs = torch.stack(step[11])[:, :, 0].cpu().numpy().astype(np.int32)
        params = {k: a for k, a in before.items() if k != "decoder_input"}
        _, want, _, _ = TT.actor_grads(params, args, data, idxs, ((R - vv) / 32).astype(np.float32))
        _, want_c, v_ref = TT.critic_grads(params, args, data, R)
        np.testing.assert_allclose(vv, v_ref, rtol=1e-4, atol=2e-6)
        got = {name: g.cpu().numpy() for g, name in step[4]}
        got_c = {name: g.cpu().numpy() for g, name in step[5]}
        assert set(got) == {k for k in params if k.startswith("Actor/")} and set(got_c) == {k for k in params if k.startswith("Critic/")}
        check(got, want, list(got))
        check_c(got_c, want_c, list(got_c), max(np.abs(w).max() for w in want_c.values()))
        for grads, lr in ((got, 1e-3), (got_c, 2e-3)):
            for k, g in grads.items():
                if k not in m:
                    m[k], v[k] = np.zeros_like(g), np.zeros_like(g)
                new, m[k], v[k] = TT.clip_adam(before[k], g, m[k], v[k], step_no, lr, 2.0)
                np.testing.assert_allclose(store.vars[k], new, rtol=1e-5, atol=2e-7, err_msg=k)
                if np.abs(g).max() > 0:
                    assert np.abs(store.vars[k] - before[k]).max() > 1e-4, k      # Adam's first steps move every variable by ~lr
        np.testing.assert_array_equal(store.vars["decoder_input"], before["decoder_input"])


def test_run_train_step_learns(tmp_path):
    """run_train_step (attention_agent.py:530-556) in the loop of main.py:107-131 on VRP-10: fresh batches from the data generator,
    in-kernel sampling; the critic's regression loss falls by an order of magnitude within 30 steps (the same schedule on the
    torch oracle: 58.8 -> 2.6), every variable stays finite, and the saved collection restores."""
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = pkg.task_specific_params.default_args("vrp10", test_size=16, random_seed=3)
    args.update(batch_size=64, actor_net_lr=1e-4, critic_net_lr=1e-3, max_grad_norm=2.0, data_dir=None, ups=False)
    DataGenerator, Env, reward_func, Actor, Critic = H.load_task_specific_components("vrp", False)
    store = H.VariableStore(seed=11)
    agent = H.RLAgent(args, None, Env(args), DataGenerator(args), reward_func, Actor, Critic, is_train=True, store=store)
    agent.Initialize(None)
    losses, rewards = [], []
    for step in range(30):
        summary = agent.run_train_step()
        _, _, actor_loss, critic_loss, agv, cgv, R, v, logprobs, probs, actions, idxs = summary
        losses.append(float(critic_loss))
        rewards.append(float(R.mean()))
        assert np.isfinite(float(actor_loss)) and np.isfinite(losses[-1])
    assert summary[0] == 30 and summary[1] == 30
    assert losses[-1] < 0.25 * losses[0], losses
    assert all(np.isfinite(a).all() for a in store.vars.values())
    agent.save_model(str(tmp_path / "model.npz"))
    store2 = H.VariableStore(seed=99)
    agent2 = H.RLAgent(dict(args, load_path=str(tmp_path)), None, Env(args), DataGenerator(args), reward_func, Actor, Critic, is_train=True,
                       store=store2)
    agent2.Initialize(None)
    for k in ("Actor/Decoder/Attention/v", "Actor/Embedding/conv1d/kernel"):
        np.testing.assert_array_equal(store2.vars[k], store.vars[k])


@pytest.mark.parametrize("task,B", [("vrptw10", 6), ("vrp20", 64)])
def test_actor_gradient_with_lstm_input_dropout(task, B):
    """vrpb200_actor_grad with d_input_scale (DropoutWrapper(input_keep_prob), shared/decode_step.py:177-179, given the mask)
    against autograd through the graph with the same scaled LSTM inputs."""
    import torch
    pkg = load_pkg()
    args, params, data, idxs, weight = make_case(task, B, 33)
    scale = dropout_scale(args, B, 5)
    eng = pkg.Engine(args, device=0)
    eng.set_weights(params)
    names = [k for k in params if k.startswith("Actor/")]
    got = eng.actor_grad(torch.from_numpy(data).cuda(), torch.from_numpy(idxs).cuda(), torch.from_numpy(weight).cuda(),
                         {k: params[k].shape for k in names}, input_scale=torch.from_numpy(scale).cuda())
    _, want, _, _ = TT.actor_grads(params, args, data, idxs, weight, scale)
    check({k: v.cpu().numpy() for k, v in got.items()}, want, names)
    with pytest.raises(pkg.VrpB200Error):
        eng.actor_grad(torch.from_numpy(data).cuda(), torch.from_numpy(idxs).cuda(), torch.from_numpy(weight).cuda(),
                       {k: params[k].shape for k in names}, input_scale=torch.from_numpy(scale[:-1]).cuda())


def test_train_step_with_dropout_decodes_and_differentiates_the_same_graph():
    """build_train_step(input_scale=...): the tour is decoded step by step with the scaled LSTM inputs (its log-probabilities are
    the teacher-forced oracle's under the same scale) and the gradient is taken through that scale; run_train_step draws the
    mask from args['dropout'] (attention_agent.py:543)."""
    import torch
    pkg = load_pkg()
    H = pkg.host
    args = O.default_args("vrptw10")
    comps = H.load_task_specific_components("vrptw", False)
    store = H.VariableStore(seed=5)
    env = comps[1](args)
    agent = H.RLAgent(args, None, env, None, comps[2], comps[3], comps[4], is_train=True, store=store)
    B, T = 32, args["decode_len"]
    data = O.create_dataset("vrptw10", B, seed=9).astype(np.float32)
    env.input_data = torch.from_numpy(data)
    rnd = np.random.RandomState(3)
    u1, u2 = rnd.uniform(size=(T, B)).astype(np.float32), rnd.uniform(size=(T, B)).astype(np.float32)
    scale = dropout_scale(args, B, 7, rate=0.25)
    step = agent.build_train_step(uniforms=u1, uniforms_retry=u2, apply=False, input_scale=scale)
    plain = agent.build_train_step(uniforms=u1, uniforms_retry=u2, apply=False)
    R, vv = step[6].cpu().numpy(), step[7].cpu().numpy()
    idxs = torch.stack(step[11])[:, :, 0].cpu().numpy().astype(np.int32)
    params = {k: a for k, a in store.vars.items() if k != "decoder_input"}
    _, want, _, lps = TT.actor_grads(params, args, data, idxs, ((R - vv) / B).astype(np.float32), scale)
    np.testing.assert_allclose(torch.stack(step[8]).cpu().numpy(), lps, rtol=2e-4, atol=2e-5)
    got = {name: g.cpu().numpy() for g, name in step[4]}
    check(got, want, list(got))
    assert not np.array_equal(idxs, torch.stack(plain[11])[:, :, 0].cpu().numpy())        # the mask changes the sampled tours
    ref = O.rollout(params, args, data, "stochastic", uniforms=u1, uniforms_retry=u2)      # and without it the fused kernel's tours
    assert (torch.stack(plain[11])[:, :, 0].cpu().numpy() == ref.idxs).all(0).mean() >= 0.9      # are the reference graph's
