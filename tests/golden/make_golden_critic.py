"""Golden fixtures for the critic value `v` of RLAgent.build_model('stochastic') (model/attention_agent.py:243-270), produced
by the reference's OWN code (AttentionVRPTWCritic / AttentionVRPCritic, unmodified, imported from /root/reference) under the
TF-1 numpy shim.  Run in the build container:  python tests/golden/make_golden_critic.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (sets up the shim + reference import path)
from make_golden import O, tf  # noqa: E402


def run(task, B, seed, bias_scale, ups=False):
    args = O.default_args(task, ups=ups)
    params = O.init_params(args, seed=seed, bias_scale=bias_scale)
    params.update(O.init_critic_params(args, seed=seed + 100, bias_scale=bias_scale))
    data = O.create_dataset(task, B, seed=2000 + seed)
    tf.reset_shim()
    tf.VARIABLE_STORE.update(params)
    Env, reward_func, Actor, Critic = G.components(args["task_name"], ups)
    a = dict(args)
    a["log_dir"] = "/tmp"
    env = Env(a)
    agent = G.RLAgent(a, G._Prt(), env, None, reward_func, Actor, Critic, is_train=False)
    summary = agent.build_model("stochastic")
    T = a["decode_len"]
    rnd = np.random.RandomState(seed)
    uni = rnd.uniform(size=(T, B)).astype(np.float32)
    uni2 = rnd.uniform(size=(T, B)).astype(np.float32)
    feed = []
    for t in range(T):
        feed += [uni[t].reshape(-1, 1), uni2[t].reshape(-1, 1)]
    tf.UNIFORM_FEED[:] = feed
    R, v, logprobs = tf.Session().run((summary[0], summary[1], summary[2]),
                                       feed_dict={env.input_data: data, agent.decodeStep.dropout: 0.0})
    requested = dict(tf.REQUESTED_VARIABLES)
    unused = sorted(k for k in params if k.startswith("Critic") and k not in requested)
    assert not unused, f"oracle defines critic variables the reference never reads: {unused}"
    return dict(data=data, v=np.asarray(v, np.float32), R=np.asarray(R, np.float32), logprobs=np.stack(logprobs).astype(np.float32),
                uniforms=uni, uniforms_retry=uni2, seed=np.int64(seed), bias_scale=np.float64(bias_scale))


if __name__ == "__main__":
    fx = {}
    for name, task, B, seed, bs, ups in (("vrptw10", "vrptw10", 16, 51, 0.3, False), ("vrp10", "vrp10", 16, 52, 0.3, False),
                                         ("vrptw20", "vrptw20", 8, 53, 0.0, False)):
        for k, v in run(task, B, seed, bs, ups).items():
            fx[f"{name}_{k}"] = v
        print(name, "v", fx[f"{name}_v"][:4], "R", fx[f"{name}_R"][:4])
    np.savez_compressed(os.path.join(HERE, "critic.npz"), **fx)
