"""OFFLINE golden-vector generator against REAL TensorFlow 1.x.  **Not executed in this repository**: TensorFlow is
not installable in the build image (no wheel, no network; the reference needs TF <= 1.15, i.e. Python <= 3.7).

Run it on a machine with `tensorflow==1.15` and a checkout of jpoulletXaccount/MIT_rl-vrptw next to this repo:

    python tools/tf1_golden.py --reference /path/to/MIT_rl-vrptw --out tests/golden_tf

For every case of tests/golden/index.json it builds the reference graph (RLAgent.build_model) under real TF,
assigns the SAME weights (oracle.restatement.init_params, by variable name) and the SAME instances, and dumps
idxs / logits / probs / masks / R with the layout of tests/golden/*.npz, so that
`pytest tests/test_oracle_golden.py --golden-dir tests/golden_tf` closes the loop that the numpy shim leaves open
(Eigen exp/tanh, MatMul summation order, tf.nn.top_k tie order).
"""
import argparse
import json
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import tensorflow as tf                                   # real TF 1.x
    assert tf.__version__.startswith("1."), "the reference needs TensorFlow 1.x"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, a.reference)
    from oracle import restatement as O
    from model.attention_agent import RLAgent
    os.makedirs(a.out, exist_ok=True)
    index = json.load(open(os.path.join(root, "tests", "golden", "index.json")))
    for name, meta in sorted(index.items()):
        fx = dict(np.load(os.path.join(root, "tests", "golden", name + ".npz")))
        args = O.default_args(meta["task"], **meta["over"])
        params = O.init_params(args, seed=int(fx["param_seed"]), bias_scale=float(fx["bias_scale"]))
        tf.reset_default_graph()
        if args["task_name"] == "vrp":
            from VRP.vrp_utils import Env, reward_func
            from VRP.vrp_attention import AttentionVRPActor as A, AttentionVRPCritic as C
        else:
            from VRPTW.vrptw_utils import Env, reward_func
            from VRPTW.vrptw_attention import AttentionVRPTWActor as A, AttentionVRPTWCritic as C
        run_args = dict(args, log_dir=a.out)

        class Prt:
            def print_out(self, *x, **k):
                pass
        env = Env(run_args)
        agent = RLAgent(run_args, Prt(), env, None, reward_func, A, C, is_train=False)
        summary = {"greedy": agent.val_summary_greedy, "beam_search": agent.val_summary_beam}.get(meta["decode_type"])
        if summary is None:
            print("skip", name, "(stochastic: TF's RNG stream cannot be fed)")
            continue
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            for v in tf.trainable_variables():
                key = v.name[:-2]
                if key in params:
                    v.load(params[key].reshape(v.shape.as_list()), sess)
            R, v_, logprobs, actions, idxs, batch, probs = sess.run(
                summary, feed_dict={env.input_data: fx["data"], agent.decodeStep.dropout: 0.0})
        np.savez_compressed(os.path.join(a.out, name + ".npz"), R=R, logprobs=np.stack(logprobs), actions=np.stack(actions),
                            idxs=np.stack(idxs)[:, :, 0], probs=np.stack(probs), data=fx["data"],
                            param_seed=fx["param_seed"], bias_scale=fx["bias_scale"])
        print(name, "R mean", float(np.mean(R)))


if __name__ == "__main__":
    main()
