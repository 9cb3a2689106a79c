"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE'S OWN PYTHON FILES.

Run once, in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports, unmodified, from /root/reference:
    VRPTW/vrptw_utils.py  VRP/vrp_utils.py            (Env, reward_func)
    VRPTW/vrptw_attention.py  VRP/vrp_attention.py    (Actor / Critic attention)
    shared/decode_step.py  shared/embeddings.py       (RNNDecodeStep, LinearEmbedding)
    model/attention_agent.py                          (RLAgent.build_model)
with ``oracle/tf1_shim`` standing in for the TensorFlow-1 package (TensorFlow itself is not
installable here - see the shim's docstring for what that does and does not pin).  Modules that
are NOT on the rollout path and cannot be imported here (configs -> scipy.misc, the GNN
embedding -> dpu_utils) are replaced by empty stubs in ``sys.modules``.

Weights come from ``oracle.restatement.init_params`` (same seed => same weights in the tests);
the shim fails if the reference asks for a variable name/shape the oracle does not define.
Fixtures hold inputs + the reference's outputs; tests compare oracle (and CUDA) against them.
"""
import os
import sys
import tempfile
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)

from oracle import restatement as O  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Dummy:
    def __init__(self, *a, **k):
        raise RuntimeError("off-path component stubbed out")


_stub("configs", ParseParams=lambda: None)
_stub("shared.graph_embedding")
_stub("shared.graph_embedding.useful_files")
_stub("shared.graph_embedding.useful_files.gnn_film_model", GNN_FiLM_Model=_Dummy)
_stub("shared.graph_embedding.useful_files.number_vehicle_task", Nb_Vehicles_Task=_Dummy)
_stub("shared.graph_embedding.full_graph_learning", FullGraphEmbedding=_Dummy)

from model.attention_agent import RLAgent  # noqa: E402


def components(task_name, ups=False, ups_old=False):
    """main.py:17-53 (load_task_specific_components): the UPS twins when args['ups'] is set; ups_old: the UPS_old package
    (nothing in main.py imports it; its Env / reward_func / actor are loaded the same way)."""
    if ups_old:
        from UPS_old.vrptw_ups_utils import Env, reward_func
        from UPS_old.vrptw_ups_attention import AttentionVRPTWActor as A, AttentionVRPTWCritic as C
    elif task_name == "vrp" and ups:
        from UPS.vrp_ups_utils import Env, reward_func
        from UPS.vrp_ups_attention import AttentionVRP_UPS_Actor as A, AttentionVRP_UPS_Critic as C
    elif task_name == "vrp":
        from VRP.vrp_utils import Env, reward_func
        from VRP.vrp_attention import AttentionVRPActor as A, AttentionVRPCritic as C
    elif ups:
        from UPS.vrptw_ups_utils import Env, reward_func
        from UPS.vrptw_ups_attention import AttentionVRPTW_UPS_Actor as A, AttentionVRPTW_UPS_Critic as C
    else:
        from VRPTW.vrptw_utils import Env, reward_func
        from VRPTW.vrptw_attention import AttentionVRPTWActor as A, AttentionVRPTWCritic as C
    return Env, reward_func, A, C


class _Prt:
    def print_out(self, *a, **k):
        pass


def params_crc(p):
    crc = 0
    for k in sorted(p):
        crc = zlib.crc32(np.ascontiguousarray(p[k]).tobytes(), crc)
    return crc


def run_reference(args, params, data, decode_type, uniforms=None, uniforms_retry=None):
    """Build the reference graph for ``decode_type`` and evaluate it on ``data``."""
    tf.reset_shim()
    tf.VARIABLE_STORE.update(params)
    Env, reward_func, Actor, Critic = components(args["task_name"], bool(args.get("ups")), bool(args.get("ups_old")))
    a = dict(args)
    a["log_dir"] = tempfile.mkdtemp()
    env = Env(a)
    agent = RLAgent(a, _Prt(), env, None, reward_func, Actor, Critic, is_train=False)
    trace = {"logit": [], "mask": []}
    orig_step = agent.decodeStep.step

    def spy(decoder_inp, context, Env_, decoder_state=None, *x, **k):
        trace["mask"].append(Env_.mask)
        out = orig_step(decoder_inp, context, Env_, decoder_state, *x, **k)
        trace["logit"].append(out[0])
        return out

    if decode_type == "greedy":
        summary = agent.val_summary_greedy
    elif decode_type == "beam_search":
        summary = agent.val_summary_beam
    else:
        summary = None
    # rebuild under the spy so that logits / masks of every step are fetchable
    agent.decodeStep.step = spy
    summary = agent.build_model(decode_type)
    if decode_type == "stochastic":
        T = a["decode_len"]
        # creation order of the tf.random_uniform nodes: first draw, then the one inside tf.cond
        feed = []
        for t in range(T):
            feed += [uniforms[t].reshape(-1, 1), uniforms_retry[t].reshape(-1, 1)]
        tf.UNIFORM_FEED[:] = feed
    sess = tf.Session()
    # the critic value `v` (summary[1]) is off the decode path and is not fetched
    summary = (summary[0], summary[2], summary[3], summary[4], summary[5], summary[6])
    fetch = (summary, trace["logit"], trace["mask"],
             (env.load, getattr(env, "time", None) if a["task_name"] == "vrptw" else env.load, env.demand, env.mask))
    (R, logprobs, actions, idxs, input_pnt, probs), logits, masks, final = sess.run(
        fetch, feed_dict={env.input_data: data, agent.decodeStep.dropout: 0.0})
    requested = dict(tf.REQUESTED_VARIABLES)
    return dict(R=R, logprobs=np.stack(logprobs), actions=np.stack(actions),
                idxs=np.stack(idxs)[:, :, 0], probs=np.stack(probs), logits=np.stack(logits),
                masks=np.stack(masks), final_load=final[0], final_time=final[1],
                final_demand=final[2], final_mask=final[3]), requested


def run_reference_env(args, data, idx_seq, parent_seq, beam_width):
    """Drive the reference Env directly with a prescribed action sequence."""
    tf.reset_shim()
    Env, _, _, _ = components(args["task_name"], False, bool(args.get("ups_old")))
    env = Env(dict(args))
    sess = tf.Session()
    out = {"load": [], "time": [], "demand": [], "mask": [], "d_sat": []}
    st = env.reset(beam_width)
    fetch = [st]
    for t in range(len(idx_seq)):
        st = env.step(tf.constant(idx_seq[t][:, None].astype(np.int64)),
                      None if parent_seq is None else tf.constant(parent_seq[t][:, None].astype(np.int64)))
        fetch.append(st)
    states = sess.run(fetch, feed_dict={env.input_data: data})
    tw = args["task_name"] == "vrptw"
    for s in states:
        out["load"].append(s.load)
        out["demand"].append(s.demand)
        out["mask"].append(s.mask)
        out["d_sat"].append(s.d_sat if s.d_sat.ndim == 1 else np.zeros_like(s.load))
        out["time"].append(s.time if tw else np.zeros_like(s.load))
    return {k: np.stack(v) for k, v in out.items()}


CASES = [
    # name, task, decode_type, B, overrides, param seed, bias_scale
    ("vrptw10_greedy", "vrptw10", "greedy", 16, {}, 11, 0.0),
    ("vrptw10_greedy_bias", "vrptw10", "greedy", 16, {}, 12, 0.3),
    ("vrp10_greedy_bias", "vrp10", "greedy", 16, {}, 13, 0.3),
    ("vrptw10_beam3", "vrptw10", "beam_search", 8, {"beam_width": 3}, 14, 0.3),
    ("vrp10_beam4", "vrp10", "beam_search", 8, {"beam_width": 4}, 15, 0.3),
    ("vrptw10_stochastic", "vrptw10", "stochastic", 16, {}, 16, 0.3),
    ("vrptw20_glimpse_tanh", "vrptw20", "greedy", 8, {"n_glimpses": 1, "use_tanh": True}, 17, 0.3),
    ("vrptw10_small_dims", "vrptw10", "greedy", 8, {"hidden_dim": 128, "embedding_dim": 8}, 18, 0.3),
    # round 2: the sizes and classes of the headline configs, run through the reference's own code as well
    ("vrptw50_greedy", "vrptw50", "greedy", 4, {}, 21, 0.0),
    ("vrptw100_greedy", "vrptw100", "greedy", 4, {}, 22, 0.0),                       # extrapolated task
    ("ups_vrptw100_greedy", "vrptw100", "greedy", 4, {"ups": True}, 23, 0.0),        # UPS/vrptw_ups_utils.Env + UPS actor, GMM instances
    ("ups_vrptw20_stochastic", "vrptw20", "stochastic", 8, {"ups": True}, 26, 0.2),
    ("ups_vrp20_greedy", "vrp20", "greedy", 8, {"ups": True}, 27, 0.2),              # UPS/vrp_ups_utils.Env + UPS actor
    ("vrptw20_tanh", "vrptw20", "greedy", 8, {"use_tanh": True}, 24, 0.3),           # C tanh(u) without glimpses
    ("vrptw10_nomaskptr", "vrptw10", "greedy", 8, {"mask_pointer": False}, 25, 0.3),
    # glimpses beyond the one-glimpse case above, and the UPS_old package (km physics, service time)
    ("vrptw10_glimpse2", "vrptw10", "greedy", 8, {"n_glimpses": 2}, 28, 0.3),
    ("vrptw10_glimpse_nomask", "vrptw10", "stochastic", 8, {"n_glimpses": 1, "mask_glimpses": False}, 29, 0.3),
    ("vrp10_glimpse", "vrp10", "greedy", 8, {"n_glimpses": 1}, 30, 0.3),
    ("upsold_vrptw20_greedy", "vrptw20", "greedy", 8, {"ups_old": True, "capacity": 30}, 31, 0.2),
    ("upsold_vrptw10_stochastic", "vrptw10", "stochastic", 8, {"ups_old": True}, 32, 0.2),
]


def main():
    import json
    only = set(sys.argv[1:])
    index = {}
    if only and os.path.exists(os.path.join(HERE, "index.json")):
        with open(os.path.join(HERE, "index.json")) as f:
            index = json.load(f)
    for name, task, dtype_, B, over, seed, bscale in CASES:
        if only and name not in only:
            continue
        args = O.default_args(task, **over)
        params = O.init_params(args, seed=seed, bias_scale=bscale)
        if over.get("ups_old"):
            data = O.create_dataset_ups_old(args["n_cust"], B, seed=1000 + seed)
        elif over.get("ups"):   # the UPS generators (GMM samples), restated in the product package's data module
            import importlib
            data = importlib.import_module("mit_rl-vrptw_b200.data").create_dataset(args["task_name"], B, args["n_cust"], 1000 + seed, ups=True)
        else:
            data = O.create_dataset(task, B, seed=1000 + seed)
        T = args["decode_len"]
        W = args["beam_width"] if dtype_ == "beam_search" else 1
        rnd = np.random.RandomState(seed)
        uni = rnd.uniform(size=(T, B * W)).astype(np.float32) if dtype_ == "stochastic" else None
        uni2 = rnd.uniform(size=(T, B * W)).astype(np.float32) if dtype_ == "stochastic" else None
        out, requested = run_reference(args, params, data, dtype_, uni, uni2)
        # every variable the oracle defines must have been asked for by the reference (names pinned both ways)
        unused = sorted(set(params) - set(requested))
        assert not unused, f"{name}: oracle defines variables the reference never reads: {unused}"
        fx = dict(out)
        fx["data"] = data
        fx["param_seed"] = np.int64(seed)
        fx["bias_scale"] = np.float64(bscale)
        fx["param_crc"] = np.int64(params_crc(params))
        if uni is not None:
            fx["uniforms"], fx["uniforms_retry"] = uni, uni2
        # keep fixtures small: probs/logits/masks as float32 [T,R,N] are fine at these sizes
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
        index[name] = dict(task=task, decode_type=dtype_, B=B, over=over)
        print(f"{name}: R mean {out['R'].mean():.5f}  idx[0][:8] {out['idxs'][0][:8]}  vars {len(requested)}")

    # Env driven directly: random feasible-ish action sequences incl. repeated depot visits and beams
    for name, task, W in [("env_vrptw10", "vrptw10", 1), ("env_vrp10", "vrp10", 1), ("env_vrptw10_beam", "vrptw10", 3),
                          ("env_upsold_vrptw10", "vrptw10", 1), ("env_upsold_vrptw10_beam", "vrptw10", 2)]:
        if only and name not in only:
            continue
        args = O.default_args(task, ups_old="upsold" in name)
        B, T = 12, 24
        data = O.create_dataset_ups_old(args["n_cust"], B, seed=77) if args["ups_old"] else O.create_dataset(task, B, seed=77)
        rnd = np.random.RandomState(5)
        R_ = B * W
        idx_seq = rnd.randint(0, args["n_nodes"], size=(T, R_))
        idx_seq[rnd.uniform(size=(T, R_)) < 0.25] = args["n_cust"]      # plenty of depot returns
        parent_seq = rnd.randint(0, W, size=(T, R_)) if W > 1 else None
        out = run_reference_env(args, data, idx_seq, parent_seq, W)
        out["data"], out["idx_seq"] = data, idx_seq
        if parent_seq is not None:
            out["parent_seq"] = parent_seq
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: mask max {out['mask'].max()}  load min {out['load'].min()}")

    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
