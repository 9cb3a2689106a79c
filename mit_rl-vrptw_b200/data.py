"""Synthetic instance generators = the input spec of the rollout path.

Restated from the reference's generators (same distributions and draw order):
  create_vrp_dataset        VRP/vrp_utils.py:41-52
  create_vrptw_dataset      VRPTW/vrptw_utils.py:46-67
  create_vrptw_ups_dataset  UPS/vrptw_ups_utils.py:44-64 (sklearn GaussianMixture.sample restated with a seeded
                            numpy stream from the parameters extracted by tools/extract_gmm.py)
All return float64 [n_problems, n_cust+1, input_dim]; the depot is the last node.
"""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def create_vrp_dataset(n_problems: int, n_cust: int, seed: int) -> np.ndarray:
    rnd = np.random.RandomState(seed)
    n = n_cust + 1
    x = rnd.uniform(0, 1, size=(n_problems, n, 2))
    d = rnd.randint(1, 10, [n_problems, n, 1])
    x = np.around(x, decimals=5)
    d[:, -1] = 0
    return np.concatenate([x, d], 2)


def create_vrptw_dataset(n_problems: int, n_cust: int, seed: int) -> np.ndarray:
    rnd = np.random.RandomState(seed)
    n = n_cust + 1
    x = np.around(rnd.uniform(0, 1, size=(n_problems, n, 2)), decimals=5)
    d = rnd.randint(1, 10, [n_problems, n, 1])
    b_tw, e_tw, dur_min, dur_max = 0, 8, 0.5, 1
    tw_begin = rnd.uniform(b_tw + dur_max, e_tw - dur_min * 2, size=(n_problems, n, 1))
    tw_end = tw_begin + rnd.uniform(dur_min, dur_max, size=(n_problems, n, 1))
    d[:, -1] = 0
    tw_begin[:, -1] = b_tw
    tw_end[:, -1] = e_tw
    return np.concatenate([x, tw_begin, tw_end, d], axis=2)


def _gmm_sample(params, n_problems: int, n_cust: int, rnd: np.random.RandomState) -> np.ndarray:
    """n_cust draws per problem from the mixture; within a problem samples are ordered by component
    (as sklearn's GaussianMixture.sample returns them)."""
    w, mu, cov = params["weights"], params["means"], params["covariances"]
    K, d = mu.shape
    counts = rnd.multinomial(n_cust, w / w.sum(), size=n_problems)          # [P, K]
    out = np.empty((n_problems, n_cust, d))
    start = np.concatenate([np.zeros((n_problems, 1), np.int64), np.cumsum(counts, 1)[:, :-1]], 1)
    for k in range(K):
        tot = int(counts[:, k].sum())
        if tot == 0:
            continue
        s = rnd.multivariate_normal(mu[k], cov[k], tot)
        prob = np.repeat(np.arange(n_problems), counts[:, k])
        within = np.arange(tot) - np.repeat(np.cumsum(counts[:, k]) - counts[:, k], counts[:, k])
        out[prob, start[prob, k] + within] = s
    return out


def create_vrptw_ups_dataset(n_problems: int, n_cust: int, seed: int) -> np.ndarray:
    params = np.load(os.path.join(_DATA, "gmm_cvrptw.npz"))
    rnd = np.random.RandomState(seed)
    speed = 0.1567
    s = _gmm_sample(params, n_problems, n_cust, rnd)                         # x, y, b_tw, e_tw, demand
    s[:, :, 4] = np.maximum(1, np.round(s[:, :, 4]))
    b = np.maximum(0, s[:, :, 2] * speed)
    e = np.maximum(0, s[:, :, 3] * speed)
    b = np.where(b < e, b, e - 50 * speed)
    e = np.minimum(e, (1000 - 50) * speed)
    s[:, :, 2], s[:, :, 3] = b, e
    depot = np.tile(np.array([0, 0, 0, 1000 * speed, 0.0])[None, None, :], [n_problems, 1, 1])
    return np.concatenate([s, depot], axis=1)


def create_vrp_ups_dataset(n_problems: int, n_cust: int, seed: int) -> np.ndarray:
    """UPS/vrp_ups_utils.py generator: GMM over (x, y, demand)."""
    params = np.load(os.path.join(_DATA, "gmm_cvrp.npz"))
    rnd = np.random.RandomState(seed)
    s = _gmm_sample(params, n_problems, n_cust, rnd)
    s[:, :, 2] = np.maximum(1, np.round(s[:, :, 2]))
    depot = np.zeros((n_problems, 1, 3))
    return np.concatenate([s, depot], axis=1)


def create_dataset(task_name: str, n_problems: int, n_cust: int, seed: int, ups: bool = False) -> np.ndarray:
    if task_name == "vrp":
        return create_vrp_ups_dataset(n_problems, n_cust, seed) if ups else create_vrp_dataset(n_problems, n_cust, seed)
    return create_vrptw_ups_dataset(n_problems, n_cust, seed) if ups else create_vrptw_dataset(n_problems, n_cust, seed)


def dataset_file_name(task_name: str, n_problems: int, n_nodes: int, data_type: str, ups: bool = False) -> str:
    """File names of the reference's generators: 'vrp-size-{B}-len-{N}-{type}.txt' (VRP/vrp_utils.py:36),
    'vrptw-size-...' (VRPTW/vrptw_utils.py:38), 'vrp-ups-size-...' / 'vrptw-ups-size-...' (UPS/*_utils.py:34-35)."""
    return "{}{}-size-{}-len-{}-{}.txt".format(task_name, "-ups" if ups else "", n_problems, n_nodes, data_type)


def load_or_create_dataset(task_name: str, n_problems: int, n_cust: int, data_dir: str, seed, data_type: str = "test",
                           ups: bool = False) -> np.ndarray:
    """create_VRP_dataset / create_VRPTW_dataset of the reference WITH their on-disk format (VRP/vrp_utils.py:10-53,
    VRPTW/vrptw_utils.py:10-67, UPS twins): an existing '<task>-size-B-len-N-<type>.txt' is loaded (np.loadtxt, one
    instance per line, N*D numbers), otherwise the instances are generated and saved with np.savetxt in that layout."""
    n_nodes = n_cust + 1
    D = 5 if task_name == "vrptw" else 3
    fname = os.path.join(data_dir, dataset_file_name(task_name, n_problems, n_nodes, data_type, ups))
    if os.path.exists(fname):
        return np.loadtxt(fname, delimiter=" ").reshape(-1, n_nodes, D)
    data = create_dataset(task_name, n_problems, n_cust, seed, ups=ups)
    os.makedirs(data_dir, exist_ok=True)
    np.savetxt(fname, data.reshape(-1, n_nodes * D))
    return data


def write_results(fname: str, all_output, task_name: str, n_nodes: int) -> None:
    """RLAgent._output_results (model/attention_agent.py:416-459): one line per decoded problem - the visited nodes
    'x y' (vrp) or 'x y b_tw e_tw' (vrptw) separated by blanks, starting at the depot, cut after the last customer has
    been seen (a node differs from the depot when |dx| >= 0.001 or |dy| >= 0.001) and closed with the depot - the format
    evaluation/benchmark reads."""
    k = 2 if task_name == "vrp" else 4
    with open(fname, "w") as f:
        for output in all_output:
            depot = output[0]
            nb_stop = 0
            for node in output:
                f.write(" ".join(str(v) for v in node[:k]) + " ")
                if abs(depot[0] - node[0]) >= 0.001 or abs(depot[1] - node[1]) >= 0.001:
                    nb_stop += 1
                if nb_stop == n_nodes - 1:
                    f.write(" ".join(str(v) for v in depot[:k]))
                    break
            f.write("\n")


def glorot_params(args: dict, seed: int = 1234) -> dict:
    """'random-init' weights: Glorot-uniform kernels and v, zero biases (tf.layers defaults),
    keyed by TF-1 variable name."""
    rnd = np.random.RandomState(seed)
    tw = args["task_name"].startswith("vrptw")
    D1, E, H = args["input_dim"] - 1, args["embedding_dim"], args["hidden_dim"]

    def glorot(fi, fo, shape):
        lim = np.sqrt(6.0 / (fi + fo))
        return rnd.uniform(-lim, lim, size=shape).astype(np.float32)

    p = {"Actor/Embedding/conv1d/kernel": glorot(D1, E, [1, D1, E]), "Actor/Embedding/conv1d/bias": np.zeros(E, np.float32)}
    lstm = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell"
    p[lstm + "/kernel"] = glorot(E + H, 4 * H, [E + H, 4 * H])
    p[lstm + "/bias"] = np.zeros(4 * H, np.float32)
    for pre in [f"Actor/Glimpse{i}" for i in range(args.get("n_glimpses", 0))] + ["Actor/Decoder/Attention"]:
        p[pre + "/v"] = glorot(1, H, [1, H])
        for blk in ["emb_d", "emb_ld"] + (["emb_time"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = glorot(1, H, [1, 1, H])
            p[f"{pre}/{blk}/bias"] = np.zeros(H, np.float32)
        for blk in ["proj_d", "proj_ld"] + (["proj_t"] if tw else []):
            p[f"{pre}/{blk}/kernel"] = glorot(H, H, [1, H, H])
            p[f"{pre}/{blk}/bias"] = np.zeros(H, np.float32)
        p[pre + "/proj_q/kernel"] = glorot(H, H, [H, H])
        p[pre + "/proj_q/bias"] = np.zeros(H, np.float32)
        p[pre + "/proj_ref/kernel"] = glorot(E, H, [1, E, H])
        p[pre + "/proj_ref/bias"] = np.zeros(H, np.float32)
    return p
