"""Fixture of the min_trucks reward branch (reward_func with a depot), produced by RUNNING THE REFERENCE'S OWN
reward_func (VRP/vrp_utils.py:257-307, VRPTW/vrptw_utils.py:361-414, UPS/vrptw_ups_utils.py:362-422,
UPS_old/vrptw_ups_utils.py:380-430) under oracle/tf1_shim.

    python tests/golden/make_golden_reward.py        # build container only (needs /root/reference)

Solutions are random walks over the nodes of random instances with frequent (and consecutive) depot visits, so that
the "count a stay at the depot once" logic is exercised."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)

from oracle import restatement as O  # noqa: E402


def make(task, n_cust, R, T, seed, ups=False, ups_old=False):
    from VRP.vrp_utils import reward_func as rf_vrp
    from VRPTW.vrptw_utils import reward_func as rf_tw
    from UPS.vrptw_ups_utils import reward_func as rf_tw_ups
    from UPS_old.vrptw_ups_utils import reward_func as rf_tw_ups_old
    rnd = np.random.RandomState(seed)
    if ups_old:     # x, y = latitude, longitude in radians, time windows in clicks - the units of the UPS_old Env fixtures
        data = O.create_dataset_ups_old(n_cust, R, seed=seed).astype(np.float32)
    else:
        data = O.create_dataset(f"{task}{n_cust}", R, seed=seed).astype(np.float32)
    pnt = data[:, :, :-1]                                    # input_pnt: x, y (, b_tw, e_tw)
    N = n_cust + 1
    idx = rnd.randint(0, N, size=(T, R))
    idx[rnd.uniform(size=(T, R)) < 0.35] = n_cust            # many depot visits, some of them consecutive
    idx[0, : R // 2] = n_cust
    actions = [pnt[np.arange(R), idx[t]] for t in range(T)]
    depot = pnt[:, n_cust]
    tf.reset_shim()
    rf = rf_tw_ups_old if ups_old else rf_tw_ups if ups else (rf_tw if task == "vrptw" else rf_vrp)
    out = rf([tf.constant(a) for a in actions], T, float(N), tf.constant(depot))
    ref = tf.Session().run(out)
    return dict(actions=np.stack(actions), depot=depot, reward=np.asarray(ref, np.float32), T=T, n_nodes=N, idx=idx)


if __name__ == "__main__":
    fx = {}
    for task, n_cust, R, T, seed in (("vrp", 10, 64, 20, 101), ("vrptw", 20, 64, 40, 102)):
        for k, v in make(task, n_cust, R, T, seed).items():
            fx[f"{task}_{k}"] = v
    for k, v in make("vrptw", 20, 64, 40, 103, ups=True).items():    # the UPS VRPTW form: 100 x visits + route length
        fx[f"vrptwups_{k}"] = v
    for k, v in make("vrptw", 20, 64, 40, 104, ups_old=True).items():   # UPS_old: 70/n x depot steps + 30 x km / Euclid
        fx[f"vrptwupsold_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "reward_min_trucks.npz"), **fx)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in fx.items()})
