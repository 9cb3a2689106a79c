"""Shared helpers for the test-suite (fixtures loading, oracle access)."""
import json
import os

import numpy as np

from oracle import restatement as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "index.json")) as f:
    GOLDEN_INDEX = json.load(f)


def load_golden(name):
    fx = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    meta = GOLDEN_INDEX.get(name)
    if meta is None:
        return fx, None, None
    args = O.default_args(meta["task"], **meta["over"])
    params = O.init_params(args, seed=int(fx["param_seed"]), bias_scale=float(fx["bias_scale"]))
    return fx, args, params


def load_pkg():
    """The product package (directory name has a hyphen)."""
    import importlib
    return importlib.import_module("mit_rl-vrptw_b200")


def first_divergence(idx_a, idx_b):
    """[T,R] index arrays -> per-row first step at which they differ (T if never)."""
    T = idx_a.shape[0]
    neq = idx_a != idx_b
    return np.where(neq.any(0), neq.argmax(0), T)
