"""The torch-CPU timing baseline (oracle/restatement_torch.py) computes the same rollout as the numpy oracle."""
import numpy as np
import pytest

from oracle import restatement as O
from oracle import restatement_torch as OT


@pytest.mark.parametrize("task,decode_type,B", [("vrp10", "greedy", 64), ("vrptw20", "greedy", 64), ("vrptw20", "beam_search", 16),
                                                ("vrptw10", "stochastic", 32), ("vrptw50", "greedy", 24)])
def test_torch_baseline_equals_numpy_oracle(task, decode_type, B):
    args = O.default_args(task, beam_width=3)
    params = O.init_params(args, seed=7, bias_scale=0.2)
    data = O.create_dataset(task, B, seed=3)
    W = 3 if decode_type == "beam_search" else 1
    rnd = np.random.RandomState(1)
    u1 = rnd.uniform(size=(args["decode_len"], B * W)).astype(np.float32)
    u2 = rnd.uniform(size=(args["decode_len"], B * W)).astype(np.float32)
    ref = O.rollout(params, args, data, decode_type, uniforms=u1, uniforms_retry=u2)
    R, idxs = OT.rollout(params, args, data, decode_type, uniforms=u1, uniforms_retry=u2)
    same = (idxs == ref.idxs).all(0)
    assert same.mean() >= 0.9          # fp32 near-ties (different BLAS / transcendental kernels) may move a tour
    np.testing.assert_allclose(R[same], ref.R[same], rtol=1e-5)
