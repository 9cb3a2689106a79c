"""Parity on the HEADLINE configurations (BASELINE.json configs 3-5: VRPTW-50, VRPTW-100 uniform, UPS VRPTW-100) against
the CPU oracle, at a thousand instances each.  Needs a B200 (``-m gpu``).

The oracle's greedy tours in fp32 AND fp64 are committed (tests/golden/oracle_tours_*.npz, made by
tests/golden/make_oracle_tours.py; the fp32 oracle itself is pinned bit-for-bit on fixtures produced by the reference's own
code at N = 51 and N = 101, UPS classes included: tests/test_oracle_golden.py).  What is checked, per workload:

  1. every instance: the CUDA tour equals the fp32 oracle's up to the first divergence, and at a divergence the state seen
     by that step (masks) is bit-exact, the logits agree within 1e-4 relative, and the two candidates are a NEAR-TIE in the
     oracle's own probabilities (|dp| <= 1e-5 p_max) - i.e. "given identical logits the choice is identical";
  2. instances without divergence: tour cost within 1e-4 relative, and (on a sample) masks bit-exact and logits within
     1e-4 relative at EVERY step;
  3. fp64 adjudication: with random-init weights two fp32 evaluation orders of the same graph disagree on a few percent of
     the VRPTW-100 instances (fp32 oracle vs fp64 oracle: see the fixture), so north_star's ">= 99.9 % identical tours" is
     not reachable by ANY fp32 implementation on these inputs; the bar used instead is that CUDA agrees with the fp64
     tours at least as often as the as-written fp32 oracle does, minus half a point;
  4. a floor on the raw agreement with the fp32 oracle (measured value minus a margin; profiles/agreement_r2.txt).
"""
import importlib
import os

import numpy as np
import pytest

from oracle import restatement as O
from tests.helpers import GOLDEN, first_divergence, load_pkg

pytestmark = pytest.mark.gpu

# workload -> floor on the fraction of instances whose whole greedy tour equals the fp32 oracle's
FLOORS = {"ups_vrptw100": 0.955, "uniform_vrptw100": 0.975, "vrptw50": 0.99}
LOGIT_RTOL, LOGIT_ATOL = 1e-4, 2e-5


def _load(name):
    fx = dict(np.load(os.path.join(GOLDEN, f"oracle_tours_{name}.npz")))
    pkg = load_pkg()
    args = O.default_args(str(fx["task"]))
    params = pkg.data.glorot_params(args, seed=int(fx["weight_seed"]))
    data = pkg.data.create_dataset(args["task_name"], int(fx["B"]), args["n_cust"], int(fx["data_seed"]), ups=bool(fx["ups"]))
    assert int(np.abs(data).sum() * 1e6) % (1 << 40) == int(fx["data_crc"]), "instances regenerated from the seed differ"
    return fx, args, params, data


@pytest.mark.parametrize("name", sorted(FLOORS))
@pytest.mark.parametrize("gemm", ["auto"])
def test_headline_parity_vs_oracle(name, gemm):
    import torch
    fx, args, params, data = _load(name)
    pkg = load_pkg()
    eng = pkg.Engine(args, gemm=gemm)
    eng.set_weights(params)
    x = torch.from_numpy(data.astype(np.float32)).cuda()
    plain = eng.rollout(x, "greedy")                                   # the measured path (fast per-row tail)
    out = eng.rollout(x, "greedy", want_logprobs=True, trace=True)     # every node evaluated, masks / logits dumped
    torch.cuda.synchronize()
    assert torch.equal(plain.idxs, out.idxs) and torch.equal(plain.R, out.R)
    idx = out.idxs.cpu().numpy().astype(np.int64)
    T, B = idx.shape
    ref32, ref64 = fx["idx32"].astype(np.int64), fx["idx64"].astype(np.int64)
    div = first_divergence(idx, ref32)
    agree32 = div == T
    # (2) costs of the agreeing instances
    np.testing.assert_allclose(out.R.cpu().numpy()[agree32], fx["R32"][agree32], rtol=1e-4)
    # (1) + (2): re-run the oracle with its per-step dumps on every diverging instance and on a sample of the others
    rows = np.concatenate([np.nonzero(~agree32)[0], np.nonzero(agree32)[0][:24]])
    ref = O.rollout(params, args, data[rows], "greedy", keep_probs=True)
    np.testing.assert_array_equal(ref.idxs, ref32[:, rows])              # the committed tours ARE the oracle's
    masks, logits = out.masks[:, rows].cpu().numpy(), out.logits[:, rows].cpu().numpy()
    worst = 0.0
    for j, r in enumerate(rows):
        upto = min(div[r] + 1, T)
        np.testing.assert_array_equal(masks[:upto, j], ref.masks[:upto, j])                       # bit-exact state
        um = ref.masks[:upto, j] == 0
        a, b = logits[:upto, j][um], ref.logits[:upto, j][um]
        np.testing.assert_allclose(a, b, rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), LOGIT_ATOL / LOGIT_RTOL))))
        if div[r] < T:
            t = div[r]
            p = ref.probs[t, j]
            assert abs(p[idx[t, r]] - p[ref32[t, r]]) <= 1e-5 * p.max(), f"{name}: instance {r} step {t} is not a near-tie"
    # (3) fp64 adjudication
    a_cuda64 = (idx == ref64).all(0).mean()
    a_3264 = (ref32 == ref64).all(0).mean()
    print(f"\n[parity {name}] B={B}: CUDA=fp32-oracle {agree32.mean():.4f}  CUDA=fp64-oracle {a_cuda64:.4f}  "
          f"fp32-oracle=fp64-oracle {a_3264:.4f}  divergences verified as near-ties {int((~agree32).sum())}  "
          f"max logit error {worst:.2e} (relative, floor {LOGIT_ATOL / LOGIT_RTOL})")
    assert a_cuda64 >= a_3264 - 0.005, f"{name}: CUDA agrees with fp64 on {a_cuda64:.4f}, the fp32 oracle on {a_3264:.4f}"
    # (4) floor
    assert agree32.mean() >= FLOORS[name], f"{name}: only {agree32.mean():.4f} identical greedy tours"
