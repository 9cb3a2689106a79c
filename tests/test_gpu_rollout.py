"""CUDA rollout (through the C ABI) against the golden fixtures produced by the reference's own code
and against the numpy oracle on seeded instances.  Needs a B200: run with ``-m gpu``.

Tolerances (BASELINE.json north_star): env state, masks and - given identical logits - chosen indices
bit-exact; logits / tour lengths within 1e-4 relative (fp32).
"""
import numpy as np
import pytest

from oracle import restatement as O
from tests.helpers import GOLDEN_INDEX, first_divergence, load_golden, load_pkg

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 1e-4      # north_star: "logits and tour lengths must agree within 1e-4 relative (fp32)"
LOGIT_ATOL = 2e-5      # u passes through 0: floor of the relative test
MASKED_ATOL = 0.02     # logit - 1e5*k: one fp32 ulp at 1e5 is 0.0078


def _engine(args, params, **kw):
    import torch
    pkg = load_pkg()
    eng = pkg.Engine(args, device=0, **kw)
    eng.set_weights(params)
    return eng, torch


def _compare_trace(out, fx, args, decode_type):
    """Per row: up to (and including) the first step at which the chosen node differs from the reference's, masks are
    bit-exact and logits / probabilities within tolerance; a divergence must be a near-tie of the two candidates in the
    reference's OWN probabilities (|dp| <= 1e-5 p_max); rows without divergence also match in cost, log-probabilities and
    final state.  The small fixtures (N <= 21) must not diverge at all."""
    idx = out.idxs.cpu().numpy().astype(np.int64)
    T, R = idx.shape
    ref_idx = fx["idxs"]
    div = first_divergence(idx, ref_idx)
    if args["n_nodes"] <= 21:
        assert (div == T).all(), f"indices diverge from the reference at steps {div[div < T]}"
    masks, logits, probs = out.masks.cpu().numpy(), out.logits.cpu().numpy(), out.probs.cpu().numpy()
    logprobs = out.logprobs.cpu().numpy()
    for r in range(R):
        upto = min(div[r] + 1, T)
        np.testing.assert_array_equal(masks[:upto, r], fx["masks"][:upto, r])                    # bit-exact
        um = fx["masks"][:upto, r] == 0
        np.testing.assert_allclose(logits[:upto, r][um], fx["logits"][:upto, r][um], rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
        np.testing.assert_allclose(logits[:upto, r][~um], fx["logits"][:upto, r][~um], rtol=0, atol=MASKED_ATOL)
        np.testing.assert_allclose(probs[:upto, r], fx["probs"][:upto, r], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(logprobs[:div[r], r], fx["logprobs"][:div[r], r], rtol=2e-4, atol=2e-5)
        if div[r] < T:
            p = fx["probs"][div[r], r]
            assert abs(p[idx[div[r], r]] - p[ref_idx[div[r], r]]) <= 1e-5 * p.max(), (r, div[r])
    ok = div == T
    assert ok.mean() >= 0.5
    np.testing.assert_allclose(out.R.cpu().numpy()[ok], fx["R"][ok], rtol=1e-4)
    np.testing.assert_array_equal(out.final_demand.cpu().numpy()[ok], fx["final_demand"][ok])   # bit-exact
    np.testing.assert_array_equal(out.final_load.cpu().numpy()[ok], fx["final_load"][ok])
    np.testing.assert_array_equal(out.final_mask.cpu().numpy()[ok], fx["final_mask"][ok])
    if args["task_name"] == "vrptw":
        np.testing.assert_array_equal(out.final_time.cpu().numpy()[ok], fx["final_time"][ok])


FUSED_CASES = [n for n, m in sorted(GOLDEN_INDEX.items())
               if m["decode_type"] in ("greedy", "stochastic") and not m["over"].get("n_glimpses") and not m["over"].get("ups_old")]
# glimpses (shared/decode_step.py:218-227) and the UPS_old Env physics: fused in the FP32-pipe kernel, whatever gemm= says
SPECIAL_CASES = [n for n, m in sorted(GOLDEN_INDEX.items())
                 if m["decode_type"] in ("greedy", "stochastic") and (m["over"].get("n_glimpses") or m["over"].get("ups_old"))]


@pytest.mark.parametrize("name", FUSED_CASES)
@pytest.mark.parametrize("rows_per_cta", [0, 64])
@pytest.mark.parametrize("gemm", ["tc4x", "tc3", "ffma"])
def test_rollout_matches_reference_fixture(name, rows_per_cta, gemm):
    fx, args, params = load_golden(name)
    meta = GOLDEN_INDEX[name]
    eng, torch = _engine(args, params, rows_per_cta=rows_per_cta, gemm=gemm)
    x = torch.from_numpy(fx["data"].astype(np.float32)).cuda()
    try:
        out = eng.rollout(x, meta["decode_type"], uniforms=fx.get("uniforms"), uniforms_retry=fx.get("uniforms_retry"),
                          want_logprobs=True, trace=True)
    except load_pkg().VrpB200Error as e:      # 64 rows per CTA do not fit shared memory at N = 101
        assert "shared memory" in str(e) and rows_per_cta == 64 and args["n_nodes"] > 51
        return
    torch.cuda.synchronize()
    _compare_trace(out, fx, args, meta["decode_type"])


@pytest.mark.parametrize("name", SPECIAL_CASES)
@pytest.mark.parametrize("rows_per_cta", [0, 32])
def test_fused_glimpses_and_ups_old_match_reference_fixture(name, rows_per_cta):
    """ONE persistent launch for the whole rollout with glimpses (1 and 2 modules, masked and un-masked, with C tanh on the
    pointer, VRP and VRPTW) and with the UPS_old Env (km travel time through cos, service time, tour length in km), against
    fixtures produced by the reference's own decode_step.py / UPS_old/vrptw_ups_utils.py."""
    fx, args, params = load_golden(name)
    meta = GOLDEN_INDEX[name]
    eng, torch = _engine(args, params, rows_per_cta=rows_per_cta)
    x = torch.from_numpy(fx["data"].astype(np.float32)).cuda()
    out = eng.rollout(x, meta["decode_type"], uniforms=fx.get("uniforms"), uniforms_retry=fx.get("uniforms_retry"),
                      want_logprobs=True, trace=True)
    torch.cuda.synchronize()
    assert eng.last_launch()["kernels"] == 1 and eng.last_launch()["family"] == 1
    _compare_trace(out, fx, args, meta["decode_type"])
    fast = eng.rollout(x, meta["decode_type"], uniforms=fx.get("uniforms"), uniforms_retry=fx.get("uniforms_retry"))
    assert torch.equal(fast.idxs, out.idxs) and torch.equal(fast.R, out.R)      # masked nodes skipped, early exit: same tours


@pytest.mark.parametrize("name", [n for n in FUSED_CASES if GOLDEN_INDEX[n]["decode_type"] == "greedy"])
def test_fast_path_equals_trace_path(name):
    """Skipping masked nodes + early exit must not change any output."""
    fx, args, params = load_golden(name)
    eng, torch = _engine(args, params)
    x = torch.from_numpy(fx["data"].astype(np.float32)).cuda()
    fast = eng.rollout(x, "greedy", want_logprobs=True, final_state=True, want_steps=True)
    full = eng.rollout(x, "greedy", want_logprobs=True, trace=True)
    torch.cuda.synchronize()
    assert torch.equal(fast.idxs, full.idxs)
    assert torch.equal(fast.R, full.R)
    assert torch.equal(fast.logprobs, full.logprobs)
    assert torch.equal(fast.final_demand, full.final_demand)
    assert torch.equal(fast.final_mask, full.final_mask)
    assert torch.equal(fast.final_load, full.final_load)
    assert int(fast.steps.max()) <= args["decode_len"]


@pytest.mark.parametrize("gemm", ["tc4x"])
@pytest.mark.parametrize("task,B,seed,ups", [("vrptw100", 1500, 21, True), ("vrptw100", 700, 22, False), ("vrptw20", 900, 23, False),
                                             ("vrp50", 500, 24, False), ("vrp100", 300, 25, True), ("vrptw50", 777, 26, True)])
def test_plain_greedy_fast_tail_equals_exact_tail(gemm, task, B, seed, ups):
    """The MEASURED path: plain greedy decoding (no log-probabilities, traces or final state requested) runs rollout4's fast
    per-row tail - arg-max over the item list, exact soft-max only on near-ties, feasibility only for customers with demand
    left.  It must give bit-identical tours and costs to the exact tail (which the other tests pin against the oracle),
    in particular on UPS instances, whose saturated logits produce ties and near-ties all the time."""
    pkg = load_pkg()
    args = O.default_args(task)
    params = pkg.data.glorot_params(args, seed=seed)
    data = pkg.data.create_dataset(args["task_name"], B, args["n_cust"], seed, ups=ups).astype(np.float32)
    eng, torch = _engine(args, params, gemm=gemm)
    x = torch.from_numpy(data).cuda()
    fast = eng.rollout(x, "greedy")
    exact = eng.rollout(x, "greedy", want_logprobs=True, final_state=True)
    torch.cuda.synchronize()
    assert torch.equal(fast.idxs, exact.idxs)
    assert torch.equal(fast.R, exact.R)


@pytest.mark.parametrize("gemm", ["tc4x"])
def test_fast_tail_hands_rows_without_unmasked_node_to_the_exact_tail(gemm):
    """Time windows that close almost immediately: after its first customer a vehicle finds every other customer too late,
    returns to the depot, and there NO node is un-masked (the depot is masked while demand is left).  The reference then
    takes the arg-max over logit - 1e5 x mask COUNT (decode_step.py:231-232); the fast tail keeps no counts and must hand
    such rows to the exact tail (masks recomputed from the row state).  Plain decoding must equal the exact tail, and the
    situation must actually occur (some instance decodes a node whose mask count is non-zero)."""
    pkg = load_pkg()
    args = O.default_args("vrptw20")
    params = pkg.data.glorot_params(args, seed=31)
    data = pkg.data.create_dataset("vrptw", 96, args["n_cust"], 32).astype(np.float32)
    data[:, :-1, 2] = 0.0          # b_tw
    data[:, :-1, 3] = 0.02         # e_tw: only reachable as the first stop after the (TW-free) initial mask ... never again
    eng, torch = _engine(args, params, gemm=gemm)
    x = torch.from_numpy(data).cuda()
    fast = eng.rollout(x, "greedy")
    exact = eng.rollout(x, "greedy", want_logprobs=True, trace=True)
    torch.cuda.synchronize()
    assert torch.equal(fast.idxs, exact.idxs)
    assert torch.equal(fast.R, exact.R)
    # the trace shows rows whose every node is masked at some step, and the chosen node has a non-zero mask count there
    allmasked = (exact.masks > 0).all(2)                       # [T, R]
    assert bool(allmasked.any())
    chosen_mask = torch.gather(exact.masks, 2, exact.idxs.long().unsqueeze(2)).squeeze(2)
    assert bool((chosen_mask[allmasked] > 0).all())
    ref = O.rollout(params, args, data.astype(np.float64), "greedy")
    same = (exact.idxs.cpu().numpy().astype(np.int64) == ref.idxs).all(0)
    assert same.mean() >= 0.95


@pytest.mark.parametrize("task,B,seed", [("vrp10", 128, 1), ("vrptw20", 512, 2), ("vrp20", 300, 3),
                                         ("vrptw50", 256, 4), ("vrp100", 96, 5), ("vrptw100", 200, 6)])
@pytest.mark.parametrize("bias", [0.0, 0.3])
@pytest.mark.parametrize("gemm", ["tc4x", "tc3", "ffma"])
def test_greedy_vs_oracle_seeded(task, B, seed, bias, gemm):
    """Ragged batch sizes, every task size; per-instance comparison up to the first near-tie."""
    args = O.default_args(task)
    params = O.init_params(args, seed=100 + seed, bias_scale=bias)
    data = O.create_dataset(task, B, seed=seed)
    eng, torch = _engine(args, params, gemm=gemm)
    out = eng.rollout(torch.from_numpy(data.astype(np.float32)).cuda(), "greedy", want_logprobs=True, trace=True)
    torch.cuda.synchronize()
    ref = O.rollout(params, args, data, "greedy", keep_probs=True)
    idx = out.idxs.cpu().numpy().astype(np.int64)
    T = idx.shape[0]
    div = first_divergence(idx, ref.idxs)
    masks, logits = out.masks.cpu().numpy(), out.logits.cpu().numpy()
    for r in range(B):
        upto = min(div[r] + 1, T)     # the step of the divergence still sees identical state
        np.testing.assert_array_equal(masks[:upto, r], ref.masks[:upto, r])
        um = ref.masks[:upto, r] == 0
        np.testing.assert_allclose(logits[:upto, r][um], ref.logits[:upto, r][um], rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
        if div[r] < T:   # a legitimate divergence is a near-tie of the two candidates in the reference's own probs
            t = div[r]
            p = ref.probs[t, r]
            assert abs(p[idx[t, r]] - p[ref.idxs[t, r]]) <= 1e-5 * p.max(), (task, r, t)
    agree = (div == T)
    np.testing.assert_allclose(out.R.cpu().numpy()[agree], ref.R[agree], rtol=1e-4)
    np.testing.assert_array_equal(out.final_demand.cpu().numpy()[agree], ref.final_state.demand[agree])
    np.testing.assert_array_equal(out.final_load.cpu().numpy()[agree], ref.final_state.load[agree])
    np.testing.assert_array_equal(out.final_mask.cpu().numpy()[agree], ref.final_state.mask[agree])
    assert agree.mean() >= 0.98, f"{task}: only {agree.mean():.3%} identical greedy tours"


def test_stochastic_given_uniforms_vs_oracle():
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=77, bias_scale=0.2)
    B, T = 256, args["decode_len"]
    data = O.create_dataset("vrptw20", B, seed=9)
    rnd = np.random.RandomState(3)
    u1 = rnd.uniform(size=(T, B)).astype(np.float32)
    u2 = rnd.uniform(size=(T, B)).astype(np.float32)
    eng, torch = _engine(args, params)
    out = eng.rollout(torch.from_numpy(data.astype(np.float32)).cuda(), "stochastic", uniforms=u1, uniforms_retry=u2,
                      want_logprobs=True, final_state=True)
    torch.cuda.synchronize()
    ref = O.rollout(params, args, data, "stochastic", uniforms=u1, uniforms_retry=u2, per_row_retry=True)
    div = first_divergence(out.idxs.cpu().numpy().astype(np.int64), ref.idxs)
    agree = div == T
    assert agree.mean() >= 0.98
    np.testing.assert_allclose(out.R.cpu().numpy()[agree], ref.R[agree], rtol=1e-4)
    np.testing.assert_allclose(out.logprobs.cpu().numpy()[:, agree], ref.logprobs[:, agree], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("gemm", ["tc4x", "tc3", "ffma"])
def test_stochastic_whole_batch_retry_is_the_reference_policy(gemm):
    """Opt-in whole-batch retry (attention_agent.py:148-167): when ANY row fails its first draw, EVERY row of the batch decides
    that step with the second set of uniforms.  Two draws are forced to fail (u = 2 > any CDF) at steps 0 and 3; the tours
    must equal the oracle's as-written policy (per_row_retry=False), differ from the per-row policy, and the library must
    report exactly the redrawn steps."""
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=78, bias_scale=0.2)
    B, T = 96, args["decode_len"]
    data = O.create_dataset("vrptw20", B, seed=10)
    rnd = np.random.RandomState(4)
    u1 = rnd.uniform(size=(T, B)).astype(np.float32)
    u2 = rnd.uniform(size=(T, B)).astype(np.float32)
    u1[0, 5] = 2.0
    u1[3, 40] = 2.0
    eng, torch = _engine(args, params, gemm=gemm)
    x = torch.from_numpy(data.astype(np.float32)).cuda()
    per_row = eng.rollout(x, "stochastic", uniforms=u1, uniforms_retry=u2)
    eng.set_sampling_retry(True)
    whole = eng.rollout(x, "stochastic", uniforms=u1, uniforms_retry=u2, want_logprobs=True)
    assert eng.last_redraws() == 2
    eng.set_sampling_retry(False)
    ref = O.rollout(params, args, data, "stochastic", uniforms=u1, uniforms_retry=u2, per_row_retry=False)
    assert sorted(ref.retried) == [0, 3]
    idx = whole.idxs.cpu().numpy().astype(np.int64)
    agree = first_divergence(idx, ref.idxs) == T
    assert agree.mean() >= 0.97
    np.testing.assert_array_equal(idx[0], ref.idxs[0])          # the redrawn first step: every row took its second uniform
    np.testing.assert_allclose(whole.R.cpu().numpy()[agree], ref.R[agree], rtol=1e-4)
    assert (per_row.idxs.cpu().numpy()[0] != idx[0]).any()      # ... which the per-row policy does not do


def test_stochastic_in_kernel_stream_is_valid_and_reproducible():
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=78)
    data = O.create_dataset("vrptw20", 300, seed=10)
    eng, torch = _engine(args, params)
    x = torch.from_numpy(data.astype(np.float32)).cuda()
    a = eng.rollout(x, "stochastic", seed=1234, final_state=True)
    b = eng.rollout(x, "stochastic", seed=1234)
    c = eng.rollout(x, "stochastic", seed=99)
    torch.cuda.synchronize()
    assert torch.equal(a.idxs, b.idxs) and torch.equal(a.R, b.R)
    assert not torch.equal(a.idxs, c.idxs)
    # every sampled tour is feasible: replay it through the oracle env, chosen nodes are never masked
    env = O.EnvOracle(args, data)
    env.reset(1)
    idx = a.idxs.cpu().numpy().astype(np.int64)
    for t in range(idx.shape[0]):
        assert (env.mask[np.arange(300), idx[t]] == 0).all(), f"masked node chosen at step {t}"
        env.step(idx[t][:, None])
    np.testing.assert_array_equal(a.final_demand.cpu().numpy(), env.demand)


def test_host_entry_point_equals_device_entry_point():
    args = O.default_args("vrptw20")
    params = O.init_params(args, seed=5)
    data = O.create_dataset("vrptw20", 1000, seed=11).astype(np.float32)
    eng, torch = _engine(args, params)
    idx_h, cost_h = eng.rollout_host(data, "greedy")
    out = eng.rollout(torch.from_numpy(data).cuda(), "greedy")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(idx_h, out.idxs.cpu().numpy())
    np.testing.assert_array_equal(cost_h, out.R.cpu().numpy())


@pytest.mark.parametrize("gemm", ["tc4x", "tc3"])
def test_tensor_core_path_equals_fp32_path_up_to_near_ties(gemm):
    """The tcgen05 (3xTF32 split) GEMMs against the FP32-pipe GEMMs of the same kernel: logits within 2e-6,
    identical tours except at near-ties, every rows_per_cta."""
    args = O.default_args("vrptw50")
    params = O.init_params(args, seed=41, bias_scale=0.2)
    data = O.create_dataset("vrptw50", 700, seed=12).astype(np.float32)
    import torch
    x = torch.from_numpy(data).cuda()
    ref_eng, _ = _engine(args, params, gemm="ffma")
    ref = ref_eng.rollout(x, "greedy", trace=True)
    pkg = load_pkg()
    for rb in (16, 32, 48, 64):
        eng, _ = _engine(args, params, gemm=gemm, rows_per_cta=rb)
        try:
            out = eng.rollout(x, "greedy", trace=True)
        except pkg.VrpB200Error as e:     # this variant does not fit that many rows in shared memory at N=51
            assert "shared memory" in str(e) and rb == 64
            continue
        torch.cuda.synchronize()
        same = (out.idxs == ref.idxs).all(0)
        assert same.float().mean().item() >= 0.97, rb
        um = (ref.masks == 0)[:, same]
        d = (out.logits[:, same] - ref.logits[:, same]).abs()[um]
        assert d.max().item() < 2e-5, (rb, d.max().item())   # north_star tolerance: 1e-4 relative
        assert torch.equal(out.masks[:, same], ref.masks[:, same])
        np.testing.assert_allclose(out.R[same].cpu().numpy(), ref.R[same].cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("name", [n for n, m in sorted(GOLDEN_INDEX.items()) if m["decode_type"] == "beam_search"])
def test_fused_beam_search_matches_reference_fixture(name):
    """Beam search inside the fused kernel (top-k over W x N candidates per instance, parent gather of the Env state,
    LSTM slot not re-ordered) against the fixture produced by the reference's own code."""
    fx, args, params = load_golden(name)
    eng, torch = _engine(args, params)
    W, B = args["beam_width"], fx["data"].shape[0]
    x = torch.from_numpy(fx["data"].astype(np.float32)).cuda()
    out = eng.rollout(x, "beam_search", beam_width=W, want_logprobs=True, final_state=True)
    torch.cuda.synchronize()
    ref = O.rollout(params, args, fx["data"], "beam_search")
    np.testing.assert_array_equal(ref.idxs, fx["idxs"])
    idx, par = out.idxs.cpu().numpy(), out.beam_parents.cpu().numpy()
    np.testing.assert_array_equal(idx, fx["idxs"])
    ok = par == ref.beam_parents          # twin beams (identical state, exact tie in the reference) may swap slots
    assert ok.mean() > 0.9
    # after a swap the two slots carry each other's (un-reordered) LSTM state: compare probabilities up to the first swap
    t_first = int(np.argmax(~ok.all(1))) if not ok.all() else ok.shape[0]
    np.testing.assert_allclose(out.logprobs.cpu().numpy()[:t_first], fx["logprobs"][:t_first], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(np.sort(out.R.cpu().numpy().reshape(W, B), 0), np.sort(fx["R"].reshape(W, B), 0), rtol=1e-4)
    if ok.all():
        np.testing.assert_allclose(out.R.cpu().numpy(), fx["R"], rtol=1e-4)
        np.testing.assert_array_equal(out.final_demand.cpu().numpy(), fx["final_demand"])
        np.testing.assert_array_equal(out.final_mask.cpu().numpy(), fx["final_mask"])


@pytest.mark.parametrize("task,B,W,seed", [("vrptw20", 300, 5, 1), ("vrptw50", 200, 10, 2), ("vrp10", 100, 2, 3), ("vrptw100", 40, 4, 4)])
def test_fused_beam_search_vs_oracle(task, B, W, seed):
    args = O.default_args(task, beam_width=W)
    params = O.init_params(args, seed=50 + seed, bias_scale=0.2)
    data = O.create_dataset(task, B, seed=seed)
    eng, torch = _engine(args, params)
    out = eng.rollout(torch.from_numpy(data.astype(np.float32)).cuda(), "beam_search", beam_width=W)
    torch.cuda.synchronize()
    ref = O.rollout(params, args, data, "beam_search")
    best = out.R.cpu().numpy().reshape(W, B).min(0)
    ref_best = O.best_of_beams(ref.R, B, W)
    close = np.isclose(best, ref_best, rtol=1e-4)
    assert close.mean() >= 0.97, f"best-of-beams differs on {1 - close.mean():.2%} of instances"
    same = (out.idxs.cpu().numpy() == ref.idxs).all(0) & (out.beam_parents.cpu().numpy() == ref.beam_parents).all(0)
    # fp32 near-ties among the W x N candidates accumulate with T: 180 steps x 404 candidates at vrptw100
    assert same.reshape(W, B).all(0).mean() >= (0.6 if task == "vrptw100" else 0.9)
    # every beam is a feasible tour: replay it (back-tracked) through the oracle Env
    idx, par = out.idxs.cpu().numpy().astype(np.int64), out.beam_parents.cpu().numpy().astype(np.int64)
    env = O.EnvOracle(args, data)
    env.reset(W)
    for t in range(idx.shape[0]):
        src = np.tile(np.arange(B), W) + B * par[t]
        assert (env.mask[src, idx[t]] == 0).all(), f"masked node chosen at step {t}"
        env.step(idx[t][:, None], par[t][:, None])


def test_results_do_not_depend_on_rows_per_cta_or_chunking():
    """rows_per_cta is a tuning knob: every reduction that feeds a decision is done in one fixed order, so the tours
    are bit-identical for every block size - and the chunked host entry point equals the one-shot device call."""
    import torch
    pkg = load_pkg()
    args = O.default_args("vrptw50")
    params = O.init_params(args, seed=61)
    data = O.create_dataset("vrptw50", 3000, seed=21).astype(np.float32)
    x = torch.from_numpy(data).cuda()
    outs = []
    for rb in (16, 32, 48, 64):
        eng, _ = _engine(args, params, rows_per_cta=rb)
        try:
            outs.append(eng.rollout(x, "greedy", want_logprobs=True))
        except pkg.VrpB200Error as e:
            assert "shared memory" in str(e)
    torch.cuda.synchronize()
    assert len(outs) >= 3
    for o in outs[1:]:
        assert torch.equal(o.idxs, outs[0].idxs) and torch.equal(o.R, outs[0].R) and torch.equal(o.logprobs, outs[0].logprobs)


def test_errors_are_reported_not_thrown():
    pkg = load_pkg()
    args = O.default_args("vrptw10")
    with pytest.raises(pkg.VrpB200Error):
        pkg.Engine(dict(args, hidden_dim=64))
    eng = pkg.Engine(args)
    import torch
    with pytest.raises(pkg.VrpB200Error, match="set_weights"):
        eng.rollout(torch.zeros(4, 11, 5).cuda())
    p = O.init_params(args, seed=1)
    p.pop("Actor/Decoder/Attention/v")
    with pytest.raises(pkg.VrpB200Error, match="Attention/v"):
        eng.set_weights(p)


@pytest.mark.parametrize("task,B,want_rb", [("vrp10", 128, 16), ("vrptw20", 2000, 16), ("vrptw20", 4096, 32), ("vrptw100", 8192, 32),
                                            ("vrptw50", 16384, 48), ("vrptw100", 30000, 48)])
def test_rows_per_cta_follow_the_wave_cost_model(task, B, want_rb):
    """The automatic block size minimises waves x (85 + rows) (profiles/r2_rows_per_cta_sweep.txt): 16-row blocks only while
    they fit one wave, 32 rows for 4 096 / 8 192 instances, 48 rows for large batches - and the choice never changes a result."""
    pkg = load_pkg()
    args = O.default_args(task)
    params = pkg.data.glorot_params(args, seed=3)
    data = pkg.data.create_dataset(args["task_name"], B, args["n_cust"], 4).astype(np.float32)
    eng, torch = _engine(args, params)
    x = torch.from_numpy(data).cuda()
    out = eng.rollout(x, "greedy")
    ll = eng.last_launch()
    assert ll["rows_per_cta"] == want_rb and ll["family"] == 6 and ll["grid"] == -(-B // want_rb)
    other = 48 if want_rb != 48 else 32
    eng2, _ = _engine(args, params, rows_per_cta=other)
    out2 = eng2.rollout(x, "greedy")
    assert torch.equal(out.idxs, out2.idxs) and torch.equal(out.R, out2.R)


@pytest.mark.parametrize("task,B,seed,ups", [("vrptw100", 1500, 31, True), ("vrptw100", 700, 32, False), ("vrptw20", 900, 33, False),
                                             ("vrp50", 500, 34, False), ("vrptw50", 777, 35, True)])
def test_sampling_fast_tail_equals_exact_tail(task, B, seed, ups):
    """Sampling without traces / final state runs rollout4's fast per-row tail (soft-max in the exact tail's summation order,
    CDF over the item list, feasibility only for customers with demand left).  Given the same uniforms - the caller's or the
    in-kernel stream's - it must draw the same nodes, costs and log-probabilities as the exact tail, bit for bit."""
    pkg = load_pkg()
    args = O.default_args(task)
    params = pkg.data.glorot_params(args, seed=seed)
    data = pkg.data.create_dataset(args["task_name"], B, args["n_cust"], seed, ups=ups).astype(np.float32)
    eng, torch = _engine(args, params)
    x = torch.from_numpy(data).cuda()
    rnd = np.random.RandomState(seed)
    T = args["decode_len"]
    u1, u2 = rnd.uniform(size=(T, B)).astype(np.float32), rnd.uniform(size=(T, B)).astype(np.float32)
    for kw in (dict(uniforms=u1, uniforms_retry=u2), dict(seed=7)):
        fast = eng.rollout(x, "stochastic", want_logprobs=True, **kw)
        plain = eng.rollout(x, "stochastic", **kw)
        exact = eng.rollout(x, "stochastic", want_logprobs=True, final_state=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(fast.idxs, exact.idxs) and torch.equal(plain.idxs, exact.idxs)
        assert torch.equal(fast.R, exact.R) and torch.equal(plain.R, exact.R)
        assert torch.equal(fast.logprobs, exact.logprobs)
    assert (fast.idxs != eng.rollout(x, "stochastic", seed=8).idxs).any()      # another seed, another sample
