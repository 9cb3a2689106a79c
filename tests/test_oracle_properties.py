"""Cross-checks of the oracle itself (CPU): the vectorised Env against an independent scalar restatement, hand-computed
known answers, and properties of decoded tours (hypothesis)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import restatement as O
from oracle.scalar_env import ScalarEnv, tour_length

ROOT_REF = None


@pytest.mark.parametrize("task", ["vrp10", "vrptw10", "vrptw20"])
def test_vector_env_equals_independent_scalar_env(task):
    args = O.default_args(task)
    tw = args["task_name"] == "vrptw"
    B, T = 24, 2 * args["n_nodes"]
    data = O.create_dataset(task, B, seed=31)
    rnd = np.random.RandomState(4)
    env = O.EnvOracle(args, data)
    env.reset(1)
    scal = [ScalarEnv(data[b], args["capacity"], tw) for b in range(B)]
    for b in range(B):
        np.testing.assert_array_equal(np.array(scal[b].reset()), env.mask[b])
    for t in range(T):
        idx = rnd.randint(0, args["n_nodes"], size=B)
        idx[rnd.uniform(size=B) < 0.3] = args["n_cust"]
        st_ = env.step(idx[:, None])
        for b in range(B):
            d_sat, mask = scal[b].step(int(idx[b]))
            assert d_sat == st_.d_sat[b]
            np.testing.assert_array_equal(np.array(mask), st_.mask[b], err_msg=f"{task} step {t} instance {b}")
            np.testing.assert_array_equal(np.array(scal[b].demand), st_.demand[b])
            assert scal[b].load == st_.load[b]
            if tw:
                assert scal[b].time == st_.time[b]


def test_env_known_answers_by_hand():
    """3 customers + depot at the origin, capacity 5; every number below was worked out by hand."""
    #            x    y    b    e    demand
    inst = [[3.0, 4.0, 0.0, 6.0, 4.0],     # 0: 5 away from the depot
            [0.0, 1.0, 2.0, 9.0, 3.0],     # 1
            [6.0, 8.0, 0.0, 9.9, 2.0],     # 2: 10 away from the depot
            [0.0, 0.0, 0.0, 99.0, 0.0]]    # depot
    args = dict(O.default_args("vrptw10"), n_nodes=4, n_cust=3, capacity=5)
    env = O.EnvOracle(args, np.array([inst]))
    s = env.reset(1)
    assert s.mask.tolist() == [[0, 0, 0, 1]] and s.load.tolist() == [5]
    s = env.step(np.array([[0]]))                       # serve node 0 fully: load 1, time 5
    assert s.d_sat.tolist() == [4] and s.load.tolist() == [1] and s.time.tolist() == [5]
    # node 0: demand 0 -> 1; node 1: arrival 5 + sqrt(9+9)=9.24 > 9 -> 1; node 2: 5 + 5 = 10 > 9.9 -> 1; depot 0
    assert s.mask.tolist() == [[1, 1, 1, 0]]
    s = env.step(np.array([[3]]))                       # back to the depot: refill, clock to zero
    assert s.load.tolist() == [5] and s.time.tolist() == [0]
    # at the depot with demand left: depot masked; node 2: 0 + 10 > 9.9 -> late
    assert s.mask.tolist() == [[1, 0, 1, 1]]
    s = env.step(np.array([[1]]))                       # node 1: wait for the window to open at 2
    assert s.time.tolist() == [2] and s.load.tolist() == [2] and s.d_sat.tolist() == [3]
    s = env.step(np.array([[2]]))                       # split delivery is legal even when late-masked
    assert s.d_sat.tolist() == [2] and s.load.tolist() == [0]
    assert s.demand.tolist() == [[0, 0, 0, 0]]
    # time is now 2 + sqrt(85) = 11.22: every customer is also late -> demand 0 + load 0 + late = 3
    assert s.mask.tolist()[0][:3] == [3, 3, 3] and s.mask.tolist()[0][3] == 0
    # closed tour depot-free: 0 -> depot -> 1 -> 2 -> (back to 0)
    pts = [(3.0, 4.0), (0.0, 0.0), (0.0, 1.0), (6.0, 8.0)]
    want = 5 + 1 + np.sqrt(36 + 49) + 5
    assert abs(float(tour_length(pts)) - want) < 1e-5
    acts = [np.array([[p[0], p[1], 0, 0]], np.float32) for p in pts]
    np.testing.assert_allclose(O.reward_func(acts, True), [want], rtol=1e-6)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), task=st.sampled_from(["vrp10", "vrptw10", "vrptw20"]), bias=st.sampled_from([0.0, 0.3]))
def test_greedy_rollout_properties(seed, task, bias):
    args = O.default_args(task)
    params = O.init_params(args, seed=seed, bias_scale=bias)
    data = O.create_dataset(task, 12, seed=seed + 1)
    res = O.rollout(params, args, data, "greedy", keep_probs=True)
    T, B = res.idxs.shape
    # a chosen node is never masked; probabilities are a distribution; masked nodes have probability exactly 0
    rows = np.arange(B)
    for t in range(T):
        assert (res.masks[t, rows, res.idxs[t]] == 0).all()
        np.testing.assert_allclose(res.probs[t].sum(1), 1.0, rtol=1e-5)
        assert (res.probs[t][res.masks[t] > 0] == 0).all()
    # demand only decreases, stays non-negative; loads within [0, capacity]; finished instances stay at the depot
    assert (res.final_state.demand >= 0).all() and (res.final_state.demand <= data[:, :, -1]).all()
    assert (res.final_state.load >= 0).all() and (res.final_state.load <= args["capacity"]).all()
    done = res.final_state.demand.sum(1) == 0
    last_customer = np.array([np.max(np.nonzero(res.idxs[:, b] != args["n_cust"])[0], initial=-1) for b in range(B)])
    for b in np.nonzero(done)[0]:
        assert (res.idxs[last_customer[b] + 1:, b] == args["n_cust"]).all()
    # the reported cost is the closed tour over the visited coordinates
    for b in range(min(B, 4)):
        pts = [tuple(data[b, i, :2].astype(np.float32)) for i in res.idxs[:, b]]
        assert abs(float(tour_length(pts)) - float(res.R[b])) <= 1e-4 * max(1.0, float(res.R[b]))


def _scalar_min_trucks(route, depot, n_nodes, xy):
    """Independent scalar restatement of reward_func with a depot (VRPTW/vrptw_utils.py:384-395, 410-412), one instance,
    plain Python floats rounded to fp32 where the reference's fp32 ops round."""
    f = np.float32
    counter, visits = f(0), f(1.0 if route[0][0] == depot[0] else 0.0)
    for p in route[1:]:
        interm = f(1.0 if p[0] == depot[0] else 0.0)
        counter = f(f(counter * interm) + interm)
        visits = f(visits + f(interm * f(1.0 if counter < 1.5 else 0.0)))
    maxlen, length, prev = f(0), f(0), route[-1]
    for p in route:
        m = f(0)
        for c in range(len(depot)):
            m = f(m + f(f(depot[c] - p[c]) ** 2))
        maxlen = f(maxlen + f(np.power(m, f(0.5))))
        s = f(0)
        for c in range(xy):
            s = f(s + f(f(prev[c] - p[c]) ** 2))
        length = f(length + f(np.power(s, f(0.5))))
        prev = p
    return f(f(70.0) * f(f(1.0 / n_nodes) * visits) + f(f(30.0) * f(length / maxlen)))


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(3, 25), st.booleans())
def test_min_trucks_reward_equals_independent_scalar_restatement(seed, T, tw):
    rnd = np.random.RandomState(seed)
    A, N = (4 if tw else 2), 7
    nodes = np.round(rnd.uniform(0, 1, size=(N, A)), 3).astype(np.float32)
    depot = nodes[-1]
    idx = rnd.randint(0, N, size=T)
    idx[rnd.uniform(size=T) < 0.4] = N - 1
    route = [nodes[i] for i in idx]
    if all((p == depot).all() for p in route):
        route[0] = nodes[0]                       # the reference divides by the summed distance to the depot
    acts = [p[None, :] for p in route]
    got = O.reward_func_min_trucks(acts, depot[None, :], T, float(N), tw)[0]
    want = _scalar_min_trucks(route, depot, float(N), 2)
    np.testing.assert_allclose(got, want, rtol=3e-6)
