"""A lazy numpy stand-in for the slice of the TensorFlow-1 API that the reference's
rollout path calls.  TEST INFRASTRUCTURE ONLY (lives under ``oracle/``).

Why it exists: TensorFlow cannot be installed in this image (no wheel, no network, and the
reference needs TF 1.x).  To pin the oracle to the reference's *own code* rather than to a
reading of it, ``tests/golden/make_golden.py`` puts this directory on ``sys.path`` and imports
the unmodified reference modules (``VRPTW/vrptw_utils.py``, ``shared/decode_step.py``,
``model/attention_agent.py`` …) from /root/reference.  ``tf.*`` calls build a small deferred
graph of numpy closures; ``Session.run(fetches, feed_dict)`` evaluates it.

What this is NOT: it is not TensorFlow.  Each op is implemented from TF-1's documented
semantics (shape/broadcast/gather rules, top_k tie order, BasicLSTMCell gate order i,j,f,o,
``_scope`` naming of tf.layers variables) with numpy float32 kernels.  Variables are not
initialised here: every ``get_variable``/layer weight is looked up BY NAME AND SHAPE in
``VARIABLE_STORE`` (filled by the caller), so a name or shape the oracle got wrong fails loudly.
"""
from __future__ import annotations

import builtins as _b
import contextlib
import itertools
import types

import numpy as np

builtins_range = _b.range
float32, float64, int32, int64, bool = np.float32, np.float64, np.int32, np.int64, np.bool_  # noqa: A001

VARIABLE_STORE: dict = {}        # name -> np.ndarray, filled by the golden generator
REQUESTED_VARIABLES: dict = {}   # name -> shape, recorded as the reference asks for them
UNIFORM_FEED: list = []          # arrays handed out, in creation order, to tf.random_uniform nodes
_ids = itertools.count()
_scope_stack: list = []
_uniform_nodes = itertools.count()


def reset_shim():
    global _ids, _uniform_nodes
    VARIABLE_STORE.clear()
    REQUESTED_VARIABLES.clear()
    UNIFORM_FEED.clear()
    _scope_stack.clear()
    _ids = itertools.count()
    _uniform_nodes = itertools.count()


# --------------------------------------------------------------------------- #
# deferred tensor
# --------------------------------------------------------------------------- #
class Tensor:
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), static_shape=None, kind="op", name=None):
        self.id = next(_ids)
        self.fn, self.inputs, self.static_shape, self.kind, self.name = fn, tuple(inputs), static_shape, kind, name

    # arithmetic ------------------------------------------------------------
    def __add__(self, o): return _op(np.add, self, o)
    def __radd__(self, o): return _op(np.add, o, self)
    def __sub__(self, o): return _op(np.subtract, self, o)
    def __rsub__(self, o): return _op(np.subtract, o, self)
    def __mul__(self, o): return _op(np.multiply, self, o)
    def __rmul__(self, o): return _op(np.multiply, o, self)
    def __truediv__(self, o): return _op(np.divide, self, o)
    def __floordiv__(self, o): return _op(np.floor_divide, self, o)
    def __mod__(self, o): return _op(np.mod, self, o)
    def __neg__(self): return _op(np.negative, self)

    def __getitem__(self, key):
        key = key if isinstance(key, tuple) else (key,)
        dyn = [k for k in key if isinstance(k, Tensor)]
        def f(x, *d):
            it = iter(d)
            k2 = tuple(int(next(it)) if isinstance(k, Tensor) else k for k in key)
            return x[k2]
        return Tensor(f, [self] + dyn)

    def __iter__(self):
        raise TypeError("shim Tensor is not iterable")

    def get_shape(self):
        return self.static_shape


def _wrap(x):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (list, tuple)) and x and all(isinstance(v, Tensor) for v in x):
        # TF's constant conversion packs a python list of tensors along a new axis 0 (ops.convert_to_tensor ->
        # array_ops.stack): reward_func subtracts the LIST sample_solution from a stacked tensor (vrptw_utils.py:394)
        return Tensor(lambda *vs: np.stack(vs, 0), list(x))
    return Tensor(lambda: x, [], kind="const")


def _match(a, b):
    """Python scalars adopt the dtype of the tensor operand (TF's constant conversion)."""
    if isinstance(a, np.ndarray) and not isinstance(b, np.ndarray) and not isinstance(b, np.generic):
        return a, np.asarray(b).astype(a.dtype)
    if isinstance(b, np.ndarray) and not isinstance(a, np.ndarray) and not isinstance(a, np.generic):
        return np.asarray(a).astype(b.dtype), b
    return a, b


def _op(npf, *args):
    def f(*vals):
        if len(vals) == 2:
            vals = _match(*vals)
        out = npf(*vals)
        return out
    return Tensor(f, [_wrap(a) for a in args])


def _lazy(fn, *tensor_args):
    return Tensor(fn, [_wrap(a) for a in tensor_args])


def _shape_arg(shape):
    """A shape given as python list possibly holding Tensors -> (Tensor inputs, builder, static)."""
    if isinstance(shape, Tensor):
        return [shape], (lambda s: tuple(int(v) for v in np.asarray(s).reshape(-1))), None
    shape = list(shape)
    dyn = [s for s in shape if isinstance(s, Tensor)]
    static = [None if isinstance(s, Tensor) else int(s) for s in shape]
    def build(*d):
        it = iter(d)
        return tuple(int(next(it)) if isinstance(s, Tensor) else int(s) for s in shape)
    return dyn, build, static


# --------------------------------------------------------------------------- #
# evaluation
# --------------------------------------------------------------------------- #
def _evaluate(targets, memo):
    need, stack = {}, list(targets)
    while stack:
        t = stack.pop()
        if t.id in need or t.id in memo:
            continue
        need[t.id] = t
        stack.extend(t.inputs)
    for tid in sorted(need):
        t = need[tid]
        if t.kind == "placeholder":
            raise KeyError(f"placeholder {t.name} was not fed")
        if t.kind == "cond":
            memo[tid] = t.fn(memo, *[memo[i.id] for i in t.inputs])
        else:
            memo[tid] = t.fn(*[memo[i.id] for i in t.inputs])


def _flatten(x, out):
    if isinstance(x, Tensor):
        out.append(x)
    elif isinstance(x, (list, tuple)):
        for y in x:
            _flatten(y, out)
    return out


def _rebuild(x, memo):
    if isinstance(x, Tensor):
        return memo[x.id]
    if isinstance(x, (list, tuple)):
        return type(x)(_rebuild(y, memo) for y in x) if not hasattr(x, "_fields") else type(x)(*[_rebuild(y, memo) for y in x])
    return x


class Session:
    def __init__(self, *a, **k):
        pass

    def run(self, fetches, feed_dict=None):
        memo = {}
        for ph, val in (feed_dict or {}).items():
            v = np.asarray(val)
            if ph.dtype is not None:
                v = v.astype(ph.dtype)
            memo[ph.id] = v
        _evaluate(_flatten(fetches, []), memo)
        return _rebuild(fetches, memo)

    def close(self):
        pass


InteractiveSession = Session


def placeholder(dtype, shape=None, name=None):
    t = Tensor(None, [], static_shape=list(shape) if shape is not None else None, kind="placeholder", name=name)
    t.dtype = dtype
    return t


# --------------------------------------------------------------------------- #
# variables / scopes
# --------------------------------------------------------------------------- #
@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scope_stack.append(name if isinstance(name, str) else name.name)
    try:
        yield types.SimpleNamespace(name="/".join(_scope_stack))
    finally:
        _scope_stack.pop()


def _scoped(name):
    return "/".join([s.rstrip("/") for s in _scope_stack if s] + [name])


def get_variable(name, shape=None, initializer=None, **k):
    full = _scoped(name)
    return Tensor(lambda: _value(full, shape), [], static_shape=list(shape), kind="variable", name=full)


def get_collection(*a, **k):
    return []


def global_variables_initializer():
    return Tensor(lambda: None, [])


def set_random_seed(*a, **k):
    pass


GraphKeys = types.SimpleNamespace(TRAINABLE_VARIABLES="trainable_variables", GLOBAL_VARIABLES="variables")


# --------------------------------------------------------------------------- #
# ops
# --------------------------------------------------------------------------- #
def shape(x, *a, **k):
    return _lazy(lambda v: np.asarray(np.shape(v), np.int32), x)


def cast(x, dtype, *a, **k):
    return _lazy(lambda v: np.asarray(v).astype(dtype), x)


def expand_dims(x, axis, *a, **k):
    return _lazy(lambda v: np.expand_dims(v, axis), x)


def squeeze(x, axis=None, *a, **k):
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return _lazy(lambda v: np.squeeze(v, ax), x)


def tile(x, multiples, *a, **k):
    dyn, build, _ = _shape_arg(multiples)
    return Tensor(lambda v, *d: np.tile(v, build(*d)), [_wrap(x)] + dyn)


def concat(values, axis, *a, **k):
    return Tensor(lambda *v: np.concatenate(v, axis), [_wrap(v) for v in values])


def stack(values, axis=0, *a, **k):
    return Tensor(lambda *v: np.stack(v, axis), [_wrap(v) for v in values])


def unstack(x, num=None, axis=0, *a, **k):
    n = num if num is not None else x.static_shape[axis]
    return [_lazy((lambda i: (lambda v: np.take(v, i, axis=axis)))(i), x) for i in builtins_range(n)]


def split(x, num_or_size_splits, axis=0, *a, **k):
    n = int(num_or_size_splits)
    return [_lazy((lambda i: (lambda v: np.split(v, n, axis=axis)[i]))(i), x) for i in builtins_range(n)]


def reshape(x, shp, *a, **k):
    dyn, build, _ = _shape_arg(shp)
    return Tensor(lambda v, *d: np.reshape(v, build(*d)), [_wrap(x)] + dyn)


def transpose(x, perm=None, *a, **k):
    return _lazy(lambda v: np.transpose(v, perm), x)


def zeros(shp, dtype=float32, *a, **k):
    dyn, build, static = _shape_arg(shp)
    return Tensor(lambda *d: np.zeros(build(*d), dtype), dyn, static_shape=static)


def ones(shp, dtype=float32, *a, **k):
    dyn, build, static = _shape_arg(shp)
    return Tensor(lambda *d: np.ones(build(*d), dtype), dyn, static_shape=static)


def zeros_like(x, *a, **k):
    return _lazy(np.zeros_like, x)


def ones_like(x, *a, **k):
    return _lazy(np.ones_like, x)


def constant(v, dtype=None, *a, **k):
    arr = np.asarray(v, dtype) if dtype is not None else (
        np.asarray(v, np.int32) if isinstance(v, int) else np.asarray(v, np.float32) if isinstance(v, float) else np.asarray(v))
    return Tensor(lambda: arr, [], kind="const")


def range(*args, **k):  # noqa: A001
    ts = [_wrap(a) for a in args]
    return Tensor(lambda *v: np.arange(*[int(x) for x in v], dtype=np.int32), ts)


def equal(a, b, *x, **k): return _op(np.equal, a, b)
def greater(a, b, *x, **k): return _op(np.greater, a, b)
def greater_equal(a, b, *x, **k): return _op(np.greater_equal, a, b)
def less(a, b, *x, **k): return _op(np.less, a, b)
def add(a, b, *x, **k): return _op(np.add, a, b)
def subtract(a, b, *x, **k): return _op(np.subtract, a, b)
def multiply(a, b, *x, **k): return _op(np.multiply, a, b)
def divide(a, b, *x, **k): return _op(np.divide, a, b)
def scalar_mul(s, x, *a, **k): return _op(np.multiply, s, x)
def minimum(a, b, *x, **k): return _op(np.minimum, a, b)
def maximum(a, b, *x, **k): return _op(np.maximum, a, b)
def square(x, *a, **k): return _op(np.square, x)
def sqrt(x, *a, **k): return _op(np.sqrt, x)
def exp(x, *a, **k): return _op(np.exp, x)
def cos(x, *a, **k): return _op(np.cos, x)


def assert_equal(a, b, *x, **k):
    """Graph-mode assert: an op nobody runs or depends on (UPS/vrptw_ups_utils.py:390 drops its result) - a no-op."""
    return None


def stop_gradient(x, *a, **k): return x
def Print(x, *a, **k): return x


def log(x, *a, **k):
    def f(v):
        with np.errstate(divide="ignore"):
            return np.log(v)
    return _lazy(f, x)


def pow(x, y, *a, **k):  # noqa: A001
    return _lazy(lambda v: np.power(v, np.asarray(y).astype(v.dtype)), x)


def reduce_sum(x, axis=None, keepdims=False, *a, **k):
    return _lazy(lambda v: np.sum(v, axis=axis, keepdims=keepdims, dtype=v.dtype), x)


def reduce_mean(x, axis=None, *a, **k):
    return _lazy(lambda v: np.mean(v, axis=axis, dtype=v.dtype), x)


def add_n(xs, *a, **k):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def cumsum(x, axis=0, *a, **k):
    return _lazy(lambda v: np.cumsum(v, axis, dtype=v.dtype), x)


def argmax(x, axis=None, *a, **k):
    return _lazy(lambda v: np.argmax(v, axis).astype(np.int64), x)


def argmin(x, axis=None, *a, **k):
    return _lazy(lambda v: np.argmin(v, axis).astype(np.int64), x)


def matmul(a, b, *x, **k):
    return _lazy(lambda u, v: np.matmul(u, v).astype(u.dtype, copy=False), a, b)


def gather_nd(params, indices, *a, **k):
    def f(p, ind):
        ind = np.asarray(ind)
        return p[tuple(ind[..., i] for i in np.arange(ind.shape[-1]))]
    return _lazy(f, params, indices)


def scatter_nd(indices, updates, shp, *a, **k):
    def f(ind, upd, s):
        out = np.zeros(tuple(int(v) for v in s), upd.dtype)
        np.add.at(out, tuple(ind[..., i] for i in np.arange(ind.shape[-1])), upd)
        return out
    return _lazy(f, indices, updates, shp)


def random_uniform(shp, *a, **k):
    slot = next(_uniform_nodes)
    dyn, build, _ = _shape_arg(shp)
    def f(*d):
        want = build(*d)
        val = np.asarray(UNIFORM_FEED[slot], np.float32).reshape(want)
        return val
    return Tensor(f, dyn)


def cond(pred, true_fn, false_fn, *a, **k):
    t_out, f_out = true_fn(), false_fn()
    box = Tensor(None, [_wrap(pred)], kind="cond")
    def run(memo, p):
        chosen = t_out if np.asarray(p).item() else f_out
        _evaluate(_flatten(chosen, []), memo)
        return _rebuild(chosen, memo)
    box.fn = run
    n = len(t_out)
    return tuple(_lazy((lambda i: (lambda v: v[i]))(i), box) for i in builtins_range(n))




# --------------------------------------------------------------------------- #
# tf.nn
# --------------------------------------------------------------------------- #
def _log_softmax(v):
    m = v.max(axis=-1, keepdims=True)
    sh = v - m
    return sh - np.log(np.exp(sh).sum(axis=-1, keepdims=True, dtype=v.dtype))


def _top_k(v, kk):
    order = np.argsort(-v, axis=-1, kind="stable")[..., :kk]
    return np.take_along_axis(v, order, -1), order.astype(np.int32)


def _nn_top_k(x, k=1, *a, **kw):
    box = _lazy(lambda v: _top_k(v, k), x)
    return _lazy(lambda t: t[0], box), _lazy(lambda t: t[1], box)


class LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return tuple.__new__(cls, (c, h))
    c = property(lambda s: s[0])
    h = property(lambda s: s[1])


def _value(full_name, shape):
    """Resolve a variable at evaluation time (its shape is only known once inputs are)."""
    REQUESTED_VARIABLES[full_name] = tuple(int(s) for s in shape)
    if full_name not in VARIABLE_STORE:
        raise KeyError(f"reference asked for variable '{full_name}' {tuple(shape)} which the oracle does not define")
    val = np.asarray(VARIABLE_STORE[full_name])
    if tuple(val.shape) != tuple(int(s) for s in shape):
        raise ValueError(f"variable '{full_name}': reference shape {tuple(shape)} != oracle shape {val.shape}")
    return val


class BasicLSTMCell:
    """TF-1 rnn_cell_impl.BasicLSTMCell: gates i, j, f, o; c' = c*sig(f+fb) + sig(i)*tanh(j)."""

    def __init__(self, num_units, forget_bias=1.0, **k):
        self.num_units, self.forget_bias = num_units, forget_bias

    def __call__(self, inputs, state):
        scope = _scoped("basic_lstm_cell")
        c, h = state
        fb, nu = self.forget_bias, self.num_units

        def gates(x, hh):
            kk = _value(scope + "/kernel", [x.shape[1] + nu, 4 * nu])
            bb = _value(scope + "/bias", [4 * nu])
            return (np.concatenate([x, hh], 1) @ kk + bb).astype(x.dtype, copy=False)
        g = _lazy(gates, inputs, h)

        def new_c(gv, cv):
            i, j, f, o = np.split(gv, 4, axis=1)
            one = gv.dtype.type(1)
            sig = lambda z: one / (one + np.exp(-z))
            return cv * sig(f + gv.dtype.type(fb)) + sig(i) * np.tanh(j)
        nc = _lazy(new_c, g, c)

        def new_h(gv, ncv):
            o = np.split(gv, 4, axis=1)[3]
            one = gv.dtype.type(1)
            return np.tanh(ncv) * (one / (one + np.exp(-o)))
        nh = _lazy(new_h, g, nc)
        return nh, LSTMStateTuple(nc, nh)


class DropoutWrapper:
    """input_keep_prob is fed 1.0 on the inference path: x / 1 * floor(1 + u) == x exactly."""

    def __init__(self, cell, input_keep_prob=1.0, **k):
        self.cell = cell

    def __call__(self, inputs, state):
        return self.cell(inputs, state)


class MultiRNNCell:
    def __init__(self, cells, **k):
        self.cells = cells

    def __call__(self, inputs, state):
        new_states, cur = [], inputs
        with variable_scope("multi_rnn_cell"):
            for i, cell in enumerate(self.cells):
                with variable_scope(f"cell_{i}"):
                    cur, ns = cell(cur, state[i])
                new_states.append(ns)
        return cur, tuple(new_states)


def _dynamic_rnn(cell, inputs, initial_state=None, scope=None, **k):
    """Only the [batch, 1, dim] single-time-step use in shared/decode_step.py:211-214."""
    x = inputs[:, 0, :]
    with variable_scope(scope if scope else "rnn"):
        out, state = cell(x, initial_state)
    return expand_dims(out, 1), state


class _Layer:
    """tf.layers.Layer: `name=` is relative to the variable scope at construction.  A string `_scope=` is entered with
    tf.variable_scope(self._scope) when the layer is first CALLED (tf/python/layers/base.py, __call__), i.e. relative to the
    variable scope that is current at that call: top level for the actor's layers ("Actor/Decoder/Attention/emb_d"), inside
    "Critic/Process" for the critic's ("Critic/Process/P0/emb_d")."""

    def __init__(self, units, activation=None, _scope=None, name=None, **k):
        self.units, self.activation = units, activation
        self._raw_scope = _scope.rstrip("/") if _scope is not None else None
        self._scope_name = None if _scope is not None else _scoped(name)

    @property
    def scope(self):
        if self._scope_name is None:       # first call: resolve against the current variable scope
            self._scope_name = _scoped(self._raw_scope)
        return self._scope_name


class Conv1D(_Layer):
    def __init__(self, filters, kernel_size, *a, **k):
        assert kernel_size == 1
        super().__init__(filters, **k)

    def __call__(self, x):
        sc, un = self.scope, self.units
        def f(v):
            kk = _value(sc + "/kernel", [1, v.shape[-1], un])
            bb = _value(sc + "/bias", [un])
            return (v @ kk[0] + bb).astype(v.dtype, copy=False)
        return _lazy(f, x)


class Dense(_Layer):
    def __call__(self, x):
        sc, un = self.scope, self.units
        def f(v):
            kk = _value(sc + "/kernel", [v.shape[-1], un])
            bb = _value(sc + "/bias", [un])
            return (v @ kk + bb).astype(v.dtype, copy=False)
        y = _lazy(f, x)
        return self.activation(y) if self.activation else y


def _dense_fn(x, units, activation=None, name=None, **k):
    return Dense(units, activation=activation, name=name)(x)


nn = types.SimpleNamespace(
    tanh=lambda x, *a, **k: _op(np.tanh, x),
    relu=lambda x, *a, **k: _op(lambda v: np.maximum(v, 0), x),
    softmax=lambda x, *a, **k: _lazy(lambda v: np.exp(_log_softmax(v)), x),
    log_softmax=lambda x, *a, **k: _lazy(_log_softmax, x),
    top_k=_nn_top_k,
    dynamic_rnn=_dynamic_rnn,
    rnn_cell=types.SimpleNamespace(BasicLSTMCell=BasicLSTMCell, MultiRNNCell=MultiRNNCell,
                                   LSTMStateTuple=LSTMStateTuple),
)
layers = types.SimpleNamespace(Conv1D=Conv1D, Dense=Dense, dense=_dense_fn)
contrib = types.SimpleNamespace(
    layers=types.SimpleNamespace(xavier_initializer=lambda *a, **k: "glorot_uniform"),
    rnn=types.SimpleNamespace(DropoutWrapper=DropoutWrapper),
)
train = types.SimpleNamespace(
    Saver=lambda *a, **k: types.SimpleNamespace(save=lambda *a, **k: None, restore=lambda *a, **k: None),
    AdamOptimizer=None, latest_checkpoint=lambda *a, **k: None)
losses = types.SimpleNamespace()
