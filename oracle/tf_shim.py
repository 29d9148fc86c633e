"""Minimal eager emulation of the TF-1.14 / TFP-0.7 API surface that the reference's hot-path
source files touch, on top of torch-CPU.  TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow 1.14 cannot be installed here (Python 3.12, no network), so the reference
cannot run as shipped.  With this shim registered as ``tensorflow`` / ``tensorflow_probability``
the UNMODIFIED files ``/root/reference/networks/{ops,utils}.py`` and
``networks/actor_critic/{actor_critic,a2c,ppo,sac}.py`` import and their graph-building methods
execute eagerly (``oracle/gen_golden.py``).  What is pinned that way is the reference's own
composition of ops (op order, custom gradients, index plumbing of the resampler); what is NOT
pinned is the inside of the third-party kernels, which this file restates:
``Normal.prob`` and ``Softmax`` from the lowered graph of the shipped ``.meta``,
``Multinomial`` (CPU functor) and ``RelaxedOneHotCategorical.sample`` from their published
algorithm.  Random ops pop caller-injected draws from ``DRAWS`` so that results are reproducible
and comparable with the CUDA kernels' verification mode.
"""
from __future__ import annotations

import contextlib
import sys
import types
from collections import namedtuple

import numpy as np
import torch

DTYPE = torch.float64  # arithmetic dtype of float tensors created by the shim
DRAWS: dict = {}       # name -> list of arrays, popped in call order
TRACE: dict = {}       # intermediate integer results recorded for the golden files


def set_dtype(dt):
    global DTYPE
    DTYPE = dt


def _pop(name):
    q = DRAWS.get(name)
    if not q:
        raise RuntimeError(f"tf_shim: no injected draw left for '{name}'")
    return q.pop(0)


class Dim(int):
    @property
    def value(self):
        return int(self)


class Shape(tuple):
    def __new__(cls, dims):
        return super().__new__(cls, [Dim(d) for d in dims])

    @property
    def ndims(self):
        return len(self)

    def num_elements(self):
        return int(np.prod(self)) if len(self) else 1

    def as_list(self):
        return [int(d) for d in self]

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return Shape(r) if isinstance(i, slice) else r


class DType:
    def __init__(self, name, tdt):
        self.name, self.t = name, tdt

    def __repr__(self):
        return f"tf.{self.name}"


float32 = DType("float32", torch.float32)
float64 = DType("float64", torch.float64)
int32 = DType("int32", torch.int32)
int64 = DType("int64", torch.int64)
bool_ = DType("bool", torch.bool)


def _tdt(dtype):
    if dtype is None:
        return DTYPE
    if isinstance(dtype, DType):
        return DTYPE if dtype.t in (torch.float32, torch.float64) else dtype.t
    if dtype in (np.float32, np.float64, float):
        return DTYPE
    return dtype


class T:
    """Eager tensor.  Deliberately has no __len__ (the reference tests hasattr(x, '__len__'))."""
    __array_priority__ = 1000

    def __init__(self, t, name=None):
        self.t = t
        self.name = name or "Tensor:0"

    @property
    def shape(self):
        return Shape(self.t.shape)

    @property
    def dtype(self):
        m = {torch.float32: float32, torch.float64: float32, torch.int32: int32, torch.int64: int64, torch.bool: bool_}
        return m[self.t.dtype]

    def read_value(self):
        return self

    def numpy(self):
        return self.t.detach().numpy()

    def _b(self, o, fn):
        return T(fn(self.t, _raw(o, like=self.t)))

    def __add__(self, o): return self._b(o, torch.add)
    def __radd__(self, o): return T(torch.add(_raw(o, like=self.t), self.t))
    def __sub__(self, o): return self._b(o, torch.sub)
    def __rsub__(self, o): return T(torch.sub(_raw(o, like=self.t), self.t))
    def __mul__(self, o): return self._b(o, torch.mul)
    def __rmul__(self, o): return T(torch.mul(_raw(o, like=self.t), self.t))
    def __truediv__(self, o): return self._b(o, torch.div)
    def __rtruediv__(self, o): return T(torch.div(_raw(o, like=self.t), self.t))
    def __pow__(self, o): return T(torch.pow(self.t, o))
    def __neg__(self): return T(-self.t)
    def __lt__(self, o): return self._b(o, torch.lt)
    def __gt__(self, o): return self._b(o, torch.gt)
    def __le__(self, o): return self._b(o, torch.le)
    def __ge__(self, o): return self._b(o, torch.ge)
    def __getitem__(self, i): return T(self.t[i])
    def __hash__(self): return id(self)
    def __eq__(self, o): return self is o


def _raw(x, like=None):
    if isinstance(x, T):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, T) for e in x):
        return torch.stack([_raw(e) for e in x])
    a = np.asarray(x)
    if a.dtype.kind == "f":
        dt = like.dtype if (like is not None and like.dtype.is_floating_point) else DTYPE
        return torch.as_tensor(a, dtype=dt)
    if a.dtype.kind in "iu" and like is not None and like.dtype.is_floating_point:
        return torch.as_tensor(a, dtype=like.dtype)
    return torch.as_tensor(a)


def _ax(axis):
    return axis


# ---------------------------------------------------------------------------------- core ops ----
def constant(v, dtype=None, name=None, shape=None):
    return T(torch.as_tensor(np.asarray(v), dtype=_tdt(dtype)))


def convert(x):
    return x if isinstance(x, T) else T(_raw(x))


def exp(x, name=None): return T(torch.exp(_raw(x)))
def log(x, name=None): return T(torch.log(_raw(x)))
def sqrt(x, name=None): return T(torch.sqrt(_raw(x)))
def square(x, name=None): return T(torch.square(_raw(x)))
def tanh(x, name=None): return T(torch.tanh(_raw(x)))
def atanh(x, name=None): return T(torch.atanh(_raw(x)))
def softplus(x, name=None): return T(torch.nn.functional.softplus(_raw(x)))
def relu6(x, name=None): return T(torch.clamp(_raw(x), 0, 6))
def relu(x, name=None): return T(torch.relu(_raw(x)))
def is_nan(x): return T(torch.isnan(_raw(x)))
def is_inf(x): return T(torch.isinf(_raw(x)))
def logical_or(a, b): return T(torch.logical_or(_raw(a), _raw(b)))
def zeros_like(x, dtype=None): return T(torch.zeros_like(_raw(x)))
def ones_like(x, dtype=None): return T(torch.ones_like(_raw(x)))
def stop_gradient(x): return T(_raw(x).detach())
def identity(x, name=None): return T(_raw(x))
def maximum(a, b, name=None): return T(torch.maximum(*_pair(a, b)))
def minimum(a, b, name=None): return T(torch.minimum(*_pair(a, b)))
def multiply(a, b, name=None): return convert(a) * b
def add(a, b, name=None): return convert(a) + b
def equal(a, b): return T(torch.eq(*_pair(a, b)))
def matmul(a, b): return T(_raw(a) @ _raw(b))
def squeeze(x, axis=None): return T(_raw(x).squeeze(axis) if axis is not None else _raw(x).squeeze())
def expand_dims(x, axis): return T(_raw(x).unsqueeze(axis))
def transpose(x, perm): return T(_raw(x).permute(*perm))
def argmax(x, axis=None): return T(torch.argmax(_raw(x), dim=axis))
def one_hot(idx, depth, dtype=None): return T(torch.nn.functional.one_hot(_raw(idx).long(), int(depth)).to(_tdt(dtype)))
def cast(x, dtype): return T(_raw(x).to(_tdt(dtype)))
def clip_by_value(x, lo, hi): return T(torch.minimum(torch.maximum(_raw(x), _raw(lo, like=_raw(x))), _raw(hi, like=_raw(x))))
def add_n(xs): return T(sum(_raw(x) for x in xs))


def _pair(a, b):
    ra = _raw(a) if isinstance(a, (T, torch.Tensor)) else None
    rb = _raw(b) if isinstance(b, (T, torch.Tensor)) else None
    like = ra if ra is not None else rb
    return (ra if ra is not None else _raw(a, like=like)), (rb if rb is not None else _raw(b, like=like))


def _reduce(fn):
    def f(x, axis=None, keepdims=False, name=None, keep_dims=None):
        r = _raw(x)
        kd = bool(keepdims or keep_dims)
        if axis is None:
            return T(fn(r))
        return T(fn(r, dim=axis, keepdim=kd))
    return f


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)
reduce_max = _reduce(torch.amax)


def softmax(x, axis=-1):
    r = _raw(x)
    e = torch.exp(r - torch.amax(r, dim=axis, keepdim=True))  # TF Softmax kernel
    return T(e / torch.sum(e, dim=axis, keepdim=True))


def log_softmax(x, axis=-1):
    r = _raw(x)
    s = r - torch.amax(r, dim=axis, keepdim=True)
    return T(s - torch.log(torch.sum(torch.exp(s), dim=axis, keepdim=True)))


def moments(x, axes):
    r = _raw(x)
    m = torch.mean(r, dim=axes)
    return T(m), T(torch.mean(torch.square(r - torch.mean(r, dim=axes, keepdim=True)), dim=axes))


def reshape(x, shape):
    shp = [int(_raw(s)) if isinstance(s, T) else int(s) for s in shape]
    return T(_raw(x).reshape(shp))


def shape(x, out_type=None):
    return T(torch.tensor(list(_raw(x).shape), dtype=_tdt(out_type) if out_type else torch.int32))


def where(cond, x=None, y=None):
    c = _raw(cond)
    if x is None:
        out = T(torch.nonzero(c))  # row-major coordinates, int64
        TRACE["where"] = out.t.clone()
        return out
    return T(torch.where(c, *_pair(x, y)))


def range_(*a, dtype=None, **kw):
    vals = [int(_raw(v)) if isinstance(v, T) else int(v) for v in a]
    return T(torch.arange(*vals, dtype=_tdt(dtype) if dtype else torch.int32))


def stack(xs, axis=0): return T(torch.stack([_raw(x) for x in xs], dim=axis))
def concat(xs, axis=0): return T(torch.cat([_raw(x) for x in xs], dim=axis))
def split(x, n, axis=0):
    return [T(p) for p in torch.split(_raw(x), n if isinstance(n, list) else _raw(x).shape[axis] // n, dim=axis)]


def meshgrid(a, b, indexing="xy"):
    ga, gb = torch.meshgrid(_raw(a), _raw(b), indexing=indexing)
    return [T(ga), T(gb)]


def gather(params, indices, batch_dims=0, axis=None, name=None):
    p, i = _raw(params), _raw(indices).long()
    if batch_dims in (-1,) or (batch_dims and batch_dims == i.dim() - 1 and batch_dims > 0):
        return T(torch.gather(p, p.dim() - 1, i))  # rank-2 params/indices: out[a, s] = p[a, i[a, s]]
    if axis in (-1, p.dim() - 1) and p.dim() > 1:
        return T(p.index_select(p.dim() - 1, i.reshape(-1)).reshape(*p.shape[:-1], *i.shape))
    return T(p[i])


def gather_nd(params, indices):
    i = _raw(indices).long()
    return T(_raw(params)[tuple(i[..., k] for k in range(i.shape[-1]))])


def batch_gather(params, indices):
    return T(torch.gather(_raw(params), 1, _raw(indices).long()))


UWC = namedtuple("UniqueWithCounts", ["y", "idx", "count"])


def unique_with_counts(x):
    v = _raw(x)
    uniq, idx, seen = [], [], {}
    for e in v.tolist():
        if e not in seen:
            seen[e] = len(uniq)
            uniq.append(e)
        idx.append(seen[e])
    cnt = [0] * len(uniq)
    for i in idx:
        cnt[i] += 1
    out = UWC(T(torch.tensor(uniq, dtype=v.dtype)), T(torch.tensor(idx, dtype=torch.int32)),
              T(torch.tensor(cnt, dtype=torch.int32)))
    TRACE["unique_in"] = v.clone()
    TRACE["unique_out"] = tuple(o.t.clone() for o in out)
    return out


def map_fn(fn, elems, dtype=None):
    e = _raw(elems)
    res = [_raw(fn(T(e[i]))) for i in range(e.shape[0])]
    out = torch.stack(res) if res else torch.zeros(0, dtype=_tdt(dtype))
    TRACE["map_fn"] = out.clone()
    return T(out)


def scatter_nd_update(ref, indices, updates):
    i = _raw(indices).long()
    with torch.no_grad():
        ref.t[tuple(i[..., k] for k in range(i.shape[-1]))] = _raw(updates).to(ref.t.dtype)
    return ref


def scatter_update(ref, indices, updates):
    with torch.no_grad():
        ref.t[_raw(indices).long()] = _raw(updates).to(ref.t.dtype)
    return ref


def assign(ref, value):
    with torch.no_grad():
        ref.t.copy_(_raw(value))
    return ref


def top_k(x, k=1, sorted=True):
    r = _raw(x)
    order = torch.argsort(-r, dim=-1, stable=True)[..., :k]
    TRACE["top_k"] = order.clone()
    return T(torch.gather(r, -1, order)), T(order.to(torch.int32))


def custom_gradient(f):
    def wrapper(*args):
        holder = {}

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *ts):
                y, g = f(*[T(t.detach()) for t in ts])
                holder["g"] = g
                return _raw(y).clone()

            @staticmethod
            def backward(ctx, dy):
                out = holder["g"](T(dy))
                out = out if isinstance(out, (tuple, list)) else (out,)
                return tuple(_raw(o) for o in out)

        return T(Fn.apply(*[_raw(a) for a in args]))
    return wrapper


# ------------------------------------------------------------------- third-party kernels --------
def multinomial_cpu(logits, num_samples, uniforms):
    """TF-1.14 Multinomial CPU functor: fp64 running CDF of exp(double(logit) - max) over the
    finite logits + upper_bound(cdf, u * total)."""
    lg = np.asarray(logits.detach().to(torch.float32).numpy(), dtype=np.float32)
    u = np.asarray(uniforms, dtype=np.float64).reshape(lg.shape[0], num_samples)
    out = np.empty((lg.shape[0], num_samples), dtype=np.int64)
    for r in range(lg.shape[0]):
        fin = np.isfinite(lg[r])
        mx = np.float64(lg[r][fin].max())
        run, cdf = 0.0, np.empty(lg.shape[1])
        for j in range(lg.shape[1]):
            if fin[j]:
                run += np.exp(np.float64(lg[r, j]) - mx)
            cdf[j] = run
        out[r] = np.minimum(np.searchsorted(cdf, u[r] * run, side="right"), lg.shape[1] - 1)
    return torch.from_numpy(out)


class Normal:
    def __init__(self, loc, scale, allow_nan_stats=True, **kw):
        self.loc, self.scale = convert(loc), convert(scale)

    def prob(self, x, name=None):  # [graph] Normal/prob_1: Sub RealDiv Square Mul Log Add Sub Exp
        z = (convert(x) - self.loc) / self.scale
        return exp(-0.5 * square(z) - (0.9189385175704956 + log(self.scale)))

    def log_prob(self, x, name=None):
        z = (convert(x) - self.loc) / self.scale
        return -0.5 * square(z) - (0.9189385175704956 + log(self.scale))

    def sample(self, n, seed=None):
        n = int(_raw(n)) if isinstance(n, T) else int(n)
        eps = torch.as_tensor(np.asarray(_pop("normal")), dtype=self.loc.t.dtype).reshape(n, *self.loc.t.shape)
        return T(eps * self.scale.t + self.loc.t)  # RandomStandardNormal * scale + loc

    def mean(self):
        return self.loc


class Categorical:
    def __init__(self, logits=None, probs=None, **kw):
        self.logits = convert(logits)
        self.probs = softmax(self.logits)

    def sample(self, n, seed=None):
        lg = self.logits.t
        flat = lg.reshape(-1, lg.shape[-1])
        idx = multinomial_cpu(flat, int(n), _pop("cat_uniform"))  # [rows, n]
        return T(idx.t().reshape(int(n), *lg.shape[:-1]).to(torch.int32))


class _ExpRelaxed:
    def __init__(self, logits):
        self.logits = convert(logits)
        self.probs = softmax(self.logits)


class RelaxedOneHotCategorical:
    """TFP 0.7: Exp bijector over ExpRelaxedOneHotCategorical._sample_n."""

    def __init__(self, temperature, logits=None, **kw):
        self.temperature = float(temperature)
        self.distribution = _ExpRelaxed(logits)

    def sample(self, n, seed=None):
        lg = self.distribution.logits.t
        u = torch.as_tensor(np.asarray(_pop("gumbel_uniform")), dtype=lg.dtype).reshape(int(n), *lg.shape)
        g = -torch.log(-torch.log(u))
        noisy = (g + lg) / self.temperature
        return T(torch.exp(_raw(log_softmax(T(noisy)))))


def random_categorical(logits, num_samples, dtype=None, seed=None):
    out = multinomial_cpu(_raw(logits), int(num_samples), _pop("resample_cat_uniform")).to(_tdt(dtype) if dtype else torch.int64)
    TRACE["categorical"] = out.clone()
    return T(out)


def random_uniform(shape_, minval=0, maxval=None, dtype=None, seed=None):
    if isinstance(dtype, DType) and dtype.t in (torch.int32, torch.int64):
        n = int(_raw(shape_[0])) if isinstance(shape_, (list, tuple)) else int(_raw(shape_)[0])
        return T(torch.as_tensor(np.asarray(_pop("int_uniform"))[:n], dtype=dtype.t))
    n = int(_raw(shape_)[0]) if isinstance(shape_, T) else int(np.prod([int(s) for s in shape_]))
    return T(torch.as_tensor(np.asarray(_pop("float_uniform"))[:n], dtype=DTYPE))


# ------------------------------------------------------------------------------ variables --------
VARIABLES: list = []
_scopes: list = []


class _Init:
    def __init__(self, fn): self.fn = fn
    def __call__(self, shape=None, dtype=None): return self.fn(shape)


def constant_initializer(value=0.0, dtype=None):
    def mk(shape):
        a = np.asarray(value, dtype=np.float64)
        return torch.as_tensor(np.broadcast_to(a, shape).copy(), dtype=DTYPE)
    return _Init(mk)


def zeros_initializer(): return constant_initializer(0.0)
def ones_initializer(): return constant_initializer(1.0)


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    return _Init(lambda shape: torch.as_tensor(np.asarray(_pop("weight_init")).reshape(shape), dtype=DTYPE))


glorot_uniform_initializer = truncated_normal_initializer
orthogonal_initializer = truncated_normal_initializer
variance_scaling_initializer = truncated_normal_initializer


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    full = "/".join(_scopes + [name]) + ":0"
    shp = [int(s) for s in shape] if shape is not None else []
    v = T(initializer(shp).clone().requires_grad_(bool(trainable)), name=full)
    v.trainable = bool(trainable)
    VARIABLES.append(v)
    return v


def Variable(initial_value, dtype=None, trainable=True, name="Variable", shape=None):
    v = T(torch.as_tensor(np.asarray(initial_value), dtype=_tdt(dtype)).clone().requires_grad_(bool(trainable)),
          name="/".join(_scopes + [name]) + ":0")
    v.trainable = bool(trainable)
    VARIABLES.append(v)
    return v


@contextlib.contextmanager
def variable_scope(name, reuse=None, **kw):
    _scopes.append(name if isinstance(name, str) else "scope")
    try:
        yield
    finally:
        _scopes.pop()


@contextlib.contextmanager
def _noop_ctx(*a, **k):
    yield


class _GraphKeys:
    TRAINABLE_VARIABLES = "trainable"
    GLOBAL_VARIABLES = "global"
    UPDATE_OPS = "update_ops"


def get_collection(key, scope=None):
    if key == _GraphKeys.TRAINABLE_VARIABLES:
        return [v for v in VARIABLES if getattr(v, "trainable", False)]
    return list(VARIABLES) if key == _GraphKeys.GLOBAL_VARIABLES else []


def placeholder(dtype, shape=None, name=None):
    raise RuntimeError("tf_shim is eager: feed tensors directly")


def install():
    """Registers the shim as `tensorflow` and `tensorflow_probability` (only if the real ones are absent)."""
    tf = types.ModuleType("tensorflow")
    g = globals()
    for k in ("constant exp log sqrt square tanh stop_gradient identity maximum minimum multiply add equal matmul "
              "squeeze expand_dims transpose argmax one_hot cast clip_by_value add_n reduce_sum reduce_mean reduce_max "
              "reshape shape where stack concat split meshgrid gather gather_nd batch_gather unique_with_counts map_fn "
              "scatter_nd_update scatter_update assign custom_gradient is_nan is_inf logical_or zeros_like ones_like "
              "constant_initializer zeros_initializer ones_initializer truncated_normal_initializer "
              "glorot_uniform_initializer orthogonal_initializer variance_scaling_initializer get_variable Variable "
              "variable_scope get_collection placeholder float32 float64 int32 int64").split():
        setattr(tf, k, g[k])
    tf.range = range_
    tf.bool = bool_
    tf.name_scope = _noop_ctx
    tf.control_dependencies = _noop_ctx
    tf.GraphKeys = _GraphKeys
    tf.make_template = lambda name, fn, **kw: fn
    tf.group = lambda *a, **k: None
    tf.no_op = lambda *a, **k: None
    tf.convert_to_tensor = convert
    tf.nn = types.SimpleNamespace(softmax=softmax, log_softmax=log_softmax, softplus=softplus, tanh=tanh, relu6=relu6,
                                  relu=relu, moments=moments, l2_loss=lambda v: T(0.5 * torch.sum(_raw(v) ** 2)))
    tf.math = types.SimpleNamespace(atanh=atanh, log=log, square=square, top_k=top_k, reduce_max=reduce_max,
                                    reduce_sum=reduce_sum, greater_equal=lambda a, b: convert(a) >= b, exp=exp)
    tf.random = types.SimpleNamespace(categorical=random_categorical, uniform=random_uniform)
    tf.distributions = types.SimpleNamespace(Normal=Normal, Categorical=Categorical)
    tf.train = types.SimpleNamespace(Optimizer=object, get_or_create_global_step=lambda: T(torch.tensor(0.)))
    tf.summary = types.SimpleNamespace()
    tfp = types.ModuleType("tensorflow_probability")
    tfp.distributions = types.SimpleNamespace(RelaxedOneHotCategorical=RelaxedOneHotCategorical, Normal=Normal)
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow_probability"] = tfp
    return tf


def import_reference(root="/root/reference"):
    """Imports the reference's network modules by path under the package name `pfpn_ref`, without
    executing the reference's package __init__ files (they import every algorithm)."""
    import importlib
    install()
    pkg = types.ModuleType("pfpn_ref")
    pkg.__path__ = [f"{root}/networks"]
    sys.modules["pfpn_ref"] = pkg
    sub = types.ModuleType("pfpn_ref.actor_critic")
    sub.__path__ = [f"{root}/networks/actor_critic"]
    sys.modules["pfpn_ref.actor_critic"] = sub
    mods = {}
    for name in ("ops", "utils", "actor_critic.actor_critic", "actor_critic.a2c", "actor_critic.ppo", "actor_critic.sac"):
        mods[name.split(".")[-1]] = importlib.import_module(f"pfpn_ref.{name}")
    return mods
