"""A small NumPy stand-in for the TensorFlow-1.x graph API, just large enough to execute the reference's hot-path
modules (trmps/mps/mps.py, trmps/optimizers/{baseoptimizer,optimizer}.py, trmps/utils.py) UNMODIFIED.

TEST INFRASTRUCTURE ONLY (see oracle/make_golden_from_reference.py).  TensorFlow 1.2 cannot be installed in this
environment, so the reference cannot be run as shipped.  What this shim preserves is everything the reference
itself wrote: its Python control flow, graph structure, index conventions, loop bounds, transposes, Armijo and
accept/revert logic.  What it replaces is the execution of each TF op by the NumPy function of the same meaning
(einsum, tensordot, svd, softmax cross-entropy, ...), in float32 like the reference graph.

Graph nodes are lazy (`Tensor`); `Session.run` evaluates them with a feed dict.  `tf.while_loop` re-traces its
body once per iteration with concrete loop variables, which is semantically what the TF runtime does.
"""
import itertools
import sys
import types

import numpy as np

float32, float64, int32, int64, bool_ = np.float32, np.float64, np.int32, np.int64, np.bool_
_uid = itertools.count()


class _Ctx(object):
    def __init__(self, feed):
        self.feed = feed
        self.cache = {}
        self.bind = {}      # values of the symbolic loop variables of the while_loop iteration being evaluated
        self.var_override = {}      # Variable uid -> value used instead of its stored one (numerical gradients)


class Tensor(object):
    def __init__(self, fn, name=None):
        self._fn = fn
        self._uid = next(_uid)
        self.name = name

    def _eval(self, ctx):
        if self._uid not in ctx.cache:
            ctx.cache[self._uid] = self._fn(ctx)
        return ctx.cache[self._uid]

    def set_shape(self, shape):
        return None

    # arithmetic
    def __add__(self, o): return _binary(np.add, self, o)
    def __radd__(self, o): return _binary(np.add, o, self)
    def __sub__(self, o): return _binary(np.subtract, self, o)
    def __rsub__(self, o): return _binary(np.subtract, o, self)
    def __mul__(self, o): return _binary(np.multiply, self, o)
    def __rmul__(self, o): return _binary(np.multiply, o, self)
    def __truediv__(self, o): return _binary(np.true_divide, self, o)
    def __rtruediv__(self, o): return _binary(np.true_divide, o, self)
    def __neg__(self): return Tensor(lambda c: -self._eval(c))

    def __getitem__(self, key):
        return Tensor(lambda c: self._eval(c)[_resolve_key(key, c)])

    def __iter__(self):
        raise TypeError("Tensor objects are not iterable (as in TensorFlow 1.x)")

    __hash__ = object.__hash__


def _val(x, ctx):
    if isinstance(x, Tensor):
        return x._eval(ctx)
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return [_val(e, ctx) for e in x]
    return x


def _resolve_key(key, ctx):
    def one(k):
        if isinstance(k, Tensor):
            return int(k._eval(ctx))
        if isinstance(k, slice):
            f = lambda v: None if v is None else int(_val(v, ctx))
            return slice(f(k.start), f(k.stop), f(k.step))
        return k
    if isinstance(key, tuple):
        return tuple(one(k) for k in key)
    return one(key)


def _weak(v):
    """Python scalars stay 'weak' (they take the dtype of the tensor they meet), as TF's constant conversion does."""
    return v


def _binary(op, a, b):
    def fn(c):
        x, y = _val(a, c), _val(b, c)
        return op(x, y)
    return Tensor(fn)


def _const(value):
    return Tensor(lambda c: value)


def convert_to_tensor(x):
    return x if isinstance(x, Tensor) else _const(x)


# ----------------------------------------------------------------------------------------------- placeholders / variables
class _Placeholder(Tensor):
    def __init__(self, dtype, shape=None, default=None, name=None):
        self.dtype, self.default = dtype, default
        import traceback
        self._where = "".join(traceback.format_stack(limit=4)[:-2])
        Tensor.__init__(self, self._get, name)

    def _get(self, ctx):
        if self in ctx.feed:
            return np.asarray(ctx.feed[self], dtype=self.dtype)
        if self.default is not None:
            return np.asarray(_val(self.default, ctx), dtype=self.dtype)
        raise ValueError("placeholder %r was not fed; created at:\n%s" % (self.name, self._where))


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape, None, name)


Placeholder = placeholder


def placeholder_with_default(input, shape, name=None):
    dtype = input.dtype if hasattr(input, "dtype") else np.float32
    return _Placeholder(dtype, shape, input, name)


_VARIABLES = []


class Variable(Tensor):
    """Mutable: tf.train.GradientDescentOptimizer.minimize (below) updates .value; an evaluation context may override it
    (the numerical gradient evaluates the cost at perturbed values)."""

    def __init__(self, initial_value, dtype=None, trainable=True, name=None):
        if isinstance(initial_value, Tensor):
            initial_value = initial_value._eval(_Ctx({}))
        self.value = np.array(initial_value, dtype=dtype if dtype is not None else None)
        self.dtype = self.value.dtype
        Tensor.__init__(self, lambda c: c.var_override.get(self._uid, self.value), name)
        _VARIABLES.append(self)


def global_variables_initializer():
    return _const(None)


class _GradientDescentOptimizer(object):
    """tf.train.GradientDescentOptimizer(lr).minimize(cost): var <- var - lr * d cost / d var for every Variable the cost
    depends on.  TensorFlow differentiates the graph symbolically; this stand-in has no graph to differentiate (a Tensor is
    a closure), so the derivative of the REFERENCE'S OWN cost graph is taken numerically: central differences in float64
    (variables and fed values cast up; step 1e-6 relative to max(1, |v|)), accurate to ~1e-9 -- far below the float32
    round-off of the update itself.  Slow (two cost evaluations per parameter), meant for small golden-vector cases."""

    def __init__(self, learning_rate, *a, **k):
        self.lr = float(learning_rate)

    def minimize(self, cost, *a, **k):
        lr = self.lr

        def fn(c):
            feed64 = {k_: (np.asarray(v, dtype=np.float64) if np.asarray(v).dtype.kind == "f" else v) for k_, v in c.feed.items()}

            def cost_at(override):
                ctx = _Ctx(feed64)
                ctx.var_override = override
                return float(np.asarray(cost._eval(ctx), dtype=np.float64))

            base = {v._uid: v.value.astype(np.float64) for v in _VARIABLES}
            f0 = cost_at(base)
            grads = {}
            for v in _VARIABLES:
                x = base[v._uid]
                g = np.zeros_like(x)
                # does the cost depend on this variable at all?
                # (a pseudo-random direction: a uniform shift of all logits leaves a softmax cost unchanged)
                probe = dict(base)
                probe[v._uid] = x + 1e-3 * np.random.RandomState(v._uid % (2 ** 31)).standard_normal(x.shape)
                if cost_at(probe) == f0:
                    continue
                it = np.nditer(x, flags=["multi_index"])
                for _ in it:
                    idx = it.multi_index
                    h = 1e-6 * max(1.0, abs(x[idx]))
                    xp, xm = x.copy(), x.copy()
                    xp[idx] += h; xm[idx] -= h
                    op, om = dict(base), dict(base)
                    op[v._uid], om[v._uid] = xp, xm
                    g[idx] = (cost_at(op) - cost_at(om)) / (2 * h)
                grads[v._uid] = g
            for v in _VARIABLES:
                if v._uid in grads:
                    v.value = (v.value.astype(np.float64) - lr * grads[v._uid]).astype(v.dtype)
            return None
        return Tensor(fn)


train = types.SimpleNamespace(GradientDescentOptimizer=_GradientDescentOptimizer)


class TensorShape(object):
    def __init__(self, dims=None):
        self.dims = dims


# ----------------------------------------------------------------------------------------------- TensorArray
class TensorArray(Tensor):
    """Value: (dict index -> ndarray, static size or None).  write() returns a new TensorArray (functional)."""

    def __init__(self, dtype=None, size=0, dynamic_size=False, clear_after_read=None, infer_shape=None, _fn=None, name=None):
        if _fn is None:
            _fn = lambda c: ({}, None if dynamic_size else int(_val(size, c)))
        Tensor.__init__(self, _fn, name)

    def write(self, index, value):
        def fn(c):
            d, static = self._eval(c)
            d = dict(d)
            d[int(_val(index, c))] = np.asarray(_val(value, c))
            return (d, static)
        return TensorArray(_fn=fn)

    def read(self, index):
        def fn(c):
            d, _ = self._eval(c)
            return d[int(_val(index, c))]
        return Tensor(fn)

    def size(self):
        def fn(c):
            d, static = self._eval(c)
            return static if static is not None else (max(d) + 1 if d else 0)
        return Tensor(fn)


class _LoopVar(Tensor):
    def __init__(self):
        Tensor.__init__(self, lambda c: c.bind[self._uid])


class _LoopArray(TensorArray):
    def __init__(self):
        TensorArray.__init__(self, _fn=lambda c: c.bind[self._uid])


# ----------------------------------------------------------------------------------------------- control flow
def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10, name=None, **kw):
    """Like TensorFlow, cond and body are traced ONCE, at graph-construction time, on symbolic loop variables (the
    reference relies on this: its bodies read Python attributes that are re-assigned later).  Evaluation then
    iterates: bind the symbolic variables, drop the cached values of every node created during the trace, evaluate."""
    n = len(loop_vars)
    first = next(_uid)
    sym = [_LoopArray() if isinstance(v, TensorArray) else _LoopVar() for v in loop_vars]
    cond_t = convert_to_tensor(cond(*sym))
    out = body(*sym)
    if not isinstance(out, (list, tuple)):
        out = [out]
    outs_t = [convert_to_tensor(o) for o in out]
    last = next(_uid)
    initial = [convert_to_tensor(v) for v in loop_vars]

    def run(ctx):
        vals = [t._eval(ctx) for t in initial]
        saved = {v._uid: ctx.bind.get(v._uid) for v in sym}
        guard = 0
        while True:
            for u in [k for k in ctx.cache if first < k < last]:
                del ctx.cache[u]
            for v, val in zip(sym, vals):
                ctx.bind[v._uid] = val
            if not bool(cond_t._eval(ctx)):
                break
            vals = [t._eval(ctx) for t in outs_t]
            guard += 1
            if guard > 1000000:
                raise RuntimeError("while_loop did not terminate")
        for u in [k for k in ctx.cache if first < k < last]:
            del ctx.cache[u]
        for k, v in saved.items():
            if v is not None:
                ctx.bind[k] = v
        return vals

    whole = Tensor(run)
    outs = []
    for i in range(n):
        if isinstance(loop_vars[i], TensorArray):
            outs.append(TensorArray(_fn=lambda c, i=i: whole._eval(c)[i]))
        else:
            outs.append(Tensor(lambda c, i=i: whole._eval(c)[i]))
    return outs


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None):
    t = convert_to_tensor((true_fn or fn1)())       # both branches are traced at construction, only one is evaluated
    f = convert_to_tensor((false_fn or fn2)())
    return Tensor(lambda c: (t if bool(_val(pred, c)) else f)._eval(c))


def case(pred_fn_pairs, default=None, exclusive=False, name=None):
    pairs = list(pred_fn_pairs.items()) if isinstance(pred_fn_pairs, dict) else list(pred_fn_pairs)
    traced = [(p, convert_to_tensor(f())) for p, f in pairs]
    dflt = convert_to_tensor(default())

    def fn(c):
        for p, t in traced:
            if bool(_val(p, c)):
                return t._eval(c)
        return dflt._eval(c)
    return Tensor(fn)


def Print(input_, data, message=None, first_n=None, summarize=None, name=None):
    return convert_to_tensor(input_)


class name_scope(object):
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


# ----------------------------------------------------------------------------------------------- ops
def _op(f):
    def wrapped(*args, **kwargs):
        kwargs.pop("name", None)

        def fn(c):
            vals = [_val(a, c) for a in args]
            try:
                return f(*vals, **{k: _val(v, c) for k, v in kwargs.items()})
            except Exception as e:
                raise type(e)("%s [tf1_shim op with args %r]" % (e, [type(v).__name__ if not isinstance(v, (int, float)) else v for v in vals]))
        return Tensor(fn)
    return wrapped


einsum = _op(lambda eq, *xs: np.einsum(eq, *xs))
tensordot = _op(lambda a, b, axes: np.tensordot(a, b, axes=axes))
transpose = _op(lambda a, perm=None: np.transpose(a, perm))
reshape = _op(lambda a, shape: np.reshape(a, [int(s) for s in shape]))
tile = _op(lambda a, multiples: np.tile(a, [int(m) for m in multiples]))
shape = _op(lambda a: np.array(np.shape(a), dtype=np.int32))
size = _op(lambda a: np.int32(np.size(a)))
reduce_sum = _op(lambda a, axis=None: np.sum(a, axis=axis, dtype=np.asarray(a).dtype))
reduce_mean = _op(lambda a, axis=None: np.mean(a, axis=axis, dtype=np.asarray(a).dtype))
square = _op(np.square)
less = _op(np.less)
greater = _op(np.greater)
equal = _op(np.equal)
logical_and = _op(np.logical_and)
add = _op(np.add)
argmax = _op(lambda a, axis=None: np.argmax(a, axis=axis))
cast = _op(lambda a, dtype: np.asarray(a).astype(dtype))
zeros = _op(lambda shape, dtype=np.float32: np.zeros([int(s) for s in shape], dtype=dtype))
zeros_like = _op(np.zeros_like)
where = _op(np.where)
is_nan = _op(np.isnan)
round = _op(np.round)
matmul = _op(np.matmul)
diag = _op(np.diag)
diag_part = _op(np.diag)
boolean_mask = _op(lambda a, mask: np.asarray(a)[np.asarray(mask, dtype=bool)])
expand_dims = _op(lambda a, axis: np.expand_dims(a, axis))
fill = _op(lambda dims, value: np.full([int(d) for d in dims], value, dtype=np.float32))
stack = _op(lambda values, axis=0: np.stack(values, axis=axis))


def svd(tensor, full_matrices=False, compute_uv=True, name=None):
    """tf.svd returns (s, u, v) with a = u diag(s) v^T."""
    def fn(c):
        u, s, vh = np.linalg.svd(_val(tensor, c), full_matrices=full_matrices)
        return (s, u, vh.T.copy())
    whole = Tensor(fn)
    return tuple(Tensor(lambda c, i=i: whole._eval(c)[i]) for i in range(3))


def confusion_matrix(labels, predictions, num_classes=None, name=None):
    def fn(c):
        y, p = np.asarray(_val(labels, c)), np.asarray(_val(predictions, c))
        cm = np.zeros((num_classes, num_classes), dtype=np.int32)
        np.add.at(cm, (y, p), 1)
        return cm
    return Tensor(fn)


def _softmax(x):
    z = x - x.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def _softmax_xent(labels=None, logits=None, name=None, dim=-1):
    def fn(c):
        y, f = np.asarray(_val(labels, c)), np.asarray(_val(logits, c))
        z = f - f.max(axis=-1, keepdims=True)
        logsm = z - np.log(np.exp(z).sum(axis=-1, keepdims=True))
        return -(y * logsm).sum(axis=-1)
    return Tensor(fn)


def _avg_pool(value, ksize=None, strides=None, padding=None, data_format="NHWC"):
    """NHWC average pooling for the one configuration the reference uses (2x2 window, stride 2, even extents:
    MNISTpreprocessing.py:44-45); the window is summed row by row, left to right, in float32."""
    if list(ksize) != [1, 2, 2, 1] or list(strides) != [1, 2, 2, 1]:
        raise NotImplementedError("tf1_shim avg_pool: only ksize = strides = [1,2,2,1]")
    x = np.asarray(value)
    b, h, w, c = x.shape
    x = x.reshape(b, h // 2, 2, w // 2, 2, c)
    return ((x[:, :, 0, :, 0] + x[:, :, 0, :, 1]) + (x[:, :, 1, :, 0] + x[:, :, 1, :, 1])) * x.dtype.type(0.25)


nn = types.SimpleNamespace(softmax=_op(_softmax), softmax_cross_entropy_with_logits=_softmax_xent, avg_pool=_op(_avg_pool))
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(flatten=_op(lambda a: np.reshape(a, (np.shape(a)[0], -1)))))


class _NotAvailable(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("not provided by the NumPy TF1 shim (outside the hot path)")




class _Dummy(object):
    FULL_TRACE = 3

    def __init__(self, *a, **k): pass
    def close(self): pass
    def add_run_metadata(self, *a, **k): pass


RunOptions = _Dummy
RunMetadata = _Dummy
summary = types.SimpleNamespace(FileWriter=_Dummy)


# ----------------------------------------------------------------------------------------------- session
class Session(object):
    def __init__(self, *a, **k):
        self.graph = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def as_default(self):
        return self

    def run(self, fetches, feed_dict=None, options=None, run_metadata=None):
        ctx = _Ctx(feed_dict or {})

        def ev(f):
            if isinstance(f, (list, tuple)):
                return [ev(x) for x in f]
            if isinstance(f, Tensor):
                v = f._eval(ctx)
                if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], dict):
                    return v[0]
                return v
            return f
        return ev(fetches)


# ----------------------------------------------------------------------------------------------- installation
def install():
    """Registers this module as `tensorflow` (plus the few sub-modules the reference imports) and stubs matplotlib,
    which trmps/utils.py imports for plotting only."""
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    pkg_python = types.ModuleType("tensorflow.python")
    pkg_client = types.ModuleType("tensorflow.python.client")
    timeline_mod = types.ModuleType("tensorflow.python.client.timeline")
    timeline_mod.Timeline = _Dummy
    pkg_client.timeline = timeline_mod
    pkg_python.client = pkg_client
    sys.modules["tensorflow.python"] = pkg_python
    sys.modules["tensorflow.python.client"] = pkg_client
    sys.modules["tensorflow.python.client.timeline"] = timeline_mod
    # trmps/preprocessing/MNISTpreprocessing.py:12 imports the tutorial loader at module level (it downloads MNIST)
    for name in ("tensorflow.examples", "tensorflow.examples.tutorials", "tensorflow.examples.tutorials.mnist"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.examples.tutorials.mnist"].input_data = types.SimpleNamespace(read_data_sets=_NotAvailable)
    try:
        import skimage.measure  # noqa: F401
    except Exception:
        # FashionMNISTpreprocessing.py:9 imports skimage for one call: block_reduce(image, (2, 2), np.max)
        def block_reduce(image, block_size, func=np.sum, cval=0):
            (bh, bw), (h, w) = block_size, image.shape
            if h % bh or w % bw:
                raise NotImplementedError("tf1_shim block_reduce: extents must be multiples of the block")
            return func(image.reshape(h // bh, bh, w // bw, bw), axis=(1, 3))
        sk = types.ModuleType("skimage")
        sk.measure = types.ModuleType("skimage.measure")
        sk.measure.block_reduce = block_reduce
        sys.modules["skimage"], sys.modules["skimage.measure"] = sk, sk.measure
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        colors = types.ModuleType("matplotlib.colors")
        colors.Normalize = _Dummy
        mpl.pyplot, mpl.colors = plt, colors
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.colors"] = colors
    return me
