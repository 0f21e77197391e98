"""A NumPy stand-in for the eager subset of TensorFlow 1.x that the reference's box / matcher / target-assigner code
touches, so that THAT code (under /root/reference, Python 2 + TF 1.7) can be executed here to produce golden vectors.

`install()` registers a fake `tensorflow` module; `load_reference_module(name)` imports a module of the reference after a
mechanical Python-2 -> 3 source transformation done in memory (dict.has_key / iteritems / xrange); nothing under
/root/reference is modified.  Every op computes immediately on ndarrays; graph-only constructs (control_dependencies,
name_scope, assert_*) are no-ops; `tf.cond` simply calls the selected branch.  Only what the golden generators need is
implemented -- an unknown attribute raises AttributeError naming it."""
import builtins
import importlib.util
import sys
import types

import numpy as np


class Dim(object):
    def __init__(self, v):
        self.value = v

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dim) else o)

    def __int__(self):
        return int(self.value)

    __index__ = __int__


class Shape(object):
    def __init__(self, dims):
        self.dims = [None if getattr(d, "value", d) is None else int(d) for d in dims]

    ndims = property(lambda self: len(self.dims))

    def as_list(self):
        return list(self.dims)

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(Dim(d) for d in self.dims)

    def __getitem__(self, i):
        return Shape(self.dims[i]) if isinstance(i, slice) else Dim(self.dims[i])

    def assert_has_rank(self, r):
        assert len(self.dims) == r, (self.dims, r)

    def assert_is_compatible_with(self, other):
        return True

    def is_fully_defined(self):
        return True

    def num_elements(self):
        return int(np.prod(self.dims))

    def merge_with(self, other):
        return self

    def with_rank(self, r):
        self.assert_has_rank(r)
        return self


class DimInt(int):
    """An element of `tensor.shape`: an int NumPy accepts, with TF's `.value`."""
    value = property(lambda self: int(self))


class ShapeTuple(tuple):
    """`tensor.shape` as TF exposes it (ndims / as_list / [i].value) while staying the plain tuple NumPy expects."""
    def __new__(cls, dims=()):
        return tuple.__new__(cls, (DimInt(d) for d in dims))

    ndims = property(lambda self: len(self))
    dims = property(lambda self: list(self))

    def as_list(self):
        return list(self)


class TT(np.ndarray):
    @property
    def shape(self):
        return ShapeTuple(np.ndarray.shape.__get__(self))

    def get_shape(self):
        return Shape(np.ndarray.shape.__get__(self))

    def set_shape(self, shape):
        return None

    # TF tensors are immutable: `x /= s` in the reference rebinds the name to a NEW tensor.  ndarray would write through
    # views (e.g. the rows `tf.unstack` returns) into the caller's data, so the augmented assignments copy instead.
    __iadd__ = lambda self, o: self + o
    __isub__ = lambda self, o: self - o
    __imul__ = lambda self, o: self * o
    __itruediv__ = lambda self, o: self / o
    __ifloordiv__ = lambda self, o: self // o


def t(a, dtype=None):
    if isinstance(a, TT) and dtype is None:
        return a
    return np.asarray(a, dtype=dtype).view(TT)


def _w(fn):
    def wrapped(*a, **k):
        k.pop("name", None)
        r = fn(*a, **k)
        if isinstance(r, (tuple, list)):
            return type(r)(t(x) if isinstance(x, np.ndarray) else x for x in r)
        return t(r)
    return wrapped


class _Ctx(object):
    def __init__(self, v=None):
        self.v = v

    def __enter__(self):
        return self.v

    def __exit__(self, *a):
        return False


def _ints(shape):
    return [int(s) for s in np.asarray(shape).reshape(-1)]


def _axis(a):
    return tuple(int(v) for v in a) if isinstance(a, (list, tuple, np.ndarray)) else a


def _where(cond, x=None, y=None):
    cond = np.asarray(cond)
    if x is None:
        return np.argwhere(cond).astype(np.int64)
    x, y = np.asarray(x), np.asarray(y)
    if cond.ndim == 1 and x.ndim > 1:                       # TF1 broadcasting rule: a vector condition picks rows
        cond = cond.reshape((-1,) + (1,) * (x.ndim - 1))
    return np.where(cond, x, y)


def _dynamic_stitch(indices, data):
    n = max(int(np.max(i)) + 1 if np.size(i) else 0 for i in indices)
    first = np.asarray(data[0])
    item_shape = None
    for i, d in zip(indices, data):
        d = np.asarray(d)
        item_shape = d.shape[np.asarray(i).ndim:]
        break
    out = np.zeros((n,) + tuple(item_shape), first.dtype)
    for i, d in zip(indices, data):
        i, d = np.asarray(i), np.asarray(d)
        out[i.reshape(-1)] = d.reshape((-1,) + tuple(item_shape))
    return out


def _split(value, num_or_size_splits, axis=0):
    value = np.asarray(value)
    if isinstance(num_or_size_splits, int):
        return [t(x) for x in np.split(value, num_or_size_splits, axis)]
    idx = np.cumsum(_ints(num_or_size_splits))[:-1]
    return [t(x) for x in np.split(value, idx, axis)]


def _setdiff1d(x, y, **k):
    x, y = np.asarray(x), np.asarray(y)
    keep = ~np.isin(x, y)
    return t(x[keep]), t(np.nonzero(keep)[0].astype(np.int32))


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool, tf.string = np.float32, np.float64, np.int32, np.int64, np.bool_, np.str_
    tf.Tensor, tf.Variable, tf.SparseTensor = TT, type("Variable", (), {}), type("SparseTensor", (), {})
    tf.TensorShape, tf.Dimension = Shape, Dim
    tf.name_scope = lambda *a, **k: _Ctx("scope")
    tf.variable_scope = lambda *a, **k: _Ctx("scope")
    tf.control_dependencies = lambda *a, **k: _Ctx()
    tf.device = lambda *a, **k: _Ctx()
    for n in ("assert_equal", "Assert", "assert_less", "assert_greater", "assert_less_equal", "assert_greater_equal",
              "assert_non_negative", "no_op", "group"):
        setattr(tf, n, lambda *a, **k: None)
    tf.contrib = types.SimpleNamespace(slim=None, framework=types.SimpleNamespace(is_tensor=lambda x: isinstance(x, np.ndarray)))
    tf.constant = _w(lambda v, dtype=None, shape=None: np.broadcast_to(np.asarray(v, dtype), _ints(shape)).copy()
                     if shape is not None else np.asarray(v, dtype))
    tf.convert_to_tensor = _w(lambda v, dtype=None: np.asarray(v, dtype))
    tf.identity = _w(lambda x: np.asarray(x))
    tf.stop_gradient = tf.identity
    tf.cast = _w(lambda x, dtype: np.asarray(x).astype(dtype))
    tf.to_float = _w(lambda x: np.asarray(x, np.float32))
    tf.to_int32 = _w(lambda x: np.asarray(x).astype(np.int32))
    tf.to_int64 = _w(lambda x: np.asarray(x).astype(np.int64))
    tf.shape = _w(lambda x, out_type=np.int32: np.asarray(np.shape(x), out_type))
    tf.size = _w(lambda x, out_type=np.int32: np.asarray(np.size(x), out_type))
    tf.rank = _w(lambda x: np.asarray(np.ndim(x), np.int32))
    def _reshape(x, shape):
        x, shape = np.asarray(x), list(_ints(shape))
        if x.size == 0 and -1 in shape and 0 in shape:
            # TF's ReshapeOp: when the requested shape holds a zero, the missing dimension is inferred from the NON-ZERO
            # dimensions of input and request
            nz = lambda dims: int(np.prod([d for d in dims if d not in (0, -1)] or [1]))
            shape[shape.index(-1)] = nz(x.shape) // nz(shape)
        return np.reshape(x, shape)
    tf.reshape = _w(_reshape)
    tf.expand_dims = _w(lambda x, axis=None, dim=None: np.expand_dims(x, axis if axis is not None else dim))
    def _squeeze(x, axis=None, squeeze_dims=None):
        ax = axis if axis is not None else squeeze_dims
        return np.squeeze(x, tuple(ax) if isinstance(ax, (list, tuple)) else ax)
    tf.squeeze = _w(_squeeze)
    tf.stack = _w(lambda xs, axis=0: np.stack([np.asarray(x) for x in xs], axis))
    tf.unstack = lambda x, num=None, axis=0: [t(v) for v in np.moveaxis(np.asarray(x), axis, 0)]
    tf.concat = _w(lambda xs, axis: np.concatenate([np.asarray(x) for x in xs], axis))
    tf.split = lambda value, num_or_size_splits, axis=0, **k: _split(value, num_or_size_splits, axis)
    tf.transpose = _w(lambda x, perm=None: np.transpose(x, perm))
    tf.tile = _w(lambda x, m: np.tile(x, _ints(m)))
    def _gather(params, indices, axis=0, **k):
        params, indices = np.asarray(params), np.asarray(indices).astype(np.int64)
        if params.size == 0 and axis == 0:
            # TF's GatherOp skips the copy (and with it the bounds check) when the slices are empty: the dummy
            # [n/q, q, 0, 0] masks of batch_multiclass_non_max_suppression are gathered with box indices >= n/q
            return np.zeros(indices.shape + params.shape[1:], params.dtype)
        return np.take(params, indices, axis)
    tf.gather = _w(_gather)
    tf.boolean_mask = _w(lambda x, mask, **k: np.asarray(x)[np.asarray(mask).astype(bool)])
    tf.where = _w(_where)
    tf.dynamic_stitch = _w(_dynamic_stitch)
    def _range(*a, start=None, limit=None, delta=None, dtype=np.int32):
        if limit is not None:
            a = (0 if start is None else start, limit) + (() if delta is None else (delta,))
        return np.arange(*[int(v) for v in a], dtype=dtype)
    tf.range = _w(_range)
    tf.ones = _w(lambda shape, dtype=np.float32: np.ones(_ints(shape), dtype))
    tf.zeros = _w(lambda shape, dtype=np.float32: np.zeros(_ints(shape), dtype))
    tf.ones_like = _w(lambda x, dtype=None: np.ones_like(x, dtype=dtype))
    tf.zeros_like = _w(lambda x, dtype=None: np.zeros_like(x, dtype=dtype))
    tf.fill = _w(lambda shape, v: np.full(_ints(shape), v))
    for n, f in (("equal", np.equal), ("not_equal", np.not_equal), ("greater", np.greater), ("greater_equal", np.greater_equal),
                 ("less", np.less), ("less_equal", np.less_equal), ("logical_and", np.logical_and),
                 ("logical_or", np.logical_or), ("logical_not", np.logical_not), ("add", np.add),
                 ("subtract", np.subtract), ("multiply", np.multiply), ("divide", np.divide), ("truediv", np.true_divide),
                 ("maximum", np.maximum), ("minimum", np.minimum), ("exp", np.exp), ("log", np.log), ("sqrt", np.sqrt),
                 ("square", np.square), ("abs", np.abs), ("is_nan", np.isnan)):
        setattr(tf, n, _w(f))
    for n, f in (("reduce_max", np.max), ("reduce_min", np.min), ("reduce_sum", np.sum), ("reduce_mean", np.mean),
                 ("reduce_any", np.any), ("reduce_all", np.all)):
        setattr(tf, n, _w(lambda x, axis=None, keep_dims=False, keepdims=False, reduction_indices=None, f=f:
                          f(np.asarray(x), axis=_axis(axis if axis is not None else reduction_indices),
                            keepdims=bool(keep_dims or keepdims))))
    tf.argmax = _w(lambda x, axis=0, output_type=np.int64, dimension=None:
                   np.argmax(np.asarray(x), axis if dimension is None else dimension).astype(output_type))
    tf.cond = lambda pred, true_fn=None, false_fn=None, fn1=None, fn2=None, **k: \
        (true_fn or fn1)() if bool(np.asarray(pred)) else (false_fn or fn2)()
    tf.setdiff1d = _setdiff1d
    tf.AUTO_REUSE = "AUTO_REUSE"
    tf.add_to_collection = lambda *a, **k: None
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                       image=lambda *a, **k: None)
    tf.pad = _w(lambda x, paddings, mode="CONSTANT", constant_values=0:
                np.pad(np.asarray(x), [tuple(int(v) for v in p) for p in np.asarray(paddings)], constant_values=constant_values))
    tf.one_hot = _w(lambda idx, depth, on_value=1.0, off_value=0.0, dtype=np.float32, axis=-1:
                    np.where(np.arange(int(depth)) == np.asarray(idx)[..., None], on_value, off_value).astype(dtype))
    tf.random_shuffle = _w(lambda x, seed=None: np.random.permutation(np.asarray(x)))
    tf.slice = _w(lambda x, begin, size: np.asarray(x)[tuple(slice(int(b), None if int(s) < 0 else int(b) + int(s))
                                                         for b, s in zip(_ints(begin), _ints(size)))])
    tf.nn = types.SimpleNamespace()
    tf.image = types.SimpleNamespace()
    tf.app = types.SimpleNamespace(flags=types.SimpleNamespace())
    return tf


def install():
    """Fake `tensorflow` + NumPy-1.x / Python-2 aliases.  Returns the module."""
    np.bool, np.float, np.int, np.NAN = bool, float, int, np.nan
    builtins.xrange = range
    tf = make_tf()
    sys.modules["tensorflow"] = tf
    return tf


def load_reference_module(name, root="/root/reference", extra_subst=()):
    """Import `name` (dotted) from the reference with has_key / iteritems / xrange rewritten in memory."""
    if name in sys.modules:
        return sys.modules[name]
    parts = name.split(".")
    for i in range(1, len(parts)):
        pkg = ".".join(parts[:i])
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [root + "/" + "/".join(parts[:i])]
            sys.modules[pkg] = m
    path = root + "/" + "/".join(parts) + ".py"
    src = open(path).read()
    for a, b in (("params.has_key('extension')", "('extension' in params)"), (".iteritems()", ".items()"),
                 (".itervalues()", ".values()")) + tuple(extra_subst):
        src = src.replace(a, b)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return mod
