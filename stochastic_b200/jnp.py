"""`jax.numpy`-shaped namespace whose functions work on traced arrays.

User drift/noise callables written for the reference use `jax.numpy` (e.g. reference
`tests/test_sde_solver.py:50-63`).  JAX is not a dependency here: the same callables are
written against this module (`import stochastic_b200.numpy as jnp`) and are traced by
calling them on a `TArray` - a numpy object array whose elements are `trace.expr.Expr`
nodes.  numpy itself provides indexing, broadcasting, reshape, sum, matmul, tensordot, roll,
stack ... for object arrays; this module adds the transcendental functions, `where`,
comparisons, `linalg.cholesky` and friends.

On concrete inputs (python numbers, numpy arrays, `DeviceArray` results of a solve) every
function falls through to numpy on the host, so post-processing code written for the
reference keeps working.
"""
import builtins

import numpy as _np

from . import config
from .trace import expr as E

pi = _np.pi
e = _np.e
inf = _np.inf
nan = _np.nan
newaxis = None
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_
ndarray = _np.ndarray


class TArray(_np.ndarray):
    """numpy object array of Expr nodes (the tracer value)."""
    __array_priority__ = 100

    def _cmp(self, other, ufunc):
        return _np.asarray(ufunc(_np.asarray(self), _np.asarray(_as_obj(other)), dtype=object)).view(TArray)

    def __gt__(self, o): return self._cmp(o, _np.greater)
    def __ge__(self, o): return self._cmp(o, _np.greater_equal)
    def __lt__(self, o): return self._cmp(o, _np.less)
    def __le__(self, o): return self._cmp(o, _np.less_equal)
    def __eq__(self, o): return self._cmp(o, _np.equal)
    def __ne__(self, o): return self._cmp(o, _np.not_equal)
    __hash__ = None

    @property
    def at(self):
        return _At(self)


class _At:
    """`x.at[idx].set(v)` / `.add(v)` of jax arrays (functional updates)."""

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Upd:
            def set(self, v):
                out = _np.array(arr, dtype=object, copy=True)
                out[idx] = _as_obj(v)
                return out.view(TArray)

            def add(self, v):
                out = _np.array(arr, dtype=object, copy=True)
                out[idx] = out[idx] + _as_obj(v)
                return out.view(TArray)
        return _Upd()


def _is_traced(x):
    if isinstance(x, E.Expr):
        return True
    if isinstance(x, _np.ndarray):
        return x.dtype == object
    if isinstance(x, (list, tuple)):
        return builtins.any(_is_traced(v) for v in x)
    return False


def _host(x):
    """Concrete value -> numpy (DeviceArray and torch tensors are copied to the host)."""
    if hasattr(x, "__sdeb200_host__"):
        return x.__sdeb200_host__()
    if type(x).__module__.startswith("torch") and hasattr(x, "detach"):
        return x.detach().cpu().numpy()          # torch tensors (any device) are accepted wherever arrays are
    if isinstance(x, (list, tuple)):
        return type(x)(_host(v) for v in x)
    return x


def _as_obj(x):
    if isinstance(x, E.Expr):
        return x
    if isinstance(x, _np.ndarray) and x.dtype == object:
        return x
    return _np.asarray(_host(x))


def _tarray(x):
    return _np.asarray(x, dtype=object).view(TArray)


def symbols(shape, start=0):
    """Fresh traced array of the given shape whose elements are var(start), var(start+1), ..."""
    n = int(_np.prod(shape, dtype=_np.int64)) if len(shape) else 1
    flat = _np.empty(n, dtype=object)
    for i in range(n):
        flat[i] = E.var(start + i)
    return flat.reshape(shape).view(TArray)


def _map1(fn, x):
    if isinstance(x, E.Expr):
        return fn(x)
    out = _np.frompyfunc(lambda v: fn(E.wrap(v)), 1, 1)(_np.asarray(x, dtype=object))
    return out.view(TArray) if isinstance(out, _np.ndarray) else out


def _map2(fn, x, y):
    if isinstance(x, E.Expr) and (isinstance(y, E.Expr) or E._is_number(y)):
        return fn(x, E.wrap(y))
    if isinstance(y, E.Expr) and E._is_number(x):
        return fn(E.wrap(x), y)
    out = _np.frompyfunc(lambda a, b: fn(E.wrap(a), E.wrap(b)), 2, 1)(
        _np.asarray(_as_obj(x), dtype=object), _np.asarray(_as_obj(y), dtype=object))
    return out.view(TArray) if isinstance(out, _np.ndarray) else out


def _unary(name, npfunc):
    def f(x):
        if _is_traced(x):
            return _map1(lambda v: E.unary(name, v), x)
        return npfunc(_host(x))
    f.__name__ = name
    return f


sin = _unary("sin", _np.sin)
cos = _unary("cos", _np.cos)
tan = _unary("tan", _np.tan)
exp = _unary("exp", _np.exp)
log = _unary("log", _np.log)
sqrt = _unary("sqrt", _np.sqrt)
tanh = _unary("tanh", _np.tanh)
sinh = _unary("sinh", _np.sinh)
cosh = _unary("cosh", _np.cosh)
arctan = _unary("atan", _np.arctan)
arcsin = _unary("asin", _np.arcsin)
arccos = _unary("acos", _np.arccos)
log1p = _unary("log1p", _np.log1p)
expm1 = _unary("expm1", _np.expm1)
floor = _unary("floor", _np.floor)
sign = _unary("sign", _np.sign)
absolute = _unary("abs", _np.abs)
abs = absolute
negative = _unary("neg", _np.negative)
logical_not = _unary("not", _np.logical_not)


def _binary(name, npfunc, fn=None):
    fn = fn or (lambda a, b: E.binary(name, a, b))

    def f(x, y):
        if _is_traced(x) or _is_traced(y):
            return _map2(fn, x, y)
        return npfunc(_host(x), _host(y))
    f.__name__ = name
    return f


maximum = _binary("max", _np.maximum)
minimum = _binary("min", _np.minimum)
fmod = _binary("fmod", _np.fmod)
arctan2 = _binary("atan2", _np.arctan2)
power = _binary("pow", _np.power, E.power)
greater = _binary("gt", _np.greater)
greater_equal = _binary("ge", _np.greater_equal)
less = _binary("lt", _np.less)
less_equal = _binary("le", _np.less_equal)
equal = _binary("eq", _np.equal)
not_equal = _binary("ne", _np.not_equal)
logical_and = _binary("and", _np.logical_and)
logical_or = _binary("or", _np.logical_or)
add = _binary("add", _np.add, E.add)
subtract = _binary("sub", _np.subtract, E.sub)
multiply = _binary("mul", _np.multiply, E.mul)
divide = _binary("div", _np.divide, E.div)
true_divide = divide


def arctanh(x):
    if _is_traced(x):
        return 0.5 * (log1p(x) - log1p(-x))
    return _np.arctanh(_host(x))


def square(x):
    return x * x


def where(c, a, b):
    if _is_traced(c) or _is_traced(a) or _is_traced(b):
        c, a, b = _np.broadcast_arrays(_np.asarray(_as_obj(c), dtype=object), _np.asarray(_as_obj(a), dtype=object),
                                       _np.asarray(_as_obj(b), dtype=object))
        out = _np.frompyfunc(lambda p, q, r: E.where(E.wrap(p), E.wrap(q), E.wrap(r)), 3, 1)(c, a, b)
        return out.view(TArray) if isinstance(out, _np.ndarray) else out
    return _np.where(_host(c), _host(a), _host(b))


select = where


def clip(x, a_min=None, a_max=None, min=None, max=None):
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    if lo is not None:
        x = maximum(x, lo)
    if hi is not None:
        x = minimum(x, hi)
    return x


# -- array construction / manipulation -------------------------------------------------------
def array(x, dtype=None, copy=True):
    if _is_traced(x):
        return _tarray(_np.array(_lists(x), dtype=object))
    return _np.array(_host(x), dtype=dtype)


def _lists(x):
    """Nested lists/arrays -> nested lists, so that np.array(..., dtype=object) stacks them."""
    if isinstance(x, _np.ndarray):
        return x.tolist() if x.ndim else x.item()
    if isinstance(x, (list, tuple)):
        return [_lists(v) for v in x]
    return x


asarray = array


def _passthrough(name):
    npfunc = getattr(_np, name)

    def f(*args, **kwargs):
        args = tuple(_host(a) for a in args)
        out = npfunc(*args, **kwargs)
        if isinstance(out, _np.ndarray) and out.dtype == object:
            return out.view(TArray)
        return out
    f.__name__ = name
    return f


for _name in ("reshape", "transpose", "swapaxes", "moveaxis", "roll", "tile", "repeat", "expand_dims", "squeeze",
              "sum", "prod", "cumsum", "dot", "matmul", "tensordot", "outer", "inner", "trace", "diag", "diagonal",
              "zeros", "ones", "eye", "identity", "zeros_like", "ones_like", "full", "arange", "linspace",
              "diag_indices", "tril_indices", "triu_indices", "tril", "triu", "ravel", "flip", "broadcast_to",
              "isscalar", "ndim", "shape", "size", "take", "meshgrid", "indices", "argmax", "argmin", "max", "min",
              "amax", "amin", "std", "var", "median", "percentile", "quantile", "isnan", "isfinite", "isclose",
              "allclose", "all", "any", "cumprod", "sort", "argsort", "histogram", "nonzero", "count_nonzero",
              "logical_xor", "rint", "round", "ceil", "hypot", "deg2rad", "rad2deg", "nan_to_num", "array_equal",
              "split", "hstack", "vstack", "atleast_1d", "atleast_2d", "issubdtype", "floating", "integer",
              "savetxt", "loadtxt", "empty", "empty_like", "copy", "unique", "interp", "kron", "cross",
              "ediff1d", "pad", "diff", "cov", "corrcoef", "average", "nanmean", "nanstd", "histogram2d", "digitize",
              "searchsorted", "bincount", "gradient", "trapezoid", "convolve", "polyfit", "polyval", "append",
              "insert", "delete", "column_stack", "dstack", "array_split", "flipud", "fliplr", "rot90", "linalg"):
    if hasattr(_np, _name) and _name not in globals():
        obj = getattr(_np, _name)
        globals()[_name] = _passthrough(_name) if callable(obj) and not isinstance(obj, type) else obj


def stack(arrays, axis=0):
    arrays = [_as_obj(a) for a in arrays]
    if builtins.any(_is_traced(a) for a in arrays):
        return _tarray(_np.stack([_np.asarray(a, dtype=object) for a in arrays], axis=axis))
    return _np.stack(arrays, axis=axis)


def concatenate(arrays, axis=0):
    arrays = [_as_obj(a) for a in arrays]
    if builtins.any(_is_traced(a) for a in arrays):
        return _tarray(_np.concatenate([_np.asarray(a, dtype=object) for a in arrays], axis=axis))
    return _np.concatenate(arrays, axis=axis)


def mean(x, axis=None, **kw):
    if _is_traced(x):
        x = _np.asarray(x, dtype=object)
        n = x.size if axis is None else x.shape[axis]
        return _np.sum(x, axis=axis) / n
    return _np.mean(_host(x), axis=axis, **kw)


def einsum(subscripts, *operands, **kw):
    ops = [_as_obj(o) for o in operands]
    if builtins.any(_is_traced(o) for o in ops):
        out = _np.einsum(subscripts, *[_np.asarray(o, dtype=object) for o in ops])
        return out.view(TArray) if isinstance(out, _np.ndarray) else out
    return _np.einsum(subscripts, *ops, **kw)


class _Linalg:
    @staticmethod
    def cholesky(a):
        """Lower Cholesky factor; symbolic Cholesky-Banachiewicz on traced input."""
        if not _is_traced(a):
            return _np.linalg.cholesky(_host(a))
        a = _np.asarray(a, dtype=object)
        n = a.shape[0]
        L = _np.empty((n, n), dtype=object)
        L[:, :] = E.ZERO
        for j in range(n):
            s = E.wrap(a[j, j])
            for k in range(j):
                s = s - L[j, k] * L[j, k]
            if config.get("library_rsqrt") and E.TRACE_F64:
                inv = E.unary("rsqrt", s)
                L[j, j] = s * inv
            else:
                d = E.unary("sqrt", s)
                L[j, j] = d
                inv = E.div(E.ONE, d)
            for i in range(j + 1, n):
                s = E.wrap(a[i, j])
                for k in range(j):
                    s = s - L[i, k] * L[j, k]
                L[i, j] = s * inv
        return L.view(TArray)

    @staticmethod
    def norm(x, axis=None):
        if _is_traced(x):
            return sqrt(_np.sum(_np.asarray(x, dtype=object) ** 2, axis=axis))
        return _np.linalg.norm(_host(x), axis=axis)

    @staticmethod
    def inv(a):
        return _np.linalg.inv(_host(a))

    @staticmethod
    def det(a):
        return _np.linalg.det(_host(a))

    @staticmethod
    def solve(a, b):
        return _np.linalg.solve(_host(a), _host(b))

    @staticmethod
    def eigh(a):
        return _np.linalg.eigh(_host(a))


linalg = _Linalg()
