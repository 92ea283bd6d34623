"""Hash-consed scalar expression DAG: the tracing IR of the callable-lowering pipeline.

The reference hands the user's drift/noise callables to JAX, which traces them
(`jax.eval_shape`, reference `pychastic/sde_problem.py:58-59`), differentiates them
(`jax.jacobian` / `jax.hessian`, reference `pychastic/sde_solver.py:29-46`) and compiles
them with XLA.  Here the same callables are run once on arrays of `Expr` leaves; every
arithmetic operation builds a node of this DAG.  The DAG is then differentiated symbolically
(`diff`), simplified by construction (constant folding, 0/1 rules, common-subexpression
sharing through hash-consing) and lowered to CUDA text by `trace.emit`.

Nodes are immutable; structurally equal nodes are the same object.
"""
import hashlib
import math

import numpy as np

_TABLE = {}
_COUNTER = [0]

TRACE_F64 = True     # working precision of the trace in progress (set by sde_solver.build_source; library routines may pick
                     # their formulation by it, e.g. jnp.linalg.cholesky)
UNARY_FUNCS = ("neg", "sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "atan",
               "asin", "acos", "abs", "sign", "log1p", "expm1", "floor", "trunc", "not", "erf", "rsqrt")
BINARY_FUNCS = ("add", "sub", "mul", "div", "pow", "max", "min", "fmod", "atan2",
                "lt", "le", "gt", "ge", "eq", "ne", "and", "or")
BOOL_OPS = ("lt", "le", "gt", "ge", "eq", "ne", "and", "or", "not")
_COMMUTATIVE = ("add", "mul", "max", "min", "eq", "ne", "and", "or")


class Expr:
    __slots__ = ("op", "args", "value", "id", "skey", "__weakref__")
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, op, args=(), value=None):
        self.op = op
        self.args = args
        self.value = value
        self.id = _COUNTER[0]
        _COUNTER[0] += 1
        # structural key: a pure function of (op, operands, value), independent of what the process traced before.
        # Commutative operands are ordered by it, so the emitted text of a problem never depends on creation order.
        text = f"{op}|{','.join(str(a.skey) for a in args)}|{value!r}"
        self.skey = int.from_bytes(hashlib.blake2b(text.encode(), digest_size=8).digest(), "little")

    # -- construction ---------------------------------------------------------------------
    @staticmethod
    def _make(op, args=(), value=None):
        key = (op, tuple(a.id for a in args), value)
        node = _TABLE.get(key)
        if node is None:
            node = Expr(op, tuple(args), value)
            _TABLE[key] = node
        return node

    @property
    def is_const(self):
        return self.op == "const"

    @property
    def is_bool(self):
        return self.op in BOOL_OPS or (self.op == "const" and isinstance(self.value, bool))

    def __repr__(self):
        if self.op == "const":
            return repr(self.value)
        if self.op == "var":
            return f"x{self.value}"
        return f"{self.op}({', '.join(map(repr, self.args))})"

    def __hash__(self):
        return self.id

    def __bool__(self):
        if self.op == "const":
            return bool(self.value)
        raise TypeError(
            "the truth value of a traced quantity is not known at trace time; use "
            "stochastic_b200.numpy.where / lax.cond instead of Python control flow")

    def __float__(self):
        if self.op == "const":
            return float(self.value)
        raise TypeError("traced quantity has no concrete value")

    # -- operators ------------------------------------------------------------------------
    # (scalar op array) broadcasts over the array: numpy defers to us because __array_ufunc__ is None
    def _bin(self, o, fn, reverse=False):
        if isinstance(o, np.ndarray) and o.ndim > 0:
            f = (lambda v: fn(wrap(v), self)) if reverse else (lambda v: fn(self, wrap(v)))
            out = np.frompyfunc(f, 1, 1)(np.asarray(o, dtype=object))
            from .. import jnp
            return out.view(jnp.TArray)
        return fn(o, self) if reverse else fn(self, o)

    def __add__(self, o): return self._bin(o, add)
    def __radd__(self, o): return self._bin(o, add, True)
    def __sub__(self, o): return self._bin(o, sub)
    def __rsub__(self, o): return self._bin(o, sub, True)
    def __mul__(self, o): return self._bin(o, mul)
    def __rmul__(self, o): return self._bin(o, mul, True)
    def __truediv__(self, o): return self._bin(o, div)
    def __rtruediv__(self, o): return self._bin(o, div, True)
    def __pow__(self, o): return self._bin(o, power)
    def __rpow__(self, o): return self._bin(o, power, True)
    def __neg__(self): return neg(self)
    def __pos__(self): return self
    def __abs__(self): return unary("abs", self)
    def __lt__(self, o): return self._bin(o, lambda a, b: binary("lt", a, b))
    def __le__(self, o): return self._bin(o, lambda a, b: binary("le", a, b))
    def __gt__(self, o): return self._bin(o, lambda a, b: binary("gt", a, b))
    def __ge__(self, o): return self._bin(o, lambda a, b: binary("ge", a, b))
    def __eq__(self, o):
        if isinstance(o, Expr) or _is_number(o):
            return binary("eq", self, o)
        return NotImplemented
    def __ne__(self, o):
        if isinstance(o, Expr) or _is_number(o):
            return binary("ne", self, o)
        return NotImplemented
    def __and__(self, o): return binary("and", self, o)
    def __rand__(self, o): return binary("and", o, self)
    def __or__(self, o): return binary("or", self, o)
    def __ror__(self, o): return binary("or", o, self)
    def __invert__(self): return unary("not", self)

    # numpy calls these methods for ufuncs applied to object arrays (np.cos(obj) -> x.cos())
    def sin(self): return unary("sin", self)
    def cos(self): return unary("cos", self)
    def tan(self): return unary("tan", self)
    def exp(self): return unary("exp", self)
    def log(self): return unary("log", self)
    def sqrt(self): return unary("sqrt", self)
    def tanh(self): return unary("tanh", self)
    def sinh(self): return unary("sinh", self)
    def cosh(self): return unary("cosh", self)
    def arctan(self): return unary("atan", self)
    def arcsin(self): return unary("asin", self)
    def arccos(self): return unary("acos", self)
    def log1p(self): return unary("log1p", self)
    def expm1(self): return unary("expm1", self)
    def floor(self): return unary("floor", self)
    def conjugate(self): return self


def _is_number(x):
    return isinstance(x, (int, float, bool, np.integer, np.floating, np.bool_))


def trim_table(max_nodes=2_000_000):
    """Bound the hash-consing table between traces: beyond `max_nodes` entries everything except constants and
    variables is forgotten.  Nodes still referenced elsewhere stay valid; they merely stop being shared with
    nodes built later (sharing is an optimisation, never a correctness requirement)."""
    if len(_TABLE) > max_nodes:
        keep = {k: v for k, v in _TABLE.items() if k[0] in ("const", "var")}
        _TABLE.clear()
        _TABLE.update(keep)


def const(v):
    if isinstance(v, (bool, np.bool_)):
        return Expr._make("const", (), bool(v))
    v = float(v)
    if v == 0.0:
        v = 0.0  # merge -0.0 and 0.0
    return Expr._make("const", (), v)


def var(i):
    return Expr._make("var", (), int(i))


def wrap(x):
    if isinstance(x, Expr):
        return x
    if _is_number(x):
        return const(x)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return wrap(x.item())
    raise TypeError(f"cannot trace value of type {type(x)}")


ZERO = const(0.0)
ONE = const(1.0)
TWO = const(2.0)
HALF = const(0.5)
MINUS_ONE = const(-1.0)

_NUMERIC = {
    "neg": lambda a: -a, "sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp,
    "log": math.log, "sqrt": math.sqrt, "tanh": math.tanh, "sinh": math.sinh, "cosh": math.cosh,
    "atan": math.atan, "asin": math.asin, "acos": math.acos, "abs": abs,
    "sign": lambda a: (a > 0) - (a < 0), "log1p": math.log1p, "expm1": math.expm1,
    "floor": math.floor, "trunc": math.trunc, "not": lambda a: not a, "erf": math.erf,
    "rsqrt": lambda a: 1.0 / math.sqrt(a),
    "add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
    "div": lambda a, b: a / b, "pow": lambda a, b: a ** b, "max": max, "min": min,
    "fmod": math.fmod, "atan2": math.atan2,
    "lt": lambda a, b: a < b, "le": lambda a, b: a <= b, "gt": lambda a, b: a > b,
    "ge": lambda a, b: a >= b, "eq": lambda a, b: a == b, "ne": lambda a, b: a != b,
    "and": lambda a, b: bool(a) and bool(b), "or": lambda a, b: bool(a) or bool(b),
}


def _fold(op, *vals):
    try:
        r = _NUMERIC[op](*vals)
    except (ValueError, ZeroDivisionError, OverflowError):
        return None
    if isinstance(r, complex):
        return None
    return const(r)


def _after(a, b):
    """Canonical order of commutative operands: by structural key (ties, i.e. 64-bit collisions, by creation id)."""
    return (a.skey, a.id) > (b.skey, b.id)


def unary(op, a):
    a = wrap(a)
    if a.is_const:
        f = _fold(op, a.value)
        if f is not None:
            return f
    if op == "neg":
        return neg(a)
    return Expr._make(op, (a,))


def binary(op, a, b):
    a, b = wrap(a), wrap(b)
    if a.is_const and b.is_const:
        f = _fold(op, a.value, b.value)
        if f is not None:
            return f
    if op in _COMMUTATIVE and _after(a, b):
        a, b = b, a
    return Expr._make(op, (a, b))


def neg(a):
    a = wrap(a)
    if a.is_const:
        return const(-a.value)
    if a.op == "neg":
        return a.args[0]
    if a.op == "sub":
        return sub(a.args[1], a.args[0])
    if a.op == "mul" and a.args[0].is_const:
        return mul(const(-a.args[0].value), a.args[1])
    return Expr._make("neg", (a,))


def add(a, b):
    a, b = wrap(a), wrap(b)
    if a.is_const and b.is_const:
        return const(a.value + b.value)
    if a.is_const and a.value == 0.0:
        return b
    if b.is_const and b.value == 0.0:
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    if a.op == "neg":
        return sub(b, a.args[0])
    if _after(a, b):
        a, b = b, a
    return Expr._make("add", (a, b))


def sub(a, b):
    a, b = wrap(a), wrap(b)
    if a.is_const and b.is_const:
        return const(a.value - b.value)
    if b.is_const and b.value == 0.0:
        return a
    if a.is_const and a.value == 0.0:
        return neg(b)
    if a is b:
        return ZERO
    if b.op == "neg":
        return add(a, b.args[0])
    return Expr._make("sub", (a, b))


def mul(a, b):
    a, b = wrap(a), wrap(b)
    if a.is_const and b.is_const:
        return const(a.value * b.value)
    if b.is_const:
        a, b = b, a
    if a.is_const:
        if a.value == 0.0:
            return ZERO
        if a.value == 1.0:
            return b
        if a.value == -1.0:
            return neg(b)
        if b.op == "mul" and b.args[0].is_const:
            return mul(const(a.value * b.args[0].value), b.args[1])
        if b.op == "neg":
            return mul(const(-a.value), b.args[0])
        return Expr._make("mul", (a, b))
    if a.op == "neg" and b.op == "neg":
        return mul(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(mul(a.args[0], b))
    if b.op == "neg":
        return neg(mul(a, b.args[0]))
    # hoist constant factors: (c*x)*y -> c*(x*y)
    if a.op == "mul" and a.args[0].is_const:
        return mul(a.args[0], mul(a.args[1], b))
    if b.op == "mul" and b.args[0].is_const:
        return mul(b.args[0], mul(a, b.args[1]))
    if _after(a, b):
        a, b = b, a
    return Expr._make("mul", (a, b))


def div(a, b):
    a, b = wrap(a), wrap(b)
    if b.is_const and b.value != 0.0:
        if a.is_const:
            return const(a.value / b.value)
        if b.value == 1.0:
            return a
        if b.value == -1.0:
            return neg(a)
        # exact only for powers of two; otherwise keep the division for fidelity
        m, _ = math.frexp(b.value)
        if abs(m) == 0.5:
            return mul(const(1.0 / b.value), a)
    if a.is_const and a.value == 0.0:
        return ZERO
    if a.op == "neg" and b.op == "neg":
        return div(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(div(a.args[0], b))
    if b.op == "neg":
        return neg(div(a, b.args[0]))
    return Expr._make("div", (a, b))


def power(a, b):
    a, b = wrap(a), wrap(b)
    if b.is_const:
        e = b.value
        if a.is_const:
            f = _fold("pow", a.value, e)
            if f is not None:
                return f
        if e == 0.0:
            return ONE
        if e == 1.0:
            return a
        if e == 0.5:
            return unary("sqrt", a)
        if e == -0.5:
            return div(ONE, unary("sqrt", a))
        if float(e).is_integer() and abs(e) <= 64:
            return powi(a, int(e))
    return Expr._make("pow", (a, b))


def powi(a, n):
    if n < 0:
        return div(ONE, powi(a, -n))
    result = None
    base = a
    while n:
        if n & 1:
            result = base if result is None else mul(result, base)
        n >>= 1
        if n:
            base = mul(base, base)
    return result if result is not None else ONE


def where(c, a, b):
    c, a, b = wrap(c), wrap(a), wrap(b)
    if c.is_const:
        return a if c.value else b
    if a is b:
        return a
    return Expr._make("where", (c, a, b))


def maximum(a, b):
    return binary("max", a, b)


def minimum(a, b):
    return binary("min", a, b)


# ---------------------------------------------------------------------------------------
# symbolic differentiation (forward accumulation over the DAG, memoised per variable)
# ---------------------------------------------------------------------------------------
def diff(e, var_index, memo=None):
    """d e / d x_{var_index}."""
    if memo is None:
        memo = {}
    return _diff(wrap(e), var_index, memo)


def _diff(e, v, memo):
    key = e.id
    r = memo.get(key)
    if r is not None:
        return r
    op = e.op
    if op == "const":
        r = ZERO
    elif op == "var":
        r = ONE if e.value == v else ZERO
    else:
        a = e.args
        d = [_diff(x, v, memo) for x in a] if op not in ("where",) else None
        if op == "add":
            r = add(d[0], d[1])
        elif op == "sub":
            r = sub(d[0], d[1])
        elif op == "neg":
            r = neg(d[0])
        elif op == "mul":
            r = add(mul(d[0], a[1]), mul(a[0], d[1]))
        elif op == "div":
            # d(a/b) = da/b - (a/b) db / b
            r = sub(div(d[0], a[1]), mul(div(e, a[1]), d[1]))
        elif op == "pow":
            # a^b: b a^(b-1) da + a^b log(a) db
            t1 = mul(mul(a[1], power(a[0], sub(a[1], ONE))), d[0])
            t2 = ZERO if (d[1].is_const and d[1].value == 0.0) else mul(mul(e, unary("log", a[0])), d[1])
            r = add(t1, t2)
        elif op == "sin":
            r = mul(unary("cos", a[0]), d[0])
        elif op == "cos":
            r = neg(mul(unary("sin", a[0]), d[0]))
        elif op == "tan":
            r = mul(add(ONE, mul(e, e)), d[0])
        elif op == "exp":
            r = mul(e, d[0])
        elif op == "expm1":
            r = mul(add(e, ONE), d[0])
        elif op == "log":
            r = div(d[0], a[0])
        elif op == "log1p":
            r = div(d[0], add(ONE, a[0]))
        elif op == "sqrt":
            r = div(d[0], mul(TWO, e))
        elif op == "rsqrt":
            r = mul(mul(const(-0.5), mul(e, mul(e, e))), d[0])       # -1/2 s^(-3/2): no division
        elif op == "tanh":
            r = mul(sub(ONE, mul(e, e)), d[0])
        elif op == "sinh":
            r = mul(unary("cosh", a[0]), d[0])
        elif op == "cosh":
            r = mul(unary("sinh", a[0]), d[0])
        elif op == "atan":
            r = div(d[0], add(ONE, mul(a[0], a[0])))
        elif op == "asin":
            r = div(d[0], unary("sqrt", sub(ONE, mul(a[0], a[0]))))
        elif op == "acos":
            r = neg(div(d[0], unary("sqrt", sub(ONE, mul(a[0], a[0])))))
        elif op == "erf":
            r = mul(mul(const(2.0 / math.sqrt(math.pi)), unary("exp", neg(mul(a[0], a[0])))), d[0])
        elif op == "abs":
            r = mul(unary("sign", a[0]), d[0])
        elif op in ("sign", "floor", "trunc") or op in BOOL_OPS:
            r = ZERO
        elif op == "max":
            r = where(binary("ge", a[0], a[1]), d[0], d[1])
        elif op == "min":
            r = where(binary("le", a[0], a[1]), d[0], d[1])
        elif op == "fmod":
            r = sub(d[0], mul(unary("trunc", div(a[0], a[1])), d[1]))
        elif op == "atan2":
            den = add(mul(a[0], a[0]), mul(a[1], a[1]))
            r = div(sub(mul(a[1], d[0]), mul(a[0], d[1])), den)
        elif op == "where":
            r = where(a[0], _diff(a[1], v, memo), _diff(a[2], v, memo))
        else:
            raise NotImplementedError(f"derivative of {op}")
    memo[key] = r
    return r


# ---------------------------------------------------------------------------------------
# substitution and numeric evaluation
# ---------------------------------------------------------------------------------------
def substitute(e, mapping, memo=None):
    """Replace var(i) by mapping[i] (Expr) throughout e."""
    if memo is None:
        memo = {}
    return _subst(wrap(e), mapping, memo)


def _subst(e, mapping, memo):
    r = memo.get(e.id)
    if r is not None:
        return r
    if e.op == "var":
        r = wrap(mapping[e.value]) if e.value in mapping else e
    elif e.op == "const":
        r = e
    else:
        args = [_subst(a, mapping, memo) for a in e.args]
        r = rebuild(e.op, args)
    memo[e.id] = r
    return r


def rebuild(op, args):
    if op == "add":
        return add(*args)
    if op == "sub":
        return sub(*args)
    if op == "mul":
        return mul(*args)
    if op == "div":
        return div(*args)
    if op == "neg":
        return neg(*args)
    if op == "pow":
        return power(*args)
    if op == "where":
        return where(*args)
    if op in UNARY_FUNCS:
        return unary(op, *args)
    return binary(op, *args)


def topo_order(roots):
    """Nodes reachable from roots, children before parents (iterative DFS)."""
    seen = set()
    order = []
    stack = [(r, False) for r in reversed(list(roots))]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if node.id in seen:
            continue
        seen.add(node.id)
        stack.append((node, True))
        for a in reversed(node.args):
            if a.id not in seen:
                stack.append((a, False))
    return order


_NP_FUNCS = {
    "neg": np.negative, "sin": np.sin, "cos": np.cos, "tan": np.tan, "exp": np.exp, "log": np.log,
    "sqrt": np.sqrt, "tanh": np.tanh, "sinh": np.sinh, "cosh": np.cosh, "atan": np.arctan,
    "asin": np.arcsin, "acos": np.arccos, "abs": np.abs, "sign": np.sign, "log1p": np.log1p,
    "expm1": np.expm1, "floor": np.floor, "trunc": np.trunc, "not": np.logical_not,
    "rsqrt": lambda a: 1.0 / np.sqrt(a),
    "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
    "max": np.maximum, "min": np.minimum, "fmod": np.fmod, "atan2": np.arctan2,
    "lt": np.less, "le": np.less_equal, "gt": np.greater, "ge": np.greater_equal,
    "eq": np.equal, "ne": np.not_equal, "and": np.logical_and, "or": np.logical_or,
}


def evaluate(roots, x, dtype=np.float64):
    """Evaluate root expressions at x (shape (..., n_vars)); returns list of arrays shaped x.shape[:-1].

    Host-side interpreter used by shape inference, the L-operator helpers and the tests;
    it is not a solver path."""
    from math import erf
    x = np.asarray(x, dtype=dtype)
    vals = {}
    with np.errstate(all="ignore"):
        for node in topo_order(roots):
            if node.op == "const":
                v = np.asarray(node.value if isinstance(node.value, bool) else dtype(node.value))
            elif node.op == "var":
                v = x[..., node.value]
            elif node.op == "where":
                c, a, b = (vals[k.id] for k in node.args)
                v = np.where(c, a, b)
            elif node.op == "erf":
                v = np.vectorize(erf)(vals[node.args[0].id])
            else:
                v = _NP_FUNCS[node.op](*(vals[k.id] for k in node.args))
            vals[node.id] = v
    shape = x.shape[:-1]
    return [np.broadcast_to(vals[r.id], shape) for r in roots]
