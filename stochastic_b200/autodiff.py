"""`jax.grad / jacobian / hessian`-shaped helpers on top of the symbolic tracer.

User problems in the reference differentiate energies inside the drift
(`jax.grad(u_ene)(x)`, reference `examples/example_two_interacting_brownian_spheres.py:20`)
and the solver differentiates drift/noise (`jax.jacobian`, `jax.hessian`, reference
`pychastic/sde_solver.py:33-34,43`).  These helpers do the same by tracing `f` on fresh
symbols, differentiating the DAG and substituting the actual argument back, so they compose
(grad inside a traced drift, grad of grad, ...).  On concrete input the result is evaluated on
the host with numpy.
"""
import itertools

import numpy as np

from . import jnp
from .trace import expr as E

_FRESH = itertools.count(1)


def _fresh_symbols(shape):
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    start = (1 << 20) * next(_FRESH)
    return jnp.symbols(shape, start=start), start, n


def _derivative_array(f, x, order):
    """d^order f / dx^order with shape f_shape + x_shape * order."""
    traced = jnp._is_traced(x)
    xin = np.asarray(x if traced else jnp._host(x), dtype=object if traced else np.float64)
    s, start, n = _fresh_symbols(xin.shape)
    y = f(s)
    y_arr = np.asarray(jnp._as_obj(y), dtype=object)
    y_shape = y_arr.shape
    cur = [E.wrap(v) for v in y_arr.ravel()]
    shape = y_shape
    for _ in range(order):
        memo = [dict() for _ in range(n)]
        cur = [E.diff(e, start + k, memo[k]) for e in cur for k in range(n)]
        shape = shape + xin.shape
    if traced:
        mapping = {start + k: E.wrap(v) for k, v in enumerate(xin.ravel())}
        memo = {}
        out = np.empty(len(cur), dtype=object)
        for i, e in enumerate(cur):
            out[i] = E.substitute(e, mapping, memo)
        out = out.reshape(shape).view(jnp.TArray)
        return out if out.ndim else out.item()
    # concrete: shift variable indices to 0..n-1 and evaluate
    mapping = {start + k: E.var(k) for k in range(n)}
    memo = {}
    roots = [E.substitute(e, mapping, memo) for e in cur]
    vals = E.evaluate(roots, xin.ravel().astype(np.float64)) if roots else []
    out = np.array([float(v) for v in vals], dtype=np.float64).reshape(shape)
    return out if out.ndim else float(out)


def jacobian(f):
    return lambda x: _derivative_array(f, x, 1)


jacfwd = jacobian
jacrev = jacobian


def hessian(f):
    return lambda x: _derivative_array(f, x, 2)


def grad(f):
    """Gradient of a scalar-valued f (same shape as x)."""
    return lambda x: _derivative_array(f, x, 1)
