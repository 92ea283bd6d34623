"""The handful of `jax.lax` control-flow helpers that the reference's examples use inside drift / noise /
step_post_processing (`jax.lax.cond`, reference `examples/examples_from_paper/brownian_rotations.py:24,38,44-53,79-86`).
On traced input both branches are evaluated and joined with a select (exactly what XLA does after vmap)."""
from . import jnp


def cond(pred, true_fun, false_fun, *operands):
    if not jnp._is_traced(pred) and not any(jnp._is_traced(o) for o in operands):
        import numpy as np
        if np.ndim(jnp._host(pred)) == 0:
            return true_fun(*operands) if bool(jnp._host(pred)) else false_fun(*operands)
    return jnp.where(pred, true_fun(*operands), false_fun(*operands))


def select(pred, on_true, on_false):
    return jnp.where(pred, on_true, on_false)


def max(x, y):
    return jnp.maximum(x, y)


def min(x, y):
    return jnp.minimum(x, y)


def clamp(lo, x, hi):
    return jnp.minimum(jnp.maximum(x, lo), hi)


def stop_gradient(x):
    return x
