"""Host-side stand-ins for the `jax` transformations the reference's example scripts apply to RESULT arrays
(`jax.vmap(process_trajectory)(trajectories)`, `@jax.jit`, `jax.tree_util.tree_map`; reference `examples/hit.py:61`,
`examples/jfm.py:89,188`, `examples/published_problems/evensen_nalum_elgsaeter_macromol_2008.py:137`).

They exist so that such post-processing code ports by changing the import; they run plain numpy on the host, one slice
at a time, and are not a solver path (the device-side equivalents are `solve_many(observables=..., first_passage=...)`).
"""
import numpy as np

from . import jnp


def tree_map(f, tree, *rest):
    """`jax.tree_util.tree_map` for dicts, lists, tuples and leaves."""
    if isinstance(tree, dict):
        return type(tree)((k, tree_map(f, v, *(r[k] for r in rest))) for k, v in tree.items())
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *(r[i] for r in rest)) for i, v in enumerate(tree))
    return f(tree, *rest)


def jit(fun=None, **_ignored):
    """`jax.jit` on host post-processing code: the identity (usable as `@jit` and as `@jit(...)`)."""
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    """`jax.vmap` for host post-processing: calls `fun` on every slice along `in_axes` and stacks the results
    (arrays, or dicts / tuples / lists of arrays) along `out_axes`."""
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap: in_axes must match the number of positional arguments")
        host = [tree_map(lambda x: np.asarray(jnp._host(x)), a) if ax is not None else a for a, ax in zip(args, axes)]
        sizes = set()
        for a, ax in zip(host, axes):
            if ax is not None:
                tree_map(lambda x, ax=ax: sizes.add(x.shape[ax]), a)
        if len(sizes) != 1:
            raise ValueError(f"vmap: mapped axes have inconsistent sizes {sorted(sizes)}")
        n = sizes.pop()
        outs = [fun(*[tree_map(lambda x, ax=ax: np.take(x, i, axis=ax), a) if ax is not None else a
                      for a, ax in zip(host, axes)]) for i in range(n)]
        if n == 0:
            raise ValueError("vmap: cannot infer the output structure of an empty map")
        return tree_map(lambda first, *others: np.stack([np.asarray(jnp._host(first))] +
                                                        [np.asarray(jnp._host(o)) for o in others], axis=out_axes),
                        outs[0], *outs[1:])
    return mapped
