"""`SDEProblem`: validated container of an Ito SDE  dx = a(x) dt + b(x) dw.

Drop-in for the reference's `pychastic.sde_problem.SDEProblem`
(reference `pychastic/sde_problem.py:8-110`): same constructor, attributes
(`a b x0 tmax dimension noise_terms exact_solution`) and `ValueError`s.  Shapes of `a(x0)` and
`b(x0)` are inferred by tracing the callables on symbolic arrays (the reference uses
`jax.eval_shape`, `sde_problem.py:58-59`).

Deviation (documented in DESIGN.md): `x0` is stored as float32 exactly like the reference
unless `stochastic_b200.config.update("enable_x64", True)` was called, in which case it is
stored as float64.  The solver's working precision is `problem.x0.dtype`, so overwriting
`problem.x0` with a float64 array after construction (the established way to run the reference
in double precision) also selects the fp64 kernels.
"""
import numpy as np

from . import config, jnp


def _traced_shape(fn, x_shape):
    x = jnp.symbols(tuple(x_shape))
    if x.ndim == 0:
        x = x.reshape(())
    return np.shape(jnp._as_obj(fn(x)))


class SDEProblem:
    def __init__(self, a, b, x0, tmax, exact_solution=None):
        if not tmax > 0:
            raise ValueError(f"tmax has to be positive, not {tmax}")
        self.tmax = tmax

        dtype = np.float64 if config.get("enable_x64") else np.float32
        x0 = np.array(jnp._host(x0), dtype=dtype)
        x_test_shape = x0.shape[1:] if x0.ndim == 2 else x0.shape
        a_shape = _traced_shape(a, x_test_shape)
        b_shape = _traced_shape(b, x_test_shape)

        if x0.ndim == len(a_shape) == len(b_shape) == 0:
            # scalar problem: wrap into a 1-dimensional one (reference sde_problem.py:61-73)
            self.x0 = x0.reshape(1)
            self.a = lambda x: jnp.reshape(a(x), (1,))
            self.b = lambda x: jnp.reshape(b(x), (1, 1))
            self.dimension = self.noise_terms = 1
        elif x0.ndim == len(a_shape) == 1 and len(b_shape) == 2:
            if not x0.shape[0] == a_shape[0] == b_shape[0]:
                raise ValueError(f"Inconsistent shapes: {x0.shape}, {a_shape}, {b_shape}")
            self.x0, self.a, self.b = x0, a, b
            self.dimension, self.noise_terms = b_shape
        elif x0.ndim == 2 and len(a_shape) == 1 and len(b_shape) == 2:
            if not x0.shape[1] == a_shape[0] == b_shape[0]:
                raise ValueError(f"Inconsistent shapes: {x0.shape}, {a_shape}, {b_shape}")
            self.x0, self.a, self.b = x0, a, b
            self.dimension, self.noise_terms = b_shape
        else:
            raise ValueError(f"Inconsistent dimensions: {x0.shape}, {a_shape}, {b_shape}")

        if not np.issubdtype(self.x0.dtype, np.floating):
            raise ValueError(f"initial condition dtype should be float, not {self.x0.dtype}.")
        self.exact_solution = exact_solution
