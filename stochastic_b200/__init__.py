"""stochastic_b200 - B200-native batched SDE integrator with the PyChastic API surface.

    import stochastic_b200 as pychastic
    import stochastic_b200.numpy as jnp          # in place of jax.numpy inside drift / noise

    problem = pychastic.sde_problem.SDEProblem(lambda x: 0.2 * x, lambda x: 0.5 * x, x0=1.0, tmax=2.0)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2 ** -8)
    solution = solver.solve_many(problem, n_trajectories=10_000)

The compute path is hand-written CUDA for sm_100a reached through the C ABI in
include/sdeb200.h; there is no CPU fallback.
"""
import sys as _sys

from . import config  # noqa: F401
from . import jnp  # noqa: F401
from . import lax  # noqa: F401
from . import prng  # noqa: F401
from . import autodiff  # noqa: F401
from . import sde_problem  # noqa: F401
from . import sde_solver  # noqa: F401
from . import vectorized_I_generation  # noqa: F401
from . import wiener_integral_moments  # noqa: F401
from .autodiff import grad, hessian, jacobian  # noqa: F401
from . import hostutil as tree_util  # noqa: F401  (tree_util.tree_map)
from .hostutil import jit, vmap  # noqa: F401  (host-side stand-ins for post-processing code, see hostutil.py)
from . import prng as random  # noqa: F401  (random.PRNGKey / split / fold_in, host side)

# `import stochastic_b200.numpy as jnp` mirrors `import jax.numpy as jnp`
_sys.modules[__name__ + ".numpy"] = jnp
numpy = jnp

__version__ = "0.1.0"
