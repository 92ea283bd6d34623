"""The reference's test / example problems written against the product API
(`stochastic_b200.numpy` in place of `jax.numpy`), paired with their oracle twins."""
import numpy as np

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from oracle import problems as oprob


def gbm(alpha=0.2, beta=0.5, x0=1.0, tmax=2.0):
    p = pychastic.sde_problem.SDEProblem(lambda x: alpha * x, lambda x: beta * x, x0=x0, tmax=tmax,
                                         exact_solution=lambda x0, t, w: x0 * np.exp((alpha - 0.5 * beta * beta) * t + beta * w))
    return p, oprob.gbm(alpha, beta, x0, tmax)


def arctan(a=1.0, tmax=1.0):
    p = pychastic.sde_problem.SDEProblem(
        a=lambda x: -(a ** 2) * jnp.sin(x) * jnp.cos(x) ** 3, b=lambda x: a * jnp.cos(x) ** 2, x0=0.0, tmax=tmax,
        exact_solution=lambda x0, t, w: np.arctan(a * w + np.tan(x0)))
    return p, oprob.arctan(a, 0.0, tmax)


def tanh_benchmark(tmax=5.0):
    p = pychastic.sde_problem.SDEProblem(lambda x: -jnp.tanh(x) * (1 - jnp.tanh(x) ** 2),
                                         lambda x: 1 - jnp.tanh(x) ** 2, x0=0.0, tmax=tmax)
    return p, oprob.tanh_benchmark(0.0, tmax)


def sinh_drift(tmax=1.0):
    p = pychastic.sde_problem.SDEProblem(a=lambda x: 0.5 * x + (x ** 2 + 1.0) ** 0.5, b=lambda x: (x ** 2 + 1.0) ** 0.5,
                                         x0=jnp.array(0.0), tmax=tmax)
    return p, oprob.sinh_drift(0.0, tmax)


def polar_walk(y_drift=0.0, x0=(1.0, 0.0), tmax=1.0):
    p = pychastic.sde_problem.SDEProblem(
        a=lambda x: jnp.array([1 / (2 * x[0]) + y_drift * jnp.sin(x[1]), y_drift * jnp.cos(x[1]) / x[0]]),
        b=lambda x: jnp.array([[jnp.cos(x[1]), jnp.sin(x[1])], [-jnp.sin(x[1]) / x[0], jnp.cos(x[1]) / x[0]]]),
        x0=jnp.array(list(x0)), tmax=tmax)
    p.to_cartesian = lambda x: np.array([x[0] * np.cos(x[1]), x[0] * np.sin(x[1])])
    return p, oprob.polar_walk(y_drift, x0, tmax)


def parabolic_walk(tmax=0.05):
    p = pychastic.sde_problem.SDEProblem(
        lambda x: jnp.array([(x[0] - x[1]) / (2 * (x[0] + x[1]) ** 2), (-x[0] + x[1]) / (2 * (x[0] + x[1]) ** 2)]),
        lambda x: jnp.array([[1 / (2 * x[0] + 2 * x[1]), x[1] / (x[0] + x[1])],
                             [-1 / (2 * x[0] + 2 * x[1]), x[0] / (x[0] + x[1])]]),
        x0=jnp.array([1.0, 0.0]), tmax=tmax)
    p.to_cartesian = lambda x: np.array([x[0] ** 2 - x[1] ** 2, x[0] + x[1]])
    return p, oprob.parabolic_walk((1.0, 0.0), tmax)


def garch(theta=0.1, omega=0.05, xi=0.1, mu=0.01, tmax=2.0):
    p = pychastic.sde_problem.SDEProblem(
        lambda x: jnp.array([mu, theta * (omega - x[1])]),
        lambda x: jnp.array([[jnp.sqrt(x[1]) * x[0], 0], [0, xi * x[1]]]),
        x0=[100.0, 0.05], tmax=tmax)
    return p, oprob.garch(theta, omega, xi, mu, (100.0, 0.05), tmax)


def exp_2d():
    p = pychastic.sde_problem.SDEProblem(
        a=lambda x: jnp.array([1.1 * jnp.exp(3.0 * x[0]), 1.2 * jnp.exp(5.0 * x[1])]),
        b=lambda x: jnp.array([[1.3 * jnp.exp(7.0 * x[0]), 1.4 * jnp.exp(11.0 * x[1])],
                               [1.5 * jnp.exp(13.0 * x[1]), 1.6 * jnp.exp(17.0 * x[0])]]),
        tmax=1.0, x0=jnp.array([0.0, 0.0]))
    return p, oprob.exp_2d()


def as_f64(problem, oracle_problem=None):
    """Run in double precision: overwrite x0 after construction (the reference's established practice,
    tests/test_sde_solver.py:119) with the exact double values the oracle twin holds."""
    src = problem.x0 if oracle_problem is None else np.asarray(oracle_problem.x0).reshape(np.shape(problem.x0))
    problem.x0 = np.asarray(src, dtype=np.float64)
    return problem


def coupled_3x3():
    """d = m = 3 test system with non-commutative noise (all Wagner-Platen tensors populated)."""
    p = pychastic.sde_problem.SDEProblem(
        a=lambda x: jnp.array([-0.5 * x[0] + 0.1 * x[1] * x[2], 0.3 * jnp.sin(x[0]) - 0.2 * x[1], 0.1 - 0.4 * x[2]]),
        b=lambda x: jnp.array([[0.3 + 0.1 * x[1], 0.05 * x[2], 0.0],
                               [0.1 * jnp.cos(x[0]), 0.2, 0.07 * x[0]],
                               [0.0, 0.1 * x[0] * x[1], 0.25 + 0.05 * jnp.sin(x[2])]]),
        x0=jnp.array([0.5, -0.25, 1.0]), tmax=0.5)
    import torch
    from oracle.solver import OracleProblem

    def oa(x):
        return torch.stack([-0.5 * x[0] + 0.1 * x[1] * x[2], 0.3 * torch.sin(x[0]) - 0.2 * x[1], 0.1 - 0.4 * x[2]])

    def ob(x):
        z = 0 * x[0]
        return torch.stack([torch.stack([0.3 + 0.1 * x[1], 0.05 * x[2], z]),
                            torch.stack([0.1 * torch.cos(x[0]), 0.2 + z, 0.07 * x[0]]),
                            torch.stack([z, 0.1 * x[0] * x[1], 0.25 + 0.05 * torch.sin(x[2])])])
    return p, OracleProblem(oa, ob, [0.5, -0.25, 1.0], 0.5, name="coupled_3x3")


def rectangular_2x3():
    """d = 2, m = 3 (more noise terms than dimensions)."""
    p = pychastic.sde_problem.SDEProblem(
        a=lambda x: jnp.array([0.1 * x[1], -0.2 * x[0]]),
        b=lambda x: jnp.array([[0.2, 0.1 * x[1], 0.05 * x[0] * x[1]], [0.1 * x[0], 0.3, 0.02]]),
        x0=jnp.array([1.0, 0.5]), tmax=0.5)
    import torch
    from oracle.solver import OracleProblem

    def oa(x):
        return torch.stack([0.1 * x[1], -0.2 * x[0]])

    def ob(x):
        z = 0 * x[0]
        return torch.stack([torch.stack([0.2 + z, 0.1 * x[1], 0.05 * x[0] * x[1]]),
                            torch.stack([0.1 * x[0], 0.3 + z, 0.02 + z])])
    return p, OracleProblem(oa, ob, [1.0, 0.5], 0.5, name="rectangular_2x3")
