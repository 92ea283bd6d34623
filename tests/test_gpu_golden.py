"""The CUDA path against the reference's own shipped golden trajectories (fp32, exact stream).

Recipes: SURVEY.md section 0 / reference `examples/hit.py:9-31`, `examples/jfm.py:16-78` with the
spring energy of `examples/published_problems/cichocki_rubin_niedzwiecka_szymczak_jfm_2019.py:29-33`.
"""
import os

import numpy as np
import pytest

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200 import grpy

import problems as P

pytestmark = pytest.mark.gpu


def test_stopping_time_trajectories(golden_dir):
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 2.0)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=2 ** -10)
    sol = solver.solve_many(prob, n_trajectories=5, seed=0, chunk_size=1, chunks_per_randomization=1)
    x = np.asarray(sol["solution_values"])
    assert x.dtype == np.float32
    for k in range(5):
        g = np.loadtxt(os.path.join(golden_dir, f"stopping_time_trajectory_{k + 1}.csv"), delimiter=",")
        assert np.abs(x[k, :len(g)] - g).max() < 2e-5


def four_bead_problem(tmax=8000.0):
    radii = np.array([3.0, 1.0, 1.0, 1.0])
    n = 4

    def u_ene(x):
        loc = jnp.reshape(x, (n, 3))
        e = 0.0
        for k in range(3):
            e = e + 5.5 * (jnp.sqrt(jnp.sum((loc[k] - loc[k + 1]) ** 2)) - 4.0) ** 2
        return e

    def drift(x):
        mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
        return jnp.matmul(mu, -pychastic.grad(u_ene)(x))

    def noise(x):
        mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
        return jnp.sqrt(2) * jnp.linalg.cholesky(mu)

    return pychastic.sde_problem.SDEProblem(
        drift, noise, x0=jnp.reshape(jnp.array([[-2.0, 0, 0], [2.0, 0, 0], [6.0, 0, 0], [10.0, 0, 0]]), (12,)), tmax=tmax)


def test_bead_single_trajectory(golden_dir):
    """4-bead RPY chain, Euler dt=0.05, seed=2, chunk_size=40, cpr=1 (far and overlapping RPY branches)."""
    prob = four_bead_problem()
    solver = pychastic.sde_solver.SDESolver(dt=0.05)
    sol = solver.solve_many(prob, n_trajectories=1, chunk_size=40, chunks_per_randomization=1, seed=2,
                            max_snapshots=200)
    x = np.asarray(sol["solution_values"])[0]
    g = np.loadtxt(os.path.join(golden_dir, "bead_single_trajectory_head.csv"), delimiter=",")
    err = np.abs(x - g).max(axis=1)
    assert err[0] < 1e-5          # 40 steps
    assert err[:10].max() < 1e-4  # 400 steps (fp32 chaos amplifies rounding later on)
    t = np.asarray(sol["time_values"])[0]
    tg = np.loadtxt(os.path.join(golden_dir, "bead_time_head.csv"))
    assert np.abs(t - tg).max() < 1e-3
