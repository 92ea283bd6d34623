"""The CTA-per-trajectory bead-chain kernel (BASELINE config C4 without the parity-unpinned twist term)
against the autodiff oracle and against the generic traced kernel on the same random stream."""
import numpy as np
import pytest

import stochastic_b200 as pychastic
from stochastic_b200 import workloads
from oracle import problems as oprob, solver as osolver

pytestmark = pytest.mark.gpu


def test_dna_loop_matches_oracle():
    prob = workloads.dna_loop(n_beads=40, tmax=6e-7, f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    got = solver.solve_many(prob, n_trajectories=3, seed=1, chunk_size=2)
    ref = osolver.solve_many(oprob.dna_loop(40, tmax=6e-7), scheme="euler", dt=1e-7, n_trajectories=3, seed=1,
                             chunk_size=2, dtype=np.float64)
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        assert np.abs(g - r).max() <= 1e-9 * max(np.abs(r).max(), 1e-12), (k, np.abs(g - r).max())
    assert np.asarray(got["solution_values"]).shape == (3, 3, 120)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 2e-4)])
def test_bead_kernel_equals_generic_traced_kernel(dtype, tol):
    """Same problem, same keys, two completely different kernels (hand-written forces + shared-memory Cholesky
    vs traced drift/noise with symbolic gradient and Cholesky)."""
    f64 = dtype == np.float64
    kw = dict(n_beads=5, beam_length=0.125, tmax=2e-6, f64=f64)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    a = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
    solver.use_bead_kernel = True
    b = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
    for k in ("time_values", "solution_values", "wiener_values"):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and x.dtype == dtype
        assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-12), (k, np.abs(x - y).max())


def test_bead_kernel_rejects_other_schemes_and_shards():
    prob = workloads.dna_loop(n_beads=40, tmax=3e-7, f64=True)
    with pytest.raises(NotImplementedError):
        pychastic.sde_solver.SDESolver(scheme="milstein", dt=1e-7).solve_many(prob, 2)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    full = solver.solve_many(prob, n_trajectories=6, seed=0, chunk_size=None)
    part = solver.solve_many(prob, n_trajectories=6, seed=0, chunk_size=None, traj_range=(2, 3))
    assert np.array_equal(np.asarray(full["solution_values"])[2:5], np.asarray(part["solution_values"]))
