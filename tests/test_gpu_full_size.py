"""Size-independent properties checked at (or near) BASELINE's full sizes, where the oracle cannot follow:

* C2 (drifted polar walk, fp64 Wagner-Platen, 10^7 trajectories): the run is a pure function of (seed, global
  trajectory index) - ragged shards of the global fan-out concatenate to the full run bit for bit and a repeat is
  bit-identical; in cartesian coordinates the process is a Brownian motion plus drift, so cart(x_T) - cart(x0) -
  drift*T must equal the driving Wiener path w_T up to the scheme's strong error (reference
  examples_from_paper/polar_random_walk.py:20-23, tests/test_sde_solver.py:50-63); the device-side moments equal the
  moments of the stored trajectories.
* C1 (GBM, 2^20 trajectories, every step stored): the exact solution x0 exp((mu - sigma^2/2) t + sigma w_t) evaluated on
  the STORED Wiener path bounds the pathwise error of every snapshot (reference tests/test_sde_solver.py:14-47), and
  time_values is the same arithmetic progression for every trajectory.
* C5 (GARCH, moments only): sharded moments add up to the unsharded moments.
* C3 (two RPY spheres on a spring, 10^6 trajectories): the separation reaches its Boltzmann law r^2 exp(-U(r)) - the
  fluctuation-dissipation balance of drift = mu F and noise = sqrt(2) chol(mu); shards equal the full run.
(C4 at its full size: tests/test_gpu_beads.py.)
"""
import numpy as np
import pytest
import torch

import stochastic_b200 as pychastic
from stochastic_b200 import workloads

pytestmark = pytest.mark.gpu


def _t(a):
    return a.torch() if hasattr(a, "torch") else torch.as_tensor(np.asarray(a), device="cuda")


def test_c2_full_size_shards_determinism_and_brownian_identity():
    n, e = 10 ** 7, 5
    y_drift, x0 = 4.0, (2.0, 0.0)
    prob = workloads.polar_walk(y_drift, x0, 1.0, f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2.0 ** -e)
    obs = workloads.polar_error_observables(y_drift, x0[0])
    full = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, observables=obs)
    X, W, T = _t(full["solution_values"]), _t(full["wiener_values"]), _t(full["time_values"])
    assert X.shape == (n, 1, 2) and W.shape == (n, 1, 2) and T.shape == (n, 1)
    assert bool((T == 1.0).all())
    # determinism: the same call again
    again = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, observables=obs)
    assert torch.equal(_t(again["solution_values"]), X) and torch.equal(_t(again["wiener_values"]), W)
    assert np.array_equal(np.asarray(again["moments"])[:, :], np.asarray(full["moments"])) or \
        np.allclose(np.asarray(again["moments"]), np.asarray(full["moments"]), rtol=1e-12)   # atomics: order-free sums
    # ragged shards of the global fan-out (what ranks of a sharded run integrate)
    cuts = [0, 1, 3_333_333, 3_333_334 + 4_999_999, n]
    for b, c in zip(cuts[:-1], cuts[1:]):
        part = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, traj_range=(b, c - b))
        assert torch.equal(_t(part["solution_values"]), X[b:c]), (b, c)
        assert torch.equal(_t(part["wiener_values"]), W[b:c]), (b, c)
    # cartesian identity: (r cos phi, r sin phi) = x0_cart + (0, y_drift) t + w_t
    r, phi = X[:, 0, 0], X[:, 0, 1]
    ex = r * torch.cos(phi) - (x0[0] + W[:, 0, 0])
    ey = r * torch.sin(phi) - (y_drift * 1.0 + W[:, 0, 1])
    err = torch.sqrt(ex * ex + ey * ey)
    finite = torch.isfinite(err)
    assert float(finite.double().mean()) > 0.9999            # a handful of paths step across r = 0 (as in the reference)
    med = float(err[finite].median())
    assert med < 0.012, med                                   # strong order 1.5: ~7e-3 at dt = 2^-5 (profiles/README.md)
    # the Wiener path itself: mean 0, variance T per component
    assert abs(float(W.mean())) < 5 / np.sqrt(2 * n) and abs(float(W.var()) - 1.0) < 5 * np.sqrt(2.0 / (2 * n))
    # device-side moments = moments of the stored states (observable 0 is the strong error, 1 is phi_T)
    mom, cnt = np.asarray(full["moments"]), full["count"]
    assert cnt == n
    phi_c = torch.fmod(phi + np.pi, 2 * np.pi) - np.pi        # the observable's canonical angle
    assert abs(mom[1, 0] / n - float(phi_c.mean())) < 1e-9
    assert abs(mom[1, 1] / n - float((phi_c ** 2).mean())) < 1e-9
    assert abs(mom[1, 0] / n - 1.10714532096375) < 5 * float(phi_c.std()) / np.sqrt(n) + 2e-3   # atan(2): weak error
    # strong-error observable: same quantity as computed above from the stored states
    assert abs(mom[0, 0] / n - float(err.mean())) <= 1e-9 * max(1.0, abs(float(err.mean()))) or not bool(finite.all())


def test_c1_full_size_snapshots_against_the_exact_solution_on_the_stored_path():
    n, steps = 2 ** 20, 512
    mu, sigma = 0.2, 0.5
    prob = workloads.gbm(f64=True)
    # mean relative pathwise error at T = 2 with dt = 2^-8 (oracle, 3 000 paths): 1.25e-2, 5.3e-4, 1.0e-5
    for scheme, bound in (("euler", 1.4e-2), ("milstein", 6e-4), ("wagner_platen", 1.2e-5)):
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -8)
        sol = solver.solve_many(prob, n_trajectories=n, seed=1, chunk_size=1)
        X, W, T = _t(sol["solution_values"])[:, :, 0], _t(sol["wiener_values"])[:, :, 0], _t(sol["time_values"])
        assert X.shape == (n, steps)
        assert torch.equal(T, T[:1].expand_as(T))
        assert torch.allclose(T[0], torch.arange(1, steps + 1, device=T.device, dtype=T.dtype) * 2.0 ** -8, rtol=1e-12)
        exact = torch.exp((mu - 0.5 * sigma * sigma) * T + sigma * W)
        rel = ((X - exact).abs() / exact)
        # mean pathwise error at the last snapshot (T = 2): strong order 1/2, 1, 3/2
        m = float(rel[:, -1].mean())
        assert m < bound, (scheme, m)
        # increments of the stored Wiener path are independent N(0, dt)
        dW = W[:, 1:] - W[:, :-1]
        assert abs(float(dW.var()) / 2.0 ** -8 - 1.0) < 1e-3
        del sol, X, W, T, exact, rel, dW
        torch.cuda.empty_cache()


def test_c5_sharded_moments_add_up():
    n = 4 * 10 ** 6
    prob = workloads.garch(f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=0.01)
    obs = workloads.call_payoff(120.0)
    full = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, observables=obs, store_trajectories=False)
    parts = [solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, observables=obs,
                               store_trajectories=False, traj_range=(b, c))
             for b, c in ((0, 1_000_001), (1_000_001, n - 1_000_001))]
    tot = sum(np.asarray(p["moments"]) for p in parts)
    assert sum(p["count"] for p in parts) == n == full["count"]
    assert np.allclose(tot, np.asarray(full["moments"]), rtol=1e-11)
    price = np.asarray(full["moments"])[0, 0] / n
    assert 5.9 < price < 6.2                                  # 6.0282 +- 0.0005 at 10^9 paths (profiles/README.md)


def test_c3_full_size_spring_length_is_boltzmann_distributed():
    """BASELINE config C3 (two RPY spheres of radii 1 and 0.2 on the spring U = (r - 4)^2, Euler dt = 0.01, 10^6 trajectories;
    reference examples/example_two_interacting_brownian_spheres.py:17-37), run for 1000 steps - twice the 500 of the config, six
    relaxation times of the spring - so that the separation has reached its equilibrium law p(r) ~ r^2 exp(-U(r)) whatever
    the (configuration dependent) mobility is: that is the fluctuation-dissipation balance between drift = mu F and
    noise = sqrt(2) chol(mu) the whole kernel exists to keep.  Also: shards of the global fan-out equal the full run."""
    n = 10 ** 6
    prob = workloads.two_spheres(tmax=10.0, f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=0.01)
    sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None)
    assert solver.last_steps == 1000
    x = _t(sol["solution_values"])[:, -1, :]
    assert x.shape == (n, 6) and bool(torch.isfinite(x).all())
    r = torch.linalg.norm(x[:, 3:] - x[:, :3], dim=1)
    grid = np.linspace(0.0, 12.0, 120001)
    p = grid ** 2 * np.exp(-(grid - 4.0) ** 2)
    p /= np.trapezoid(p, grid)
    mean = np.trapezoid(grid * p, grid)
    var = np.trapezoid((grid - mean) ** 2 * p, grid)
    assert abs(float(r.mean()) - mean) < 0.01 * mean, (float(r.mean()), mean)          # 4.24; sampling error 0.0007
    assert abs(float(r.var()) / var - 1.0) < 0.03, (float(r.var()), var)                # Euler bias ~ mu_rel U'' dt / 2 = 0.3 %
    # the spheres diffuse with very different mobilities: the small one has moved much further
    msd_big, msd_small = float((x[:, :3] ** 2).sum(1).mean()), float(((x[:, 3:] - torch.tensor([0.0, 0.0, 4.0], device=x.device)) ** 2).sum(1).mean())
    assert msd_small > 2.0 * msd_big > 0
    part = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, traj_range=(333_333, 1001))
    assert torch.equal(_t(part["solution_values"])[:, -1, :], x[333_333:334_334])
