"""The CUDA path against NUMBERS THE REFERENCE HOLDS for schemes beyond Euler.

* The hitting-time convergence tables of its paper (`tests/golden/{strong,weak}_convergence_test.csv`, produced by
  reference `examples/hit.py:61-145`): the only reference-held data that exercise Milstein and Wagner-Platen STEPPING
  through the solver.  Running the script's own recipe (10 000 fp32 trajectories, seed 0,
  chunks_per_randomization=3, n_steps = int(1.5**k)) on the CUDA path reproduces the tables DRAW FOR DRAW: most rows
  agree to 4 digits, the rest to a few per cent (a single trajectory grazing the r = 0 singularity of the polar
  coordinates changes an RMSE by that much when the last bits of fp32 arithmetic differ from XLA's).  Decoded columns:
  strong = RMSE of (hitting time of the solution - hitting time of the exact Wiener path), its sigma column = std / 3;
  weak = |mean of that difference| (not |mean - 1/2| as the script prints).
* The analytic mean of `examples/examples_from_paper/sinh_drift.py:25`.
* The ensemble mean-square displacements and the long-time diffusion coefficient 0.2898 of `examples/jfm.py:134-221`
  (device-side per-snapshot observables; no (n_traj, n_snap, 12) array).
* The rotational-diffusion curves of `examples/published_problems/evensen_nalum_elgsaeter_macromol_2008.py`
  (`rotational/sources/curve_{1,3,5}.csv`) and their closed form.
"""
import os

import numpy as np
import pytest
import torch

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200 import workloads

import problems as P

pytestmark = pytest.mark.gpu

Y_DRIFT, Y_BARRIER = 4.0, 2.0


def hitting_time_errors(scheme, n_steps, n_samples, seed=0):
    """reference examples/hit.py:61-112 `analyze_solver`, with its post-processing (including its wrap-around
    `trajectory[hit_index - 1]` for a hit at the first snapshot) evaluated with torch on the stored snapshots."""
    prob, _ = P.polar_walk(Y_DRIFT, (2.0, 0.0), 2.0)           # fp32, like the reference
    dt = prob.tmax / n_steps
    solver = pychastic.sde_solver.SDESolver(dt=dt, scheme=scheme)
    sol = solver.solve_many(prob, n_samples, seed=seed, chunks_per_randomization=3)
    x, w, t = (sol[k].torch() for k in ("solution_values", "wiener_values", "time_values"))
    y_real = w[..., 1] + Y_DRIFT * t[0]
    x_real = w[..., 0] + 2
    x_sol = x[..., 0] * torch.cos(x[..., 1])
    y_sol = x[..., 0] * torch.sin(x[..., 1])
    both = (y_sol > Y_BARRIER).any(dim=1) & (y_real > Y_BARRIER).any(dim=1)

    def hit_index(xs, ys):
        hit = (ys > Y_BARRIER).to(torch.int8).argmax(dim=1)
        before = (hit - 1) % ys.shape[1]                        # python's negative index wraps to the last snapshot
        rows = torch.arange(ys.shape[0], device=ys.device)
        yb, ya = ys[rows, before], ys[rows, hit]
        return hit + (Y_BARRIER - yb) / (ya - yb)

    t_sol = hit_index(x_sol[both], y_sol[both]) * dt
    t_real = hit_index(x_real[both], y_real[both]) * dt
    err = (t_sol - t_real).double()
    return {"rmse": float(torch.sqrt((err ** 2).mean())), "mean": float(t_sol.double().mean()),
            "mean_real": float(t_real.double().mean()), "mean_err": float(err.mean()), "err_std": float(err.std()),
            "mean_std": float(t_sol.double().std() / np.sqrt(len(t_sol))), "n": int(both.sum())}


N_STEPS = [int(v) for v in 1.5 ** np.arange(3, 20)]                      # hit.py:116


@pytest.mark.parametrize("scheme,col", [("euler", 1), ("milstein", 3), ("wagner_platen", 5)])
def test_hitting_time_convergence_tables_are_reproduced(scheme, col, golden_dir):
    strong = np.loadtxt(os.path.join(golden_dir, "strong_convergence_test.csv"), delimiter=",")
    weak = np.loadtxt(os.path.join(golden_dir, "weak_convergence_test.csv"), delimiter=",")
    assert np.allclose(strong[:, 0], [2.0 / n for n in N_STEPS[::-1]]) and np.allclose(weak[:, 0], strong[:, 0])
    close_strong = close_weak = 0
    for n_steps in N_STEPS:
        row = len(N_STEPS) - 1 - N_STEPS.index(n_steps)
        r = hitting_time_errors(scheme, n_steps, 10_000, seed=0)          # hit.py:33,62: n_samples = 10000, seed=0
        ref_strong, ref_sigma, ref_weak = strong[row, col], strong[row, col + 1], weak[row, col]
        # Every row within 12 % / two standard errors of the reference's number.  Exception: Wagner-Platen below
        # dt = 2/129, where its error (< 5e-3) is no longer discretisation but float32 rounding noise, which differs
        # between XLA's and this kernel's evaluation order (measured: 0.0043 vs 0.0047, ..., 4e-5 vs 3.3e-4 at the
        # finest step) - there the requirement is "not worse than the reference".
        if scheme == "wagner_platen" and n_steps > 129:
            assert r["rmse"] < 1.3 * ref_strong, (scheme, n_steps, r, ref_strong)
        else:
            assert abs(r["rmse"] - ref_strong) < 0.12 * ref_strong, (scheme, n_steps, r, ref_strong)
            assert abs(r["err_std"] / 3 - ref_sigma) < 0.12 * ref_sigma
        assert abs(abs(r["mean_err"]) - ref_weak) < 2.0 * r["err_std"] / np.sqrt(r["n"]) + 1e-5, (scheme, n_steps, r, ref_weak)
        close_strong += abs(r["rmse"] - ref_strong) < 0.01 * ref_strong
        close_weak += abs(abs(r["mean_err"]) - ref_weak) < 0.02 * ref_weak + 6e-6
    # ... and most rows to the digits the table prints (same seed, same draws): measured 15 / 13 / 6 of 17 (strong)
    need = {"euler": (14, 13), "milstein": (12, 11), "wagner_platen": (6, 5)}[scheme]
    assert close_strong >= need[0] and close_weak >= need[1], (scheme, close_strong, close_weak)


def test_hitting_time_orders_from_the_device_side_first_passage():
    """The same study without storing trajectories: `first_passage` locates the crossing on the device (2 x 10^6
    trajectories); strong RMSE against the exact path's crossing falls with the orders the paper's figure shows
    (dashed guides of plot_strong_convergence.py:38-40: dt^(1/4)... for Euler, dt^(1/2) Milstein, dt^1 Wagner-Platen)."""
    n = 2_000_000
    rm = {}
    for scheme in ("euler", "milstein", "wagner_platen"):
        for n_steps in (64, 256):
            prob = workloads.polar_walk(Y_DRIFT, (2.0, 0.0), 2.0, f64=True)
            solver = pychastic.sde_solver.SDESolver(dt=prob.tmax / n_steps, scheme=scheme)
            kw = dict(n_trajectories=n, seed=0, chunk_size=None, store_trajectories=False)
            sol = solver.solve_many(prob, first_passage=lambda x, w, t: x[0] * jnp.sin(x[1]) - Y_BARRIER, **kw)
            real = solver.solve_many(prob, first_passage=lambda x, w, t: w[1] + Y_DRIFT * t - Y_BARRIER, **kw)
            both = sol["hit_flag"].torch() & real["hit_flag"].torch()
            err = (sol["hit_time"].torch() - real["hit_time"].torch())[both]
            rm[scheme, n_steps] = float(torch.sqrt((err ** 2).mean()))
            assert both.float().mean() > 0.99
            assert abs(float(real["hit_time"].torch()[both].mean()) - 0.5) < 0.03       # discrete monitoring bias only
    assert rm["wagner_platen", 256] < rm["milstein", 256] < rm["euler", 256]
    slope = {s: np.log2(rm[s, 64] / rm[s, 256]) / 2 for s in ("euler", "milstein", "wagner_platen")}
    # measured: 0.34 / 0.54 / 0.65 (the linear interpolation of the crossing limits the highest order)
    assert slope["euler"] > 0.25 and slope["milstein"] > 0.45 and slope["wagner_platen"] > 0.55, slope
    assert rm["wagner_platen", 256] < 0.5 * rm["milstein", 256] and rm["milstein", 256] < 0.6 * rm["euler", 256]


@pytest.mark.parametrize("scheme,dt_exp,tol_bias", [("milstein", 7, 0.02), ("wagner_platen", 5, 0.01), ("euler", 8, 0.03)])
def test_sinh_drift_analytic_mean(scheme, dt_exp, tol_bias):
    """reference examples/examples_from_paper/sinh_drift.py:6-25: dx = (x/2 + sqrt(x^2+1)) dt + sqrt(x^2+1) dW,
    x0 = 0, E[x_1] = 1.937579205326131, Var = 9.64532433084641; the mean is reduced on the device."""
    prob, _ = P.sinh_drift(1.0)
    P.as_f64(prob)
    n = 1 << 21
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -dt_exp)
    sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, store_trajectories=False,
                            observables=lambda x, w, t: jnp.array([x[0], jnp.sinh(w[0] + t)]))
    mom = np.asarray(sol["moments"])
    mean, var = mom[0, 0] / n, mom[0, 1] / n - (mom[0, 0] / n) ** 2
    sigma = np.sqrt(9.64532433084641 / n)
    # heavy-tailed (sinh of a Gaussian) and biased low by the discretisation: measured 9.36 / 9.22 / 9.63
    assert 9.0 < var < 10.2 and (scheme != "wagner_platen" or abs(var - 9.64532433084641) < 0.15)
    assert abs(mean - 1.937579205326131) < 4 * sigma + tol_bias, (mean, sigma)
    # the exact solution is x_t = sinh(w_t + t): same sampling noise, no discretisation bias
    assert abs(mom[1, 0] / n - 1.937579205326131) < 4 * sigma


# ---- jfm.py: ensemble mean-square displacements on the device ---------------------------------------------------
def _jfm_observables(weights):
    wts = [float(v) for v in weights]

    def msd(x, w, t, x_ref):           # reference examples/jfm.py:134-171 (displacement from the FIRST SNAPSHOT)
        d = x - x_ref
        c = [sum(wts[a] * d[3 * a + k] for a in range(4)) for k in range(3)]
        return jnp.array([d[0] ** 2 + d[1] ** 2 + d[2] ** 2, d[9] ** 2 + d[10] ** 2 + d[11] ** 2,
                          c[0] ** 2 + c[1] ** 2 + c[2] ** 2])
    return msd


def _trace_mobility(x, w, t):          # jfm.py:82-95: particle-wise trace of the RPY mobility
    from stochastic_b200 import grpy
    mu = grpy.muTT(jnp.reshape(x, (4, 3)), np.array([3.0, 1.0, 1.0, 1.0]))
    return jnp.array([mu[3 * a][3 * b] + mu[3 * a + 1][3 * b + 1] + mu[3 * a + 2][3 * b + 2]
                      for a in range(4) for b in range(4)])


def test_per_snapshot_msd_equals_msd_of_stored_trajectories_and_the_reference_csv(golden_dir):
    prob = workloads.four_beads(tmax=400.0, half_spring=True, f64=True)
    solver = pychastic.sde_solver.SDESolver(dt=0.05)
    weights = [0.7, 0.06, 0.11, 0.13]
    kw = dict(n_trajectories=3000, chunk_size=40, chunks_per_randomization=1, seed=0)
    dev = solver.solve_many(prob, observables=_jfm_observables(weights), observe_at="snapshots",
                            store_trajectories=False, **kw)
    mom = np.asarray(dev["moments"])
    assert mom.shape == (200, 3, 2) and dev["count"] == 3000
    x = np.asarray(solver.solve_many(prob, **kw)["solution_values"])           # (3000, 200, 12)
    d = x - x[:, :1]
    centre = np.einsum("a,ntak->ntk", np.array(weights), d.reshape(3000, 200, 4, 3))
    vals = np.stack([(d[:, :, 0:3] ** 2).sum(-1), (d[:, :, 9:12] ** 2).sum(-1), (centre ** 2).sum(-1)], axis=-1)
    assert np.allclose(mom[:, :, 0], vals.sum(0), rtol=1e-10, atol=1e-9)
    assert np.allclose(mom[:, :, 1], (vals ** 2).sum(0), rtol=1e-10, atol=1e-9)
    # the reference's own curves (256 trajectories: ~6 % noise per point, correlated along the curve)
    msd = mom[:, :, 0] / 3000
    for k, name in enumerate(("big_bead", "small_bead")):
        ref = np.loadtxt(os.path.join(golden_dir, f"bead_{name}_msd_head.csv"))
        assert ref[0] == 0.0 and msd[0, k] == 0.0
        assert np.abs(msd[1:, k] / ref[1:200] - 1).max() < 0.25, name


def test_four_bead_long_time_diffusion_coefficient():
    """jfm.py:203-221: pi x (slope of the diffusion-centre MSD) = 0.2898 (Cichocki et al. 2019).  Two launches of
    2^14 trajectories x 160 000 Euler steps: the first reduces the mean trace-mobility matrix at the final time (optimal
    weights, jfm.py:89-103), the second the three MSD curves at 4000 snapshots - nothing is stored per trajectory."""
    prob = workloads.four_beads(tmax=8000.0, half_spring=True, f64=True)
    solver = pychastic.sde_solver.SDESolver(dt=0.05)
    n = 1 << 14
    kw = dict(n_trajectories=n, chunk_size=40, chunks_per_randomization=1, seed=0, store_trajectories=False)
    tm = (np.asarray(solver.solve_many(prob, observables=_trace_mobility, **kw)["moments"])[:, 0] / n).reshape(4, 4)
    inv = np.linalg.inv(tm)
    weights = inv.sum(-1) / inv.sum()
    assert abs(6.0 / inv.sum() - 0.2787) < 0.004                      # minimal short-time diffusion coefficient
    mom = np.asarray(solver.solve_many(prob, observables=_jfm_observables(weights), observe_at="snapshots", **kw)["moments"])
    msd = mom[:, :, 0] / n
    t = 2.0 * np.arange(len(msd))
    slope = [np.polyfit(t[500:], msd[500:, k], 1)[0] * np.pi for k in range(3)]
    assert abs(slope[2] - 0.2898) < 0.006, slope                      # measured 0.2903
    assert abs(slope[0] - slope[2]) < 0.01 and abs(slope[1] - slope[2]) < 0.01   # every bead diffuses with the centre


# ---- rotational Brownian motion (Evensen et al. 2008) --------------------------------------------------------------
def test_rotational_diffusion_curves(golden_dir):
    """`step_post_processing` + per-snapshot observables: <delta u_k^2>(t) = 1/6 - 5/12 exp(-6 D t) + 1/4 exp(-2 D t)
    (published_problems/evensen_nalum_elgsaeter_macromol_2008.py:137-162, D = 1).  2^18 fp32 trajectories stay within
    1.5e-3 of the closed form at every one of the 250 snapshots; the reference's own curves (its paper figure,
    ~10^3 trajectories) scatter by 0.02 around it and are met to that accuracy."""
    prob, post, obs = workloads.evensen_rotations(f64=False)
    solver = pychastic.sde_solver.SDESolver(dt=0.008)
    n = 1 << 18
    sol = solver.solve_many(prob, n_trajectories=n, step_post_processing=post, chunks_per_randomization=2, seed=0,
                            store_trajectories=False, observables=obs, observe_at="snapshots")
    mom = np.asarray(sol["moments"])[:250, :, 0] / n
    t = 0.008 * np.arange(1, 251)
    theory = 1 / 6 - 5 / 12 * np.exp(-6 * t) + 1 / 4 * np.exp(-2 * t)
    assert np.abs(mom - theory[:, None]).max() < 2.5e-3
    for k, c in enumerate((1, 3, 5)):
        ref = np.loadtxt(os.path.join(golden_dir, f"rotational_curve_{c}.csv"), delimiter=",")
        assert np.allclose(ref[:, 0], t, atol=1e-4)
        assert np.abs(mom[:, k] - ref[:, 1]).max() < 0.03
