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
    for regtile in (None, False):       # the register-tiled kernel (default; fp32 problems: fp32 in / out, fp64 inside) and the shared-memory one
        solver.bead_regtile = regtile
        b = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
        src = solver._bead_plan(workloads.dna_loop(**kw), f64, "original").ks.source
        assert f"SDE_BEAD_REGTILE {1 if regtile is None else 0}" in src
        assert f"SDE_BEAD_IO_F32 {1 if regtile is None and not f64 else 0}" in src
        for k in ("time_values", "solution_values", "wiener_values"):
            x, y = np.asarray(a[k]), np.asarray(b[k])
            assert x.shape == y.shape and x.dtype == dtype
            assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-12), (k, regtile, np.abs(x - y).max())


def test_bead_kernel_rejects_other_schemes_and_shards():
    prob = workloads.dna_loop(n_beads=40, tmax=3e-7, f64=True)
    with pytest.raises(NotImplementedError):
        pychastic.sde_solver.SDESolver(scheme="milstein", dt=1e-7).solve_many(prob, 2)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    full = solver.solve_many(prob, n_trajectories=6, seed=0, chunk_size=None)
    part = solver.solve_many(prob, n_trajectories=6, seed=0, chunk_size=None, traj_range=(2, 3))
    assert np.array_equal(np.asarray(full["solution_values"])[2:5], np.asarray(part["solution_values"]))


def writhed_ring(n, beam_length=1.0, amplitude=0.3):
    """Closed ring of circumference ~beam_length lifted out of its plane (non-zero writhe and twist force)."""
    k = 2 * np.pi * np.arange(n) / n
    return (beam_length / (2 * np.pi)) * np.stack([np.cos(k), np.sin(k), amplitude * np.sin(2 * k)], axis=1)


def test_dna_loop_with_twist_matches_oracle():
    """BASELINE config C4 with the full energy of reference examples/example_writhed_dna_loop.py:49-61.  The kernel
    evaluates the writhe as Van Oosterom-Strackee triangle solid angles with an analytic gradient; the oracle as
    spherical excesses differentiated by torch autograd (pywrithe itself is unavailable: parity unpinned)."""
    x0 = writhed_ring(40)
    prob = workloads.dna_loop(n_beads=40, tmax=6e-7, f64=True, twist=True, x0=x0)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    got = solver.solve_many(prob, n_trajectories=3, seed=1, chunk_size=2)
    ref = osolver.solve_many(oprob.dna_loop(40, tmax=6e-7, twist=True, x0=x0), scheme="euler", dt=1e-7,
                             n_trajectories=3, seed=1, chunk_size=2, dtype=np.float64)
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        assert np.abs(g - r).max() <= 1e-9 * max(np.abs(r).max(), 1e-12), (k, np.abs(g - r).max())
    # the twist term must actually matter in this comparison: without it the trajectories differ visibly
    plain = solver.solve_many(workloads.dna_loop(n_beads=40, tmax=6e-7, f64=True, x0=x0), n_trajectories=3, seed=1,
                              chunk_size=2)
    assert np.abs(np.asarray(plain["solution_values"]) - ref["solution_values"]).max() > 1e-7


@pytest.mark.parametrize("twist", [False, True])
def test_dna_loop_mid_size_matches_oracle(twist):
    """12 trajectories x 60 steps from a writhed ring with jittered beads (stretched / compressed springs, overlapping and
    non-overlapping RPY pairs, beads inside the steric contact range): every stored snapshot against the oracle."""
    rng = np.random.default_rng(5)
    x0 = writhed_ring(40) + 0.1 / 40 * rng.standard_normal((40, 3))
    prob = workloads.dna_loop(n_beads=40, tmax=60e-7, f64=True, twist=twist, x0=x0)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    got = solver.solve_many(prob, n_trajectories=12, seed=11, chunk_size=15)
    ref = osolver.solve_many(oprob.dna_loop(40, tmax=60e-7, twist=twist, x0=x0), scheme="euler", dt=1e-7,
                             n_trajectories=12, seed=11, chunk_size=15, dtype=np.float64)
    assert ref["solution_values"].shape == (12, 4, 120)
    assert np.abs(ref["solution_values"][:, -1] - x0.reshape(-1)).max() > 0.5 / 40      # the beads really moved
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        assert np.abs(g - r).max() <= 1e-9 * max(np.abs(r).max(), 1e-12), (k, np.abs(g - r).max())


@pytest.mark.parametrize("regtile", [None, False])
def test_strongly_overlapping_beads_match_oracle(regtile):
    """A ring 2.5 times shorter than the reference's: neighbours sit at 0.78 hydrodynamic radii and every bead overlaps its four
    nearest neighbours (the overlapping-spheres branch of RPY everywhere near the diagonal, steric contact for two neighbours
    on each side), mobility condition number 140 - the factorisation's pivots shrink to a twelfth of the diagonal."""
    kw = dict(n_beads=40, beam_length=0.4, tmax=8e-8)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=2e-8)
    solver.bead_regtile = regtile
    got = solver.solve_many(workloads.dna_loop(f64=True, **kw), n_trajectories=3, seed=2, chunk_size=2)
    ref = osolver.solve_many(oprob.dna_loop(40, beam_length=0.4, tmax=8e-8), scheme="euler", dt=2e-8, n_trajectories=3,
                             seed=2, chunk_size=2, dtype=np.float64)
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        assert np.abs(g - r).max() <= 1e-9 * max(np.abs(r).max(), 1e-12), (k, np.abs(g - r).max())


@pytest.mark.parametrize("twist", [False, True])
def test_fp32_dna_loop_on_the_register_tiled_kernel(twist):
    """The reference's default dtype on BASELINE config C4: fp32 initial condition, snapshots, time accumulator and random
    draws (the reference's own 32-bit stream), fp64 tensor-core arithmetic in between (SDE_BEAD_IO_F32).  Against the fp32
    oracle and against the all-fp32 shared-memory kernel: within fp32 rounding; the Wiener path and the times bit for bit
    ... up to the one rounding of each accumulated sum."""
    x0 = writhed_ring(40)
    prob = workloads.dna_loop(n_beads=40, tmax=8e-7, f64=False, twist=twist, x0=x0)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    got = solver.solve_many(prob, n_trajectories=4, seed=1, chunk_size=2)
    assert "SDE_BEAD_IO_F32 1" in solver._bead_plan(prob, False, "original").ks.source
    solver.bead_regtile = False
    smem = solver.solve_many(prob, n_trajectories=4, seed=1, chunk_size=2)
    ref = osolver.solve_many(oprob.dna_loop(40, tmax=8e-7, twist=twist, x0=x0), scheme="euler", dt=1e-7, n_trajectories=4,
                             seed=1, chunk_size=2, dtype=np.float32)
    for k, tol in (("time_values", 1e-6), ("wiener_values", 2e-6), ("solution_values", 2e-5)):
        g, r, m = np.asarray(got[k]), ref[k], np.asarray(smem[k])
        assert g.dtype == np.float32 and g.shape == r.shape == m.shape, (k, g.dtype, g.shape)
        scale = max(np.abs(r).max(), 1e-12)
        assert np.abs(g - r).max() <= tol * scale, (k, "oracle", np.abs(g - r).max() / scale)
        assert np.abs(g - m).max() <= tol * scale, (k, "shared-memory kernel", np.abs(g - m).max() / scale)


def test_small_rings_on_the_register_tiled_kernel():
    """5 and 6 beads (2 and 3 tile rows, mostly padding) on the register-tiled kernel against the generic traced one."""
    for nb in (5, 6):
        kw = dict(n_beads=nb, beam_length=nb / 40, tmax=2e-6, f64=True, twist=(nb == 6), x0=writhed_ring(nb, nb / 40))
        solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
        a = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
        solver.use_bead_kernel = solver.bead_regtile = True
        b = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
        for k in ("time_values", "solution_values", "wiener_values"):
            x, y = np.asarray(a[k]), np.asarray(b[k])
            assert np.abs(x - y).max() <= 1e-9 * max(np.abs(x).max(), 1e-12), (nb, k, np.abs(x - y).max())


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 5e-4)])
def test_twisted_ring_bead_kernel_equals_generic_traced_kernel(dtype, tol):
    """Six beads with the twist term: hand-written writhe gradient (bead kernel) vs the symbolic derivative of
    `stochastic_b200.writhe.writhe` through the tracer (generic kernel), same keys."""
    f64 = dtype == np.float64
    kw = dict(n_beads=6, beam_length=0.15, tmax=2e-6, f64=f64, twist=True, x0=writhed_ring(6, 0.15))
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    a = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
    solver.use_bead_kernel = True
    for regtile in (None, False):
        solver.bead_regtile = regtile
        b = solver.solve_many(workloads.dna_loop(**kw), n_trajectories=7, seed=3, chunk_size=5)
        for k in ("time_values", "solution_values", "wiener_values"):
            x, y = np.asarray(a[k]), np.asarray(b[k])
            assert x.shape == y.shape and x.dtype == y.dtype == dtype
            assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-12), (k, regtile, np.abs(x - y).max())


@pytest.mark.parametrize("twist,n_beads", [(False, 40), (True, 40), (False, 42), (True, 41), (False, 11)])
def test_register_tiled_and_shared_memory_variants_agree(twist, n_beads):
    """fp64: the register-tiled kernel (mobility and Cholesky factor in tensor-core accumulator fragments,
    fraction-free factorisation of the diagonal tiles, row solves as tensor-core products, drift from the pair table,
    noise folded into the row solves) against the shared-memory variant.  40 beads = 15 tile rows (chain warp + 7 bulk
    warps with two tile rows each); 41 / 42 beads = 16 tile rows (the last bulk warp owns one row; 42: odd / even ring for
    the once-per-pair writhe sums); 11 beads = 5 tile rows with padding."""
    prob = workloads.dna_loop(n_beads=n_beads, beam_length=n_beads / 40, tmax=1.2e-6, f64=True, twist=twist,
                              x0=writhed_ring(n_beads, n_beads / 40))
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    solver.use_bead_kernel = True
    solver.bead_regtile = False
    b = solver.solve_many(prob, n_trajectories=5, seed=4, chunk_size=4)
    assert "SDE_BEAD_REGTILE 0" in solver._bead_plan(prob, True, "original").ks.source
    solver.bead_regtile = True
    a = solver.solve_many(prob, n_trajectories=5, seed=4, chunk_size=4)
    plan = solver._bead_plan(prob, True, "original")
    nt = -(-3 * n_beads // 8)
    assert "SDE_BEAD_REGTILE 1" in plan.ks.source and plan.ks.block == 32 * (1 + nt // 2)      # chain warp + NT / 2 bulk warps
    for k in ("time_values", "solution_values", "wiener_values"):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape == ((5, 3) if k == "time_values" else (5, 3, 3 * n_beads))
        assert np.abs(x - y).max() <= 1e-11 * max(np.abs(x).max(), 1e-12), (k, np.abs(x - y).max())


@pytest.mark.parametrize("twist", [False, True])
def test_ctas_that_integrate_several_trajectories_in_a_row(twist):
    """The bead kernel launches at most two CTAs per SM and every CTA loops over trajectories: a run with more
    trajectories than CTAs must equal the same trajectories integrated in shards that fit in one wave (every CTA exactly
    one trajectory) - nothing may leak from one trajectory of a CTA into its next."""
    prob = workloads.dna_loop(n_beads=40, tmax=3.5e-7, f64=True, twist=twist, x0=writhed_ring(40))
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    n = 700                                             # 148 SMs x 2 CTAs = 296 CTAs: three trajectories for some CTAs
    full = solver.solve_many(prob, n_trajectories=n, seed=9, chunk_size=None)
    x_full = np.asarray(full["solution_values"])
    assert x_full.shape == (n, 1, 120) and np.isfinite(x_full).all()
    for begin in range(0, n, 175):
        part = solver.solve_many(prob, n_trajectories=n, seed=9, chunk_size=None, traj_range=(begin, 175))
        for k in ("time_values", "solution_values", "wiener_values"):
            assert np.array_equal(np.asarray(full[k])[begin:begin + 175], np.asarray(part[k])), (k, begin)
    # and the trajectories differ from each other (distinct keys)
    assert np.unique(x_full[:, 0, 0]).size == n


@pytest.mark.parametrize("twist", [False, True])
def test_one_step_of_1e5_rings_has_the_mobility_as_covariance(twist):
    """BASELINE config C4's trajectory count, size-independent property: one Euler step from x0 is
    x0 + a(x0) dt + sqrt(2 dt) chol(mu(x0)) xi, so over 10^5 trajectories the displacement has mean a(x0) dt and covariance
    2 dt mu(x0) (mu and a from the oracle).  Sampling errors: Wishart, E|C_hat - C|_F^2 = (|C|_F^2 + tr(C)^2) / n."""
    import torch
    n, dt = 100_000, 1e-7
    x0 = writhed_ring(40)
    prob = workloads.dna_loop(n_beads=40, tmax=1.5e-7, f64=True, twist=twist, x0=x0)
    sol = pychastic.sde_solver.SDESolver(scheme="euler", dt=dt).solve_many(prob, n_trajectories=n, seed=7, chunk_size=None)
    x1 = np.asarray(sol["solution_values"])[:, -1, :]
    assert x1.shape == (n, 120) and np.isfinite(x1).all()
    oracle = oprob.dna_loop(40, tmax=1.5e-7, twist=twist, x0=x0)
    xt = torch.as_tensor(x0.reshape(-1))
    drift, noise = oracle.a(xt).numpy(), oracle.b(xt).numpy()
    cov = dt * noise @ noise.T                                       # = 2 dt mu
    disp = x1 - x0.reshape(-1)
    mean_err = np.abs(disp.mean(0) - drift * dt) / np.sqrt(np.diag(cov) / n)
    assert mean_err.max() < 5.0, mean_err.max()                      # 120 components, 5 standard errors
    c_hat = np.cov(disp, rowvar=False)
    expected = np.sqrt((np.sum(cov ** 2) + np.trace(cov) ** 2) / n)
    err = np.linalg.norm(c_hat - cov)
    assert err < 1.5 * expected, (err, expected)
    assert err > 0.5 * expected, (err, expected)                     # (and not suspiciously exact: the draws are independent)
    # the Wiener increments returned beside the state are the xi of that step: x1 - x0 - a dt = b(x0) w
    w = np.asarray(sol["wiener_values"])[:, -1, :]
    assert np.abs(disp - drift * dt - w @ noise.T).max() < 1e-12


def test_baseline_config_c4_full_size_keeps_the_rings_intact():
    """BASELINE config C4 as stated (10^5 trajectories x 1000 steps, dt = 1e-7, final state only): every ring is finite,
    connected (bonds within 30 % of the rest length, thermal width as the stretch modulus says), distinct, and the ensemble has not drifted
    (centre of mass: pure diffusion, mean within 5 standard errors of zero)."""
    n, steps = 100_000, 1000
    prob = workloads.dna_loop(n_beads=40, tmax=(steps + 0.5) * 1e-7, f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    sol = solver.solve_many(prob, n_trajectories=n, seed=3, chunk_size=None)
    assert solver.last_steps == steps
    x = np.asarray(sol["solution_values"])[:, -1, :].reshape(n, 40, 3)
    assert np.isfinite(x).all()
    bonds = np.linalg.norm(x - np.roll(x, 1, axis=1), axis=2)
    assert 0.7 / 40 < bonds.min() and bonds.max() < 1.3 / 40, (bonds.min(), bonds.max())
    # the bonds have equilibrated (relaxation time 1 / (2 mu ks) ~ a step): Boltzmann width sqrt(kT / ks), widened by the
    # Euler step (stationary variance (kT / ks) / (1 - lambda dt / 2) with lambda dt ~ 0.5 - 1)
    width = np.sqrt(1.0 / prob.bead_ring["ks"])
    assert 0.9 * width < bonds.std() < 1.5 * width, (bonds.std(), width)
    cm = x.mean(1) - np.asarray(prob.x0).reshape(40, 3).mean(0)
    assert np.abs(cm.mean(0)).max() < 5 * cm.std(0).max() / np.sqrt(n)
    assert cm.std(0).min() > 0 and len(np.unique(x[:, 0, 0])) == n
    t = np.asarray(sol["time_values"])
    assert np.allclose(t[:, -1], steps * 1e-7, rtol=1e-12)
