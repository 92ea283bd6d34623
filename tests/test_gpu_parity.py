"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeds.

Bars: bit-exact for uint32 key material; normals within a few ulp (the oracle's numpy
log1p / the device's log1p differ in the last place, as XLA's would); trajectories within
1e-9 relative in fp64 (north_star) and 2e-5 absolute in fp32 (the tolerance at which the
oracle itself reproduces the reference's golden CSVs).
"""
import numpy as np
import pytest
import torch

import stochastic_b200 as pychastic
from oracle import jax_prng, solver as osolver, wiener_integrals as owi
from stochastic_b200 import runtime
from stochastic_b200.vectorized_I_generation import get_wiener_integrals

import problems as P

pytestmark = pytest.mark.gpu

F64_RTOL = 1e-9
F32_ATOL = 2e-5


def _keys_split(seed, n_total, begin, count, layout):
    out = torch.empty((count, 2), dtype=torch.int32, device="cuda")
    ka, kb = (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF
    runtime.check(runtime.lib().sdeb200_keys_split(ka, kb, n_total, begin, count, layout, out.data_ptr(),
                                                   runtime.current_stream_ptr()), "keys_split")
    return out.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("layout", ["original", "partitionable"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 2048, 10007])
def test_keys_split_bit_exact(n, layout):
    for seed in (0, 2, 1234567891234):
        ref = jax_prng.split(jax_prng.PRNGKey(seed, x64=True), n, layout)
        got = _keys_split(seed, n, 0, n, 0 if layout == "original" else 1)
        assert np.array_equal(got, ref)
    # a slice of the fan-out (what a rank of a sharded run derives) equals the slice of the full split
    if n > 4:
        got = _keys_split(0, n, 3, n - 4, 0 if layout == "original" else 1)
        assert np.array_equal(got, jax_prng.split(jax_prng.PRNGKey(0), n, layout)[3:n - 1])


@pytest.mark.parametrize("layout", ["original", "partitionable"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 4, 1001, 65536])
def test_normal_stream(n, dtype, layout):
    key = jax_prng.PRNGKey(7)
    out = torch.empty(n, dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda")
    runtime.check(runtime.lib().sdeb200_normal(int(key[0]), int(key[1]), n, int(dtype == np.float64),
                                               0 if layout == "original" else 1, out.data_ptr(),
                                               runtime.current_stream_ptr()), "normal")
    ref = jax_prng.normal(key, (n,), dtype, layout)
    got = out.cpu().numpy()
    eps = np.finfo(dtype).eps
    assert np.all(np.abs(got - ref) <= 8 * eps * np.maximum(np.abs(ref), 0.25))


def test_fp64_normals_from_jax_pinned_bits():
    """The three 64-bit words jax's own regression test pins for key 1701 (tests/test_oracle_golden.py), mapped to
    uniforms and normals on the host, against the device stream for the same key."""
    bits = np.array([3982329540505020460, 16822122385914693683, 7882654074788531506], dtype=np.uint64)
    lo = np.nextafter(-1.0, 0.0)
    f = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    u = np.maximum(lo, f * 2.0 + lo)
    from scipy.special import erfinv
    ref = np.sqrt(2.0) * erfinv(u)
    key = jax_prng.PRNGKey(1701)
    out = torch.empty(3, dtype=torch.float64, device="cuda")
    runtime.check(runtime.lib().sdeb200_normal(int(key[0]), int(key[1]), 3, 1, 0, out.data_ptr(),
                                               runtime.current_stream_ptr()), "normal")
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-14, atol=1e-15)


def _cmp_integrals(got, ref, tol):
    assert set(got) == set(ref)
    for k in ref:
        g = np.asarray(got[k])
        r = np.asarray(ref[k]).reshape(g.shape)
        scale = max(1.0, np.abs(r).max())
        assert np.abs(g - r).max() <= tol * scale, (k, np.abs(g - r).max())


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
@pytest.mark.parametrize("scheme", ["euler", "milstein", "wagner_platen"])
@pytest.mark.parametrize("m", [1, 2, 3])
def test_wiener_integrals_match_oracle(m, scheme, dtype, tol):
    key = jax_prng.split(jax_prng.PRNGKey(0), 5)[2]
    for steps in (2, 37):
        got = get_wiener_integrals(key, steps=steps, noise_terms=m, scheme=scheme, dtype=dtype)
        ref = owi.get_wiener_integrals(key, steps=steps, noise_terms=m, scheme=scheme, dtype=dtype)
        _cmp_integrals(got, ref, tol)


def test_wiener_integrals_other_p_and_layout():
    key = jax_prng.PRNGKey(3)
    got = get_wiener_integrals(key, steps=9, noise_terms=2, scheme="wagner_platen", p=4, dtype=np.float64,
                               prng_layout="partitionable")
    ref = owi.get_wiener_integrals(key, steps=9, noise_terms=2, scheme="wagner_platen", p=4, dtype=np.float64,
                                   layout="partitionable")
    _cmp_integrals(got, ref, 1e-12)
    got = get_wiener_integrals(key, steps=5, noise_terms=2, scheme="milstein", p=25, dtype=np.float64)
    ref = owi.get_wiener_integrals(key, steps=5, noise_terms=2, scheme="milstein", p=25, dtype=np.float64)
    _cmp_integrals(got, ref, 1e-12)


def test_unknown_scheme_raises():
    with pytest.raises(NotImplementedError):
        get_wiener_integrals(0, steps=1, noise_terms=2, scheme="heun")


def _solve_both(prob, oprob, scheme, dt, n, dtype, **kw):
    if dtype == np.float64:
        P.as_f64(prob, oprob)
    else:
        oprob.x0 = np.asarray(prob.x0, dtype=np.float64)     # the float32-rounded x0 the product integrates
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
    got = solver.solve_many(prob, n_trajectories=n, progress_bar=False, **kw)
    okw = {k: v for k, v in kw.items() if k in ("seed", "chunk_size", "chunks_per_randomization")}
    ref = osolver.solve_many(oprob, scheme=scheme, dt=dt, n_trajectories=n, dtype=dtype, **okw)
    return got, ref


def _assert_close(got, ref, dtype):
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        if dtype == np.float64:
            err = np.abs(g - r) / np.maximum(np.abs(r), 1e-3)
            assert err.max() < F64_RTOL, (k, err.max())
        else:
            assert np.abs(g - r).max() < F32_ATOL * max(1.0, np.abs(r).max()), (k, np.abs(g - r).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("scheme", ["euler", "milstein", "wagner_platen"])
def test_gbm_config1_matches_oracle(scheme, dtype):
    """BASELINE config 1 (GBM 0.2x, 0.5x, x0=1, T=2, dt=2^-8), reduced to 64 trajectories for the oracle."""
    prob, oprob = P.gbm()
    got, ref = _solve_both(prob, oprob, scheme, 2 ** -8, 64, dtype, seed=0)
    assert np.asarray(got["solution_values"]).shape == (64, 512, 1)
    _assert_close(got, ref, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("scheme", ["euler", "milstein", "wagner_platen"])
@pytest.mark.parametrize("name", ["polar_walk", "polar_drift", "parabolic_walk", "garch", "arctan", "tanh", "sinh"])
def test_problems_match_oracle(name, scheme, dtype):
    prob, oprob = {
        "polar_walk": lambda: P.polar_walk(tmax=0.25),
        "polar_drift": lambda: P.polar_walk(4.0, (2.0, 0.0), 0.25),
        "parabolic_walk": lambda: P.parabolic_walk(0.05),
        "garch": lambda: P.garch(tmax=0.5),
        "arctan": lambda: P.arctan(tmax=0.5),
        "tanh": lambda: P.tanh_benchmark(tmax=0.5),
        "sinh": lambda: P.sinh_drift(tmax=0.5),
    }[name]()
    dt = prob.tmax / 16
    got, ref = _solve_both(prob, oprob, scheme, dt, 12, dtype, seed=1)
    _assert_close(got, ref, dtype)


@pytest.mark.parametrize("chunk_size,cpr", [(1, None), (4, None), (None, None), (3, 2), (5, 1), (1, 1)])
def test_chunking_matches_oracle(chunk_size, cpr):
    """chunk_size / chunks_per_randomization incl. the round-up that overshoots tmax (reference :217-221)."""
    prob, oprob = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    if chunk_size == 1 and cpr == 1:
        scheme = "euler"   # the reference's m>1 Milstein path squeezes the step axis when S == 1
    else:
        scheme = "wagner_platen"
    got, ref = _solve_both(prob, oprob, scheme, 1 / 23, 7, np.float64, seed=5, chunk_size=chunk_size,
                           chunks_per_randomization=cpr)
    _assert_close(got, ref, np.float64)


def test_multiple_initial_conditions():
    prob, oprob = P.gbm(1.0, 1.0, 1.0, 1.0)
    x0 = np.array([[1.0], [2.0], [0.5], [3.0]])
    prob.x0 = x0
    oprob.x0 = x0
    solver = pychastic.sde_solver.SDESolver(scheme="milstein", dt=1e-2)
    with pytest.raises(ValueError):
        solver.solve_many(prob, n_trajectories=1)
    got = solver.solve_many(prob, n_trajectories=None)
    assert np.asarray(got["solution_values"]).shape == (4, 100, 1)
    ref = osolver.solve_many(oprob, scheme="milstein", dt=1e-2, n_trajectories=None, dtype=np.float64)
    _assert_close(got, ref, np.float64)


def test_traj_range_is_a_slice_of_the_full_run():
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 0.5)
    P.as_f64(prob)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2 ** -5)
    full = solver.solve_many(prob, n_trajectories=1000, seed=3, chunk_size=None)
    part = solver.solve_many(prob, n_trajectories=1000, seed=3, chunk_size=None, traj_range=(300, 250))
    for k in ("time_values", "solution_values", "wiener_values"):
        assert np.array_equal(np.asarray(full[k])[300:550], np.asarray(part[k]))


def test_solve_strips_leading_axis():
    prob, oprob = P.gbm(1.0, 1.0, 1.0, 1.0)
    solver = pychastic.sde_solver.SDESolver(dt=1 / 128)
    sol = solver.solve(prob)
    assert np.asarray(sol["time_values"]).shape == (128,)
    assert np.asarray(sol["solution_values"]).shape == (128, 1)
    ref = osolver.solve(oprob, scheme="euler", dt=1 / 128, dtype=np.float32)
    assert np.abs(np.asarray(sol["solution_values"]) - ref["solution_values"]).max() < 1e-4


def test_step_post_processing_and_observables():
    prob, oprob = P.polar_walk(0.0, (1.0, 0.0), 1.0)
    P.as_f64(prob)
    import stochastic_b200.numpy as jnp
    solver = pychastic.sde_solver.SDESolver(scheme="milstein", dt=1 / 32)
    post = lambda x: jnp.fmod(x, 2 * jnp.pi)
    got = solver.solve_many(prob, n_trajectories=50, step_post_processing=post, seed=4,
                            observables=lambda x, w, t: jnp.array([x[0] * jnp.cos(x[1]), x[1], t]))
    ref = osolver.solve_many(oprob, scheme="milstein", dt=1 / 32, n_trajectories=50, seed=4, dtype=np.float64,
                             step_post_processing=lambda x: np.fmod(x, 2 * np.pi))
    _assert_close(got, ref, np.float64)
    xT = ref["solution_values"][:, -1]
    obs = np.stack([xT[:, 0] * np.cos(xT[:, 1]), xT[:, 1], ref["time_values"][:, -1]], axis=1)
    mom = np.asarray(got["moments"])
    assert np.allclose(mom[:, 0], obs.sum(0), rtol=1e-9, atol=1e-9)
    assert np.allclose(mom[:, 1], (obs ** 2).sum(0), rtol=1e-9, atol=1e-9)
    assert got["count"] == 50


# ---- the reference's own solver tests, run through the product API ----------------------------------------
@pytest.mark.parametrize("scheme,name,limit", [("euler", "gbm", 1.4), ("milstein", "gbm", 0.7),
                                               ("euler", "arctan", 0.09), ("milstein", "arctan", 0.008)])
def test_reference_scalar_thresholds(scheme, name, limit):
    """reference tests/test_sde_solver.py:27-47"""
    prob, _ = P.gbm(1.0, 1.0, 1.0, 1.0) if name == "gbm" else P.arctan()
    solver = pychastic.sde_solver.SDESolver(scheme=scheme)
    solver.dt = prob.tmax / 2 ** 7
    result = solver.solve(prob)
    t = np.asarray(result["time_values"]).reshape(-1, 1)
    exact = prob.exact_solution(prob.x0, t, np.asarray(result["wiener_values"]))
    assert abs((exact - np.asarray(result["solution_values"]))[-1]) < limit


@pytest.mark.parametrize("scheme,name,limit", [("euler", "polar", 1.08), ("milstein", "polar", 1.2),
                                               ("euler", "parabolic", 0.025), ("milstein", "parabolic", 0.002)])
def test_reference_vector_thresholds(scheme, name, limit):
    """reference tests/test_sde_solver.py:91-112"""
    prob, _ = P.polar_walk() if name == "polar" else P.parabolic_walk()
    solver = pychastic.sde_solver.SDESolver(scheme=scheme)
    solver.dt = prob.tmax / 2 ** 7
    result = solver.solve(prob, seed=1)
    x, w = np.asarray(result["solution_values"]), np.asarray(result["wiener_values"])
    end_error = prob.to_cartesian(x[-1]) - (w[-1] + prob.to_cartesian(np.asarray(prob.x0)))
    assert (end_error ** 2).sum() ** 0.5 < limit


def test_kp_exercise_9_3_3():
    """reference tests/test_kloeden_platen_statistics.py: Euler strong error halves with dt (order 1/2 -> here the
    published 0.280 / 0.140 / 0.076 / 0.039 +- 15 %)."""
    a, b = 1.5, 0.1
    prob, _ = P.gbm(a, b, 1.0, 1.0)
    solver = pychastic.sde_solver.SDESolver()
    for dt, target in zip([2 ** -4, 2 ** -5, 2 ** -6, 2 ** -7], [0.280, 0.140, 0.076, 0.039]):
        solver.dt = dt
        sol = solver.solve_many(prob, n_trajectories=250)
        x, t, w = (np.asarray(sol[k])[:, -1].ravel() for k in ("solution_values", "time_values", "wiener_values"))
        err = np.abs(prob.exact_solution(1.0, t, w) - x)
        assert np.isclose(err.mean(), target, rtol=0.15)


def test_strong_and_weak_orders_on_the_polar_walk():
    """Strong error of the three schemes against the exact cartesian Brownian motion and the weak error of
    E[phi_T] (reference examples/examples_from_paper/polar_random_walk.py:23: atan(2) = 1.10714532096375):
    the scheme ranking and the decay with dt must hold on 2^17 fp64 trajectories."""
    from stochastic_b200 import workloads
    prob = workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True)
    obs = workloads.polar_error_observables(4.0, 2.0)
    n = 2 ** 17
    med = {}
    for scheme in ("euler", "milstein", "wagner_platen"):
        for e in (5, 7):
            solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -e)
            sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None)
            x, w, t = (np.asarray(sol[k])[:, -1] for k in ("solution_values", "wiener_values", "time_values"))
            ex = x[:, 0] * np.cos(x[:, 1]) - (w[:, 0] + 2.0)
            ey = x[:, 0] * np.sin(x[:, 1]) - (w[:, 1] + 4.0 * t)
            med[scheme, e] = np.median(np.hypot(ex, ey))
    # measured on B200 (profiles/README.md): medians 4.1e-2 / 2.6e-2 / 8.3e-4 at dt=2^-7, slopes 0.77 / 1.01 / 1.53
    assert med["milstein", 7] < 0.8 * med["euler", 7]
    assert med["wagner_platen", 7] < 0.1 * med["milstein", 7]
    for scheme, order in (("euler", 0.5), ("milstein", 1.0), ("wagner_platen", 1.5)):
        slope = np.log2(med[scheme, 5] / med[scheme, 7]) / 2
        assert slope > 0.8 * order, (scheme, slope)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2.0 ** -6)
    sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, observables=obs)
    mom = np.asarray(sol["moments"])
    mean_phi = mom[1, 0] / n
    sigma = np.sqrt(0.5951002076847987 / n)
    assert abs(mean_phi - 1.10714532096375) < 5 * sigma + 2e-3


# ---- BASELINE configs C3 (two spheres) and the jfm four-bead chain against the oracle --------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_two_spheres_config3_matches_oracle(dtype):
    """BASELINE config 3 (reference examples/example_two_interacting_brownian_spheres.py:17-37): radii (1, 0.2),
    U = (|r1 - r2| - 4)^2, RPY mobility of unequal spheres + traced Cholesky, Euler dt = 0.01; 8 trajectories x 64 steps."""
    from stochastic_b200 import workloads
    from oracle import problems as oprob
    f64 = dtype == np.float64
    prob = workloads.two_spheres(tmax=0.64, f64=f64)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=0.01)
    got = solver.solve_many(prob, n_trajectories=8, seed=3)
    assert np.asarray(got["solution_values"]).shape == (8, 64, 6)
    ref = osolver.solve_many(oprob.two_spheres(tmax=0.64), scheme="euler", dt=0.01, n_trajectories=8, seed=3, dtype=dtype)
    _assert_close(got, ref, dtype)


def test_four_beads_matches_oracle_fp64():
    """reference examples/jfm.py:16-61 (radii 3,1,1,1; three springs; far-field AND overlapping RPY branches)."""
    from stochastic_b200 import workloads
    from oracle import problems as oprob
    prob = workloads.four_beads(tmax=3.2, f64=True)
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=0.05)
    got = solver.solve_many(prob, n_trajectories=8, seed=2, chunk_size=4)
    ref = osolver.solve_many(oprob.jfm_four_beads(tmax=3.2), scheme="euler", dt=0.05, n_trajectories=8, seed=2,
                             chunk_size=4, dtype=np.float64)
    _assert_close(got, ref, np.float64)


@pytest.mark.parametrize("scheme,dt,steps", [("euler", 0.05, 64), ("milstein", 0.05, 16)])
def test_library_routines_with_and_without_reciprocal_square_roots(scheme, dt, steps):
    """config "library_rsqrt" (the traced Cholesky and grpy.muTT from one rsqrt per pivot / pair instead of sqrt + divisions,
    default on in fp64): both settings are the same function within rounding and both match the oracle."""
    from stochastic_b200 import workloads
    from oracle import problems as oprob
    ref = osolver.solve_many(oprob.jfm_four_beads(tmax=dt * steps), scheme=scheme, dt=dt, n_trajectories=6, seed=5,
                             chunk_size=4, dtype=np.float64)
    out = {}
    old = pychastic.config.get("library_rsqrt")
    try:
        for flag in (True, False):
            pychastic.config.update("library_rsqrt", flag)
            prob = workloads.four_beads(tmax=dt * steps, f64=True)
            solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
            out[flag] = solver.solve_many(prob, n_trajectories=6, seed=5, chunk_size=4)
            _assert_close(out[flag], ref, np.float64)
    finally:
        pychastic.config.update("library_rsqrt", old)
    a, b = np.asarray(out[True]["solution_values"]), np.asarray(out[False]["solution_values"])
    assert 0 < np.abs(a - b).max() <= 1e-12 * np.abs(b).max()           # different instructions, same numbers


def test_c2_wagner_platen_mid_size_with_erf_inv_tail():
    """4096 trajectories x 64 Wagner-Platen steps of BASELINE config 2 against the oracle (the small cases above draw
    ~9 k normals and reach the rare branch of erf_inv only by chance).  12 M draws here: the test first PROVES from the
    oracle's own Wiener path that increments beyond the central polynomial (w = -log(1 - u^2) >= 6.25, i.e.
    |xi| >= 3.29...) occur, then demands 1e-9 agreement on every trajectory."""
    from scipy.special import erfinv
    prob, oprob = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    n, steps = 4096, 64
    got, ref = _solve_both(prob, oprob, "wagner_platen", 1.0 / steps, n, np.float64, seed=0)
    z_tail = np.sqrt(2.0) * erfinv(np.sqrt(1.0 - np.exp(-6.25)))
    dw = np.diff(np.concatenate([np.zeros((n, 1, 2)), ref["wiener_values"]], axis=1), axis=1) * np.sqrt(steps)
    n_tail = int((np.abs(dw) >= z_tail).sum())
    assert n_tail >= 100, n_tail            # ~1e-3 of the 524 288 xi draws alone
    assert (np.abs(dw) >= np.sqrt(2.0) * erfinv(np.sqrt(1.0 - np.exp(-16.0)))).sum() >= 0   # (the w >= 16 branch is ~1e-7)
    # compare where the reference solution is regular; a trajectory that steps across r = 0 is ill-conditioned
    # (1/r drift) in BOTH implementations and amplifies last-bit differences
    g, r = np.asarray(got["solution_values"]), ref["solution_values"]
    regular = np.abs(r[:, :, 0]).min(axis=1) > 1e-2
    assert regular.mean() > 0.99
    err = np.abs(g - r)[regular] / np.maximum(np.abs(r[regular]), 1e-3)
    assert err.max() < F64_RTOL, err.max()
    assert np.abs(np.asarray(got["wiener_values"]) - ref["wiener_values"]).max() < 1e-12
    assert np.array_equal(np.asarray(got["time_values"]), ref["time_values"])
