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
        ref = jax_prng.split(jax_prng.PRNGKey(seed), n, layout)
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
