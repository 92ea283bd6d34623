"""The oracle against everything that pins it (CPU only): Random123 / JAX known answers, the reference's
shipped golden trajectories, its own test thresholds and the analytic moments of the iterated integrals."""
import itertools
import os

import numpy as np
import pytest

from oracle import jax_prng, problems, solver, wiener_integrals as owi
from stochastic_b200 import wiener_integral_moments as wim


def test_threefry_random123_known_answers():
    cases = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
             ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
             ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, out in cases:
        y0, y1 = jax_prng.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(y0), int(y1)) == out


def test_jax_key_known_answers():
    K0 = jax_prng.PRNGKey(0)
    assert jax_prng.split(K0, 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert jax_prng.split(K0, 3).tolist() == [[2467461003, 428148500], [3186719485, 3840466878],
                                              [2562233961, 1946702221]]
    assert jax_prng.split(jax_prng.PRNGKey(2), 1).tolist() == [[637334850, 3278974502]]
    assert jax_prng.fold_in(K0, 1).tolist() == [928981903, 3453687069]
    assert jax_prng.random_bits(K0, 32, (4,)).tolist() == [4146024105, 967050713, 2718843009, 1272950319]
    assert jax_prng.random_bits(K0, 32, (3,)).tolist() == [4146024105, 1351547692, 2718843009]
    n = jax_prng.normal(K0, (3,), np.float32)
    assert np.allclose(n, [1.8160863, -0.48262316, 0.33988908], atol=2e-7)
    # jax tests/random_test.py::testPRNGValues
    assert jax_prng.split(K0, 4).tolist() == [[2285895361, 1501764800], [1518642379, 4090693311],
                                              [433833334, 4221794875], [839183663, 3740430601]]
    assert jax_prng.fold_in(K0, 4).tolist() == [2285895361, 433833334]
    # scalar draws printed in jax's documentation (random-numbers tutorial, `jax.random` module docs)
    assert abs(float(jax_prng.uniform(K0, (1,), np.float32)[0]) - 0.41845703) < 1e-8
    assert abs(float(jax_prng.normal(K0, (1,), np.float32)[0]) - (-0.20584226)) < 2e-8
    assert abs(float(jax_prng.normal(jax_prng.PRNGKey(42), (1,), np.float32)[0]) - (-0.18471177)) < 2e-8
    chain = jax_prng.split(jax_prng.split(K0, 5)[0], 2048)[0]
    assert chain.tolist() == [3338368460, 4177929024]


def test_jax_random_bits_known_answers_32_and_64():
    """jax's own regression vectors (jax tests/random_test.py::testRngRandomBits, key 1701, threefry, x64 enabled):
    the 64-bit case is what pins the counter layout of the fp64 stream."""
    key = jax_prng.PRNGKey(1701)
    assert jax_prng.random_bits(key, 32, (3,)).tolist() == [56197195, 4200222568, 961309823]
    assert jax_prng.random_bits(key, 64, (3,)).tolist() == [3982329540505020460, 16822122385914693683,
                                                            7882654074788531506]
    # the same test's expectation with x64 disabled is the truncation of those words to 32 bits
    assert [b & 0xFFFFFFFF for b in jax_prng.random_bits(key, 64, (3,)).tolist()] == [676898860, 3164047411, 4010691890]


def test_erfinv_against_scipy():
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001)
    # Giles' polynomials: ~1e-16 in the central branch, < 1e-12 relative in the far tails (double); ~3e-6 (single)
    assert np.max(np.abs(jax_prng.erf_inv64(x) - erfinv(x)) / np.maximum(np.abs(erfinv(x)), 1e-300)) < 1e-12
    xc = np.linspace(-0.99, 0.99, 2001)
    assert np.max(np.abs(jax_prng.erf_inv64(xc) - erfinv(xc)) / np.maximum(np.abs(erfinv(xc)), 1e-300)) < 1e-14
    x32 = x.astype(np.float32)
    ref = erfinv(x32.astype(np.float64))
    assert np.max(np.abs(jax_prng.erf_inv32(x32) - ref) / np.maximum(np.abs(ref), 1.0)) < 4e-6


def test_golden_stopping_time_trajectories(golden_dir):
    """hit.py recipe: Euler, dt=2^-10, 5 trajectories, seed 0, chunk_size=1, cpr=1 (fp32)."""
    prob = problems.polar_walk(y_drift=4.0, x0=(2.0, 0.0), tmax=2.0)
    sol = solver.solve_many(prob, scheme="euler", dt=2 ** -10, n_trajectories=5, seed=0, chunk_size=1,
                            chunks_per_randomization=1, dtype=np.float32, max_fragments=480)
    for k in range(5):
        g = np.loadtxt(os.path.join(golden_dir, f"stopping_time_trajectory_{k + 1}.csv"), delimiter=",")
        assert np.abs(sol["solution_values"][k][:len(g)] - g).max() < 1e-5


def test_golden_bead_trajectory(golden_dir):
    """jfm.py recipe with the un-halved spring energy: pins RPY mobility (far + overlapping) and Cholesky."""
    prob = problems.jfm_four_beads()
    sol = solver.solve_many(prob, scheme="euler", dt=0.05, n_trajectories=1, seed=2, chunk_size=40,
                            chunks_per_randomization=1, dtype=np.float32, max_fragments=6)
    g = np.loadtxt(os.path.join(golden_dir, "bead_single_trajectory_head.csv"), delimiter=",")[:6]
    assert np.abs(sol["solution_values"][0] - g).max() < 2e-5
    tg = np.loadtxt(os.path.join(golden_dir, "bead_time_head.csv"))[:6]
    assert np.abs(sol["time_values"][0] - tg).max() < 1e-5


def test_D_matrix_forms_agree():
    """Same three fills as the reference's tests/test_higher_correlation_matrices_logic.py plus random ones."""
    m, p = 2, 3
    base = np.arange(m * p, dtype=np.float64).reshape(m, p)
    for ce, cz in [(0, 2), (2, 0), (3, 5)]:
        assert np.allclose(owi.make_D_mat(ce * base, cz * base), owi.make_D_mat_loopy(ce * base, cz * base))
    rng = np.random.default_rng(0)
    for m, p in [(2, 10), (3, 4)]:
        eta, zeta = rng.normal(size=(m, p)), rng.normal(size=(m, p))
        assert np.allclose(owi.make_D_mat(eta, zeta), owi.make_D_mat_loopy(eta, zeta), atol=1e-13)
        assert np.allclose(owi.batched_C_mat(eta[None], zeta[None])[0], owi.make_C_mat(eta, zeta), atol=1e-13)


def _moment_check(I, labels, n):
    tol = 5 * 2 ** (-n / 2)
    for (ia, va), (ib, vb) in itertools.product(labels, labels):
        assert abs(np.mean(va * vb) - wim.E2(ia, ib)(1)) < tol, (ia, ib)
    for ia, va in labels:
        assert abs(np.mean(va) - wim.E(ia)(1)) < tol, ia


def test_wiener_integral_moments_wagner_platen():
    """reference tests/test_wiener.py: WP m=2 with p=50, 2^12 samples, seed 0, tolerance 5 * 2^(-n/2)."""
    n = 12
    I = owi.get_wiener_integrals(jax_prng.PRNGKey(0), steps=2 ** n, noise_terms=2, scheme="wagner_platen", p=50,
                                 dtype=np.float32)
    labels = [([1], I["d_w"][:, 0]), ([2], I["d_w"][:, 1]), ([0], np.ones(2 ** n)),
              ([1, 0], I["d_wt"][:, 0, 0]), ([2, 0], I["d_wt"][:, 1, 0]),
              ([0, 1], I["d_tw"][:, 0, 0]), ([0, 2], I["d_tw"][:, 0, 1])]
    labels += [([a, b], I["d_ww"][:, a - 1, b - 1]) for a in (1, 2) for b in (1, 2)]
    labels += [([a, b, c], I["d_www"][:, a - 1, b - 1, c - 1]) for a in (1, 2) for b in (1, 2) for c in (1, 2)]
    _moment_check(I, [(i, v.astype(np.float64)) for i, v in labels], n)


def test_wiener_integral_moments_milstein_and_scalar():
    n = 14
    I = owi.get_wiener_integrals(jax_prng.PRNGKey(0), steps=2 ** n, noise_terms=2, scheme="milstein")
    labels = [([a, b], I["d_ww"][:, a - 1, b - 1]) for a in (1, 2) for b in (1, 2)]
    labels += [([1], I["d_w"][:, 0]), ([2], I["d_w"][:, 1]), ([0], np.ones(2 ** n))]
    _moment_check(I, [(i, v.astype(np.float64)) for i, v in labels], n)
    I = owi.get_wiener_integrals(jax_prng.PRNGKey(0), steps=2 ** n, noise_terms=1, scheme="wagner_platen")
    labels = [([1], I["d_w"][:, 0]), ([1, 1], I["d_ww"][:, 0, 0]), ([1, 0], I["d_wt"][:, 0, 0]),
              ([0, 1], I["d_tw"][:, 0, 0]), ([1, 1, 1], I["d_www"][:, 0, 0, 0]), ([0], np.ones(2 ** n))]
    _moment_check(I, [(i, v.astype(np.float64)) for i, v in labels], n)


@pytest.mark.parametrize("scheme,problem,limit", [
    ("euler", "gbm", 1.4), ("milstein", "gbm", 0.7), ("euler", "arctan", 0.09), ("milstein", "arctan", 0.008)])
def test_reference_scalar_thresholds(scheme, problem, limit):
    """reference tests/test_sde_solver.py:27-47 (seed 0, 128 steps)."""
    prob = problems.gbm(1.0, 1.0, 1.0, 1.0) if problem == "gbm" else problems.arctan()
    sol = solver.solve(prob, scheme=scheme, dt=1 / 128, seed=0, dtype=np.float32)
    t, x, w = sol["time_values"][-1], sol["solution_values"][-1, 0], sol["wiener_values"][-1, 0]
    exact = np.exp(0.5 * t + w) if problem == "gbm" else np.arctan(w)
    assert abs(exact - x) < limit


@pytest.mark.parametrize("scheme,problem,limit", [
    ("euler", "polar", 1.08), ("milstein", "polar", 1.2), ("euler", "parabolic", 0.025), ("milstein", "parabolic", 0.002)])
def test_reference_vector_thresholds(scheme, problem, limit):
    """reference tests/test_sde_solver.py:91-112 (seed 1, 128 steps)."""
    if problem == "polar":
        prob = problems.polar_walk()
        cart = lambda x: np.array([x[0] * np.cos(x[1]), x[0] * np.sin(x[1])])
    else:
        prob = problems.parabolic_walk()
        cart = lambda x: np.array([x[0] ** 2 - x[1] ** 2, x[0] + x[1]])
    sol = solver.solve(prob, scheme=scheme, dt=prob.tmax / 128, seed=1, dtype=np.float32)
    err = cart(sol["solution_values"][-1]) - (sol["wiener_values"][-1] + cart(prob.x0))
    assert np.sqrt((err ** 2).sum()) < limit


def test_kloeden_platen_exercise_9_3_3():
    """reference tests/test_kloeden_platen_statistics.py: Euler strong error 0.280/0.140/0.076/0.039 +-15%."""
    a, b = 1.5, 0.1
    prob = problems.gbm(a, b, 1.0, 1.0)
    expect = [0.280, 0.140, 0.076, 0.039]
    for dt, target in zip([2 ** -4, 2 ** -5, 2 ** -6, 2 ** -7], expect):
        sol = solver.solve_many(prob, scheme="euler", dt=dt, n_trajectories=250, seed=0, dtype=np.float32,
                                use_closed_form=True)
        x, t, w = sol["solution_values"][:, -1, 0], sol["time_values"][:, -1], sol["wiener_values"][:, -1, 0]
        err = np.abs(np.exp((a - b * b / 2) * t + b * w) - x)
        assert np.isclose(err.mean(), target, rtol=0.15)


def test_closed_form_coefficients_match_autodiff():
    prob = problems.gbm()
    kw = dict(scheme="wagner_platen", dt=2 ** -5, n_trajectories=6, seed=3, dtype=np.float64)
    a = solver.solve_many(prob, use_closed_form=False, **kw)
    b = solver.solve_many(prob, use_closed_form=True, **kw)
    assert np.allclose(a["solution_values"], b["solution_values"], rtol=1e-13)


def test_traj_slice_is_a_slice():
    prob = problems.polar_walk(4.0, (2.0, 0.0), 0.25)
    kw = dict(scheme="milstein", dt=2 ** -4, n_trajectories=10, seed=2, dtype=np.float64)
    full = solver.solve_many(prob, **kw)
    part = solver.solve_many(prob, traj_slice=(3, 4), **kw)
    assert np.array_equal(full["solution_values"][3:7], part["solution_values"])
