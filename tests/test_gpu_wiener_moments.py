"""Means and second moments of the generated iterated integrals against their exact values - the reference's own
acceptance test for `get_wiener_integrals` (reference tests/test_wiener.py:11-330: seed 0, 2^14 samples for
Euler / Milstein / scalar Wagner-Platen, 2^12 samples and p = 50 for vector Wagner-Platen, tolerance
5 * 2^(-n/2)), run through the product API: the device generator and `wiener_integral_moments.E / E2`."""
import itertools

import numpy as np
import pytest

import stochastic_b200 as pychastic
from stochastic_b200.vectorized_I_generation import PRNGKey

pytestmark = pytest.mark.gpu

Z_SCORE_CUTOFF = 5


def _check(columns, indices, labels, samples_exponent):
    E, E2 = pychastic.wiener_integral_moments.E, pychastic.wiener_integral_moments.E2
    tol = Z_SCORE_CUTOFF * 2 ** (-samples_exponent / 2)
    columns = [np.asarray(c, dtype=np.float64) for c in columns]
    means_error = {lab: float(np.mean(c) - E(idx)(1)) for c, idx, lab in zip(columns, indices, labels)}
    assert all(abs(e) <= tol for e in means_error.values()), "Expected values incorrect\n" + str(means_error)
    products_error = {
        (la, lb): float(np.mean(ca * cb) - E2(ia, ib)(1))
        for (ca, ia, la), (cb, ib, lb) in itertools.product(list(zip(columns, indices, labels)), repeat=2)}
    bad = {k: v for k, v in products_error.items() if abs(v) > tol}
    assert not bad, "Expected products incorrect\n" + str(bad)


def _draw(scheme, steps, noise_terms, **kw):
    I = pychastic.vectorized_I_generation.get_wiener_integrals(PRNGKey(0), scheme=scheme, steps=steps,
                                                               noise_terms=noise_terms, **kw)
    return {k: np.asarray(v) for k, v in I.items()}


def test_integral_generation_euler():
    n = 14
    I = _draw("euler", 2 ** n, 2)
    assert set(I) == {"d_w"} and I["d_w"].shape == (2 ** n, 2)
    _check([I["d_w"][:, 0], I["d_w"][:, 1], np.ones(2 ** n)], [[1], [2], [0]], ["d_w1", "d_w2", "d_t"], n)


def test_integral_generation_milstein():
    n = 14
    I = _draw("milstein", 2 ** n, 2)
    assert I["d_ww"].shape == (2 ** n, 2, 2)
    cols = [I["d_ww"][:, 0, 0], I["d_ww"][:, 0, 1], I["d_ww"][:, 1, 0], I["d_ww"][:, 1, 1],
            I["d_w"][:, 0], I["d_w"][:, 1], np.ones(2 ** n)]
    _check(cols, [[1, 1], [1, 2], [2, 1], [2, 2], [1], [2], [0]],
           ["d_w1 d_w1", "d_w1 d_w2", "d_w2 d_w1", "d_w2 d_w2", "d_w1", "d_w2", "d_t"], n)


def test_integral_generation_wagner_platen_1d():
    n = 14
    I = _draw("wagner_platen", 2 ** n, 1)
    assert I["d_www"].shape == (2 ** n, 1, 1, 1) and I["d_wt"].shape == (2 ** n, 1, 1) and I["d_tw"].shape == (2 ** n, 1, 1)
    cols = [I["d_www"][:, 0, 0, 0], I["d_tw"][:, 0, 0], I["d_wt"][:, 0, 0], I["d_ww"][:, 0, 0], I["d_w"][:, 0],
            np.ones(2 ** n), 0.5 * np.ones(2 ** n)]
    _check(cols, [[1, 1, 1], [0, 1], [1, 0], [1, 1], [1], [0], [0, 0]],
           ["d_111", "d_01", "d_10", "d_11", "d_1", "d_0", "d_00"], n)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_integral_generation_wagner_platen(dtype):
    n = 12
    I = _draw("wagner_platen", 2 ** n, 2, p=50, dtype=dtype)
    cols, idx, lab = [], [], []
    for i, j, k in itertools.product((1, 2), repeat=3):
        cols.append(I["d_www"][:, i - 1, j - 1, k - 1]); idx.append([i, j, k]); lab.append(f"d_{i}{j}{k}")
    for i, j in itertools.product((0, 1, 2), repeat=2):
        if i == 0 and j == 0:
            continue
        if i == 0:
            cols.append(I["d_tw"][:, 0, j - 1])
        elif j == 0:
            cols.append(I["d_wt"][:, i - 1, 0])
        else:
            cols.append(I["d_ww"][:, i - 1, j - 1])
        idx.append([i, j]); lab.append(f"d_{i}{j}")
    for i in (1, 2):
        cols.append(I["d_w"][:, i - 1]); idx.append([i]); lab.append(f"d_{i}")
    cols += [np.ones(2 ** n), 0.5 * np.ones(2 ** n)]
    idx += [[0], [0, 0]]
    lab += ["d_0", "d_00"]
    _check(cols, idx, lab, n)
