"""Exact first and second moments of multiple Ito integrals over [0, t] as polynomials in t.

Same public surface as the reference's `pychastic.wiener_integral_moments`
(reference `pychastic/wiener_integral_moments.py:6-83`): `E(alpha)` and `E2(alpha, beta)` return
`numpy.polynomial.Polynomial`s; multi-index entries are 0 for time and j >= 1 for the j-th
Wiener process.  Host-side numpy (as in the reference, where it is a test oracle for the
integral generator), implemented from Kloeden-Platen lemma 5.7.2:

    E[I_alpha] = t^l / l!   if alpha = (0, ..., 0) of length l, else 0
    E[I_alpha I_beta] satisfies, with a- = alpha without its last entry,
        d/dt E[I_alpha I_beta] = [a_l = b_l != 0] E[I_a- I_b-] + [a_l = 0] E[I_a- I_beta] + [b_l = 0] E[I_alpha I_b-]
"""
from functools import lru_cache

import numpy as np

Polynomial = np.polynomial.Polynomial


def E(alpha):
    alpha = tuple(alpha)
    if any(alpha):
        return Polynomial([0])
    out = Polynomial([1])
    for _ in alpha:
        out = out.integ()
    return out


@lru_cache(maxsize=None)
def _second_moment(alpha, beta):
    if not alpha:
        return E(beta)
    if not beta:
        return E(alpha)
    total = Polynomial([0])
    la, lb = alpha[-1], beta[-1]
    if la == lb and la != 0:
        total = total + _second_moment(alpha[:-1], beta[:-1]).integ()
    if la == 0:
        total = total + _second_moment(alpha[:-1], beta).integ()
    if lb == 0:
        total = total + _second_moment(alpha, beta[:-1]).integ()
    return total


def E2(alpha, beta=None):
    beta = alpha if beta is None else beta
    return _second_moment(tuple(alpha), tuple(beta))
