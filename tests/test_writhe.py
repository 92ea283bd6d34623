"""Writhe of a closed polygon (the twist term of BASELINE config C4, reference
`examples/example_writhed_dna_loop.py:57`): the oracle's spherical-excess evaluation, the paper's arcsin formula,
brute-force quadrature of the Gauss integral and the product's Van Oosterom-Strackee evaluation must agree, and the
known properties of the writhe must hold.  pywrithe itself is unavailable (parity unpinned, see oracle/writhe.py)."""
import numpy as np
import pytest
import torch

import stochastic_b200.numpy as jnp
from stochastic_b200 import writhe as pw
from stochastic_b200.autodiff import grad
from oracle import writhe as ow


def knotty_polygon(n, seed=0, wobble=0.05):
    """Non-planar closed polygon winding around a (2,3) torus-knot-like path."""
    rng = np.random.default_rng(seed)
    k = 2 * np.pi * np.arange(n) / n
    loc = np.stack([np.cos(k) * (1 + 0.5 * np.cos(3 * k)), np.sin(k) * (1 + 0.5 * np.cos(3 * k)), 0.5 * np.sin(3 * k)], 1)
    return loc + wobble * rng.standard_normal(loc.shape)


@pytest.mark.parametrize("n", [4, 5, 9, 16, 40])
def test_three_exact_evaluations_and_quadrature_agree(n):
    loc = knotty_polygon(n, seed=n)
    w_oracle = float(ow.writhe(torch.tensor(loc)))
    assert abs(w_oracle - ow.writhe_klenin_langowski(loc)) < 1e-12
    assert abs(w_oracle - pw.writhe(loc)) < 1e-12
    q1, q2 = ow.writhe_quadrature(loc, 30), ow.writhe_quadrature(loc, 60)
    assert abs(q2 - w_oracle) < 0.3 * abs(q1 - w_oracle) + 1e-9       # midpoint rule: error falls ~4x per halving
    assert abs(q2 - w_oracle) < 2e-4 * max(1.0, abs(w_oracle))


def test_nearly_planar_rings_have_no_branch_jumps():
    """Slightly lifted rings (the C4 parity test's initial condition): many segment pairs are almost coplanar, the
    quadrilaterals almost degenerate; every evaluation must stay on the right branch (an early oracle returned -3)."""
    for n, amp in ((40, 0.3), (40, 1e-6), (12, 0.05), (6, 0.3)):
        k = 2 * np.pi * np.arange(n) / n
        loc = np.stack([np.cos(k), np.sin(k), amp * np.sin(2 * k)], axis=1) / (2 * np.pi)
        w = pw.writhe(loc)
        assert abs(w - float(ow.writhe(torch.tensor(loc)))) < 1e-12
        assert abs(w - ow.writhe_quadrature(loc, 24)) < 1e-4
        assert n % 4 or abs(w) < 1e-10     # n divisible by 4: the polygon is its own mirror image after a quarter turn
    loc = knotty_polygon(40, seed=5, wobble=0.0) * np.array([1.0, 1.0, 1e-3])     # flattened knot: still |Wr| ~ 3 crossings
    assert abs(pw.writhe(loc) - float(ow.writhe(torch.tensor(loc)))) < 1e-11
    assert abs(pw.writhe(loc) - ow.writhe_klenin_langowski(loc)) < 1e-9


def test_known_values_and_symmetries():
    n = 24
    k = 2 * np.pi * np.arange(n) / n
    planar = np.stack([np.cos(k), 0.7 * np.sin(k), 0 * k], 1)
    assert abs(pw.writhe(planar)) < 1e-14 and abs(float(ow.writhe(torch.tensor(planar)))) < 1e-14
    # a curve on a sphere that is symmetric under the antipodal map ... simpler exact statements:
    loc = knotty_polygon(n, seed=3)
    w = pw.writhe(loc)
    mirrored = loc * np.array([1.0, 1.0, -1.0])
    assert abs(pw.writhe(mirrored) + w) < 1e-12                       # chirality: mirror image flips the sign
    assert abs(pw.writhe(loc[::-1]) - w) < 1e-12                      # orientation reversal does not
    assert abs(pw.writhe(np.roll(loc, 5, axis=0)) - w) < 1e-12        # nor does the starting vertex
    rot = np.linalg.qr(np.random.default_rng(1).standard_normal((3, 3)))[0]
    rot *= np.sign(np.linalg.det(rot))
    assert abs(pw.writhe(2.5 * loc @ rot.T + 3.0) - w) < 1e-12        # rigid motions and scale
    # a figure-eight pressed almost flat has writhe close to +-1 (one crossing), its mirror image the opposite
    t = 2 * np.pi * (np.arange(200) + 0.5) / 200
    eight = np.stack([np.sin(t), np.sin(t) * np.cos(t), 0.02 * np.cos(t)], 1)
    w8 = pw.writhe(eight)
    assert abs(abs(w8) - 1.0) < 0.05


def test_strand_passage_changes_the_writhe_by_two():
    """Moving one strand through another is the only way the writhe of a closed curve jumps, and it jumps by 2."""
    t = 2 * np.pi * (np.arange(120) + 0.5) / 120
    def eight(h):
        return np.stack([np.sin(t), np.sin(t) * np.cos(t), h * np.cos(t)], 1)
    assert abs(abs(pw.writhe(eight(1e-3)) - pw.writhe(eight(-1e-3))) - 2.0) < 0.02


@pytest.mark.parametrize("n", [5, 8])
def test_traced_gradient_matches_oracle_autograd(n):
    """The generic (traced) kernel differentiates `writhe` symbolically; the bead kernel uses the analytic gradient of the
    same formula; the oracle differentiates a different formula by autograd.  Here: traced vs autograd."""
    from stochastic_b200.trace import expr as E
    loc = knotty_polygon(n, seed=11)
    x = jnp.symbols((3 * n,))
    g = grad(lambda v: pw.writhe(jnp.reshape(v, (n, 3))))(x)
    got = np.array([float(v) for v in E.evaluate([E.wrap(e) for e in np.asarray(g, dtype=object).ravel()], loc.ravel())])
    xt = torch.tensor(loc, requires_grad=True)
    ref = torch.autograd.grad(ow.writhe(xt), xt)[0].numpy().ravel()
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_gradient_is_finite_on_the_planar_circle():
    """The reference's initial condition (a planar circle) is where the paper's arcsin formula is not differentiable."""
    n = 40
    k = 2 * np.pi * np.arange(n) / n
    xt = torch.tensor(np.stack([np.cos(k), np.sin(k), 0 * k], 1), requires_grad=True)
    g = torch.autograd.grad(ow.writhe(xt), xt)[0]
    assert torch.isfinite(g).all() and g.abs().max() < 1e-12
