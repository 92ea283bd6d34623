"""Writhe of a closed polygon: stands in for `pywrithe.writhe_jax`, which the reference's DNA-loop example calls
(reference `examples/example_writhed_dna_loop.py:57`) and which is an un-vendored, unpinned dependency with no source
in the reference tree (**parity unpinned**: the published algorithm is implemented, not pywrithe's code).

Wr = (1 / 4 pi) sum over ordered pairs of non-adjacent segments of the Gauss double integral; for two straight segments
p1 -> p2 and p3 -> p4 that integral is the signed area of the spherical quadrilateral with corners
a = p3 - p1, b = p3 - p2, c = p4 - p2, d = p4 - p1 (Klenin & Langowski, Biopolymers 54 (2000), method 1a).  The area is
evaluated as two triangle solid angles (Van Oosterom & Strackee 1983):

    Omega(a, b, c) = 2 atan2( a . (b x c),  |a||b||c| + (a.b)|c| + (a.c)|b| + (b.c)|a| )

which - unlike the paper's arcsin form - stays differentiable for coplanar segments (the planar circle the
example starts from).  Written against `stochastic_b200.numpy`, so it traces (symbolic gradient through
`autodiff.grad`) for small rings and evaluates with numpy on concrete arrays.  Rings too large to trace use the same
formula with its analytic gradient inside csrc/sde_bead_kernel.cuh (SDE_BEAD_TWIST).
"""
import numpy as np

from . import jnp


def _dot(u, v):
    return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]


def _cross(u, v):
    return [u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]]


def triangle_solid_angle(a, b, c):
    la, lb, lc = jnp.sqrt(_dot(a, a)), jnp.sqrt(_dot(b, b)), jnp.sqrt(_dot(c, c))
    num = _dot(a, _cross(b, c))
    den = la * lb * lc + _dot(a, b) * lc + _dot(a, c) * lb + _dot(b, c) * la
    return 2.0 * jnp.arctan2(num, den)


def segment_pairs(n):
    """Unordered pairs (i, j), i < j, of non-adjacent segments of a closed n-gon (segment k: vertex k -> k+1 mod n)."""
    return [(i, j) for i in range(n) for j in range(i + 2, n) if not (i == 0 and j == n - 1)]


def writhe(locations):
    """Writhe of the closed polygon through the rows of `locations` (n, 3)."""
    n = np.shape(locations)[0]
    total = 0.0
    for i, j in segment_pairs(n):
        p1, p2, p3, p4 = locations[i], locations[(i + 1) % n], locations[j], locations[(j + 1) % n]
        a, b, c, d = p3 - p1, p3 - p2, p4 - p2, p4 - p1
        total = total + triangle_solid_angle(a, b, c) + triangle_solid_angle(a, c, d)
    return total / (2.0 * np.pi)


writhe_jax = writhe        # the name the reference imports
