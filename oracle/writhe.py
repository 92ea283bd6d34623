"""Oracle: writhe of a closed polygon and the twist energy of the DNA-loop example (TEST INFRASTRUCTURE ONLY).

The reference calls `pywrithe.writhe_jax(locations)` (reference `examples/example_writhed_dna_loop.py:57`):
an UN-VENDORED dependency (PyPI `pywrithe`, unpinned; no source, fixture or call site other than that line in
/root/reference).  **Parity unpinned**: what is restated here is the published algorithm - the writhe of a closed
polygon is the Gauss double integral over all pairs of its straight segments, and for two straight segments that
integral is exactly (signed area of a spherical quadrilateral) / 4 pi (Klenin & Langowski, Biopolymers 54 (2000)
307-317, "method 1a") - not pywrithe's code.  The writhe of a polygon is a geometric quantity, so any exact
evaluation agrees with any other up to rounding; that is what the tests below rely on:

* `writhe` (this file, used as the oracle of the CUDA kernel): spherical excess of the quadrilateral with vertices
  r13, r23, r24, r14 on the unit sphere, interior angles from atan2 - a different formula from the kernel's
  (Van Oosterom-Strackee triangle solid angles), differentiated by torch autograd;
* `writhe_klenin_langowski` (numpy): the paper's arcsin formula with its sign rule, for non-degenerate inputs;
* `writhe_quadrature`: brute-force midpoint rule of the Gauss integral on the subdivided polygon.
"""
import numpy as np
import torch


def _pairs(n):
    """Unordered pairs (i, j), i < j, of non-adjacent segments of a closed n-gon (segment k: vertex k -> k+1 mod n)."""
    ii, jj = [], []
    for i in range(n):
        for j in range(i + 2, n):
            if i == 0 and j == n - 1:
                continue
            ii.append(i)
            jj.append(j)
    return np.array(ii), np.array(jj)


def _angle_at(v, prev, nxt):
    """Interior angle of a spherical polygon at unit vector v between the great-circle arcs to prev and nxt,
    measured counter-clockwise (seen from outside the sphere) from the arc v->nxt to the arc v->prev.

    The image of a pair of straight segments is a CONVEX spherical quadrilateral (every interior angle in [0, pi]:
    the paper's arcsin form is pi/2 + asin(.) per corner), so a value that atan2 returns just below 0 is a rounded 0
    (degenerate, e.g. coplanar segments) and only values near -pi are rounded pi."""
    tp = prev - (prev * v).sum(-1, keepdim=True) * v           # tangent towards prev
    tn = nxt - (nxt * v).sum(-1, keepdim=True) * v             # tangent towards nxt
    s = (torch.linalg.cross(tn, tp) * v).sum(-1)
    c = (tn * tp).sum(-1)
    ang = torch.atan2(s, c)
    return torch.where(ang < -0.5 * np.pi, ang + 2 * np.pi, ang)


def pair_solid_angles(loc):
    """Signed solid angle Omega_ij (= 4 pi x Gauss integral) of every non-adjacent segment pair, shape (n_pairs,)."""
    n = loc.shape[0]
    ii, jj = _pairs(n)
    p1, p2 = loc[ii], loc[(ii + 1) % n]
    p3, p4 = loc[jj], loc[(jj + 1) % n]
    unit = lambda v: v / torch.linalg.norm(v, dim=-1, keepdim=True)
    # image of the parameter square (s, t) -> (r2(t) - r1(s)) / |.|, corners in the order (0,0), (1,0), (1,1), (0,1)
    a, b, c, d = unit(p3 - p1), unit(p3 - p2), unit(p4 - p2), unit(p4 - p1)
    # the quadrilateral is traversed a -> b -> c -> d; its orientation decides the sign
    vol = (a * torch.linalg.cross(b, c)).sum(-1) + (a * torch.linalg.cross(c, d)).sum(-1)
    ccw = vol >= 0
    # interior angles for a counter-clockwise traversal; for a clockwise one traverse backwards
    def excess(q0, q1, q2, q3):
        return (_angle_at(q0, q3, q1) + _angle_at(q1, q0, q2) + _angle_at(q2, q1, q3) + _angle_at(q3, q2, q0)
                - 2 * np.pi)
    pos = excess(a, b, c, d)
    neg = excess(a, d, c, b)
    return torch.where(ccw, pos, -neg)


def writhe(loc):
    """Writhe of the closed polygon through the rows of `loc` (n, 3); torch, differentiable."""
    return pair_solid_angles(loc).sum() / (2 * np.pi)


def writhe_klenin_langowski(loc):
    """Klenin-Langowski method 1a, literally (numpy; undefined for coplanar segment pairs)."""
    loc = np.asarray(loc, dtype=np.float64)
    n = len(loc)
    total = 0.0
    ii, jj = _pairs(n)
    for i, j in zip(ii, jj):
        p1, p2, p3, p4 = loc[i], loc[(i + 1) % n], loc[j], loc[(j + 1) % n]
        r12, r34, r13, r14, r23, r24 = p2 - p1, p4 - p3, p3 - p1, p4 - p1, p3 - p2, p4 - p2
        un = lambda v: v / np.linalg.norm(v)
        n1, n2, n3, n4 = un(np.cross(r13, r14)), un(np.cross(r14, r24)), un(np.cross(r24, r23)), un(np.cross(r23, r13))
        cl = lambda x: np.clip(x, -1.0, 1.0)
        om = np.arcsin(cl(n1 @ n2)) + np.arcsin(cl(n2 @ n3)) + np.arcsin(cl(n3 @ n4)) + np.arcsin(cl(n4 @ n1))
        total += om * np.sign(np.cross(r34, r12) @ r13)
    return total / (2 * np.pi)


def writhe_quadrature(loc, sub=40):
    """(1 / 4 pi) sum over sub-segment midpoints of (dr1 x dr2) . (r1 - r2) / |r1 - r2|^3, non-adjacent segments only
    (adjacent and identical straight segments contribute exactly zero)."""
    loc = np.asarray(loc, dtype=np.float64)
    n = len(loc)
    s = (np.arange(sub) + 0.5) / sub
    seg = np.roll(loc, -1, axis=0) - loc
    mid = loc[:, None, :] + s[None, :, None] * seg[:, None, :]           # (n, sub, 3)
    dr = np.broadcast_to(seg[:, None, :] / sub, mid.shape)
    total = 0.0
    ii, jj = _pairs(n)
    for i, j in zip(ii, jj):
        r = mid[i][:, None, :] - mid[j][None, :, :]
        cr = np.cross(dr[i][:, None, :], dr[j][None, :, :])
        total += 2.0 * ((cr * r).sum(-1) / np.linalg.norm(r, axis=-1) ** 3).sum()
    return total / (4 * np.pi)


def twist_energy(loc, twisting_constant, linking_number):
    """reference `examples/example_writhed_dna_loop.py:57`."""
    return twisting_constant * 0.5 * (4.0 * np.pi * np.pi) * (linking_number - writhe(loc)) ** 2
