"""Oracle: the SDE problems used by the reference's tests, examples and BASELINE configs,
written as torch callables for `oracle.taylor_ito` (TEST INFRASTRUCTURE ONLY).

Sources (reference paths):
* GBM                    `README.md:17`, `tests/test_sde_solver.py:9-15`, `tests/test_kloeden_platen_statistics.py:6-14`
* arctan (KP 4.27)       `tests/test_sde_solver.py:17-24`
* polar random walk      `tests/test_sde_solver.py:50-63`; drifted variant `examples/hit.py:9-31`
* parabolic random walk  `tests/test_sde_solver.py:69-84`
* exp test systems       `tests/test_coefficient_function_logic_1d.py:8-13`, `..._2d.py:8-19`
* GARCH / "Heston"       `docs/source/optionpricing.rst:117-123`
* tanh benchmark         `benchmarks/benchmark_sde_solver.py:12-18`
* sinh drift             `examples/examples_from_paper/sinh_drift.py:6-11`
* bead systems (RPY)     `examples/example_two_interacting_brownian_spheres.py:12-34`,
                         `examples/jfm.py:16-61`, `examples/published_problems/cichocki_...jfm_2019.py:29-33`
"""
import numpy as np
import torch

from .solver import OracleProblem
from . import rpy


def _stack(rows):
    return torch.stack([torch.stack(list(r)) if isinstance(r, (list, tuple)) else r for r in rows])


def gbm(alpha=0.2, beta=0.5, x0=1.0, tmax=2.0):
    def coeffs_numpy(x, scheme):
        n = x.shape[0]
        out = {"f_t": alpha * x, "f_w": (beta * x)[:, :, None]}
        if scheme != "euler":
            out["f_ww"] = (beta * beta * x).reshape(n, 1, 1, 1)
        if scheme == "wagner_platen":
            out["f_tw"] = (alpha * beta * x).reshape(n, 1, 1)
            out["f_wt"] = (alpha * beta * x).reshape(n, 1, 1)
            out["f_www"] = (beta ** 3 * x).reshape(n, 1, 1, 1, 1)
            out["f_tt"] = alpha * alpha * x
        return out
    return OracleProblem(lambda x: alpha * x, lambda x: (beta * x).reshape(1, 1), [x0], tmax,
                         coeffs_numpy=coeffs_numpy, name="gbm")


def arctan(a=1.0, x0=0.0, tmax=1.0):
    return OracleProblem(lambda x: -(a ** 2) * torch.sin(x) * torch.cos(x) ** 3,
                         lambda x: (a * torch.cos(x) ** 2).reshape(1, 1), [x0], tmax, name="arctan")


def tanh_benchmark(x0=0.0, tmax=5.0):
    return OracleProblem(lambda x: -(torch.tanh(x)) * (1 - torch.tanh(x) ** 2),
                         lambda x: (1 - torch.tanh(x) ** 2).reshape(1, 1), [x0], tmax, name="tanh")


def sinh_drift(x0=0.0, tmax=1.0):
    return OracleProblem(lambda x: 0.5 * x + (x ** 2 + 1.0) ** 0.5,
                         lambda x: ((x ** 2 + 1.0) ** 0.5).reshape(1, 1), [x0], tmax, name="sinh_drift")


def exp_1d():
    return OracleProblem(lambda x: torch.exp(3.0 * x), lambda x: torch.exp(7.0 * x).reshape(1, 1),
                         [0.0], 1.0, name="exp_1d")


def exp_2d():
    def a(x):
        return torch.stack([1.1 * torch.exp(3.0 * x[0]), 1.2 * torch.exp(5.0 * x[1])])

    def b(x):
        return _stack([[1.3 * torch.exp(7.0 * x[0]), 1.4 * torch.exp(11.0 * x[1])],
                       [1.5 * torch.exp(13.0 * x[1]), 1.6 * torch.exp(17.0 * x[0])]])
    return OracleProblem(a, b, [0.0, 0.0], 1.0, name="exp_2d")


def polar_walk(y_drift=0.0, x0=(1.0, 0.0), tmax=1.0, sigma=1.0):
    def a(x):
        r, ph = x[0], x[1]
        return torch.stack([1 / (2 * r) + y_drift * torch.sin(ph), y_drift * torch.cos(ph) / r])

    def b(x):
        r, ph = x[0], x[1]
        return sigma * _stack([[torch.cos(ph), torch.sin(ph)],
                               [-torch.sin(ph) / r, torch.cos(ph) / r]])
    return OracleProblem(a, b, list(x0), tmax, name="polar_walk")


def parabolic_walk(x0=(1.0, 0.0), tmax=0.05):
    def a(x):
        u, v = x[0], x[1]
        return torch.stack([(u - v) / (2 * (u + v) ** 2), (-u + v) / (2 * (u + v) ** 2)])

    def b(x):
        u, v = x[0], x[1]
        return _stack([[1 / (2 * u + 2 * v), v / (u + v)], [-1 / (2 * u + 2 * v), u / (u + v)]])
    return OracleProblem(a, b, list(x0), tmax, name="parabolic_walk")


def garch(theta=0.1, omega=0.05, xi=0.1, mu=0.01, x0=(100.0, 0.05), tmax=2.0):
    def a(x):
        return torch.stack([mu + 0 * x[0], theta * (omega - x[1])])

    def b(x):
        z = 0 * x[0]
        return _stack([[torch.sqrt(x[1]) * x[0], z], [z, xi * x[1]]])
    return OracleProblem(a, b, list(x0), tmax, name="garch")


# ---------------------------------------------------------------------------------------
# bead systems: a = mu(x) (-grad U), b = sqrt(2) chol(mu(x))
# ---------------------------------------------------------------------------------------
def bead_chain(radii, springs, x0, tmax, name="beads"):
    """springs: list of (i, j, k, r0) with U = sum k (|r_i - r_j| - r0)^2 (no 1/2 inside k)."""
    radii_t = torch.as_tensor(np.asarray(radii, dtype=np.float64))
    n = len(radii)

    def energy(x):
        loc = x.reshape(n, 3)
        e = 0
        for (i, j, k, r0) in springs:
            e = e + k * (torch.sqrt(torch.sum((loc[i] - loc[j]) ** 2)) - r0) ** 2
        return e

    def a(x):
        mu = rpy.muTT_torch(x.reshape(n, 3), radii_t.to(x.dtype))
        force = -torch.func.grad(energy)(x)
        return mu @ force

    def b(x):
        mu = rpy.muTT_torch(x.reshape(n, 3), radii_t.to(x.dtype))
        return (2 ** 0.5) * torch.linalg.cholesky(mu)
    return OracleProblem(a, b, list(x0), tmax, name=name)


def two_spheres(tmax=500.0):
    return bead_chain([1.0, 0.2], [(0, 1, 1.0, 4.0)], [0, 0, 0, 0, 0, 4.0], tmax, name="two_spheres")


def jfm_four_beads(tmax=8000.0, half=False):
    """Golden recipe: spring energy WITHOUT the 1/2 (published_problems variant)."""
    k = 5.5 * (0.5 if half else 1.0)
    return bead_chain([3.0, 1.0, 1.0, 1.0], [(0, 1, k, 4.0), (1, 2, k, 4.0), (2, 3, k, 4.0)],
                      [-2.0, 0, 0, 2.0, 0, 0, 6.0, 0, 0, 10.0, 0, 0], tmax, name="jfm_four_beads")


def dna_loop(n_beads=40, beam_length=1.0, hydrodynamic_diameter=28.4 / 1111., steric_diameter=20.0 / 1111., kbT=1.0,
             persistence_length=500.0 / 1111., stretch_length=0.037 / 1111., overlap_strength=70.0, tmax=6000.0e-6,
             twist=False, linking_number=2.7, x0=None):
    """reference `examples/example_writhed_dna_loop.py:25-84`: energy written as in the example, differentiated by
    torch autograd.  `twist=True` includes its twist term (`:57`) with the polygon writhe of `oracle.writhe`
    (pywrithe itself is unavailable: parity unpinned for that term)."""
    from . import writhe as owrithe
    twisting_constant = kbT * persistence_length * (2.0 / 3.0) * beam_length
    n = n_beads
    d0 = beam_length / n
    ks = kbT / (stretch_length * d0)
    kb = kbT * persistence_length * d0
    ko = overlap_strength * kbT
    delta = 0.6 * steric_diameter
    radii = torch.full((n,), hydrodynamic_diameter / 2.0, dtype=torch.float64)

    def energy(x):
        loc = x.reshape(n, 3)
        dist = torch.sum((loc - torch.roll(loc, 1, 0)) ** 2, dim=1) ** 0.5
        ext = ks * 0.5 * torch.sum((dist - d0) ** 2)
        curv = torch.sum((loc - 2 * torch.roll(loc, 1, 0) + torch.roll(loc, 2, 0)) ** 2, dim=1) ** 0.5 / d0 ** 2
        bend = kb * 0.5 * torch.sum(curv ** 2)
        r2 = torch.sum((loc[:, None, :] - loc[None, :, :]) ** 2, dim=-1)
        over = ko * (0.5 * torch.sum(torch.tanh((steric_diameter ** 2 - r2) / delta ** 2) + 1.0) - n)
        if twist:
            return ext + bend + over + owrithe.twist_energy(loc, twisting_constant, linking_number)
        return ext + bend + over

    def a(x):
        mu = rpy.muTT_torch(x.reshape(n, 3), radii.to(x.dtype))
        return mu @ (-torch.func.grad(energy)(x))

    def b(x):
        mu = rpy.muTT_torch(x.reshape(n, 3), radii.to(x.dtype))
        return (2 ** 0.5) * torch.linalg.cholesky(mu)

    if x0 is None:
        k = np.arange(n)
        x0 = (beam_length / (2 * np.pi)) * np.stack([np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), np.zeros(n)], 1)
    return OracleProblem(a, b, np.asarray(x0, dtype=np.float64).ravel(), tmax, name="dna_loop")


def brownian_rotations(tmax=20.0):
    """reference `examples/examples_from_paper/brownian_rotations.py:7-86` with `jax.lax.cond` written as
    `torch.where` (what vmap turns it into).  Returns (problem, canonicalize_coordinates on numpy batches)."""
    mobility = 2.0 * torch.eye(3, dtype=torch.float64)
    mobility_d = torch.linalg.cholesky(mobility)
    eye = torch.eye(3, dtype=torch.float64)

    def spin(q):
        z = 0 * q[0]
        return torch.stack([torch.stack([z, -q[2], q[1]]), torch.stack([q[2], z, -q[0]]), torch.stack([-q[1], q[0], z])])

    def rotation_matrix(q):
        phi = torch.sqrt(torch.sum(q ** 2))
        rot = ((torch.sin(phi) / phi) * spin(q) + torch.cos(phi) * eye.to(q.dtype)
               + ((1.0 - torch.cos(phi)) / phi ** 2) * q.reshape(1, 3) * q.reshape(3, 1))
        return torch.where(phi > 0.01, rot, eye.to(q.dtype))

    def transformation_matrix(q):
        phi = torch.sqrt(torch.sum(q ** 2))
        trans = (0.5 * (1.0 / phi ** 2 - (torch.sin(phi) / (2.0 * phi * (1.0 - torch.cos(phi))))) * q.reshape(1, 3)
                 * q.reshape(3, 1) + spin(q) + (phi * torch.sin(phi) / (1.0 - torch.cos(phi))) * eye.to(q.dtype))
        return torch.where(phi > 0.01, trans, eye.to(q.dtype))

    def metric_force(q):
        phi = torch.sqrt(torch.sum(q ** 2))
        scale = torch.where(phi < 0.01, -phi / 6.0, torch.sin(phi) / (1.0 - torch.cos(phi)) - 2.0 / phi)
        return torch.where(phi > 0.0, (q / phi) * scale, torch.zeros(3, dtype=q.dtype))

    def t_mobility(q):
        T = transformation_matrix(q)
        return T @ mobility.to(q.dtype) @ T.T

    def drift(q):
        return t_mobility(q) @ metric_force(q) + torch.einsum("iji->j", torch.func.jacfwd(t_mobility)(q))

    def noise(q):
        return (2 ** 0.5) * transformation_matrix(q) @ rotation_matrix(q).T @ mobility_d.to(q.dtype)

    def canonicalize(x):
        phi = np.sqrt(np.sum(x ** 2, axis=-1, keepdims=True))
        canonical = np.fmod(phi + np.pi, 2 * np.pi) - np.pi
        return np.where(phi > np.pi, (canonical / phi) * x, x)

    return OracleProblem(drift, noise, [1.0, 0.0, 0.0], tmax, name="brownian_rotations"), canonicalize
