"""The BASELINE workloads as ready-made `SDEProblem`s (the "built-in functors").

Each is the reference's own definition, written against `stochastic_b200.numpy`:

* C1 `gbm`            reference `README.md:17`, `docs/source/examples.rst:36`
* C2 `polar_walk`     reference `tests/test_sde_solver.py:50-63`; drifted `examples/hit.py:9-31`,
                      `examples/examples_from_paper/polar_random_walk.py:8-16`
* C3 `two_spheres`    reference `examples/example_two_interacting_brownian_spheres.py:12-34`
*    `four_beads`     reference `examples/jfm.py:16-61` (spring energy selectable with / without 1/2)
* C5 `garch`          reference `docs/source/optionpricing.rst:117-123`
"""
import numpy as np

from . import grpy, jnp
from .autodiff import grad
from .sde_problem import SDEProblem


def _finish(problem, f64, x0=None):
    """fp64 runs: overwrite x0 with the exact double values (SDEProblem stores float32 like the reference)."""
    if f64:
        src = problem.x0 if x0 is None else np.asarray(x0, dtype=np.float64).reshape(np.shape(problem.x0))
        problem.x0 = np.asarray(src, dtype=np.float64)
    return problem


def gbm(alpha=0.2, beta=0.5, x0=1.0, tmax=2.0, f64=False):
    p = SDEProblem(lambda x: alpha * x, lambda x: beta * x, x0=x0, tmax=tmax,
                   exact_solution=lambda x0, t, w: x0 * np.exp((alpha - 0.5 * beta * beta) * t + beta * w))
    return _finish(p, f64, x0)


def polar_walk(y_drift=0.0, x0=(1.0, 0.0), tmax=1.0, f64=False):
    p = SDEProblem(
        a=lambda x: jnp.array([1 / (2 * x[0]) + y_drift * jnp.sin(x[1]), y_drift * jnp.cos(x[1]) / x[0]]),
        b=lambda x: jnp.array([[jnp.cos(x[1]), jnp.sin(x[1])], [-jnp.sin(x[1]) / x[0], jnp.cos(x[1]) / x[0]]]),
        x0=jnp.array(list(x0)), tmax=tmax)
    p.to_cartesian = lambda x: np.array([x[0] * np.cos(x[1]), x[0] * np.sin(x[1])])
    return _finish(p, f64, x0)


def garch(theta=0.1, omega=0.05, xi=0.1, mu=0.01, x0=(100.0, 0.05), tmax=2.0, f64=False):
    p = SDEProblem(lambda x: jnp.array([mu, theta * (omega - x[1])]),
                   lambda x: jnp.array([[jnp.sqrt(x[1]) * x[0], 0], [0, xi * x[1]]]),
                   x0=list(x0), tmax=tmax)
    return _finish(p, f64, x0)


def bead_chain(radii, springs, x0, tmax, f64=False):
    """a = mu(x) (-grad U), b = sqrt(2) chol(mu(x)), U = sum_k k_s (|r_i - r_j| - r0)^2 over `springs`
    = [(i, j, k_s, r0), ...]; mu = RPY mobility."""
    radii = np.asarray(radii, dtype=np.float64)
    n = len(radii)

    def u_ene(x):
        loc = jnp.reshape(x, (n, 3))
        e = 0.0
        for (i, j, ks, r0) in springs:
            e = e + ks * (jnp.sqrt(jnp.sum((loc[i] - loc[j]) ** 2)) - r0) ** 2
        return e

    def drift(x):
        mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
        return jnp.matmul(mu, -grad(u_ene)(x))

    def noise(x):
        mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
        return jnp.sqrt(2) * jnp.linalg.cholesky(mu)

    p = SDEProblem(drift, noise, x0=jnp.array(np.asarray(x0, dtype=np.float64).ravel()), tmax=tmax)
    p.u_ene = u_ene
    return _finish(p, f64, x0)


def two_spheres(tmax=500.0, f64=False):
    return bead_chain([1.0, 0.2], [(0, 1, 1.0, 4.0)], [[0, 0, 0], [0, 0, 4.0]], tmax, f64)


def four_beads(tmax=8000.0, half_spring=False, f64=False):
    k = 5.5 * (0.5 if half_spring else 1.0)
    return bead_chain([3.0, 1.0, 1.0, 1.0], [(0, 1, k, 4.0), (1, 2, k, 4.0), (2, 3, k, 4.0)],
                      [[-2.0, 0, 0], [2.0, 0, 0], [6.0, 0, 0], [10.0, 0, 0]], tmax, f64)


class BeadRingProblem:
    """Closed chain of N equal beads with RPY hydrodynamic interactions: the writhed-DNA-loop model of
    reference `examples/example_writhed_dna_loop.py:25-84`.  `twist=True` adds its twist energy
    `kt/2 (2 pi)^2 (Lk - Wr)^2` (`:36,57`), with the writhe Wr of the bead polygon from `stochastic_b200.writhe`
    (the reference's `pywrithe.writhe_jax` has neither source nor fixture here: that term is parity-unpinned).

    Carries the same attributes as `SDEProblem` (`a`, `b`, `x0`, `tmax`, `dimension`, `noise_terms`) - `a` and
    `b` are the reference's drift / noise written against `stochastic_b200.numpy`, usable by the generic traced
    kernel for small N - plus `bead_ring`, the parameter block of the dedicated CTA-per-trajectory kernel
    (csrc/sde_bead_kernel.cuh) that `SDESolver` dispatches to when the state is too large to trace (d > 24)."""

    def __init__(self, n_beads=40, beam_length=1.0, hydrodynamic_diameter=28.4 / 1111., steric_diameter=20.0 / 1111.,
                 kbT=1.0, persistence_length=500.0 / 1111., stretch_length=0.037 / 1111., overlap_strength=70.0,
                 x0=None, tmax=6000.0e-6, f64=False, twist=False, linking_number=2.7):
        n = int(n_beads)
        d0 = beam_length / n
        ks = kbT / (stretch_length * d0)
        kb = kbT * persistence_length * d0
        ko = overlap_strength * kbT
        delta = 0.6 * steric_diameter
        radius = hydrodynamic_diameter / 2.0
        kt = kbT * persistence_length * (2.0 / 3.0) * beam_length       # twisting stiffness (reference :36)
        twist_k = kt * 4.0 * np.pi * np.pi if twist else 0.0
        if twist and n < 4:
            raise ValueError("the twist term needs at least 4 beads (a triangle has no non-adjacent segments)")
        self.bead_ring = dict(n_beads=n, radius=radius, ks=ks, d0=d0, kb_over_d04=kb / d0 ** 4, ko=ko,
                              sigma2=steric_diameter ** 2, inv_delta2=1.0 / delta ** 2, twist=bool(twist),
                              twist_k=twist_k, linking_number=float(linking_number))
        radii = np.full(n, radius)

        def u_ene(x):
            loc = jnp.reshape(x, (n, 3))
            dist = jnp.sum((loc - jnp.roll(loc, 1, axis=0)) ** 2, axis=1) ** 0.5
            ext = ks * 0.5 * jnp.sum((dist - d0) ** 2)
            curv2 = jnp.sum((loc - 2 * jnp.roll(loc, 1, axis=0) + jnp.roll(loc, 2, axis=0)) ** 2, axis=1) / d0 ** 4
            bend = kb * 0.5 * jnp.sum(curv2)
            diff = loc[:, None, :] - loc[None, :, :]
            over = ko * (0.5 * jnp.sum(jnp.tanh((steric_diameter ** 2 - jnp.sum(diff ** 2, axis=-1)) / delta ** 2) + 1.0) - n)
            if twist:
                from .writhe import writhe
                return ext + bend + over + 0.5 * twist_k * (linking_number - writhe(loc)) ** 2
            return ext + bend + over

        def drift(x):
            mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
            return jnp.matmul(mu, -grad(u_ene)(x))

        def noise(x):
            mu = grpy.muTT(jnp.reshape(x, (n, 3)), radii)
            return jnp.sqrt(2) * jnp.linalg.cholesky(mu)

        self.u_ene, self.a, self.b = u_ene, drift, noise
        if x0 is None:   # circle of circumference beam_length (reference :79-81)
            k = np.arange(n)
            x0 = (beam_length / (2 * np.pi)) * np.stack([np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n),
                                                        np.zeros(n)], axis=1)
        self.x0 = np.asarray(x0, dtype=np.float64 if f64 else np.float32).reshape(-1)
        if self.x0.size != 3 * n:
            raise ValueError(f"Inconsistent shapes: x0 has {self.x0.size} entries, expected {3 * n}")
        if not tmax > 0:
            raise ValueError(f"tmax has to be positive, not {tmax}")
        self.tmax = tmax
        self.dimension = self.noise_terms = 3 * n
        self.exact_solution = None


def dna_loop(n_beads=40, tmax=6000.0e-6, f64=False, **kw):
    """BASELINE config C4; `twist=True` for the full energy of the reference example (the twist term is
    parity-unpinned, see `BeadRingProblem`)."""
    return BeadRingProblem(n_beads=n_beads, tmax=tmax, f64=f64, **kw)


def call_payoff(strike=120.0):
    """European call payoff max(S_T - K, 0) as a device-side observable."""
    return lambda x, w, t: jnp.array([jnp.maximum(x[0] - strike, 0.0)])


def polar_error_observables(y_drift=4.0, x_start=2.0):
    """Per-trajectory strong error |cart(x_T) - (w_T + cart(x0) + drift t)| and canonical angle phi_T of the
    drifted polar walk (cartesian motion is exact Brownian motion with drift: reference `examples/hit.py:69-80`,
    `examples/examples_from_paper/polar_random_walk.py:30-33`)."""
    def observables(x, w, t):
        ex = x[0] * jnp.cos(x[1]) - (w[0] + x_start)
        ey = x[0] * jnp.sin(x[1]) - (w[1] + y_drift * t)
        phi = jnp.fmod(x[1] + jnp.pi, 2 * jnp.pi) - jnp.pi
        return jnp.array([jnp.sqrt(ex * ex + ey * ey), phi])
    return observables


def precompile_all():
    """NVRTC-compile (no GPU needed) the kernels of the BASELINE workloads into the cubin cache."""
    from .sde_solver import SDESolver
    n = 0
    jobs = []
    for f64 in (True, False):
        for scheme in ("euler", "milstein", "wagner_platen"):
            jobs.append((gbm(f64=f64), scheme, None))
            jobs.append((polar_walk(4.0, (2.0, 0.0), f64=f64), scheme, None))
            jobs.append((polar_walk(f64=f64), scheme, None))
            jobs.append((garch(f64=f64), scheme, None))
        jobs.append((two_spheres(f64=f64), "euler", None))
        jobs.append((four_beads(f64=f64), "euler", None))
    jobs.append((garch(f64=True), "wagner_platen", call_payoff()))
    jobs.append((polar_walk(4.0, (2.0, 0.0), f64=True), "wagner_platen", polar_error_observables(4.0, 2.0)))
    for twist in (False, True):            # C4: the CTA-per-trajectory bead kernels
        for f64 in (True, False):
            SDESolver(scheme="euler", dt=1e-7).compile(dna_loop(n_beads=40, f64=f64, twist=twist))
            n += 1
    for prob, scheme, obs in jobs:
        # final-only launches (chunk_size=None) and every-step snapshots (the reference's default chunk_size=1)
        for chunk_size in (None, 1):
            SDESolver(scheme=scheme, dt=2.0 ** -7).compile(prob, observables=obs, chunk_size=chunk_size,
                                                            store_trajectories=(obs is None or chunk_size is None))
            n += 1
    return n


def brownian_rotations(tmax=20.0, f64=False):
    """Rotational Brownian motion on SO(3) in rotation-vector coordinates (Evensen et al. 2008), the
    `step_post_processing` showcase of the reference (`examples/examples_from_paper/brownian_rotations.py:7-93`).
    Returns (problem, canonicalize_coordinates)."""
    from . import lax
    from .autodiff import jacobian
    mobility = 2.0 * np.eye(3)
    mobility_d = np.linalg.cholesky(mobility)

    def spin_matrix(q):
        return jnp.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])

    def rotation_matrix(q):
        phi = jnp.sqrt(jnp.sum(q ** 2))
        rot = ((jnp.sin(phi) / phi) * spin_matrix(q) + jnp.cos(phi) * jnp.eye(3)
               + ((1.0 - jnp.cos(phi)) / phi ** 2) * q.reshape(1, 3) * q.reshape(3, 1))
        return lax.cond(phi > 0.01, lambda: rot, lambda: 1.0 * jnp.eye(3))

    def transformation_matrix(q):
        phi = jnp.sqrt(jnp.sum(q ** 2))
        trans = (0.5 * (1.0 / phi ** 2 - (jnp.sin(phi) / (2.0 * phi * (1.0 - jnp.cos(phi))))) * q.reshape(1, 3)
                 * q.reshape(3, 1) + spin_matrix(q) + (phi * jnp.sin(phi) / (1.0 - jnp.cos(phi))) * jnp.eye(3))
        return lax.cond(phi > 0.01, lambda: trans, lambda: 1.0 * jnp.eye(3))

    def metric_force(q):
        phi = jnp.sqrt(jnp.sum(q ** 2))
        scale = lax.cond(phi < 0.01, lambda t: -t / 6.0, lambda t: jnp.sin(t) / (1.0 - jnp.cos(t)) - 2.0 / t, phi)
        return lax.cond(phi > 0.0, lambda: (q / phi) * scale, lambda: jnp.array([0.0, 0.0, 0.0]))

    def t_mobility(q):
        return transformation_matrix(q) @ mobility @ (transformation_matrix(q).T)

    def drift(q):
        return t_mobility(q) @ metric_force(q) + jnp.einsum("iji->j", jacobian(t_mobility)(q))

    def noise(q):
        return jnp.sqrt(2) * transformation_matrix(q) @ (rotation_matrix(q).T) @ mobility_d

    def canonicalize_coordinates(q):
        phi = jnp.sqrt(jnp.sum(q ** 2))
        max_phi = jnp.pi
        canonical_phi = jnp.fmod(phi + max_phi, 2.0 * max_phi) - max_phi
        return lax.cond(phi > max_phi, lambda canonical_phi, phi, q: (canonical_phi / phi) * q,
                        lambda canonical_phi, phi, q: q, canonical_phi, phi, q)

    p = SDEProblem(drift, noise, tmax=tmax, x0=jnp.array([1.0, 0.0, 0.0]))
    return _finish(p, f64, [1.0, 0.0, 0.0]), canonicalize_coordinates


def evensen_rotations(tmax=2.0, f64=False):
    """Rotational Brownian motion of a sphere, started at the identity, in rotation-vector coordinates with
    branch-free small-angle guards: the published problem of reference
    `examples/published_problems/evensen_nalum_elgsaeter_macromol_2008.py:14-124` (mobility = identity, D = 1).

    Returns (problem, canonicalize_coordinates, rotation_variance) where `rotation_variance` is the device-side
    observable of its post-processing (`:137-151`): the squared components of delta_u = -1/2 eps_kij R_ij(q), whose
    ensemble mean is 1/6 - 5/12 exp(-6 D t) + 1/4 exp(-2 D t)."""
    from .autodiff import jacobian
    mobility = np.eye(3)
    mobility_d = np.linalg.cholesky(mobility)

    def spin_matrix(q):
        return jnp.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])

    def rotation_matrix(q):
        unsafe_phi_squared = jnp.sum(q ** 2)
        phi_squared = jnp.maximum(unsafe_phi_squared, 0.01 ** 2)
        phi = jnp.sqrt(phi_squared)
        rot = ((jnp.sin(phi) / phi) * spin_matrix(q) + jnp.cos(phi) * jnp.eye(3)
               + ((1.0 - jnp.cos(phi)) / phi ** 2) * q.reshape(1, 3) * q.reshape(3, 1))
        return jnp.where(phi_squared == unsafe_phi_squared, rot,
                         (1.0 - 0.5 * unsafe_phi_squared) * jnp.eye(3) + spin_matrix(q)
                         + 0.5 * q.reshape(1, 3) * q.reshape(3, 1))

    def transformation_matrix(q):
        unsafe_phi_squared = jnp.sum(q ** 2)
        phi_squared = jnp.maximum(unsafe_phi_squared, 0.01 ** 2)
        phi = jnp.sqrt(phi_squared)
        c = phi * jnp.sin(phi) / (1.0 - jnp.cos(phi))
        return jnp.where(phi_squared == unsafe_phi_squared,
                         ((1.0 - 0.5 * c) / (phi ** 2)) * q.reshape(1, 3) * q.reshape(3, 1) + 0.5 * spin_matrix(q)
                         + 0.5 * c * jnp.eye(3),
                         (1.0 / 12.0) * q.reshape(1, 3) * q.reshape(3, 1) + 0.5 * spin_matrix(q) + jnp.eye(3))

    def metric_force(q):
        unsafephi = jnp.sqrt(jnp.sum(q ** 2))
        phi = jnp.maximum(unsafephi, 0.01)
        scale = jnp.where(phi == unsafephi, jnp.sin(phi) / (1.0 - jnp.cos(phi)) - 2.0 / phi, -unsafephi / 6.0)
        return jnp.where(phi == unsafephi, (q / phi) * scale, jnp.array([0.0, 0.0, 0.0]))

    def t_mobility(q):
        return (transformation_matrix(q) @ (rotation_matrix(q).T) @ mobility @ rotation_matrix(q)
                @ (transformation_matrix(q).T))

    def drift(q):
        return t_mobility(q) @ metric_force(q) + jnp.einsum("iji->j", jacobian(t_mobility)(q))

    def noise(q):
        return jnp.sqrt(2) * transformation_matrix(q) @ (rotation_matrix(q).T) @ mobility_d

    def canonicalize_coordinates(q):
        unsafephi = jnp.sqrt(jnp.sum(q ** 2))
        phi = jnp.maximum(unsafephi, 0.01)
        max_phi = jnp.pi
        canonical_phi = jnp.fmod(phi + max_phi, 2.0 * max_phi) - max_phi
        return jnp.where(phi > max_phi, (canonical_phi / phi) * q, q)

    def rotation_variance(q, w, t):
        rot = rotation_matrix(q)           # R(x0)^T = identity
        du = [-0.5 * (rot[1][2] - rot[2][1]), -0.5 * (rot[2][0] - rot[0][2]), -0.5 * (rot[0][1] - rot[1][0])]
        return jnp.array([du[0] ** 2, du[1] ** 2, du[2] ** 2])

    p = SDEProblem(drift, noise, tmax=tmax, x0=jnp.array([0.0, 0.0, 0.0]))
    return _finish(p, f64, [0.0, 0.0, 0.0]), canonicalize_coordinates, rotation_variance
