"""Oracle: `SDESolver.solve_many` restated on the CPU (numpy + torch.func).

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).

Follows reference `pychastic/sde_solver.py`:

* initial-condition tiling / multi-IC rule           `:141-149`
* the stochastic Taylor step                          `:186-215`
* chunk arithmetic (`int(tmax/dt)`, round-up)         `:217-221`
* `scan_func` (rescaling by dt^1/2, dt, dt^3/2; t, x, w update order)  `:228-250`
* fragment / chunk / step nest and snapshot emission  `:263-295`
* key fan-out  PRNGKey(seed) -> split(n_traj) -> split(n_fragments)    `:223,292,299`

Coefficient tensors come either from `oracle.taylor_ito` (autodiff of torch callables, the
faithful path) or from a problem's hand-derived numpy closed forms (`coeffs_numpy`, used
for large CPU-baseline runs and cross-checked against autodiff in the tests).
"""

import numpy as np
import torch

from . import jax_prng
from .taylor_ito import batched_coefficient_functions
from .wiener_integrals import get_wiener_integrals_many

SCHEMES = ("euler", "milstein", "wagner_platen")


def chunk_arithmetic(tmax, dt, chunk_size, chunks_per_randomization):
    """Reference `sde_solver.py:217-221`; returns (chunk_size, n_chunks, cpr, n_fragments)."""
    steps_needed = int(tmax / dt)
    tmp = chunks_per_randomization or 1
    chunk_size = chunk_size or steps_needed
    n_chunks = ((steps_needed // (chunk_size * tmp)) + (1 if steps_needed % (chunk_size * tmp) else 0)) * tmp
    cpr = chunks_per_randomization or n_chunks
    return chunk_size, n_chunks, cpr, n_chunks // cpr


class OracleProblem:
    """a, b: torch callables (d,)->(d,), (d,)->(d,m);  coeffs_numpy optional closed forms."""

    def __init__(self, a, b, x0, tmax, coeffs_numpy=None, name=""):
        self.a, self.b, self.tmax, self.name = a, b, tmax, name
        self.x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
        self.coeffs_numpy = coeffs_numpy
        xt = torch.as_tensor(self.x0 if self.x0.ndim == 1 else self.x0[0], dtype=torch.float64)
        self.dimension, self.noise_terms = tuple(b(xt).shape)


def _autodiff_coeffs(problem, scheme, dtype):
    # Coefficient tensors are always evaluated in float64 and rounded to the working dtype: nested
    # torch.func.jacfwd mixes dtypes in float32, and a correctly rounded coefficient is within half an ulp
    # of whatever float32 evaluation order the reference's XLA program would use.
    fns = batched_coefficient_functions(problem.a, problem.b)
    tdt = torch.float64
    out_dt = np.dtype(dtype)
    need = {"euler": ("f_t", "f_w"),
            "milstein": ("f_t", "f_w", "f_ww"),
            "wagner_platen": ("f_t", "f_w", "f_ww", "f_tw", "f_wt", "f_www", "f_tt")}[scheme]

    def coeffs(x):
        xt = torch.as_tensor(x, dtype=tdt)
        return {k: fns[k](xt).numpy().astype(out_dt) for k in need}
    return coeffs


def solve_many(problem, scheme="euler", dt=0.01, n_trajectories=1, seed=0, chunk_size=1,
               chunks_per_randomization=None, dtype=np.float32, layout="original",
               step_post_processing=None, use_closed_form=False, traj_slice=None, p=10,
               max_fragments=None):
    """Returns dict(time_values (n,K), solution_values (n,K,d), wiener_values (n,K,m)).

    `traj_slice=(begin, count)` integrates only that range of the n_trajectories-wide key
    fan-out (used to check sharded runs).  `max_fragments` stops after that many fragments
    (keys are still derived for the full run, so a prefix of a long run is reproduced)."""
    if scheme not in SCHEMES:
        raise ValueError(scheme)
    dt_np = np.dtype(dtype)
    f = dt_np.type
    x0 = np.asarray(problem.x0)
    if x0.ndim == 2:
        if n_trajectories is not None:
            raise ValueError("n_trajectories option not supported with multiple initial conditions")
        init = x0
    else:
        init = np.tile(x0.reshape(1, -1), (n_trajectories, 1))
    n_total = init.shape[0]
    d, m = problem.dimension, problem.noise_terms

    chunk, n_chunks, cpr, n_frag = chunk_arithmetic(problem.tmax, dt, chunk_size, chunks_per_randomization)
    S = chunk * cpr

    keys = jax_prng.split(jax_prng.PRNGKey(seed, x64=(np.dtype(dtype) == np.float64)), n_total, layout)
    if traj_slice is not None:
        b0, cnt = traj_slice
        keys = keys[b0:b0 + cnt]
        init = init[b0:b0 + cnt]
    n = init.shape[0]

    coeffs = (problem.coeffs_numpy if (use_closed_form and problem.coeffs_numpy is not None)
              else _autodiff_coeffs(problem, scheme, dt_np))

    sqrt_dt = np.sqrt(f(dt))                 # jnp.sqrt(self.dt) in the working dtype
    dt_w = f(dt)
    dt32 = f(dt ** (3 / 2))                  # python double, then weak-typed into dtype
    x = init.astype(dt_np)
    w = np.zeros((n, m), dtype=dt_np)
    t = np.zeros((n,), dtype=dt_np)
    out_t = np.zeros((n, n_chunks), dtype=dt_np)
    out_x = np.zeros((n, n_chunks, d), dtype=dt_np)
    out_w = np.zeros((n, n_chunks, m), dtype=dt_np)

    frag_keys = np.stack([jax_prng.split(k, n_frag, layout) for k in keys])      # (n, n_frag, 2)
    n_run = n_frag if max_fragments is None else min(n_frag, max_fragments)
    for fr in range(n_run):
        I = get_wiener_integrals_many(frag_keys[:, fr], S, m, scheme, p=p, dtype=dt_np, layout=layout)
        for c in range(cpr):
            for k in range(chunk):
                s = c * chunk + k
                dW = sqrt_dt * I["d_w"][:, s].reshape(n, m)
                cf = coeffs(x) if not (use_closed_form and problem.coeffs_numpy is not None) \
                    else problem.coeffs_numpy(x, scheme)
                t = t + dt_w
                new_x = x + ((cf["f_t"].reshape(n, d) * dt_w) + np.einsum("nij,nj->ni", cf["f_w"], dW))
                if scheme in ("milstein", "wagner_platen"):
                    Iww = dt_w * I["d_ww"][:, s]
                    new_x = new_x + np.einsum("nijk,njk->ni", cf["f_ww"], Iww)
                if scheme == "wagner_platen":
                    Iwt = dt32 * I["d_wt"][:, s].reshape(n, m)
                    Itw = dt32 * I["d_tw"][:, s].reshape(n, m)
                    Iwww = dt32 * I["d_www"][:, s]
                    new_x = new_x + (np.einsum("nij,nj->ni", cf["f_tw"].reshape(n, d, m), Itw)
                                     + np.einsum("nij,nj->ni", cf["f_wt"].reshape(n, d, m), Iwt)
                                     + np.einsum("nijkl,njkl->ni", cf["f_www"], Iwww)
                                     + cf["f_tt"].reshape(n, d) * dt_w * dt_w / f(2))
                x = new_x.astype(dt_np)
                if step_post_processing is not None:
                    x = step_post_processing(x).astype(dt_np)
                w = (w + dW).astype(dt_np)
            ci = fr * cpr + c
            out_t[:, ci] = t
            out_x[:, ci] = x
            out_w[:, ci] = w
    K = n_run * cpr
    return dict(time_values=out_t[:, :K], solution_values=out_x[:, :K], wiener_values=out_w[:, :K])


def solve(problem, **kw):
    kw.pop("n_trajectories", None)
    sol = solve_many(problem, n_trajectories=1, **kw)
    return {k: v[0] for k, v in sol.items()}
