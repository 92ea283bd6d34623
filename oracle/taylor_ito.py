"""Oracle: Taylor-Ito coefficient functions by automatic differentiation (torch.func, CPU).

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).

Restates reference `pychastic/sde_solver.py:16-46` (`contract_all`, `L_t_operator`,
`L_w_operator`) and the closure construction of `sde_solver.py:157-184` with
`torch.func.jacfwd` standing in for `jax.jacobian` / `jax.hessian`.  It is deliberately
independent of the product's symbolic differentiator so that it can check it.

Index convention (reference comments at `sde_solver.py:36,44`): [spatial, noise/time, ...].
"""
import torch
from torch.func import jacfwd, vmap


def L_t_operator(f, a, b):
    """L^0 f = df.a + 1/2 d2f : (b b^T), with a singleton time axis at position 1."""
    def wrapped(x):
        bv = b(x)
        jac = jacfwd(f)(x)
        hes = jacfwd(jacfwd(f))(x)
        val = torch.tensordot(jac, a(x), dims=1) + 0.5 * torch.tensordot(hes, bv @ bv.T, dims=2)
        return val.unsqueeze(1)
    return wrapped


def L_w_operator(f, a, b):
    """L^j f = df.b, the new noise index moved to position 1."""
    def wrapped(x):
        val = torch.tensordot(jacfwd(f)(x), b(x), dims=1)          # (d, ..., m)
        return torch.movedim(val, -1, 1)                           # (d, m, ...)
    return wrapped


def L(f, idx, a, b):
    for c in reversed(idx):
        if c == "t":
            f = L_t_operator(f, a, b)
        elif c == "w":
            f = L_w_operator(f, a, b)
        else:
            raise ValueError(idx)
    return f


def coefficient_functions(a, b):
    """The seven closures of reference `sde_solver.py:172-184`, single-state versions."""
    ident = lambda x: x
    return {name: L(ident, name[2:], a, b)
            for name in ("f_t", "f_w", "f_ww", "f_tw", "f_wt", "f_www", "f_tt")}


def batched_coefficient_functions(a, b):
    """Same closures vmapped over a leading trajectory axis."""
    return {k: vmap(v) for k, v in coefficient_functions(a, b).items()}
