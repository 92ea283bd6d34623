"""Taylor-Ito coefficient tensors and the one-step update as symbolic DAGs.

Replaces the trace-time closure construction of the reference
(`L_t_operator`, `L_w_operator`, `L`, `step`; reference `pychastic/sde_solver.py:29-46,157-215`):

    L^0 f = sum_k a_k d_k f + 1/2 sum_kl (b b^T)_kl d_k d_l f          L^j f = sum_k b_kj d_k f
    f_t = a, f_w = b, f_ww[i,j1,j2] = L^j1 b[i,j2], f_wt[i,j] = L^j a_i, f_tw[i,j] = L^0 b[i,j],
    f_www[i,j1,j2,j3] = L^j1 L^j2 b[i,j3], f_tt[i] = L^0 a_i

but as hash-consed scalar expressions (shared subexpressions, structural zeros dropped), so
that the emitted device function evaluates a(x), b(x) and each derivative exactly once.
"""
import numpy as np

from . import expr as E
from .. import jnp

SCHEMES = ("euler", "milstein", "wagner_platen")


class TracedProblem:
    """a (d,), b (d, m) as arrays of Expr in the state variables var(0..d-1)."""

    def __init__(self, a_fn, b_fn, d):
        x = jnp.symbols((d,))
        a = np.asarray(jnp._as_obj(a_fn(x)), dtype=object)
        b = np.asarray(jnp._as_obj(b_fn(x)), dtype=object)
        if a.shape != (d,):
            raise ValueError(f"drift returned shape {a.shape}, expected {(d,)}")
        if b.ndim != 2 or b.shape[0] != d:
            raise ValueError(f"noise returned shape {b.shape}, expected ({d}, m)")
        self.d = d
        self.m = b.shape[1]
        self.a = [E.wrap(v) for v in a]
        self.b = [[E.wrap(v) for v in row] for row in b]
        self._dmemo = [dict() for _ in range(d)]
        self._bbT = None

    def D(self, f, k):
        return E.diff(f, k, self._dmemo[k])

    def bbT(self, k, l):
        if self._bbT is None:
            self._bbT = {}
        key = (min(k, l), max(k, l))
        v = self._bbT.get(key)
        if v is None:
            v = E.ZERO
            for j in range(self.m):
                v = v + self.b[k][j] * self.b[l][j]
            self._bbT[key] = v
        return v

    def Lw(self, f, j):
        out = E.ZERO
        for k in range(self.d):
            out = out + self.b[k][j] * self.D(f, k)
        return out

    def Lt(self, f):
        out = E.ZERO
        for k in range(self.d):
            dk = self.D(f, k)
            out = out + self.a[k] * dk
            if dk.is_const and dk.value == 0.0:
                continue
            for l in range(self.d):
                out = out + E.HALF * (self.bbT(k, l) * self.D(dk, l))
        return out

    def coefficients(self, scheme):
        """dict of nested lists of Expr: f_t[i], f_w[i][j], f_ww[i][j1][j2], f_tw[i][j], f_wt[i][j],
        f_www[i][j1][j2][j3], f_tt[i] (only those the scheme uses)."""
        if scheme not in SCHEMES:
            raise ValueError(f"unknown scheme {scheme!r}")
        d, m = self.d, self.m
        out = {"f_t": list(self.a), "f_w": [list(r) for r in self.b]}
        if scheme == "euler":
            return out
        out["f_ww"] = [[[self.Lw(self.b[i][j2], j1) for j2 in range(m)] for j1 in range(m)] for i in range(d)]
        if scheme == "milstein":
            return out
        out["f_wt"] = [[self.Lw(self.a[i], j) for j in range(m)] for i in range(d)]
        out["f_tw"] = [[self.Lt(self.b[i][j]) for j in range(m)] for i in range(d)]
        out["f_www"] = [[[[self.Lw(out["f_ww"][i][j2][j3], j1) for j3 in range(m)] for j2 in range(m)]
                         for j1 in range(m)] for i in range(d)]
        out["f_tt"] = [self.Lt(self.a[i]) for i in range(d)]
        return out


class StepVars:
    """Variable numbering of the emitted step function and the C expression of each."""

    def __init__(self, d, m):
        self.d, self.m = d, m
        self.names = {}
        n = 0
        for i in range(d):
            self.names[n] = f"x[{i}]"
            n += 1
        self.dt = n
        self.names[n] = "dt"
        n += 1
        self.dW = list(range(n, n + m))
        for j in range(m):
            self.names[n + j] = f"I.dW[{j}]"
        n += m
        self.Iww = [[0] * m for _ in range(m)]
        for j1 in range(m):
            for j2 in range(m):
                self.Iww[j1][j2] = n
                self.names[n] = f"I.Iww[{j1}][{j2}]"
                n += 1
        self.Iwt, self.Itw = [], []
        for j in range(m):
            self.Iwt.append(n)
            self.names[n] = f"I.Iwt[{j}]"
            n += 1
        for j in range(m):
            self.Itw.append(n)
            self.names[n] = f"I.Itw[{j}]"
            n += 1
        self.Iwww = [[[0] * m for _ in range(m)] for _ in range(m)]
        for j1 in range(m):
            for j2 in range(m):
                for j3 in range(m):
                    self.Iwww[j1][j2][j3] = n
                    self.names[n] = f"I.Iwww[{j1}][{j2}][{j3}]"
                    n += 1
        self.count = n


def step_expressions(tp, scheme):
    """New state x_i + sum_alpha f_alpha,i(x) I_alpha as Exprs over StepVars (reference
    `sde_solver.py:186-215`, same grouping of the additions)."""
    d, m = tp.d, tp.m
    sv = StepVars(d, m)
    cf = tp.coefficients(scheme)
    dt = E.var(sv.dt)
    outs = []
    for i in range(d):
        inc = cf["f_t"][i] * dt
        for j in range(m):
            inc = inc + cf["f_w"][i][j] * E.var(sv.dW[j])
        new = E.var(i) + inc
        if scheme in ("milstein", "wagner_platen"):
            inc = E.ZERO
            for j1 in range(m):
                for j2 in range(m):
                    inc = inc + cf["f_ww"][i][j1][j2] * E.var(sv.Iww[j1][j2])
            new = new + inc
        if scheme == "wagner_platen":
            inc = E.ZERO
            for j in range(m):
                inc = inc + cf["f_tw"][i][j] * E.var(sv.Itw[j])
            for j in range(m):
                inc = inc + cf["f_wt"][i][j] * E.var(sv.Iwt[j])
            for j1 in range(m):
                for j2 in range(m):
                    for j3 in range(m):
                        inc = inc + cf["f_www"][i][j1][j2][j3] * E.var(sv.Iwww[j1][j2][j3])
            inc = inc + cf["f_tt"][i] * (dt * dt * E.HALF)
            new = new + inc
        outs.append(new)
    return outs, sv
