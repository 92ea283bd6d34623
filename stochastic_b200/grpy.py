"""Generalised Rotne-Prager-Yamakawa translational mobility `muTT(centres, radii)`.

Built-in counterpart of `pygrpy.jax_grpy_tensors.muTT`, which the reference's bead examples
call inside drift and noise (reference `examples/example_two_interacting_brownian_spheres.py:19,25`,
`examples/jfm.py:38,45`).  Written against `stochastic_b200.numpy`, so it can be used inside
traced drift/noise callables (every pair branch becomes a select in the emitted kernel) as well
as on concrete numpy input.

Formulas (viscosity 1): far field Rotne-Prager-Yamakawa for unequal radii; overlapping and
immersed spheres after Zuk, Wajnryb, Mizerski, Szymczak, J. Fluid Mech. 741 (2014) R5.
Layout: mu[3 i + alpha, 3 j + beta].
"""
import math

import numpy as np

from . import config, jnp
from .trace import expr as E


def muTT(centres, radii):
    centres = jnp._as_obj(centres)
    radii = np.asarray(jnp._host(radii), dtype=np.float64)
    n = len(radii)
    traced = jnp._is_traced(centres)
    mu = np.zeros((3 * n, 3 * n), dtype=object if traced else np.float64)
    pi = math.pi
    for i in range(n):
        for a in range(3):
            mu[3 * i + a, 3 * i + a] = 1.0 / (6 * pi * radii[i])
    for i in range(n):
        for j in range(i + 1, n):
            ai, aj = float(radii[i]), float(radii[j])
            rv = [centres[i][k] - centres[j][k] for k in range(3)]
            r2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]
            s2 = ai * ai + aj * aj
            dm2 = (ai - aj) ** 2
            if traced and E.TRACE_F64 and config.get("library_rsqrt"):
                # fp64 kernels: every power of the distance from ONE reciprocal square root, no division (the same values
                # within rounding; an fp64 division or square root costs ~25-50 GPU instructions, rsqrt ~12)
                inv_r = E.unary("rsqrt", E.wrap(r2))
                r = r2 * inv_r
                inv_r2 = inv_r * inv_r
                pref = inv_r * (1.0 / (8 * pi))
                far_i = pref * (1.0 + (s2 / 3.0) * inv_r2)
                far_r = pref * (1.0 - s2 * inv_r2)
                r3 = r2 * r
                q = (1.0 / (6 * pi * ai * aj) / 32.0) * (inv_r * inv_r2)
            else:
                r = jnp.sqrt(r2)
                inv_r = 1.0 / r
                # far field
                pref = inv_r * (1.0 / (8 * pi))
                far_i = pref * (1.0 + (s2 / 3.0) / r2)
                far_r = pref * (1.0 - s2 / r2)
                r3 = r2 * r
                q = 1.0 / (6 * pi * ai * aj) / (32.0 * r3)
            # overlapping spheres
            ov_i = q * (16.0 * r3 * (ai + aj) - (dm2 + 3.0 * r2) ** 2)
            ov_r = q * 3.0 * (dm2 - r2) ** 2
            # smaller sphere immersed in the larger one
            im_i = 1.0 / (6 * pi * max(ai, aj))
            is_far = jnp.greater(r, ai + aj)
            is_ov = jnp.greater(r, abs(ai - aj))
            ci = jnp.where(is_far, far_i, jnp.where(is_ov, ov_i, im_i))
            cr = jnp.where(is_far, far_r, jnp.where(is_ov, ov_r, 0.0))
            rh = [v * inv_r for v in rv]
            for a in range(3):
                for b in range(3):
                    v = cr * rh[a] * rh[b]
                    if a == b:
                        v = v + ci
                    mu[3 * i + a, 3 * j + b] = v
                    mu[3 * j + b, 3 * i + a] = v
    return mu.view(jnp.TArray) if traced else mu
