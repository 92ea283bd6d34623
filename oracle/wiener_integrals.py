"""Oracle: unit-step Wiener increments and Kloeden-Platen iterated integrals.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).

Restates `pychastic/vectorized_I_generation.py` of the reference in numpy:

* `get_wiener_integrals`  <- reference `vectorized_I_generation.py:243-473`
* `make_C_mat`            <- reference `vectorized_I_generation.py:22-49`
* `make_D_mat_loopy`      <- reference `vectorized_I_generation.py:55-132` (readable spec)
* `make_D_mat`            <- reference `vectorized_I_generation.py:135-236` (what runs)

including the three corrections to Kloeden-Platen eq. (5.8.11) recorded in the
reference's `found_typos.md:5-8`.  Random draws come from `oracle.jax_prng` so the stream is
the reference's stream (key order: Milstein xi,mu,eta,zeta <- split(key,4); Wagner-Platen
xi,zeta,eta,mu,phi <- split(key,5)).  All outputs are for a UNIT time step.
"""
import numpy as np

from . import jax_prng


def make_C_mat(eta, zeta):
    """C^p_{j1 j2} for one step; eta, zeta of shape (m, p)."""
    eta = np.asarray(eta)
    zeta = np.asarray(zeta)
    m, p = eta.shape
    C = np.zeros((m, m), dtype=np.result_type(eta.dtype, np.float32))
    for j1 in range(m):
        for j2 in range(m):
            acc = 0.0
            for l in range(1, p + 1):
                for r in range(1, p + 1):
                    if l == r:
                        continue
                    acc += r / (r * r - l * l) * (
                        zeta[j1, r - 1] * zeta[j2, l - 1] / l + eta[j1, r - 1] * eta[j2, l - 1] / r
                    )
            C[j1, j2] = -acc / (2 * np.pi ** 2)
    return C


def make_D_mat_loopy(eta, zeta):
    """D^p_{j1 j2 j3} for one step, plain loops; eta, zeta of shape (m, p)."""
    eta = np.asarray(eta, dtype=np.float64)
    zeta = np.asarray(zeta, dtype=np.float64)
    m, p = eta.shape

    def z(j, k):
        return zeta[j, k - 1] if 1 <= k <= p else 0.0

    def e(j, k):
        return eta[j, k - 1] if 1 <= k <= p else 0.0

    D = np.zeros((m, m, m))
    for j1 in range(m):
        for j2 in range(m):
            for j3 in range(m):
                s = 0.0
                for r in range(1, p + 1):
                    for l in range(1, p + 1):
                        s -= (z(j2, l) * (z(j3, l + r) * e(j1, r) - z(j3, r) * e(j1, l + r))
                              + e(j2, l) * (z(j1, r) * z(j3, l + r) + e(j1, r) * e(j3, l + r))
                              ) / (r * (l + r))
                for l in range(1, p + 1):
                    for r in range(1, l):
                        s += (z(j2, l) * (z(j1, r) * e(j3, l - r) + z(j3, l - r) * e(j1, r))
                              - e(j2, l) * (z(j1, r) * z(j3, l - r) - e(j1, r) * e(j3, l - r))
                              ) / (r * (l - r))
                for l in range(1, p + 1):
                    for r in range(l + 1, 2 * p + 1):
                        s += (z(j2, l) * (z(j3, r - l) * e(j1, r) - z(j1, r) * e(j3, r - l))
                              + e(j2, l) * (z(j1, r) * z(j3, r - l) + e(j1, r) * e(j3, r - l))
                              ) / (r * (r - l))
                D[j1, j2, j3] = s
    return D / (np.pi ** 2 * 2 ** 2.5)


def _pad(x, p):
    """(S, m, p) -> (S, m, 2p+1) with x at indices 1..p and zeros elsewhere."""
    out = np.zeros(x.shape[:-1] + (2 * p + 1,), dtype=x.dtype)
    out[..., 1:p + 1] = x
    return out


def batched_C_mat(eta, zeta):
    """C for a batch: eta, zeta (S, m, p) -> (S, m, m)."""
    S, m, p = eta.shape
    dt = eta.dtype
    l = np.arange(1, p + 1).reshape(p, 1).astype(np.float64)
    r = np.arange(1, p + 1).reshape(1, p).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        base = np.where(l != r, r / (r * r - l * l), 0.0)
    cz = (base / l).astype(dt)     # multiplies zeta[j1,r] zeta[j2,l]
    ce = (base / r).astype(dt)     # multiplies eta[j1,r]  eta[j2,l]
    acc = np.einsum("lr,sar,sbl->sab", cz, zeta, zeta) + np.einsum("lr,sar,sbl->sab", ce, eta, eta)
    return (dt.type(-1.0 / (2 * np.pi ** 2)) * acc).astype(dt)


def batched_D_mat(eta, zeta):
    """D for a batch: eta, zeta (S, m, p) -> (S, m, m, m); same sums as `make_D_mat_loopy`."""
    S, m, p = eta.shape
    dt = eta.dtype
    Z = _pad(zeta, p)
    E = _pad(eta, p)
    l1 = np.arange(1, p + 1).reshape(p, 1)
    r1 = np.arange(1, p + 1).reshape(1, p)
    r2 = np.arange(1, 2 * p + 1).reshape(1, 2 * p)
    zl, el = Z[..., 1:p + 1], E[..., 1:p + 1]                # [s, j, l]
    zr, er = Z[..., 1:p + 1], E[..., 1:p + 1]                # [s, j, r], r <= p
    zr2, er2 = Z[..., 1:2 * p + 1], E[..., 1:2 * p + 1]      # [s, j, r], r <= 2p (zero beyond p)

    # first sum, l, r in 1..p, index l + r
    c = (1.0 / (r1 * (l1 + r1))).astype(dt)
    zs, es = Z[..., l1 + r1], E[..., l1 + r1]                # [s, j, l, r]
    s1 = (np.einsum("lr,sbl,sclr,sar->sabc", c, zl, zs, er, optimize=True)
          - np.einsum("lr,sbl,scr,salr->sabc", c, zl, zr, es, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, el, zr, zs, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, el, er, es, optimize=True))

    # second sum, r < l, index l - r
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(r2 < l1, 1.0 / (r2 * np.where(r2 < l1, l1 - r2, 1)), 0.0).astype(dt)
    idx = np.clip(l1 - r2, 0, 2 * p)
    zd, ed = Z[..., idx], E[..., idx]
    s2 = (np.einsum("lr,sbl,sar,sclr->sabc", c, zl, zr2, ed, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, zl, er2, zd, optimize=True)
          - np.einsum("lr,sbl,sar,sclr->sabc", c, el, zr2, zd, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, el, er2, ed, optimize=True))

    # third sum, r > l (r up to 2p, reads beyond p are zero), index r - l
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(r2 > l1, 1.0 / (r2 * np.where(r2 > l1, r2 - l1, 1)), 0.0).astype(dt)
    idx = np.clip(r2 - l1, 0, 2 * p)
    zd, ed = Z[..., idx], E[..., idx]
    s3 = (np.einsum("lr,sbl,sclr,sar->sabc", c, zl, zd, er2, optimize=True)
          - np.einsum("lr,sbl,sar,sclr->sabc", c, zl, zr2, ed, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, el, zr2, zd, optimize=True)
          + np.einsum("lr,sbl,sar,sclr->sabc", c, el, er2, ed, optimize=True))

    return (dt.type(1.0 / (np.pi ** 2 * 2 ** 2.5)) * (-s1 + s2 + s3)).astype(dt)


def make_D_mat(eta, zeta):
    """Single-step D (m, p) -> (m, m, m) through the batched einsum form."""
    eta = np.asarray(eta, dtype=np.result_type(np.asarray(eta).dtype, np.float32))
    zeta = np.asarray(zeta, dtype=eta.dtype)
    return batched_D_mat(eta[None], zeta[None])[0]


def _fill_diag(mat, vec):
    out = mat.copy()
    m = mat.shape[-1]
    idx = np.arange(m)
    out[..., idx, idx] = vec
    return out


def draw_normals(key, steps, noise_terms, scheme, p=10, dtype=np.float32, layout="original"):
    """The raw normal draws of one key, in the reference's key order (dict of arrays with leading axis S)."""
    dt = np.dtype(dtype)
    m, S = noise_terms, steps

    def nrm(k, shape):
        return jax_prng.normal(k, shape, dt, layout)

    if m == 1:
        if scheme in ("euler", "milstein"):
            return {"u": nrm(key, (S, 1))}
        if scheme == "wagner_platen":
            u = nrm(key, (2, S, 1))
            return {"u0": u[0], "u1": u[1]}
    if scheme == "euler":
        return {"u": nrm(key, (S, m))}
    if scheme == "milstein":
        k1, k2, k3, k4 = jax_prng.split(key, 4, layout)
        return {"xi": nrm(k1, (S, 1, m))[:, 0], "mu": nrm(k2, (S, 1, m))[:, 0], "eta": nrm(k3, (S, p, m)),
                "zeta": nrm(k4, (S, p, m))}
    if scheme == "wagner_platen":
        k1, k2, k3, k4, k5 = jax_prng.split(key, 5, layout)
        return {"xi": nrm(k1, (S, m)), "zeta": nrm(k2, (S, p, m)), "eta": nrm(k3, (S, p, m)), "mu": nrm(k4, (S, m)),
                "phi": nrm(k5, (S, m))}
    raise NotImplementedError(scheme)


def assemble_integrals(draws, noise_terms, scheme, p=10, dtype=np.float32):
    """Draws (leading axis = any batch of steps) -> unit-step integrals, reference shapes with leading axis S."""
    dt = np.dtype(dtype)
    f = dt.type
    m = noise_terms
    if m == 1:
        if scheme == "euler":
            return {"d_w": draws["u"]}
        if scheme == "milstein":
            dW = draws["u"]
            return {"d_w": dW, "d_ww": (f(0.5) * (dW ** 2 - f(1)))[..., None]}
        if scheme == "wagner_platen":
            u0, u1 = draws["u0"], draws["u1"]
            dW = u0
            dZ = (f(0.5) * (u0 + f(3 ** -0.5) * u1))[..., None]
            return {
                "d_w": dW,
                "d_ww": (f(0.5) * (dW ** 2 - f(1)))[..., None],
                "d_wt": dZ,
                "d_tw": dW[..., None] - dZ,
                "d_www": (f(0.5) * (f(1.0 / 3.0) * dW ** 2 - f(1)) * dW)[..., None, None],
            }
    if scheme == "euler":
        return {"d_w": draws["u"]}

    rec = (1.0 / np.arange(1, p + 1)).astype(dt)
    if scheme == "milstein":
        x, u, eta, zeta = draws["xi"], draws["mu"], draws["eta"], draws["zeta"]
        rho = f(1 / 12 - (rec.astype(np.float64) ** 2).sum() / (2 * np.pi ** 2))
        a = f(np.sqrt(2)) * x[:, None, :] + eta                                   # (S, p, m)
        series = np.einsum("r,sri,srj->sij", rec, zeta, a) - np.einsum("r,sri,srj->sij", rec, a, zeta)
        I = (x[:, :, None] * x[:, None, :] / f(2)
             + np.sqrt(rho) * (x[:, :, None] * u[:, None, :] - u[:, :, None] * x[:, None, :])
             + f(1.0 / (2 * np.pi)) * series).astype(dt)
        I = _fill_diag(I, f(0.5) * (x ** 2 - f(1)))
        return {"d_w": x, "d_ww": I}

    if scheme == "wagner_platen":
        xi, zeta, eta, mu, phi = draws["xi"], draws["zeta"], draws["eta"], draws["mu"], draws["phi"]
        rec64 = rec.astype(np.float64)
        rho = f(1 / 12 - (rec64 ** 2).sum() / (2 * np.pi ** 2))
        alpha = f(np.pi ** 2 / 180 - (0.5 / np.pi ** 2) * (rec64 ** 4).sum())
        pi = f(np.pi)
        a = (-f(np.sqrt(2)) / pi * np.einsum("r,srj->sj", rec, zeta) - f(2) * np.sqrt(rho) * mu).astype(dt)
        b = (f(1) / (f(np.sqrt(2)) * pi) * np.einsum("r,srj->sj", rec ** 2, eta) + np.sqrt(alpha) * phi).astype(dt)
        A = (f(0.5) / pi * (np.einsum("r,sri,srj->sij", rec, zeta, eta)
                            - np.einsum("r,sri,srj->sij", rec, eta, zeta))).astype(dt)
        B = (f(1.0 / (4 * np.pi ** 2)) * (np.einsum("r,sri,srj->sij", rec ** 2, zeta, zeta)
                                          + np.einsum("r,sri,srj->sij", rec ** 2, eta, eta))).astype(dt)
        eta_t = np.ascontiguousarray(eta.transpose(0, 2, 1))
        zeta_t = np.ascontiguousarray(zeta.transpose(0, 2, 1))
        C = batched_C_mat(eta_t, zeta_t)
        D = batched_D_mat(eta_t, zeta_t)

        d_wt = f(0.5) * (xi + a)                                      # (S, m)  I_(j,0)
        d_tw = xi - d_wt                                              # (S, m)  I_(0,j)
        xx = xi[:, :, None] * xi[:, None, :]
        J_ww = (f(0.5) * xx - f(0.5) * (xi[:, :, None] * a[:, None, :] - xi[:, None, :] * a[:, :, None]) + A)
        d_ww = _fill_diag(J_ww, f(0.5) * (xi ** 2 - f(1)))
        J_tww = (f(1.0 / 6.0) * xx
                 - (f(1) / pi) * xi[:, None, :] * b[:, :, None]
                 + B
                 - f(0.25) * a[:, None, :] * xi[:, :, None]
                 + (f(0.5) / pi) * xi[:, :, None] * b[:, None, :]
                 + C
                 + f(0.5) * A)
        J = (xi[:, :, None, None] * J_tww[:, None, :, :]
             + f(0.5) * a[:, :, None, None] * J_ww[:, None, :, :]
             + (f(0.5) / pi) * b[:, :, None, None] * xi[:, None, :, None] * xi[:, None, None, :]
             - xi[:, None, :, None] * B[:, :, None, :]
             + xi[:, None, None, :] * (f(0.5) * A - C.transpose(0, 2, 1))[:, :, :, None]
             + D)
        eye = np.eye(m, dtype=dt)
        d_www = J - f(0.5) * (eye[None, :, :, None] * d_tw[:, None, None, :]
                              + eye[None, None, :, :] * d_wt[:, :, None, None])
        return {
            "d_w": xi,
            "d_ww": d_ww.astype(dt),
            "d_wt": d_wt[:, :, None].astype(dt),
            "d_tw": d_tw[:, None, :].astype(dt),
            "d_www": d_www.astype(dt),
        }
    raise NotImplementedError(scheme)


def get_wiener_integrals(key, steps=1, noise_terms=1, scheme="euler", p=10,
                         dtype=np.float32, layout="original"):
    """Unit-step integrals for `steps` steps; same dict keys and shapes as the reference."""
    draws = draw_normals(key, steps, noise_terms, scheme, p, dtype, layout)
    out = assemble_integrals(draws, noise_terms, scheme, p, dtype)
    if noise_terms > 1 and scheme in ("milstein", "wagner_platen"):
        out["d_w"] = out["d_w"].squeeze()        # the reference's xi.squeeze() (drops the step axis when steps == 1)
    return out


def get_wiener_integrals_many(keys, steps, noise_terms, scheme, p=10, dtype=np.float32, layout="original"):
    """One call per key for the draws (the reference vmaps over keys), ONE batched assembly for all keys.
    Returns arrays with leading axes (n_keys, steps)."""
    per = [draw_normals(k, steps, noise_terms, scheme, p, dtype, layout) for k in keys]
    cat = {name: np.concatenate([d[name] for d in per], axis=0) for name in per[0]}
    flat = assemble_integrals(cat, noise_terms, scheme, p, dtype)
    n = len(keys)
    return {k: v.reshape((n, steps) + v.shape[1:]) for k, v in flat.items()}
