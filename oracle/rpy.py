"""Oracle: generalised Rotne-Prager-Yamakawa translational mobility `muTT(centres, radii)`.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).

The reference calls `pygrpy.jax_grpy_tensors.muTT` (un-vendored dependency `PyGRPY>=0.0.5`,
reference `requirements.txt:5`; call sites `examples/jfm.py:38,45`,
`examples/example_two_interacting_brownian_spheres.py:19,25`).  Its source is not available
here, so the published formulas are restated (Rotne & Prager 1969; Yamakawa 1970; unequal
and overlapping spheres: Zuk, Wajnryb, Mizerski, Szymczak, J. Fluid Mech. 741 (2014) R5),
viscosity eta = 1.  Pinned by the reference's shipped `single_trajectory.csv`
(far-field and overlapping branches both exercised) - see `tests/test_oracle_golden.py`.

Layout: mu[(3 i + alpha), (3 j + beta)].
"""
import numpy as np
import torch


def muTT_torch(centres, radii):
    """centres (N,3), radii (N,) -> (3N,3N); branch-free (torch.where) so it vmaps."""
    n = centres.shape[0]
    dtype = centres.dtype
    eye3 = torch.eye(3, dtype=dtype)
    rij = centres[:, None, :] - centres[None, :, :]                       # (N,N,3)
    offdiag = ~torch.eye(n, dtype=torch.bool)
    dist2 = (rij ** 2).sum(-1)
    dist = torch.sqrt(torch.where(offdiag, dist2, torch.ones_like(dist2)))
    rhat = rij / dist[..., None]
    rr = rhat[..., :, None] * rhat[..., None, :]                          # (N,N,3,3)
    ai = radii[:, None].expand(n, n)
    aj = radii[None, :].expand(n, n)
    pi = np.pi

    # far field
    s2 = ai * ai + aj * aj
    far_i = (1.0 / (8 * pi * dist)) * (1 + s2 / (3 * dist2_safe(dist)))
    far_r = (1.0 / (8 * pi * dist)) * (1 - s2 / dist2_safe(dist))
    # overlapping
    dm2 = (ai - aj) ** 2
    d3 = dist ** 3
    pref = 1.0 / (6 * pi * ai * aj)
    ov_i = pref * (16 * d3 * (ai + aj) - (dm2 + 3 * dist * dist) ** 2) / (32 * d3)
    ov_r = pref * 3 * (dm2 - dist * dist) ** 2 / (32 * d3)
    # one sphere inside the other
    im_i = 1.0 / (6 * pi * torch.maximum(ai, aj))
    im_r = torch.zeros_like(im_i)

    is_far = dist > ai + aj
    is_ov = dist > torch.abs(ai - aj)
    ci = torch.where(is_far, far_i, torch.where(is_ov, ov_i, im_i))
    cr = torch.where(is_far, far_r, torch.where(is_ov, ov_r, im_r))
    self_i = 1.0 / (6 * pi * ai)
    ci = torch.where(offdiag, ci, self_i)
    cr = torch.where(offdiag, cr, torch.zeros_like(cr))
    blocks = ci[..., None, None] * eye3 + cr[..., None, None] * rr        # (N,N,3,3)
    return blocks.permute(0, 2, 1, 3).reshape(3 * n, 3 * n)


def dist2_safe(dist):
    return dist * dist


def muTT_numpy(centres, radii):
    c = torch.as_tensor(np.asarray(centres, dtype=np.float64))
    r = torch.as_tensor(np.asarray(radii, dtype=np.float64))
    return muTT_torch(c, r).numpy()
