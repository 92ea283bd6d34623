"""Host model of the register-tiled bead kernel (csrc/sde_bead_kernel.cuh, SDE_BEAD_REGTILE): the same tile ownership,
fragment layouts, buffer indexing and phase order, executed lane by lane in numpy.  Checks that the data flow of the
kernel - mobility tiles built from the pair table in accumulator-fragment order, the factorisation of the diagonal tile
with its inverse as rank-one tensor-core updates, the row solves as tensor-core products with a permuted contraction index, the
double-buffered panel, the catch-up of the diagonal tiles, noise folded into the row solves, drift from the pair
table - computes chol(mu) xi and mu F.  (The arithmetic itself is checked on the GPU against the oracle.)"""
import re

import numpy as np
import pytest

from oracle import rpy

LANES = np.arange(32)
ROW, QD = LANES >> 2, LANES & 3
OFF = 8 * ROW + 2 * QD                      # a lane's two entries of a row-major 8x8 tile


def dmma884(c, a, b):
    """mma.m8n8k4.row.col.f64 on 32-lane fragments: c (32, 2), a (32,) = A[l/4][l%4], b (32,) = B[l%4][l/4]."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    A[ROW, QD] = a
    B[QD, ROW] = b
    P = A @ B
    out = c.copy()
    out[:, 0] += P[ROW, 2 * QD]
    out[:, 1] += P[ROW, 2 * QD + 1]
    return out


def shfl(v, src):
    """shfl.sync.idx of a 32-lane register: lane l reads lane src[l] (or the one lane src)."""
    return v[np.broadcast_to(src, (32,))]


def factor_rank_one(frag, rng):
    """`factor_rank_one` of the header, lane by lane: one tile (S in the lower triangle, T = B^T in the strict upper one),
    one DMMA per pivot on the masked column register, the next pivot predicted from two broadcasts, reciprocals with the
    accuracy of rcp.approx + the two-term correction.  Returns the fragments of L and of inv(L) TRANSPOSED."""
    i, j0 = ROW, 2 * QD
    cx = np.where(j0 > i, 0.0, frag[:, 0])
    cy = np.where(j0 + 1 > i, 0.0, frag[:, 1])

    def recip(d):
        a = (1.0 / d) * (1.0 + 2.0 ** -22 * rng.uniform(-1, 1))          # rcp.approx: ~22 good bits
        r = 1.0 - d * a
        a = a + a * r
        return a + a * (r * r)

    d = shfl(cx, 0)
    px, py = d.copy(), np.ones(32)
    a = recip(d)
    for k in range(7):
        ck = cy if k & 1 else cx
        mine = QD == (k >> 1)
        ub = np.where(mine & (i > k), ck, 0.0)
        ua = np.where(mine, np.where(i == k, 1.0, ck), 0.0)
        P = dmma884(np.zeros((32, 2)), ua, ub)
        e = shfl(cy if (k + 1) & 1 else cx, 4 * (k + 1) + ((k + 1) >> 1))
        w = shfl(ck, 4 * (k + 1) + (k >> 1))
        dn = e - a * w * w
        if (k + 1) & 1:
            py = np.where(j0 == k, dn, py)
        else:
            px = np.where(j0 == k + 1, dn, px)
        cx = np.where((i > k) & (j0 > i), cx, cx - a * P[:, 0])
        cy = np.where((i > k) & (j0 + 1 > i), cy, cy - a * P[:, 1])
        a = recip(dn)
    sx, sy = 1.0 / np.sqrt(px), 1.0 / np.sqrt(py)
    L = np.stack([np.where(j0 <= i, cx * sx, 0.0), np.where(j0 + 1 <= i, cy * sy, 0.0)], axis=1)
    WT = np.stack([np.where(j0 > i, cx * sx, np.where(j0 == i, sx, 0.0)),
                   np.where(j0 + 1 > i, cy * sy, np.where(j0 + 1 == i, sy, 0.0))], axis=1)
    return L, WT


def test_rank_one_factorisation_of_a_tile():
    """L L^T = C and inv(L) L = 1 for random SPD tiles, from the lane-level model of the tensor-core factorisation."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        G = rng.standard_normal((8, 12))
        C = G @ G.T / 12 + 0.5 * np.eye(8)
        frag = np.stack([C[ROW, 2 * QD], C[ROW, 2 * QD + 1]], axis=1)
        Lf, WTf = factor_rank_one(frag, rng)
        L, WT = np.zeros((8, 8)), np.zeros((8, 8))
        L[ROW, 2 * QD], L[ROW, 2 * QD + 1] = Lf[:, 0], Lf[:, 1]
        WT[ROW, 2 * QD], WT[ROW, 2 * QD + 1] = WTf[:, 0], WTf[:, 1]
        assert np.allclose(L, np.linalg.cholesky(C), rtol=1e-12, atol=1e-14)
        assert np.allclose(WT.T @ L, np.eye(8), rtol=0, atol=1e-12)


def slots_of(b, NT):
    """Bulk warp b: slot s -> (I, J) or None, as the RT_IN_HI / RT_SLOT_J / RT_SLOT_OK macros of the header."""
    hi, lo = NT - 1 - b, b + 1
    out = []
    for s in range(NT):
        in_hi = s < hi
        ok = in_hi or lo < hi
        out.append(((hi if in_hi else lo), (s if in_hi else NT - 1 - s)) if ok else None)
    return hi, lo, out


@pytest.mark.parametrize("NT", list(range(2, 17)))
def test_tile_ownership(NT):
    """Every off-diagonal tile of the lower block triangle is owned by exactly one slot of one bulk warp, a slot's column
    is one of two compile-time values, and the active slots of a panel form the contiguous range the chunk skipping of
    the trailing update assumes."""
    nbw = max(1, NT // 2)
    seen = []
    for b in range(nbw):
        hi, lo, sl = slots_of(b, NT)
        for s, t in enumerate(sl):
            if t is not None:
                I, J = t
                assert 0 <= J < I < NT and J in (s, NT - 1 - s)
                seen.append(t)
        for p in range(NT - 1):
            active = [s for s, t in enumerate(sl) if t is not None and t[1] > p]
            top = max(hi, NT - 1 - p)
            assert all(p < s < top for s in active)
            # every valid slot inside the range is active (invalid ones multiply by zero)
            assert active == [s for s in range(p + 1, top) if sl[s] is not None]
            solved = [t for t in sl if t is not None and t[1] == p]
            assert sorted(I for I, _ in solved) == sorted(([hi] if hi > p else []) + ([lo] if p < lo < hi else []))
    assert sorted(seen) == [(I, J) for I in range(NT) for J in range(I)]


@pytest.mark.parametrize("nb", [5, 6, 11, 40, 42])
def test_register_tiled_cholesky_data_flow(nb):
    rng = np.random.default_rng(nb)
    n3, NP = 3 * nb, -(-3 * nb // 8) * 8
    NT = NP // 8
    NOFF = NT * (NT - 1) // 2
    nbw = max(1, NT // 2)
    a = 0.05
    k = 2 * np.pi * np.arange(nb) / nb
    loc = 0.3 * np.stack([np.cos(k), np.sin(k), 0.2 * np.sin(2 * k)], 1) + 0.01 * rng.standard_normal((nb, 3))
    loc[1] = loc[0] + np.array([0.04, 0.0, 0.0])            # one overlapping pair (|r| < 2a)
    pos = np.zeros(NP)
    pos[:n3] = loc.ravel()
    frc = rng.standard_normal(n3)
    xi = np.zeros(NP)
    xi[:n3] = rng.standard_normal(n3)
    mu_self = 1.0 / (6 * np.pi * a)

    # value table: [0, 1, mu_self, -] then 6 entries per bead pair q = i (i - 1) / 2 + j, from `rpy_pair` (one reciprocal
    # square root, no division)
    val = np.zeros(4 + 6 * (nb * (nb - 1) // 2))
    val[1], val[2] = 1.0, mu_self
    for q in range(nb * (nb - 1) // 2):
        i = int((np.sqrt(np.float32(8.0 * q + 1.0)) + 1.0) * 0.5)
        while i * (i - 1) // 2 > q:
            i -= 1
        while (i + 1) * i // 2 <= q:
            i += 1
        j = q - i * (i - 1) // 2
        assert 0 <= j < i < nb
        d = loc[i] - loc[j]
        r2 = d @ d
        inv_r = 1 / np.sqrt(r2)
        if r2 > 4 * a * a:
            pref = inv_r * (0.125 / np.pi)
            ci, crp = pref * (1 + (2 * a * a / 3) * inv_r ** 2), pref * (1 - 2 * a * a * inv_r ** 2) * inv_r ** 2
        else:
            ci, crp = (mu_self / a) * (a - (9 / 32) * r2 * inv_r), (mu_self / a) * (3 / 32) * inv_r
        val[4 + 6 * q:10 + 6 * q] = [crp * d[0] * d[0] + ci, crp * d[0] * d[1], crp * d[0] * d[2],
                                     crp * d[1] * d[1] + ci, crp * d[1] * d[2], crp * d[2] * d[2] + ci]

    def mobility_index(R, C):
        if R >= n3 or C >= n3:
            return 1 if R == C else 0
        bi, al, bj, be = R // 3, R % 3, C // 3, C % 3
        if bi == bj:
            return 2 if al == be else 0
        hi, lo = max(bi, bj), min(bi, bj)
        a0, a1 = min(al, be), max(al, be)
        return 4 + 6 * (hi * (hi - 1) // 2 + lo) + (a1 if a0 == 0 else a0 + a1 + 1)

    mu_ref = np.asarray(rpy.muTT_numpy(loc, np.full(nb, a)))
    mu_model = np.array([[val[mobility_index(R, C)] for C in range(n3)] for R in range(n3)])
    assert np.allclose(mu_model, mu_ref, rtol=1e-12, atol=1e-14)

    def fragment(I, J):
        return np.array([[val[mobility_index(8 * I + (l >> 2), 8 * J + 2 * (l & 3) + e)] for e in (0, 1)] for l in LANES])

    # meta numbering of the header: off-diagonal tile (I, J) is I (I - 1) / 2 + J, diagonal tile k is NOFF + k
    assert sorted(I * (I - 1) // 2 + J for I in range(NT) for J in range(I)) == list(range(NOFF))

    bulk = []
    for b in range(nbw):
        hi, lo, sl = slots_of(b, NT)
        bulk.append(dict(hi=hi, lo=lo, slots=sl, acc=[fragment(*t) if t is not None else np.zeros((32, 2)) for t in sl],
                         dgh=fragment(hi, hi), dgl=fragment(lo, lo), yph=np.zeros(32), ypl=np.zeros(32),
                         nah=np.zeros((32, 2)), nal=np.zeros((32, 2))))
    Pan = np.zeros((NT, 64))
    Linv, nextd = np.zeros(64), np.zeros(64)
    yd4, yb = np.zeros(4 * NP), np.zeros(NP)
    cur = fragment(0, 0)
    for p in range(NT):
        # chain warp: factor, publish the inverse, the diagonal rows' terms of y (four partial sums per row)
        L, WT = factor_rank_one(cur, rng)
        Linv[8 * (2 * QD) + ROW], Linv[8 * (2 * QD + 1) + ROW] = WT[:, 0], WT[:, 1]      # published row-major: inv(L)[j][i]
        yd4[4 * 8 * p + LANES] = L[:, 0] * xi[8 * p + 2 * QD] + L[:, 1] * xi[8 * p + 2 * QD + 1]
        if p + 1 == NT:
            break
        # ---- barrier 1 ----
        lf = Linv[OFF], Linv[OFF + 1]
        x2 = xi[8 * p + 2 * QD], xi[8 * p + 2 * QD + 1]
        for w in bulk:
            # `solve_column<P>`: slot P of the hi row, slot NT-1-P of the lo row
            todo = []
            if p < w["hi"]:
                todo.append((p, w["hi"], "h"))
            if w["lo"] < w["hi"] and NT - 1 - p >= w["hi"]:
                todo.append((NT - 1 - p, w["lo"], "l"))
            for s, I, tag in todo:
                assert w["slots"][s] == (I, p)
                out = dmma884(dmma884(np.zeros((32, 2)), w["acc"][s][:, 0], lf[0]), w["acc"][s][:, 1], lf[1])
                Pan[I, OFF], Pan[I, OFF + 1] = out[:, 0], out[:, 1]
                w["na" + tag] = -out
                w["dg" + tag] = dmma884(dmma884(w["dg" + tag], -out[:, 0], out[:, 0]), -out[:, 1], out[:, 1])
                if I == p + 1:
                    nextd[OFF], nextd[OFF + 1] = w["dg" + tag][:, 0], w["dg" + tag][:, 1]
                w["yp" + tag] += out[:, 0] * x2[0] + out[:, 1] * x2[1]
        # ---- barrier 2 ----
        cur = np.stack([nextd[OFF], nextd[OFF + 1]], axis=1)
        for w in bulk:
            top = max(w["hi"], NT - 1 - p)
            for c0 in range(0, NT, 5):
                if c0 + 4 > p and c0 < top:
                    for s in range(c0, min(c0 + 5, NT)):
                        in_hi = s < w["hi"]
                        J = s if in_hi else NT - 1 - s
                        on = w["slots"][s] is not None and J > p
                        na = (w["nah"] if in_hi else w["nal"]) * (1.0 if on else 0.0)
                        w["acc"][s] = dmma884(dmma884(w["acc"][s], na[:, 0], Pan[J, OFF]), na[:, 1], Pan[J, OFF + 1])
    for w in bulk:
        for name, I in (("yph", w["hi"]), ("ypl", w["lo"])):
            if name == "ypl" and not w["lo"] < w["hi"]:
                continue
            v = w[name] + w[name][LANES ^ 1]
            v = v + v[LANES ^ 2]
            yb[8 * I + ROW[QD == 0]] = v[QD == 0]
    Lref = np.linalg.cholesky(mu_ref)
    yd = yd4.reshape(NP, 4).sum(1)
    assert np.allclose((yb + yd)[:n3], Lref @ xi[:n3], rtol=1e-10, atol=1e-12)

    # drift from the pair table, thread (i, chunk); the pair {i, i} reads another pair's block times zero
    from stochastic_b200.sde_solver import bead_regtile_block
    dch = max(3, min(6, bead_regtile_block(nb) // nb))
    part = np.zeros((dch, NP))
    for tid in range(dch * nb):
        i, chunk = tid % nb, tid // nb
        o = mu_self * frc[3 * i:3 * i + 3] if chunk == 0 else np.zeros(3)
        for j in range(chunk, nb, dch):
            jj = (1 if i == 0 else 0) if j == i else j
            on = 0.0 if j == i else 1.0
            hi, lo = max(i, jj), min(i, jj)
            blk = val[4 + 6 * (hi * (hi - 1) // 2 + lo):][:6]
            B = np.array([[blk[0], blk[1], blk[2]], [blk[1], blk[3], blk[4]], [blk[2], blk[4], blk[5]]])
            o = o + B @ (on * frc[3 * j:3 * j + 3])
        part[chunk, 3 * i:3 * i + 3] = o
    assert np.allclose(part.sum(0)[:n3], mu_ref @ frc, rtol=1e-11, atol=1e-13)


def test_shared_memory_size_and_block_formulas_match_the_header():
    from stochastic_b200.sde_solver import bead_regtile_block, bead_regtile_smem_bytes
    # (8 + drift chunks) NP + value table + overlap coefficients + 2 NT tiles + inverse block + next diagonal tile
    assert bead_regtile_smem_bytes(40) == 8 * (14 * 120 + 4684 + 780 + 1920 + 128)
    assert bead_regtile_smem_bytes(5) == 8 * (14 * 16 + 64 + 10 + 256 + 128)
    assert bead_regtile_block(40) == 256 and bead_regtile_block(42) == 288 and bead_regtile_block(5) == 64
    assert 2 * (bead_regtile_smem_bytes(40) + 1024) <= 227 * 1024
    # with the twist term: + [NB / 2 - 1][NB][6] writhe gradients of the second segments; still two CTAs per SM
    assert bead_regtile_smem_bytes(40, twist=True) == bead_regtile_smem_bytes(40) + 8 * 19 * 40 * 6
    assert 2 * (bead_regtile_smem_bytes(40, twist=True) + 2 * 1024 + 64) <= 228 * 1024
    import os
    import stochastic_b200
    header = open(os.path.join(os.path.dirname(stochastic_b200.__file__), "csrc", "sde_bead_kernel.cuh")).read()
    # the carve-up the formula mirrors
    for piece in ("RT_OFF_PART + RT_DCH * SDE_NP", "RT_OFF_YD + 4 * SDE_NP", "RT_OFF_VAL + RT_NVAL",
                  "RT_OFF_OVL + RT_NPAIR + (RT_NPAIR & 1)", "RT_OFF_PAN + 2 * RT_NT * 64", "RT_OFF_LINV + 64", "RT_OFF_NEXTD + 64", "RT_OFF_TWSC + RT_TWSC"):
        assert piece in header, piece
    assert re.search(r"SDE_BLOCK == 32 \* \(1 \+ RT_NBW\)", header)
