"""Host model of the register-tiled bead kernel (csrc/sde_bead_kernel.cuh, SDE_BEAD_REGTILE): the same tile map,
fragment layouts, buffer indexing and phase order, executed lane by lane in numpy.  Checks that the data flow of the
kernel - mobility tiles built from the pair table in accumulator-fragment order, raw buffer, diagonal factor, row
solves written in operand-fragment order, DMMA trailing updates, noise folded into the row solves, drift from the pair
table - computes chol(mu) xi and mu F.  (The arithmetic itself is checked on the GPU against the oracle.)"""
import numpy as np
import pytest

from oracle import rpy


def dmma884(c, a, b):
    """mma.m8n8k4.row.col.f64 on 32-lane fragments: c (32, 2), a (32,), b (32,)."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for l in range(32):
        A[l >> 2, l & 3] = a[l]
        B[l & 3, l >> 2] = b[l]
    P = A @ B
    out = c.copy()
    for l in range(32):
        out[l, 0] += P[l >> 2, 2 * (l & 3)]
        out[l, 1] += P[l >> 2, 2 * (l & 3) + 1]
    return out


@pytest.mark.parametrize("nb", [5, 6, 11, 40])
def test_register_tiled_cholesky_data_flow(nb):
    rng = np.random.default_rng(nb)
    n3, NP = 3 * nb, -(-3 * nb // 8) * 8
    NT = NP // 8
    ntiles = NT * (NT + 1) // 2
    tpw = (ntiles + 3) // 4
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

    # pair table (ci, cr / r^2), q = i (i - 1) / 2 + j
    PT = np.zeros((nb * (nb - 1) // 2, 2))
    for q in range(len(PT)):
        i = int((np.sqrt(np.float32(8.0 * q + 1.0)) + 1.0) * 0.5)
        while i * (i - 1) // 2 > q:
            i -= 1
        while (i + 1) * i // 2 <= q:
            i += 1
        j = q - i * (i - 1) // 2
        assert 0 <= j < i < nb
        d = loc[i] - loc[j]
        r2 = d @ d
        rr = np.sqrt(r2)
        if rr > 2 * a:
            pref = 1 / (8 * np.pi * rr)
            ci, cr = pref * (1 + 2 * a * a / (3 * r2)), pref * (1 - 2 * a * a / r2)
        else:
            qq = 1 / (6 * np.pi * a * a) / (32 * r2 * rr)
            ci, cr = qq * (16 * r2 * rr * 2 * a - 9 * r2 * r2), qq * 3 * r2 * r2
        PT[q] = ci, cr / r2

    def entry(R, C):
        if R >= n3 or C >= n3:
            return 1.0 if R == C else 0.0
        bi, al, bj, be = R // 3, R % 3, C // 3, C % 3
        if bi == bj:
            return mu_self if al == be else 0.0
        hi, lo = max(bi, bj), min(bi, bj)
        ci, crp = PT[hi * (hi - 1) // 2 + lo]
        return crp * (pos[3 * bi + al] - pos[3 * bj + al]) * (pos[3 * bi + be] - pos[3 * bj + be]) + (ci if al == be else 0.0)

    mu_ref = np.asarray(rpy.muTT_numpy(loc, np.full(nb, a)))
    mu_model = np.array([[entry(R, C) for C in range(n3)] for R in range(n3)])
    assert np.allclose(mu_model, mu_ref, rtol=1e-12, atol=1e-14)

    # tile map
    tmap = []
    for s in range(ntiles):
        J, rem = 0, s
        while rem >= NT - J:
            rem -= NT - J
            J += 1
        tmap.append((J + rem, J))
    assert sorted(tmap) == sorted((I, J) for J in range(NT) for I in range(J, NT)) and tmap[:NT] == [(I, 0) for I in range(NT)]

    lanes = np.arange(32)
    acc = np.zeros((4, tpw, 32, 2))
    for w in range(4):
        for kk in range(tpw):
            sl = w + 4 * kk
            if sl < ntiles:
                I, J = tmap[sl]
                for l in lanes:
                    R, C = 8 * I + (l >> 2), 8 * J + 2 * (l & 3)
                    acc[w, kk, l] = entry(R, C), entry(R, C + 1)
    Raw = np.zeros(NT * 80)
    Pan = np.full(NT * 64, np.nan)
    y = np.zeros(128)

    def store_raw(I, frag):
        for l in lanes:
            o = I * 80 + (l >> 2) * 10 + 2 * (l & 3)
            Raw[o:o + 2] = frag[l]
    for w in range(4):
        for kk in range(tpw):
            if w + 4 * kk < NT:
                store_raw(w + 4 * kk, acc[w, kk])
    for p in range(NT):
        row0 = 8 * p
        D = np.array([[Raw[p * 80 + i * 10 + j] if j <= i else 0.0 for j in range(8)] for i in range(8)])
        dinv = np.zeros(8)
        for kq in range(8):
            dinv[kq] = 1 / np.sqrt(D[kq, kq])
            D[kq:, kq] *= dinv[kq]
            for j in range(kq + 1, 8):
                D[j:, j] -= D[j:, kq] * D[j, kq]
        for tid in range(row0, NP):
            I, r = tid >> 3, tid & 7
            if I == p:
                v = D[r].copy()
            else:
                v = Raw[I * 80 + r * 10: I * 80 + r * 10 + 8].copy()
                for kq in range(8):
                    v[kq] *= dinv[kq]
                    v[kq + 1:] -= v[kq] * D[kq + 1:, kq]
            y[tid] += v @ xi[row0:row0 + 8]
            if I > p:
                dst = I * 64 + r * 4
                Pan[dst:dst + 4] = v[:4]
                Pan[dst + 32:dst + 36] = v[4:]
        for w in range(4):
            for kk in range(tpw):
                sl = w + 4 * kk
                if sl < ntiles:
                    I, J = tmap[sl]
                    if J > p:
                        for h in range(2):
                            acc[w, kk] = dmma884(acc[w, kk], -Pan[I * 64 + 32 * h + lanes], Pan[J * 64 + 32 * h + lanes])
                        if J == p + 1:
                            store_raw(I, acc[w, kk])
    L = np.linalg.cholesky(mu_ref)
    assert np.allclose(y[:n3], L @ xi[:n3], rtol=1e-10, atol=1e-12)

    # drift from the pair table, thread (i, chunk)
    part = np.zeros((3, NP))
    for tid in range(3 * nb):
        i, chunk = tid % nb, tid // nb
        o = mu_self * frc[3 * i:3 * i + 3] if chunk == 0 else np.zeros(3)
        for j in range(chunk, nb, 3):
            if j == i:
                continue
            hi, lo = max(i, j), min(i, j)
            ci, crp = PT[hi * (hi - 1) // 2 + lo]
            d = loc[i] - loc[j]
            o = o + crp * (d @ frc[3 * j:3 * j + 3]) * d + ci * frc[3 * j:3 * j + 3]
        part[chunk, 3 * i:3 * i + 3] = o
    assert np.allclose(part.sum(0)[:n3], mu_ref @ frc, rtol=1e-11, atol=1e-13)


def test_shared_memory_size_formula_matches_the_header():
    from stochastic_b200.sde_solver import bead_regtile_smem_bytes
    # RT_OFF_* of csrc/sde_bead_kernel.cuh for N = 40: 6 NP + value table (4 + 6 NPAIR) + 80 NT + 64 NT + 72
    # + tile maps (3 x 36 bytes); three CTAs of the 192-thread variant fit the 227 KB of an SM
    assert bead_regtile_smem_bytes(40) == 8 * (720 + 4684 + 1200 + 960 + 72 + 14)
    assert bead_regtile_smem_bytes(5) == 8 * (96 + 64 + 160 + 128 + 72 + 3)
    assert 3 * (bead_regtile_smem_bytes(40, 192) + 1024) <= 227 * 1024


@pytest.mark.parametrize("nbw", [3, 4, 5, 6, 7])
@pytest.mark.parametrize("nb", [5, 6, 11, 40, 42])
def test_bulk_warp_tile_maps(nb, nbw):
    """The packed tile maps of the bulk warps (sde_bead_kernel: `map[q]`): every off-diagonal tile exactly once, dealt
    round-robin in column-major order, so that for each warp the column index never decreases along its slots - the
    property the chunk-wise skipping of finished tiles relies on - and unused slots are (0, 0)."""
    NT = -(-3 * nb // 8)
    noff = NT * (NT - 1) // 2
    opw = -(-noff // nbw)
    mapw = (-(-opw // 5) * 5 + 3) // 4
    seen = []
    for w in range(nbw):
        last_j = 0
        slots = []
        for k in range(mapw * 4):
            o = w + nbw * k
            v = 0
            if k < opw and o < noff:
                J, rem = 0, o
                while rem >= NT - 1 - J:
                    rem -= NT - 1 - J
                    J += 1
                v = (J + 1 + rem) | (J << 4)
            I, J = v & 15, v >> 4
            slots.append((I, J, bool(v)))
            if v:
                assert NT > I > J >= last_j
                last_j = J
                seen.append((I, J))
            else:
                assert (I, J) == (0, 0)
        # the chunk-activity test of the trailing update (`jlast`): for every panel p a chunk is entered iff it
        # holds an active tile
        for p in range(NT):
            for ch in range(-(-opw // 5)):
                last = min(ch * 5 + 4, opw - 1)
                jlast = slots[last][1] if (w + nbw * last < noff or last == 0) else slots[max(last - 1, 0)][1]
                truth = any(valid and J > p for (I, J, valid) in slots[ch * 5:ch * 5 + 5])
                assert (jlast > p) == truth, (nb, nbw, w, p, ch)
    assert sorted(seen) == sorted((I, J) for J in range(NT) for I in range(J + 1, NT))


def test_value_table_indices_and_lane_parallel_factorisation():
    """`mobility_index` (symmetric 3x3 pair blocks in a value table) and the lane-parallel Cholesky of the diagonal
    block on its accumulator fragment (`factor_diagonal`), modelled lane by lane with explicit shuffles."""
    nb, n3 = 7, 21
    rng = np.random.default_rng(1)
    blocks = rng.standard_normal((nb * (nb - 1) // 2, 6))
    val = np.concatenate([[0.0, 1.0, 0.25, 0.0], blocks.ravel()])

    def index(R, C):
        if R >= n3 or C >= n3:
            return 1 if R == C else 0
        bi, al, bj, be = R // 3, R % 3, C // 3, C % 3
        if bi == bj:
            return 2 if al == be else 0
        hi, lo = max(bi, bj), min(bi, bj)
        a0, a1 = min(al, be), max(al, be)
        return 4 + 6 * (hi * (hi - 1) // 2 + lo) + (a1 if a0 == 0 else a0 + a1 + 1)
    sym = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
    M = np.array([[val[index(R, C)] for C in range(24)] for R in range(24)])
    assert np.allclose(M, M.T) and np.allclose(M[21:, 21:], np.eye(3)) and np.all(M[21:, :21] == 0)
    for bi in range(nb):
        for bj in range(bi):
            blk = np.array([[blocks[bi * (bi - 1) // 2 + bj][sym[min(a, b), max(a, b)]] for b in range(3)] for a in range(3)])
            assert np.array_equal(M[3 * bi:3 * bi + 3, 3 * bj:3 * bj + 3], blk)
        assert np.array_equal(M[3 * bi:3 * bi + 3, 3 * bi:3 * bi + 3], 0.25 * np.eye(3))

    # lane-parallel factorisation
    B = rng.standard_normal((8, 8))
    S = B @ B.T + 8 * np.eye(8)
    c = np.array([[S[l >> 2, 2 * (l & 3)], S[l >> 2, 2 * (l & 3) + 1]] for l in range(32)])
    lanes = np.arange(32)
    i, jp = lanes >> 2, lanes & 3
    j0 = 2 * jp
    rinvs = np.zeros(8)
    for k in range(8):
        comp = k & 1
        d = c[4 * k + (k >> 1), comp]
        rinv = 1 / np.sqrt(d)
        rinvs[k] = rinv
        c[jp == (k >> 1), comp] *= rinv
        if k < 7:
            src = c[:, comp].copy()
            lik = src[4 * i + (k >> 1)]
            lj0 = src[np.minimum(4 * j0 + (k >> 1), 31)]
            lj1 = src[np.minimum(4 * (j0 + 1) + (k >> 1), 31)]
            m0, m1 = j0 > k, j0 + 1 > k
            c[m0, 0] -= (lik * lj0)[m0]
            c[m1, 1] -= (lik * lj1)[m1]
    Lm = np.zeros((8, 8))
    for l in range(32):
        Lm[i[l], j0[l]] = c[l, 0] if j0[l] <= i[l] else 0.0
        Lm[i[l], j0[l] + 1] = c[l, 1] if j0[l] + 1 <= i[l] else 0.0
    assert np.allclose(Lm, np.linalg.cholesky(S), rtol=1e-12, atol=1e-13)
    assert np.allclose(rinvs, 1 / np.diag(Lm))
