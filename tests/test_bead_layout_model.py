"""Host model of the shared-memory layout and the tensor-core tiling of the bead kernel's Cholesky trailing update
(csrc/sde_bead_kernel.cuh: `pbase`, `poff`, `lidx`, the 16x16 super-tile loop under SDE_BEAD_DMMA), CPU only.

The CUDA code was written against this model: it mirrors the index functions and the lane -> (row, column) maps of the
`mma.sync.m8n8k4.f64` fragments and checks that
* every stored element of the panel-packed lower triangle has its own address inside SDE_LSIZE,
* the super-tile decomposition (with its row permutation, masking of rows past the end and of the block above the
  diagonal) reproduces the dense right-looking update  A[i][j] -= sum_k L[i][c0+k] L[j][c0+k]  on the stored triangle,
* every fragment / accumulator access of a warp takes the minimum number of shared-memory wavefronts (no bank conflicts).
"""
import numpy as np
import pytest

NP = 120                                     # 3 * 40 beads


def pbase(p):
    return 17 * (p * NP - 4 * p * (p - 1)) // 2 + 2 * (p & 1)


def poff(r):
    return 8 * r + 2 * (r >> 2)


def lidx(i, j):
    p = j >> 3
    return pbase(p) + poff(i - 8 * p) + (j & 7)


LSIZE = 17 * ((NP // 8) * NP - 4 * (NP // 8) * ((NP // 8) - 1)) // 2 + 2


def row_of(tile, fr):                        # matrix row (column) of MMA-tile row (column) fr inside a 16-block
    return 2 * tile + (fr & 1) + 8 * ((fr >> 1) & 1) + 4 * (fr >> 2)


def wavefronts(word_addrs, width_words, lanes_per_wavefront):
    total = 0
    for g in range(0, 32, lanes_per_wavefront):
        rows_per_bank = {}
        for a in word_addrs[g:g + lanes_per_wavefront]:
            if a is None:
                continue
            for w in range(width_words):
                rows_per_bank.setdefault((a + w) % 32, set()).add((a + w) // 32)
        total += max((len(v) for v in rows_per_bank.values()), default=0)
    return total


def test_every_stored_element_has_its_own_address():
    seen = {}
    for j in range(NP):
        for i in range(8 * (j >> 3), NP):
            a = lidx(i, j)
            assert 0 <= a < LSIZE and a not in seen, (i, j, seen.get(a))
            seen[a] = (i, j)
            assert a % 2 == (j & 1)          # even columns are 16-byte aligned (double2 accumulator accesses)


@pytest.mark.parametrize("p", [0, 1, 6, 13])
def test_super_tiles_reproduce_the_dense_update_without_bank_conflicts(p):
    rng = np.random.default_rng(p)
    A = np.tril(rng.normal(size=(NP, NP)))
    L = np.full(LSIZE, np.nan)
    for j in range(NP):
        for i in range(8 * (j >> 3), NP):
            L[lidx(i, j)] = A[i, j]
    c0, t0 = 8 * p, 8 * p + 8
    T16 = (NP - t0 + 15) // 16
    for bi in range(T16):
        for bj in range(bi + 1):
            i0, j0 = t0 + 16 * bi, t0 + 16 * bj
            for t in (0, 1):
                for u in (0, 1):
                    a_addr = [[None] * 32 for _ in range(2)]
                    b_addr = [[None] * 32 for _ in range(2)]
                    c_addr = [None] * 32
                    for lane in range(32):
                        fr, fk = lane >> 2, lane & 3
                        row, coln = i0 + row_of(t, fr), j0 + row_of(u, fr)
                        for h in (0, 1):     # operand fragments: 64-bit loads, one k index per lane
                            if row < NP:
                                a_addr[h][lane] = 2 * (pbase(p) + poff(row - c0) + 4 * h + fk)
                            if coln < NP:
                                b_addr[h][lane] = 2 * (pbase(p) + poff(coln - c0) + 4 * h + fk)
                        col = j0 + row_of(u, 2 * fk)          # accumulator: columns col, col + 1 (one double2)
                        assert row_of(u, 2 * fk + 1) == row_of(u, 2 * fk) + 1
                        if row < NP and col < NP and row >= 8 * (col >> 3):
                            c_addr[lane] = 2 * lidx(row, col)
                            for e in (0, 1):
                                L[lidx(row, col + e)] -= float(A[row, c0:c0 + 8] @ A[col + e, c0:c0 + 8])
                    for h in (0, 1):
                        assert wavefronts(a_addr[h], 2, 16) <= 2 and wavefronts(b_addr[h], 2, 16) <= 2
                    assert wavefronts(c_addr, 4, 8) <= 4
    want = A - A[:, c0:c0 + 8] @ A[:, c0:c0 + 8].T
    for i in range(t0, NP):
        for j in range(t0, i + 1):
            assert abs(L[lidx(i, j)] - want[i, j]) < 1e-12, (i, j)
