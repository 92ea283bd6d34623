// Latency of the lane-parallel fraction-free 8x8 factorisation (csrc/sde_bead_kernel.cuh: factor_fraction_free) for one
// warp alone on an SM, and of variants that drop pieces of it - where do the cycles of a pivot go?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o factor_latency factor_latency.cu && ./factor_latency
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double shfl_real(const double v, const int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
__device__ __forceinline__ double pivot_rsqrt(const double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h, y * y, 1.5);
    y = y * fma(-h, y * y, 1.5);
    return y;
}

template <int VARIANT>
__device__ __forceinline__ void factor(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double bx = (j0 == i) ? 1.0 : 0.0, by = (j0 + 1 == i) ? 1.0 : 0.0;
    double Q = 1.0, px = 1.0, py = 1.0, pi = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double ck = (k & 1) ? cy : cx;
        const double ckk = shfl_real(ck, 4 * k + (k >> 1));
        const double cik = shfl_real(ck, 4 * i + (k >> 1));
        const int rowk = 4 * k + qd;
        const double cj0 = shfl_real(cx, rowk), cj1 = shfl_real(cy, rowk);
        double bk0 = 0.0, bk1 = 0.0;
        if (VARIANT != 1) { bk0 = shfl_real(bx, rowk); bk1 = shfl_real(by, rowk); }
        double alpha = 1.0;
        if (VARIANT != 3) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(alpha) : "d"(ckk));
        const double pq = Q * ckk;
        if (VARIANT != 2) {
            px = (j0 == k) ? pq : px;
            py = (j0 + 1 == k) ? pq : py;
            pi = (i == k) ? pq : pi;
        }
        const double nx = fma(ckk, cx, -(cik * cj0)) * alpha, ny = fma(ckk, cy, -(cik * cj1)) * alpha;
        const bool below = i > k;
        cx = (below && j0 > k) ? nx : cx;
        cy = (below && j0 + 1 > k) ? ny : cy;
        if (VARIANT != 1) {
            const double nbx = fma(ckk, bx, -(cik * bk0)) * alpha, nby = fma(ckk, by, -(cik * bk1)) * alpha;
            bx = below ? nbx : bx;
            by = below ? nby : by;
        }
        Q *= ckk * alpha;
    }
    double sx = 1.0, sy = 1.0, si = 1.0;
    if (VARIANT != 2) { sx = pivot_rsqrt(px); sy = pivot_rsqrt(py); si = pivot_rsqrt(pi); }
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = (j0 <= i) ? bx * si : 0.0;
    wy = (j0 + 1 <= i) ? by * si : 0.0;
}


// the version the kernel uses: power-of-two scale from the exponent bits (no MUFU, exact), scale folded into the two row
// factors, predicated FMAs instead of selects, one reciprocal square root per lane + shuffles instead of three
#define PRED_FMA(c, thr, K, g, nt) \
    asm("{\n .reg .pred p;\n setp.gt.s32 p, %1, " #K ";\n @p fma.rn.f64 %0, %2, %0, %3;\n}" : "+d"(c) : "r"(thr), "d"(g), "d"(nt))
template <int K>
__device__ __forceinline__ void pivot_step(double& cx, double& cy, double& bx, double& by, double& Q, double& mypq, const int lane,
                                           const int i, const int qd, const int mx, const int my) {
    const double ck = (K & 1) ? cy : cx;
    const double ckk = shfl_real(ck, 4 * K + (K >> 1));
    const double cik = shfl_real(ck, 4 * i + (K >> 1));
    const int rowk = 4 * K + qd;
    const double cj0 = shfl_real(cx, rowk), cj1 = shfl_real(cy, rowk);
    const double bk0 = shfl_real(bx, rowk), bk1 = shfl_real(by, rowk);
    // alpha = 2^-e for c_kk = m 2^e, 1 <= m < 2: exponent arithmetic, and every product with it is exact
    const double alpha = __hiloint2double(0x7fe00000 - (__double2hiint(ckk) & 0x7ff00000), 0);
    const double pq = Q * ckk;
    mypq = (lane == K) ? pq : mypq;
    const double g = ckk * alpha, nh = -(cik * alpha);
    const double tx = nh * cj0, ty = nh * cj1, tbx = nh * bk0, tby = nh * bk1;
    if (K == 0) { PRED_FMA(cx, mx, 0, g, tx); PRED_FMA(cy, my, 0, g, ty); PRED_FMA(bx, i, 0, g, tbx); PRED_FMA(by, i, 0, g, tby); }
    if (K == 1) { PRED_FMA(cx, mx, 1, g, tx); PRED_FMA(cy, my, 1, g, ty); PRED_FMA(bx, i, 1, g, tbx); PRED_FMA(by, i, 1, g, tby); }
    if (K == 2) { PRED_FMA(cx, mx, 2, g, tx); PRED_FMA(cy, my, 2, g, ty); PRED_FMA(bx, i, 2, g, tbx); PRED_FMA(by, i, 2, g, tby); }
    if (K == 3) { PRED_FMA(cx, mx, 3, g, tx); PRED_FMA(cy, my, 3, g, ty); PRED_FMA(bx, i, 3, g, tbx); PRED_FMA(by, i, 3, g, tby); }
    if (K == 4) { PRED_FMA(cx, mx, 4, g, tx); PRED_FMA(cy, my, 4, g, ty); PRED_FMA(bx, i, 4, g, tbx); PRED_FMA(by, i, 4, g, tby); }
    if (K == 5) { PRED_FMA(cx, mx, 5, g, tx); PRED_FMA(cy, my, 5, g, ty); PRED_FMA(bx, i, 5, g, tbx); PRED_FMA(by, i, 5, g, tby); }
    if (K == 6) { PRED_FMA(cx, mx, 6, g, tx); PRED_FMA(cy, my, 6, g, ty); PRED_FMA(bx, i, 6, g, tbx); PRED_FMA(by, i, 6, g, tby); }
    if (K == 7) { PRED_FMA(cx, mx, 7, g, tx); PRED_FMA(cy, my, 7, g, ty); PRED_FMA(bx, i, 7, g, tbx); PRED_FMA(by, i, 7, g, tby); }
    Q *= g;
}
template <>
__device__ __forceinline__ void factor<4>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    const int mx = i < j0 ? i : j0, my = i < j0 + 1 ? i : j0 + 1;     // entry (i, j) is updated by the pivots k < min(i, j)
    double bx = (j0 == i) ? 1.0 : 0.0, by = (j0 + 1 == i) ? 1.0 : 0.0;
    double Q = 1.0, mypq = 1.0;
    pivot_step<0>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<1>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<2>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<3>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<4>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<5>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<6>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    pivot_step<7>(cx, cy, bx, by, Q, mypq, lane, i, qd, mx, my);
    const double sg = pivot_rsqrt(mypq);                                  // lane k < 8: sigma_k
    const double sx = shfl_real(sg, j0), sy = shfl_real(sg, j0 + 1), si = shfl_real(sg, i);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = (j0 <= i) ? bx * si : 0.0;
    wy = (j0 + 1 <= i) ? by * si : 0.0;
}

// pivot row through shared memory instead of shuffles: the four lanes of row k publish their entries of both fragments,
// every lane reads c_kk, c_ik = c_ki (symmetry), its two columns of row k of C and of B: 2 STS + 4 LDS per pivot instead
// of 12 SHFL and the register moves that pair their 32-bit halves
template <>
__device__ __forceinline__ void factor<5>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    __shared__ __align__(16) double rowbuf[2][16];
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double bx = (j0 == i) ? 1.0 : 0.0, by = (j0 + 1 == i) ? 1.0 : 0.0;
    double Q = 1.0, px = 1.0, py = 1.0, pi = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double* rb = rowbuf[k & 1];
        if (i == k) {
            *reinterpret_cast<double2*>(rb + j0) = make_double2(cx, cy);
            *reinterpret_cast<double2*>(rb + 8 + j0) = make_double2(bx, by);
        }
        __syncwarp();
        const double ckk = rb[k], cik = rb[i];
        const double2 cj = *reinterpret_cast<const double2*>(rb + j0), bk = *reinterpret_cast<const double2*>(rb + 8 + j0);
        double alpha;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(alpha) : "d"(ckk));
        const double pq = Q * ckk;
        px = (j0 == k) ? pq : px;
        py = (j0 + 1 == k) ? pq : py;
        pi = (i == k) ? pq : pi;
        const double nx = fma(ckk, cx, -(cik * cj.x)) * alpha, ny = fma(ckk, cy, -(cik * cj.y)) * alpha;
        const double nbx = fma(ckk, bx, -(cik * bk.x)) * alpha, nby = fma(ckk, by, -(cik * bk.y)) * alpha;
        const bool below = i > k;
        cx = (below && j0 > k) ? nx : cx;
        cy = (below && j0 + 1 > k) ? ny : cy;
        bx = below ? nbx : bx;
        by = below ? nby : by;
        Q *= ckk * alpha;
    }
    const double sx = pivot_rsqrt(px), sy = pivot_rsqrt(py), si = pivot_rsqrt(pi);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = (j0 <= i) ? bx * si : 0.0;
    wy = (j0 + 1 <= i) ? by * si : 0.0;
}

// variant 5 + power-of-two scale from the exponent bits folded into the two row factors + one reciprocal square root per
// lane, exchanged through shared memory
template <>
__device__ __forceinline__ void factor<6>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    __shared__ __align__(16) double rowbuf[2][16];
    __shared__ __align__(16) double sig[8];
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double bx = (j0 == i) ? 1.0 : 0.0, by = (j0 + 1 == i) ? 1.0 : 0.0;
    double Q = 1.0, mypq = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double* rb = rowbuf[k & 1];
        if (i == k) {
            *reinterpret_cast<double2*>(rb + j0) = make_double2(cx, cy);
            *reinterpret_cast<double2*>(rb + 8 + j0) = make_double2(bx, by);
        }
        __syncwarp();
        const double ckk = rb[k], cik = rb[i];
        const double2 cj = *reinterpret_cast<const double2*>(rb + j0), bk = *reinterpret_cast<const double2*>(rb + 8 + j0);
        const double alpha = __hiloint2double(0x7fe00000 - (__double2hiint(ckk) & 0x7ff00000), 0);   // 2^-e: exact products
        mypq = (lane == k) ? Q * ckk : mypq;
        const double g = ckk * alpha, nh = -(cik * alpha);
        const double nx = fma(g, cx, nh * cj.x), ny = fma(g, cy, nh * cj.y);
        const double nbx = fma(g, bx, nh * bk.x), nby = fma(g, by, nh * bk.y);
        const bool below = i > k;
        cx = (below && j0 > k) ? nx : cx;
        cy = (below && j0 + 1 > k) ? ny : cy;
        bx = below ? nbx : bx;
        by = below ? nby : by;
        Q *= g;
    }
    if (lane < 8) sig[lane] = pivot_rsqrt(mypq);
    __syncwarp();
    const double2 sxy = *reinterpret_cast<const double2*>(sig + j0);
    const double si = sig[i];
    lx = (j0 <= i) ? cx * sxy.x : 0.0;
    ly = (j0 + 1 <= i) ? cy * sxy.y : 0.0;
    wx = (j0 <= i) ? bx * si : 0.0;
    wy = (j0 + 1 <= i) ? by * si : 0.0;
}


// ---- rank-one updates on the tensor cores -------------------------------------------------------------------------------
// Pivot k of the elimination is C -= (u / c_kk) u^T with u = column k of C below the pivot: ONE mma.m8n8k4 whose operands
// are non-zero in contraction slot 0 only.  Column k lives in the quad's lane k / 2, so one quad broadcast (2 SHFL) feeds
// both operands (C is symmetric); the inverse is tracked TRANSPOSED (T = B^T, T -= T[:, k] (u / c_kk)^T), so its operand is
// a quad broadcast as well.  1 / c_kk: rcp.approx + Newton.  Per pivot 6 SHFL + ~8 FP64 + 2 DMMA instead of 12 SHFL + 15 FP64
// + 12 SEL.  Outputs: lx, ly = L[i][j0], L[i][j0+1]; wx, wy = inv(L)^T[i][j0], inv(L)^T[i][j0+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NEWTON, bool EARLY_SIGMA>
__device__ __forceinline__ void factor_dmma(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double tx = (j0 == i) ? 1.0 : 0.0, ty = (j0 + 1 == i) ? 1.0 : 0.0;
    double piv = 1.0, sx = 1.0, sy = 1.0;
    const int qbase = lane & ~3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double ck = (k & 1) ? cy : cx, tk = (k & 1) ? ty : tx;
        const double ckk = shfl_real(ck, 4 * k + (k >> 1));
        const double u = shfl_real(ck, qbase + (k >> 1));           // c[i][k]
        const double v = shfl_real(tk, qbase + (k >> 1));           // T[i][k] = B[k][i]
        double alpha;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(alpha) : "d"(ckk));
#pragma unroll
        for (int n = 0; n < NEWTON; ++n) alpha = fma(alpha, fma(-ckk, alpha, 1.0), alpha);
        const bool on = qd == 0 && i > k;
        const double ub = on ? u : 0.0;
        const double m = ub * alpha;                                 // u_i / c_kk below the pivot, slot 0
        dmma884(cx, cy, -m, ub);
        dmma884(tx, ty, qd == 0 ? -v : 0.0, m);
        if (EARLY_SIGMA) {
            const double s = pivot_rsqrt(ckk);
            sx = (j0 == k) ? s : sx;
            sy = (j0 + 1 == k) ? s : sy;
        } else {
            piv = (lane == k) ? ckk : piv;
        }
    }
    if (!EARLY_SIGMA) {
        const double sg = pivot_rsqrt(piv);
        sx = shfl_real(sg, j0);
        sy = shfl_real(sg, j0 + 1);
    }
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = tx * sx;                                                    // inv(L)^T[i][j] = sigma_j B[j][i]
    wy = ty * sy;
}
template <> __device__ __forceinline__ void factor<7>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) { factor_dmma<2, false>(cx, cy, lane, lx, ly, wx, wy); }
template <> __device__ __forceinline__ void factor<8>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) { factor_dmma<2, true>(cx, cy, lane, lx, ly, wx, wy); }
template <> __device__ __forceinline__ void factor<9>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) { factor_dmma<1, true>(cx, cy, lane, lx, ly, wx, wy); }


// ---- the product first, the reciprocal beside it ------------------------------------------------------------------------
// P = u u^T needs no shuffle at all: column k of the symmetric tile, masked to the quad lane that holds it, IS both operands
// of the DMMA (contraction slot s <-> column 2 s + (k & 1)).  The DMMA runs while 1 / c_kk is refined (rcp.approx, then
// (1 + r)(1 + r^2) with r = 1 - c_kk a0: three dependent DFMA), and C -= a P closes the pivot.  Dependent chain per pivot:
// SHFL (c_kk) - MUFU - 3 DFMA - DFMA.
template <int CORR>
__device__ __forceinline__ void factor_pfirst(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double tx = (j0 == i) ? 1.0 : 0.0, ty = (j0 + 1 == i) ? 1.0 : 0.0;
    double piv = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double ck = (k & 1) ? cy : cx, tk = (k & 1) ? ty : tx;
        const bool mine = qd == (k >> 1);
        const double uc = (mine && i > k) ? ck : 0.0;               // c[i][k] below the pivot, in contraction slot k / 2
        const double vt = mine ? tk : 0.0;                           // T[i][k] = B[k][i]
        double pcx = 0.0, pcy = 0.0, ptx = 0.0, pty = 0.0;
        dmma884(pcx, pcy, uc, uc);                                   // c[i][k] c[j][k]
        dmma884(ptx, pty, vt, uc);                                   // T[j][k] c[i][k]
        const double ckk = shfl_real(ck, 4 * k + (k >> 1));
        double a;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(a) : "d"(ckk));
        const double r = fma(-ckk, a, 1.0);
        a = fma(a, r, a);
        if (CORR == 2) a = fma(a, r * r, a);
        cx = fma(-a, pcx, cx);
        cy = fma(-a, pcy, cy);
        tx = fma(-a, ptx, tx);
        ty = fma(-a, pty, ty);
        piv = (lane == k) ? ckk : piv;
    }
    const double sg = pivot_rsqrt(piv);
    const double sx = shfl_real(sg, j0), sy = shfl_real(sg, j0 + 1);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = tx * sx;
    wy = ty * sy;
}
template <> __device__ __forceinline__ void factor<10>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) { factor_pfirst<2>(cx, cy, lane, lx, ly, wx, wy); }
template <> __device__ __forceinline__ void factor<11>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) { factor_pfirst<1>(cx, cy, lane, lx, ly, wx, wy); }


// ... and the NEXT pivot predicted in every lane: d_{k+1} = S_k[k+1][k+1] - a_k S_k[k+1][k]^2 needs two broadcasts of S_k
// (issued when S_k is complete, long before a_k) and one DFMA after a_k - the reciprocal chain (DFMA, MUFU, 3 DFMA: ~49
// cycles) is all that is left between consecutive pivots; the tile update trails it.
__device__ __forceinline__ double recip_exact(const double d) {
    double a;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(a) : "d"(d));
    const double r = fma(-d, a, 1.0);
    a = fma(a, r, a);
    return fma(a, r * r, a);
}
template <>
__device__ __forceinline__ void factor<12>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double tx = (j0 == i) ? 1.0 : 0.0, ty = (j0 + 1 == i) ? 1.0 : 0.0;
    double piv = 1.0;
    double d = shfl_real(cx, 0);
    double a = recip_exact(d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double ck = (k & 1) ? cy : cx, tk = (k & 1) ? ty : tx;
        const bool mine = qd == (k >> 1);
        const double uc = (mine && i > k) ? ck : 0.0;
        const double vt = mine ? tk : 0.0;
        double pcx = 0.0, pcy = 0.0, ptx = 0.0, pty = 0.0;
        dmma884(pcx, pcy, uc, uc);
        dmma884(ptx, pty, vt, uc);
        double dn = 1.0, an = 1.0;
        if (k < 7) {
            const double e = shfl_real(((k + 1) & 1) ? cy : cx, 4 * (k + 1) + ((k + 1) >> 1));
            const double w = shfl_real(ck, 4 * (k + 1) + (k >> 1));
            dn = fma(-a, w * w, e);
            an = recip_exact(dn);
        }
        cx = fma(-a, pcx, cx);
        cy = fma(-a, pcy, cy);
        tx = fma(-a, ptx, tx);
        ty = fma(-a, pty, ty);
        piv = (lane == k) ? d : piv;
        d = dn;
        a = an;
    }
    const double sg = pivot_rsqrt(piv);
    const double sx = shfl_real(sg, j0), sy = shfl_real(sg, j0 + 1);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = tx * sx;
    wy = ty * sy;
}


// ... and the epilogue off the chain: every lane keeps the pivots of its own two columns (taken from the prediction, so the
// last reciprocal square root starts before the last tile update), no DMMA in the last pivot (nothing lies below it)
template <>
__device__ __forceinline__ void factor<13>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    double tx = (j0 == i) ? 1.0 : 0.0, ty = (j0 + 1 == i) ? 1.0 : 0.0;
    double d = shfl_real(cx, 0);
    double px = d, py = 1.0;                                          // pivots of columns j0, j0 + 1 (column 0: lanes with qd == 0)
    double a = recip_exact(d);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double ck = (k & 1) ? cy : cx, tk = (k & 1) ? ty : tx;
        const bool mine = qd == (k >> 1);
        const double uc = (mine && i > k) ? ck : 0.0;
        const double vt = mine ? tk : 0.0;
        double pcx = 0.0, pcy = 0.0, ptx = 0.0, pty = 0.0;
        dmma884(pcx, pcy, uc, uc);
        dmma884(ptx, pty, vt, uc);
        const double e = shfl_real(((k + 1) & 1) ? cy : cx, 4 * (k + 1) + ((k + 1) >> 1));
        const double w = shfl_real(ck, 4 * (k + 1) + (k >> 1));
        const double dn = fma(-a, w * w, e);
        if ((k + 1) & 1) py = (j0 == k) ? dn : py; else px = (j0 == k + 1) ? dn : px;
        const double an = (k < 6) ? recip_exact(dn) : 0.0;
        cx = fma(-a, pcx, cx);
        cy = fma(-a, pcy, cy);
        tx = fma(-a, ptx, tx);
        ty = fma(-a, pty, ty);
        a = an;
    }
    const double sx = pivot_rsqrt(px), sy = pivot_rsqrt(py);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = tx * sx;
    wy = ty * sy;
}


// ... and ONE tile, ONE DMMA per pivot: the lower triangle (diagonal included) holds S, the strict upper triangle holds T =
// B^T, whose diagonal is 1 and whose column k is non-zero only in the rows above k - exactly the rows S has finished with.
// Column k of that tile with its diagonal entry replaced by 1 is the A operand of both updates at once.
template <>
__device__ __forceinline__ void factor<14>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    cx = (j0 > i) ? 0.0 : cx;                                        // T starts as the identity: nothing above its diagonal
    cy = (j0 + 1 > i) ? 0.0 : cy;
    double d = shfl_real(cx, 0);
    double px = d, py = 1.0;
    double a = recip_exact(d);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double ck = (k & 1) ? cy : cx;
        const bool mine = qd == (k >> 1);
        const double ub = (mine && i > k) ? ck : 0.0;                // S_k[i][k] below the pivot
        const double ua = mine ? ((i == k) ? 1.0 : ck) : 0.0;        // ... and T[i][k] above it, 1 on it
        double px2 = 0.0, py2 = 0.0;
        dmma884(px2, py2, ua, ub);
        const double e = shfl_real(((k + 1) & 1) ? cy : cx, 4 * (k + 1) + ((k + 1) >> 1));
        const double w = shfl_real(ck, 4 * (k + 1) + (k >> 1));
        const double dn = fma(-a, w * w, e);
        if ((k + 1) & 1) py = (j0 == k) ? dn : py; else px = (j0 == k + 1) ? dn : px;
        const double an = (k < 6) ? recip_exact(dn) : 0.0;
        // rows below the pivot: S, columns up to the diagonal; rows up to the pivot: T, every column (the product is zero left of k)
        const double nx = fma(-a, px2, cx), ny = fma(-a, py2, cy);
        cx = (i > k && j0 > i) ? cx : nx;
        cy = (i > k && j0 + 1 > i) ? cy : ny;
        a = an;
    }
    const double sx = pivot_rsqrt(px), sy = pivot_rsqrt(py);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = (j0 > i) ? cx * sx : (j0 == i) ? sx : 0.0;
    wy = (j0 + 1 > i) ? cy * sy : (j0 + 1 == i) ? sy : 0.0;
}


// ... the same single tile and predicted pivots, the product from shuffles instead of a DMMA (inside the kernel the FP64
// tensor pipe is saturated by the bulk warps' trailing update while the chain warp factorises)
template <>
__device__ __forceinline__ void factor<15>(double cx, double cy, const int lane, double& lx, double& ly, double& wx, double& wy) {
    const int i = lane >> 2, qd = lane & 3, j0 = 2 * qd;
    cx = (j0 > i) ? 0.0 : cx;
    cy = (j0 + 1 > i) ? 0.0 : cy;
    double d = shfl_real(cx, 0);
    double px = d, py = 1.0;
    double a = recip_exact(d);
    const int qbase = lane & ~3;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double ck = (k & 1) ? cy : cx;
        const double ar = shfl_real(ck, qbase + (k >> 1));           // column k at this lane's row: S below, T above the pivot
        const double b0 = shfl_real(ck, 4 * j0 + (k >> 1));          // S_k[j0][k], S_k[j0 + 1][k]
        const double b1 = shfl_real(ck, 4 * (j0 + 1) + (k >> 1));
        const double e = shfl_real(((k + 1) & 1) ? cy : cx, 4 * (k + 1) + ((k + 1) >> 1));
        const double w = shfl_real(ck, 4 * (k + 1) + (k >> 1));
        const double dn = fma(-a, w * w, e);
        if ((k + 1) & 1) py = (j0 == k) ? dn : py; else px = (j0 == k + 1) ? dn : px;
        const double an = (k < 6) ? recip_exact(dn) : 0.0;
        const double m = -a * ((i == k) ? 1.0 : ar);
        const double nx = fma(m, b0, cx), ny = fma(m, b1, cy);
        // columns right of the pivot only; rows below the pivot: S, up to the diagonal; rows up to the pivot: T
        cx = (j0 > k && !(i > k && j0 > i)) ? nx : cx;
        cy = (j0 + 1 > k && !(i > k && j0 + 1 > i)) ? ny : cy;
        a = an;
    }
    const double sx = pivot_rsqrt(px), sy = pivot_rsqrt(py);
    lx = (j0 <= i) ? cx * sx : 0.0;
    ly = (j0 + 1 <= i) ? cy * sy : 0.0;
    wx = (j0 > i) ? cx * sx : (j0 == i) ? sx : 0.0;
    wy = (j0 + 1 > i) ? cy * sy : (j0 + 1 == i) ? sy : 0.0;
}

// single-warp issue rates and latencies the chain depends on
__global__ void unit_rates(long long* cyc, double* out) {
    const int lane = threadIdx.x;
    double a0 = 1.0 + lane, a1 = 2.0 + lane, a2 = 3.0 + lane, a3 = 4.0 + lane, a4 = 5.0, a5 = 6.0, a6 = 7.0, a7 = 8.0;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;                                 // 2048 independent-ish DFMA (8 chains)
    double x = 1.5 + lane;
    t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        double y;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        x = y;
    }
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;                                 // 256 dependent rcp.approx
    double p0 = 0.0, p1 = 0.0;
    t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        dmma884(p0, p1, a0, a1);
        const double s = shfl_real(p0, (lane + 4) & 31);
        a0 = s * 1e-3;
    }
    t1 = clock64();
    if (lane == 0) cyc[2] = t1 - t0;                                 // 256 x (DMMA -> SHFL -> DMUL)
    out[lane] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + x + p0 + p1;
}

// one factorisation, fragments written out: L and W (W = inv(L) for variants < 7, inv(L)^T from 7 on)
template <int VARIANT>
__global__ void dump(double* out) {
    const int lane = threadIdx.x & 31, i = lane >> 2, j0 = 2 * (lane & 3);
    double cx = (i == j0) ? 2.0 : 0.1 / (1 + abs(i - j0)), cy = (i == j0 + 1) ? 2.0 : 0.1 / (1 + abs(i - j0 - 1));
    double lx, ly, wx, wy;
    factor<VARIANT>(cx, cy, lane, lx, ly, wx, wy);
    out[8 * i + j0] = lx;
    out[8 * i + j0 + 1] = ly;
    out[64 + 8 * i + j0] = wx;
    out[64 + 8 * i + j0 + 1] = wy;
}
template <int VARIANT>
void verify(const char* name, double* out, const bool transposed) {
    dump<VARIANT><<<1, 32>>>(out);
    double o[128];
    cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    double e1 = 0.0, e2 = 0.0;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double llt = 0.0, lw = 0.0;
            for (int k = 0; k < 8; ++k) {
                llt += o[8 * i + k] * o[8 * j + k];
                lw += o[8 * i + k] * (transposed ? o[64 + 8 * j + k] : o[64 + 8 * k + j]);
            }
            const double c = (i == j) ? 2.0 : 0.1 / (1 + abs(i - j));
            e1 = fmax(e1, fabs(llt - c));
            e2 = fmax(e2, fabs(lw - (i == j ? 1.0 : 0.0)));
        }
    printf("%-52s |L L^T - C| %.2e  |L W - 1| %.2e (%s)\n", name, e1, e2, cudaGetErrorString(cudaGetLastError()));
}

template <int VARIANT>
__global__ void k(double* out, long long* cyc, int iters) {
    const int lane = threadIdx.x & 31, i = lane >> 2, j0 = 2 * (lane & 3);
    // a well-conditioned SPD tile: 2 on the diagonal, 0.1 / (1 + |i - j|) off it
    double cx = (i == j0) ? 2.0 : 0.1 / (1 + abs(i - j0)), cy = (i == j0 + 1) ? 2.0 : 0.1 / (1 + abs(i - j0 - 1));
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double lx, ly, wx, wy;
        factor<VARIANT>(cx + 1e-9 * acc, cy, lane, lx, ly, wx, wy);       // the next call depends on this one's result
        acc += lx + ly + wx + wy;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / iters;
    out[threadIdx.x] = acc;
}

template <int VARIANT>
void run(const char* name, double* out, long long* cyc) {
    k<VARIANT><<<1, 32>>>(out, cyc, 200);
    k<VARIANT><<<1, 32>>>(out, cyc, 200);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double o[32];
    cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    printf("%-52s %6lld cycles per factorisation (%s, checksum %.6f)\n", name, h, cudaGetErrorString(cudaGetLastError()), o[9] + o[18]);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 128 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("full: L and inv(L)", out, cyc);
    run<1>("without the inverse (no second fragment)", out, cyc);
    run<2>("without the reciprocal square roots", out, cyc);
    run<3>("without the reciprocal approximation (alpha = 1)", out, cyc);
    run<4>("predicated FMAs, power-of-two scale", out, cyc);
    run<5>("pivot row through shared memory instead of shuffles", out, cyc);
    run<6>("... + power-of-two scale folded, one rsqrt per lane", out, cyc);
    run<7>("rank-one updates as DMMA, 2 Newton steps", out, cyc);
    run<8>("... + reciprocal square roots off the chain", out, cyc);
    run<9>("... with 1 Newton step", out, cyc);
    run<10>("product first (no u shuffle), exact reciprocal", out, cyc);
    run<11>("... with one correction step", out, cyc);
    run<12>("... and the next pivot predicted in every lane", out, cyc);
    verify<12>("... and the next pivot predicted in every lane", out, true);
    run<13>("... and the epilogue off the chain", out, cyc);
    verify<13>("... and the epilogue off the chain", out, true);
    run<14>("... one tile, one DMMA per pivot", out, cyc);
    verify<14>("... one tile, one DMMA per pivot", out, true);
    run<15>("... one tile, product from shuffles (no DMMA)", out, cyc);
    verify<15>("... one tile, product from shuffles (no DMMA)", out, true);
    verify<10>("product first (no u shuffle), exact reciprocal", out, true);
    verify<11>("... with one correction step", out, true);
    {
        long long* c3;
        cudaMalloc(&c3, 3 * sizeof(long long));
        unit_rates<<<1, 32>>>(c3, out);
        unit_rates<<<1, 32>>>(c3, out);
        long long h[3];
        cudaMemcpy(h, c3, sizeof(h), cudaMemcpyDeviceToHost);
        printf("one warp: %.2f cycles per independent DFMA (8 chains), %.1f per dependent rcp.approx.f64, %.1f per DMMA -> SHFL -> DMUL\n",
               h[0] / 2048.0, h[1] / 256.0, h[2] / 256.0);
    }
    verify<0>("full: L and inv(L)", out, false);
    verify<5>("pivot row through shared memory", out, false);
    verify<7>("rank-one updates as DMMA, 2 Newton steps", out, true);
    verify<8>("... + reciprocal square roots off the chain", out, true);
    verify<9>("... with 1 Newton step", out, true);
    return 0;
}
