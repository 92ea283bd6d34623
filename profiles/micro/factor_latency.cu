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
    cudaMalloc(&out, 64 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("full: L and inv(L)", out, cyc);
    run<1>("without the inverse (no second fragment)", out, cyc);
    run<2>("without the reciprocal square roots", out, cyc);
    run<3>("without the reciprocal approximation (alpha = 1)", out, cyc);
    run<4>("predicated FMAs, power-of-two scale", out, cyc);
    run<5>("pivot row through shared memory instead of shuffles", out, cyc);
    run<6>("... + power-of-two scale folded, one rsqrt per lane", out, cyc);
    return 0;
}
