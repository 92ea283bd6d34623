// sde_wiener_kernel.cuh - unit-step Wiener-integral generator as a standalone kernel: one thread
// per step s of a single key.  Device counterpart of the reference's
// get_wiener_integrals(key, steps, noise_terms, scheme, p)
// (pychastic/vectorized_I_generation.py:243-473); it runs the SAME WienerGen code the stepping
// kernel uses, so checking it checks the in-kernel generator.  Include after sde_core.cuh.
#pragma once

namespace sdeb {
struct WienerArgs {
    u32 key_a, key_b;
    u32 steps;
    u32 _pad;
    real* d_w;    // (steps, M)
    real* d_ww;   // (steps, M, M)
    real* d_wt;   // (steps, M)
    real* d_tw;   // (steps, M)
    real* d_www;  // (steps, M, M, M)
};
}  // namespace sdeb

extern __shared__ __align__(16) unsigned char sde_dyn_smem[];

extern "C" __global__ void __launch_bounds__(SDE_BLOCK) sde_wiener_kernel(const sdeb::WienerArgs A) {
    using namespace sdeb;
    real* const stage = reinterpret_cast<real*>(sde_dyn_smem) + threadIdx.x;
    const u32 s = blockIdx.x * SDE_BLOCK + threadIdx.x;
    if (s >= A.steps) return;
    Key key;
    key.a = A.key_a;
    key.b = A.key_b;
    WienerGen gen;
    gen.init(key, A.steps);
    StepScale sc;
    sc.dt = RC(1.0);
    sc.sqrt_dt = RC(1.0);
    sc.dt32 = RC(1.0);
    Integrals Iblk[SDE_STEP_BLOCK];      // the block generator is the production path; only its first step is kept
    gen.generate_block(s, sc, Iblk, stage);
    const Integrals& I = Iblk[0];
    constexpr int M = SDE_M;
#pragma unroll
    for (int j = 0; j < M; ++j) A.d_w[(u64)s * M + j] = I.dW[j];
#if SDE_SCHEME >= SDE_MILSTEIN
    if (A.d_ww) {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) A.d_ww[((u64)s * M + i) * M + j] = I.Iww[i][j];
    }
#endif
#if SDE_SCHEME >= SDE_WAGNER_PLATEN
#pragma unroll
    for (int j = 0; j < M; ++j) {
        if (A.d_wt) A.d_wt[(u64)s * M + j] = I.Iwt[j];
        if (A.d_tw) A.d_tw[(u64)s * M + j] = I.Itw[j];
    }
    if (A.d_www) {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int k = 0; k < M; ++k) A.d_www[(((u64)s * M + i) * M + j) * M + k] = I.Iwww[i][j][k];
    }
#endif
}
