// sde_hooks.cuh - device-side reductions shared by the stepping kernels (sde_kernel.cuh, sde_kernel_ws.cuh).
//
// Replaces host post-processing that the reference's examples run over stored trajectories - ensemble means per
// snapshot such as the mean-square displacements of reference examples/jfm.py:134-171 - with sums accumulated
// while the trajectories are integrated, so that (n_traj, n_snapshots, d) never has to exist.
//
// Included after the emitted sde_observe(x, w, t, x_ref, obs).
#pragma once

#if SDE_NOBS > 0
namespace sdeb {

// Sum and sum of squares of every observable over the 32 trajectories of the calling warp, added to
// moments[k][0..1] with one atomicAdd pair per observable per warp.  Must be called by all lanes of the warp.
__device__ __forceinline__ void observe_accumulate_warp(const real* x, const real* w, const real t, const real* xr,
                                                        const bool active, double* moments) {
    double obs[SDE_NOBS];
    sde_observe(x, w, t, xr, obs);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < SDE_NOBS; ++k) {
        double s1 = active ? obs[k] : 0.0;
        double s2 = s1 * s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            atomicAdd(&moments[2 * k], s1);
            atomicAdd(&moments[2 * k + 1], s2);
        }
    }
}

}  // namespace sdeb
#endif
