// sde_kernel.cuh - the trajectory-stepping kernel (one thread per trajectory, state and step
// loop in registers, only requested snapshots written out).
//
// Replaces the vmapped 3-level lax.scan nest of the reference
// (pychastic/sde_solver.py:228-300: scan_func / chunk_function / get_solution_fragment /
// get_solution + vmap) for low-dimensional problems.
//
// Included AFTER sde_core.cuh and after the problem's device functions:
//   __device__ void sde_step(real (&x)[SDE_D], const sdeb::Integrals& I, real dt);   // required
//   __device__ void sde_post(real (&x)[SDE_D]);                      // if SDE_HAS_POST
//   __device__ void sde_observe(const real* x, const real* w, real t, const real* x_ref, double (&obs)[SDE_NOBS]);  // if SDE_NOBS > 0
//   __device__ real sde_event(const real* x, const real* w, real t);                      // if SDE_HAS_EVENT
#pragma once

#ifndef SDE_NOBS
#define SDE_NOBS 0
#endif
#ifndef SDE_OBS_SNAP
#define SDE_OBS_SNAP 0    // 1: observables are accumulated at EVERY snapshot into moments[snapshot][k][2]
#endif
#ifndef SDE_OBS_REF
#define SDE_OBS_REF 0     // 1: sde_observe also reads x_ref, the state at the first snapshot
#endif
#ifndef SDE_HAS_POST
#define SDE_HAS_POST 0
#endif
#ifndef SDE_HAS_EVENT
#define SDE_HAS_EVENT 0   // 1: sde_event(x, w, t) is watched for its first crossing of zero
#endif
#ifndef SDE_MIN_BLOCKS
#define SDE_MIN_BLOCKS 1
#endif
#ifndef SDE_SNAP_TILE
#define SDE_SNAP_TILE 1     // > 1: stage this many consecutive snapshots per trajectory in shared memory
#endif
#ifndef SDE_SNAP_TMA
#define SDE_SNAP_TMA 0      // > 0: snapshot tiles of this many snapshots leave through the TMA unit (sde_snap_tma.cuh)
#endif
#define SDE_SNAP_FIELDS (1 + SDE_D + SDE_M)
#define SDE_SNAP_PITCH 33   // lane pitch of a staged row (bank-conflict padding)
#ifndef SDE_LOCKSTEP
#define SDE_LOCKSTEP (SDE_STAGE_COUNT > 0)
#endif

#include "sde_hooks.cuh"
#include "sde_snap_tma.cuh"

namespace sdeb {

// Plain-old-data argument block; mirrored field by field by `sdeb200_solve_args_t`
// in include/sdeb200.h (the host copies it into the kernel parameter space).
struct SolveArgs {
    u32 key_a, key_b;         // jax.random.PRNGKey(seed)
    u32 n_fragments;          // number_of_chunks // chunks_per_randomization
    u32 cpr;                  // chunks per randomization (chunks per fragment)
    u32 chunk;                // steps per chunk
    u32 n_snapshots;          // chunks actually integrated and stored per trajectory (<= n_fragments*cpr)
    u64 n_traj_total;         // N of split(PRNGKey(seed), N): decides every trajectory key
    u64 traj_begin;           // global index of this launch's first trajectory
    u64 traj_count;           // trajectories in this launch
    const real* x0;           // initial conditions of this launch's trajectories
    long long x0_stride;      // elements between consecutive ICs (0: one IC broadcast)
    double dt;                // step
    double dt32;              // dt ** 1.5 (evaluated in double on the host, as the reference does)
    real* out_t;              // (traj_count, n_snapshots)            or nullptr
    real* out_x;              // (traj_count, n_snapshots, SDE_D)     or nullptr
    real* out_w;              // (traj_count, n_snapshots, SDE_M)     or nullptr
    double* moments;          // (SDE_NOBS, 2): sum, sum of squares; (n_snapshots, SDE_NOBS, 2) if SDE_OBS_SNAP   or nullptr
    real* out_hit;            // (traj_count, 2 + SDE_D): flag, time, state of the first passage   or nullptr
};

}  // namespace sdeb

// dynamic shared memory: SDE_STAGE_COUNT rows of SDE_BLOCK reals (the normals staging area)
extern __shared__ __align__(16) unsigned char sde_dyn_smem[];

#if SDE_SNAP_TMA
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_MIN_BLOCKS) sde_solve_kernel(const sdeb::SolveArgs A,
                                                                                        const __grid_constant__ sdeb::SnapMaps TM) {
#else
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, SDE_MIN_BLOCKS) sde_solve_kernel(const sdeb::SolveArgs A) {
#endif
    using namespace sdeb;
#if SDE_LOGTAB_IN_SMEM
    sde_fill_logtab();
#endif
    real* const stage = reinterpret_cast<real*>(sde_dyn_smem) + threadIdx.x;
#if SDE_SNAP_TMA
    SnapWriter snapw;
    snapw.init(sde_dyn_smem + (size_t)SDE_STAGE_COUNT * SDE_BLOCK * sizeof(real), threadIdx.x >> 5, threadIdx.x & 31,
               (u64)blockIdx.x * SDE_BLOCK + (threadIdx.x & ~31u));
#elif SDE_SNAP_TILE > 1
    // Snapshot staging: the thread-per-trajectory mapping would store 8 bytes at a stride of a whole trajectory
    // row.  Instead every warp collects SDE_SNAP_TILE consecutive snapshots of its 32 trajectories in shared
    // memory and writes them out transposed, so that one store instruction covers contiguous
    // SDE_SNAP_TILE * sizeof(real) runs of the (trajectory, snapshot, component) output rows.
    real* const snapbuf = reinterpret_cast<real*>(sde_dyn_smem) + SDE_STAGE_COUNT * SDE_BLOCK +
                          (threadIdx.x >> 5) * (SDE_SNAP_FIELDS * SDE_SNAP_TILE * SDE_SNAP_PITCH);
    const int lane = threadIdx.x & 31;
    const u64 warp_first = (u64)blockIdx.x * SDE_BLOCK + (threadIdx.x & ~31u);
    u32 tile_fill = 0, tile_snap0 = 0;
#endif
    const u64 tid_raw = (u64)blockIdx.x * SDE_BLOCK + threadIdx.x;
    const bool active = tid_raw < A.traj_count;
    // Threads past the end integrate trajectory 0 again (no stores): every thread of the CTA then runs the
    // same number of steps, which the per-step barrier below relies on.
    const u64 tid = active ? tid_raw : 0;
#if SDE_NOBS > 0 && !SDE_OBS_SNAP
    double obs[SDE_NOBS];
#pragma unroll
    for (int k = 0; k < SDE_NOBS; ++k) obs[k] = 0.0;
#endif
    {
        Key root;
        root.a = A.key_a;
        root.b = A.key_b;
        const Key tkey = split_key(root, A.n_traj_total, A.traj_begin + tid);

        real x[SDE_D], w[SDE_M];
        real t = RC(0.0);
#pragma unroll
        for (int i = 0; i < SDE_D; ++i) x[i] = A.x0[tid * (u64)A.x0_stride + i];
#pragma unroll
        for (int j = 0; j < SDE_M; ++j) w[j] = RC(0.0);
#if SDE_OBS_REF
        real xref[SDE_D];     // state at the first snapshot (what the reference's MSD post-processing subtracts)
#pragma unroll
        for (int i = 0; i < SDE_D; ++i) xref[i] = x[i];
#else
        const real* const xref = x;
#endif

        StepScale sc;
        sc.dt = (real)A.dt;
        sc.sqrt_dt = sqrt((real)A.dt);
        sc.dt32 = (real)A.dt32;
#if SDE_HAS_EVENT
        // first-passage observer: the first step after which sde_event(x, w, t) > 0, located by linear interpolation
        // between the states before and after that step (reference examples/hit.py:40-58, done on the device)
        real ev_prev = sde_event(x, w, t);
        real hit_t = RC(0.0), hit_x[SDE_D];
        bool hit = false;
#pragma unroll
        for (int i = 0; i < SDE_D; ++i) hit_x[i] = RC(0.0);
#endif
        const u32 S = A.chunk * A.cpr;
        u32 snap = 0;
        for (u32 f = 0; f < A.n_fragments && snap < A.n_snapshots; ++f) {
            WienerGen gen;
            gen.init(split_key(tkey, A.n_fragments, f), S);
            u32 in_chunk = 0;   // steps taken since the last snapshot
            for (u32 s0 = 0; s0 < S && snap < A.n_snapshots; s0 += SDE_STEP_BLOCK) {
#if SDE_LOCKSTEP
                // Keep the warps of the CTA on the same stretch of the (large, mostly straight-line) step body so
                // that one instruction-cache fill serves all of them.
                __syncthreads();
#endif
                Integrals Iblk[SDE_STEP_BLOCK];
                gen.generate_block(s0, sc, Iblk, stage);
#pragma unroll
                for (int b = 0; b < SDE_STEP_BLOCK; ++b) {
                    if (s0 + b >= S || snap >= A.n_snapshots) break;
                    const Integrals& I = Iblk[b];
#if SDE_HAS_EVENT
                    real x_before[SDE_D];
#pragma unroll
                    for (int i = 0; i < SDE_D; ++i) x_before[i] = x[i];
#endif
                    t += sc.dt;
                    sde_step(x, I, sc.dt);
#if SDE_HAS_POST
                    sde_post(x);
#endif
#pragma unroll
                    for (int j = 0; j < SDE_M; ++j) w[j] += I.dW[j];
#if SDE_HAS_EVENT
                    {
                        const real ev_now = sde_event(x, w, t);
                        if (!hit && ev_now > RC(0.0)) {
                            const real wgt = (RC(0.0) - ev_prev) / (ev_now - ev_prev);
                            hit = true;
                            hit_t = (t - sc.dt) + wgt * sc.dt;
#pragma unroll
                            for (int i = 0; i < SDE_D; ++i) hit_x[i] = x[i] * wgt + x_before[i] * (RC(1.0) - wgt);
                        }
                        ev_prev = ev_now;
                    }
#endif
                    if (++in_chunk < A.chunk) continue;
                    in_chunk = 0;
                    // ---- snapshot `snap` -----------------------------------------------------------------
#if SDE_OBS_REF
                    if (snap == 0) {
#pragma unroll
                        for (int i = 0; i < SDE_D; ++i) xref[i] = x[i];
                    }
#endif
#if SDE_NOBS > 0 && SDE_OBS_SNAP
                    if (A.moments) observe_accumulate_warp(x, w, t, xref, active, A.moments + (u64)snap * (2 * SDE_NOBS));
#endif
#if SDE_SNAP_TMA
                    snapw.snapshot(A, TM, snap, t, x, w);
#elif SDE_SNAP_TILE > 1
                {
                    real* col = snapbuf + tile_fill * SDE_SNAP_PITCH + lane;
                    col[0] = t;
#pragma unroll
                    for (int i = 0; i < SDE_D; ++i) col[(1 + i) * SDE_SNAP_TILE * SDE_SNAP_PITCH] = x[i];
#pragma unroll
                    for (int j = 0; j < SDE_M; ++j) col[(1 + SDE_D + j) * SDE_SNAP_TILE * SDE_SNAP_PITCH] = w[j];
                    if (tile_fill == 0) tile_snap0 = snap;
                    ++tile_fill;
                    if (tile_fill == SDE_SNAP_TILE || snap + 1 == A.n_snapshots) {
                        __syncwarp();
                        // rows of this warp's trajectories start at row0 + l * n_snapshots (l = lane of the owner)
                        const u64 row0 = warp_first * (u64)A.n_snapshots + tile_snap0;
                        const u32 lmax = (A.traj_count > warp_first)
                                             ? (u32)((A.traj_count - warp_first < 32) ? (A.traj_count - warp_first) : 32)
                                             : 0u;
                        if (A.out_t) {   // SDE_SNAP_TILE contiguous values per trajectory
                            real* const dst = A.out_t + row0;
#pragma unroll 4
                            for (u32 idx = lane; idx < 32u * SDE_SNAP_TILE; idx += 32) {
                                const u32 l = idx / SDE_SNAP_TILE, kk = idx % SDE_SNAP_TILE;
                                if (l < lmax && kk < tile_fill) dst[l * A.n_snapshots + kk] = snapbuf[kk * SDE_SNAP_PITCH + l];
                            }
                        }
                        if (A.out_x) {   // SDE_SNAP_TILE * SDE_D contiguous values per trajectory
                            real* const dst = A.out_x + row0 * SDE_D;
#pragma unroll 4
                            for (u32 idx = lane; idx < 32u * SDE_SNAP_TILE * SDE_D; idx += 32) {
                                const u32 l = idx / (SDE_SNAP_TILE * SDE_D), rem = idx % (SDE_SNAP_TILE * SDE_D);
                                const u32 kk = rem / SDE_D, i = rem % SDE_D;
                                if (l < lmax && kk < tile_fill)
                                    dst[l * A.n_snapshots * SDE_D + rem] =
                                        snapbuf[((1 + i) * SDE_SNAP_TILE + kk) * SDE_SNAP_PITCH + l];
                            }
                        }
                        if (A.out_w) {
                            real* const dst = A.out_w + row0 * SDE_M;
#pragma unroll 4
                            for (u32 idx = lane; idx < 32u * SDE_SNAP_TILE * SDE_M; idx += 32) {
                                const u32 l = idx / (SDE_SNAP_TILE * SDE_M), rem = idx % (SDE_SNAP_TILE * SDE_M);
                                const u32 kk = rem / SDE_M, j = rem % SDE_M;
                                if (l < lmax && kk < tile_fill)
                                    dst[l * A.n_snapshots * SDE_M + rem] =
                                        snapbuf[((1 + SDE_D + j) * SDE_SNAP_TILE + kk) * SDE_SNAP_PITCH + l];
                            }
                        }
                        __syncwarp();
                        tile_fill = 0;
                    }
                }
#else
                    if (active) {
                        const u64 row = tid * (u64)A.n_snapshots + snap;
                        if (A.out_t) A.out_t[row] = t;
                        if (A.out_x) {
#pragma unroll
                            for (int i = 0; i < SDE_D; ++i) A.out_x[row * SDE_D + i] = x[i];
                        }
                        if (A.out_w) {
#pragma unroll
                            for (int j = 0; j < SDE_M; ++j) A.out_w[row * SDE_M + j] = w[j];
                        }
                    }
#endif
                    ++snap;
                }
            }
        }
#if SDE_HAS_EVENT
        if (active && A.out_hit) {
            real* o = A.out_hit + tid * (u64)(2 + SDE_D);
            o[0] = hit ? RC(1.0) : RC(0.0);
            o[1] = hit_t;
#pragma unroll
            for (int i = 0; i < SDE_D; ++i) o[2 + i] = hit_x[i];
        }
#endif
#if SDE_NOBS > 0 && !SDE_OBS_SNAP
        sde_observe(x, w, t, xref, obs);
#endif
#if SDE_SNAP_TMA
        snapw.finish();
#endif
    }
#if SDE_NOBS > 0 && !SDE_OBS_SNAP
    // block reduction of sum and sum of squares, one atomicAdd pair per observable per block
    __shared__ double red[2 * SDE_NOBS][SDE_BLOCK / 32];
    const int red_lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < SDE_NOBS; ++k) {
        double s1 = active ? obs[k] : 0.0;
        double s2 = s1 * s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        if (red_lane == 0) {
            red[2 * k][wid] = s1;
            red[2 * k + 1][wid] = s2;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * SDE_NOBS && A.moments) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < SDE_BLOCK / 32; ++v) s += red[threadIdx.x][v];
        atomicAdd(&A.moments[threadIdx.x], s);
    }
#endif
}
