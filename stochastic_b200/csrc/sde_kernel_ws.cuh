// sde_kernel_ws.cuh - warp-specialised variant of the stepping kernel for schemes whose per-step random
// input is large (Wagner-Platen with m > 1: 2mp + 3m normals per step).
//
// Same entry point, arguments and results as sde_kernel.cuh.  The CTA is split into
//   * one CONSUMER warpgroup (128 threads = 128 trajectories): keeps (t, x, w) in registers, turns the staged
//     normals of a step into iterated integrals (WienerGen::wp_assemble) and applies the emitted sde_step;
//   * SDE_WS_PGROUPS PRODUCER warpgroups (3 x 128 threads by default): nothing but threefry2x32 + erf_inv,
//     writing the normals of the next step into a two-stage shared-memory ring.
// The roles need very different register files, so the producers shrink theirs and the consumers grow theirs
// with `setmaxnreg` (Hopper/Blackwell register re-allocation between warpgroups); the hand-off uses named
// barriers (full[stage]: producers arrive, consumers wait; empty[stage]: the reverse).  The integer-heavy RNG
// and the FP64-heavy integrals then overlap on every scheduler by construction instead of by the accident of
// two CTAs drifting out of phase, and the tight producer loop keeps issuing while the consumers' long
// straight-line code waits on instruction fetch.
//
// Included AFTER sde_core.cuh (with SDE_STAGE_STRIDE = 128 defined before it) and the problem's device functions.
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
#define SDE_HAS_EVENT 0
#endif
#ifndef SDE_WS_CONSUMER_REGS
#define SDE_WS_CONSUMER_REGS 160
#endif
#ifndef SDE_WS_PRODUCER_REGS
#define SDE_WS_PRODUCER_REGS 40
#endif
#ifndef SDE_WS_GROUP
#define SDE_WS_GROUP 2                // normals a producer thread generates together
#endif
#ifndef SDE_WS_PGROUPS
#define SDE_WS_PGROUPS 2              // producer warpgroups per consumer warpgroup
#endif
#ifndef SDE_WS_CTAIL
#define SDE_WS_CTAIL 0                // of the 3M normals xi, mu, phi: how many (the last ones) the consumer draws itself
#endif
#ifndef SDE_WS_SPLIT_NORMAL
#define SDE_WS_SPLIT_NORMAL 0         // 1: the producers draw UNIFORM variates (threefry: integer work only), the consumers
                                      // turn them into normals (erf_inv: FP64 work, 46 independent chains per step)
#endif
#define SDE_WS_NC 128                 // consumer threads = trajectories per CTA
#define SDE_WS_NP (128 * SDE_WS_PGROUPS)   // producer threads
static_assert(SDE_BLOCK == SDE_WS_NC + SDE_WS_NP, "SDE_BLOCK must be 128 consumers + 128 per producer warpgroup");
static_assert(SDE_STAGE_STRIDE == SDE_WS_NC, "staging rows hold one value per consumer");
static_assert(SDE_STAGE_COUNT > 0, "warp specialisation is for staged generators");

#ifndef SDE_SNAP_TMA
#define SDE_SNAP_TMA 0                // > 0: the consumers' snapshots leave in tiles through the TMA unit (sde_snap_tma.cuh)
#endif

#include "sde_hooks.cuh"
#include "sde_snap_tma.cuh"

namespace sdeb {

struct SolveArgs {   // identical to sde_kernel.cuh
    u32 key_a, key_b;
    u32 n_fragments, cpr, chunk, n_snapshots;
    u64 n_traj_total, traj_begin, traj_count;
    const real* x0;
    long long x0_stride;
    double dt, dt32;
    real* out_t;
    real* out_x;
    real* out_w;
    double* moments;
    real* out_hit;
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void st_shared(u32 addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void st_shared(u32 addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

}  // namespace sdeb

extern __shared__ __align__(16) unsigned char sde_dyn_smem[];

#if SDE_SNAP_TMA
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 2) sde_solve_kernel(const sdeb::SolveArgs A,
                                                                            const __grid_constant__ sdeb::SnapMaps TM) {
#else
extern "C" __global__ void __launch_bounds__(SDE_BLOCK, 2) sde_solve_kernel(const sdeb::SolveArgs A) {
#endif
    using namespace sdeb;
    constexpr int NC = SDE_WS_NC, NP = SDE_WS_NP, ALL = SDE_BLOCK;
    constexpr int FULL0 = 1, EMPTY0 = 3, CONS = 5;            // named barrier ids (0 is __syncthreads)
    real* const ring = reinterpret_cast<real*>(sde_dyn_smem);  // [2][SDE_STAGE_COUNT][NC]
    const bool is_consumer = threadIdx.x < NC;
    const int c = threadIdx.x & (NC - 1);                      // trajectory slot served by this thread
    const u64 tid_raw = (u64)blockIdx.x * NC + c;
    const bool active = tid_raw < A.traj_count;
    const u64 tid = active ? tid_raw : 0;                      // slots past the end redo trajectory 0, no stores
    const u32 S = A.chunk * A.cpr;
    // total steps of this launch: whole chunks only, at most n_snapshots of them
    const u64 total_steps = (u64)A.n_snapshots * A.chunk;

    Key root;
    root.a = A.key_a;
    root.b = A.key_b;
    const Key tkey = split_key(root, A.n_traj_total, A.traj_begin + tid);
#if SDE_LOGTAB_IN_SMEM
    sde_fill_logtab();
#endif

#if SDE_WS_SPLIT_NORMAL
#define SDE_WS_DRAW uniform_elem
    static_assert(SDE_WS_GROUP == 1 && SDE_WS_CTAIL == 0, "split normals: one draw at a time, none by the consumers");
#else
#define SDE_WS_DRAW normal_elem
#endif
    if (!is_consumer) {
        // ================================ producers =====================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(SDE_WS_PRODUCER_REGS));
        constexpr int PG = SDE_WS_PGROUPS;
        const int half = (threadIdx.x - NC) >> 7;              // producer warpgroup pg serves slots n = pg (mod PG)
        WienerGen gen;
        u32 f = 0, s = S;                                      // fragment number, step within the fragment
        // shared-window byte address of this thread's first slot of stage 0; a running address per step keeps the
        // store a single STS off one register instead of re-deriving the column every normal
        const u32 col0 = (u32)__cvta_generic_to_shared(ring) + (u32)((half * NC + c) * sizeof(real));
        constexpr u32 STAGE_BYTES = SDE_STAGE_COUNT * NC * sizeof(real), SLOT_BYTES = NC * sizeof(real);
        for (u64 g = 0; g < total_steps; ++g, ++s) {
            if (s == S) {                                      // first step of a fragment: re-randomise
                s = 0;
                gen.init(split_key(tkey, A.n_fragments, f++), S);
            }
            const int st = (int)(g & 1);
            if (g >= 2) bar_sync(EMPTY0 + st, ALL);            // consumers have drained this stage
            u32 dst = col0 + (u32)st * STAGE_BYTES;
            // this thread's slots: n = half, half + PG, ... ; SDE_WS_GROUP of them per iteration (independent chains)
            constexpr int G = SDE_WS_GROUP;
            constexpr int SERIES = 2 * SDE_P * SDE_M;             // zeta then eta, SDE_P * SDE_M each
            const u32 base = s * (u32)(SDE_P * SDE_M), size = S * (u32)(SDE_P * SDE_M);
            if constexpr (G == 1) {
                // one series after the other: the key is loop-invariant and the element index a running counter
                int n = half;
                u32 e = base + (u32)half;
#pragma unroll 1
                for (int ser = 0; ser < 2; ++ser) {
                    const Key k = ser ? gen.k_eta : gen.k_zeta;
                    const int end = (ser + 1) * SDE_P * SDE_M;
#pragma unroll 1
                    for (; n < end; n += PG, e += PG, dst += PG * SLOT_BYTES) st_shared(dst, SDE_WS_DRAW(k, e, size));
                    e -= (u32)(SDE_P * SDE_M);
                }
                // then the xi, mu, phi slots the consumer does not draw itself, same round-robin
#pragma unroll 1
                for (; n < SERIES + 3 * SDE_M - SDE_WS_CTAIL; n += PG, dst += PG * SLOT_BYTES) {
                    const int q = n - SERIES;
                    const Key k = (q < SDE_M) ? gen.k_xi : ((q < 2 * SDE_M) ? gen.k_mu : gen.k_phi);
                    st_shared(dst, SDE_WS_DRAW(k, s * SDE_M + (u32)(q % SDE_M), S * SDE_M));
                }
            } else
#pragma unroll 1
            for (int n = half; n < SERIES; n += PG * G) {
                Key kk[G];
                u32 ee[G], ss[G];
                real nn[G];
                int slot[G];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int ni = (n + PG * i < SERIES) ? n + PG * i : n;    // past the end: repeat (rare)
                    const bool first = ni < SDE_P * SDE_M;
                    slot[i] = ni;
                    kk[i] = first ? gen.k_zeta : gen.k_eta;
                    ee[i] = base + (u32)(first ? ni : ni - SDE_P * SDE_M);
                    ss[i] = size;
                }
                normal_group<G>(kk, ee, ss, nn);
#pragma unroll
                for (int i = 0; i < G; ++i) st_shared(dst + (u32)(slot[i] - n) * SLOT_BYTES, nn[i]);
                dst += PG * G * SLOT_BYTES;
            }
            if constexpr (G != 1) {   // xi, mu, phi: 3 SDE_M slots, this thread takes those with its parity
                static_assert(G == 1 || SDE_WS_CTAIL == 0, "consumer-drawn normals need SDE_WS_GROUP == 1");
                constexpr int T = (3 * SDE_M + PG - 1) / PG;
                Key kk[T];
                u32 ee[T], ss[T];
                real nn[T];
                int slot[T];
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    int q = half + PG * i;
                    if (q >= 3 * SDE_M) q = half;
                    slot[i] = SERIES + q;
                    kk[i] = (q < SDE_M) ? gen.k_xi : ((q < 2 * SDE_M) ? gen.k_mu : gen.k_phi);
                    ee[i] = s * SDE_M + (u32)(q % SDE_M);
                    ss[i] = S * SDE_M;
                }
                normal_group<T>(kk, ee, ss, nn);
                const u32 tail = col0 + (u32)st * STAGE_BYTES - (u32)half * SLOT_BYTES;
#pragma unroll
                for (int i = 0; i < T; ++i) st_shared(tail + (u32)slot[i] * SLOT_BYTES, nn[i]);
            }
            __threadfence_block();
            bar_arrive(FULL0 + st, ALL);
        }
        return;
    }

    // ==================================== consumers =====================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(SDE_WS_CONSUMER_REGS));
#if SDE_NOBS > 0 && !SDE_OBS_SNAP
    double obs[SDE_NOBS];
#pragma unroll
    for (int k = 0; k < SDE_NOBS; ++k) obs[k] = 0.0;
#endif
    {
        real x[SDE_D], w[SDE_M];
        real t = RC(0.0);
#pragma unroll
        for (int i = 0; i < SDE_D; ++i) x[i] = A.x0[tid * (u64)A.x0_stride + i];
#pragma unroll
        for (int j = 0; j < SDE_M; ++j) w[j] = RC(0.0);
#if SDE_OBS_REF
        real xref[SDE_D];     // state at the first snapshot
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
#if SDE_SNAP_TMA
        // snapshot tiles of the four consumer warps live behind the ring
        SnapWriter snapw;
        snapw.init(sde_dyn_smem + (size_t)2 * SDE_STAGE_COUNT * NC * sizeof(real), threadIdx.x >> 5, threadIdx.x & 31,
                   (u64)blockIdx.x * NC + (threadIdx.x & ~31u));
#endif
        WienerGen gen;                                         // without consumer-drawn normals only S is used here
        gen.S = S;
        u32 snap = 0, in_chunk = 0;
#if SDE_WS_CTAIL > 0
        u32 f = 0, s = S;
#endif
        for (u64 g = 0; g < total_steps; ++g) {
            const int st = (int)(g & 1);
#if SDE_WS_CTAIL > 0
            // the consumer's share of the step's normals, drawn while the producers finish theirs
            if (s == S) {
                s = 0;
                gen.init(split_key(tkey, A.n_fragments, f++), S);
            }
            real own[SDE_WS_CTAIL];
            {
                constexpr int CT = SDE_WS_CTAIL;
                Key kk[CT];
                u32 ee[CT], ss[CT];
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const int q = 3 * SDE_M - CT + i;
                    kk[i] = (q < SDE_M) ? gen.k_xi : ((q < 2 * SDE_M) ? gen.k_mu : gen.k_phi);
                    ee[i] = s * SDE_M + (u32)(q % SDE_M);
                    ss[i] = S * SDE_M;
                }
                normal_group<CT>(kk, ee, ss, own);
            }
            ++s;
#else
            const real* own = nullptr;
#endif
            bar_sync(FULL0 + st, ALL);                         // normals of step g are in the ring
#if SDE_WS_SPLIT_NORMAL
            {   // uniforms -> normals in place, four independent chains at a time (this thread's column only)
                real* const col = ring + (size_t)st * SDE_STAGE_COUNT * NC + c;
                constexpr int NQ = SDE_STAGE_COUNT / 4 * 4;
#pragma unroll 1
                for (int n = 0; n < NQ; n += 4) {
                    real u4[4], z4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) u4[q] = col[(n + q) * NC];
                    normals_from_uniforms<4>(u4, z4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) col[(n + q) * NC] = z4[q];
                }
                if constexpr (SDE_STAGE_COUNT % 4 != 0) {
                    constexpr int R = SDE_STAGE_COUNT % 4;
                    real ur[R], zr[R];
#pragma unroll
                    for (int q = 0; q < R; ++q) ur[q] = col[(NQ + q) * NC];
                    normals_from_uniforms<R>(ur, zr);
#pragma unroll
                    for (int q = 0; q < R; ++q) col[(NQ + q) * NC] = zr[q];
                }
            }
#endif
            Integrals I;
            // the stage is released as soon as its values sit in registers (only if a producer will wait for it)
            const bool release = g + 2 < total_steps;
            gen.wp_assemble<SDE_WS_CTAIL>(sc, I, ring + (size_t)st * SDE_STAGE_COUNT * NC + c, [&] {
                if (release) bar_arrive(EMPTY0 + st, ALL);
            }, own);
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
#if SDE_SNAP_TMA
        snapw.finish();
#endif
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
    }
#if SDE_NOBS > 0 && !SDE_OBS_SNAP
    __shared__ double red[2 * SDE_NOBS][SDE_WS_NC / 32];
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
    bar_sync(CONS, NC);
    if (threadIdx.x < 2 * SDE_NOBS && A.moments) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < SDE_WS_NC / 32; ++v) s += red[threadIdx.x][v];
        atomicAdd(&A.moments[threadIdx.x], s);
    }
#endif
}
