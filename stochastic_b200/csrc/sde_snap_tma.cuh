// sde_snap_tma.cuh - snapshot write-out through the Tensor Memory Accelerator (TMA).
//
// The reference returns (trajectory, snapshot, component) arrays (pychastic/sde_solver.py:279-286,295).  With one
// thread per trajectory a plain store puts 8 bytes at a stride of a whole trajectory row.  Here every warp collects
// SDE_SNAP_TMA consecutive snapshots of its 32 trajectories in shared memory - one row of SDE_SNAP_TMA reals per
// trajectory and field column, exactly the box [32 trajectories] x [SDE_SNAP_TMA elements] of a 2-D tensor view
// (trajectory, n_snapshots * width) of the output array - and ONE elected lane hands each box to the TMA unit
// (cp.async.bulk.tensor.2d, SASS UTMASTG): no transposing write-out loop, no per-element address arithmetic, rows and
// columns past the end of the arrays are clipped by the hardware.  The rows are 16..128 bytes long; the box is written
// with the TMA swizzle pattern of that width (16-byte chunk index XOR row bits), so that the 32 lanes storing the
// same column of 32 different rows spread over all banks (4 wavefronts per 8-byte store instead of 32).
//
// The host side (sdeb200_solve, csrc/sdeb200.cu) finds `sde_snap_tma_info` in the module, encodes the three tensor
// maps for the launch's output arrays and passes them as the kernel's second parameter.
//
// Macros: SDE_SNAP_TMA (snapshots per tile, power of two, row = SDE_SNAP_TMA * sizeof(real) in {16, 32, 64, 128} bytes),
//         SDE_SNAP_SWZ (CUtensorMapSwizzle of the maps: 0 none (16-byte rows), 1 32B, 2 64B, 3 128B).
#pragma once

#ifndef SDE_SNAP_SWZ
#define SDE_SNAP_SWZ 0
#endif
#if SDE_SNAP_TMA
namespace sdeb {

struct alignas(64) TensorMap { unsigned long long opaque[16]; };   // CUtensorMap
struct SnapMaps { TensorMap t, x, w; };                             // views of out_t / out_x / out_w

constexpr int SNAP_TILE = SDE_SNAP_TMA;
constexpr int SNAP_ROWB = SNAP_TILE * (int)sizeof(real);            // bytes per box row
constexpr int SNAP_BOXB = 32 * SNAP_ROWB;                           // bytes per box (32 trajectories)
constexpr int SNAP_NBOX = 1 + SDE_D + SDE_M;                        // boxes per warp: t, x components, w components
constexpr int SNAP_WARP_BYTES = SNAP_NBOX * SNAP_BOXB;
constexpr u32 SNAP_SWZ_MASK = SDE_SNAP_SWZ == 3 ? 7u : (SDE_SNAP_SWZ == 2 ? 3u : (SDE_SNAP_SWZ == 1 ? 1u : 0u));
static_assert((SNAP_TILE & (SNAP_TILE - 1)) == 0 && SNAP_ROWB >= 16 && SNAP_ROWB <= 128, "tile rows are 16..128 bytes");
static_assert(SDE_SNAP_SWZ == 0 || SNAP_ROWB == (16 << SDE_SNAP_SWZ), "the swizzle span is the row length");

struct SnapWriter {
    unsigned char* base;   // this warp's boxes (generic pointer into shared memory, 1024-byte aligned)
    u32 srow;              // shared-window address of this lane's row in box 0
    u32 xr;                // the row's swizzle term: chunk bits [4, 7) of an address are XORed with its bits [7, 10)
    u32 fill, snap0;       // snapshots in the tile, number of its first snapshot
    u64 warp_first;        // launch-relative index of the warp's first trajectory
    int lane;

    // `area`: first byte after the kernel's other dynamic shared memory; `warp_slot`: this warp's number among the
    // warps that store snapshots
    __device__ __forceinline__ void init(unsigned char* area, const int warp_slot, const int, const u64 warp_first_) {
        const u32 a = (u32)__cvta_generic_to_shared(area);
        base = area + (((a + 1023u) & ~1023u) - a) + (size_t)warp_slot * SNAP_WARP_BYTES;
        // the lane number through a volatile read: srow and xr then live in two registers for the whole run instead of
        // being re-derived from the thread index at every snapshot (five instructions of the ~20 a snapshot cost)
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        const u32 row = (u32)lane * SNAP_ROWB;
        srow = (u32)__cvta_generic_to_shared(base) + row;
        xr = ((row >> 7) & SNAP_SWZ_MASK) << 4;
        fill = 0;
        snap0 = 0;
        warp_first = warp_first_;
    }
    // the previous tile must have left shared memory before its rows are overwritten
    __device__ __forceinline__ void begin_tile(const u32 snap) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        snap0 = snap;
    }
    __device__ __forceinline__ static void st_shared(const u32 addr, const double v) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
    }
    __device__ __forceinline__ static void st_shared(const u32 addr, const float v) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
    }
    // element `i` of a field of width W whose first box is `box0`, for snapshot `fill` of the tile
    template <int W>
    __device__ __forceinline__ void put(const int box0, const int i, const real v) {
        if constexpr (W == 1) {   // one box: the swizzled column offset is the same for every scalar field of the snapshot
            st_shared(srow + (u32)box0 * SNAP_BOXB + ((fill * (u32)sizeof(real)) ^ xr), v);
        } else {
            const u32 e = fill * W + i;
            st_shared(srow + (u32)(box0 + e / SNAP_TILE) * SNAP_BOXB + (((e % SNAP_TILE) * (u32)sizeof(real)) ^ xr), v);
        }
    }
    __device__ __forceinline__ static void store_box(const TensorMap* map, const unsigned char* box, const u32 c0, const u32 c1) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<u64>(map)), "r"((u32)__cvta_generic_to_shared(box)), "r"(c0), "r"(c1)
                     : "memory");
    }
    // hand the tile (its first `fill` snapshots) to the TMA unit
    template <class Args>
    __device__ __forceinline__ void flush(const Args& A, const SnapMaps& TM) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the TMA unit
        __syncwarp();
        if (lane == 0 && warp_first < A.traj_count) {
            const u32 c1 = (u32)warp_first;
            if (A.out_t) store_box(&TM.t, base, snap0, c1);
            if (A.out_x) {
#pragma unroll
                for (int b = 0; b < SDE_D; ++b)
                    if ((u32)b * SNAP_TILE < fill * SDE_D) store_box(&TM.x, base + (1 + b) * SNAP_BOXB, snap0 * SDE_D + b * SNAP_TILE, c1);
            }
            if (A.out_w) {
#pragma unroll
                for (int b = 0; b < SDE_M; ++b)
                    if ((u32)b * SNAP_TILE < fill * SDE_M)
                        store_box(&TM.w, base + (1 + SDE_D + b) * SNAP_BOXB, snap0 * SDE_M + b * SNAP_TILE, c1);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        fill = 0;
    }
    // snapshot `snap` of this lane's trajectory; flushes when the tile is full or the run ends
    template <class Args>
    __device__ __forceinline__ void snapshot(const Args& A, const SnapMaps& TM, const u32 snap, const real t,
                                             const real (&x)[SDE_D], const real (&w)[SDE_M]) {
        if (fill == 0) begin_tile(snap);
        put<1>(0, 0, t);
#pragma unroll
        for (int i = 0; i < SDE_D; ++i) put<SDE_D>(1, i, x[i]);
#pragma unroll
        for (int j = 0; j < SDE_M; ++j) put<SDE_M>(1 + SDE_D, j, w[j]);
        ++fill;
        if (fill == SNAP_TILE || snap + 1 == A.n_snapshots) flush(A, TM);
    }
    // shared memory must outlive the TMA unit's reads
    __device__ __forceinline__ void finish() {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
};

}  // namespace sdeb

// read by the host at module load: {tile, sizeof(real), d, m, swizzle}
extern "C" __device__ int sde_snap_tma_info[5] = {SDE_SNAP_TMA, (int)sizeof(real), SDE_D, SDE_M, SDE_SNAP_SWZ};
#endif
