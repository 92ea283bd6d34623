/* sdeb200.h - C ABI of the B200-native batched SDE integrator (libsdeb200.so).
 *
 * The reference (PyChastic, pure Python on JAX) has no FFI/plugin boundary of its own; its
 * drop-in surface is the Python API `SDESolver.solve_many` (reference
 * pychastic/sde_solver.py:82-304).  This header is the native boundary underneath the
 * re-implemented Python API: plain pointers and sizes, no torch/JAX types.  A maintainer of
 * the reference would bind these with ctypes (stub in INTEGRATION.md) and call
 * `sdeb200_solve` from `solve_many` in place of the `jax.vmap(get_solution)` dispatch at
 * pychastic/sde_solver.py:297-300.
 *
 * Conventions: every function returns int (0 = ok, <0 = invalid argument, >0 = CUDA / NVRTC
 * failure; text via sdeb200_last_error()).  The caller owns all buffers; device pointers are
 * passed as raw addresses; launches are asynchronous on the given cudaStream_t (passed as
 * void*), no hidden host synchronisation.
 */
#ifndef SDEB200_H
#define SDEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEB200_ABI_VERSION 4   /* 4: TMA-storing programs (second kernel parameter built by sdeb200_solve), bead workspace carries zeroed CTA counters */

typedef struct sdeb200_program sdeb200_program_t; /* a compiled (problem, scheme, dtype) kernel */

/* Mirrors sdeb::SolveArgs (stochastic_b200/csrc/sde_kernel.cuh) field by field.
 * Replaces the arguments of the reference's get_solution/vmap dispatch:
 *   key fan-out            pychastic/sde_solver.py:223,292,299
 *   chunk arithmetic       pychastic/sde_solver.py:217-221
 *   snapshot arrays        pychastic/sde_solver.py:279-286,295                      */
typedef struct sdeb200_solve_args {
    uint32_t key_a, key_b;   /* jax.random.PRNGKey(seed) = (seed >> 32, seed & 0xffffffff) */
    uint32_t n_fragments;    /* number_of_chunks / chunks_per_randomization */
    uint32_t cpr;            /* chunks per randomization */
    uint32_t chunk;          /* steps per chunk (snapshot stride) */
    uint32_t n_snapshots;    /* snapshots integrated and stored (<= n_fragments * cpr) */
    uint64_t n_traj_total;   /* N of split(PRNGKey(seed), N) - fixes every trajectory key */
    uint64_t traj_begin;     /* global index of the first trajectory of this launch (sharding) */
    uint64_t traj_count;     /* trajectories integrated by this launch */
    const void* x0;          /* device: initial conditions, working dtype */
    int64_t x0_stride;       /* elements between consecutive ICs; 0 = broadcast one IC */
    double dt;
    double dt32;             /* dt ** 1.5 evaluated in double (pychastic/sde_solver.py:239-241) */
    void* out_t;             /* device (traj_count, n_snapshots)      or NULL */
    void* out_x;             /* device (traj_count, n_snapshots, d)   or NULL */
    void* out_w;             /* device (traj_count, n_snapshots, m)   or NULL */
    double* moments;         /* device (n_obs, 2) sums / sums of squares, accumulated - (n_snapshots, n_obs, 2) for a
                                program built with SDE_OBS_SNAP (observables at every snapshot); or NULL */
    void* out_hit;           /* device (traj_count, 2 + d): first-passage flag, time, state (kernels built with an
                                event function; replaces the host post-processing of examples/hit.py:40-58); or NULL */
} sdeb200_solve_args_t;

/* Mirrors sdeb::WienerArgs: replaces get_wiener_integrals
 * (pychastic/vectorized_I_generation.py:243-473) for `steps` unit steps of one key. */
typedef struct sdeb200_wiener_args {
    uint32_t key_a, key_b;
    uint32_t steps;
    uint32_t _pad;
    void* d_w;    /* (steps, m) */
    void* d_ww;   /* (steps, m, m)     or NULL */
    void* d_wt;   /* (steps, m)        or NULL */
    void* d_tw;   /* (steps, m)        or NULL */
    void* d_www;  /* (steps, m, m, m)  or NULL */
} sdeb200_wiener_args_t;

/* Mirrors sdeb::BeadArgs (stochastic_b200/csrc/sde_bead_kernel.cuh): Euler-Maruyama Brownian dynamics of a
 * closed chain of N equal beads with RPY mobility, drift = mu (-grad U), noise = sqrt(2) chol(mu) - the built-in
 * functor for reference examples/example_writhed_dna_loop.py:49-86 (stretch + bend + overlap energy, and - for a
 * program compiled with SDE_BEAD_TWIST - the writhe-dependent twist energy of :57). */
typedef struct sdeb200_bead_args {
    uint32_t key_a, key_b;
    uint32_t n_fragments, cpr, chunk, n_snapshots;
    uint64_t n_traj_total, traj_begin, traj_count;
    const void* x0;          /* device (n_ic, 3N) working dtype */
    int64_t x0_stride;       /* 0 = broadcast */
    double dt;
    double radius;           /* hydrodynamic radius of every bead */
    double ks, d0;           /* stretch stiffness and rest length */
    double kb_over_d04;      /* bending stiffness / d0^4 */
    double ko;               /* overlap stiffness */
    double sigma2;           /* steric diameter squared */
    double inv_delta2;       /* 1 / (overlap smear width)^2 */
    double twist_k;          /* twisting stiffness x (2 pi)^2: U_twist = twist_k / 2 (Lk - Wr)^2 (ignored without SDE_BEAD_TWIST) */
    double linking_number;   /* Lk */
    void* workspace;         /* device scratch of the register-tiled variant (SDE_BEAD_REGTILE): 128 bytes per 8x8 tile of
                                the lower block triangle (written and read by the kernel itself), followed by 1024
                                uint32 CTA arrival counters that the caller zeroes before every launch; NULL otherwise */
    void* out_t;             /* device (traj_count, n_snapshots)      or NULL */
    void* out_x;             /* device (traj_count, n_snapshots, 3N)  or NULL */
    void* out_w;             /* device (traj_count, n_snapshots, 3N)  or NULL */
} sdeb200_bead_args_t;

int sdeb200_abi_version(void);
const char* sdeb200_last_error(void);

/* Select the CUDA device for the calling thread (cudaSetDevice + primary context). */
int sdeb200_set_device(int device);

/* Compile CUDA C++ `source` for sm_100a with NVRTC (include path `include_dir` holds
 * sde_core.cuh / sde_kernel.cuh) and return the cubin in a malloc'ed buffer the caller frees
 * with sdeb200_free.  Needs no GPU.  `log` (optional) receives the NVRTC log. */
int sdeb200_compile(const char* source, const char* include_dir, const char* arch,
                    void** cubin, size_t* cubin_size, char** log);
void sdeb200_free(void* p);
/* Which NVRTC sdeb200_compile uses (it is bound with dlopen: $SDEB200_NVRTC, then the CUDA toolkit's
 * libnvrtc, then the soname - independent of what the host process loaded before).  `path` (optional)
 * receives a pointer to a static string. */
int sdeb200_nvrtc_version(int* major, int* minor, const char** path);

/* Load a cubin on the current device and look up the kernel `entry`. */
int sdeb200_program_load(const void* cubin, size_t cubin_size, const char* entry,
                         sdeb200_program_t** out);
int sdeb200_program_destroy(sdeb200_program_t* prog);
/* Registers per thread / static shared memory of the loaded kernel (diagnostics). */
int sdeb200_program_info(sdeb200_program_t* prog, int* regs, int* smem_bytes, int* max_threads);

/* Launch the stepping kernel - replaces `jax.vmap(get_solution)(keys, initial_conditions)`,
 * pychastic/sde_solver.py:288-300, i.e. the scan nest :228-286 for every trajectory of the launch.
 * `block` threads per CTA (must equal the SDE_BLOCK the program was compiled
 * with), `traj_per_block` trajectories per CTA (0 = block: one thread per trajectory; 128 for the
 * warp-specialised variant whose 512-thread CTAs hold 128 consumer + 384 producer threads),
 * `dyn_smem_bytes` of dynamic shared memory (the staging area of the in-kernel normal generator). */
int sdeb200_solve(sdeb200_program_t* prog, const sdeb200_solve_args_t* args, int block, int traj_per_block,
                  size_t dyn_smem_bytes, void* stream);

/* Launch the bead-chain kernel (entry `sde_bead_kernel`) - the same dispatch for the drift / noise functors of
 * examples/example_writhed_dna_loop.py:49-74 (RPY mobility, Cholesky noise): one CTA of `block` threads per trajectory, at most
 * `max_ctas` CTAs (grid-stride over trajectories; 0 = one CTA per trajectory). */
int sdeb200_bead_solve(sdeb200_program_t* prog, const sdeb200_bead_args_t* args, int block,
                       size_t dyn_smem_bytes, unsigned max_ctas, void* stream);

/* Launch the unit-step Wiener-integral generator kernel (one thread per step) - replaces
 * get_wiener_integrals, pychastic/vectorized_I_generation.py:243-473. */
int sdeb200_wiener_integrals(sdeb200_program_t* prog, const sdeb200_wiener_args_t* args, int block,
                             size_t dyn_smem_bytes, void* stream);

/* jax.random.split(PRNGKey(seed), n_total)[begin : begin+count] -> device uint32 (count, 2).
 * layout 0 = original (jax<0.5), 1 = partitionable.  pychastic/sde_solver.py:299. */
int sdeb200_keys_split(uint32_t key_a, uint32_t key_b, uint64_t n_total, uint64_t begin, uint64_t count,
                       int layout, uint32_t* out_keys, void* stream);

/* jax.random.normal(key, (n,), dtype) -> device buffer; dtype_f64 0/1; layout as above. */
int sdeb200_normal(uint32_t key_a, uint32_t key_b, uint64_t n, int dtype_f64, int layout, void* out,
                   void* stream);

/* Measured FP64 FMA throughput of the current device (dependent-chain-free DFMA loop),
 * TFLOP/s counting 2 flop per FMA: the roofline denominator for the stepping kernels. */
int sdeb200_fp64_peak(int iters, double* tflops, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* SDEB200_H */
