/* cabi_client.c - the C ABI of include/sdeb200.h used from plain C, no Python and no torch:
 * geometric Brownian motion dx = 0.2 x dt + 0.5 x dW, x0 = 1, Euler-Maruyama, fp64, final state only.
 *
 *   gcc -O2 examples/cabi_client.c -Iinclude -I/usr/local/cuda/include -Lstochastic_b200 -lsdeb200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/stochastic_b200 -o cabi_client
 *   ./cabi_client stochastic_b200/csrc [n_trajectories]
 *
 * The step function below is what the Python front end emits for this problem (stochastic_b200/trace/emit.py); a host
 * written in another language would generate or hand-write the same few lines.  Without a CUDA device the program
 * still compiles the kernel with NVRTC (needs no GPU) and stops there.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "sdeb200.h"

static const char* SOURCE =
    "#define SDE_D 1\n#define SDE_M 1\n#define SDE_SCHEME 0\n#define SDE_P 10\n#define SDE_F64 1\n"
    "#define SDE_LAYOUT 0\n#define SDE_BLOCK 128\n"
    "#include \"sde_core.cuh\"\n"
    "__device__ __forceinline__ void sde_step(real (&x)[SDE_D], const sdeb::Integrals& I, const real dt) {\n"
    "    x[0] = x[0] + (0.2 * (x[0] * dt) + 0.5 * (x[0] * I.dW[0]));\n"
    "}\n"
    "#include \"sde_kernel.cuh\"\n";

#define CHECK(call)                                                               \
    do {                                                                          \
        int st_ = (call);                                                         \
        if (st_ != 0) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, st_, sdeb200_last_error()); \
            return 1;                                                             \
        }                                                                         \
    } while (0)

int main(int argc, char** argv) {
    const char* include_dir = argc > 1 ? argv[1] : "stochastic_b200/csrc";
    const size_t n = argc > 2 ? (size_t)atoll(argv[2]) : 1u << 20;
    int major = 0, minor = 0;
    const char* path = NULL;
    CHECK(sdeb200_nvrtc_version(&major, &minor, &path));
    printf("sdeb200 ABI %d, NVRTC %d.%d (%s)\n", sdeb200_abi_version(), major, minor, path);

    void* cubin = NULL;
    size_t cubin_size = 0;
    char* log = NULL;
    CHECK(sdeb200_compile(SOURCE, include_dir, "sm_100a", &cubin, &cubin_size, &log));
    printf("compiled sde_solve_kernel for sm_100a: %zu bytes of cubin\n", cubin_size);
    sdeb200_free(log);

    int devices = 0;
    if (cudaGetDeviceCount(&devices) != cudaSuccess || devices == 0) {
        printf("no CUDA device: stopping after the compile step\n");
        sdeb200_free(cubin);
        return 0;
    }
    CHECK(sdeb200_set_device(0));
    sdeb200_program_t* prog = NULL;
    CHECK(sdeb200_program_load(cubin, cubin_size, "sde_solve_kernel", &prog));
    sdeb200_free(cubin);

    const double x0_host = 1.0, dt = 1.0 / 256, tmax = 2.0;
    const unsigned steps = (unsigned)(tmax / dt);
    double *x0 = NULL, *out_x = NULL;
    if (cudaMalloc((void**)&x0, sizeof(double)) != cudaSuccess || cudaMalloc((void**)&out_x, n * sizeof(double)) != cudaSuccess) {
        fprintf(stderr, "cudaMalloc failed\n");
        return 1;
    }
    cudaMemcpy(x0, &x0_host, sizeof(double), cudaMemcpyHostToDevice);

    sdeb200_solve_args_t args;
    memset(&args, 0, sizeof args);
    args.key_a = 0;              /* jax.random.PRNGKey(0) */
    args.key_b = 0;
    args.n_fragments = 1;        /* one randomisation, one chunk of `steps` steps, one snapshot: the final state */
    args.cpr = 1;
    args.chunk = steps;
    args.n_snapshots = 1;
    args.n_traj_total = n;
    args.traj_begin = 0;
    args.traj_count = n;
    args.x0 = x0;
    args.x0_stride = 0;          /* broadcast one initial condition */
    args.dt = dt;
    args.dt32 = pow(dt, 1.5);
    args.out_x = out_x;          /* out_t, out_w, moments, out_hit stay NULL */
    CHECK(sdeb200_solve(prog, &args, 128, 0, 0, NULL));
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    double* host = (double*)malloc(n * sizeof(double));
    cudaMemcpy(host, out_x, n * sizeof(double), cudaMemcpyDeviceToHost);
    double mean = 0.0;
    for (size_t i = 0; i < n; ++i) mean += host[i];
    mean /= (double)n;
    printf("%zu trajectories x %u steps: E[x_T] = %.5f (exact exp(0.2 T) = %.5f)\n", n, steps, mean, exp(0.2 * tmax));
    free(host);
    cudaFree(x0);
    cudaFree(out_x);
    CHECK(sdeb200_program_destroy(prog));
    return 0;
}
