// sde_helpers.cu - small statically compiled kernels for checking the PRNG layers in isolation
// (jax.random.split / jax.random.normal restated on the device).  Compiled once per
// (SDE_F64, SDE_LAYOUT) combination; SDE_SUFFIX names the instantiation.
#define SDE_M 1
#define SDE_SCHEME 0
#include <cuda_runtime.h>

#include "sde_core.cuh"

#define SDE_CAT2(a, b) a##b
#define SDE_CAT(a, b) SDE_CAT2(a, b)

#define KEYS_KERNEL SDE_CAT(keys_split_kernel_, SDE_SUFFIX)
#define NORMAL_KERNEL SDE_CAT(normal_kernel_, SDE_SUFFIX)

__global__ void KEYS_KERNEL(sdeb::Key root, u64 n_total, u64 begin, u64 count, u32* out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const sdeb::Key k = sdeb::split_key(root, n_total, begin + i);
    out[2 * i] = k.a;
    out[2 * i + 1] = k.b;
}

__global__ void NORMAL_KERNEL(sdeb::Key key, u64 n, real* out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = sdeb::normal_elem(key, (u32)i, (u32)n);
}

extern "C" int SDE_CAT(sdeb200_keys_split_, SDE_SUFFIX)(u32 ka, u32 kb, u64 n_total, u64 begin, u64 count, u32* out,
                                                      void* stream) {
    sdeb::Key k;
    k.a = ka;
    k.b = kb;
    const unsigned grid = (unsigned)((count + 255) / 256);
    KEYS_KERNEL<<<grid, 256, 0, (cudaStream_t)stream>>>(k, n_total, begin, count, out);
    return (int)cudaGetLastError();
}

extern "C" int SDE_CAT(sdeb200_normal_, SDE_SUFFIX)(u32 ka, u32 kb, u64 n, void* out, void* stream) {
    sdeb::Key k;
    k.a = ka;
    k.b = kb;
    const unsigned grid = (unsigned)((n + 255) / 256);
    NORMAL_KERNEL<<<grid, 256, 0, (cudaStream_t)stream>>>(k, n, (real*)out);
    return (int)cudaGetLastError();
}
