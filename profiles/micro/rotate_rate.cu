// Throughput of the threefry2x32 round (x0 += x1; x1 = rotl(x1, r) ^ x0) with the rotation spelled three ways:
//   0: funnel shift (SHF.L.W) + LOP3            (what nvcc emits for __funnelshift_l / the usual shift-or idiom)
//   1: 32 x 32 -> 64 bit multiply by 2^r (IMAD.WIDE: low word = x << r, high word = x >> (32 - r)) + ONE three-input
//      LOP3 (lo | hi) ^ x0
//   2: byte permute (PRMT) for rotations by 16 and 24, funnel shift otherwise
// 8 warps per scheduler, four independent chains per thread, the 8 rotation constants of threefry2x32.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rotate_rate rotate_rate.cu && ./rotate_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int V, int R>
__device__ __forceinline__ void round_fn(unsigned& x0, unsigned& x1) {
    x0 += x1;
    if (V == 0 || (V == 2 && R != 16 && R != 24)) {
        x1 = __funnelshift_l(x1, x1, R) ^ x0;
    } else if (V == 1) {
        unsigned lo, hi;
        asm("{\n .reg .b64 w;\n mul.wide.u32 w, %2, %3;\n mov.b64 {%0, %1}, w;\n}" : "=r"(lo), "=r"(hi) : "r"(x1), "r"(1u << R));
        asm("lop3.b32 %0, %1, %2, %3, 0x56;" : "=r"(x1) : "r"(lo), "r"(hi), "r"(x0));      // (a | b) ^ c
    } else {
        x1 = __byte_perm(x1, x1, R == 16 ? 0x1032 : 0x0321) ^ x0;                             // rotl by 16 / 24
    }
}

template <int V>
__global__ void __launch_bounds__(512, 2) k(unsigned* out, int iters) {
    unsigned a[4], b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { a[c] = threadIdx.x * 7 + c; b[c] = threadIdx.x * 13 + 5 * c + blockIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            round_fn<V, 13>(a[c], b[c]); round_fn<V, 15>(a[c], b[c]); round_fn<V, 26>(a[c], b[c]); round_fn<V, 6>(a[c], b[c]);
            round_fn<V, 17>(a[c], b[c]); round_fn<V, 29>(a[c], b[c]); round_fn<V, 16>(a[c], b[c]); round_fn<V, 24>(a[c], b[c]);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] ^ a[1] ^ a[2] ^ a[3] ^ b[0] ^ b[1] ^ b[2] ^ b[3];
}

template <int V>
void run(const char* name, unsigned* out, int sms, double mhz) {
    const int iters = 4000, blocks = 2 * sms;
    k<V><<<blocks, 512>>>(out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<blocks, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned h;
    cudaMemcpy(&h, out + 77, 4, cudaMemcpyDeviceToHost);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-44s %.3f rounds / cycle / scheduler  (checksum %08x, %s)\n", name, 8.0 * iters * 32 / cycles, h, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned* out;
    cudaMalloc(&out, sizeof(unsigned) * p.multiProcessorCount * 1024);
    run<0>("funnel shift + LOP3", out, p.multiProcessorCount, khz / 1e3);
    run<1>("IMAD.WIDE by 2^r + three-input LOP3", out, p.multiProcessorCount, khz / 1e3);
    run<2>("PRMT for 16 / 24, funnel shift otherwise", out, p.multiProcessorCount, khz / 1e3);
    return 0;
}
