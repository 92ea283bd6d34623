// Which hardware warp slot (%warpid) - and hence which of the SM's four schedulers (slot mod 4) - does each warp of a CTA get?
// Two 512-thread CTAs per SM (like the warp-specialised stepping kernel), printed for the CTAs on SM 0.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu && ./warp_slots
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ unsigned char smem[];
__global__ void __launch_bounds__(512, 2) k(int* out) {
    unsigned smid, slot;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    smem[threadIdx.x] = 1;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && smid < 2) out[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 2] = (int)smid, out[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 2 + 1] = (int)slot;
    // keep the CTA alive long enough for the second wave to be co-resident
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {}
}
int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = 2 * sms;
    int* out;
    cudaMalloc(&out, grid * 32 * sizeof(int));
    cudaMemset(out, 0xff, grid * 32 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k<<<grid, 512, 100 * 1024>>>(out);
        int* h = new int[grid * 32];
        cudaMemcpy(h, out, grid * 32 * sizeof(int), cudaMemcpyDeviceToHost);
        for (int b = 0; b < grid; ++b)
            if (h[b * 32] == 0 || h[b * 32] == 1) {
                printf("launch %d CTA %3d on SM %d, warp -> slot (scheduler):", rep, b, h[b * 32]);
                for (int w = 0; w < 16; ++w) printf(" %d(%d)", h[(b * 16 + w) * 2 + 1], h[(b * 16 + w) * 2 + 1] & 3);
                printf("\n");
            }
        cudaMemset(out, 0xff, grid * 32 * sizeof(int));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
