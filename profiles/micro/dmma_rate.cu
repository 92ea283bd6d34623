// FP64 tensor-core (DMMA) instruction throughput on sm_100a by MMA shape, against plain DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
    double a[8], b[4], c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x - i);
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[t][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (SHAPE == 884)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a[0]), "d"(b[0]));
            if (SHAPE == 1684)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            if (SHAPE == 1688)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            if (SHAPE == 16816)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            if (SHAPE == 0) {   // plain DFMA: 8 independent chains x 4
#pragma unroll
                for (int i = 0; i < 4; ++i) c[t][i] = fma(a[i], b[i], c[t][i]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double fma_per_instr_per_warp, double* out, int sms, double mhz) {
    const int iters = 20000, blocks = sms * 4;
    dmma_kernel<SHAPE><<<blocks, 256>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    dmma_kernel<SHAPE><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)blocks * 8 /*warps*/ * iters * 8;
    const double fma = instr * fma_per_instr_per_warp;
    printf("%-14s %8.2f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM  %6.1f clk per warp-instr per SM  (%s)\n", name, ms,
           2 * fma / ms / 1e9, fma / (ms * 1e-3) / (mhz * 1e6) / sms, (ms * 1e-3) * (mhz * 1e6) * sms / instr,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    double* out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 256);
    printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    run<0>("dfma x4", 32.0 * 4, out, p.multiProcessorCount, mhz);
    run<884>("dmma m8n8k4", 8 * 8 * 4, out, p.multiProcessorCount, mhz);
    run<1684>("dmma m16n8k4", 16 * 8 * 4, out, p.multiProcessorCount, mhz);
    run<1688>("dmma m16n8k8", 16 * 8 * 8, out, p.multiProcessorCount, mhz);
    run<16816>("dmma m16n8k16", 16 * 8 * 16, out, p.multiProcessorCount, mhz);
    return 0;
}
