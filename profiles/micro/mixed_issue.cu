// Issue ceiling of a mixed INT32 / FP64 instruction stream on one SM sub-partition: does a DFMA cost one issue slot or the
// ~2.2 cycles its half-rate pipe needs?  1024 threads per SM (8 warps per scheduler, like the warp-specialised stepping
// kernel), four independent chains per thread, loop bodies with a fixed ratio of integer (IADD3 / LOP3 / SHF, the
// threefry mix) and FP64 (DFMA) instructions.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mixed_issue mixed_issue.cu && ./mixed_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int NF>
__global__ void __launch_bounds__(512, 2) k(double* out, unsigned* iout, int iters) {
    unsigned a[4], b[4];
    double x[4], y = 1.0000001, z = 1e-9;
#pragma unroll
    for (int c = 0; c < 4; ++c) { a[c] = threadIdx.x * 7 + c; b[c] = threadIdx.x * 13 + 5 * c; x[c] = 1.0 + 1e-3 * (threadIdx.x + c); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < NI / 12; ++r) {        // 12 integer instructions: 4 chains x (add, rotate, xor)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a[c] += b[c];
                b[c] = __funnelshift_l(b[c], b[c], 13) ^ a[c];
            }
        }
#pragma unroll
        for (int r = 0; r < NF / 4; ++r) {         // 4 DFMA: one per chain
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = fma(x[c], y, z);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
    iout[blockIdx.x * blockDim.x + threadIdx.x] = a[0] ^ a[1] ^ a[2] ^ a[3] ^ b[0] ^ b[1] ^ b[2] ^ b[3];
}

template <int NI, int NF>
void run(double* out, unsigned* iout, int sms, double mhz) {
    const int iters = 4000, blocks = 2 * sms;
    k<NI, NF><<<blocks, 512>>>(out, iout, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NI, NF><<<blocks, 512>>>(out, iout, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * mhz * 1e6;                       // per SM sub-partition: 8 warps each
    const double instr = 8.0 * iters * (NI + NF);                      // warp-instructions per sub-partition (loop overhead not counted)
    printf("%3d integer + %3d DFMA per iteration (%4.1f %% FP64): %.3f instructions / cycle / scheduler, %.3f DFMA / cycle / scheduler  (%s)\n",
           NI, NF, 100.0 * NF / (NI + NF), instr / cycles, 8.0 * iters * NF / cycles, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    double* out;
    unsigned* iout;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    cudaMalloc(&iout, sizeof(unsigned) * p.multiProcessorCount * 1024);
    printf("%s, %d SMs, %.0f MHz, 8 warps per scheduler\n", p.name, p.multiProcessorCount, mhz);
    run<96, 0>(out, iout, p.multiProcessorCount, mhz);
    run<0, 40>(out, iout, p.multiProcessorCount, mhz);
    run<96, 8>(out, iout, p.multiProcessorCount, mhz);
    run<96, 24>(out, iout, p.multiProcessorCount, mhz);
    run<96, 40>(out, iout, p.multiProcessorCount, mhz);
    run<96, 68>(out, iout, p.multiProcessorCount, mhz);     // 41.5 % FP64: the mix of the headline kernel
    run<96, 96>(out, iout, p.multiProcessorCount, mhz);
    run<48, 96>(out, iout, p.multiProcessorCount, mhz);
    return 0;
}
