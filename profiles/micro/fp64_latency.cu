// Dependent-issue latencies (cycles) of the instructions on the bead kernel's critical path, one warp on one SM:
// DFMA, DMUL, DMMA.8x8x4 (accumulator chain and independent issue), 64-bit shuffle, rsqrt / rcp approximations,
// shared-memory load, CTA barrier of 128 threads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

#define N 256
__device__ __forceinline__ double shfl64(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

template <int OP>
__global__ void lat(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = seed + threadIdx.x;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-3, y = 1.0 + seed * 1e-9, z = seed;
    double2 c = make_double2(seed, seed), c2 = c, c3 = c, c4 = c;
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, z);
        if (OP == 1) x = x * y;
        if (OP == 2) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.x), "+d"(c.y) : "d"(y), "d"(z));
        if (OP == 3) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.x), "+d"(c.y) : "d"(y), "d"(z));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2.x), "+d"(c2.y) : "d"(y), "d"(z));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3.x), "+d"(c3.y) : "d"(y), "d"(z));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c4.x), "+d"(c4.y) : "d"(y), "d"(z));
        }
        if (OP == 4) x = shfl64(x, (idx + 1) & 31);
        if (OP == 5) asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
        if (OP == 6) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
        if (OP == 7) { idx = (int)sm[idx] & 63; }
        if (OP == 8) asm volatile("bar.sync 1, 128;" ::: "memory");
        if (OP == 9) x = rsqrt(x);
        if (OP == 10) x = 1.0 / x;
        if (OP == 11) x = sqrt(x);
        if (OP == 12) { double h = 0.5 * x, r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); r = r * fma(-h, r * r, 1.5); x = r * fma(-h, r * r, 1.5); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + c.x + c.y + c2.x + c3.x + c4.x + idx;
}

template <int OP>
void run(const char* name, double* out, long long* cyc, int threads, double seed, int per_iter = 1) {
    lat<OP><<<1, threads>>>(out, cyc, seed);
    lat<OP><<<1, threads>>>(out, cyc, seed);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles per op   (%s)\n", name, (double)h / N / per_iter, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("DFMA dependent", out, cyc, 32, 1.0);
    run<1>("DMUL dependent", out, cyc, 32, 1.0);
    run<2>("DMMA.8x8x4 dependent (accumulator chain)", out, cyc, 32, 1e-3);
    run<3>("DMMA.8x8x4, 4 independent accumulators", out, cyc, 32, 1e-3, 4);
    run<4>("64-bit shuffle (2 SHFL) dependent", out, cyc, 32, 1.0);
    run<5>("rsqrt.approx.ftz.f64 dependent", out, cyc, 32, 1.0);
    run<6>("rcp.approx.ftz.f64 dependent", out, cyc, 32, 1.0);
    run<7>("LDS.64 + F2I dependent (pointer chase)", out, cyc, 32, 1.0);
    run<8>("bar.sync of 128 threads", out, cyc, 128, 1.0);
    run<9>("rsqrt() library", out, cyc, 32, 1.7);
    run<10>("1.0 / x", out, cyc, 32, 1.7);
    run<11>("sqrt()", out, cyc, 32, 1.7);
    run<12>("rsqrt approx + 2 Newton (pivot_rsqrt)", out, cyc, 32, 1.7);
    return 0;
}
