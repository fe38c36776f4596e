// Dependent-issue latency of DFMA / DMMA / rsqrt(double) on one warp (clock64), and the DFMA issue interval of
// one warp with 1, 2, 4, 8 independent chains. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH> __global__ void dfma_lat(double *out, long long *clk, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void dmma_lat(double *out, long long *clk, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) dmma884(c0, c1, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void rsqrt_lat(double *out, long long *clk, int iters, double a) {
    double x = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = rsqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void lds_lat(double *out, long long *clk, int iters) {
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) next[i] = (i + 33) & 1023;
    __syncwarp();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) j = next[j];
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    double *out; long long *clk, h;
    cudaMalloc(&out, 4096); cudaMalloc(&clk, 8);
    const int iters = 4096;
    printf("{");
#define RUN(name, launch, per) launch; cudaDeviceSynchronize(); launch; cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("\"%s\": %.2f, ", name, (double)h / (iters * per));
    RUN("dfma_dependent_clk", (dfma_lat<1><<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 1)
    RUN("dfma_2chains_clk_per_instr", (dfma_lat<2><<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 2)
    RUN("dfma_4chains_clk_per_instr", (dfma_lat<4><<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 4)
    RUN("dfma_8chains_clk_per_instr", (dfma_lat<8><<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 8)
    RUN("dfma_16chains_clk_per_instr", (dfma_lat<16><<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 16)
    RUN("dmma_dependent_clk", (dmma_lat<<<1, 32>>>(out, clk, iters, 1.0000001, 1e-9)), 1)
    RUN("rsqrt_f64_dependent_clk", (rsqrt_lat<<<1, 32>>>(out, clk, iters, 1e-9)), 1)
    RUN("lds_dependent_clk", (lds_lat<<<1, 32>>>(out, clk, iters)), 1)
    printf("\"iters\": %d}\n", iters);
    return cudaDeviceSynchronize() != cudaSuccess;
}
