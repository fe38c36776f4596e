// Measures the FP64 roofline denominators that MEASURED_PEAKS.json lacks: DFMA (vector pipe)
// and DMMA (mma.sync f64, the "FP64 tensor pipe") issue-bound throughput on the whole chip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma884_kernel(double *out, int iters, double a, double b) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef HAVE_M16
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void dmma16816_kernel(double *out, int iters, double av, double bv) {
    double c[8][4], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = av + i; for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + j; }
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = bv + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// Do DMMA and DFMA overlap (separate pipes) or share one FP64 datapath? Even warps issue DMMA, odd warps DFMA;
// if the combined rate stays at the single-pipe rate the two share the unit and a fused producer/consumer kernel
// can only hide latency, not add throughput.
__global__ void mixed_kernel(double *out, int iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double x[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i][0] = threadIdx.x * 1e-9 + i; x[i][1] = i; }
    if (warp & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int i = 0; i < 16; ++i) { x[i][0] = fma(x[i][0], a, b); x[i][1] = fma(x[i][1], a, b); }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(x[i][0], x[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i][0] + x[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 512);
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int wpb : {4, 8, 16}) {
        int threads = wpb * 32, blocks = sms * 4;
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 16 * iters * (double)threads * blocks;
        printf(", \"dfma_tflops_w%d\": %.3f", wpb, flops / ms * 1e-9);
        ms = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        flops = 2.0 * 256 * 16 * iters * (double)wpb * blocks;
        printf(", \"dmma884_tflops_w%d\": %.3f", wpb, flops / ms * 1e-9);
        // mixed: half the warps DMMA (2*256*16 flop per warp-iteration), half DFMA (2*128*32 flop per warp-iteration: the same pipe time if shared)
        ms = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        flops = (2.0 * 256 * 16 + 2.0 * 128 * 32) * iters * (double)(wpb / 2) * blocks;
        printf(", \"mixed_dmma_dfma_tflops_w%d\": %.3f", wpb, flops / ms * 1e-9);
#ifdef HAVE_M16
        ms = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        flops = 2.0 * 16 * 8 * 16 * 8 * iters * (double)wpb * blocks;
        printf(", \"dmma16816_tflops_w%d\": %.3f", wpb, flops / ms * 1e-9);
#endif
    }
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
