// How the FP64 datapath of a B200 SM sub-partition is shared between DMMA (mma.sync m8n8k4 f64) warps and warps that
// run dependent DFMA chains (the consumer / producer roles of the fused model + chi2 kernel).
// One CTA per SM; warp w runs on sub-partition w % 4. Each class does a fixed amount of work; the cycles until the
// class's slowest warp finishes are reported alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix tools/fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double &x, double a, double b) { asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(x) : "d"(a), "d"(b)); }

// mode bit 0: DMMA warps work, bit 1: DFMA warps work. n_dmma_warps first warps are DMMA warps.
// interleave > 0: every warp runs 1 DMMA + `interleave` DFMAs (2 chains) in one instruction stream.
__global__ void __launch_bounds__(512, 1) mix(double *out, long long *clk, int n_dmma_warps, int mode, int dmma_iters, int dfma_iters,
                                              int chains, int interleave, double a, double b, int dep) {
    const int warp = threadIdx.x >> 5;
    __shared__ long long t_end[16];
    __syncthreads();
    const long long t0 = clock64();
    double s = 0;
    if (interleave > 0) {
        double c[28][2], x0 = threadIdx.x * 1e-9, x1 = 1 + x0;
#pragma unroll
        for (int i = 0; i < 28; ++i) c[i][0] = c[i][1] = i;
        for (int it = 0; it < dmma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 28; ++i) {
                dmma884(c[i][0], c[i][1], a, b);
                if (interleave == 4) { dfma(x0, a, b); dfma(x1, a, b); dfma(x0, a, b); dfma(x1, a, b); }
                if (interleave == 2) { dfma(x0, a, b); dfma(x1, a, b); }
            }
        }
#pragma unroll
        for (int i = 0; i < 28; ++i) s += c[i][0] + c[i][1];
        s += x0 + x1;
    } else if (warp < n_dmma_warps) {
        if (mode & 1) {
            double c[28][2];
#pragma unroll
            for (int i = 0; i < 28; ++i) c[i][0] = c[i][1] = i;
            if (dep == 1)
                for (int it = 0; it < dmma_iters; ++it) {
#pragma unroll
                    for (int i = 0; i < 28; ++i) dmma884(c[i][0], c[i][1], a, b);
                }
            else if (dep == 2)  // runs of two dependent DMMAs (two k-steps per accumulator fragment)
                for (int it = 0; it < dmma_iters; it += 2) {
#pragma unroll
                    for (int i = 0; i < 28; ++i) { dmma884(c[i][0], c[i][1], a, b); dmma884(c[i][0], c[i][1], a, b); }
                }
            else  // runs of four
                for (int it = 0; it < dmma_iters; it += 4) {
#pragma unroll
                    for (int i = 0; i < 28; ++i) { dmma884(c[i][0], c[i][1], a, b); dmma884(c[i][0], c[i][1], a, b); dmma884(c[i][0], c[i][1], a, b); dmma884(c[i][0], c[i][1], a, b); }
                }
#pragma unroll
            for (int i = 0; i < 28; ++i) s += c[i][0] + c[i][1];
        }
    } else if (mode & 2) {
        double x0 = threadIdx.x * 1e-9, x1 = 1 + x0, x2 = 2 + x0, x3 = 3 + x0;
        if (chains == 2)
            for (int it = 0; it < dfma_iters; ++it) {
#pragma unroll
                for (int u = 0; u < 8; ++u) { dfma(x0, a, b); dfma(x1, a, b); }
            }
        else
            for (int it = 0; it < dfma_iters; ++it) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { dfma(x0, a, b); dfma(x1, a, b); dfma(x2, a, b); dfma(x3, a, b); }
            }
        s = x0 + x1 + x2 + x3;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) t_end[warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long m0 = 0, m1 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            if (w < n_dmma_warps || interleave > 0) m0 = max(m0, t_end[w]); else m1 = max(m1, t_end[w]);
        }
        clk[0] = m0; clk[1] = m1;
    }
}
int main() {
    double *out; long long *clk, h[2];
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&clk, 16);
    const int DI = 400;  // 400 x 28 DMMAs per DMMA warp
    printf("{");
    auto run = [&](const char *name, int warps, int ndw, int mode, int dmma_it, int dfma_it, int chains, int inter, int dep = 1) {
        for (int rep = 0; rep < 2; ++rep) { mix<<<148, warps * 32>>>(out, clk, ndw, mode, dmma_it, dfma_it, chains, inter, 1.0000001, 1e-9, dep); cudaDeviceSynchronize(); }
        cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
        const double dmmas = (double)dmma_it * 28, dfmas = (double)dfma_it * 16;
        printf("\"%s\": {\"dmma_class_clk\": %lld, \"dfma_class_clk\": %lld, \"clk_per_dmma_per_warp\": %.2f, \"clk_per_dfma_per_warp\": %.2f},\n ", name, h[0], h[1],
               h[0] / dmmas, h[1] / dfmas);
    };
    // work sized so that each class alone takes about the same time
    run("dmma_1_per_smsp_alone", 16, 4, 1, DI, 0, 2, 0);
    run("dfma_3_per_smsp_2chains_alone", 16, 4, 2, 0, DI * 28, 2, 0);          // 16 DFMA per iter
    run("mixed_1dmma_3dfma_per_smsp", 16, 4, 3, DI, DI * 28, 2, 0);
    run("dmma_2_per_smsp_alone", 16, 8, 1, DI, 0, 2, 0);
    run("dfma_2_per_smsp_2chains_alone", 16, 8, 2, 0, DI * 28, 2, 0);
    run("mixed_2dmma_2dfma_per_smsp", 16, 8, 3, DI, DI * 28, 2, 0);
    run("dfma_3_per_smsp_4chains_alone", 16, 4, 2, 0, DI * 28, 4, 0);
    run("mixed_1dmma_3dfma4chains_per_smsp", 16, 4, 3, DI, DI * 28, 4, 0);
    run("interleaved_1dmma_4dfma_8warps", 8, 8, 3, DI, 0, 2, 4);
    run("interleaved_1dmma_2dfma_8warps", 8, 8, 3, DI, 0, 2, 2);
    run("interleaved_1dmma_4dfma_4warps", 4, 4, 3, DI, 0, 2, 4);
    run("dmma_dep2_alone", 16, 4, 1, DI, 0, 2, 0, 2);
    run("dmma_dep4_alone", 16, 4, 1, DI, 0, 2, 0, 4);
    run("mixed_dep2_3dfma", 16, 4, 3, DI, DI * 28, 2, 0, 2);
    run("mixed_dep4_3dfma", 16, 4, 3, DI, DI * 28, 2, 0, 4);
    run("mixed_dep4_3dfma_4chains", 16, 4, 3, DI, DI * 28, 4, 0, 4);
    printf("\"dmma_per_warp\": %d}\n", DI * 28);
    return cudaDeviceSynchronize() != cudaSuccess;
}
