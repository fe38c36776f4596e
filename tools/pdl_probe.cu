// Micro-measurements of programmatic dependent launch (griddepcontrol) on one stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_probe tools/pdl_probe.cu
// (1) cost of one kernel boundary in a chain of small dependent kernels: plain launches, PDL with the trigger at the top of
//     the kernel (launch_dependents; wait), PDL with the wait only (the trigger is then implicit at grid completion);
// (2) a chain of MULTI-WAVE kernels (each CTA busy for ~10 us, 10 waves at full occupancy): does an early trigger let parked
//     CTAs of the next grid take the slots of waves of the current grid that have not started yet?
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: plain, 1: trigger at top + wait, 2: wait only
template <int MODE>
__global__ void link_kernel(const int *gate, double *x, int n, int want, long long spin) {
    if (MODE == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (MODE >= 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (spin > 0) {
        long long t0 = clock64();
        while (clock64() - t0 < spin) { }
    }
    if (*gate != want) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] * 1.0000001 + 1.0;
}

template <int MODE>
static void launch(cudaStream_t s, const int *gate, double *x, int n, int want, int blocks, int threads, long long spin) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = MODE >= 1 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, link_kernel<MODE>, gate, x, n, want, spin);
}

template <int MODE>
static float chain(cudaStream_t s, const int *gate, double *x, int n, int want, int blocks, int threads, long long spin, int links, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 2; ++r) for (int l = 0; l < links; ++l) launch<MODE>(s, gate, x, n, want, blocks, threads, spin);
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int r = 0; r < reps; ++r) for (int l = 0; l < links; ++l) launch<MODE>(s, gate, x, n, want, blocks, threads, spin);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / (reps * links);
}

int main() {
    cudaStream_t s;
    cudaStreamCreate(&s);
    int *gate; double *x;
    const int n = 37888;
    cudaMalloc(&gate, 4); cudaMemset(gate, 0, 4);
    cudaMalloc(&x, n * sizeof(double)); cudaMemset(x, 0, n * sizeof(double));
    printf("{\"boundary_us\": {");
    const int blocks_list[3] = {1, 148, 592};
    for (int bi = 0; bi < 3; ++bi) {
        int blocks = blocks_list[bi];
        printf("%s\"blocks_%d\": {", bi ? ", " : "", blocks);
        for (int want = 1; want >= 0; --want) {
            float a = chain<0>(s, gate, x, n, want, blocks, 256, 0, 20, 50);
            float b = chain<1>(s, gate, x, n, want, blocks, 256, 0, 20, 50);
            float c = chain<2>(s, gate, x, n, want, blocks, 256, 0, 20, 50);
            printf("%s\"%s\": {\"plain\": %.3f, \"trigger_at_top\": %.3f, \"wait_only\": %.3f}", want ? "" : ", ", want ? "gated_off" : "small_work", a, b, c);
        }
        printf("}");
    }
    // multi-wave chain: 148 SMs x 2 CTAs of 1024 threads x 10 waves, ~10 us per CTA
    const long long spin = 19000;
    const int blocks = 148 * 2 * 10;
    float a = chain<0>(s, gate, x, n, 1, blocks, 1024, spin, 6, 5);
    float b = chain<1>(s, gate, x, n, 1, blocks, 1024, spin, 6, 5);
    float c = chain<2>(s, gate, x, n, 1, blocks, 1024, spin, 6, 5);
    printf("}, \"multiwave_kernel_us\": {\"plain\": %.2f, \"trigger_at_top\": %.2f, \"wait_only\": %.2f}", a, b, c);
    // one-wave primary followed by a multi-wave dependent and back (mixed chain)
    printf("}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
