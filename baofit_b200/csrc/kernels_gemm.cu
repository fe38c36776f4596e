// FP64 tensor-pipe (DMMA, mma.sync.m8n8k4.f64) GEMM kernels.
//
// sm_100a has no tcgen05 kind for f64; the Blackwell FP64 tensor path is DMMA.8x8x4 (SASS), which
// ptxas emits for mma.sync.aligned.m8n8k4.f64 (larger PTX shapes are split into the same SASS op).
// Measured issue-bound rate on B200: 37.1 TFLOP/s (profiles/fp64_peaks_r01.json), cuBLAS DGEMM
// 35.5 TFLOP/s.
//
// All three contractions of the hot path have the same shape: a small, L2-resident left matrix
// (Hankel weights, kept rows of the distortion matrix, W = L^-1 of the covariance) times a wide
// bin-major panel whose columns are the parameter vectors:
//     C[M x N] = A[M x K] . B[K x N],   N = batch.
// Tiles: CTA 64 x 64, k depth 16, three cp.async stages; 4 warps in a 2 x 2 grid, each warp a
// 32 x 32 tile = 4 x 4 DMMA fragments. Shared-memory rows are padded by 4 doubles so that every
// fragment load is bank-conflict free.
#include <algorithm>

#include "kernels.h"

namespace bf {

namespace {
constexpr int BM = kGemmBM, BN = kGemmBN, BK = kGemmBK;
constexpr int STAGES = 3;
constexpr int LDA_S = BK + 4;  // 20 doubles: conflict-free A fragments
// Tile geometry: TB x TB output tile per CTA of 4 warps (2 x 2, each TB/2 x TB/2 = (TB/16)^2 DMMA fragments). TB = 64 is the
// throughput shape; TB = 32 is used when a batch has so few columns that 64 x 64 tiles leave most SMs without a CTA: a
// tile's K loop is a serial chain (64 DMMAs per warp and k-tile at TB = 64, 16 at TB = 32), and the lock-step rounds of a
// scan spend their time in exactly these chains. The order of the sums over k does not depend on TB.
template <int TB> struct Tile {
    static constexpr int LDB = TB + 4, A_STAGE = TB * LDA_S, B_STAGE = BK * LDB, FR = TB / 16;
    static constexpr int SMEM = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);
};
constexpr int SMEM_BYTES = Tile<BM>::SMEM;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// loads one (TB x BK) tile of A and one (BK x TB) tile of B into a pipeline stage
template <int TB>
__device__ __forceinline__ void load_stage(double *As, double *Bs, const double *__restrict__ A,
                                           const double *__restrict__ Bm, int lda, int ldb, int m0, int n0, int k0,
                                           int tid) {
#pragma unroll
    for (int c = tid; c < TB * (BK / 2); c += 128) {
        int row = c / (BK / 2), ch = c % (BK / 2);
        cp_async16(As + row * LDA_S + ch * 2, A + (size_t)(m0 + row) * lda + k0 + ch * 2);
    }
#pragma unroll
    for (int c = tid; c < BK * (TB / 2); c += 128) {
        int row = c / (TB / 2), ch = c % (TB / 2);
        cp_async16(Bs + row * Tile<TB>::LDB + ch * 2, Bm + (size_t)(k0 + row) * ldb + n0 + ch * 2);
    }
}

template <int TB>
__device__ __forceinline__ void compute_stage(const double *As, const double *Bs, double (&acc)[TB / 16][TB / 16][2], int wm,
                                              int wn, int lane) {
    constexpr int FR = TB / 16, LDB = Tile<TB>::LDB;
    const int l4 = lane & 3, l8 = lane >> 2;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
        double a[FR], b[FR];
#pragma unroll
        for (int mi = 0; mi < FR; ++mi) a[mi] = As[(wm * (TB / 2) + mi * 8 + l8) * LDA_S + kk + l4];
#pragma unroll
        for (int ni = 0; ni < FR; ++ni) b[ni] = Bs[(kk + l4) * LDB + wn * (TB / 2) + ni * 8 + l8];
#pragma unroll
        for (int mi = 0; mi < FR; ++mi)
#pragma unroll
            for (int ni = 0; ni < FR; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}
}  // namespace

// ---- plain store epilogue: C = A.B --------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(128) dgemm_store_kernel(GemmArgs g) {
    pdl_sync();
    BF_GATE(g.gate);
    extern __shared__ __align__(16) double smem[];
    constexpr int FR = TB / 16, A_STAGE = Tile<TB>::A_STAGE, B_STAGE = Tile<TB>::B_STAGE;
    double *As = smem, *Bs = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
    const int n0 = blockIdx.x * TB, m0 = blockIdx.y * TB;
    const double *A = g.A + (size_t)blockIdx.z * g.strideA;
    const double *Bm = g.Bm + (size_t)blockIdx.z * g.strideB;
    double *C = g.C + (size_t)blockIdx.z * g.strideC;
    // k tiles of this row tile: the triangular part up to the diagonal, then the dense tail columns
    const int ktri = g.lower_triangular ? (g.k_tri > 0 ? g.k_tri : g.K) : g.K;
    const int Ttri = min(ktri, g.lower_triangular ? m0 + TB : ktri) / BK, T = Ttri + (g.K - ktri) / BK;
    auto kofs = [&](int t) { return t < Ttri ? t * BK : ktri + (t - Ttri) * BK; };
    double acc[FR][FR][2];
#pragma unroll
    for (int i = 0; i < FR; ++i)
#pragma unroll
        for (int j = 0; j < FR; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < T) load_stage<TB>(As + s * A_STAGE, Bs + s * B_STAGE, A, Bm, g.lda, g.ldb, m0, n0, kofs(s), tid);
        cp_async_commit();
    }
    for (int t = 0; t < T; ++t) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int tn = t + STAGES - 1;
        if (tn < T)
            load_stage<TB>(As + (tn % STAGES) * A_STAGE, Bs + (tn % STAGES) * B_STAGE, A, Bm, g.lda, g.ldb, m0, n0,
                           kofs(tn), tid);
        cp_async_commit();
        compute_stage<TB>(As + (t % STAGES) * A_STAGE, Bs + (t % STAGES) * B_STAGE, acc, wm, wn, lane);
    }
    cp_async_wait<0>();
    const int l4 = lane & 3, l8 = lane >> 2;
#pragma unroll
    for (int mi = 0; mi < FR; ++mi) {
        int row = m0 + wm * (TB / 2) + mi * 8 + l8;
#pragma unroll
        for (int ni = 0; ni < FR; ++ni) {
            int col = n0 + wn * (TB / 2) + ni * 8 + 2 * l4;
            *reinterpret_cast<double2 *>(C + (size_t)row * g.ldc + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
    }
}

// ---- chi-square epilogue: chi2[n] = sum_m (A.B)[m][n]^2 ----------------------------------------
// One CTA owns a TB-column panel and walks all row tiles, so the row reduction never leaves the
// CTA: registers -> warp shuffles -> one shared-memory exchange between the two row-warps.
template <int TB>
__global__ void __launch_bounds__(128) dgemm_chi2_kernel(GemmArgs g) {
    pdl_sync();
    BF_GATE(g.gate);
    extern __shared__ __align__(16) double smem[];
    constexpr int FR = TB / 16, A_STAGE = Tile<TB>::A_STAGE, B_STAGE = Tile<TB>::B_STAGE;
    double *As = smem, *Bs = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
    const int n0 = blockIdx.x * TB;
    // gridDim.y == 1: this CTA walks all row tiles; otherwise (small batches) only row tile blockIdx.y
    const bool split = gridDim.y > 1;
    const int mt0 = split ? blockIdx.y : 0, mt1 = split ? blockIdx.y + 1 : g.M / TB;
    // flattened (row tile, k tile) sequence so that the cp.async pipeline never drains
    // k tiles of a row tile: the triangular part up to the diagonal, then the dense tail columns (GemmArgs::k_tri)
    const int ktri = g.lower_triangular ? (g.k_tri > 0 ? g.k_tri : g.K) : g.K, tail = (g.K - ktri) / BK;
    auto tritiles = [&](int mt) { return (g.lower_triangular ? min(ktri, (mt + 1) * TB) : ktri) / BK; };
    auto ktiles = [&](int mt) { return tritiles(mt) + tail; };
    int total = 0;
    for (int mt = mt0; mt < mt1; ++mt) total += ktiles(mt);
    int lmt = mt0, lkt = 0;  // next tile to load
    auto issue = [&](int slot) {
        const int tt = tritiles(lmt);
        load_stage<TB>(As + slot * A_STAGE, Bs + slot * B_STAGE, g.A, g.Bm, g.lda, g.ldb, lmt * TB, n0,
                       lkt < tt ? lkt * BK : ktri + (lkt - tt) * BK, tid);
        if (++lkt == ktiles(lmt)) { lkt = 0; ++lmt; }
    };
    double colsum[FR][2];
#pragma unroll
    for (int j = 0; j < FR; ++j) colsum[j][0] = colsum[j][1] = 0.0;
    double acc[FR][FR][2];
#pragma unroll
    for (int i = 0; i < FR; ++i)
#pragma unroll
        for (int j = 0; j < FR; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < total) issue(s);
        cp_async_commit();
    }
    int cmt = mt0, ckt = 0;
    for (int t = 0; t < total; ++t) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (t + STAGES - 1 < total) issue((t + STAGES - 1) % STAGES);
        cp_async_commit();
        compute_stage<TB>(As + (t % STAGES) * A_STAGE, Bs + (t % STAGES) * B_STAGE, acc, wm, wn, lane);
        if (++ckt == ktiles(cmt)) {
            // row tile finished: fold its squares into the column sums, restart accumulators
#pragma unroll
            for (int mi = 0; mi < FR; ++mi)
#pragma unroll
                for (int ni = 0; ni < FR; ++ni) {
                    colsum[ni][0] = fma(acc[mi][ni][0], acc[mi][ni][0], colsum[ni][0]);
                    colsum[ni][1] = fma(acc[mi][ni][1], acc[mi][ni][1], colsum[ni][1]);
                    acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                }
            ckt = 0;
            ++cmt;
        }
    }
    cp_async_wait<0>();
    // reduce over the 8 row-lanes of each fragment column (lanes with equal lane%4)
#pragma unroll
    for (int ni = 0; ni < FR; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = colsum[ni][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            colsum[ni][e] = v;
        }
    __syncthreads();
    double *red = smem;  // [2 row-warps][TB columns]
    if (lane < 4) {
#pragma unroll
        for (int ni = 0; ni < FR; ++ni) {
            int col = wn * (TB / 2) + ni * 8 + 2 * lane;
            red[wm * TB + col] = colsum[ni][0];
            red[wm * TB + col + 1] = colsum[ni][1];
        }
    }
    __syncthreads();
    if (tid < TB) {
        int n = n0 + tid;
        double v = red[tid] + red[TB + tid];
        if (split) {
            g.chi2_part[(size_t)blockIdx.y * g.part_ld + n] = v;  // part_ld covers whole panels
        } else if (n < g.N) {
            if (g.status && g.status[n]) v = nan("");
            g.chi2[n] = v;
        }
    }
}

// chi2[n] = sum over row tiles of the partial sums, in tile order (deterministic)
__global__ void chi2_part_sum_kernel(const double *__restrict__ part, int part_ld, int tiles, int N,
                                     const int *__restrict__ status, double *__restrict__ chi2, Gate gate) {
    pdl_sync();
    BF_GATE(gate);
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double v = 0.0;
    for (int t = 0; t < tiles; ++t) v += part[(size_t)t * part_ld + n];
    if (status && status[n]) v = nan("");
    chi2[n] = v;
}

// ---- full-M chi-square GEMM ---------------------------------------------------------------------
// dgemm_chi2_kernel above walks the row tiles of A one after the other and re-reads its 64-column panel of B
// for each of them (7 times for the 448 x 528 fused tail of the DR11 QSO model: 525 MB of DRAM traffic for a
// 175 MB panel). Here one CTA holds accumulators for ALL rows of a 32-column panel (M/32 warps, each a
// 32 x 32 tile), so B is read exactly once and only the small, L2-resident left matrix is streamed per CTA.
// A is stored pre-tiled as At[k chunk][M][12], i.e. one pipeline stage is ONE contiguous TMA bulk copy
// (cp.async.bulk, SASS UBLKCP) plus 8 row copies of the panel; full/empty mbarriers, no CTA-wide barrier in
// the k loop; persistent CTAs keep the pipeline running across panels.
namespace {
constexpr int FKC = kFullmKC, FLDA = kFullmLdA, FBN = kFullmBN, FLDB = FBN + 4, FSTAGES = 4;

__device__ __forceinline__ void f_mbar_init(unsigned long long *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void f_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_arrive(unsigned long long *bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar), ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void f_tma_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
}  // namespace

__global__ void __launch_bounds__(kFullmMaxM, 1) dgemm_chi2_fullm_kernel(GemmArgs g) {
    pdl_sync();
    BF_GATE(g.gate);
    extern __shared__ __align__(128) double fsm[];
    const int M = g.M, K = g.K;
    const int a_stage = M * FLDA, b_stage = FKC * FLDB;
    double *As = fsm;                       // [FSTAGES][M][FLDA]
    double *Bs = As + FSTAGES * a_stage;    // [FSTAGES][FKC][FLDB]
    double *red = Bs + FSTAGES * b_stage;   // [M/32][FBN]
    __shared__ __align__(8) unsigned long long full_bar[FSTAGES], empty_bar[FSTAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int l4 = lane & 3, l8 = lane >> 2;
    const int nchunks = K / FKC, ntiles = (g.N + FBN - 1) / FBN;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total = my_tiles * nchunks;
    const unsigned stage_bytes = (unsigned)((a_stage + FKC * FBN) * sizeof(double));

    auto issue = [&](int gc) {  // one thread: chunk gc of this CTA's (tile, k chunk) sequence
        const int slot = gc % FSTAGES, it = gc / nchunks, chunk = gc - it * nchunks;
        const int n0 = (blockIdx.x + it * gridDim.x) * FBN;
        f_mbar_expect_tx(&full_bar[slot], stage_bytes);
        f_tma_g2s(As + slot * a_stage, g.A + (size_t)chunk * a_stage, (unsigned)(a_stage * sizeof(double)), &full_bar[slot]);
#pragma unroll
        for (int kr = 0; kr < FKC; ++kr)
            f_tma_g2s(Bs + slot * b_stage + kr * FLDB, g.Bm + (size_t)(chunk * FKC + kr) * g.ldb + n0, (unsigned)(FBN * sizeof(double)),
                      &full_bar[slot]);
    };
    if (tid == 0) {
        for (int st = 0; st < FSTAGES; ++st) { f_mbar_init(&full_bar[st], 1); f_mbar_init(&empty_bar[st], nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (my_tiles == 0) return;
    if (tid == 0)
        for (int c = 0; c < FSTAGES && c < total; ++c) issue(c);

    for (int it = 0; it < my_tiles; ++it) {
        const int n0 = (blockIdx.x + it * gridDim.x) * FBN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int t = 0; t < nchunks; ++t) {
            const int gc = it * nchunks + t, slot = gc % FSTAGES;
            if (tid == 0 && gc >= 1 && gc - 1 + FSTAGES < total) {
                f_mbar_wait(&empty_bar[(gc - 1) % FSTAGES], ((gc - 1) / FSTAGES) & 1);
                issue(gc - 1 + FSTAGES);
            }
            __syncwarp();
            f_mbar_wait(&full_bar[slot], (gc / FSTAGES) & 1);
            const double *as = As + slot * a_stage + (warp * 32 + l8) * FLDA + l4;
            const double *bs = Bs + slot * b_stage + l4 * FLDB + l8;
#pragma unroll
            for (int kk = 0; kk < FKC; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = as[mi * 8 * FLDA + kk];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = bs[kk * FLDB + ni * 8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&empty_bar[slot]);
        }
        // column sums of squares over this warp's 32 rows, then over the warps
        double cs[4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = 0.0;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) v = fma(acc[mi][ni][e], acc[mi][ni][e], v);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                cs[ni][e] = v;
            }
        __syncthreads();  // the previous tile's reduction buffer has been consumed
        if (lane < 4) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                red[warp * FBN + ni * 8 + 2 * lane] = cs[ni][0];
                red[warp * FBN + ni * 8 + 2 * lane + 1] = cs[ni][1];
            }
        }
        __syncthreads();
        if (tid < FBN) {
            const int n = n0 + tid;
            if (n < g.N) {
                double v = 0.0;
                for (int w = 0; w < nwarps; ++w) v += red[w * FBN + tid];
                if (g.status && g.status[n]) v = nan("");
                g.chi2[n] = v;
            }
        }
    }
}

void gemm_fullm_tile_A(const double *A, int M, int K, int lda, double *At) {
    const int nch = K / kFullmKC;
    for (int c = 0; c < nch; ++c)
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < kFullmLdA; ++k)
                At[((size_t)c * M + m) * kFullmLdA + k] = k < kFullmKC ? A[(size_t)m * lda + c * kFullmKC + k] : 0.0;
}

void launch_gemm_chi2_fullm(cudaStream_t s, const GemmArgs &g, int sm_count) {
    const size_t smem = ((size_t)FSTAGES * (g.M * FLDA + FKC * FLDB) + (size_t)(g.M / 32) * FBN) * sizeof(double);
    static SmemOptIn opt;
    opt.ensure(dgemm_chi2_fullm_kernel, smem);
    const int ntiles = (g.N + FBN - 1) / FBN;
    Launch(std::min(ntiles, sm_count), g.M, smem, s)(dgemm_chi2_fullm_kernel, g);
}

// ---- toy-MC noise ------------------------------------------------------------------------------
// Counter-based normal deviates keyed by (seed, toy id, element): realisations are independent of
// how toys are sharded over GPUs (replaces likely::CovarianceMatrix::sample, call site
// baofit/CorrelationAnalyzer.cc:271).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double counter_normal(unsigned long long seed, unsigned long long toy, unsigned j) {
    unsigned long long k = splitmix64(seed ^ splitmix64(toy * 0x100000001B3ull + j));
    unsigned long long k2 = splitmix64(k);
    double u1 = ((k >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0,1]
    double u2 = (k2 >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// Toy realisations d = truth + scale * L g (likely::CovarianceMatrix::sample, call site baofit/CorrelationAnalyzer.cc:271)
// as a lower-triangular DMMA GEMM over a panel of deviates: G[j][t] = deviate j of toy t (bin-major, toys contiguous),
// Y = L . G through dgemm_store_kernel (k-tiles above the diagonal skipped), then a tiled transpose into the
// toy-major table the refits gather from.
__global__ void toy_deviates_kernel(int n, unsigned long long seed, long long first_toy, int ntoy, int ldg, double *__restrict__ G) {
    pdl_sync();
    const int t = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (t >= ldg) return;
    G[(size_t)j * ldg + t] = (j < n && t < ntoy) ? counter_normal(seed, (unsigned long long)(first_toy + t), (unsigned)j) : 0.0;
}

__global__ void toy_finish_kernel(const double *__restrict__ Y, int n, int ntoy, int ldg, const double *__restrict__ truth, double scale,
                                  double *__restrict__ out) {
    pdl_sync();
    __shared__ double tile[32][33];
    const int t0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && t < ntoy) ? Y[(size_t)i * ldg + t] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, i = i0 + threadIdx.x;
        if (t < ntoy && i < n) out[(size_t)t * n + i] = fma(scale, tile[threadIdx.x][r], truth[i]);
    }
}

// fewer 64 x 64 tiles than SMs: 32 x 32 tiles quarter the serial work of a tile's K loop
static bool use_small_tiles(int M, int N, int batch, int sm_count) {
    return sm_count > 0 && (long)((N + BN - 1) / BN) * (M / BM) * batch < sm_count && M % 32 == 0;
}

void launch_gemm_store(cudaStream_t s, const GemmArgs &g, int batch, int sm_count) {
    static SmemOptIn opt, opt32;
    if (use_small_tiles(g.M, g.N, batch, sm_count)) {
        opt32.ensure(dgemm_store_kernel<32>, (size_t)Tile<32>::SMEM);
        Launch(dim3((g.N + 31) / 32, g.M / 32, batch), 128, Tile<32>::SMEM, s)(dgemm_store_kernel<32>, g);
        return;
    }
    opt.ensure(dgemm_store_kernel<BM>, (size_t)SMEM_BYTES);
    dim3 grid((g.N + BN - 1) / BN, g.M / BM, batch);
    Launch(grid, 128, SMEM_BYTES, s)(dgemm_store_kernel<BM>, g);
}

// Row tiles are split over CTAs while all (panel, row tile) CTAs fit in about two waves of 4 resident CTAs
// per SM: a 64-column panel walked by one CTA is a serial chain of M/64 * K/16 pipeline steps (183 us for the
// 448 x 528 fused tail), far longer than the rest of a small-batch call.
int gemm_chi2_split_columns(int M, int sm_count) {
    const int tiles = M / BM;
    if (tiles < 2 || sm_count <= 0) return 0;
    return (8 * sm_count / tiles) * BN;
}

void launch_gemm_chi2(cudaStream_t s, const GemmArgs &g, int sm_count) {
    static SmemOptIn opt, opt32;
    const int panels = (g.N + BN - 1) / BN, tiles = g.M / BM;
    if (g.chi2_part && g.N <= gemm_chi2_split_columns(g.M, sm_count) && g.part_ld >= panels * BN) {
        if (use_small_tiles(g.M, g.N, 1, sm_count)) {  // scratch holds M / 32 partial rows (gemm_chi2_scratch)
            opt32.ensure(dgemm_chi2_kernel<32>, (size_t)Tile<32>::SMEM);
            Launch(dim3((g.N + 31) / 32, g.M / 32, 1), 128, Tile<32>::SMEM, s)(dgemm_chi2_kernel<32>, g);
            Launch((g.N + 127) / 128, 128, 0, s)(chi2_part_sum_kernel, g.chi2_part, g.part_ld, g.M / 32, g.N, g.status, g.chi2, g.gate);
            return;
        }
        opt.ensure(dgemm_chi2_kernel<BM>, (size_t)SMEM_BYTES);
        Launch(dim3(panels, tiles, 1), 128, SMEM_BYTES, s)(dgemm_chi2_kernel<BM>, g);
        Launch((g.N + 127) / 128, 128, 0, s)(chi2_part_sum_kernel, g.chi2_part, g.part_ld, tiles, g.N, g.status, g.chi2, g.gate);
        return;
    }
    opt.ensure(dgemm_chi2_kernel<BM>, (size_t)SMEM_BYTES);
    Launch(dim3(panels, 1, 1), 128, SMEM_BYTES, s)(dgemm_chi2_kernel<BM>, g);
}

size_t toy_noise_work_doubles(int n_pad, int ntoy) {
    const int cols = std::min(round_up(ntoy, BN), kToyPassCols);
    return (size_t)2 * n_pad * cols;
}

void launch_toy_noise(cudaStream_t s, const double *L, int n, int n_pad, const double *truth, unsigned long long seed, long long first_toy,
                      int ntoy, double scale, double *out, double *work) {
    const int cols = std::min(round_up(ntoy, BN), kToyPassCols);
    double *G = work, *Y = work + (size_t)n_pad * cols;
    for (int t0 = 0; t0 < ntoy; t0 += cols) {
        const int nt = std::min(cols, ntoy - t0), ldg = round_up(nt, BN);
        Launch(dim3((ldg + 127) / 128, n_pad), 128, 0, s)(toy_deviates_kernel, n, seed, first_toy + t0, nt, ldg, G);
        GemmArgs g{};
        g.A = L; g.Bm = G; g.C = Y;
        g.M = n_pad; g.N = nt; g.K = n_pad;
        g.lda = n_pad; g.ldb = ldg; g.ldc = ldg;
        g.lower_triangular = 1;
        launch_gemm_store(s, g, 1, 0);
        Launch(dim3((nt + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, s)(toy_finish_kernel, Y, n, nt, ldg, truth, scale, out + (size_t)t0 * n);
    }
}

}  // namespace bf
