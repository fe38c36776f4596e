// Fused model + distortion matrix + chi-square kernel of the headline path (BASELINE config 4).
//
//   chi2[b] = || WA . [ xi_u(b) ; coefficient rows(b) ] ||^2 ,   WA = W . [D_kept | broadband basis | -d]
//
// In the unfused chain the undistorted correlation xi_u of every parameter vector on every grid bin is
// written to HBM by two model passes (cosmology, metals) and read back -- several times -- by the DMMA GEMM.
// Here the model evaluation is the PRODUCER of the GEMM's right-hand operand: xi_u never leaves the SM.
//
// One persistent CTA of 16 warps per SM walks 32-column panels (32 parameter vectors):
//   * 8 consumer warps, two per sub-partition: each owns a (M/8 = 56) x 32 tile of the product over ALL M rows
//     (7 x 4 DMMA m8n8k4 fragments, 112 accumulator registers), so a panel needs no second pass and the 28
//     independent DMMAs of a k-step keep the FP64 pipe of the sub-partition busy on their own.
//   * 8 producer warps, two per sub-partition: lanes = the 32 vectors of the panel, warp p produces K row 8c + p of
//     every stage c: anisotropic dilation, two cubic-spline look-ups in the shared-memory basis tables and the
//     bicubic metal patch from a window of template rows staged by the warp's own cp.async copies. Rows past the
//     grid are the per-vector broadband coefficient rows written by fold_coef_kernel.
//   * lane 0 of the first producer warp also streams the left matrix, pre-tiled as At[K/4][M][4] so that a stage
//     is ONE contiguous TMA bulk copy (cp.async.bulk, SASS UBLKCP) and every fragment load is bank-conflict free
//     without padding.
// The kernel is launched at 128 registers per thread; the consumer warpgroups then take 160 and the producer
// warpgroups give back to 96 (setmaxnreg). Stages are handed over with full/empty mbarriers (full: loader's
// expect_tx + one arrival per produced row; empty: one arrival per consumer warp); there is no CTA-wide barrier in
// the main loop.
//
// Why the producer looks the way it does (measured on B200, profiles/fp64_latency_r02.json and the r02 notes in
// DESIGN.md): DFMA and DMMA share one FP64 datapath (mixed issue 35 TFLOP/s, not 70), so fusion cannot add
// arithmetic throughput; a dependent DFMA issues every 8.4 cycles, rsqrt(double) costs 75, and a first version that
// reused the unfused kernels' per-bin code (~950 instructions per row, global loads at the head of the chain) needed
// 4200-5900 cycles per row and warp -- six such warps could not feed the consumers (1.8 ms per step against 1.24 ms
// unfused). Hence: bin coordinates in shared memory, halo'd template rows so that no index ever wraps, branch-free
// spline interval search, and everything per-vector hoisted out of the row loop.
#include <algorithm>
#include <cstdlib>

#include "fast_device.cuh"

namespace bf {

namespace {
constexpr int FU_KC = 8;        // K rows per stage = producer warps
constexpr int FU_P = 8;         // producer warps
constexpr int FU_CW = 8;        // consumer warps
constexpr int FU_BN = 32;       // columns per panel
constexpr int FU_LDB = FU_BN + 4;
constexpr int FU_NS = 4;        // pipeline stages
constexpr int FU_WIN = 7;       // rows of one staged metal window (the 4-row stencils of the panel's 32 vectors:
                                // cell indices that differ by at most 3; eight rows would not fit the 227 KB)
constexpr int FU_WS = 4;        // metal windows per producer warp: the two bins being evaluated and the next two

__device__ __forceinline__ void dmma_f(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init_nofence(unsigned long long *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar), ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}

// 1/sqrt(x) for x well inside the normal range (squared separations in (Mpc/h)^2): hardware seed (2^-22) and two
// Newton steps, WITHOUT the special-case branch of rsqrt() -- that branch ends the basic block, and the two grid
// bins a producer evaluates at a time only interleave (the point of evaluating two) inside one block.
__device__ __forceinline__ double rsqrt_nr(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}

// sum_l L_l(mu) sum_n coef_n X_{l,n}(r) from the shared-memory basis tables {y, y''/2}: natural cubic spline with
// end-point clamp, the interval found without branches (x is clamped to the table, so u = 0 / u = 1 at the ends
// reproduce the end-point rows exactly). Same arithmetic as basis_correlation (fast_device.cuh) otherwise.
__device__ __forceinline__ double basis_lean(const double2 *__restrict__ tab, int nr, double r0, double inv_dr, double dx2,
                                             double r, double musq, double c1, double c2) {
    double x = (r - r0) * inv_dr;
    x = fmin(fmax(x, 0.0), (double)(nr - 1));
    const int j = min(__double2int_rz(x), nr - 2);
    const double u = x - (double)j;
    const double w0 = 1.0 - u, w1 = u, du = dx2 * u;
    const double w2 = du * fma(u, fma(-u, 1.0 / 3.0, 1.0), -2.0 / 3.0);
    const double w3 = (du * (1.0 / 3.0)) * fma(u, u, -1.0);
    const double L2 = fma(1.5, musq, -0.5), L4 = fma(musq, fma(4.375, musq, -3.75), 0.375);
    const double2 *lo = tab + j * 9, *hi = lo + 9;
    double ylo, yhi, clo, chi;
    {
        const double2 a0 = lo[0], a1 = lo[1], a2 = lo[2], b0 = hi[0], b1 = hi[1], b2 = hi[2];
        ylo = fma(c2, a2.x, fma(c1, a1.x, a0.x));
        clo = fma(c2, a2.y, fma(c1, a1.y, a0.y));
        yhi = fma(c2, b2.x, fma(c1, b1.x, b0.x));
        chi = fma(c2, b2.y, fma(c1, b1.y, b0.y));
    }
#pragma unroll
    for (int l = 1; l < 3; ++l) {
        const double Ll = l == 1 ? L2 : L4;
        const double2 a0 = lo[3 * l], a1 = lo[3 * l + 1], a2 = lo[3 * l + 2], b0 = hi[3 * l], b1 = hi[3 * l + 1], b2 = hi[3 * l + 2];
        ylo = fma(Ll, fma(c2, a2.x, fma(c1, a1.x, a0.x)), ylo);
        clo = fma(Ll, fma(c2, a2.y, fma(c1, a1.y, a0.y)), clo);
        yhi = fma(Ll, fma(c2, b2.x, fma(c1, b1.x, b0.x)), yhi);
        chi = fma(Ll, fma(c2, b2.y, fma(c1, b1.y, b0.y)), chi);
    }
    return fma(w0, ylo, w1 * yhi) + fma(w2, clo, w3 * chi);
}

// 4 template rows x NT templates against the per-vector normalisation factors, then the Catmull-Rom weights in r_par
template <int NT>
__device__ __forceinline__ double metal_patch(const double *__restrict__ base, const double (&mn)[NT > 0 ? NT : 1],
                                              const double (&wy)[4]) {
    double wsum = 0.0;
#pragma unroll
    for (int jy = 0; jy < 4; ++jy) {
        const double *row = base + jy * NT;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int q = 0; q < NT; q += 2) {
            const double2 u = *reinterpret_cast<const double2 *>(row + q);
            d0 = fma(mn[q], u.x, d0);
            d1 = fma(mn[q + 1], u.y, d1);
        }
        wsum = fma(wy[jy], d0 + d1, wsum);
    }
    return wsum;
}
}  // namespace

template <int M, int NT>
__global__ void __launch_bounds__((FU_CW + FU_P) * 32, 1) fused_model_chi2_kernel(FusedArgs a) {
    BF_GATE(a.gate);
    extern __shared__ __align__(128) unsigned char fsm_raw[];
    const DevModel &m = a.m;
    constexpr int MF = M / (8 * FU_CW);  // row fragments per consumer warp
    static_assert(M % (8 * FU_CW) == 0, "M must split into 8-row fragments over the consumer warps");
    constexpr int NTS = NT > 0 ? NT : 2;
    const unsigned tab_bytes = 2u * m.nr * 9u * (unsigned)sizeof(double2);
    const unsigned tab_span = (tab_bytes + 127u) & ~127u;
    double2 *tab = reinterpret_cast<double2 *>(fsm_raw);
    double *As = reinterpret_cast<double *>(fsm_raw + tab_span);     // [NS][KC/4][M][4]
    double *Bs = As + FU_NS * FU_KC * M;                               // [NS][KC][LDB]
    double *red = Bs + FU_NS * FU_KC * FU_LDB;                         // [2][CW][BN]
    double *crd = red + 2 * FU_CW * FU_BN;                             // [2][kgrid] r_par, r_perp^2 of the grid bins
    double *win = crd + 2 * a.kgrid;                                   // [P][WS][WIN * NT] metal row windows
    __shared__ __align__(8) unsigned long long full_bar[FU_NS], empty_bar[FU_NS], tab_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = a.K / FU_KC, ntiles = (a.B + FU_BN - 1) / FU_BN;
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = my_tiles * nchunks;
    if (tid == 0) {
        for (int s = 0; s < FU_NS; ++s) { mbar_init_nofence(&full_bar[s], 1 + FU_KC); mbar_init_nofence(&empty_bar[s], FU_CW); }
        mbar_init_nofence(&tab_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp < FU_P) {
        // =============================== producers ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;\n");
        if (my_tiles == 0) return;
        const int pw = warp;
        const PointTab &pt = a.pt;
        const int ldb = a.ldb, npts = pt.npts, kgrid = a.kgrid;
        const bool aniso = m.flags & BF_FLAG_ANISOTROPIC, decoupled = m.flags & BF_FLAG_DECOUPLED;
        // bin-static coordinates of the grid bins: shared memory, once per CTA
        for (int i = tid; i < kgrid; i += FU_P * 32) {
            const double rt = i < npts ? pt.rperp0[i] : 0.0;
            crd[i] = i < npts ? pt.rpar0[i] : 0.0;
            crd[kgrid + i] = rt * rt;
        }
        named_bar_sync(2, FU_P * 32);
        mbar_wait(&tab_bar, 0);
        const double2 *tab_pk = tab, *tab_nw = tab + m.nr * 9;
        const int nr = m.nr, ny = pt.mx_ny, nyh = ny + 3;
        const double tr_r0 = m.tr_r0, tr_inv_dr = m.tr_inv_dr, dx2 = m.tr_dr * m.tr_dr, rlo = m.rlo, rhi = m.rhi;
        const double my0 = m.metal_y0, mymax = m.metal_ymax, minv_dy = m.metal_inv_dy;
        double *mywin = win + pw * (FU_WS * FU_WIN * NTS);
        const int model_chunks = npts / FU_KC;  // chunks whose eight rows are all grid bins
        for (int it = 0; it < my_tiles; ++it) {
            const int n0 = ((int)blockIdx.x + it * (int)gridDim.x) * FU_BN;
            const int b = min(n0 + lane, a.B - 1);
            const double *pT = a.pT + b, *S = a.S + b;
            // ---- per-vector state (single redshift class: the evolution factors are per-vector constants) ----
            double amp_pk, amp_sm, c1, c2, dpi, s_par2, s_perp2;
            double mn[NT > 0 ? NT : 1];
            {
                FastVec v;
                load_fast_vec(m, pT, ldb, S[(size_t)(a.zc_trigger * 3 + 1) * ldb], v);
                const double eb = S[0], ebeta = S[(size_t)ldb], es = S[(size_t)2 * ldb];
                amp_sm = v.biasSq * eb;
                amp_pk = amp_sm * v.ampl;
                c1 = v.c1; c2 = v.c2;
                dpi = v.dv * a.vfac[0];
                const double apar = (aniso ? v.apar : v.iso) * es, aperp = (aniso ? v.aperp : v.iso) * es;
                s_par2 = apar * apar;
                s_perp2 = aperp * aperp;
                if constexpr (NT > 0) {
#pragma unroll
                    for (int c = 0; c < NT / 3; ++c) cross_metal_norms(m, pT, ldb, c, eb, ebeta, mn[3 * c], mn[3 * c + 1], mn[3 * c + 2]);
                }
            }
            // range of the velocity shift over the panel: fixes, per bin, the window of template rows its vectors touch
            double dpi_min = dpi, dpi_max = dpi;
            bool staged = false;
            if constexpr (NT > 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dpi_min = fmin(dpi_min, __shfl_xor_sync(0xffffffffu, dpi_min, o));
                    dpi_max = fmax(dpi_max, __shfl_xor_sync(0xffffffffu, dpi_max, o));
                }
                // stencil rows iy-1 .. iy+2 of all 32 vectors fit one window when their cell indices differ by <= 3
                staged = (dpi_max - dpi_min) * minv_dy < 2.5 && ny >= FU_WIN;
            }
            // The window of bin i goes into slot seq % WS of the warp's private ring with 16-byte cp.async copies issued by
            // all lanes (halo coordinates: stencil row iy - 1 is halo row iy, so a window that starts at the smallest cell
            // index of the panel never wraps). Plain cp.async, not TMA: the SM's bulk-copy queue carries the 28 KB stages
            // of the left matrix. Returns the first halo row of the window.
            auto stage_window = [&](int i, int seq) {
                const double y = fmin(fmax(crd[i] + dpi_min, my0), mymax);
                const int wlo = min(__double2int_rz((y - my0) * minv_dy), nyh - FU_WIN);
                const double *src = pt.mxh + ((size_t)i * nyh + wlo) * NTS;
                double *dst = mywin + (seq % FU_WS) * (FU_WIN * NTS);
#pragma unroll
                for (int c = lane; c < FU_WIN * NTS / 2; c += 32) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + 2 * c) : "memory");
                }
                return wlo;
            };
            // One grid bin for the 32 vectors of the panel (branch free, so that two bins interleave in one basic block).
            // mbase: template rows iy-1 .. iy+2 start at mbase + iy * NT (shared-memory window or the global array).
            auto eval_row = [&](int k, const double *mbase) {
                const double rpar = crd[k] + dpi, rt2 = crd[kgrid + k];
                const double rp2 = rpar * rpar, r2 = rp2 + rt2;
                const double ap2 = s_par2 * rp2, q2 = fma(s_perp2, rt2, ap2);
                const double rs = rsqrt_nr(r2), qs = rsqrt_nr(q2);
                const double rprime = r2 * rs, rBAO = q2 * qs;
                const double mu2 = rp2 * (rs * rs), muBAO2 = ap2 * (qs * qs);
                // BaoKSpaceCorrelationModel.cc:466-469,485-488
                const bool zero = (rprime < rlo || rprime > rhi) || (rBAO < rlo || rBAO > rhi);
                const double peak = basis_lean(tab_pk, nr, tr_r0, tr_inv_dr, dx2, rBAO, muBAO2, c1, c2);
                const double smooth = basis_lean(tab_nw, nr, tr_r0, tr_inv_dr, dx2, decoupled ? rprime : rBAO, decoupled ? mu2 : muBAO2, c1, c2);
                double val = fma(amp_pk, peak, amp_sm * smooth);
                if constexpr (NT > 0) {
                    // r_perp (hence the x stencil) is bin-static and already folded into the template rows
                    const double y = fmin(fmax(rpar, my0), mymax);
                    const double fy = (y - my0) * minv_dy;
                    const int iy = __double2int_rz(fy);
                    double wy[4];
                    catmull_weights(fy - (double)iy, wy);
                    val += metal_patch<NT>(mbase + iy * NTS, mn, wy);
                }
                return zero ? 0.0 : val;
            };
            auto publish = [&](int t, double val) {  // row 8 t + pw of the panel goes into its stage
                const int gc = it * nchunks + t, slot = gc % FU_NS;
                if (gc >= FU_NS) mbar_wait(&empty_bar[slot], ((gc / FU_NS) - 1) & 1);
                Bs[(slot * FU_KC + pw) * FU_LDB + lane] = val;
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[slot]);
            };
            // ---- grid bins, two per iteration: a single bin is a ~2000-cycle chain of dependent FP64 operations ----
            int t = 0;
            if (NT == 0 || staged) {
                int w0 = 0, w1 = 0, seq = 0;  // window starts of the two bins in flight
                if constexpr (NT > 0) {
                    if (model_chunks > 0) w0 = stage_window(pw, 0);
                    if (model_chunks > 1) w1 = stage_window(FU_KC + pw, 1);
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                }
#pragma unroll 1
                for (; t + 1 < model_chunks; t += 2) {
                    const int k0 = t * FU_KC + pw, k1 = k0 + FU_KC;
                    int w2 = 0, w3 = 0;
                    if constexpr (NT > 0) {
                        // windows of the next pair into the two slots that are not being read (WS = 4 slots)
                        if (t + 2 < model_chunks) w2 = stage_window(k1 + FU_KC, seq + 2);
                        if (t + 3 < model_chunks) w3 = stage_window(k1 + 2 * FU_KC, seq + 3);
                        asm volatile("cp.async.commit_group;\n" ::: "memory");
                        asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // this pair's windows have landed
                        __syncwarp();
                    }
                    const double *m0 = mywin + (seq % FU_WS) * (FU_WIN * NTS) - w0 * NTS;
                    const double *m1 = mywin + ((seq + 1) % FU_WS) * (FU_WIN * NTS) - w1 * NTS;
                    const double v0 = eval_row(k0, m0), v1 = eval_row(k1, m1);
                    publish(t, v0);
                    publish(t + 1, v1);
                    if constexpr (NT > 0) {
                        __syncwarp();  // both windows are free again
                        w0 = w2; w1 = w3; seq += 2;
                    }
                }
                if (t < model_chunks) {  // odd number of model chunks
                    if constexpr (NT > 0) {
                        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                        __syncwarp();
                    }
                    publish(t, eval_row(t * FU_KC + pw, mywin + (seq % FU_WS) * (FU_WIN * NTS) - w0 * NTS));
                    ++t;
                }
            } else {
                // velocity shifts too far apart for one window per bin: template rows straight from global memory
#pragma unroll 1
                for (; t < model_chunks; ++t) {
                    const int k = t * FU_KC + pw;
                    publish(t, eval_row(k, pt.mxh + (size_t)k * nyh * NTS));
                }
            }
            // ---- remaining rows: a partial chunk of grid bins, zero padding, the broadband coefficient rows ----
#pragma unroll 1
            for (; t < nchunks; ++t) {
                const int k = t * FU_KC + pw;
                double val = 0.0;
                if (k < npts) val = eval_row(k, pt.mxh + (size_t)k * nyh * NTS);
                else if (k >= kgrid) val = a.rows[(size_t)(k - kgrid) * ldb + b];
                publish(t, val);
            }
        }
        return;
    }

    // =============================== consumers ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;\n");
    if (my_tiles == 0) return;
    const int l4 = lane & 3, l8 = lane >> 2, cw = warp - FU_P;
    // lane 0 of the first consumer warp streams the left matrix: a stage is refilled right after this warp has released it,
    // as soon as the other consumer warps have too
    constexpr unsigned stage_bytes = (unsigned)(FU_KC * M * sizeof(double));
    const bool loader = cw == 0 && lane == 0;
    auto load_chunk = [&](int gc) {
        const int slot = gc % FU_NS;
        mbar_expect_tx(&full_bar[slot], stage_bytes);
        tma_bulk_g2s(As + (size_t)slot * FU_KC * M, a.At + (size_t)(gc % nchunks) * FU_KC * M, stage_bytes, &full_bar[slot]);
    };
    if (loader) {
        mbar_expect_tx(&tab_bar, tab_bytes);
        for (unsigned off = 0; off < tab_bytes; off += 32768u)
            tma_bulk_g2s(fsm_raw + off, reinterpret_cast<const unsigned char *>(a.tabs) + off, min(32768u, tab_bytes - off), &tab_bar);
        for (int g = 0; g < FU_NS && g < total; ++g) load_chunk(g);
    }
    for (int it = 0; it < my_tiles; ++it) {
        const int n0 = ((int)blockIdx.x + it * (int)gridDim.x) * FU_BN;
        double acc[MF][4][2];
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
        for (int t = 0; t < nchunks; ++t) {
            const int gc = it * nchunks + t, slot = gc % FU_NS;
            mbar_wait(&full_bar[slot], (gc / FU_NS) & 1);
            const double *as = As + slot * (FU_KC * M) + (cw * (MF * 8) + l8) * 4 + l4;
            const double *bs = Bs + slot * (FU_KC * FU_LDB) + l4 * FU_LDB + l8;
            if (a.debug != 2) {
#pragma unroll
                for (int kk = 0; kk < FU_KC / 4; ++kk) {
                    double af[MF];
#pragma unroll
                    for (int mi = 0; mi < MF; ++mi) af[mi] = as[(kk * M + mi * 8) * 4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const double bf_ = bs[kk * 4 * FU_LDB + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < MF; ++mi) dmma_f(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf_);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            if (loader && gc + FU_NS < total) {
                mbar_wait(&empty_bar[slot], (gc / FU_NS) & 1);  // all eight consumer warps are done with this stage
                load_chunk(gc + FU_NS);
            }
        }
        // column sums of squares over this warp's rows, then over the consumer warps
        double *rbuf = red + (it & 1) * (FU_CW * FU_BN);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double s = 0.0;
#pragma unroll
                for (int mi = 0; mi < MF; ++mi) s = fma(acc[mi][ni][e], acc[mi][ni][e], s);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (lane < 4) rbuf[cw * FU_BN + ni * 8 + 2 * lane + e] = s;
            }
        named_bar_sync(1, FU_CW * 32);
        if (cw == 0) {
            const int n = n0 + lane;
            if (n < a.B) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < FU_CW; ++w) s += rbuf[w * FU_BN + lane];
                if (a.status && a.status[n]) s = nan("");
                a.chi2[n] = s;
            }
        }
    }
}

void gemm_fused_tile_A(const double *A, int M, int K, int lda, double *At) {
    for (int c = 0; c < K / 4; ++c)
        for (int r = 0; r < M; ++r)
            for (int k = 0; k < 4; ++k) At[((size_t)c * M + r) * 4 + k] = A[(size_t)r * lda + c * 4 + k];
}

bool fused_chi2_supported(const FusedArgs &a) {
    const DevModel &m = a.m;
    const int nt = m.metal_kind == BF_METAL_CROSS_BICUBIC ? a.pt.mx_nt : 0;
    return a.M == 448 && a.kgrid <= 512 && a.K % FU_KC == 0 && a.kgrid % FU_KC == 0 && (nt == 0 || (nt == 12 && a.pt.mxh)) &&
           (m.metal_kind == BF_METAL_NONE || m.metal_kind == BF_METAL_CROSS_BICUBIC) && !(m.flags & BF_FLAG_RADIATION_MODEL) &&
           !m.bb[BF_BB_DM_DIST_ADD].enabled && !m.bb[BF_BB_DM_DIST_MUL].enabled && m.nr * 9 * 2 * (int)sizeof(double2) <= 73728;
}

void launch_fused_model_chi2(cudaStream_t s, const FusedArgs &a_in, int sm_count) {
    FusedArgs a = a_in;
    static const int dbg = std::getenv("BAOFIT_FUSED_DEBUG") ? std::atoi(std::getenv("BAOFIT_FUSED_DEBUG")) : 0;
    a.debug = dbg;
    constexpr int M = 448;
    const unsigned tab_bytes = 2u * a.m.nr * 9u * (unsigned)sizeof(double2), tab_span = (tab_bytes + 127u) & ~127u;
    const int nt = a.m.metal_kind == BF_METAL_CROSS_BICUBIC ? 12 : 2;
    const size_t smem = tab_span + ((size_t)FU_NS * FU_KC * M + (size_t)FU_NS * FU_KC * FU_LDB + (size_t)2 * FU_CW * FU_BN + (size_t)2 * a.kgrid +
                                    (size_t)FU_P * FU_WS * FU_WIN * nt) * sizeof(double);
    const int ntiles = (a.B + FU_BN - 1) / FU_BN, grid = std::min(ntiles, sm_count), threads = (FU_CW + FU_P) * 32;
    static SmemOptIn opt0, opt12;
    if (a.m.metal_kind == BF_METAL_CROSS_BICUBIC) {
        opt12.ensure(fused_model_chi2_kernel<M, 12>, smem);
        fused_model_chi2_kernel<M, 12><<<grid, threads, smem, s>>>(a);
    } else {
        opt0.ensure(fused_model_chi2_kernel<M, 0>, smem);
        fused_model_chi2_kernel<M, 0><<<grid, threads, smem, s>>>(a);
    }
}

}  // namespace bf
