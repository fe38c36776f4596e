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
    pdl_sync();
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

// ------------------------------------------------------------------------------------------------------------------
// Third generation: 32-column panels, 12 producer warps + 4 consumer warps, half of the accumulators parked in TMEM.
//
// What the first kernel's ncu capture showed (profiles/ncu_fused_r02_summary.json): the FP64 datapath is shared by the
// producers' DFMA chains and the consumers' DMMAs and was busy 66 % of the time (DMMA sub-pipe 44 %, FP64 pipe 21 %).
// Consumer warps spent half of their samples waiting for a full stage; producer warps spent 44 % of theirs on the math
// pipe throttle: whenever a stage was ready the two consumer warps of a sub-partition kept the pipe busy with 16-cycle
// DMMAs and the two producer warps -- four dependent chains per sub-partition -- got a quarter of the issue slots, then
// ran alone, latency bound, while the consumers waited. Producers need more independent chains, and the register file
// (64 K) is what limits them: a 448 x 32 accumulator panel takes 28 K registers.
// A second kernel with 16-column panels (half the accumulators, twelve producer warps whose half-warps evaluated two
// grid bins) was correct but slower, 1.31 ms: with two bins per warp every table load touches twice as many spline rows
// and the shared-memory pipe saturates (producers alone 0.84 ms against 0.71 ms). The lanes of a warp must share a bin.
// Hence: the panel keeps 32 columns, but each of the four consumer warps (one per sub-partition) owns 112 rows x 32
// columns as TWO 56-row halves and works on one half at a time, 16 K-rows per visit; the other half's 56 accumulator
// fragments wait in tensor memory (tcgen05.st / tcgen05.ld, 112 columns per lane, one swap per 16 K-rows; TMEM has its
// own datapath, the swap costs no FP64 issue slot and no shared-memory bandwidth). The visits alternate h0 h1 | h1 h0 |
// h0 h1 ... so that only one swap per chunk is needed. That frees the registers for twelve producer warps, three per
// sub-partition, with one bin per warp as in the first kernel.
namespace {
constexpr int F3_BN = 32;            // columns per panel
constexpr int F3_LDB = F3_BN + 4;
constexpr int F3_CH = 16;            // K rows per chunk = one visit of a half
constexpr int F3_P = 12;             // producer warps
constexpr int F3_CW = 4;             // consumer warps
constexpr int F3_NSA = 3;            // stages of the left-matrix ring: one stage = one visit (16 K-rows x 224 rows, 28 KB)
constexpr int F3_NSB = 6;            // chunks of the right-hand ring
constexpr int F3_RPI = 2 * F3_P;     // rows per producer iteration: 2 evaluations x 12 warps
constexpr int F3_WIN = 7;            // template rows of one staged metal window
constexpr int F3_WS = 4;             // windows per producer warp: the two bins being evaluated and the next two
constexpr int F3_TMEM_COLS = 256;    // two parking regions of 112 columns per lane

// 56 accumulator doubles of a lane <-> 112 tensor-memory columns of the lane (32x32b shape: lane i of the warp
// addresses TMEM lane base + i)
__device__ __forceinline__ void tmem_park(unsigned taddr, const double (&a)[56]) {
    asm volatile("{\n"
                 ".reg .b32 t<112>;\n"
                 "mov.b64 {t0,t1}, %0;\n"
                 "mov.b64 {t2,t3}, %1;\n"
                 "mov.b64 {t4,t5}, %2;\n"
                 "mov.b64 {t6,t7}, %3;\n"
                 "mov.b64 {t8,t9}, %4;\n"
                 "mov.b64 {t10,t11}, %5;\n"
                 "mov.b64 {t12,t13}, %6;\n"
                 "mov.b64 {t14,t15}, %7;\n"
                 "mov.b64 {t16,t17}, %8;\n"
                 "mov.b64 {t18,t19}, %9;\n"
                 "mov.b64 {t20,t21}, %10;\n"
                 "mov.b64 {t22,t23}, %11;\n"
                 "mov.b64 {t24,t25}, %12;\n"
                 "mov.b64 {t26,t27}, %13;\n"
                 "mov.b64 {t28,t29}, %14;\n"
                 "mov.b64 {t30,t31}, %15;\n"
                 "mov.b64 {t32,t33}, %16;\n"
                 "mov.b64 {t34,t35}, %17;\n"
                 "mov.b64 {t36,t37}, %18;\n"
                 "mov.b64 {t38,t39}, %19;\n"
                 "mov.b64 {t40,t41}, %20;\n"
                 "mov.b64 {t42,t43}, %21;\n"
                 "mov.b64 {t44,t45}, %22;\n"
                 "mov.b64 {t46,t47}, %23;\n"
                 "mov.b64 {t48,t49}, %24;\n"
                 "mov.b64 {t50,t51}, %25;\n"
                 "mov.b64 {t52,t53}, %26;\n"
                 "mov.b64 {t54,t55}, %27;\n"
                 "mov.b64 {t56,t57}, %28;\n"
                 "mov.b64 {t58,t59}, %29;\n"
                 "mov.b64 {t60,t61}, %30;\n"
                 "mov.b64 {t62,t63}, %31;\n"
                 "mov.b64 {t64,t65}, %32;\n"
                 "mov.b64 {t66,t67}, %33;\n"
                 "mov.b64 {t68,t69}, %34;\n"
                 "mov.b64 {t70,t71}, %35;\n"
                 "mov.b64 {t72,t73}, %36;\n"
                 "mov.b64 {t74,t75}, %37;\n"
                 "mov.b64 {t76,t77}, %38;\n"
                 "mov.b64 {t78,t79}, %39;\n"
                 "mov.b64 {t80,t81}, %40;\n"
                 "mov.b64 {t82,t83}, %41;\n"
                 "mov.b64 {t84,t85}, %42;\n"
                 "mov.b64 {t86,t87}, %43;\n"
                 "mov.b64 {t88,t89}, %44;\n"
                 "mov.b64 {t90,t91}, %45;\n"
                 "mov.b64 {t92,t93}, %46;\n"
                 "mov.b64 {t94,t95}, %47;\n"
                 "mov.b64 {t96,t97}, %48;\n"
                 "mov.b64 {t98,t99}, %49;\n"
                 "mov.b64 {t100,t101}, %50;\n"
                 "mov.b64 {t102,t103}, %51;\n"
                 "mov.b64 {t104,t105}, %52;\n"
                 "mov.b64 {t106,t107}, %53;\n"
                 "mov.b64 {t108,t109}, %54;\n"
                 "mov.b64 {t110,t111}, %55;\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%56], {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%57], {t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%58], {t32,t33,t34,t35,t36,t37,t38,t39,t40,t41,t42,t43,t44,t45,t46,t47};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%59], {t48,t49,t50,t51,t52,t53,t54,t55,t56,t57,t58,t59,t60,t61,t62,t63};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%60], {t64,t65,t66,t67,t68,t69,t70,t71,t72,t73,t74,t75,t76,t77,t78,t79};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%61], {t80,t81,t82,t83,t84,t85,t86,t87,t88,t89,t90,t91,t92,t93,t94,t95};\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%62], {t96,t97,t98,t99,t100,t101,t102,t103,t104,t105,t106,t107,t108,t109,t110,t111};\n"
                 "tcgen05.wait::st.sync.aligned;\n"
                 "}\n"
                 :
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(a[8]), "d"(a[9]), "d"(a[10]), "d"(a[11]), "d"(a[12]), "d"(a[13]), "d"(a[14]), "d"(a[15]), "d"(a[16]), "d"(a[17]), "d"(a[18]), "d"(a[19]), "d"(a[20]), "d"(a[21]), "d"(a[22]), "d"(a[23]), "d"(a[24]), "d"(a[25]), "d"(a[26]), "d"(a[27]), "d"(a[28]), "d"(a[29]), "d"(a[30]), "d"(a[31]), "d"(a[32]), "d"(a[33]), "d"(a[34]), "d"(a[35]), "d"(a[36]), "d"(a[37]), "d"(a[38]), "d"(a[39]), "d"(a[40]), "d"(a[41]), "d"(a[42]), "d"(a[43]), "d"(a[44]), "d"(a[45]), "d"(a[46]), "d"(a[47]), "d"(a[48]), "d"(a[49]), "d"(a[50]), "d"(a[51]), "d"(a[52]), "d"(a[53]), "d"(a[54]), "d"(a[55]), "r"(taddr + 0), "r"(taddr + 16), "r"(taddr + 32), "r"(taddr + 48), "r"(taddr + 64), "r"(taddr + 80), "r"(taddr + 96)
                 : "memory");
}
__device__ __forceinline__ void tmem_unpark(unsigned taddr, double (&a)[56]) {
    asm volatile("{\n"
                 ".reg .b32 t<112>;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}, [%56];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31}, [%57];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t32,t33,t34,t35,t36,t37,t38,t39,t40,t41,t42,t43,t44,t45,t46,t47}, [%58];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t48,t49,t50,t51,t52,t53,t54,t55,t56,t57,t58,t59,t60,t61,t62,t63}, [%59];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t64,t65,t66,t67,t68,t69,t70,t71,t72,t73,t74,t75,t76,t77,t78,t79}, [%60];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t80,t81,t82,t83,t84,t85,t86,t87,t88,t89,t90,t91,t92,t93,t94,t95}, [%61];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t96,t97,t98,t99,t100,t101,t102,t103,t104,t105,t106,t107,t108,t109,t110,t111}, [%62];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 "mov.b64 %0, {t0,t1};\n"
                 "mov.b64 %1, {t2,t3};\n"
                 "mov.b64 %2, {t4,t5};\n"
                 "mov.b64 %3, {t6,t7};\n"
                 "mov.b64 %4, {t8,t9};\n"
                 "mov.b64 %5, {t10,t11};\n"
                 "mov.b64 %6, {t12,t13};\n"
                 "mov.b64 %7, {t14,t15};\n"
                 "mov.b64 %8, {t16,t17};\n"
                 "mov.b64 %9, {t18,t19};\n"
                 "mov.b64 %10, {t20,t21};\n"
                 "mov.b64 %11, {t22,t23};\n"
                 "mov.b64 %12, {t24,t25};\n"
                 "mov.b64 %13, {t26,t27};\n"
                 "mov.b64 %14, {t28,t29};\n"
                 "mov.b64 %15, {t30,t31};\n"
                 "mov.b64 %16, {t32,t33};\n"
                 "mov.b64 %17, {t34,t35};\n"
                 "mov.b64 %18, {t36,t37};\n"
                 "mov.b64 %19, {t38,t39};\n"
                 "mov.b64 %20, {t40,t41};\n"
                 "mov.b64 %21, {t42,t43};\n"
                 "mov.b64 %22, {t44,t45};\n"
                 "mov.b64 %23, {t46,t47};\n"
                 "mov.b64 %24, {t48,t49};\n"
                 "mov.b64 %25, {t50,t51};\n"
                 "mov.b64 %26, {t52,t53};\n"
                 "mov.b64 %27, {t54,t55};\n"
                 "mov.b64 %28, {t56,t57};\n"
                 "mov.b64 %29, {t58,t59};\n"
                 "mov.b64 %30, {t60,t61};\n"
                 "mov.b64 %31, {t62,t63};\n"
                 "mov.b64 %32, {t64,t65};\n"
                 "mov.b64 %33, {t66,t67};\n"
                 "mov.b64 %34, {t68,t69};\n"
                 "mov.b64 %35, {t70,t71};\n"
                 "mov.b64 %36, {t72,t73};\n"
                 "mov.b64 %37, {t74,t75};\n"
                 "mov.b64 %38, {t76,t77};\n"
                 "mov.b64 %39, {t78,t79};\n"
                 "mov.b64 %40, {t80,t81};\n"
                 "mov.b64 %41, {t82,t83};\n"
                 "mov.b64 %42, {t84,t85};\n"
                 "mov.b64 %43, {t86,t87};\n"
                 "mov.b64 %44, {t88,t89};\n"
                 "mov.b64 %45, {t90,t91};\n"
                 "mov.b64 %46, {t92,t93};\n"
                 "mov.b64 %47, {t94,t95};\n"
                 "mov.b64 %48, {t96,t97};\n"
                 "mov.b64 %49, {t98,t99};\n"
                 "mov.b64 %50, {t100,t101};\n"
                 "mov.b64 %51, {t102,t103};\n"
                 "mov.b64 %52, {t104,t105};\n"
                 "mov.b64 %53, {t106,t107};\n"
                 "mov.b64 %54, {t108,t109};\n"
                 "mov.b64 %55, {t110,t111};\n"
                 "}\n"
                 : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]), "=d"(a[4]), "=d"(a[5]), "=d"(a[6]), "=d"(a[7]), "=d"(a[8]), "=d"(a[9]), "=d"(a[10]), "=d"(a[11]), "=d"(a[12]), "=d"(a[13]), "=d"(a[14]), "=d"(a[15]), "=d"(a[16]), "=d"(a[17]), "=d"(a[18]), "=d"(a[19]), "=d"(a[20]), "=d"(a[21]), "=d"(a[22]), "=d"(a[23]), "=d"(a[24]), "=d"(a[25]), "=d"(a[26]), "=d"(a[27]), "=d"(a[28]), "=d"(a[29]), "=d"(a[30]), "=d"(a[31]), "=d"(a[32]), "=d"(a[33]), "=d"(a[34]), "=d"(a[35]), "=d"(a[36]), "=d"(a[37]), "=d"(a[38]), "=d"(a[39]), "=d"(a[40]), "=d"(a[41]), "=d"(a[42]), "=d"(a[43]), "=d"(a[44]), "=d"(a[45]), "=d"(a[46]), "=d"(a[47]), "=d"(a[48]), "=d"(a[49]), "=d"(a[50]), "=d"(a[51]), "=d"(a[52]), "=d"(a[53]), "=d"(a[54]), "=d"(a[55])
                 : "r"(taddr + 0), "r"(taddr + 16), "r"(taddr + 32), "r"(taddr + 48), "r"(taddr + 64), "r"(taddr + 80), "r"(taddr + 96)
                 : "memory");
}
}  // namespace

// mbarrier / bulk-copy wrappers on 32-bit shared addresses computed once from the dynamic shared-memory base (taking the
// address of a __shared__ variable costs an S2R SR_CgaCtaId + LEA each time on sm_100)
__device__ __forceinline__ void mb_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mb_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *gsrc, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}

// One visit of a consumer warp: 16 K-rows against the first NA row fragments of its half (NA x 4 column fragments per k-step).
template <int NA, int MH>
__device__ __forceinline__ void f3_visit(double (&acc)[56], const double *__restrict__ as, const double *__restrict__ bs) {
    double bf_ = bs[0];
#pragma unroll
    for (int ks = 0; ks < F3_CH / 4; ++ks) {
        double af[NA];
#pragma unroll
        for (int mi = 0; mi < NA; ++mi) af[mi] = as[(ks * MH + mi * 8) * 4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const double bcur = bf_;
            const int nx = ks * 4 + ni + 1;  // next (k-step, column fragment), fetched before this group's DMMAs
            if (nx < F3_CH) bf_ = bs[(nx >> 2) * 4 * F3_LDB + (nx & 3) * 8];
#pragma unroll
            for (int mi = 0; mi < NA; ++mi) dmma_f(acc[(mi * 4 + ni) * 2], acc[(mi * 4 + ni) * 2 + 1], af[mi], bcur);
        }
    }
}

template <int M, int NT>
__global__ void __launch_bounds__((F3_CW + F3_P) * 32, 1) fused3_model_chi2_kernel(FusedArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    extern __shared__ __align__(128) unsigned char fsm_raw[];
    const DevModel &m = a.m;
    constexpr int MH = M / 2;                 // rows of a half
    constexpr int MF = MH / (8 * F3_CW);      // row fragments of a consumer warp's half tile (7)
    static_assert(MH % (8 * F3_CW) == 0 && MF * 4 * 2 == 56, "a half tile is 56 accumulator doubles per lane");
    constexpr int NTS = NT > 0 ? NT : 2;
    constexpr int WSZ = F3_WIN * NTS;         // doubles per window
    const unsigned tab_bytes = 2u * m.nr * 9u * (unsigned)sizeof(double2);
    const unsigned tab_span = (tab_bytes + 127u) & ~127u;
    // dynamic shared memory: [barriers 1 KB][tables][As][Bs][red][crd][win]
    constexpr unsigned BAR_SPAN = 1024;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(fsm_raw);
    const unsigned full_a = sm0, empty_a = sm0 + 8 * F3_NSA, full_b = sm0 + 16 * F3_NSA, empty_b = full_b + 8 * F3_NSB, tab_bar = empty_b + 8 * F3_NSB;
    unsigned *tmem_base_s = reinterpret_cast<unsigned *>(fsm_raw + BAR_SPAN - 16);
    double2 *tab = reinterpret_cast<double2 *>(fsm_raw + BAR_SPAN);
    double *As = reinterpret_cast<double *>(fsm_raw + BAR_SPAN + tab_span);  // [NSA][CH/4][MH][4]
    double *Bs = As + F3_NSA * F3_CH * MH;                             // [NSB][CH][LDB]
    double *red = Bs + F3_NSB * F3_CH * F3_LDB;                        // [2][CW][BN]
    double *crd = red + 2 * F3_CW * F3_BN;                             // [2][kgrid] r_par, r_perp^2 of the grid bins
    double *win = crd + 2 * a.kgrid;                                   // [P][WS][WIN * NT] metal row windows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = a.K / F3_CH, ntiles = (a.B + F3_BN - 1) / F3_BN;
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = my_tiles * nchunks * 2;  // left-matrix stages: one per visit
    if (tid == 0) {
        for (int s = 0; s < F3_NSA; ++s) { mb_init(full_a + 8 * s, 1); mb_init(empty_a + 8 * s, F3_CW); }
        for (int s = 0; s < F3_NSB; ++s) { mb_init(full_b + 8 * s, F3_CH); mb_init(empty_b + 8 * s, F3_CW); }
        mb_init(tab_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == F3_P) {  // first consumer warp: tensor memory for the parked accumulator halves
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sm0 + BAR_SPAN - 16), "n"(F3_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    if (warp < F3_P) {
        // =============================== producers ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;\n");
        if (my_tiles == 0) return;
        const int pw = warp;
        const PointTab &pt = a.pt;
        const int ldb = a.ldb, npts = pt.npts, kgrid = a.kgrid, K = a.K;
        const bool aniso = m.flags & BF_FLAG_ANISOTROPIC, decoupled = m.flags & BF_FLAG_DECOUPLED;
        for (int i = tid; i < kgrid; i += F3_P * 32) {
            const double rt = i < npts ? pt.rperp0[i] : 0.0;
            crd[i] = i < npts ? pt.rpar0[i] : 0.0;
            crd[kgrid + i] = rt * rt;
        }
        named_bar_sync(2, F3_P * 32);
        mb_wait(tab_bar, 0);
        const double2 *tab_pk = tab, *tab_nw = tab + m.nr * 9;
        const int nr = m.nr, ny = pt.mx_ny, nyh = ny + 3;
        const double tr_r0 = m.tr_r0, tr_inv_dr = m.tr_inv_dr, dx2 = m.tr_dr * m.tr_dr, rlo = m.rlo, rhi = m.rhi;
        const double my0 = m.metal_y0, mymax = m.metal_ymax, minv_dy = m.metal_inv_dy;
        double *mywin = win + pw * (F3_WS * WSZ);
        const int full_iters = npts / F3_RPI;  // iterations whose 24 rows are all grid bins
        for (int it = 0; it < my_tiles; ++it) {
            const int n0 = ((int)blockIdx.x + it * (int)gridDim.x) * F3_BN;
            const int b = min(n0 + lane, a.B - 1);
            const double *pT = a.pT + b, *S = a.S + b;
            // ---- per-vector state (single redshift class: the evolution factors are per-vector constants) ----
            double amp_pk, amp_sm, c1, c2, dpi, s_par2, s_perp2;
            double mn[NT > 0 ? NT : 1];
            {
                FastVec fv;
                load_fast_vec(m, pT, ldb, S[(size_t)(a.zc_trigger * 3 + 1) * ldb], fv);
                const double eb = S[0], ebeta = S[(size_t)ldb], es = S[(size_t)2 * ldb];
                amp_sm = fv.biasSq * eb;
                amp_pk = amp_sm * fv.ampl;
                c1 = fv.c1; c2 = fv.c2;
                dpi = fv.dv * a.vfac[0];
                const double apar = (aniso ? fv.apar : fv.iso) * es, aperp = (aniso ? fv.aperp : fv.iso) * es;
                s_par2 = apar * apar;
                s_perp2 = aperp * aperp;
                if constexpr (NT > 0) {
#pragma unroll
                    for (int c = 0; c < NT / 3; ++c) cross_metal_norms(m, pT, ldb, c, eb, ebeta, mn[3 * c], mn[3 * c + 1], mn[3 * c + 2]);
                }
            }
            // range of the velocity shift over the panel: fixes, per bin, the window of template rows its vectors touch
            double dpi_min = dpi;
            bool staged = false;
            if constexpr (NT > 0) {
                double dpi_max = dpi;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dpi_min = fmin(dpi_min, __shfl_xor_sync(0xffffffffu, dpi_min, o));
                    dpi_max = fmax(dpi_max, __shfl_xor_sync(0xffffffffu, dpi_max, o));
                }
                // stencil rows iy-1 .. iy+2 of all 32 vectors fit one window when their cell indices differ by <= 3
                staged = (dpi_max - dpi_min) * minv_dy < 2.5 && ny >= F3_WIN;
            }
            // window of bin i into slot seq % WS of the warp's private ring (16-byte cp.async copies by all lanes; halo
            // coordinates as in the first kernel). Returns the first halo row of the window.
            auto stage_window = [&](int i, int seq) {
                const double y = fmin(fmax(crd[i] + dpi_min, my0), mymax);
                const int wlo = min(__double2int_rz((y - my0) * minv_dy), nyh - F3_WIN);
                const double *src = pt.mxh + ((size_t)i * nyh + wlo) * NTS;
                double *dst = mywin + (seq % F3_WS) * WSZ;
#pragma unroll
                for (int c = lane; c < WSZ / 2; c += 32) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + 2 * c) : "memory");
                }
                return wlo;
            };
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
                    const double y = fmin(fmax(rpar, my0), mymax);
                    const double fy = (y - my0) * minv_dy;
                    const int iy = __double2int_rz(fy);
                    double wy[4];
                    catmull_weights(fy - (double)iy, wy);
                    val += metal_patch<NT>(mbase + iy * NTS, mn, wy);
                }
                return zero ? 0.0 : val;
            };
            auto publish = [&](int k, double val) {  // row k of the panel goes into its chunk
                const int g = it * nchunks + k / F3_CH, slot = g % F3_NSB;
                if (g >= F3_NSB) mb_wait(empty_b + 8 * slot, ((g / F3_NSB) - 1) & 1);
                Bs[(slot * F3_CH + (k % F3_CH)) * F3_LDB + lane] = val;
                __syncwarp();
                if (lane == 0) mb_arrive(full_b + 8 * slot);
            };
            int r = 0;  // first row of the current iteration
            if (NT == 0 || staged) {
                int w0 = 0, w1 = 0, seq = 0;
                if constexpr (NT > 0) {
                    if (full_iters > 0) {
                        w0 = stage_window(pw, 0);
                        w1 = stage_window(F3_P + pw, 1);
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                }
#pragma unroll 1
                for (int i = 0; i < full_iters; ++i, r += F3_RPI) {
                    const int k0 = r + pw, k1 = k0 + F3_P;
                    int w2 = 0, w3 = 0;
                    if constexpr (NT > 0) {
                        if (i + 1 < full_iters) {
                            w2 = stage_window(k0 + F3_RPI, seq + 2);
                            w3 = stage_window(k1 + F3_RPI, seq + 3);
                        }
                        asm volatile("cp.async.commit_group;\n" ::: "memory");
                        asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // this iteration's windows have landed
                        __syncwarp();
                    }
                    const double *m0 = mywin + (seq % F3_WS) * WSZ - w0 * NTS;
                    const double *m1 = mywin + ((seq + 1) % F3_WS) * WSZ - w1 * NTS;
                    const double v0 = eval_row(k0, m0), v1 = eval_row(k1, m1);
                    publish(k0, v0);
                    publish(k1, v1);
                    if constexpr (NT > 0) {
                        __syncwarp();  // both windows are free again
                        w0 = w2; w1 = w3; seq += 2;
                    }
                }
                if constexpr (NT > 0) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
            // ---- remaining rows: grid bins with template rows straight from global memory (velocity shifts too far
            // apart for one window per bin, or a partial iteration), zero padding, the broadband coefficient rows ----
#pragma unroll 1
            for (; r < K; r += F3_RPI) {
#pragma unroll 1
                for (int e = 0; e < 2; ++e) {
                    const int k = r + e * F3_P + pw;
                    if (k >= K) break;
                    double val = 0.0;
                    if (k < npts) val = eval_row(k, pt.mxh + (size_t)k * nyh * NTS);
                    else if (k >= kgrid) val = a.rows[(size_t)(k - kgrid) * ldb + b];
                    publish(k, val);
                }
            }
        }
        return;
    }

    // =============================== consumers ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;\n");
    const int l4 = lane & 3, l8 = lane >> 2, cw = warp - F3_P;
    const unsigned tmem0 = *tmem_base_s + ((unsigned)(32 * cw) << 16);  // this warp's lanes; region h at column 112 h
    if (my_tiles > 0) {
        constexpr unsigned stage_bytes = (unsigned)(F3_CH * MH * sizeof(double));
        const bool loader = cw == 0 && lane == 0;
        const unsigned as_u32 = sm0 + BAR_SPAN + tab_span;
        const int per_tile = nchunks * 2;
        auto load_stage = [&](int g) {
            const int slot = g % F3_NSA;
            mb_expect_tx(full_a + 8 * slot, stage_bytes);
            bulk_g2s(as_u32 + slot * stage_bytes, a.At3 + (size_t)(g % per_tile) * F3_CH * MH, stage_bytes, full_a + 8 * slot);
        };
        if (loader) {
            mb_expect_tx(tab_bar, tab_bytes);
            for (unsigned off = 0; off < tab_bytes; off += 32768u)
                bulk_g2s(sm0 + BAR_SPAN + off, reinterpret_cast<const unsigned char *>(a.tabs) + off, min(32768u, tab_bytes - off), tab_bar);
            for (int g = 0; g < F3_NSA && g < total; ++g) load_stage(g);
        }
        const double *as0 = As + (cw * (MF * 8) + l8) * 4 + l4;
        const double *bs0 = Bs + l4 * F3_LDB + l8;
        for (int it = 0; it < my_tiles; ++it) {
            const int n0 = ((int)blockIdx.x + it * (int)gridDim.x) * F3_BN;
            double acc[56];  // [mi][ni][e] of the half being worked on
#pragma unroll
            for (int i = 0; i < 56; ++i) acc[i] = 0.0;
            int cur = 0;  // the half in registers
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                const int gb = it * nchunks + c, sb = gb % F3_NSB;
                mb_wait(full_b + 8 * sb, (gb / F3_NSB) & 1);
                const double *bs = bs0 + sb * (F3_CH * F3_LDB);
#pragma unroll 1
                for (int visit = 0; visit < 2; ++visit) {
                    const int g = gb * 2 + visit, sa = g % F3_NSA;
                    mb_wait(full_a + 8 * sa, (g / F3_NSA) & 1);
                    const double *as = as0 + sa * (F3_CH * MH);
                    if (a.debug != 2) {
                        // trapezoidal left matrix: slot mi of this warp and half holds fragment f = 8 mi + 4 half + warp, whose rows
                        // start at column 8 f: the visit runs the variant with as many slots as have started by its last column
                        // (one jump per visit; predicated-off DMMAs would still cost their 16 issue cycles each)
                        const int half_now = (c + visit) & 1, frag0 = 32 * half_now + 8 * cw, kend = c * F3_CH + F3_CH - 1;
                        const int nact = a.tri ? min(MF, kend >= frag0 ? (kend - frag0) / 64 + 1 : 0) : MF;
                        switch (nact) {
                        case 7: f3_visit<7, MH>(acc, as, bs); break;
                        case 6: f3_visit<6, MH>(acc, as, bs); break;
                        case 5: f3_visit<5, MH>(acc, as, bs); break;
                        case 4: f3_visit<4, MH>(acc, as, bs); break;
                        case 3: f3_visit<3, MH>(acc, as, bs); break;
                        case 2: f3_visit<2, MH>(acc, as, bs); break;
                        case 1: f3_visit<1, MH>(acc, as, bs); break;
                        default: break;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mb_arrive(empty_a + 8 * sa);
                    // the left-matrix stage released one visit ago is refilled now: by this time the other consumer
                    // warps have normally released it too, so the loader does not wait
                    if (loader && g >= 1 && g - 1 + F3_NSA < total) {
                        mb_wait(empty_a + 8 * ((g - 1) % F3_NSA), ((g - 1) / F3_NSA) & 1);
                        load_stage(g - 1 + F3_NSA);
                    }
                    if (visit == 0) {  // swap halves: park the one just worked on, fetch the other
                        tmem_park(tmem0 + 112u * cur, acc);
                        if (c == 0) {
#pragma unroll
                            for (int i = 0; i < 56; ++i) acc[i] = 0.0;
                        } else
                            tmem_unpark(tmem0 + 112u * (cur ^ 1), acc);
                        cur ^= 1;
                    }
                }
                __syncwarp();
                if (lane == 0) mb_arrive(empty_b + 8 * sb);
            }
            // column sums of squares over this warp's rows (both halves), then over the consumer warps
            double ss[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ss[q] = 0.0;
#pragma unroll
            for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                for (int q = 0; q < 8; ++q) ss[q] = fma(acc[mi * 8 + q], acc[mi * 8 + q], ss[q]);
            tmem_unpark(tmem0 + 112u * (cur ^ 1), acc);
#pragma unroll
            for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                for (int q = 0; q < 8; ++q) ss[q] = fma(acc[mi * 8 + q], acc[mi * 8 + q], ss[q]);
            double *rbuf = red + (it & 1) * (F3_CW * F3_BN);
#pragma unroll
            for (int q = 0; q < 8; ++q) {  // q = ni * 2 + e
                double s = ss[q];
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (lane < 4) rbuf[cw * F3_BN + (q >> 1) * 8 + 2 * lane + (q & 1)] = s;
            }
            named_bar_sync(1, F3_CW * 32);
            if (cw == 0) {
                const int n = n0 + lane;
                if (n < a.B) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < F3_CW; ++w) s += rbuf[w * F3_BN + lane];
                    if (a.status && a.status[n]) s = nan("");
                    a.chi2[n] = s;
                }
            }
        }
    }
    // all tensor-memory traffic of the consumer warps is complete (tcgen05.wait inside park / unpark)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    named_bar_sync(1, F3_CW * 32);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (cw == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(*tmem_base_s), "n"(F3_TMEM_COLS) : "memory");
}

// left matrix in the order the third kernel streams it: [chunk c][visit v][k-step][row of half (c + v) & 1][4]
void gemm_fused_tile_A3(const double *A, int M, int K, int lda, double *At) {
    const int MH = M / 2;
    size_t o = 0;
    for (int c = 0; c < K / F3_CH; ++c)
        for (int v = 0; v < 2; ++v) {
            const int half = (c + v) & 1;
            for (int ks = 0; ks < F3_CH / 4; ++ks)
                for (int r = 0; r < MH; ++r)
                    for (int k = 0; k < 4; ++k) At[o++] = A[(size_t)(half * MH + r) * lda + c * F3_CH + ks * 4 + k];
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
    static const bool gen1 = std::getenv("BAOFIT_FUSED_GEN1") != nullptr;  // A/B measurements: the 32-column kernel
    a.debug = dbg;
    constexpr int M = 448;
    const unsigned tab_bytes = 2u * a.m.nr * 9u * (unsigned)sizeof(double2), tab_span = (tab_bytes + 127u) & ~127u;
    const int nt = a.m.metal_kind == BF_METAL_CROSS_BICUBIC ? 12 : 2;
    if (!gen1 && a.At3 && a.K % F3_CH == 0) {
        const size_t smem = 1024 + tab_span + ((size_t)F3_NSA * F3_CH * (M / 2) + (size_t)F3_NSB * F3_CH * F3_LDB + (size_t)2 * F3_CW * F3_BN + (size_t)2 * a.kgrid +
                                               (size_t)F3_P * F3_WS * F3_WIN * nt) * sizeof(double);
        const int ntiles = (a.B + F3_BN - 1) / F3_BN, grid = std::min(ntiles, sm_count), threads = (F3_CW + F3_P) * 32;
        static SmemOptIn o0, o12;
        if (a.m.metal_kind == BF_METAL_CROSS_BICUBIC) {
            o12.ensure(fused3_model_chi2_kernel<M, 12>, smem);
            Launch(grid, threads, smem, s)(fused3_model_chi2_kernel<M, 12>, a);
        } else {
            o0.ensure(fused3_model_chi2_kernel<M, 0>, smem);
            Launch(grid, threads, smem, s)(fused3_model_chi2_kernel<M, 0>, a);
        }
        return;
    }
    const size_t smem = tab_span + ((size_t)FU_NS * FU_KC * M + (size_t)FU_NS * FU_KC * FU_LDB + (size_t)2 * FU_CW * FU_BN + (size_t)2 * a.kgrid +
                                    (size_t)FU_P * FU_WS * FU_WIN * nt) * sizeof(double);
    const int ntiles = (a.B + FU_BN - 1) / FU_BN, grid = std::min(ntiles, sm_count), threads = (FU_CW + FU_P) * 32;
    static SmemOptIn opt0, opt12;
    if (a.m.metal_kind == BF_METAL_CROSS_BICUBIC) {
        opt12.ensure(fused_model_chi2_kernel<M, 12>, smem);
        Launch(grid, threads, smem, s)(fused_model_chi2_kernel<M, 12>, a);
    } else {
        opt0.ensure(fused_model_chi2_kernel<M, 0>, smem);
        Launch(grid, threads, smem, s)(fused_model_chi2_kernel<M, 0>, a);
    }
}

}  // namespace bf
