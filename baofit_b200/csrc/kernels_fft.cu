// 3-D FFT k-space model (baofit/BaoKSpaceFftCorrelationModel.cc) without a 3-D FFT.
//
// The reference fills an nx x ny x nz grid with D(k, mu_k) P(k) (y = line of sight), runs an
// inverse FFT (cosmo::DistortedPowerCorrelationFft::transform, call sites :204,:212) and reads xi
// off the (r_perp, r_par) plane through the origin (:241-243). Only that plane is ever used, and
// D P is even in every k component, so
//
//   xi(ix, iy) = norm  sum_my w_my cos(2 pi my iy / ny)  sum_mx w_mx cos(2 pi mx ix / nx)  G(mx, my),
//   G(mx, my)  = sum_mz w_mz  D P (k = |(kx,ky,kz)|, mu = ky / k)
//
// with w = 1 at m = 0 and at the Nyquist index, 2 in between: the z transform degenerates to a sum
// and the other two are small dense cosine transforms. D factorises (:120-153) as
//
//   D = (1 + c1 mu^2 + c2 mu^4) . cont(|k_y|; kc, pc) . K(k, mu)
//
// where c1, c2 depend only on beta (and beta2), cont only on k_y, and K (non-linear broadening,
// non-linear correction, P itself) only on parameters that a fit keeps fixed. Hence
//   * per key (SigmaNL, 1+f, nlcorr, z_eff):  H_j(mx,my) = sum_mz w K P mu^2j  and its x transform
//     T_j[my][ix]  (fft_hsum_kernel, fft_xtrans_kernel; cached across calls),
//   * per parameter vector: one (npy x my) . (my x npx) product per component,
//       plane_b[iy][ix] = sum_my Cy[iy][my] . cont_b[my] (T_0 + c1_b T_1 + c2_b T_2)[my][ix],
//     on the FP64 tensor pipe (fft_plane_kernel, DMMA m8n8k4), the right operand built on the fly
//     in shared memory,
//   * per (bin, vector): two bicubic look-ups on the planes (fft_eval_kernel).
#include <algorithm>
#include <cstdlib>

#include "device_math.cuh"
#include "kernels.h"

namespace bf {

namespace {

// ---- tabulated P(k): cubic spline in ln k inside the table, power law outside ---------------
__device__ __forceinline__ double fft_table_power(const DevFft &f, int which, double k) {
    if (k < f.pk_k0) return f.pk_end[which][0] * pow(k / f.pk_k0, f.pk_slope[which][0]);
    if (k > f.pk_k1) return f.pk_end[which][1] * pow(k / f.pk_k1, f.pk_slope[which][1]);
    const double lk = log(k);
    int lo = 0, hi = f.n_pk - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (f.lnk[mid] > lk) hi = mid; else lo = mid;
    }
    const double dx = lk - f.lnk[lo];
    const double4 c = *reinterpret_cast<const double4 *>(f.pk_coef + ((size_t)which * (f.n_pk - 1) + lo) * 4);
    return c.x + dx * (c.y + dx * (c.z + dx * c.w));
}

// NonLinearCorrectionModel::_evaluateKSpace, baofit/NonLinearCorrectionModel.cc:49-86
__device__ __forceinline__ double fft_nl_correction(unsigned flags, const FftKeyState &ks, double k, double mu_k,
                                                    double pk) {
    if (flags & (BF_FLAG_NL_CORRECTION | BF_FLAG_FIT_NL_CORRECTION)) {
        pk = pk * ks.nl_pk_scale;
        double deltaSq = k * k * k * pk / (2.0 * M_PI * M_PI);
        double growth = ks.nl_tab[0] * deltaSq;
        double pec = pow(k / ks.nl_tab[1], ks.nl_tab[2]) * pow(fabs(mu_k), ks.nl_tab[3]);
        double pressure = (k / ks.nl_tab[4]) * (k / ks.nl_tab[4]);
        return exp(growth * (1.0 - pec) - pressure);
    }
    if (flags & BF_FLAG_NL_CORRECTION_ALT) {
        const double knl = 6.4, pnl = 0.569, kpp = 15.3, pp = 2.01, kv0 = 1.22, pv = 1.5, kvi = 0.923, pvi = 0.451;
        double kvel = kv0 * pow(1.0 + k / kvi, pvi);
        double growth = pow(k / knl, pnl), pressure = pow(k / kpp, pp);
        double pec = pow(fabs(k * mu_k) / kvel, pv);
        return exp(growth - pressure - pec);
    }
    return 1.0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// per key: H[comp][j][my][mx] = sum_mz w_mz K(k,mu) P_comp(k) mu^2j
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fft_hsum_kernel(DevModel m, FftKeyState ks, double *__restrict__ H) {
    pdl_sync();
    const DevFft &f = m.fft;
    const int mx = blockIdx.x * blockDim.x + threadIdx.x, my = blockIdx.y, comp = blockIdx.z;
    if (mx >= f.mx) return;
    const double two_pi = 2.0 * M_PI;
    const double kx = mx * (two_pi / (f.nx * f.delta)), ky = my * (two_pi / (f.ny * f.delta));
    const double dkz = two_pi / (f.nz * f.delta);
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    // :207-210 the no-wiggles component is broadened only with nl-broadband
    const bool broaden = comp == 0 || (m.flags & BF_FLAG_NL_BROADBAND);
    const double sperp2 = broaden ? ks.snl_perp2 : 0.0, spar2 = broaden ? ks.snl_par2 : 0.0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int mz = 0; mz < f.mz; ++mz) {
        const double kz = mz * dkz;
        const double k = sqrt(kx * kx + ky * ky + kz * kz);
        if (k == 0.0) continue;
        const double mu = ky / k, mu2 = mu * mu;
        const double pn = fft_table_power(f, 1, k);
        const double pk = comp == 0 ? fft_table_power(f, 0, k) - pn : pn;
        const double snl2 = spar2 * mu2 + sperp2 * (1.0 - mu2);
        const double nonlinear = exp(-0.5 * snl2 * k * k);
        double nlc = fft_nl_correction(m.flags, ks, k, mu, pk);
        if (cross) nlc = sqrt(nlc);
        const double w = (mz == 0 || 2 * mz == f.nz) ? 1.0 : 2.0;
        const double v = w * nonlinear * nlc * pk;
        a0 += v;
        a1 = fma(v, mu2, a1);
        a2 = fma(v, mu2 * mu2, a2);
    }
    const size_t plane = (size_t)f.my * f.mx, o = (size_t)my * f.mx + mx;
    H[(size_t)(comp * 3 + 0) * plane + o] = a0;
    H[(size_t)(comp * 3 + 1) * plane + o] = a1;
    H[(size_t)(comp * 3 + 2) * plane + o] = a2;
}

// T[comp*3+j][my][ix] = norm w_my sum_mx w_mx cos(2 pi mx ix / nx) H[comp*3+j][my][mx], stored in the stage
// layout of fft_plane_kernel: [comp][ctile][chunk][j][kr][kFftLdT]; rows my >= My and columns >= npx are zero
__global__ void __launch_bounds__(128) fft_xtrans_kernel(DevModel m, const double *__restrict__ H, double *__restrict__ T) {
    pdl_sync();
    const DevFft &f = m.fft;
    const int ixl = threadIdx.x, ctile = blockIdx.x, my = blockIdx.y, t = blockIdx.z;
    if (ixl >= kFftLdT) return;
    const int ix = ctile * f.ncf * 8 + ixl;
    double s = 0.0;
    if (my < f.my && ix < f.npx && ixl < f.ncf * 8) {
        const double *h = H + ((size_t)t * f.my + my) * f.mx;
        for (int mx = 0; mx < f.mx; ++mx) {
            const double w = (mx == 0 || 2 * mx == f.nx) ? 1.0 : 2.0;
            s = fma(w * f.cosx[(int)(((long long)mx * ix) % f.nx)], h[mx], s);
        }
        s *= f.norm * ((my == 0 || 2 * my == f.ny) ? 1.0 : 2.0);
    }
    const int comp = t / 3, j = t - comp * 3, chunk = my / kFftKC, kr = my - chunk * kFftKC, nchunks = f.my_pad / kFftKC;
    T[((((size_t)(comp * f.ctiles + ctile) * nchunks + chunk) * 3 + j) * kFftKC + kr) * kFftLdT + ixl] = s;
}

void launch_fft_tables(cudaStream_t s, const DevModel &m, const FftKeyState &k, double *H, double *T) {
    const DevFft &f = m.fft;
    Launch(dim3((f.mx + 127) / 128, f.my, 2), 128, 0, s)(fft_hsum_kernel, m, k, H);
    Launch(dim3(f.ctiles, f.my_pad, 6), 128, 0, s)(fft_xtrans_kernel, m, H, T);
}

// ------------------------------------------------------------------------------------------
// per vector: key check, tracer coefficients, continuum-fitting distortion along k_par
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double fft_zcorr(const DevFft &f, double r, double mu, double z) {
    // BaoKSpaceFftCorrelationModel.cc:165-168
    if (f.zcorr0 > 0.0) {
        double rpar = fabs(r * mu) / 100.;
        z = f.zcorr0 + f.zcorr1 * rpar + f.zcorr2 * rpar * rpar;
    }
    return z;
}

__device__ __forceinline__ double fft_dpi(const DevModel &m, double dv, double z) {
    // AbsCorrelationModel.cc:146-149
    double zp1 = 1.0 + z;
    return (dv / 100.) * zp1 / sqrt(1.0 - m.omega_matter + m.omega_matter * zp1 * zp1 * zp1);
}

__device__ __forceinline__ double fft_trigger_z(const FftVecArgs &a, const double *__restrict__ pT) {
    if (a.use_ztrig) return a.ztrig;
    double r = a.r0, mu = a.mu0;
    if (a.m.idx_delta_v >= 0) velocity_shift(fft_dpi(a.m, pT[(size_t)a.m.idx_delta_v * a.ldb], a.z0), r, mu);
    return fft_zcorr(a.m.fft, r, mu, a.z0);
}

constexpr int kFftVecRows = 16;  // k_par rows per thread of fft_vec_kernel

__global__ void __launch_bounds__(128) fft_vec_kernel(FftVecArgs a) {
    pdl_sync();
    const DevModel &m = a.m;
    const DevFft &f = m.fft;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const double *pT = a.pT + b;
    const int ldb = a.ldb;
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    if (blockIdx.y == 0) {
        const double ztrig = fft_trigger_z(a, pT);
        // key: everything the per-key tables read
        double key[kFftKeyMax - 1], key0[kFftKeyMax - 1];
        int n = 0;
        const double *p0 = a.pT;
        key[n] = pT[(size_t)m.idx_nl * ldb]; key0[n++] = p0[(size_t)m.idx_nl * ldb];
        key[n] = pT[(size_t)(m.idx_nl + 1) * ldb]; key0[n++] = p0[(size_t)(m.idx_nl + 1) * ldb];
        if (m.flags & BF_FLAG_FIT_NL_CORRECTION) {
            key[n] = pT[(size_t)m.idx_nlcorr * ldb]; key0[n++] = p0[(size_t)m.idx_nlcorr * ldb];
            key[n] = pT[(size_t)(m.idx_nlcorr + 1) * ldb]; key0[n++] = p0[(size_t)(m.idx_nlcorr + 1) * ldb];
        } else {
            key[n] = 0.0; key0[n++] = 0.0;
            key[n] = 0.0; key0[n++] = 0.0;
        }
        key[n] = ztrig; key0[n++] = fft_trigger_z(a, p0);
        bool differs = false;
        for (int q = 0; q < n; ++q) differs |= key[q] != key0[q];
        if (differs) a.keyflag[0] = 1.0;
        if (b == 0)
            for (int q = 0; q < n; ++q) a.keyflag[1 + q] = key0[q];
        for (int q = 0; q < n; ++q) a.keys[(size_t)q * ldb + b] = key[q];
        // tracer factor (1 + beta_z mu^2)(1 + beta2_z mu^2), :122-125 with :171-173
        const double gbeta = pT[(size_t)m.idx_gamma_beta * ldb];
        const double zf = (1.0 + ztrig) / (1.0 + m.zref);
        const double ev = gbeta == 0.0 ? 1.0 : pow(zf, gbeta);
        const double betaz = pT[(size_t)m.idx_beta * ldb] * ev;
        double c1, c2;
        if (cross) {
            const double beta2z = pT[(size_t)m.idx_bb2 * ldb] / pT[(size_t)m.idx_bias2 * ldb] * ev;
            c1 = betaz + beta2z;
            c2 = betaz * beta2z;
        } else {
            c1 = 2.0 * betaz;
            c2 = betaz * betaz;
        }
        a.c12[b] = c1;
        a.c12[(size_t)ldb + b] = c2;
    }
    // continuum-fitting distortion, :128-143 (a function of k_par = |k mu_k| = |k_y| only)
    const double kc = pT[(size_t)m.idx_cont * ldb], pc = pT[(size_t)(m.idx_cont + 1) * ldb];
    const double dky = 2.0 * M_PI / (f.ny * f.delta);
    const int r0 = blockIdx.y * kFftVecRows;
    for (int my = r0; my < r0 + kFftVecRows && my < f.my_pad; ++my) {
        double c = 0.0;
        if (my < f.my) {
            const double kpar = my * dky;
            if (m.flags & BF_FLAG_NO_DISTORTION) c = 1.0;
            else if (m.flags & BF_FLAG_DISTORTION_ALT) c = tanh(pow(kpar / kc, pc));
            else {
                const double k1 = pow(kpar / kc + 1.0, 0.75);
                c = pow((k1 - 1.0 / k1) / (k1 + 1.0 / k1), pc);
            }
            if (cross) c = sqrt(c);
        }
        a.cont[(size_t)my * ldb + b] = c;
    }
}

void launch_fft_vec(cudaStream_t s, const FftVecArgs &a) {
    dim3 grid((a.B + 127) / 128, (a.m.fft.my_pad + kFftVecRows - 1) / kFftVecRows);
    Launch(grid, 128, 0, s)(fft_vec_kernel, a);
}

// ------------------------------------------------------------------------------------------
// per vector: the y transform as a DMMA product, right operand built in shared memory
// ------------------------------------------------------------------------------------------
namespace {
constexpr int PV = 4;        // parameter vectors per CTA
constexpr int PKC = kFftKC;  // k depth (k_par rows) per pipeline stage
constexpr int PSTAGES = 5;   // at most 5 TMA stages of the Cy / T_j tiles (fewer if k_par tables crowd shared memory)
constexpr int PLDA = kFftLdA, PLDT = kFftLdT, P_A = kFftStageA, P_T = kFftStageT;

__device__ __forceinline__ void p_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void p_mbar_init(unsigned long long *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void p_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_arrive(unsigned long long *bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void p_mbar_wait(unsigned long long *bar, unsigned parity) {
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
// TMA bulk copy global -> shared, completion counted on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void p_tma_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
}  // namespace

struct FftPlaneArgs {
    DevFft f;
    const double *T, *c12, *cont;
    int ldb;
    const int *cols;  // columns of the vectors to transform (null: 0..count-1)
    int count;
    int nstages;  // pipeline depth actually used, 2..PSTAGES
    double *planes;
};

// CTA = PV vectors x one component x one (<= 72 x 72) tile of the plane; warp w works on vector
// w & 3 and the column group w >> 2 (three 8-column fragments) over all row fragments of the tile:
// 9 x 3 accumulator fragments, 9 A + 9 T fragment loads per 27 DMMAs.
//   * The per-vector operands are never materialised: the continuum-fitting factor cont_v[my] scales
//     the A fragments (a column scaling of Cy) and each B fragment T0 + c1_v T1 + c2_v T2 is combined
//     in registers from the three static tiles (2 DFMA per B fragment feeding 9 DMMAs).
//   * Cy and T_j are stored in global memory in exactly the padded stage layout, so one stage is
//     two contiguous TMA bulk copies (cp.async.bulk, SASS UBLKCP) issued by one thread; "full"
//     mbarriers carry the byte counts, "empty" mbarriers count the consumer warps. No CTA-wide
//     barrier inside the k loop: warps drift apart by up to PSTAGES-1 chunks, so the FP64 tensor
//     pipe never drains at chunk boundaries.
// NCF (column fragments per tile) is a template parameter: all tile index arithmetic is constant.
template <int NCF>
__global__ void __launch_bounds__(PV * 3 * 32, 1) fft_plane_kernel(FftPlaneArgs a) {
    pdl_sync();
    extern __shared__ __align__(128) double psm[];
    const DevFft &f = a.f;
    constexpr int NCOLS = NCF * 8, NRF = 9;
    const int NS = a.nstages;
    double *As = psm;               // [NS][72][PLDA]
    double *Ts = As + NS * P_A;     // [NS][3][PKC][PLDT]
    double *conts = Ts + NS * P_T;  // [2 item parities][PV][my_pad]
    __shared__ double c12s[2][2][PV];
    __shared__ __align__(8) unsigned long long full_bar[PSTAGES], empty_bar[PSTAGES];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int l4 = lane & 3, l8 = lane >> 2;
    const int wv = warp & (PV - 1), cf0 = (warp >> 2) * 3;  // vector, first column fragment
    const int ncf = min(3, NCF - cf0);                       // column fragments of this warp
    const int rtile = blockIdx.y, ctile = blockIdx.z;
    const int row0 = rtile * f.wr * 8, col0 = ctile * NCOLS;
    const int nrows = min(f.wr * 8, f.npy_pad - row0);
    const int nrf = nrows >> 3;
    const int ncols_valid = min(NCOLS, f.npx_pad - col0);
    const int nchunks = f.my_pad / PKC;
    const double *gA = f.Cy + (size_t)rtile * nchunks * P_A;
    constexpr unsigned kStageBytes = (unsigned)((P_A + P_T) * sizeof(double));

    // Persistent CTA: work items (vector group, component) i = blockIdx.x, blockIdx.x + gridDim.x, ...
    // The stage pipeline runs across item boundaries (global chunk counter gc), so the prologue,
    // the epilogue stores and the TMA latency of one item hide behind its neighbours.
    const bool full_tile = nrf == NRF && ncf == 3;  // warp-uniform: the common case runs unpredicated
    const int ngroups = (a.count + PV - 1) / PV, nitems = 2 * ngroups;
    const int my_items = blockIdx.x < nitems ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total_chunks = my_items * nchunks;
    auto issue = [&](int gc) {  // one thread
        const int slot = gc % NS, it = gc / nchunks, chunk = gc - it * nchunks;
        const int comp = (blockIdx.x + it * gridDim.x) & 1;
        const double *gT = a.T + (size_t)(comp * f.ctiles + ctile) * nchunks * P_T;
        p_mbar_expect_tx(&full_bar[slot], kStageBytes);
        p_tma_g2s(As + slot * P_A, gA + (size_t)chunk * P_A, (unsigned)(P_A * sizeof(double)), &full_bar[slot]);
        p_tma_g2s(Ts + slot * P_T, gT + (size_t)chunk * P_T, (unsigned)(P_T * sizeof(double)), &full_bar[slot]);
    };
    auto item_col = [&](int it, int v) {  // column of vector slot v of item it (clamped; extra slots are not stored)
        const int i = min(((blockIdx.x + it * gridDim.x) >> 1) * PV + v, a.count - 1);
        return a.cols ? a.cols[i] : i;
    };

    if (tid == 0) {
        for (int st = 0; st < NS; ++st) { p_mbar_init(&full_bar[st], 1); p_mbar_init(&empty_bar[st], nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (my_items == 0) return;
    if (tid == 0)
        for (int c = 0; c < NS && c < total_chunks; ++c) issue(c);
    // per-vector scalars of the first item; later items are prefetched one item ahead
    constexpr int kContRegs = 3;  // registers per thread for that prefetch; larger k_par tables are copied in place
    const bool pre_ok = PV * f.my_pad <= kContRegs * nthr;
    for (int i = tid; i < PV * f.my_pad; i += nthr) {
        int v = i / f.my_pad, my = i - v * f.my_pad;
        conts[i] = a.cont[(size_t)my * a.ldb + item_col(0, v)];
    }
    if (tid < 2 * PV) c12s[0][tid / PV][tid % PV] = a.c12[(size_t)(tid / PV) * a.ldb + item_col(0, tid % PV)];

    for (int it = 0; it < my_items; ++it) {
        const int par = it & 1;
        __syncthreads();  // scalars of this item are in place; everyone has left item it-1
        const int item = blockIdx.x + it * gridDim.x, b0 = (item >> 1) * PV, comp = item & 1;
        const double c1 = c12s[par][0][wv], c2 = c12s[par][1][wv];
        const int mycol = item_col(it, wv);
        // prefetch the scalars of the next item into registers; stored after this item's k loop
        double cpre[kContRegs], c12pre = 0.0;
        const bool have_next = it + 1 < my_items;
        if (have_next && pre_ok) {
#pragma unroll
            for (int q = 0; q < kContRegs; ++q) {
                const int i = tid + q * nthr;
                if (i < PV * f.my_pad) {
                    int v = i / f.my_pad, my = i - v * f.my_pad;
                    cpre[q] = a.cont[(size_t)my * a.ldb + item_col(it + 1, v)];
                }
            }
            if (tid < 2 * PV) c12pre = a.c12[(size_t)(tid / PV) * a.ldb + item_col(it + 1, tid % PV)];
        } else if (have_next && tid < 2 * PV)
            c12pre = a.c12[(size_t)(tid / PV) * a.ldb + item_col(it + 1, tid % PV)];

        double acc[NRF][3][2];
#pragma unroll
        for (int mi = 0; mi < NRF; ++mi)
#pragma unroll
            for (int ni = 0; ni < 3; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int t = 0; t < nchunks; ++t) {
            const int gc = it * nchunks + t, slot = gc % NS;
            if (tid == 0 && gc >= 1 && gc - 1 + NS < total_chunks) {
                // refill the stage of chunk gc-1 once every warp has released it
                p_mbar_wait(&empty_bar[(gc - 1) % NS], ((gc - 1) / NS) & 1);
                issue(gc - 1 + NS);
            }
            __syncwarp();
            p_mbar_wait(&full_bar[slot], (gc / NS) & 1);
            if (ncf > 0) {
                const double *as = As + slot * P_A + l8 * PLDA + l4;
                const double *ts = Ts + slot * P_T + l4 * PLDT + cf0 * 8 + l8;
                const double *cs = conts + (par * PV + wv) * f.my_pad + t * PKC + l4;
                if (full_tile) {
#pragma unroll
                    for (int kk = 0; kk < PKC; kk += 4) {
                        const double cv = cs[kk];
                        double bf[3];
#pragma unroll
                        for (int ni = 0; ni < 3; ++ni) {
                            const double *tp = ts + kk * PLDT + ni * 8;
                            bf[ni] = fma(c2, tp[2 * PKC * PLDT], fma(c1, tp[PKC * PLDT], tp[0]));
                        }
#pragma unroll
                        for (int mi = 0; mi < NRF; ++mi) {
                            const double af = as[mi * 8 * PLDA + kk] * cv;
#pragma unroll
                            for (int ni = 0; ni < 3; ++ni) p_dmma(acc[mi][ni][0], acc[mi][ni][1], af, bf[ni]);
                        }
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < PKC; kk += 4) {
                        const double cv = cs[kk];
                        double bf[3];
#pragma unroll
                        for (int ni = 0; ni < 3; ++ni) {
                            const double *tp = ts + kk * PLDT + (ni < ncf ? ni * 8 : 0);
                            bf[ni] = fma(c2, tp[2 * PKC * PLDT], fma(c1, tp[PKC * PLDT], tp[0]));
                        }
#pragma unroll
                        for (int mi = 0; mi < NRF; ++mi) {
                            if (mi < nrf) {
                                const double af = as[mi * 8 * PLDA + kk] * cv;
#pragma unroll
                                for (int ni = 0; ni < 3; ++ni)
                                    if (ni < ncf) p_dmma(acc[mi][ni][0], acc[mi][ni][1], af, bf[ni]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) p_mbar_arrive(&empty_bar[slot]);
        }
        if (have_next) {
            if (pre_ok) {
#pragma unroll
                for (int q = 0; q < kContRegs; ++q) {
                    const int i = tid + q * nthr;
                    if (i < PV * f.my_pad) conts[(par ^ 1) * PV * f.my_pad + i] = cpre[q];
                }
            } else {
                for (int i = tid; i < PV * f.my_pad; i += nthr) {
                    int v = i / f.my_pad, my = i - v * f.my_pad;
                    conts[(par ^ 1) * PV * f.my_pad + i] = a.cont[(size_t)my * a.ldb + item_col(it + 1, v)];
                }
            }
            if (tid < 2 * PV) c12s[par ^ 1][tid / PV][tid % PV] = c12pre;
        }
        if (ncf > 0 && b0 + wv < a.count) {
#pragma unroll
            for (int mi = 0; mi < NRF; ++mi) {
                if (mi >= nrf) continue;
                const int row = row0 + mi * 8 + l8;
                double *dst = a.planes + (((size_t)mycol * 2 + comp) * f.npy_pad + row) * f.npx_pad + col0 + cf0 * 8 + 2 * l4;
#pragma unroll
                for (int ni = 0; ni < 3; ++ni)
                    if (ni < ncf && (cf0 + ni) * 8 < ncols_valid)
                        *reinterpret_cast<double2 *>(dst + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    }
}

template <int NCF>
static void launch_fft_planes_t(cudaStream_t s, const FftPlaneArgs &a, dim3 grid, int threads, size_t smem) {
    static SmemOptIn opt;
    opt.ensure(fft_plane_kernel<NCF>, (size_t)226 * 1024);
    Launch(grid, threads, smem, s)(fft_plane_kernel<NCF>, a);
}

void launch_fft_planes(cudaStream_t s, const DevModel &m, const double *T, const double *c12, const double *cont,
                       int ldb, const int *cols, int count, double *planes) {
    const DevFft &f = m.fft;
    FftPlaneArgs a;
    a.f = f; a.T = T; a.c12 = c12; a.cont = cont; a.ldb = ldb; a.cols = cols; a.count = count; a.planes = planes;
    const size_t smem_max = 226 * 1024, conts_bytes = (size_t)2 * PV * f.my_pad * sizeof(double), stage_bytes = (size_t)(P_A + P_T) * sizeof(double);
    a.nstages = (int)std::min<size_t>(PSTAGES, conts_bytes < smem_max ? (smem_max - conts_bytes) / stage_bytes : 0);
    const size_t smem = a.nstages * stage_bytes + conts_bytes;
    // persistent CTAs: one per SM (the kernel needs most of an SM's registers and shared memory)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = f.rtiles * f.ctiles;
    dim3 grid(std::max(1, std::min(2 * ((count + PV - 1) / PV), sms / tiles)), f.rtiles, f.ctiles);
    const int threads = PV * ((f.ncf + 2) / 3) * 32;  // (vector, column group of three fragments)
    switch (f.ncf) {
    case 1: launch_fft_planes_t<1>(s, a, grid, threads, smem); break;
    case 2: launch_fft_planes_t<2>(s, a, grid, threads, smem); break;
    case 3: launch_fft_planes_t<3>(s, a, grid, threads, smem); break;
    case 4: launch_fft_planes_t<4>(s, a, grid, threads, smem); break;
    case 5: launch_fft_planes_t<5>(s, a, grid, threads, smem); break;
    case 6: launch_fft_planes_t<6>(s, a, grid, threads, smem); break;
    case 7: launch_fft_planes_t<7>(s, a, grid, threads, smem); break;
    case 8: launch_fft_planes_t<8>(s, a, grid, threads, smem); break;
    default: launch_fft_planes_t<9>(s, a, grid, threads, smem); break;
    }
}

// ------------------------------------------------------------------------------------------
// per (bin, vector): BaoKSpaceFftCorrelationModel::_evaluate, :155-250
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ double catmull(double pm1, double p0, double p1, double p2, double t) {
    return p0 + 0.5 * t * (p1 - pm1 + t * (2.0 * pm1 - 5.0 * p0 + 4.0 * p1 - p2 + t * (3.0 * (p0 - p1) + p2 - pm1)));
}

// bicubic (Catmull-Rom) interpolation on an even plane; false if the stencil leaves the plane
template <bool LDG = true>
__device__ __forceinline__ bool plane_lookup(const DevFft &f, const double *__restrict__ pl, double r, double mu,
                                             double &out) {
    auto ld = [](const double *p) { return LDG ? __ldg(p) : *p; };
    const double x = r * sqrt(1.0 - mu * mu) * f.inv_delta, y = fabs(r * mu) * f.inv_delta;
    const int ix = (int)floor(x), iy = (int)floor(y);
    if (ix + 2 > f.npx - 1 || iy + 2 > f.npy - 1 || !(x >= 0.0)) { out = nan(""); return false; }
    const double tx = x - ix, ty = y - iy;
    double col[4];
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
        const double *row = pl + (size_t)abs(iy - 1 + bq) * f.npx_pad;
        col[bq] = catmull(ld(row + abs(ix - 1)), ld(row + ix), ld(row + ix + 1), ld(row + ix + 2), tx);
    }
    out = catmull(col[0], col[1], col[2], col[3], ty);
    return true;
}
constexpr int kEvalBinsPerBlock = 512;  // 32 bin lanes x 16 iterations
constexpr int kEvalVecPerBlock = 8;    // one warp per parameter vector
}  // namespace

__global__ void __launch_bounds__(256, 4) fft_eval_kernel(FftEvalArgs a) {
    pdl_sync();
    const DevModel &m = a.m;
    const DevFft &f = m.fft;
    // One warp per parameter vector, lanes = 32 consecutive bins: consecutive bins are neighbours in r_perp, so the
    // 16 stencil loads of a plane look-up each touch ~10 consecutive sectors of one plane row per warp instead of
    // 32 scattered ones (lanes = vectors would read 32 different planes): 1.35 -> 1.20 ms for 18 944 vectors x 1515
    // bins. The price is a strided store of xi; staging it through a shared tile for 64-byte runs measured no
    // better (1.24 ms), the loop is bound by load latency (long scoreboard) at 37 % occupancy, not by the stores.
    const int warp = threadIdx.x >> 5, bin_lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * kEvalVecPerBlock, b = min(b0 + warp, a.B - 1);
    const bool valid = b0 + warp < a.B;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    const double beta = pT[(size_t)m.idx_beta * ldb];
    const double bias = pT[(size_t)m.idx_bb * ldb] / (1.0 + beta);
    const double biasSq = cross ? bias * pT[(size_t)m.idx_bias2 * ldb] : bias * bias;
    const double gammaBias = pT[(size_t)m.idx_gamma_bias * ldb];
    const double dv = cross ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    const double ampl = pT[(size_t)m.idx_bao * ldb], s_iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    const double s_par = pT[(size_t)(m.idx_bao + 2) * ldb], s_perp = pT[(size_t)(m.idx_bao + 3) * ldb];
    const double gamma_scale = pT[(size_t)(m.idx_bao + 4) * ldb];
    const double *pl_peak = a.planes + (size_t)b * 2 * f.npy_pad * f.npx_pad;
    const double *pl_smooth = pl_peak + (size_t)f.npy_pad * f.npx_pad;
    int st = 0;
    const int i0 = blockIdx.y * kEvalBinsPerBlock, i1 = min(i0 + kEvalBinsPerBlock, a.npts);
    for (int ib = i0; ib < i1; ib += 32) {
        const int i = ib + bin_lane;
        if (i < i1 && valid) {
            double r = a.r[i], mu = a.mu[i], z = a.z[i];
            if (cross) velocity_shift(fft_dpi(m, dv, z), r, mu);  // AbsCorrelationModel.cc:25
            z = fft_zcorr(f, r, mu, z);
            const double zf = (1.0 + z) / (1.0 + m.zref);
            const double evb = gammaBias == 0.0 ? 1.0 : pow(zf, gammaBias);
            const double evs = gamma_scale == 0.0 ? 1.0 : pow(zf, gamma_scale);
            double rBAO, muBAO;
            if (m.flags & BF_FLAG_ANISOTROPIC) {
                const double apar = s_par * evs, aperp = s_perp * evs, musq = mu * mu;
                const double scale = sqrt(apar * apar * musq + aperp * aperp * (1.0 - musq));
                rBAO = r * scale;
                muBAO = apar * mu / scale;
            } else {
                rBAO = r * (s_iso * evs);
                muBAO = mu;
            }
            double peak, smooth;
            bool ok = plane_lookup(f, pl_peak, rBAO, muBAO, peak);
            if (m.flags & BF_FLAG_DECOUPLED) ok &= plane_lookup(f, pl_smooth, r, mu, smooth);
            else ok &= plane_lookup(f, pl_smooth, rBAO, muBAO, smooth);
            if (!ok) st |= BF_STATUS_FFT_RANGE;
            double xi = (biasSq * evb) * (ampl * peak + smooth);
            if (m.bb[BF_BB_DIST_MUL].enabled) xi *= 1.0 + broadband_eval(m.bb[BF_BB_DIST_MUL], pT, ldb, r, mu, z);
            if (m.bb[BF_BB_DIST_ADD].enabled) xi += broadband_eval(m.bb[BF_BB_DIST_ADD], pT, ldb, r, mu, z) * evb;
            const double sub = a.dov ? a.dov[(size_t)b * a.npts + i] : (a.dsub ? a.dsub[i] : 0.0);
            a.out[(size_t)i * a.ldo + b] = xi - sub;
        }
    }
    st = __reduce_or_sync(0xffffffffu, st);
    if (st && valid && bin_lane == 0) atomicOr(a.status + b, st);
}

// The same evaluation with the two planes of a parameter vector staged in shared memory: one CTA per vector, the 83 KB of
// planes arrive by TMA bulk copies (two CTAs per SM), and the 2 x 16 stencil reads of every bin are shared-memory loads
// instead of L2 round trips. fft_eval_kernel was bound by load latency (1.14 ms for 18 944 vectors x 1515 bins at 37 %
// occupancy, its 8 vectors per CTA sweep 660 KB of planes through L1); here each plane byte crosses L2 once.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) fft_eval_smem_kernel(FftEvalArgs a) {
    pdl_sync();
    extern __shared__ __align__(128) double esm[];
    __shared__ __align__(8) unsigned long long bar;
    const DevModel &m = a.m;
    const DevFft &f = m.fft;
    const int b = blockIdx.x, tid = threadIdx.x, ldb = a.ldb;
    const unsigned plane_doubles = (unsigned)(f.npy_pad * f.npx_pad), bytes = 2u * plane_doubles * (unsigned)sizeof(double);
    if (tid == 0) {
        p_mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        p_mbar_expect_tx(&bar, bytes);
        const char *src = reinterpret_cast<const char *>(a.planes + (size_t)b * 2 * plane_doubles);
        for (unsigned off = 0; off < bytes; off += 32768u)
            p_tma_g2s(reinterpret_cast<char *>(esm) + off, src + off, min(32768u, bytes - off), &bar);
    }
    const double *pT = a.pT + b;
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    const double beta = pT[(size_t)m.idx_beta * ldb];
    const double bias = pT[(size_t)m.idx_bb * ldb] / (1.0 + beta);
    const double biasSq = cross ? bias * pT[(size_t)m.idx_bias2 * ldb] : bias * bias;
    const double gammaBias = pT[(size_t)m.idx_gamma_bias * ldb];
    const double dv = cross ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    const double ampl = pT[(size_t)m.idx_bao * ldb], s_iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    const double s_par = pT[(size_t)(m.idx_bao + 2) * ldb], s_perp = pT[(size_t)(m.idx_bao + 3) * ldb];
    const double gamma_scale = pT[(size_t)(m.idx_bao + 4) * ldb];
    const double *pl_peak = esm, *pl_smooth = esm + plane_doubles;
    __syncthreads();  // the barrier is initialised
    p_mbar_wait(&bar, 0);
    int st = 0;
    for (int i = tid; i < a.npts; i += blockDim.x) {
        double r = a.r[i], mu = a.mu[i], z = a.z[i];
        if (cross) velocity_shift(fft_dpi(m, dv, z), r, mu);  // AbsCorrelationModel.cc:25
        z = fft_zcorr(f, r, mu, z);
        const double zf = (1.0 + z) / (1.0 + m.zref);
        // ((1+z)/(1+zref))^gamma as exp(gamma log zf): one log serves both exponents and the pair costs less than half of a
        // double-precision pow (which dominated this kernel's FP64 instruction count); relative difference ~1e-16
        const double lzf = (gammaBias != 0.0 || gamma_scale != 0.0) ? log(zf) : 0.0;
        const double evb = gammaBias == 0.0 ? 1.0 : exp(gammaBias * lzf);
        const double evs = gamma_scale == 0.0 ? 1.0 : exp(gamma_scale * lzf);
        double rBAO, muBAO;
        if (m.flags & BF_FLAG_ANISOTROPIC) {
            const double apar = s_par * evs, aperp = s_perp * evs, musq = mu * mu;
            const double scale = sqrt(apar * apar * musq + aperp * aperp * (1.0 - musq));
            rBAO = r * scale;
            muBAO = apar * mu / scale;
        } else {
            rBAO = r * (s_iso * evs);
            muBAO = mu;
        }
        double peak, smooth;
        bool ok = plane_lookup<false>(f, pl_peak, rBAO, muBAO, peak);
        if (m.flags & BF_FLAG_DECOUPLED) ok &= plane_lookup<false>(f, pl_smooth, r, mu, smooth);
        else ok &= plane_lookup<false>(f, pl_smooth, rBAO, muBAO, smooth);
        if (!ok) st |= BF_STATUS_FFT_RANGE;
        double xi = (biasSq * evb) * (ampl * peak + smooth);
        if (m.bb[BF_BB_DIST_MUL].enabled) xi *= 1.0 + broadband_eval(m.bb[BF_BB_DIST_MUL], pT, ldb, r, mu, z);
        if (m.bb[BF_BB_DIST_ADD].enabled) xi += broadband_eval(m.bb[BF_BB_DIST_ADD], pT, ldb, r, mu, z) * evb;
        const double sub = a.dov ? a.dov[(size_t)b * a.npts + i] : (a.dsub ? a.dsub[i] : 0.0);
        a.out[(size_t)i * a.ldo + b] = xi - sub;
    }
    st = __reduce_or_sync(0xffffffffu, st);
    if (st && (tid & 31) == 0) atomicOr(a.status + b, st);
}

void launch_fft_eval(cudaStream_t s, const FftEvalArgs &a) {
    const DevFft &f = a.m.fft;
    const size_t smem = (size_t)2 * f.npy_pad * f.npx_pad * sizeof(double);
    static const bool no_smem = std::getenv("BAOFIT_FFT_EVAL_GLOBAL") != nullptr;  // A/B measurements
    if (!no_smem && smem <= 110 * 1024 && smem % 16 == 0 && a.B >= 64) {
        // 384 threads per CTA (two CTAs per SM by shared memory: 24 resident warps at <= 85 registers, a few spilled words)
        // against 256 at 128 registers: 1.03 -> 0.94 ms for 18 944 x 1515; 320 / 448 / 512 threads measured 0.98 / 0.94 / 0.95.
        // The loop is bound by dependency latency, not by issue slots. BAOFIT_FFT_EVAL_THREADS=256 keeps the old shape (A/B).
        static const bool t256 = std::getenv("BAOFIT_FFT_EVAL_THREADS") && std::atoi(std::getenv("BAOFIT_FFT_EVAL_THREADS")) == 256;
        static SmemOptIn opt256, opt384;
        if (t256) {
            opt256.ensure(fft_eval_smem_kernel<256>, smem);
            Launch(a.B, 256, smem, s)(fft_eval_smem_kernel<256>, a);
        } else {
            opt384.ensure(fft_eval_smem_kernel<384>, smem);
            Launch(a.B, 384, smem, s)(fft_eval_smem_kernel<384>, a);
        }
        return;
    }
    dim3 grid((a.B + kEvalVecPerBlock - 1) / kEvalVecPerBlock, (a.npts + kEvalBinsPerBlock - 1) / kEvalBinsPerBlock);
    Launch(grid, 256, 0, s)(fft_eval_kernel, a);
}

}  // namespace bf
