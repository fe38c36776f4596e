// Model-evaluation kernels: xi(r, mu, z) of B parameter vectors over all bins.
//
// Thread mapping everywhere: consecutive lanes = consecutive parameter vectors b, so that the
// bin-major outputs out[bin*ld + b] are written with coalesced, full-sector fp64 stores and the
// transposed parameters pT[j*ld + b] are read coalesced. Bin coordinates are warp-uniform.
#include <cstdlib>
#include "device_math.cuh"
#include "kernels.h"

namespace bf {

// ------------------------------------------------------------------------------------------
// prep: transpose parameters, per-(class, b) redshift-evolution factors, clear status
// ------------------------------------------------------------------------------------------
__global__ void prep_kernel(DevModel m, const double *__restrict__ params, int B, int ldb, int nz,
                            const double *__restrict__ zvals, const int *__restrict__ rank, double *__restrict__ pT,
                            double *__restrict__ S, int *__restrict__ status, int z_retrigger, double *__restrict__ key_flag) {
    pdl_sync();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) key_flag[0] = 0.0;  // kspace_key_check_kernel, which follows, raises it (a memset node here would end the dependent-launch chain)
    if (b >= B) return;
    const double *p = params + (size_t)b * m.npar;
    const int col = rank ? rank[b] : b;  // column of this vector in the (possibly sorted) panels
    for (int j = 0; j < m.npar; ++j) pT[(size_t)j * ldb + col] = p[j];
    double gb = p[m.idx_gamma_bias], gbeta = p[m.idx_gamma_beta], gs = p[m.idx_bao + 4];
    for (int c = 0; c < nz; ++c) {
        // redshiftEvolution, baofit/AbsCorrelationModel.cc:159-161
        double x = (1.0 + zvals[c]) / (1.0 + m.zref);
        S[(size_t)(c * 3 + 0) * ldb + col] = gb == 0.0 ? 1.0 : pow(x, gb);
        S[(size_t)(c * 3 + 1) * ldb + col] = gbeta == 0.0 ? 1.0 : pow(x, gbeta);
        S[(size_t)(c * 3 + 2) * ldb + col] = gs == 0.0 ? 1.0 : pow(x, gs);
    }
    // gamma-beta != 0 on data where the reference would re-transform in the middle of its sweep (bf_data::Plan::z_retrigger)
    status[col] = (z_retrigger && gbeta != 0.0) ? BF_STATUS_Z_RETRIGGER : 0;
}

// ------------------------------------------------------------------------------------------
// Coherence sort. Lanes of a warp are consecutive parameter vectors; their table look-ups hit the
// same shared-memory rows only if their dilation parameters are close. A counting sort on the
// quantised (alpha_parallel, alpha_perp) makes warps coherent for arbitrary batches (scans and
// lock-step fits already are). Results do not depend on the order: every column is independent.
// ------------------------------------------------------------------------------------------
constexpr int kSortQA = 32, kSortQP = 64, kSortBuckets = kSortQA * kSortQP;

__device__ __forceinline__ int sort_bucket(const DevModel &m, const double *__restrict__ p) {
    double a, t;
    if (m.flags & BF_FLAG_ANISOTROPIC) {
        a = p[m.idx_bao + 2];
        t = (m.flags & BF_FLAG_COMBINED_SCALE) ? p[m.idx_comb_scale] * a : p[m.idx_bao + 3];
    } else
        a = t = p[m.idx_bao + 1];
    int qa = (int)((a - 0.5) * kSortQA), qt = (int)((t - 0.5) * kSortQP);
    qa = max(0, min(kSortQA - 1, qa));
    qt = max(0, min(kSortQP - 1, qt));
    return qa * kSortQP + qt;
}

__global__ void sort_hist_kernel(DevModel m, const double *__restrict__ params, int B, int *__restrict__ hist) {
    pdl_sync();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    atomicAdd(hist + sort_bucket(m, params + (size_t)b * m.npar), 1);
}

__global__ void sort_scan_kernel(int *__restrict__ hist, int *__restrict__ base) {
    pdl_sync();
    // exclusive prefix sum over kSortBuckets (= 2 per thread of a 1024-thread block)
    __shared__ int sh[kSortBuckets];
    int t = threadIdx.x;
    sh[2 * t] = hist[2 * t];
    sh[2 * t + 1] = hist[2 * t + 1];
    __syncthreads();
    for (int off = 1; off < kSortBuckets; off <<= 1) {
        int v0 = 2 * t >= off ? sh[2 * t - off] : 0, v1 = 2 * t + 1 >= off ? sh[2 * t + 1 - off] : 0;
        __syncthreads();
        sh[2 * t] += v0;
        sh[2 * t + 1] += v1;
        __syncthreads();
    }
    base[2 * t] = sh[2 * t] - hist[2 * t];
    base[2 * t + 1] = sh[2 * t + 1] - hist[2 * t + 1];
    __syncthreads();
    hist[2 * t] = 0;  // reused as the per-bucket cursor
    hist[2 * t + 1] = 0;
}

__global__ void sort_rank_kernel(DevModel m, const double *__restrict__ params, int B, const int *__restrict__ base,
                                 int *__restrict__ cursor, int *__restrict__ rank) {
    pdl_sync();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int k = sort_bucket(m, params + (size_t)b * m.npar);
    rank[b] = base[k] + atomicAdd(cursor + k, 1);
}

__global__ void unpermute_kernel(int B, const int *__restrict__ rank, const double *__restrict__ chi2_sorted,
                                 const int *__restrict__ status_sorted, double *__restrict__ chi2, int *__restrict__ status) {
    pdl_sync();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int c = rank[b];
    chi2[b] = chi2_sorted[c];
    if (status) status[b] = status_sorted[c];
}

// ------------------------------------------------------------------------------------------
// shared tail: dilation of (r,mu), radiation, broadband
// ------------------------------------------------------------------------------------------
struct Dilated {
    double rBAO, muBAO, scalez;
};

// Everything a thread needs about its parameter vector, loaded once before the bin loop.
struct VecCtx {
    Internal in;
    double ampl, dv, apar, aperp, iso, rad[4];
    MetalVec mv;
};

template <bool METALS = true>
__device__ __forceinline__ void load_vec(const DevModel &m, const double *__restrict__ pT, int ldb, bool want_rad,
                                         VecCtx &v) {
    v.in = load_internal(m, pT, ldb);
    v.ampl = pT[(size_t)m.idx_bao * ldb];
    v.iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    v.apar = pT[(size_t)(m.idx_bao + 2) * ldb];
    v.aperp = (m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * v.apar
                                                 : pT[(size_t)(m.idx_bao + 3) * ldb];
    v.dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v.rad[k] = want_rad ? pT[(size_t)(m.idx_rad + k) * ldb] : 0.0;
    if (METALS && m.metal_kind != BF_METAL_NONE) metal_load(m, v.in, pT, ldb, v.mv);
}

__device__ __forceinline__ Dilated dilate(const DevModel &m, const VecCtx &v, double es, double r, double mu) {
    // baofit/BaoCorrelationModel.cc:104-127, baofit/BaoKSpaceCorrelationModel.cc:403-417
    Dilated d;
    if (m.flags & BF_FLAG_ANISOTROPIC) {
        double apar = v.apar * es, aperp = v.aperp * es;
        double musq = mu * mu;
        d.scalez = sqrt(apar * apar * musq + aperp * aperp * (1.0 - musq));
        d.rBAO = r * d.scalez;
        d.muBAO = apar * mu / d.scalez;
    } else {
        d.scalez = v.iso * es;
        d.rBAO = r * d.scalez;
        d.muBAO = mu;
    }
    return d;
}

__device__ __forceinline__ double post_broadband(const DevModel &m, const double *__restrict__ pT, int ldb,
                                                 double eb, double xi, double r, double mu, double z, int slot_add,
                                                 int slot_mul) {
    if (m.bb[slot_mul].enabled) xi *= 1.0 + broadband_eval(m.bb[slot_mul], pT, ldb, r, mu, z);
    if (m.bb[slot_add].enabled) xi += broadband_eval(m.bb[slot_add], pT, ldb, r, mu, z) * eb;
    return xi;
}

// ------------------------------------------------------------------------------------------
// r-space model, baofit/BaoCorrelationModel.cc:90-179 (+ AbsCorrelationModel.cc:21-41)
// ------------------------------------------------------------------------------------------
// METALS = false is the kernel for models without a metal sub-model: without the inlined template sums and the
// per-vector metal state it runs at 64 registers (a few spills to L1) and twice as many warps per SM, which is
// what this latency-bound loop needs: 1.00 -> 0.83 ms for 37 888 vectors x 440 bins (80 / 96 registers were slower).
template <bool METALS>
__global__ void __launch_bounds__(128, METALS ? 4 : 8) rspace_eval_kernel(EvalArgs a) {
    pdl_sync();
    const DevModel &m = a.m;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    VecCtx v;
    load_vec<METALS>(m, pT, ldb, m.idx_rad >= 0, v);
    const Internal &in = v.in;
    const double ampl = v.ampl;
    const double *rad = v.rad;
    int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, a.npts_pad);
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        if (i < a.npts) {
            double r = a.r[i], mu = a.mu[i], z = a.z[i];
            int zc = a.zc[i];
            double eb = S[(size_t)(zc * 3 + 0) * ldb], ebeta = S[(size_t)(zc * 3 + 1) * ldb],
                   es = S[(size_t)(zc * 3 + 2) * ldb];
            if (m.idx_delta_v >= 0) velocity_shift(v.dv * a.vfac[zc], r, mu);
            Dilated d = dilate(m, v, es, r, mu);
            double n0, n2, n4;
            norm_factors(in.biasSq * eb, in.betaAvg * ebeta, in.betaProd * ebeta * ebeta, n0, n2, n4);
            double musq = d.muBAO * d.muBAO;
            double L2 = (-1.0 + 3.0 * musq) * 0.5, L4 = (3.0 + musq * (-30.0 + 35.0 * musq)) * 0.125;
            double f[3], w[3];
            spline3_eval(m.fid, m.fid.coef, d.rBAO, f);
            spline3_eval(m.nw, m.nw.coef, d.rBAO, w);
            double fid = n0 * f[0] + n2 * L2 * f[1] + n4 * L4 * f[2];
            double nw = n0 * w[0] + n2 * L2 * w[1] + n4 * L4 * w[2];
            double peak = ampl * (fid - nw), smooth = nw;
            if (m.flags & BF_FLAG_DECOUPLED) {
                double m2 = mu * mu;
                double L2a = (-1.0 + 3.0 * m2) * 0.5, L4a = (3.0 + m2 * (-30.0 + 35.0 * m2)) * 0.125;
                spline3_eval(m.nw, m.nw.coef, r, w);
                smooth = n0 * w[0] + n2 * L2a * w[1] + n4 * L4a * w[2];
            }
            double xi = peak + smooth;
            if (METALS && m.metal_kind != BF_METAL_NONE)
                xi += metal_eval(m, in, v.mv, pT, ldb, S, a.zc_metal, a.zc_stride, i, a.gidx[i], r, mu);
            xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DIST_ADD, BF_BB_DIST_MUL);
            if (m.idx_rad >= 0 && rad[0] > 0.0 && r > 0.0) {  // :165-176
                double rd = rad[0] / (r * r);
                rd *= exp(-r / rad[2]);
                rd *= (1.0 - rad[1] * (1.0 - mu * mu));
                rd *= exp(-(r * (1.0 - mu) / (1.0 + z)) / rad[3]);
                xi += rd;
            }
            val = xi;
            if (a.dov) val -= a.dov[(size_t)b * a.npts + i];
            else if (a.dsub) val -= a.dsub[i];
        }
        a.out[(size_t)i * a.ldo + b] = val;
    }
}

// ------------------------------------------------------------------------------------------
// k-space: multipoles of D(k, mu_k), baofit/BaoKSpaceCorrelationModel.cc:215-283 and
// baofit/NonLinearCorrectionModel.cc:49-86
// ------------------------------------------------------------------------------------------
struct KPar {
    double betaz, beta2z, snl_par2, snl_perp2;
    double uv[3], hcd[3], bs[2], sg, sl;
    double qnl, kv, av, bv, kp;
};

__device__ __forceinline__ double nl_correction(const DevModel &m, const ProjArgs &a, const KPar &kp, double k,
                                                double mu_k, double pk) {
    if (m.flags & (BF_FLAG_NL_CORRECTION | BF_FLAG_FIT_NL_CORRECTION)) {
        pk = pk * a.nl_pk_scale;
        double deltaSq = k * k * k * pk / (2.0 * M_PI * M_PI);
        double growth = kp.qnl * deltaSq;
        double pec = pow(k / kp.kv, kp.av) * pow(fabs(mu_k), kp.bv);
        double pressure = (k / kp.kp) * (k / kp.kp);
        return exp(growth * (1.0 - pec) - pressure);
    }
    if (m.flags & BF_FLAG_NL_CORRECTION_ALT) {
        const double knl = 6.4, pnl = 0.569, kpp = 15.3, pp = 2.01, kv0 = 1.22, pv = 1.5, kvi = 0.923, pvi = 0.451;
        double kvel = kv0 * pow(1.0 + k / kvi, pvi);
        double growth = pow(k / knl, pnl), pressure = pow(k / kpp, pp);
        double pec = pow(fabs(k * mu_k) / kvel, pv);
        return exp(growth - pressure - pec);
    }
    return 1.0;
}

// Everything in D(k, mu_k) except the linear tracer factor (bin smoothing, Gaussian / Lorentzian
// smoothing, non-linear broadening, non-linear correction); :246-282
__device__ __forceinline__ double kspace_nonpoly(const DevModel &m, const ProjArgs &a, const KPar &kp, double k,
                                                 double mu_k, double pk) {
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    double mu2 = mu_k * mu_k, kpar = fabs(k * mu_k);
    double smoothbin = 1.0;
    if (m.flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) {
        if (m.flags & BF_FLAG_BIN_SMOOTH) {
            double kperp = k * sqrt(1.0 - mu2);
            double xp = 0.5 * kp.bs[0] * kpar, xt = 0.5 * kp.bs[1] * kperp;
            smoothbin = (sin(xp) / xp) * (sin(xt) / xt);
        }
        if (m.flags & BF_FLAG_BIN_SMOOTH_ALT) {
            double bs2 = kp.bs[0] * kp.bs[0] * mu2 + kp.bs[1] * kp.bs[1] * (1.0 - mu2);
            smoothbin = exp(-0.5 * bs2 * k * k);
        }
        smoothbin *= smoothbin;
    }
    double gaussmooth = (m.flags & BF_FLAG_SMOOTH_GAUSS) ? exp(-0.5 * kp.sg * kp.sg * kpar * kpar) : 1.0;
    double lorsmooth = (m.flags & BF_FLAG_SMOOTH_LORENTZ) ? sqrt(1.0 / (1.0 + kp.sl * kp.sl * kpar * kpar)) : 1.0;
    double snl2 = kp.snl_par2 * mu2 + kp.snl_perp2 * (1.0 - mu2);
    double nonlinear = exp(-0.5 * snl2 * k * k);
    double nlc = nl_correction(m, a, kp, k, mu_k, pk);
    if (cross) nlc = sqrt(nlc);
    return nonlinear * nlc * smoothbin * gaussmooth * lorsmooth;
}

// Linear tracer factor incl. the UV and HCD corrections; :216-245
__device__ __forceinline__ double kspace_linear(const DevModel &m, const KPar &kp, double k, double mu_k) {
    const bool cross = m.flags & BF_FLAG_CROSS_CORRELATION;
    double mu2 = mu_k * mu_k, kpar = fabs(k * mu_k);
    double tracer1 = 1.0 + kp.betaz * mu2;
    double tracer2 = cross ? 1.0 + kp.beta2z * mu2 : tracer1;
    double linear = tracer1 * tracer2;
    if (m.flags & BF_FLAG_UVFLUCTUATION) {
        double s = kp.uv[2] * k, Ws = atan(s) / s;
        double uvFunc = 1.0 + kp.uv[0] * Ws / (1.0 + kp.uv[1] * Ws);
        tracer1 = uvFunc + kp.betaz * mu2;
        tracer2 = cross ? 1.0 + kp.beta2z * mu2 : tracer1;
        linear = tracer1 * tracer2;
    }
    if (m.flags & (BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT)) {
        double x = kp.hcd[2] * kpar;
        double hcdFunc = (m.flags & BF_FLAG_HCD_MODEL) ? sin(x) / x : exp(-x * x / 2.0);
        double tracerHcd = kp.hcd[1] * (1.0 + kp.hcd[0] * mu2) * hcdFunc;
        linear = cross ? tracer1 * tracer2 + tracer2 * tracerHcd
                       : tracer1 * tracer1 + 2.0 * tracer1 * tracerHcd + tracerHcd * tracerHcd;
    }
    return linear;
}

__device__ __forceinline__ double kspace_distortion(const DevModel &m, const ProjArgs &a, const KPar &kp, double k,
                                                    double mu_k, double pk) {
    return kspace_nonpoly(m, a, kp, k, mu_k, pk) * kspace_linear(m, kp, k, mu_k);
}

__device__ __forceinline__ void load_kpar(const DevModel &m, const ProjArgs &a, const double *__restrict__ pT, int ldb,
                                          int b, int comp, KPar &kp) {
    double ebeta = a.S[(size_t)(a.zc_trigger * 3 + 1) * ldb + b];
    double beta = pT[(size_t)m.idx_beta * ldb];
    kp.betaz = beta * ebeta;
    kp.beta2z = 0.0;
    if (m.flags & BF_FLAG_CROSS_CORRELATION)
        kp.beta2z = pT[(size_t)m.idx_bb2 * ldb] / pT[(size_t)m.idx_bias2 * ldb] * ebeta;
    double snlPerp = pT[(size_t)m.idx_nl * ldb], snlPar = snlPerp * pT[(size_t)(m.idx_nl + 1) * ldb];
    kp.snl_perp2 = snlPerp * snlPerp;
    kp.snl_par2 = snlPar * snlPar;
    if (comp == 1 && !(m.flags & BF_FLAG_NL_BROADBAND)) kp.snl_perp2 = kp.snl_par2 = 0.0;  // :364-367
    for (int k = 0; k < 3; ++k) {
        kp.uv[k] = m.idx_uv >= 0 ? pT[(size_t)(m.idx_uv + k) * ldb] : 0.0;
        kp.hcd[k] = m.idx_hcd >= 0 ? pT[(size_t)(m.idx_hcd + k) * ldb] : 0.0;
    }
    for (int k = 0; k < 2; ++k) kp.bs[k] = m.idx_bs >= 0 ? pT[(size_t)(m.idx_bs + k) * ldb] : 0.0;
    kp.sg = m.idx_smgaus >= 0 ? pT[(size_t)m.idx_smgaus * ldb] : 0.0;
    kp.sl = m.idx_smlor >= 0 ? pT[(size_t)m.idx_smlor * ldb] : 0.0;
    kp.qnl = a.nl_tab[0]; kp.kv = a.nl_tab[1]; kp.av = a.nl_tab[2]; kp.bv = a.nl_tab[3]; kp.kp = a.nl_tab[4];
    if (m.flags & BF_FLAG_FIT_NL_CORRECTION) {
        kp.qnl = pT[(size_t)m.idx_nlcorr * ldb];
        kp.kp = pT[(size_t)(m.idx_nlcorr + 1) * ldb];
    }
}

// G[((comp*3 + l)*nk_pad + i)*ldb + b] = (2l+1) Int_0^1 L_l(mu) D_comp(k_i, mu; p_b) dmu
__global__ void __launch_bounds__(128) kspace_project_kernel(ProjArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const int comp = blockIdx.z;
    KPar kp;
    load_kpar(m, a, pT, ldb, b, comp, kp);
    const int lmax = m.ell_max / 2;
    int i0 = blockIdx.y * a.k_per_block, i1 = min(i0 + a.k_per_block, m.nk_pad);
    for (int i = i0; i < i1; ++i) {
        double s0 = 0.0, s2 = 0.0, s4 = 0.0;
        if (i < m.nk) {
            double k = m.kc[i], pk = m.pk_kc[comp * m.nk + i];
#pragma unroll 4
            for (int q = 0; q < kMuNodes; ++q) {
                double mu = m.mu_x[q], mu2 = mu * mu;
                double wD = m.mu_w[q] * kspace_distortion(m, a, kp, k, mu, pk);
                s0 += wD;
                s2 += wD * ((-1.0 + 3.0 * mu2) * 0.5);
                s4 += wD * ((3.0 + mu2 * (-30.0 + 35.0 * mu2)) * 0.125);
            }
        }
        size_t o = ((size_t)(comp * 3) * m.nk_pad + i) * ldb + b;
        size_t st = (size_t)m.nk_pad * ldb;
        a.G[o] = s0;
        a.G[o + st] = lmax >= 1 ? 5.0 * s2 : 0.0;
        a.G[o + 2 * st] = lmax >= 2 ? 9.0 * s4 : 0.0;
    }
}

// Natural-spline second derivatives (gsl "c" = y''/2) of the per-vector multipole tables on the
// uniform radial grid: constant (1,4,1) tridiagonal system, factors precomputed on the host.
// One thread per (b, table); consecutive lanes = consecutive b => coalesced column sweeps.
__global__ void __launch_bounds__(128) kspace_thomas_kernel(int B, int ldb, int nr, int nr_pad, double h,
                                                            const double *__restrict__ th_w,
                                                            const double *__restrict__ th_inv_diag,
                                                            const double *__restrict__ T, double *__restrict__ C2, Gate gate) {
    pdl_sync();
    BF_GATE(gate);
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *y = T + (size_t)blockIdx.y * nr_pad * ldb + b;
    double *c = C2 + (size_t)blockIdx.y * nr_pad * ldb + b;
    const int msz = nr - 2;
    const double inv_h = 1.0 / h;
    c[0] = 0.0;
    double y0 = y[0], y1 = y[ldb], gprev = 0.0;
    for (int i = 0; i < msz; ++i) {
        double y2 = y[(size_t)(i + 2) * ldb];
        double g = 3.0 * ((y2 - y1) * inv_h - (y1 - y0) * inv_h);
        if (i > 0) g -= th_w[i] * gprev;
        c[(size_t)(i + 1) * ldb] = g;
        gprev = g;
        y0 = y1;
        y1 = y2;
    }
    c[(size_t)(nr - 1) * ldb] = 0.0;
    double cnext = 0.0;
    for (int i = msz - 1; i >= 0; --i) {
        double ci = (c[(size_t)(i + 1) * ldb] - h * cnext) * th_inv_diag[i];
        c[(size_t)(i + 1) * ldb] = ci;
        cnext = ci;
    }
}

// cosmo::DistortedPowerCorrelation::getCorrelation: sum_l L_l(mu) spline_l(r) for one component
__device__ __forceinline__ double kspace_correlation(const EvalArgs &a, int comp, int b, double r, double mu) {
    const DevModel &m = a.m;
    const int nr = m.nr;
    const size_t st = (size_t)m.nr_pad * a.ldb;
    const double *T = a.T + (size_t)(comp * 3) * st + b;
    const double *C = a.C2 + (size_t)(comp * 3) * st + b;
    double musq = mu * mu;
    double Lw[3] = {1.0, (-1.0 + 3.0 * musq) * 0.5, (3.0 + musq * (-30.0 + 35.0 * musq)) * 0.125};
    double rlast = m.tr_r0 + (nr - 1) * m.tr_dr;
    double out = 0.0;
    if (r <= m.tr_r0 || r >= rlast) {
        size_t row = r <= m.tr_r0 ? 0 : (size_t)(nr - 1) * a.ldb;
#pragma unroll
        for (int l = 0; l < 3; ++l) out += Lw[l] * T[l * st + row];
        return out;
    }
    int j = (int)((r - m.tr_r0) * m.tr_inv_dr);
    j = max(0, min(j, nr - 2));
    double xl = m.tr_r0 + j * m.tr_dr;
    if (r < xl && j > 0) { --j; xl = m.tr_r0 + j * m.tr_dr; }
    double xh = m.tr_r0 + (j + 1) * m.tr_dr;
    if (r >= xh && j < nr - 2) { ++j; xl = xh; xh = m.tr_r0 + (j + 1) * m.tr_dr; }
    double dx = xh - xl, t = r - xl;
    size_t row = (size_t)j * a.ldb;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        double ylo = T[l * st + row], yhi = T[l * st + row + a.ldb];
        double clo = C[l * st + row], chi = C[l * st + row + a.ldb];
        double bb = (yhi - ylo) / dx - dx * (chi + 2.0 * clo) / 3.0;
        double dd = (chi - clo) / (3.0 * dx);
        out += Lw[l] * (ylo + t * (bb + t * (clo + t * dd)));
    }
    return out;
}

// k-space model at a set of points, baofit/BaoKSpaceCorrelationModel.cc:285-538.
//   mode 0: data bins, no distortion matrix: full model incl. post broadband (:392-455,:529-535)
//   mode 1: grid bins for the distortion matrix: undistorted xi_u with the range zeroing and the
//           pre-matrix broadband (:460-521)
__global__ void __launch_bounds__(128) kspace_eval_kernel(EvalArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    VecCtx v;
    load_vec(m, pT, ldb, (m.flags & BF_FLAG_RADIATION_MODEL) != 0, v);
    const Internal &in = v.in;
    const double ampl = v.ampl;
    const double *rad = v.rad;
    int st_bits = 0;
    int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, a.npts_pad);
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        if (i < a.npts) {
            double r = a.r[i], mu = a.mu[i], z = a.z[i];
            int zc = a.zc[i];
            double eb = S[(size_t)(zc * 3 + 0) * ldb], es = S[(size_t)(zc * 3 + 2) * ldb];
            if (m.idx_delta_v >= 0) velocity_shift(v.dv * a.vfac[zc], r, mu);
            Dilated d = dilate(m, v, es, r, mu);
            bool zero = false;
            if (a.mode == 1) {
                zero = (r < m.rlo || r > m.rhi) || (d.rBAO < m.rlo || d.rBAO > m.rhi);
            } else {
                if (d.scalez < m.dilmin) st_bits |= BF_STATUS_DILATION_MIN;
                else if (d.scalez > m.dilmax) st_bits |= BF_STATUS_DILATION_MAX;
            }
            if (!zero) {
                double peak = kspace_correlation(a, 0, b, d.rBAO, d.muBAO);
                double smooth = (m.flags & BF_FLAG_DECOUPLED) ? kspace_correlation(a, 1, b, r, mu)
                                                              : kspace_correlation(a, 1, b, d.rBAO, d.muBAO);
                double xi = in.biasSq * eb * (ampl * peak + smooth);
                if (m.metal_kind != BF_METAL_NONE)
                    xi += metal_eval(m, in, v.mv, pT, ldb, S, a.zc_metal, a.zc_stride, i, a.gidx ? a.gidx[i] : i, r, mu);
                if ((m.flags & BF_FLAG_RADIATION_MODEL) && r > 0.0) {  // :441-455
                    double rd = rad[0] / (r * r);
                    if (rad[1] > 0.0) rd *= (1.0 - rad[1] * (1.0 - mu * mu));
                    if (rad[2] > 0.0) rd *= exp(-r / rad[2]);
                    if (rad[3] > 0.0) rd *= exp(-(r * (1.0 - mu) / (1.0 + z)) / rad[3]);
                    xi += rd;
                }
                if (a.mode == 1) xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DM_DIST_ADD, BF_BB_DM_DIST_MUL);
                else xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DIST_ADD, BF_BB_DIST_MUL);
                val = xi;
            }
            if (a.mode == 0) {
                if (st_bits) val = nan("");
                if (a.dov) val -= a.dov[(size_t)b * a.npts + i];
                else if (a.dsub) val -= a.dsub[i];
            }
        }
        a.out[(size_t)i * a.ldo + b] = val;
    }
    if (st_bits) atomicOr(a.status + b, st_bits);
}

// ------------------------------------------------------------------------------------------
// Shared-transform fast path.
//
// Without the UV / HCD corrections the tracer factor of D(k,mu) is the polynomial
//     (1 + beta mu^2)(1 + beta2 mu^2) = 1 + c1 mu^2 + c2 mu^4
// and everything else, K(k,mu), depends only on "key" parameters (SigmaNL-perp, 1+f, bin / Gauss /
// Lorentz smoothing scales, nlcorr) that fits normally keep fixed. The whole chain projection ->
// Hankel transform -> natural spline is linear in the multipoles of D, so when every vector of a
// batch has the same key
//     xi_l^comp(r; b) = X_{l,0}(r) + c1(b) X_{l,1}(r) + c2(b) X_{l,2}(r)
// with 18 basis tables X computed once per batch. The per-vector transform disappears and the
// tables become small and static: they are staged in shared memory with one TMA bulk copy.
// ------------------------------------------------------------------------------------------
__global__ void kspace_key_check_kernel(DevModel m, const double *__restrict__ pT, int ldb, int B, double *keyout) {
    pdl_sync();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bool diff = false;
    int nkey = 0;
    auto cmp = [&](int idx) {
        double v0 = pT[(size_t)idx * ldb];
        diff |= (pT[(size_t)idx * ldb + b] != v0);
        if (b == 0) keyout[1 + nkey] = v0;
        ++nkey;
    };
    cmp(m.idx_nl);
    cmp(m.idx_nl + 1);
    if (m.flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) { cmp(m.idx_bs); cmp(m.idx_bs + 1); }
    if (m.flags & BF_FLAG_SMOOTH_GAUSS) cmp(m.idx_smgaus);
    if (m.flags & BF_FLAG_SMOOTH_LORENTZ) cmp(m.idx_smlor);
    if (m.flags & BF_FLAG_FIT_NL_CORRECTION) { cmp(m.idx_nlcorr); cmp(m.idx_nlcorr + 1); }
    if (b == 0)
        for (int k = nkey; k < kKeyMax - 1; ++k) keyout[1 + k] = 0.0;
    if (diff) keyout[0] = 1.0;
}

// The decision the host used to take after copying the flag back (and waiting for it): do all vectors share their key,
// and if so, were the cached basis tables built for this key? state[0..17] = key of the cached tables, state[31] = 1 once
// tables exist. Runs as one thread; the kernels that follow read gate[] (struct Gate).
__global__ void kspace_key_decide_kernel(const double *__restrict__ flag, double *__restrict__ state, int *__restrict__ gate, KeyExtras ex) {
    pdl_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool shared = flag[0] == 0.0;
    bool same = state[31] == 1.0;
    for (int k = 0; k < kKeyMax - 1; ++k) same &= state[k] == flag[1 + k];
    for (int k = 0; k < 7; ++k) same &= state[kKeyMax - 1 + k] == ex.v[k];
    const bool rebuild = shared && !same;
    if (rebuild) {
        for (int k = 0; k < kKeyMax - 1; ++k) state[k] = flag[1 + k];
        for (int k = 0; k < 7; ++k) state[kKeyMax - 1 + k] = ex.v[k];
        state[31] = 1.0;
    }
    gate[0] = shared ? 1 : 0;
    gate[1] = rebuild ? 1 : 0;
}

// G[((comp*3 + l)*nk_pad + i)*kPadN + n] = (2l+1) Int_0^1 L_l(mu) mu^(2n) K_comp(k_i, mu) dmu
__global__ void __launch_bounds__(128) kspace_basis_project_kernel(ProjArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, comp = blockIdx.y;
    if (i >= m.nk_pad) return;
    KPar kp;
    load_kpar(m, a, a.pT, a.ldb, 0, comp, kp);
    double s[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (i < m.nk) {
        double k = m.kc[i], pk = m.pk_kc[comp * m.nk + i];
        for (int q = 0; q < kMuNodes; ++q) {
            double mu = m.mu_x[q], mu2 = mu * mu;
            double w = m.mu_w[q] * kspace_nonpoly(m, a, kp, k, mu, pk);
            double L[3] = {1.0, (-1.0 + 3.0 * mu2) * 0.5, (3.0 + mu2 * (-30.0 + 35.0 * mu2)) * 0.125};
            double pw = w;
#pragma unroll
            for (int n = 0; n < 3; ++n) {
#pragma unroll
                for (int l = 0; l < 3; ++l) s[l][n] += pw * L[l];
                pw *= mu2;
            }
        }
    }
    const int lmax = m.ell_max / 2;
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int n = 0; n < 3; ++n)
            a.G[((size_t)(comp * 3 + l) * m.nk_pad + i) * kPadN + n] = l <= lmax ? (2 * l * 2 + 1) * s[l][n] : 0.0;
    // the padding columns the Hankel GEMM's tiles read (written here rather than by a memset node before the kernel)
    for (int l = 0; l < 3; ++l)
        for (int n = 3; n < kPadN; ++n) a.G[((size_t)(comp * 3 + l) * m.nk_pad + i) * kPadN + n] = 0.0;
}

// Natural-spline second derivatives of the 18 basis tables and packing for the evaluation
// kernel: tabs[(comp*nr + j)*9 + l*3 + n] = { X_{l,n}(r_j), X''_{l,n}(r_j)/2 }.
// One CTA; the 18 serial (1,4,1) sweeps run in shared memory instead of global memory.
__global__ void __launch_bounds__(256) kspace_basis_finish_kernel(DevModel m, const double *__restrict__ T,
                                                                  double2 *__restrict__ tabs, Gate gate) {
    pdl_sync();
    BF_GATE(gate);
    extern __shared__ double sh[];  // y[18][nr], c[18][nr]
    const int nr = m.nr;
    double *y = sh, *c = sh + 18 * nr;
    for (int idx = threadIdx.x; idx < 18 * nr; idx += blockDim.x) {
        int t = idx / nr, j = idx % nr;  // t = comp*9 + l*3 + n
        int comp = t / 9, l = (t % 9) / 3, n = t % 3;
        y[idx] = T[((size_t)(comp * 3 + l) * m.nr_pad + j) * kPadN + n];
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        double *yy = y + threadIdx.x * nr, *cc = c + threadIdx.x * nr;
        const int msz = nr - 2;
        const double h = m.tr_dr, inv_h = 1.0 / h;
        cc[0] = 0.0;
        cc[nr - 1] = 0.0;
        double gprev = 0.0;
        for (int i = 0; i < msz; ++i) {
            double g = 3.0 * ((yy[i + 2] - yy[i + 1]) * inv_h - (yy[i + 1] - yy[i]) * inv_h);
            if (i > 0) g -= m.th_w[i] * gprev;
            cc[i + 1] = g;
            gprev = g;
        }
        double cnext = 0.0;
        for (int i = msz - 1; i >= 0; --i) {
            double ci = (cc[i + 1] - h * cnext) * m.th_inv_diag[i];
            cc[i + 1] = ci;
            cnext = ci;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 18 * nr; idx += blockDim.x) {
        int t = idx / nr, j = idx % nr;
        int comp = t / 9, q = t % 9;
        tabs[((size_t)comp * nr + j) * 9 + q] = make_double2(y[idx], c[idx]);
    }
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
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

// sum_l L_l(mu) sum_n coef_n X_{l,n}(r) for one component from the shared-memory basis tables
__device__ __forceinline__ double shared_correlation(const DevModel &m, const double2 *__restrict__ tab, double r,
                                                     double mu, double c1, double c2) {
    const int nr = m.nr;
    const double rlast = m.tr_r0 + (nr - 1) * m.tr_dr;
    double w0, w1, w2, w3;
    int j;
    if (r <= m.tr_r0) { j = 0; w0 = 1.0; w1 = w2 = w3 = 0.0; }
    else if (r >= rlast) { j = nr - 2; w1 = 1.0; w0 = w2 = w3 = 0.0; }
    else {
        j = (int)((r - m.tr_r0) * m.tr_inv_dr);
        j = max(0, min(j, nr - 2));
        double xl = m.tr_r0 + j * m.tr_dr;
        if (r < xl && j > 0) { --j; xl = m.tr_r0 + j * m.tr_dr; }
        else if (r >= xl + m.tr_dr && j < nr - 2) { ++j; xl = m.tr_r0 + j * m.tr_dr; }
        const double dx = m.tr_dr, t = r - xl, u = t * m.tr_inv_dr, t3 = t * t * t * m.tr_inv_dr * (1.0 / 3.0);
        // gsl cspline evaluation rearranged as weights of (y_lo, y_hi, c_lo, c_hi)
        w0 = 1.0 - u;
        w1 = u;
        w2 = t * t - (2.0 / 3.0) * t * dx - t3;
        w3 = t3 - (1.0 / 3.0) * t * dx;
    }
    const double musq = mu * mu;
    const double L2 = (-1.0 + 3.0 * musq) * 0.5, L4 = (3.0 + musq * (-30.0 + 35.0 * musq)) * 0.125;
    const double g[9] = {1.0, c1, c2, L2, L2 * c1, L2 * c2, L4, L4 * c1, L4 * c2};
    const double2 *lo = tab + (size_t)j * 9, *hi = lo + 9;
    double ylo = 0.0, yhi = 0.0, clo = 0.0, chi = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        double2 a = lo[q], b = hi[q];
        ylo = fma(g[q], a.x, ylo);
        clo = fma(g[q], a.y, clo);
        yhi = fma(g[q], b.x, yhi);
        chi = fma(g[q], b.y, chi);
    }
    return w0 * ylo + w1 * yhi + w2 * clo + w3 * chi;
}

// Same contract as kspace_eval_kernel, basis tables in shared memory (TMA bulk copy + mbarrier).
__global__ void __launch_bounds__(256, 2) kspace_eval_shared_kernel(EvalArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tab = reinterpret_cast<double2 *>(smem_raw);
    __shared__ __align__(8) unsigned long long bar;
    const DevModel &m = a.m;
    const unsigned tab_bytes = 2u * m.nr * 9u * (unsigned)sizeof(double2);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, tab_bytes);
        const unsigned chunk = 32768;
        for (unsigned off = 0; off < tab_bytes; off += chunk)
            tma_bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char *>(a.tabs) + off,
                         min(chunk, tab_bytes - off), &bar);
    }
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = b < a.B;
    if (!active) b = a.B - 1;  // keep the whole CTA alive for the barrier; results are not stored
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    VecCtx v;
    load_vec(m, pT, ldb, (m.flags & BF_FLAG_RADIATION_MODEL) != 0, v);
    const Internal &in = v.in;
    const double ampl = v.ampl;
    const double *rad = v.rad;
    double c1, c2;
    {
        double ebt = a.S[(size_t)(a.zc_trigger * 3 + 1) * ldb + b];
        double betaz = in.beta * ebt;
        if (m.flags & BF_FLAG_CROSS_CORRELATION) {
            double beta2z = in.beta2 * ebt;
            c1 = betaz + beta2z;
            c2 = betaz * beta2z;
        } else {
            c1 = 2.0 * betaz;
            c2 = betaz * betaz;
        }
    }
    mbar_wait(&bar, 0);
    const double2 *tab_pk = tab, *tab_nw = tab + (size_t)m.nr * 9;
    int st_bits = 0;
    int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, a.npts_pad);
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        if (i < a.npts) {
            double r = a.r[i], mu = a.mu[i], z = a.z[i];
            int zc = a.zc[i];
            double eb = S[(size_t)(zc * 3 + 0) * ldb], es = S[(size_t)(zc * 3 + 2) * ldb];
            if (m.idx_delta_v >= 0) velocity_shift(v.dv * a.vfac[zc], r, mu);
            Dilated d = dilate(m, v, es, r, mu);
            bool zero = false;
            if (a.mode == 1) {
                zero = (r < m.rlo || r > m.rhi) || (d.rBAO < m.rlo || d.rBAO > m.rhi);
            } else {
                if (d.scalez < m.dilmin) st_bits |= BF_STATUS_DILATION_MIN;
                else if (d.scalez > m.dilmax) st_bits |= BF_STATUS_DILATION_MAX;
            }
            if (!zero) {
                double peak = shared_correlation(m, tab_pk, d.rBAO, d.muBAO, c1, c2);
                double smooth = (m.flags & BF_FLAG_DECOUPLED) ? shared_correlation(m, tab_nw, r, mu, c1, c2)
                                                              : shared_correlation(m, tab_nw, d.rBAO, d.muBAO, c1, c2);
                double xi = in.biasSq * eb * (ampl * peak + smooth);
                if (m.metal_kind != BF_METAL_NONE)
                    xi += metal_eval(m, in, v.mv, pT, ldb, S, a.zc_metal, a.zc_stride, i, a.gidx ? a.gidx[i] : i, r, mu);
                if ((m.flags & BF_FLAG_RADIATION_MODEL) && r > 0.0) {
                    double rd = rad[0] / (r * r);
                    if (rad[1] > 0.0) rd *= (1.0 - rad[1] * (1.0 - mu * mu));
                    if (rad[2] > 0.0) rd *= exp(-r / rad[2]);
                    if (rad[3] > 0.0) rd *= exp(-(r * (1.0 - mu) / (1.0 + z)) / rad[3]);
                    xi += rd;
                }
                if (a.mode == 1) xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DM_DIST_ADD, BF_BB_DM_DIST_MUL);
                else xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DIST_ADD, BF_BB_DIST_MUL);
                val = xi;
            }
            if (a.mode == 0) {
                if (st_bits) val = nan("");
                if (a.dov) val -= a.dov[(size_t)b * a.npts + i];
                else if (a.dsub) val -= a.dsub[i];
            }
        }
        if (active) a.out[(size_t)i * a.ldo + b] = val;
    }
    if (active && st_bits) atomicOr(a.status + b, st_bits);
}

// Distortion-matrix tail for data bins: Z = Y (1 + dist mul) + dist add * evol - d, plus the
// dilation-limit check of the data bin itself (baofit/BaoKSpaceCorrelationModel.cc:420-427,
// :529-535). Y = D_kept . xi_u comes from the DMMA GEMM.
__global__ void __launch_bounds__(128, 4) dist_post_kernel(EvalArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    VecCtx v;
    v.in = Internal();
    v.iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    v.apar = pT[(size_t)(m.idx_bao + 2) * ldb];
    v.aperp = (m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * v.apar
                                                 : pT[(size_t)(m.idx_bao + 3) * ldb];
    v.dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    int st_bits = 0;
    int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, a.npts_pad);
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        if (i < a.npts) {
            double r = a.r[i], mu = a.mu[i], z = a.z[i];
            int zc = a.zc[i];
            double eb = S[(size_t)(zc * 3 + 0) * ldb], es = S[(size_t)(zc * 3 + 2) * ldb];
            if (m.idx_delta_v >= 0) velocity_shift(v.dv * a.vfac[zc], r, mu);
            Dilated d = dilate(m, v, es, r, mu);
            if (d.scalez < m.dilmin) st_bits |= BF_STATUS_DILATION_MIN;
            else if (d.scalez > m.dilmax) st_bits |= BF_STATUS_DILATION_MAX;
            double xi = a.in[(size_t)i * a.ldo + b];
            xi = post_broadband(m, pT, ldb, eb, xi, r, mu, z, BF_BB_DIST_ADD, BF_BB_DIST_MUL);
            val = st_bits ? nan("") : xi;
            if (a.dov) val -= a.dov[(size_t)b * a.npts + i];
            else if (a.dsub) val -= a.dsub[i];
        }
        a.out[(size_t)i * a.ldo + b] = val;
    }
    if (st_bits) atomicOr(a.status + b, st_bits);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int pick_pts_per_block(int npts_pad, int B, int sm_count) {
    // enough CTAs for >= 4 waves when B is small, long bin loops when B is large
    int bx = (B + 127) / 128;
    int want = sm_count * 8;
    int by = max(1, min(npts_pad, (want + bx - 1) / bx));
    int ppb = (npts_pad + by - 1) / by;
    return max(1, ppb);
}

void launch_prep(cudaStream_t s, const DevModel &m, const double *params, int B, int ldb, int nz,
                 const double *zvals, const int *rank, double *pT, double *S, int *status, bool z_retrigger, double *key_flag) {
    Launch((B + 127) / 128, 128, 0, s)(prep_kernel, m, params, B, ldb, nz, zvals, rank, pT, S, status, z_retrigger ? 1 : 0, key_flag);
}

void launch_coherence_sort(cudaStream_t s, const DevModel &m, const double *params, int B, int *hist, int *base, int *rank) {
    cudaMemsetAsync(hist, 0, sizeof(int) * kSortBuckets, s);
    Launch((B + 255) / 256, 256, 0, s)(sort_hist_kernel, m, params, B, hist);
    Launch(1, kSortBuckets / 2, 0, s)(sort_scan_kernel, hist, base);
    Launch((B + 255) / 256, 256, 0, s)(sort_rank_kernel, m, params, B, base, hist, rank);
}

void launch_unpermute(cudaStream_t s, int B, const int *rank, const double *chi2_sorted, const int *status_sorted,
                      double *chi2, int *status) {
    Launch((B + 255) / 256, 256, 0, s)(unpermute_kernel, B, rank, chi2_sorted, status_sorted, chi2, status);
}

void launch_rspace_eval(cudaStream_t s, EvalArgs &a, int sm_count) {
    a.pts_per_block = pick_pts_per_block(a.npts_pad, a.B, sm_count);
    dim3 grid((a.B + 127) / 128, (a.npts_pad + a.pts_per_block - 1) / a.pts_per_block);
    if (a.m.metal_kind != BF_METAL_NONE) Launch(grid, 128, 0, s)(rspace_eval_kernel<true>, a);
    else Launch(grid, 128, 0, s)(rspace_eval_kernel<false>, a);
}

void launch_kspace_project(cudaStream_t s, ProjArgs &a, int sm_count) {
    a.k_per_block = pick_pts_per_block(a.m.nk_pad, a.B, sm_count);
    dim3 grid((a.B + 127) / 128, (a.m.nk_pad + a.k_per_block - 1) / a.k_per_block, 2);
    Launch(grid, 128, 0, s)(kspace_project_kernel, a);
}

void launch_kspace_thomas(cudaStream_t s, int B, int ldb, int nr, int nr_pad, double h, const double *th_w,
                          const double *th_inv_diag, const double *T, double *C2, Gate gate) {
    dim3 grid((B + 127) / 128, 6);
    Launch(grid, 128, 0, s)(kspace_thomas_kernel, B, ldb, nr, nr_pad, h, th_w, th_inv_diag, T, C2, gate);
}

void launch_kspace_eval(cudaStream_t s, EvalArgs &a, int sm_count) {
    a.pts_per_block = pick_pts_per_block(a.npts_pad, a.B, sm_count);
    dim3 grid((a.B + 127) / 128, (a.npts_pad + a.pts_per_block - 1) / a.pts_per_block);
    Launch(grid, 128, 0, s)(kspace_eval_kernel, a);
}

void launch_kspace_key_check(cudaStream_t s, const DevModel &m, const double *pT, int ldb, int B, double *keyout) {
    Launch((B + 255) / 256, 256, 0, s)(kspace_key_check_kernel, m, pT, ldb, B, keyout);
}

void launch_kspace_key_decide(cudaStream_t s, const double *flag, double *state, int *gate, const KeyExtras &ex) {
    Launch(1, 32, 0, s)(kspace_key_decide_kernel, flag, state, gate, ex);
}

void launch_kspace_basis_project(cudaStream_t s, ProjArgs &a) {
    dim3 grid((a.m.nk_pad + 127) / 128, 2);
    Launch(grid, 128, 0, s)(kspace_basis_project_kernel, a);
}

void launch_kspace_basis_finish(cudaStream_t s, const DevModel &m, const double *T, double2 *tabs, Gate gate) {
    size_t smem = (size_t)36 * m.nr * sizeof(double);
    static SmemOptIn opt;
    opt.ensure(kspace_basis_finish_kernel, smem);
    Launch(1, 256, smem, s)(kspace_basis_finish_kernel, m, T, tabs, gate);
}

void launch_kspace_eval_shared(cudaStream_t s, EvalArgs &a, int sm_count) {
    size_t smem = (size_t)2 * a.m.nr * 9 * sizeof(double2);
    static SmemOptIn opt;
    opt.ensure(kspace_eval_shared_kernel, smem);
    int bx = (a.B + 255) / 256;
    int want = sm_count * 6;
    int by = max(1, min(a.npts_pad, (want + bx - 1) / bx));
    a.pts_per_block = (a.npts_pad + by - 1) / by;
    dim3 grid(bx, (a.npts_pad + a.pts_per_block - 1) / a.pts_per_block);
    Launch(grid, 256, smem, s)(kspace_eval_shared_kernel, a);
}

void launch_dist_post(cudaStream_t s, EvalArgs &a, int sm_count) {
    a.pts_per_block = pick_pts_per_block(a.npts_pad, a.B, sm_count);
    dim3 grid((a.B + 127) / 128, (a.npts_pad + a.pts_per_block - 1) / a.pts_per_block);
    Launch(grid, 128, 0, s)(dist_post_kernel, a);
}

// ------------------------------------------------------------------------------------------
// multipole-binned data: AbsCorrelationModel::evaluate(r, multipole, z, ...) = Gauss-Legendre
// projection of the (r, mu) model (AbsCorrelationModel.cc:65-79)
// ------------------------------------------------------------------------------------------
__global__ void multipole_project_kernel(const double *__restrict__ X, int ldx, const double *__restrict__ proj, int n, int nmu,
                                         int B, double *__restrict__ out, int ldo, const double *__restrict__ dsub,
                                         const double *__restrict__ dov) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (b >= B) return;
    double s = 0.0;
    for (int q = 0; q < nmu; ++q) s = fma(proj[i * nmu + q], X[(size_t)(i * nmu + q) * ldx + b], s);
    const double sub = dov ? dov[(size_t)b * n + i] : (dsub ? dsub[i] : 0.0);
    out[(size_t)i * ldo + b] = s - sub;
}

void launch_multipole_project(cudaStream_t s, const double *X, int ldx, const double *proj, int n, int nmu, int B,
                              double *out, int ldo, const double *dsub, const double *dov) {
    Launch(dim3((B + 127) / 128, n), 128, 0, s)(multipole_project_kernel, X, ldx, proj, n, nmu, B, out, ldo, dsub, dov);
}

__global__ void gather_rows_kernel(const double *__restrict__ table, const int *__restrict__ index, int B, int n,
                                   double *__restrict__ out) {
    pdl_sync();
    const int b = blockIdx.y;
    const double *src = table + (size_t)index[b] * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[(size_t)b * n + i] = src[i];
}

void launch_gather_rows(cudaStream_t s, const double *table, const int *index, int B, int n, double *out) {
    Launch(dim3((n + 255) / 256, B), 256, 0, s)(gather_rows_kernel, table, index, B, n, out);
}

// ------------------------------------------------------------------------------------------
// Gauss-Newton normal equations: one CTA per fit
// ------------------------------------------------------------------------------------------
constexpr int kGramRows = 32;  // G <= 96 (checked by the callers): at most 19 pairs per thread
constexpr int kProbeSplit = 8;  // row slices of the normal-equation kernels

__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ Y, int ldy, int nrows, int G, int nfit,
                                                   double *__restrict__ out) {
    pdl_sync();
    extern __shared__ double gsm[];  // [kGramRows][G]
    const int f = blockIdx.x, tid = threadIdx.x;
    const int npairs = G * (G + 1) / 2;
    // every thread owns the pairs tid, tid + 256, ... (i >= j), at most kGramMaxG*(kGramMaxG+1)/2/256 = 19
    double acc[19];
    int pi[19], pj[19], np = 0;
    for (int p = tid; p < npairs && np < 19; p += blockDim.x) {
        int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > p) --i;
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        pi[np] = i; pj[np] = p - i * (i + 1) / 2; acc[np] = 0.0; ++np;
    }
    const double *base = Y + (size_t)f * G;
    for (int r0 = 0; r0 < nrows; r0 += kGramRows) {
        const int nr = min(kGramRows, nrows - r0);
        __syncthreads();
        for (int e = tid; e < nr * G; e += blockDim.x) {
            int r = e / G, c = e - r * G;
            const double *row = base + (size_t)(r0 + r) * ldy;
            gsm[r * G + c] = c == 0 ? row[0] : row[c] - row[0];
        }
        __syncthreads();
        for (int q = 0; q < np; ++q) {
            double s = acc[q];
            const int i = pi[q], j = pj[q];
            for (int r = 0; r < nr; ++r) s = fma(gsm[r * G + i], gsm[r * G + j], s);
            acc[q] = s;
        }
    }
    double *o = out + (size_t)f * G * G;
    for (int q = 0; q < np; ++q) { o[pi[q] * G + pj[q]] = acc[q]; o[pj[q] * G + pi[q]] = acc[q]; }
}

// Normal equations of central probes. The G = 2 n + 1 columns of a fit are y0, y(x + h_1), y(x - h_1), ...; with
// a_i = y(x + h_i) - y(x - h_i) and s_i = y(x + h_i) + y(x - h_i) - 2 y0 a minimiser needs only
//   M = [y0 a_1 .. a_n]^T [y0 a_1 .. a_n]   ((n+1)^2, symmetric)   and   t_i = y0 . s_i   (n values),
// a quarter of the full Gram matrix to compute and to send back: out[f] = M (row-major) followed by t.
// The rows are cut into kProbeSplit slices (always, so that the sums do not depend on how many fits share a call): CTA
// (f, s) writes its partial sums to part[(f * kProbeSplit + s) * npairs + p], normal_probes_sum_kernel adds them in order.
__global__ void __launch_bounds__(256) normal_probes_kernel(const double *__restrict__ Y, int ldy, int nrows, int n, int nfit,
                                                            double *__restrict__ part) {
    pdl_sync();
    extern __shared__ double nsm[];  // [kGramRows][R + n]: y0, a_1..a_n, s_1..s_n
    const int f = blockIdx.x, tid = threadIdx.x, G = 2 * n + 1, R = n + 1, W = R + n;
    const int npairs = R * (R + 1) / 2 + n;  // pairs (i >= j) of M, then the n dot products y0 . s_i
    double acc[6];                           // R <= 48: at most 1224 values over 256 threads
    int pi[6], pj[6], np = 0;
    for (int p = tid; p < npairs && np < 6; p += blockDim.x) {
        int i, j;
        if (p < R * (R + 1) / 2) {
            i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > p) --i;
            while ((i + 1) * (i + 2) / 2 <= p) ++i;
            j = p - i * (i + 1) / 2;
        } else {
            i = R + (p - R * (R + 1) / 2);  // column of s
            j = 0;
        }
        pi[np] = i; pj[np] = j; acc[np] = 0.0; ++np;
    }
    const double *base = Y + (size_t)f * G;
    const int per_slice = (nrows + kProbeSplit - 1) / kProbeSplit, r_lo = min(nrows, (int)blockIdx.y * per_slice), r_hi = min(nrows, r_lo + per_slice);
    for (int r0 = r_lo; r0 < r_hi; r0 += kGramRows) {
        const int nr = min(kGramRows, r_hi - r0);
        __syncthreads();
        for (int e = tid; e < nr * R; e += blockDim.x) {
            const int r = e / R, c = e - r * R;
            const double *row = base + (size_t)(r0 + r) * ldy;
            if (c == 0) nsm[r * W] = row[0];
            else {
                const double yp = row[2 * c - 1], ym = row[2 * c];
                nsm[r * W + c] = yp - ym;
                nsm[r * W + R + c - 1] = (yp - row[0]) + (ym - row[0]);
            }
        }
        __syncthreads();
        for (int q = 0; q < np; ++q) {
            double s = acc[q];
            const int i = pi[q], j = pj[q];
            for (int r = 0; r < nr; ++r) s = fma(nsm[r * W + i], nsm[r * W + j], s);
            acc[q] = s;
        }
    }
    double *o = part + ((size_t)f * kProbeSplit + blockIdx.y) * npairs;
    for (int q = 0; q < np; ++q) o[tid + q * blockDim.x] = acc[q];
}

// out[f] = M (row-major, symmetric) followed by t, from the kProbeSplit partial sums of every pair, added in slice order
__global__ void __launch_bounds__(256) normal_probes_sum_kernel(const double *__restrict__ part, int n, int nfit, double *__restrict__ out) {
    pdl_sync();
    const int f = blockIdx.x, R = n + 1, npairs = R * (R + 1) / 2 + n;
    double *o = out + (size_t)f * (R * R + n);
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        double v = 0.0;
        for (int s = 0; s < kProbeSplit; ++s) v += part[((size_t)f * kProbeSplit + s) * npairs + p];
        if (p < R * (R + 1) / 2) {
            int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > p) --i;
            while ((i + 1) * (i + 2) / 2 <= p) ++i;
            const int j = p - i * (i + 1) / 2;
            o[i * R + j] = v;
            o[j * R + i] = v;
        } else
            o[R * R + p - R * (R + 1) / 2] = v;
    }
}

// normal_probes_kernel with the columns of linear parameters synthesised (struct NormalLinArgs)
constexpr int kLinMaxTerms = 16;  // static columns one synthesised column is a combination of (redshift classes, or powers of v)

__global__ void __launch_bounds__(256) normal_probes_lin_kernel(NormalLinArgs a) {
    pdl_sync();
    extern __shared__ double lsm[];  // [kGramRows][R + n] as in normal_probes_kernel, then per-column metadata
    const int f = blockIdx.x, tid = threadIdx.x, n = a.n, R = n + 1, W = R + n, G = 2 * a.n_nl + 1;
    // column i of a linear parameter: scale_i * sum_q cw[i][q] * WA[row][kleft + cbase_i + q * cstride]
    double *scale = lsm + kGramRows * W;                       // [n]   2 h_i for linear columns
    double *cw = scale + n;                                    // [n][kLinMaxTerms]
    int *kind = reinterpret_cast<int *>(cw + n * kLinMaxTerms);  // [n]   >= 0: linear, -1 - q: q-th non-linear probe
    int *cbase = kind + n, *cnum = cbase + n;                  // [n] each
    const double *c = a.centres + (size_t)f * a.npar, *h = a.steps + (size_t)f * a.npar;
    const int cstride = a.fold_case == 2 ? a.nrt : a.nb1;
    if (tid == 0) {
        const double gb = c[a.idx_gamma_bias];
        double eb[kLinMaxTerms];
        for (int k = 0; k < a.ncls; ++k) {
            const double x = (1.0 + a.zvals[a.ncls == 1 ? a.zc0 : k]) / (1.0 + a.zref);
            eb[k] = gb == 0.0 ? 1.0 : pow(x, gb);  // as prep_kernel (AbsCorrelationModel.cc:159-161)
        }
        // velocity shift of this fit's centre in units of r0 (fold_coef_kernel, case 2)
        const double v = a.fold_case == 2 ? c[a.idx_delta_v] * a.vfac[a.zc0] * a.inv_r0 : 0.0;
        int col = 0, q = 0;
        for (int j = 0; j < a.npar && col < n; ++j)
            if (h[j] != 0.0) {
                const int k = a.lin[j];
                if (k >= 0) {
                    kind[col] = k;
                    scale[col] = 2.0 * h[j];
                    if (a.fold_case == 2) {
                        // coefficient (zi, ip, t) feeds the rows (zi, m <= p, t) with C(p, m) v^(p - m)
                        const int t = k % a.nrt, ip = (k / a.nrt) % a.n_rp, zi = k / (a.nrt * a.n_rp), p = a.rp_min + ip * a.rp_step;
                        cbase[col] = zi * (a.pmax + 1) * a.nrt + t;
                        cnum[col] = p + 1;
                        for (int mm = 0; mm <= p; ++mm) {
                            double binom = 1.0, vp = 1.0;
                            for (int u = 0; u < mm; ++u) binom = binom * (p - u) / (u + 1);
                            for (int u = 0; u < p - mm; ++u) vp *= v;
                            cw[col * kLinMaxTerms + mm] = eb[0] * binom * vp;
                        }
                    } else {
                        cbase[col] = k;
                        cnum[col] = a.ncls;
                        for (int u = 0; u < a.ncls; ++u) cw[col * kLinMaxTerms + u] = eb[u];
                    }
                } else { kind[col] = -1 - q; scale[col] = 0.0; cnum[col] = 0; cbase[col] = 0; ++q; }
                ++col;
            }
    }
    const int npairs = R * (R + 1) / 2 + n;
    double acc[6];
    int pi[6], pj[6], np = 0;
    for (int p = tid; p < npairs && np < 6; p += blockDim.x) {
        int i, j;
        if (p < R * (R + 1) / 2) {
            i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > p) --i;
            while ((i + 1) * (i + 2) / 2 <= p) ++i;
            j = p - i * (i + 1) / 2;
        } else {
            i = R + (p - R * (R + 1) / 2);
            j = 0;
        }
        pi[np] = i; pj[np] = j; acc[np] = 0.0; ++np;
    }
    const double *base = a.Y + (size_t)f * G;
    const int per_slice = (a.nrows + kProbeSplit - 1) / kProbeSplit, r_lo = min(a.nrows, (int)blockIdx.y * per_slice),
              r_hi = min(a.nrows, r_lo + per_slice);
    for (int r0 = r_lo; r0 < r_hi; r0 += kGramRows) {
        const int nr = min(kGramRows, r_hi - r0);
        __syncthreads();
        for (int e = tid; e < nr * R; e += blockDim.x) {
            const int r = e / R, col = e - r * R;
            const double *row = base + (size_t)(r0 + r) * a.ldy;
            if (col == 0) lsm[r * W] = row[0];
            else {
                const int kd = kind[col - 1];
                if (kd >= 0) {
                    const double *wa = a.WA + (size_t)(r0 + r) * a.lda + a.kleft + cbase[col - 1];
                    const double *w = cw + (col - 1) * kLinMaxTerms;
                    double v = 0.0;
                    for (int u = 0; u < cnum[col - 1]; ++u) v = fma(w[u], wa[u * cstride], v);
                    lsm[r * W + col] = scale[col - 1] * v;
                    lsm[r * W + R + col - 1] = 0.0;
                } else {
                    const int q = -1 - kd;
                    const double yp = row[2 * q + 1], ym = row[2 * q + 2];
                    lsm[r * W + col] = yp - ym;
                    lsm[r * W + R + col - 1] = (yp - row[0]) + (ym - row[0]);
                }
            }
        }
        __syncthreads();
        for (int q = 0; q < np; ++q) {
            double s = acc[q];
            const int i = pi[q], j = pj[q];
            for (int r = 0; r < nr; ++r) s = fma(lsm[r * W + i], lsm[r * W + j], s);
            acc[q] = s;
        }
    }
    double *o = a.part + ((size_t)f * kProbeSplit + blockIdx.y) * npairs;
    for (int q = 0; q < np; ++q) o[tid + q * blockDim.x] = acc[q];
}

size_t normal_probes_part_doubles(int n, int nfit) { return (size_t)nfit * kProbeSplit * ((size_t)(n + 1) * (n + 2) / 2 + n); }

void launch_normal_probes_lin(cudaStream_t s, const NormalLinArgs &a) {
    const size_t smem = ((size_t)kGramRows * (2 * a.n + 1) + a.n + (size_t)a.n * kLinMaxTerms) * sizeof(double) + (size_t)3 * a.n * sizeof(int);
    static SmemOptIn opt;
    opt.ensure(normal_probes_lin_kernel, smem);
    Launch(dim3(a.nfit, kProbeSplit), 256, smem, s)(normal_probes_lin_kernel, a);
    Launch(a.nfit, 256, 0, s)(normal_probes_sum_kernel, a.part, a.n, a.nfit, a.out);
}

void launch_normal_probes(cudaStream_t s, const double *Y, int ldy, int nrows, int n, int nfit, double *out, double *part) {
    Launch(dim3(nfit, kProbeSplit), 256, (size_t)kGramRows * (2 * n + 1) * sizeof(double), s)(normal_probes_kernel, Y, ldy, nrows, n, nfit, part);
    Launch(nfit, 256, 0, s)(normal_probes_sum_kernel, part, n, nfit, out);
}

void launch_gram(cudaStream_t s, const double *Y, int ldy, int nrows, int G, int nfit, double *out) {
    Launch(nfit, 256, (size_t)kGramRows * G * sizeof(double), s)(gram_kernel, Y, ldy, nrows, G, nfit, out);
}

__global__ void expand_probes_kernel(const double *__restrict__ centres, const double *__restrict__ steps, int nfit, int npar,
                                     int nprobe, double *__restrict__ params, const int *__restrict__ skip) {
    pdl_sync();
    const int G = 2 * nprobe + 1;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= nfit * G) return;
    const int f = col / G, g = col - f * G;
    const double *c = centres + (size_t)f * npar, *h = steps + (size_t)f * npar;
    double *out = params + (size_t)col * npar;
    // the probed parameter of column g is the ((g-1)/2)-th parameter with a non-zero step
    int want = g == 0 ? -1 : (g - 1) / 2, seen = 0;
    const double sgn = (g & 1) ? 1.0 : -1.0;
    for (int j = 0; j < npar; ++j) {
        double v = c[j];
        if (h[j] != 0.0 && !(skip && skip[j] >= 0)) {  // skip: parameters whose columns are synthesised (NormalLinArgs::lin)
            if (seen == want) v += sgn * h[j];
            ++seen;
        }
        out[j] = v;
    }
}

__global__ void reduce_status_kernel(const int *__restrict__ status, int nfit, int G, int *__restrict__ status_fit) {
    pdl_sync();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfit) return;
    int s = 0;
    for (int g = 0; g < G; ++g) s |= status[(size_t)f * G + g];
    status_fit[f] = s;
}

void launch_expand_probes(cudaStream_t s, const double *centres, const double *steps, int nfit, int npar, int nprobe, double *params,
                          const int *skip) {
    const int total = nfit * (2 * nprobe + 1);
    Launch((total + 127) / 128, 128, 0, s)(expand_probes_kernel, centres, steps, nfit, npar, nprobe, params, skip);
}

void launch_reduce_status(cudaStream_t s, const int *status, int nfit, int G, int *status_fit) {
    Launch((nfit + 127) / 128, 128, 0, s)(reduce_status_kernel, status, nfit, G, status_fit);
}

}  // namespace bf
