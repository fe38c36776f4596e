// Device-side building blocks shared by the model kernels.
#pragma once
#include "bf_internal.h"

namespace bf {

__device__ __forceinline__ double ipow(double x, int n) {
    // x^n for integer |n| <= 31 (checked when the model is created) by binary powering, written
    // without a data-dependent loop so that it stays a dozen instructions when inlined.
    // The reference calls std::pow with an integral exponent (baofit/BroadbandModel.cc:205-213);
    // results agree to ~1 ulp.
    int m = n < 0 ? -n : n;
    double r = (m & 1) ? x : 1.0;
    double p = x * x;
    if (m & 2) r *= p;
    p *= p;
    if (m & 4) r *= p;
    p *= p;
    if (m & 8) r *= p;
    p *= p;
    if (m & 16) r *= p;
    return n < 0 ? 1.0 / r : r;
}

__device__ __forceinline__ double legendre_p(int ell, double mu) {  // baofit/BroadbandModel.cc:171-195
    double m2 = mu * mu;
    switch (ell) {
    case 0: return 1.0;
    case 1: return mu;
    case 2: return (-1.0 + 3.0 * m2) * 0.5;
    case 3: return mu * (-3.0 + 5.0 * m2) * 0.5;
    case 4: return (3.0 + m2 * (-30.0 + 35.0 * m2)) * 0.125;
    case 5: return mu * (15.0 + m2 * (-70.0 + 63.0 * m2)) * 0.125;
    case 6: return (-5.0 + m2 * (105.0 + m2 * (-315.0 + 231.0 * m2))) * 0.0625;
    case 7: return mu * (-35.0 + m2 * (315.0 + m2 * (-693.0 + 429.0 * m2))) * 0.0625;
    default: return (35.0 + m2 * (-1260.0 + m2 * (6930.0 + m2 * (-12012.0 + 6435.0 * m2)))) * 0.0078125;
    }
}

// One broadband axis in Horner form. The reference multiplies every coefficient by
// pow(e > 0 ? x - 1 : x, e) (baofit/BroadbandModel.cc:209-213); summed over a run of exponents that is
// a polynomial in 1/x^step for the non-positive exponents and one in (x-1)^step for the positive ones,
// so an axis costs one fma per term and at most one division per evaluation point instead of a power
// per term. AxisPlan (host-built, warp-uniform) holds the split of the run.
__device__ __forceinline__ double small_pow(double x, int k) {  // k >= 0, almost always 0 or 1
    return k == 0 ? 1.0 : k == 1 ? x : ipow(x, k);
}

struct AxisVal {
    double us, tail, vs, head;
    __device__ __forceinline__ void init(const AxisPlan &p, double x) {
        us = tail = vs = head = 1.0;
        if (p.nneg > 1 || p.elast > 0) {
            const double u = 1.0 / x;
            us = small_pow(u, p.step);
            tail = small_pow(u, p.elast);
        }
        if (p.n > p.nneg) {
            const double v = x - 1.0;
            vs = small_pow(v, p.step);
            head = small_pow(v, p.efirst);
        }
    }
};

// sum_k term(base + k*stride) * pow(e_k > 0 ? x - 1 : x, e_k) over the exponent run of the plan; the coefficient
// pointer is stepped (forwards over the first run, backwards over the second) so that a term costs no multiply
template <class Term>
__device__ __forceinline__ double axis_sum(const AxisPlan &p, const AxisVal &v, const double *base, size_t stride, Term term) {
    double acc = 0.0;
    if (p.nneg > 0) {
        const double *q = base;
        acc = term(q);
#pragma unroll 1
        for (int k = 1; k < p.nneg; ++k) {
            q += stride;
            acc = fma(acc, v.us, term(q));
        }
        if (p.elast > 0) acc *= v.tail;
    }
    if (p.n > p.nneg) {
        const double *q = base + (size_t)(p.n - 1) * stride;
        double pos = term(q);
#pragma unroll 1
        for (int k = p.n - 2; k >= p.nneg; --k) {
            q -= stride;
            pos = fma(pos, v.vs, term(q));
        }
        acc = fma(pos, v.head, acc);
    }
    return acc;
}

// Broadband polynomial, baofit/BroadbandModel.cc:197-226. pT = transposed parameters
// (pT[j*ldb + b]), already offset to column b.
__device__ __forceinline__ double broadband_eval(const DevBB &s, const double *__restrict__ pT, int ldb, double r,
                                                 double mu, double z) {
    const double rr = r * s.inv_r0;
    AxisVal vr, vp, vt;
    if (s.r_denom == 1) vr.init(s.plan_r, rr);
    vp.init(s.plan_rp, r * mu * s.inv_r0);
    if (s.plan_rt.n > 1 || s.rt_min != 0) vt.init(s.plan_rt, r * sqrt(1.0 - mu * mu) * s.inv_r0);
    else vt.us = vt.tail = vt.vs = vt.head = 1.0;
    const bool z_axis = !(s.z_min == 0 && s.z_max == 0);
    const double zz = z_axis ? (1.0 + z) * s.inv_1pz0 : 1.0;
    double xi = 0.0;
    const double *pc = pT + (size_t)s.idx_base * ldb;  // coefficient blocks, rT fastest
    const size_t st_t = (size_t)ldb, st_p = st_t * s.plan_rt.n, st_r = st_p * s.plan_rp.n, st_m = st_r * s.plan_r.n;
#pragma unroll 1
    for (int zi = s.z_min; zi <= s.z_max; zi += s.z_step) {
        const double zF = zi == 0 ? 1.0 : ipow(zz, zi);
#pragma unroll 1
        for (int mi = s.mu_min; mi <= s.mu_max; mi += s.mu_step, pc += st_m) {
            const double zmF = mi == 0 ? zF : zF * legendre_p(mi, mu);
            auto p_block = [&](const double *pr) {
                return axis_sum(s.plan_rp, vp, pr, st_p, [&](const double *pp) {
                    return axis_sum(s.plan_rt, vt, pp, st_t, [](const double *q) { return __ldg(q); });
                });
            };
            double sum_r;
            if (s.r_denom == 1) {
                sum_r = axis_sum(s.plan_r, vr, pc, st_r, p_block);
            } else {  // fractional r exponents (r-index denominator), one pow per term as in the reference
                sum_r = 0.0;
                int kr = 0;
#pragma unroll 1
                for (int ri = s.r_min; ri <= s.r_max; ri += s.r_step, ++kr)
                    sum_r = fma(p_block(pc + kr * st_r), pow(ri > 0 ? rr - 1.0 : rr, (double)ri / s.r_denom), sum_r);
            }
            xi = fma(sum_r, zmF, xi);
        }
    }
    return xi;
}

// natural cubic spline with end-point clamp on per-interval coefficients
__device__ __forceinline__ int spline_interval(const DevSplineSet &t, double x) {
    int j;
    if (t.uniform) {
        j = (int)((x - t.x0) * t.inv_dx);
        j = max(0, min(j, t.n - 2));
        // guard against rounding at knots
        if (x < t.x[j] && j > 0) --j;
        else if (x >= t.x[j + 1] && j < t.n - 2) ++j;
    } else {
        int lo = 0, hi = t.n - 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (t.x[mid] > x) hi = mid; else lo = mid;
        }
        j = lo;
    }
    return j;
}

__device__ __forceinline__ void spline3_eval(const DevSplineSet &t, const double *__restrict__ coef, double x,
                                             double out[3]) {
    if (x <= t.x0) { out[0] = t.ylo[0]; out[1] = t.ylo[1]; out[2] = t.ylo[2]; return; }
    if (x >= t.xlast) { out[0] = t.yhi[0]; out[1] = t.yhi[1]; out[2] = t.yhi[2]; return; }
    int j = spline_interval(t, x);
    double dx = x - t.x[j];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double4 c = *reinterpret_cast<const double4 *>(coef + ((size_t)l * (t.n - 1) + j) * 4);
        out[l] = c.x + dx * (c.y + dx * (c.z + dx * c.w));
    }
}

// Kaiser normalisation factors, baofit/AbsCorrelationModel.cc:196-204 / MetalCorrelationModel.cc:359-363
__device__ __forceinline__ void norm_factors(double biasSq, double betaAvg, double betaProd, double &n0, double &n2,
                                             double &n4) {
    n0 = biasSq * (1.0 + (2.0 / 3.0) * betaAvg + (1.0 / 5.0) * betaProd);
    n2 = biasSq * ((4.0 / 3.0) * betaAvg + (4.0 / 7.0) * betaProd);
    n4 = biasSq * betaProd * (8.0 / 35.0);
}

// Velocity shift, baofit/AbsCorrelationModel.cc:146-157 with dpi = dv * vfac(z)
__device__ __forceinline__ void velocity_shift(double dpi, double &r, double &mu) {
    double rnew = sqrt(r * r + 2.0 * r * mu * dpi + dpi * dpi);
    double munew = (r * mu + dpi) / rnew;
    r = rnew;
    mu = munew;
}

struct Internal {  // AbsCorrelationModel::_updateInternalParameters, AbsCorrelationModel.cc:132-144
    double beta, bias, bias2, beta2, biasSq, betaAvg, betaProd;
};

__device__ __forceinline__ Internal load_internal(const DevModel &m, const double *__restrict__ pT, int ldb) {
    Internal in;
    in.beta = pT[(size_t)m.idx_beta * ldb];
    in.bias = pT[(size_t)m.idx_bb * ldb] / (1.0 + in.beta);
    if (m.idx_beta_bias >= 0) in.bias = pT[(size_t)m.idx_beta_bias * ldb] / in.beta;
    in.bias2 = in.beta2 = 0.0;
    if (m.flags & BF_FLAG_CROSS_CORRELATION) {
        in.bias2 = pT[(size_t)m.idx_bias2 * ldb];
        in.beta2 = pT[(size_t)m.idx_bb2 * ldb] / in.bias2;
        in.betaAvg = (in.beta + in.beta2) / 2.0;
        in.betaProd = in.beta * in.beta2;
        in.biasSq = in.bias * in.bias2;
    } else {
        in.betaAvg = in.beta;
        in.betaProd = in.beta * in.beta;
        in.biasSq = in.bias * in.bias;
    }
    return in;
}

// Catmull-Rom weights (tensor-product form of the bicubic patch with centred differences)
__device__ __forceinline__ void catmull_weights(double t, double w[4]) {
    double t2 = t * t, t3 = t2 * t;
    w[0] = 0.5 * (-t + 2.0 * t2 - t3);
    w[1] = 0.5 * (2.0 - 5.0 * t2 + 3.0 * t3);
    w[2] = 0.5 * (t + 4.0 * t2 - 3.0 * t3);
    w[3] = 0.5 * (-t2 + t3);
}
__device__ __forceinline__ int wrap_index(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// Metal correlations, baofit/MetalCorrelationModel.cc:235-351.
// Everything that depends only on the parameter vector (bias products, beta averages of each line
// combination) is gathered once per thread into MetalVec; the per-bin part is the redshift
// evolution, the normalisation factors and the template look-up. No dynamically indexed local
// arrays (those would live in local memory).
constexpr int kMaxCombos = 5;            // bicubic cross templates: 4 lines + CIV
constexpr int kMaxCrossTemplates = 15;   // x 3 multipoles

struct MetalVec {
    double biasSq[kMaxCombos], betaAvg[kMaxCombos], betaProd[kMaxCombos];
    double toy_ampl[4], toy_width0;
};

__device__ __forceinline__ void metal_pair(const DevModel &m, const Internal &in, bool cross,
                                           const double *__restrict__ pT, int ldb, int a, double &beta, double &bias) {
    if (a == 0) {  // :247-248, :282-283
        beta = cross ? in.beta2 : in.beta;
        bias = cross ? in.bias2 : in.bias;
    } else {
        beta = pT[(size_t)(m.idx_metal + 2 * (a - 1)) * ldb];
        bias = pT[(size_t)(m.idx_metal + 2 * (a - 1) + 1) * ldb];
    }
}

__device__ __forceinline__ void metal_load(const DevModel &m, const Internal &in, const double *__restrict__ pT, int ldb,
                                           MetalVec &mv) {
    if (m.metal_kind == BF_METAL_TOY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mv.toy_ampl[i] = pT[(size_t)(m.idx_metal + i) * ldb];
        mv.toy_width0 = pT[(size_t)(m.idx_metal + 4) * ldb];
    } else if (m.metal_kind == BF_METAL_CROSS_BICUBIC) {
#pragma unroll
        for (int c = 0; c < kMaxCombos; ++c) {
            mv.biasSq[c] = mv.betaAvg[c] = mv.betaProd[c] = 0.0;
            if (c < m.metal_ncomb) {
                double ba, bia, bb, bib;
                metal_pair(m, in, true, pT, ldb, m.metal_p1[c], ba, bia);
                metal_pair(m, in, true, pT, ldb, m.metal_p2[c], bb, bib);
                mv.biasSq[c] = bia * bib;
                mv.betaAvg[c] = 0.5 * (ba + bb);
                mv.betaProd[c] = ba * bb;
            }
        }
    }
}

template <int NT>
__device__ __forceinline__ void bicubic_accumulate(const double *__restrict__ grid, int nx, int ny, int ix, int iy,
                                                   const double (&wx)[4], const double (&wy)[4], double (&acc)[NT]) {
    int cx[4], cy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cx[j] = wrap_index(ix - 1 + j, nx) * NT;
        cy[j] = wrap_index(iy - 1 + j, ny) * nx * NT;
    }
#pragma unroll
    for (int jy = 0; jy < 4; ++jy)
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
            const double *cell = grid + cy[jy] + cx[jx];
            const double w = wx[jx] * wy[jy];
            if (NT % 2 == 0) {
#pragma unroll
                for (int t = 0; t < NT; t += 2) {
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(cell + t));
                    acc[t] = fma(w, v.x, acc[t]);
                    acc[t + 1] = fma(w, v.y, acc[t + 1]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < NT; ++t) acc[t] = fma(w, __ldg(cell + t), acc[t]);
            }
        }
}

// S = per-class evolution factors S[(c*3+k)*ldb] for column b; zc_metal[c*stride + bin] = class ids.
__device__ __forceinline__ double metal_eval(const DevModel &m, const Internal &in, const MetalVec &mv,
                                             const double *__restrict__ pT, int ldb, const double *__restrict__ S,
                                             const int *__restrict__ zc_metal, int zc_stride, int bin, int grid_index,
                                             double r, double mu) {
    if (m.metal_kind == BF_METAL_TOY) {
        const double rpar0[4] = {59.533, 111.344, 134.998, 174.525};
        const double sigpar[4] = {0.0, 4.88626, 5.489, 8.48942};
        const double sigperp[4] = {5.55309, 4.13032, 6.30484, 7.85636};
        double rperp = r * sqrt(1.0 - mu * mu), rpar = fabs(r * mu), xi = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double sp = i == 0 ? mv.toy_width0 : sigpar[i];
            double ep = (rpar - rpar0[i]) / sp, et = (rperp - 2.0) / sigperp[i];
            if (fabs(ep) < 5.0 && fabs(et) < 10.0) xi += 1e-4 * mv.toy_ampl[i] * exp(-ep * ep - et);
        }
        return xi;
    }
    double xi = 0.0;
    if (m.metal_kind == BF_METAL_CROSS_BICUBIC) {
        // bicubic templates: the 16 patch weights are shared by all templates
        double rperp = r * sqrt(1.0 - mu * mu), rpar = r * mu;
        rperp = fmin(fmax(rperp, m.metal_x0), m.metal_xmax);
        rpar = fmin(fmax(rpar, m.metal_y0), m.metal_ymax);
        double fx = (rperp - m.metal_x0) * m.metal_inv_dx, fy = (rpar - m.metal_y0) * m.metal_inv_dy;
        double flx = floor(fx), fly = floor(fy);
        int ix = (int)flx, iy = (int)fly;
        double wx[4], wy[4];
        catmull_weights(fx - flx, wx);
        catmull_weights(fy - fly, wy);
        double acc[kMaxCrossTemplates];
#pragma unroll
        for (int t = 0; t < kMaxCrossTemplates; ++t) acc[t] = 0.0;
        if (m.metal_ncomb == 4) {
            double (&a12)[12] = reinterpret_cast<double (&)[12]>(acc);
            bicubic_accumulate<12>(m.metal_cross, m.metal_nx, m.metal_ny, ix, iy, wx, wy, a12);
        } else if (m.metal_ncomb == 5) {
            bicubic_accumulate<15>(m.metal_cross, m.metal_nx, m.metal_ny, ix, iy, wx, wy, acc);
        } else {
            const int nt = m.metal_ncomb * 3;
            for (int jy = 0; jy < 4; ++jy) {
                const double *row = m.metal_cross + (size_t)wrap_index(iy - 1 + jy, m.metal_ny) * m.metal_nx * nt;
                for (int jx = 0; jx < 4; ++jx) {
                    const double *cell = row + (size_t)wrap_index(ix - 1 + jx, m.metal_nx) * nt;
                    double w = wx[jx] * wy[jy];
#pragma unroll
                    for (int t = 0; t < kMaxCrossTemplates; ++t)
                        if (t < nt) acc[t] = fma(w, __ldg(cell + t), acc[t]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxCombos; ++c) {
            if (c < m.metal_ncomb) {
                int zc = zc_metal[c * zc_stride + bin];
                double eb = S[(size_t)(zc * 3 + 0) * ldb], ebeta = S[(size_t)(zc * 3 + 1) * ldb];
                double n0, n2, n4;
                norm_factors(mv.biasSq[c] * eb, mv.betaAvg[c] * ebeta, mv.betaProd[c] * ebeta * ebeta, n0, n2, n4);
                xi += n0 * acc[3 * c] + n2 * acc[3 * c + 1] + n4 * acc[3 * c + 2];
            }
        }
        return xi;
    }
    // auto-correlation templates tabulated per grid bin (up to 20 combinations)
    for (int c = 0; c < m.metal_ncomb; ++c) {
        double ba, bia, bb, bib;
        metal_pair(m, in, false, pT, ldb, m.metal_p1[c], ba, bia);
        metal_pair(m, in, false, pT, ldb, m.metal_p2[c], bb, bib);
        int zc = zc_metal[c * zc_stride + bin];
        double eb = S[(size_t)(zc * 3 + 0) * ldb], ebeta = S[(size_t)(zc * 3 + 1) * ldb];
        double n0, n2, n4;
        norm_factors(bia * bib * eb, 0.5 * (ba + bb) * ebeta, ba * bb * ebeta * ebeta, n0, n2, n4);
        const double *t = m.metal_auto + (size_t)(3 * c) * m.metal_ngrid + grid_index;
        xi += n0 * t[0] + n2 * t[m.metal_ngrid] + n4 * t[2 * (size_t)m.metal_ngrid];
    }
    return xi;
}

}  // namespace bf
