// Device helpers shared by the cartesian-form model kernels (kernels_fast.cu) and the fused
// model + chi-square kernel (kernels_fused.cu): mbarrier / TMA bulk-copy wrappers, the per-vector state
// of the fast pass, and the basis-table correlation look-up.
#pragma once
#include "device_math.cuh"
#include "kernels.h"

namespace bf {

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
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

// per-vector state, loaded once (kept small: it lives in registers across the bin loop)
struct FastVec {
    double biasSq, ampl, dv, apar, aperp, iso, c1, c2;
};

__device__ __forceinline__ void load_fast_vec(const DevModel &m, const double *__restrict__ pT, int ldb, double ebeta_trigger,
                                              FastVec &v) {
    const Internal in = load_internal(m, pT, ldb);
    v.biasSq = in.biasSq;
    v.ampl = pT[(size_t)m.idx_bao * ldb];
    v.iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    v.apar = pT[(size_t)(m.idx_bao + 2) * ldb];
    v.aperp = (m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * v.apar
                                                 : pT[(size_t)(m.idx_bao + 3) * ldb];
    v.dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    double betaz = in.beta * ebeta_trigger;
    if (m.flags & BF_FLAG_CROSS_CORRELATION) {
        double beta2z = in.beta2 * ebeta_trigger;
        v.c1 = betaz + beta2z;
        v.c2 = betaz * beta2z;
    } else {
        v.c1 = 2.0 * betaz;
        v.c2 = betaz * betaz;
    }
}

// normalisation factors of the cross metal templates for redshift-evolution factors (eb, ebeta)
__device__ __forceinline__ void cross_metal_norms(const DevModel &m, const double *__restrict__ pT, int ldb, int c,
                                                  double eb, double ebeta, double &n0, double &n2, double &n4) {
    const Internal in = load_internal(m, pT, ldb);
    double ba, bia, bb, bib;
    metal_pair(m, in, true, pT, ldb, m.metal_p1[c], ba, bia);
    metal_pair(m, in, true, pT, ldb, m.metal_p2[c], bb, bib);
    norm_factors(bia * bib * eb, 0.5 * (ba + bb) * ebeta, ba * bb * ebeta * ebeta, n0, n2, n4);
}

// sum_l L_l(mu) sum_n coef_n X_{l,n}(r) from the shared-memory basis tables; musq = mu^2
__device__ __forceinline__ double basis_correlation(const DevModel &m, const double2 *__restrict__ tab, double r, double musq,
                                                    double c1, double c2) {
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
    const double L2 = (-1.0 + 3.0 * musq) * 0.5, L4 = (3.0 + musq * (-30.0 + 35.0 * musq)) * 0.125;
    const double2 *lo = tab + (size_t)j * 9, *hi = lo + 9;
    // sum_l L_l (X_l0 + c1 X_l1 + c2 X_l2) for the four spline quantities: 32 FMAs, no products of weights
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
    return w0 * ylo + w1 * yhi + w2 * clo + w3 * chi;
}


__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}

}  // namespace bf
