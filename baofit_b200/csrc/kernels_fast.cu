// Hot kernels of the shared-transform path, written in "cartesian" form.
//
// Everything that depends only on the bin is tabulated once per (model, data) plan on the host
// (PointTab): r_par, r_perp, the r_T- and z-axis factors of every broadband polynomial, and the
// metal templates already collapsed along r_perp. What is left per (vector b, bin) is what really
// depends on b: the velocity shift is an addition on r_par (r_perp is invariant under it), the
// anisotropic dilation needs one rsqrt, two spline look-ups in the shared-memory basis tables,
// a 4-row x 12-template metal patch, and the r_P powers of the broadband polynomial.
// Lanes = consecutive b (coalesced bin-major stores); bin quantities are warp-uniform loads.
#include <cstdlib>
#include "fast_device.cuh"

namespace bf {

namespace {

// Broadband polynomial with the bin-static r_T and z factors read from the plan tables.
// rrP = r_par/r0 of this (b, bin); rr and mu are only touched if their axes are non-trivial.
__device__ __forceinline__ double broadband_fast(const DevBB &s, const double *__restrict__ pT, int ldb,
                                                 const double *__restrict__ rt_pow, int n_rt,
                                                 const double *__restrict__ z_pow, double rr, double mu, double rrP) {
    const bool r_axis = !(s.r_min == 0 && s.r_max == 0);
    AxisVal vr, vp;
    vr.us = vr.tail = vr.vs = vr.head = 1.0;
    if (r_axis && s.r_denom == 1) vr.init(s.plan_r, rr);
    vp.init(s.plan_rp, rrP);
    double xi = 0.0;
    const double *pc = pT + (size_t)s.idx_base * ldb;  // coefficient blocks, rT fastest
    const size_t st_p = (size_t)ldb * n_rt, st_r = st_p * s.plan_rp.n, st_m = st_r * s.plan_r.n;
    int kz = 0;
#pragma unroll 1
    for (int zi = s.z_min; zi <= s.z_max; zi += s.z_step, ++kz) {
        const double zF = z_pow[kz];
#pragma unroll 1
        for (int mi = s.mu_min; mi <= s.mu_max; mi += s.mu_step, pc += st_m) {
            const double zmF = mi == 0 ? zF : zF * legendre_p(mi, mu);
            auto p_block = [&](const double *pr) {
                return axis_sum(s.plan_rp, vp, pr, st_p, [&](const double *pp) {
                    double acc = 0.0;
#pragma unroll 5
                    for (int kt = 0; kt < n_rt; ++kt) acc = fma(__ldg(pp + (size_t)kt * ldb), rt_pow[kt], acc);
                    return acc;
                });
            };
            double sum_r;
            if (s.r_denom == 1) {
                sum_r = axis_sum(s.plan_r, vr, pc, st_r, p_block);
            } else {
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

__device__ __forceinline__ double broadband_pair(const DevModel &m, const PointTab &pt, int slot_add, int slot_mul, int i,
                                                 const double *__restrict__ pT, int ldb, double eb, double xi, double rprime,
                                                 double muprime, double rpar) {
    if (m.bb[slot_mul].enabled) {
        const DevBB &s = m.bb[slot_mul];
        xi *= 1.0 + broadband_fast(s, pT, ldb, pt.bb_rt[slot_mul] + (size_t)i * pt.n_rt[slot_mul], pt.n_rt[slot_mul],
                                   pt.bb_z[slot_mul] + (size_t)i * pt.n_z[slot_mul], rprime * s.inv_r0, muprime, rpar * s.inv_r0);
    }
    if (m.bb[slot_add].enabled) {
        const DevBB &s = m.bb[slot_add];
        xi += eb * broadband_fast(s, pT, ldb, pt.bb_rt[slot_add] + (size_t)i * pt.n_rt[slot_add], pt.n_rt[slot_add],
                                  pt.bb_z[slot_add] + (size_t)i * pt.n_z[slot_add], rprime * s.inv_r0, muprime, rpar * s.inv_r0);
    }
    return xi;
}

// Cross metals from the x-collapsed templates MX[i][row][t]: 4 rows x NT templates, then the
// normalisation factors n[t] of this vector. Returns sum_t n[t] * T_t(r_perp, r_par).
template <int NT>
__device__ __forceinline__ double metal_rows_apply(const double *__restrict__ mx_point, int ny, int iy, const double (&wy)[4],
                                                   const double (&n)[kMaxCrossTemplates]) {
    double acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.0;
#pragma unroll
    for (int jy = 0; jy < 4; ++jy) {
        const double *row = mx_point + (size_t)wrap_index(iy - 1 + jy, ny) * NT;
        if (NT % 2 == 0) {
#pragma unroll
            for (int t = 0; t < NT; t += 2) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(row + t));
                acc[t] = fma(wy[jy], v.x, acc[t]);
                acc[t + 1] = fma(wy[jy], v.y, acc[t + 1]);
            }
        } else {
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t] = fma(wy[jy], __ldg(row + t), acc[t]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) s = fma(n[t], acc[t], s);
    return s;
}

}  // namespace

// mode 0: data bins without distortion matrix (full model incl. post broadband, residual)
// mode 1: grid bins for the distortion matrix (undistorted xi_u, range zeroing, pre-matrix broadband)
// kLean: no radiation term, metals none or bicubic cross templates (the BASELINE configurations);
// the full variant also carries the toy / auto-table metals and the radiation model.
// kPart: 0 = everything in one pass; 1 = cosmological part only (peak + smooth, status);
// 2 = the rest (metals, radiation, broadband, residual) added to what part 1 stored. Splitting
// halves the live registers of each kernel, which is what bounds the occupancy of this
// latency-bound code.
// kRS: the r-space model (BaoCorrelationModel.cc:90-179) through the same pass. Its Kaiser normalisation factors
// are polynomials in (beta_avg, beta_prod) = (c1/2, c2), so sum_l L_l n_l xi_l(r) has exactly the basis-table form
// with tables built once from the tabulated multipoles (bf_model_create); what differs is the per-bin redshift
// evolution of beta and the radiation term, which r-space models always carry.
template <bool kLean, int kPart, bool kRS = false>
__global__ void __launch_bounds__(kPart == 1 ? 512 : 256, 2) kspace_fast_eval_kernel(FastArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tab = reinterpret_cast<double2 *>(smem_raw);
    __shared__ __align__(8) unsigned long long bar;
    const DevModel &m = a.m;
    const PointTab &pt = a.pt;
    const unsigned tab_bytes = 2u * m.nr * 9u * (unsigned)sizeof(double2);
    if (kPart != 2 && threadIdx.x == 0) mbar_init(&bar, 1);
    if (kPart != 2) __syncthreads();
    if (kPart != 2 && threadIdx.x == 0) {
        mbar_expect_tx(&bar, tab_bytes);
        const unsigned chunk = 32768;
        for (unsigned off = 0; off < tab_bytes; off += chunk)
            tma_bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char *>(a.tabs) + off, min(chunk, tab_bytes - off), &bar);
    }
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = b < a.B;
    if (!active) b = a.B - 1;  // keep the whole CTA alive for the barrier; results are not stored
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    FastVec v;
    load_fast_vec(m, pT, ldb, (kRS && a.nz != 1) ? 1.0 : S[(size_t)(a.zc_trigger * 3 + 1) * ldb], v);
    const bool aniso = m.flags & BF_FLAG_ANISOTROPIC, decoupled = m.flags & BF_FLAG_DECOUPLED;
    const bool cross_metal = m.metal_kind == BF_METAL_CROSS_BICUBIC;
    const double dmin2 = m.dilmin * m.dilmin, dmax2 = m.dilmax * m.dilmax;
    // with a single redshift class the evolution factors and metal norms are per-vector constants
    double eb1 = 1.0, es1 = 1.0, dpi1 = 0.0;
    double mn[kMaxCrossTemplates];
#pragma unroll
    for (int t = 0; t < kMaxCrossTemplates; ++t) mn[t] = 0.0;
    if (a.nz == 1) {
        eb1 = S[0];
        es1 = S[(size_t)2 * ldb];
        dpi1 = v.dv * a.vfac[0];
        double ebeta = S[(size_t)ldb];
        if (kPart != 1) {
#pragma unroll
            for (int c = 0; c < kMaxCombos; ++c)
                if (cross_metal && c < m.metal_ncomb) cross_metal_norms(m, pT, ldb, c, eb1, ebeta, mn[3 * c], mn[3 * c + 1], mn[3 * c + 2]);
        }
    }
    // r-space: c1, c2 follow the redshift class of the bin; v.c1, v.c2 were loaded for ebeta = 1 if there are several
    double rad0 = 0.0;
    if (kRS && m.idx_rad >= 0) rad0 = pT[(size_t)m.idx_rad * ldb];
    if (kPart != 2) mbar_wait(&bar, 0);
    const double2 *tab_pk = tab, *tab_nw = tab + (size_t)m.nr * 9;
    int st_bits = 0;
    const int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, pt.npts_pad);
    // software pipelining of the bin-static loads: the coordinates of bin i+1 are fetched, and the
    // metal rows it will need are prefetched into L1, while bin i is being computed
    double nx_rperp = pt.rperp0[i0], nx_rpar0 = pt.rpar0[i0];
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        const double rperp = nx_rperp, rpar0 = nx_rpar0;
        if (i + 1 < i1) {
            nx_rperp = pt.rperp0[i + 1];
            nx_rpar0 = pt.rpar0[i + 1];
            if (kPart != 1 && cross_metal && a.nz == 1 && i + 1 < pt.npts) {
                double y = fmin(fmax(nx_rpar0 + dpi1, m.metal_y0), m.metal_ymax);
                int iy = (int)floor((y - m.metal_y0) * m.metal_inv_dy);
                const double *row = pt.mx + ((size_t)(i + 1) * pt.mx_ny + max(iy - 1, 0)) * pt.mx_nt;
#pragma unroll
                for (int q = 0; q < 4; ++q) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + q * 16));
            }
        }
        if (i < pt.npts) {
            double eb = eb1, es = es1, dpi = dpi1, c1 = v.c1, c2 = v.c2;
            if (a.nz != 1) {
                int zc = pt.zc[i];
                eb = S[(size_t)(zc * 3 + 0) * ldb];
                es = S[(size_t)(zc * 3 + 2) * ldb];
                dpi = v.dv * a.vfac[zc];
                if (kRS) {
                    const double ebeta = S[(size_t)(zc * 3 + 1) * ldb];
                    c1 *= ebeta;
                    c2 *= ebeta * ebeta;
                }
            }
            const double rpar = rpar0 + dpi;  // the velocity shift only moves r_parallel
            const double rp2 = rpar * rpar, rt2 = rperp * rperp, r2 = rp2 + rt2;
            const double rs = rsqrt(r2);
            const double rprime = r2 * rs, muprime = rpar * rs, mu2 = muprime * muprime;
            double rBAO, muBAO2, rBAO2;
            if (aniso) {
                const double apar = v.apar * es, aperp = v.aperp * es;
                const double ap2 = apar * apar * rp2;
                rBAO2 = ap2 + aperp * aperp * rt2;
                const double rsb = rsqrt(rBAO2);
                rBAO = rBAO2 * rsb;
                muBAO2 = ap2 * rsb * rsb;
            } else {
                const double sc = v.iso * es;
                rBAO = rprime * sc;
                rBAO2 = rBAO * rBAO;
                muBAO2 = mu2;
            }
            bool zero = false;
            if (a.mode == 1) {
                zero = (rprime < m.rlo || rprime > m.rhi) || (rBAO < m.rlo || rBAO > m.rhi);
            } else {
                if (rBAO2 < dmin2 * r2) st_bits |= BF_STATUS_DILATION_MIN;
                else if (rBAO2 > dmax2 * r2) st_bits |= BF_STATUS_DILATION_MAX;
            }
            if (!zero) {
                double xi;
                if (kPart != 2) {
                    const double peak = basis_correlation(m, tab_pk, rBAO, muBAO2, c1, c2);
                    const double smooth = decoupled ? basis_correlation(m, tab_nw, rprime, mu2, c1, c2)
                                                    : basis_correlation(m, tab_nw, rBAO, muBAO2, c1, c2);
                    xi = v.biasSq * eb * (v.ampl * peak + smooth);
                    if (kRS && rad0 > 0.0 && rprime > 0.0) {  // BaoCorrelationModel.cc:165-176
                        const double rad1 = pT[(size_t)(m.idx_rad + 1) * ldb], rad2 = pT[(size_t)(m.idx_rad + 2) * ldb];
                        const double rad3 = pT[(size_t)(m.idx_rad + 3) * ldb];
                        double rd = rad0 / r2;
                        rd *= exp(-rprime / rad2);
                        rd *= (1.0 - rad1 * (1.0 - mu2));
                        rd *= exp(-(rprime * (1.0 - muprime) / (1.0 + pt.z[i])) / rad3);
                        xi += rd;
                    }
                } else {
                    xi = a.out[(size_t)i * a.ldo + b];
                }
                if (kPart == 1) {
                    val = xi;
                } else {
                if (cross_metal) {
                    // r_perp (hence the x stencil) is bin-static and already folded into pt.mx
                    double y = fmin(fmax(rpar, m.metal_y0), m.metal_ymax);
                    double fy = (y - m.metal_y0) * m.metal_inv_dy, fly = floor(fy);
                    double wy[4];
                    catmull_weights(fy - fly, wy);
                    const double *mxp = pt.mx + (size_t)i * pt.mx_ny * pt.mx_nt;
                    if (a.nz != 1) {
                        // several redshift classes: the normalisation factors depend on the bin
#pragma unroll
                        for (int c = 0; c < kMaxCombos; ++c)
                            if (c < m.metal_ncomb) {
                                int zcm = pt.zc_metal[c * pt.zc_stride + i];
                                cross_metal_norms(m, pT, ldb, c, S[(size_t)(zcm * 3 + 0) * ldb], S[(size_t)(zcm * 3 + 1) * ldb],
                                                  mn[3 * c], mn[3 * c + 1], mn[3 * c + 2]);
                            }
                    }
                    if (pt.mx_nt == 12) xi += metal_rows_apply<12>(mxp, pt.mx_ny, (int)fly, wy, mn);
                    else xi += metal_rows_apply<15>(mxp, pt.mx_ny, (int)fly, wy, mn);
                } else if (!kLean && m.metal_kind != BF_METAL_NONE) {
                    const Internal in = load_internal(m, pT, ldb);
                    MetalVec mv;
                    metal_load(m, in, pT, ldb, mv);
                    xi += metal_eval(m, in, mv, pT, ldb, S, pt.zc_metal, pt.zc_stride, i, pt.gidx ? pt.gidx[i] : i, rprime, muprime);
                }
                if (!kLean && (m.flags & BF_FLAG_RADIATION_MODEL) && rprime > 0.0) {  // BaoKSpaceCorrelationModel.cc:441-455
                    const double rad0 = pT[(size_t)m.idx_rad * ldb], rad1 = pT[(size_t)(m.idx_rad + 1) * ldb];
                    const double rad2 = pT[(size_t)(m.idx_rad + 2) * ldb], rad3 = pT[(size_t)(m.idx_rad + 3) * ldb];
                    double rd = rad0 / r2;
                    if (rad1 > 0.0) rd *= (1.0 - rad1 * (1.0 - mu2));
                    if (rad2 > 0.0) rd *= exp(-rprime / rad2);
                    if (rad3 > 0.0) rd *= exp(-(rprime * (1.0 - muprime) / (1.0 + pt.z[i])) / rad3);
                    xi += rd;
                }
                if (a.mode == 1) xi = broadband_pair(m, pt, BF_BB_DM_DIST_ADD, BF_BB_DM_DIST_MUL, i, pT, ldb, eb, xi, rprime, muprime, rpar);
                else xi = broadband_pair(m, pt, BF_BB_DIST_ADD, BF_BB_DIST_MUL, i, pT, ldb, eb, xi, rprime, muprime, rpar);
                val = xi;
                }
            }
            if (a.mode == 0 && kPart != 1) {
                if (st_bits) val = nan("");
                if (a.dov) val -= a.dov[(size_t)b * pt.npts + i];
                else if (a.dsub) val -= a.dsub[i];
            }
            if (a.mode == 0 && kPart == 1 && st_bits) val = nan("");
            if (a.mode == 0 && kPart == 1 && a.sub_dov_in_part1) val -= a.dov[(size_t)b * pt.npts + i];
        }
        if (active) a.out[(size_t)i * a.ldo + b] = val;
    }
    if (kPart != 2 && active && st_bits) atomicOr(a.status + b, st_bits);
}

// Specialised second pass for the common case "bicubic cross metals, one redshift class, no
// broadband in this pass": out[i][b] += sum_row wy_row(b,i) * dot(n(b), MX[i][row][:]).
// The normalisation factors n(b) are per-vector constants, so each row is a 12-term dot product;
// two bins are processed per iteration to double the loads in flight (this pass is latency bound).
// The four template rows a (vector, bin) pair needs depend on the vector only through its velocity
// shift, which moves r_par by at most a cell or two: for every bin the rows of all 256 vectors of
// the CTA lie in one window of kMetalWin consecutive rows of MX[i], 768 contiguous bytes. The
// window of bin i + kMetalStages is fetched by one TMA bulk copy (cp.async.bulk + mbarrier) while
// bin i is being computed, so the pass reads shared memory instead of waiting on L1/L2 (it was
// long-scoreboard bound). Bins whose window would wrap or exceed kMetalWin rows take the global path.
constexpr int kMetalWin = 8, kMetalStages = 8;


template <int NT>
__global__ void __launch_bounds__(256, 3) metal_add_kernel(FastArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    __shared__ __align__(128) double win[kMetalStages][kMetalWin * NT];
    __shared__ __align__(8) unsigned long long full_bar[kMetalStages], empty_bar[kMetalStages];
    __shared__ double red_min[8], red_max[8];
    __shared__ int win_lo[kMetalStages];
    const DevModel &m = a.m;
    const PointTab &pt = a.pt;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = b < a.B;
    if (!active) b = a.B - 1;  // keep the whole CTA alive for the barriers; results are not stored
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    const double eb = S[0], ebeta = S[(size_t)ldb], es = S[(size_t)2 * ldb];
    const double dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    const double dpi = dv * a.vfac[0];
    const bool aniso = m.flags & BF_FLAG_ANISOTROPIC;
    const double iso = pT[(size_t)(m.idx_bao + 1) * ldb] * es;
    const double apar = pT[(size_t)(m.idx_bao + 2) * ldb] * es;
    const double aperp = ((m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * pT[(size_t)(m.idx_bao + 2) * ldb]
                                                              : pT[(size_t)(m.idx_bao + 3) * ldb]) * es;
    double n[NT];
#pragma unroll
    for (int c = 0; c < NT / 3; ++c) cross_metal_norms(m, pT, ldb, c, eb, ebeta, n[3 * c], n[3 * c + 1], n[3 * c + 2]);
    const int ny = pt.mx_ny;
    const int i0 = blockIdx.y * a.pts_per_block, i1 = min(min(i0 + a.pts_per_block, pt.npts_pad), pt.npts);

    // range of the velocity shift over the CTA
    double lo = dpi, hi = dpi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { red_min[warp] = lo; red_max[warp] = hi; }
    if (tid == 0)
        for (int st = 0; st < kMetalStages; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], 8); }
    __syncthreads();
    double dpi_min = red_min[0], dpi_max = red_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { dpi_min = fmin(dpi_min, red_min[w]); dpi_max = fmax(dpi_max, red_max[w]); }

    // first row of the window of bin i, or -1 if the bin cannot be staged
    auto window = [&](int i) {
        const double r0 = pt.rpar0[i];
        const double ya = fmin(fmax(r0 + dpi_min, m.metal_y0), m.metal_ymax), yb = fmin(fmax(r0 + dpi_max, m.metal_y0), m.metal_ymax);
        const int ia = (int)floor((ya - m.metal_y0) * m.metal_inv_dy), ib = (int)floor((yb - m.metal_y0) * m.metal_inv_dy);
        const int first = ia - 1, last = ib + 2;
        if (NT % 2 != 0 || ny < kMetalWin || first < 0 || last > ny - 1 || last - first + 1 > kMetalWin) return -1;  // odd NT: rows are not 16-byte aligned
        return min(first, ny - kMetalWin);
    };
    auto issue = [&](int i) {  // thread 0
        const int st = (i - i0) % kMetalStages, w = window(i);
        win_lo[st] = w;
        if (w >= 0) {
            mbar_expect_tx(&full_bar[st], (unsigned)(kMetalWin * NT * sizeof(double)));
            tma_bulk_g2s(&win[st][0], pt.mx + ((size_t)i * ny + w) * NT, (unsigned)(kMetalWin * NT * sizeof(double)), &full_bar[st]);
        } else {
            // nothing to copy: complete the phase right away so that consumers fall through
            mbar_expect_tx(&full_bar[st], 0);
        }
    };
    if (tid == 0)
        for (int i = i0; i < i1 && i < i0 + kMetalStages; ++i) issue(i);
    __syncthreads();  // win_lo of the first stages is visible

    double x_next = (active && i0 < i1) ? a.out[(size_t)i0 * a.ldo + b] : 0.0;
    for (int i = i0; i < i1; ++i) {
        const int k = i - i0, st = k % kMetalStages;
        // the read-modify-write of the xi panel: the value of the next bin is fetched a whole iteration ahead
        const double x_cur = x_next;
        if (active && i + 1 < i1) x_next = a.out[(size_t)(i + 1) * a.ldo + b];
        if (tid == 0 && k >= 1 && i - 1 + kMetalStages < i1) {
            // refill the stage of bin i-1 once all eight warps have released it
            mbar_wait(&empty_bar[(k - 1) % kMetalStages], ((k - 1) / kMetalStages) & 1);
            issue(i - 1 + kMetalStages);
        }
        __syncwarp();
        mbar_wait(&full_bar[st], (k / kMetalStages) & 1);
        const int wlo = win_lo[st];
        const double rperp = pt.rperp0[i], rpar = pt.rpar0[i] + dpi;
        bool zero = false;
        if (a.mode == 1) {  // same range zeroing as the cosmological pass (BaoKSpaceCorrelationModel.cc:466-469,485-488)
            const double rp2 = rpar * rpar, rt2 = rperp * rperp, r2 = rp2 + rt2;
            const double rprime = r2 * rsqrt(r2);
            double rBAO;
            if (aniso) {
                const double q = apar * apar * rp2 + aperp * aperp * rt2;
                rBAO = q * rsqrt(q);
            } else
                rBAO = rprime * iso;
            zero = (rprime < m.rlo || rprime > m.rhi) || (rBAO < m.rlo || rBAO > m.rhi);
        }
        const double y = fmin(fmax(rpar, m.metal_y0), m.metal_ymax);
        const double fy = (y - m.metal_y0) * m.metal_inv_dy, fly = floor(fy);
        double wy[4];
        catmull_weights(fy - fly, wy);
        const int iy = (int)fly;
        double wsum = 0.0;
        if (wlo >= 0) {
            const double *base = &win[st][0] + (iy - 1 - wlo) * NT;
#pragma unroll
            for (int jy = 0; jy < 4; ++jy) {
                const double *row = base + jy * NT;
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                if (NT % 4 == 0) {
#pragma unroll
                    for (int t = 0; t < NT; t += 4) {
                        const double2 u = *reinterpret_cast<const double2 *>(row + t), v = *reinterpret_cast<const double2 *>(row + t + 2);
                        d0 = fma(n[t], u.x, d0);
                        d1 = fma(n[t + 1], u.y, d1);
                        d2 = fma(n[t + 2], v.x, d2);
                        d3 = fma(n[t + 3], v.y, d3);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < NT; ++t) d0 = fma(n[t], row[t], d0);
                }
                wsum = fma(wy[jy], (d0 + d1) + (d2 + d3), wsum);
            }
        } else {
            const double *mxp = pt.mx + (size_t)i * ny * NT;
#pragma unroll
            for (int jy = 0; jy < 4; ++jy) {
                const double *row = mxp + (size_t)wrap_index(iy - 1 + jy, ny) * NT;
                double d0 = 0.0;
#pragma unroll
                for (int t = 0; t < NT; ++t) d0 = fma(n[t], __ldg(row + t), d0);
                wsum = fma(wy[jy], d0, wsum);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        if (active && !zero) a.out[(size_t)i * a.ldo + b] = x_cur + wsum;
    }
}

// Distortion-matrix tail on the data bins: Z = Y (1 + dist mul) + dist add * evol - d and the
// dilation-limit check of the data bin itself (BaoKSpaceCorrelationModel.cc:420-427, :529-535).
__global__ void __launch_bounds__(256, 2) fast_post_kernel(FastArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    const PointTab &pt = a.pt;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    const double iso = pT[(size_t)(m.idx_bao + 1) * ldb];
    const double apar0 = pT[(size_t)(m.idx_bao + 2) * ldb];
    const double aperp0 = (m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * apar0
                                                             : pT[(size_t)(m.idx_bao + 3) * ldb];
    const double dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    const bool aniso = m.flags & BF_FLAG_ANISOTROPIC;
    const double dmin2 = m.dilmin * m.dilmin, dmax2 = m.dilmax * m.dilmax;
    const bool need_r = [&] {
        bool n = false;
        for (int s = 0; s < 2; ++s) {
            const DevBB &q = m.bb[s];
            if (q.enabled) n |= !(q.r_min == 0 && q.r_max == 0) || !(q.mu_min == 0 && q.mu_max == 0);
        }
        return n;
    }();
    int st_bits = 0;
    const int i0 = blockIdx.y * a.pts_per_block, i1 = min(i0 + a.pts_per_block, pt.npts_pad);
    for (int i = i0; i < i1; ++i) {
        double val = 0.0;
        if (i < pt.npts) {
            int zc = a.nz == 1 ? 0 : pt.zc[i];
            const double eb = S[(size_t)(zc * 3 + 0) * ldb], es = S[(size_t)(zc * 3 + 2) * ldb];
            const double rperp = pt.rperp0[i], rpar = pt.rpar0[i] + dv * a.vfac[zc];
            const double rp2 = rpar * rpar, rt2 = rperp * rperp, r2 = rp2 + rt2;
            double scale2;  // (r_BAO / r)^2
            if (aniso) {
                const double apar = apar0 * es, aperp = aperp0 * es;
                scale2 = apar * apar * rp2 + aperp * aperp * rt2;
            } else {
                const double sc = iso * es;
                scale2 = sc * sc * r2;
            }
            if (scale2 < dmin2 * r2) st_bits |= BF_STATUS_DILATION_MIN;
            else if (scale2 > dmax2 * r2) st_bits |= BF_STATUS_DILATION_MAX;
            double rprime = 0.0, muprime = 0.0;
            if (need_r) {
                const double rs = rsqrt(r2);
                rprime = r2 * rs;
                muprime = rpar * rs;
            }
            double xi = a.in[(size_t)i * a.ldo + b];
            xi = broadband_pair(m, pt, BF_BB_DIST_ADD, BF_BB_DIST_MUL, i, pT, ldb, eb, xi, rprime, muprime, rpar);
            val = st_bits ? nan("") : xi;
            if (a.dov) val -= a.dov[(size_t)b * pt.npts + i];
            else if (a.dsub) val -= a.dsub[i];
        }
        a.out[(size_t)i * a.ldo + b] = val;
    }
    if (st_bits) atomicOr(a.status + b, st_bits);
}

// Fused tail: per-vector coefficient rows appended below the xi_u panel, so that the post-matrix
// additive broadband and the "- d" become extra K columns of the distortion GEMM. Also performs
// the dilation-limit check of every data bin (BaoKSpaceCorrelationModel.cc:420-427).
//   case 1 (no velocity shift): row k = eb * c_k                      (basis fully bin-static)
//   case 2 (velocity shift):    rP factors are (u_i + v_b)^p with v_b = dpi_b / r0, expanded
//           binomially: row (z, m, t) = eb * sum_{p >= m} C(p,m) c_{z,p,t} v_b^(p-m)
__global__ void __launch_bounds__(128) fold_coef_kernel(FoldArgs a) {
    pdl_sync();
    BF_GATE(a.gate);
    const DevModel &m = a.m;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int ldb = a.ldb;
    const double *pT = a.pT + b;
    const double *S = a.S + b;
    const double eb = S[(size_t)(a.zc_data * 3 + 0) * ldb], es = S[(size_t)(a.zc_data * 3 + 2) * ldb];
    const double dv = m.idx_delta_v >= 0 ? pT[(size_t)m.idx_delta_v * ldb] : 0.0;
    const double dpi = dv * a.vfac[a.zc_data];
    double *rows = a.rows + b;
    const DevBB &s = m.bb[BF_BB_DIST_ADD];
    if (a.fold_case == 1 && a.ncls > 1) {
        const int nb1 = a.nb / a.ncls;
        for (int c = 0; c < a.ncls; ++c) {
            const double ebc = S[(size_t)(c * 3 + 0) * ldb];
            for (int k = 0; k < nb1; ++k) rows[(size_t)(c * nb1 + k) * ldb] = ebc * pT[(size_t)(s.idx_base + k) * ldb];
        }
    } else if (a.fold_case == 1) {
        for (int k = 0; k < a.nb; ++k) rows[(size_t)k * ldb] = eb * pT[(size_t)(s.idx_base + k) * ldb];
    } else if (a.fold_case == 2) {
        const double v = dpi * s.inv_r0;
        const int n_rp = (s.rp_max - s.rp_min) / s.rp_step + 1;
        for (int zi = 0; zi < a.nz; ++zi)
            for (int mm = 0; mm <= a.pmax; ++mm)
                for (int t = 0; t < a.nrt; ++t) {
                    double acc = 0.0;
                    for (int ip = 0; ip < n_rp; ++ip) {
                        int p = s.rp_min + ip * s.rp_step;
                        if (p < mm) continue;
                        // C(p, mm) v^(p-mm); exponent 0 terms use x, positive ones x-1: p = 0 only feeds mm = 0
                        double binom = 1.0;
                        for (int q = 0; q < mm; ++q) binom = binom * (p - q) / (q + 1);
                        acc += binom * ipow(v, p - mm) * pT[(size_t)(s.idx_base + (zi * n_rp + ip) * a.nrt + t) * ldb];
                    }
                    rows[(size_t)((zi * (a.pmax + 1) + mm) * a.nrt + t) * ldb] = eb * acc;
                }
    }
    rows[(size_t)a.nb * ldb] = a.minus_d_row_value ? 1.0 : 0.0;
    for (int k = a.nb + 1; k < a.kext; ++k) rows[(size_t)k * ldb] = 0.0;
    // dilation limits: scalez^2 = apar^2 mu^2 + aperp^2 (1 - mu^2) lies between the two squares, so
    // bins only need checking when one of the two scales is outside the limits
    const bool aniso = m.flags & BF_FLAG_ANISOTROPIC;
    const double iso = pT[(size_t)(m.idx_bao + 1) * ldb] * es;
    const double apar = pT[(size_t)(m.idx_bao + 2) * ldb] * es;
    const double aperp = ((m.flags & BF_FLAG_COMBINED_SCALE) ? pT[(size_t)m.idx_comb_scale * ldb] * pT[(size_t)(m.idx_bao + 2) * ldb]
                                                              : pT[(size_t)(m.idx_bao + 3) * ldb]) * es;
    int st = 0;
    if (!aniso) {
        if (iso < m.dilmin) st = BF_STATUS_DILATION_MIN;
        else if (iso > m.dilmax) st = BF_STATUS_DILATION_MAX;
    } else if (fmin(apar, aperp) < m.dilmin || fmax(apar, aperp) > m.dilmax) {
        const double dmin2 = m.dilmin * m.dilmin, dmax2 = m.dilmax * m.dilmax;
        for (int i = 0; i < a.npts; ++i) {
            const double rpar = a.rpar0[i] + dpi, rperp = a.rperp0[i];
            const double rp2 = rpar * rpar, rt2 = rperp * rperp, r2 = rp2 + rt2;
            const double sc2 = apar * apar * rp2 + aperp * aperp * rt2;
            if (sc2 < dmin2 * r2) st |= BF_STATUS_DILATION_MIN;
            else if (sc2 > dmax2 * r2) st |= BF_STATUS_DILATION_MAX;
        }
    }
    if (st) atomicOr(a.status + b, st);
}

void launch_fold_coef(cudaStream_t s, const FoldArgs &a) { Launch((a.B + 127) / 128, 128, 0, s)(fold_coef_kernel, a); }

static void pick_grid(FastArgs &a, int sm_count, dim3 &grid, int threads = 256) {
    int bx = (a.B + threads - 1) / threads;
    int want = sm_count * (threads == 256 ? 6 : 4);
    int by = max(1, min(a.pt.npts_pad, (want + bx - 1) / bx));
    a.pts_per_block = (a.pt.npts_pad + by - 1) / by;
    grid = dim3(bx, (a.pt.npts_pad + a.pts_per_block - 1) / a.pts_per_block);
}

void launch_kspace_fast_eval(cudaStream_t s, FastArgs &a, int sm_count, int part) {
    size_t smem = (size_t)2 * a.m.nr * 9 * sizeof(double2);
    static SmemOptIn opt_lean, opt_full, opt_rs;
    opt_lean.ensure(kspace_fast_eval_kernel<true, 1>, smem);
    opt_full.ensure(kspace_fast_eval_kernel<false, 0>, smem);
    opt_rs.ensure(kspace_fast_eval_kernel<true, 1, true>, smem);
    dim3 grid;
    pick_grid(a, sm_count, grid);
    if (part == 1) {
        // the cosmology pass is bound by shared-memory wavefronts and its 72 KB table allows 3 CTAs per SM at most:
        // 512 threads at 64 registers (2 CTAs, 32 warps per SM instead of 24) measured 0.344 -> 0.332 ms
        pick_grid(a, sm_count, grid, 512);
        if (a.rspace) Launch(grid, 512, smem, s)(kspace_fast_eval_kernel<true, 1, true>, a);
        else Launch(grid, 512, smem, s)(kspace_fast_eval_kernel<true, 1>, a);
    }
    else if (part == 2) {
        const int sa = a.mode == 1 ? BF_BB_DM_DIST_ADD : BF_BB_DIST_ADD, sm = a.mode == 1 ? BF_BB_DM_DIST_MUL : BF_BB_DIST_MUL;
        const bool metals_only = a.m.metal_kind == BF_METAL_CROSS_BICUBIC && a.nz == 1 && !a.m.bb[sa].enabled &&
                                 !a.m.bb[sm].enabled && !(a.mode == 0 && (a.dsub || a.dov)) &&
                                 (a.pt.mx_nt == 12 || a.pt.mx_nt == 15);
        if (metals_only && a.pt.mx_nt == 12) Launch(grid, 256, 0, s)(metal_add_kernel<12>, a);
        else if (metals_only) Launch(grid, 256, 0, s)(metal_add_kernel<15>, a);
        else Launch(grid, 256, 0, s)(kspace_fast_eval_kernel<true, 2>, a);
    }
    else Launch(grid, 256, smem, s)(kspace_fast_eval_kernel<false, 0>, a);
}

void launch_fast_post(cudaStream_t s, FastArgs &a, int sm_count) {
    dim3 grid;
    pick_grid(a, sm_count, grid);
    Launch(grid, 256, 0, s)(fast_post_kernel, a);
}

}  // namespace bf

namespace bf {
// Lean models (no radiation, metals none or bicubic cross) run as two passes with small register
// footprints: the cosmological part, then everything else (if there is anything else).
bool fast_eval_is_lean(const FastArgs &a) {
    return !(a.m.flags & BF_FLAG_RADIATION_MODEL) &&
           (a.m.metal_kind == BF_METAL_NONE || a.m.metal_kind == BF_METAL_CROSS_BICUBIC);
}
bool fast_eval_has_extras(const FastArgs &a) {
    const int sa = a.mode == 1 ? BF_BB_DM_DIST_ADD : BF_BB_DIST_ADD, sm = a.mode == 1 ? BF_BB_DM_DIST_MUL : BF_BB_DIST_MUL;
    return a.m.metal_kind != BF_METAL_NONE || a.m.bb[sa].enabled || a.m.bb[sm].enabled || (a.mode == 0 && (a.dsub || a.dov));
}
}  // namespace bf
