/*
 * oracle.c -- CPU restatement (plain C99) of igmhub/baofit's fit-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle.h for who may load this and for the pinning status
 * of each part. Deliberately written bin-by-bin and scalar, in the shape of the reference
 * (CorrelationFitter::getPrediction -> AbsCorrelationModel::evaluate -> _evaluate), so that it
 * shares no arithmetic structure with the CUDA kernels it checks. Citations are
 * file:line under /root/reference.
 *
 * Third-party algorithms restated because their sources are not in the reference tree
 * (un-vendored, unpinned dependencies, configure.ac:25-43):
 *   - gsl_interp_cspline (GSL, natural cubic spline; used through likely::Interpolator)
 *   - likely::BiCubicInterpolator (bicubic patch with centred finite-difference derivatives
 *     and periodic index wrap == tensor-product Catmull-Rom; layout UNVERIFIED)
 *   - cosmo::DistortedPowerCorrelation (replaced by our own fine-grid Filon quadrature,
 *     documented in DESIGN.md "k-space transform"; UNPINNED)
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static char g_err[512];
const char *orc_last_error(void) { return g_err; }
static void set_err(const char *s) { snprintf(g_err, sizeof(g_err), "%s", s); }

/* ------------------------------------------------------------------------------------ */
/* natural cubic spline (gsl_interp_cspline restated)                                    */
/* ------------------------------------------------------------------------------------ */

void orc_cspline_coeffs(int n, const double *x, const double *y, double *c) {
    /* GSL cspline_init: c[0] = c[n-1] = 0, interior from the symmetric tridiagonal system
     *   h_{i} c_{i} + 2 (h_{i} + h_{i+1}) c_{i+1} + h_{i+1} c_{i+2} = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i) */
    int i, m = n - 2;
    c[0] = 0.0;
    c[n - 1] = 0.0;
    if (m <= 0) return;
    double *diag = (double *)malloc(sizeof(double) * m);
    double *off = (double *)malloc(sizeof(double) * m);
    double *g = (double *)malloc(sizeof(double) * m);
    for (i = 0; i < m; ++i) {
        double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
        double yd_i = y[i + 1] - y[i], yd_ip1 = y[i + 2] - y[i + 1];
        double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
        off[i] = h_ip1;
        diag[i] = 2.0 * (h_ip1 + h_i);
        g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
    }
    /* Thomas algorithm on (diag, off, off) */
    for (i = 1; i < m; ++i) {
        double w = off[i - 1] / diag[i - 1];
        diag[i] -= w * off[i - 1];
        g[i] -= w * g[i - 1];
    }
    c[m] = g[m - 1] / diag[m - 1];
    for (i = m - 2; i >= 0; --i) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
    free(diag);
    free(off);
    free(g);
}

double orc_cspline_eval(int n, const double *x, const double *y, const double *c, double xq) {
    /* likely::Interpolator returns the end-point value outside the tabulated range
     * (confirmed by the corners of data/BOSSDR11QSOLyaF.scan) */
    if (xq <= x[0]) return y[0];
    if (xq >= x[n - 1]) return y[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (x[mid] > xq) hi = mid; else lo = mid;
    }
    /* gsl cspline_eval */
    double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo], delx = xq - x[lo];
    double b = dy / dx - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
    double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
    return y[lo] + delx * (b + delx * (c[lo] + delx * d));
}

typedef struct {
    int n;
    double *x, *y, *c;
} spline_t;

static void spline_make(spline_t *s, int n, const double *x, const double *y) {
    s->n = n;
    s->x = (double *)malloc(sizeof(double) * n);
    s->y = (double *)malloc(sizeof(double) * n);
    s->c = (double *)malloc(sizeof(double) * n);
    memcpy(s->x, x, sizeof(double) * n);
    memcpy(s->y, y, sizeof(double) * n);
    orc_cspline_coeffs(n, s->x, s->y, s->c);
}
static void spline_free(spline_t *s) {
    free(s->x); free(s->y); free(s->c);
    s->x = s->y = s->c = NULL;
}
static double spline_eval(const spline_t *s, double x) {
    return orc_cspline_eval(s->n, s->x, s->y, s->c, x);
}

/* ------------------------------------------------------------------------------------ */
/* small helpers restating reference functions                                           */
/* ------------------------------------------------------------------------------------ */

double orc_redshift_evolution(double p0, double gamma, double z, double zref) {
    return p0 * pow((1 + z) / (1 + zref), gamma); /* AbsCorrelationModel.cc:159-161 */
}

double orc_legendre(int ell, double mu) { /* BroadbandModel.cc:171-195 */
    double musq = mu * mu;
    switch (ell) {
    case 0: return 1;
    case 1: return mu;
    case 2: return (-1 + 3 * musq) / 2.;
    case 3: return mu * (-3 + 5 * musq) / 2.;
    case 4: return (3 + musq * (-30 + 35 * musq)) / 8.;
    case 5: return mu * (15 + musq * (-70 + 63 * musq)) / 8.;
    case 6: return (-5 + musq * (105 + musq * (-315 + 231 * musq))) / 16.;
    case 7: return mu * (-35 + musq * (315 + musq * (-693 + 429 * musq))) / 16.;
    case 8: return (35 + musq * (-1260 + musq * (6930 + musq * (-12012 + 6435 * musq)))) / 128.;
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------ */
/* model state                                                                           */
/* ------------------------------------------------------------------------------------ */

#define NFINE_SUB 4 /* fine k nodes per native P(k) interval */
#define NMU_PROJ 20 /* Gauss-Legendre nodes on mu in [0,1] for the multipoles of D(k,mu) */

struct orc_model {
    bf_model_desc d; /* shallow copy; pointers below own deep copies */
    spline_t fid[3], nw[3];
    int *metal_p1, *metal_p2;
    double *metal_zgrid, *metal_auto, *metal_cross, *dist_matrix;
    /* k-space */
    int n_pk;
    double *pk_lnk;
    spline_t pk_fid, pk_nw; /* natural cubic splines of P versus ln k */
    int nk;                 /* coarse nodes for the multipoles of D */
    double *kc;             /* coarse nodes */
    int nfine;
    double *kf, *pf_pk, *pf_nw; /* fine nodes and P values (peak = fid - nw, nw) */
    int nr;
    double r0, dr, rlo, rhi;
    double gl_x[NMU_PROJ], gl_w[NMU_PROJ];
    /* per-evaluation scratch that mirrors the reference's mutable members */
    double betaz, beta2z, snl_par2, snl_perp2, zeff;
    /* cached transform tables: xi_ell(r_j) splines for [comp][ell] */
    int have_tables;
    double *tab_r;
    double *tab_y[2][3], *tab_c[2][3];
    double *last_params;
    double last_ztrig;
    /* 3-D FFT model (BaoKSpaceFftCorrelationModel.cc) */
    int fnx, fny, fnz;      /* grid sizes, y = line of sight */
    int fmx, fmy, fmz;      /* number of non-negative frequencies per axis */
    int fnpx, fnpy;         /* extent of the tabulated plane */
    double fdelta;
    double pk_slope[2][2];  /* power-law extrapolation exponents [fid,nw][below,above] */
    double *fplane[2];      /* xi_comp(ix*delta, iy*delta), [iy*fnpx + ix] */
    double fkey[16];
    int fhave;
};

struct orc_data {
    bf_data_desc d;
    double *r, *mu, *z, *data, *cov, *grid_r, *grid_mu, *grid_z;
    int *index;
    double *L; /* Cholesky factor of C (row-major lower) when a covariance was given */
};

static double *dupd(const double *p, size_t n) {
    if (!p || n == 0) return NULL;
    double *q = (double *)malloc(sizeof(double) * n);
    memcpy(q, p, sizeof(double) * n);
    return q;
}
static int *dupi(const int *p, size_t n) {
    if (!p || n == 0) return NULL;
    int *q = (int *)malloc(sizeof(int) * n);
    memcpy(q, p, sizeof(int) * n);
    return q;
}

static void gauss_legendre_01(int n, double *x, double *w) {
    /* nodes/weights on [0,1] by Newton iteration on P_n */
    int i, j;
    for (i = 0; i < n; ++i) {
        double t = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0, p1;
        int it;
        for (it = 0; it < 100; ++it) {
            double p2 = 0;
            p1 = 1;
            for (j = 0; j < n; ++j) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * t * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (t * p1 - p2) / (t * t - 1.0);
            double dt = p1 / pp;
            t -= dt;
            if (fabs(dt) < 1e-16) break;
        }
        x[i] = 0.5 * (1.0 - t);
        w[i] = 1.0 / ((1.0 - t * t) * pp * pp);
    }
}

static int kspace_setup(orc_model *m) {
    const bf_model_desc *d = &m->d;
    int i, s, n = d->n_pk;
    if (n < 4 || !d->pk_k || !d->pk_fid || !d->pk_nw) { set_err("orc: k-space model needs P(k) tables"); return -1; }
    m->n_pk = n;
    m->pk_lnk = (double *)malloc(sizeof(double) * n);
    for (i = 0; i < n; ++i) m->pk_lnk[i] = log(d->pk_k[i]);
    spline_make(&m->pk_fid, n, m->pk_lnk, d->pk_fid);
    spline_make(&m->pk_nw, n, m->pk_lnk, d->pk_nw);
    double klo = d->pk_k[0], khi = d->pk_k[n - 1];
    /* BaoKSpaceCorrelationModel.cc:140-141 */
    m->nk = (int)ceil(log10(khi / klo) * d->samples_per_decade);
    m->kc = (double *)malloc(sizeof(double) * m->nk);
    for (i = 0; i < m->nk; ++i) m->kc[i] = klo * pow(khi / klo, (double)i / (m->nk - 1));
    m->kc[m->nk - 1] = khi;
    /* fine nodes: every native interval split into NFINE_SUB equal parts in k */
    m->nfine = (n - 1) * NFINE_SUB + 1;
    m->kf = (double *)malloc(sizeof(double) * m->nfine);
    m->pf_pk = (double *)malloc(sizeof(double) * m->nfine);
    m->pf_nw = (double *)malloc(sizeof(double) * m->nfine);
    for (i = 0; i < n - 1; ++i)
        for (s = 0; s < NFINE_SUB; ++s) {
            double k = d->pk_k[i] + (d->pk_k[i + 1] - d->pk_k[i]) * s / NFINE_SUB;
            int f = i * NFINE_SUB + s;
            m->kf[f] = k;
            if (s == 0) {
                m->pf_nw[f] = d->pk_nw[i];
                m->pf_pk[f] = d->pk_fid[i] - d->pk_nw[i];
            } else {
                double lk = log(k), pn = spline_eval(&m->pk_nw, lk);
                m->pf_nw[f] = pn;
                m->pf_pk[f] = spline_eval(&m->pk_fid, lk) - pn;
            }
        }
    m->kf[m->nfine - 1] = d->pk_k[n - 1];
    m->pf_nw[m->nfine - 1] = d->pk_nw[n - 1];
    m->pf_pk[m->nfine - 1] = d->pk_fid[n - 1] - d->pk_nw[n - 1];
    /* radial nodes, BaoKSpaceCorrelationModel.cc:160-166 */
    m->rlo = d->rmin * d->dilmin;
    m->rhi = d->rmax * d->dilmax;
    m->nr = (int)ceil(m->rhi - m->rlo);
    if (m->nr < 4) { set_err("orc: k-space radial range too small"); return -1; }
    m->r0 = m->rlo;
    m->dr = (m->rhi - m->rlo) / (m->nr - 1);
    gauss_legendre_01(NMU_PROJ, m->gl_x, m->gl_w);
    m->tab_r = (double *)malloc(sizeof(double) * m->nr);
    for (i = 0; i < m->nr; ++i) m->tab_r[i] = m->r0 + i * m->dr;
    for (i = 0; i < 2; ++i)
        for (s = 0; s < 3; ++s) {
            m->tab_y[i][s] = (double *)calloc(m->nr, sizeof(double));
            m->tab_c[i][s] = (double *)calloc(m->nr, sizeof(double));
        }
    m->last_params = (double *)malloc(sizeof(double) * d->npar);
    m->have_tables = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* 3-D FFT model set-up (BaoKSpaceFftCorrelationModel.cc:64-97)                          */
/* ------------------------------------------------------------------------------------ */

static double extrap_slope(double k0, double p0, double k1, double p1) {
    /* cosmo::createTabulatedPower(..., extrapolateBelow, extrapolateAbove, ...) continues the table as
     * a power law through its last two points (UNVERIFIED dep); a sign change or a zero leaves a
     * constant */
    if (p0 * p1 <= 0) return 0;
    return log(p1 / p0) / log(k1 / k0);
}

static int fft_setup(orc_model *m) {
    const bf_model_desc *d = &m->d;
    int n = d->n_pk, i, c;
    if (n < 4 || !d->pk_k || !d->pk_fid || !d->pk_nw) { set_err("orc: FFT model needs P(k) tables"); return -1; }
    if (!(d->fft_spacing > 0) || d->fft_nx < 4 || d->fft_ny < 4 || d->fft_nz < 4) { set_err("orc: bad FFT grid"); return -1; }
    m->n_pk = n;
    m->pk_lnk = (double *)malloc(sizeof(double) * n);
    for (i = 0; i < n; ++i) m->pk_lnk[i] = log(d->pk_k[i]);
    spline_make(&m->pk_fid, n, m->pk_lnk, d->pk_fid);
    spline_make(&m->pk_nw, n, m->pk_lnk, d->pk_nw);
    for (c = 0; c < 2; ++c) {
        const double *pk = c == 0 ? d->pk_fid : d->pk_nw;
        m->pk_slope[c][0] = extrap_slope(d->pk_k[0], pk[0], d->pk_k[1], pk[1]);
        m->pk_slope[c][1] = extrap_slope(d->pk_k[n - 2], pk[n - 2], d->pk_k[n - 1], pk[n - 1]);
    }
    m->fnx = d->fft_nx; m->fny = d->fft_ny; m->fnz = d->fft_nz;
    m->fmx = m->fnx / 2 + 1; m->fmy = m->fny / 2 + 1; m->fmz = m->fnz / 2 + 1;
    m->fdelta = d->fft_spacing;
    double rmax = d->fft_rmax > 0 ? d->fft_rmax : d->rmax * d->dilmax;
    int np = (int)floor(rmax / m->fdelta) + 3;
    m->fnpx = np < m->fmx ? np : m->fmx;
    m->fnpy = np < m->fmy ? np : m->fmy;
    for (c = 0; c < 2; ++c) m->fplane[c] = (double *)calloc((size_t)m->fnpx * m->fnpy, sizeof(double));
    m->fhave = 0;
    return 0;
}

orc_model *orc_model_create(const bf_model_desc *desc) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->d = *desc;
    const bf_model_desc *d = &m->d;
    if (d->kind == BF_MODEL_RSPACE) {
        if (d->n_fid < 3 || d->n_nw < 3) { set_err("orc: r-space model needs tables"); free(m); return NULL; }
        spline_make(&m->fid[0], d->n_fid, d->fid_r, d->fid_xi0);
        spline_make(&m->fid[1], d->n_fid, d->fid_r, d->fid_xi2);
        spline_make(&m->fid[2], d->n_fid, d->fid_r, d->fid_xi4);
        spline_make(&m->nw[0], d->n_nw, d->nw_r, d->nw_xi0);
        spline_make(&m->nw[1], d->n_nw, d->nw_r, d->nw_xi2);
        spline_make(&m->nw[2], d->n_nw, d->nw_r, d->nw_xi4);
    } else if (d->kind == BF_MODEL_KSPACE) {
        if (kspace_setup(m)) { free(m); return NULL; }
    } else if (d->kind == BF_MODEL_KSPACE_FFT) {
        if (fft_setup(m)) { free(m); return NULL; }
    } else {
        set_err("orc: model kind not restated");
        free(m);
        return NULL;
    }
    if (d->metal_kind == BF_METAL_AUTO_TABLE || d->metal_kind == BF_METAL_CROSS_BICUBIC) {
        m->metal_p1 = dupi(d->metal_p1, d->metal_ncomb);
        m->metal_p2 = dupi(d->metal_p2, d->metal_ncomb);
        m->metal_zgrid = dupd(d->metal_zgrid, (size_t)d->metal_ncomb * d->metal_ngrid);
        if (d->metal_kind == BF_METAL_AUTO_TABLE)
            m->metal_auto = dupd(d->metal_auto, (size_t)d->metal_ncomb * 3 * d->metal_ngrid);
        else
            m->metal_cross = dupd(d->metal_cross, (size_t)d->metal_ncomb * 3 * d->metal_nx * d->metal_ny);
    }
    if (d->dist_matrix && d->dist_matrix_order > 0)
        m->dist_matrix = dupd(d->dist_matrix, (size_t)d->dist_matrix_order * d->dist_matrix_order);
    return m;
}

void orc_model_destroy(orc_model *m) {
    int i, s;
    if (!m) return;
    if (m->d.kind == BF_MODEL_RSPACE)
        for (i = 0; i < 3; ++i) { spline_free(&m->fid[i]); spline_free(&m->nw[i]); }
    if (m->d.kind == BF_MODEL_KSPACE) {
        free(m->pk_lnk); spline_free(&m->pk_fid); spline_free(&m->pk_nw);
        free(m->kc); free(m->kf); free(m->pf_pk); free(m->pf_nw); free(m->tab_r);
        for (i = 0; i < 2; ++i) for (s = 0; s < 3; ++s) { free(m->tab_y[i][s]); free(m->tab_c[i][s]); }
        free(m->last_params);
    }
    if (m->d.kind == BF_MODEL_KSPACE_FFT) {
        free(m->pk_lnk); spline_free(&m->pk_fid); spline_free(&m->pk_nw);
        free(m->fplane[0]); free(m->fplane[1]);
    }
    free(m->metal_p1); free(m->metal_p2); free(m->metal_zgrid); free(m->metal_auto);
    free(m->metal_cross); free(m->dist_matrix);
    free(m);
}

orc_data *orc_data_create(const bf_data_desc *desc) {
    orc_data *o = (orc_data *)calloc(1, sizeof(orc_data));
    int n = desc->n_data, i, j, k;
    o->d = *desc;
    o->r = dupd(desc->r, n); o->mu = dupd(desc->mu, n); o->z = dupd(desc->z, n);
    o->index = dupi(desc->index, n);
    o->data = dupd(desc->data, n);
    o->cov = dupd(desc->cov, (size_t)n * n);
    o->grid_r = dupd(desc->grid_r, desc->n_grid);
    o->grid_mu = dupd(desc->grid_mu, desc->n_grid);
    o->grid_z = dupd(desc->grid_z, desc->n_grid);
    if (o->cov && !desc->cov_is_inverse) {
        /* likely::CovarianceMatrix inverts through a Cholesky factorisation (dpptrf/dpptri);
         * we keep the factor and solve instead, which is the same quadratic form */
        double *L = (double *)calloc((size_t)n * n, sizeof(double));
        for (i = 0; i < n; ++i)
            for (j = 0; j <= i; ++j) {
                double s = o->cov[(size_t)i * n + j];
                for (k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
                if (i == j) {
                    if (s <= 0) { set_err("orc: covariance not positive definite"); free(L); orc_data_destroy(o); return NULL; }
                    L[(size_t)i * n + i] = sqrt(s);
                } else
                    L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
            }
        o->L = L;
    }
    return o;
}

void orc_data_destroy(orc_data *o) {
    if (!o) return;
    free(o->r); free(o->mu); free(o->z); free(o->index); free(o->data); free(o->cov);
    free(o->grid_r); free(o->grid_mu); free(o->grid_z); free(o->L);
    free(o);
}

/* ------------------------------------------------------------------------------------ */
/* AbsCorrelationModel pieces                                                            */
/* ------------------------------------------------------------------------------------ */

typedef struct { /* AbsCorrelationModel::_updateInternalParameters, AbsCorrelationModel.cc:132-144 */
    double beta, bias, gamma_bias, gamma_beta, bias2, beta2;
} internal_t;

static void update_internal(const orc_model *m, const double *p, internal_t *in) {
    const bf_model_desc *d = &m->d;
    in->beta = p[d->idx_beta];
    in->bias = p[d->idx_bb] / (1 + in->beta);
    in->gamma_bias = p[d->idx_gamma_bias];
    in->gamma_beta = p[d->idx_gamma_beta];
    in->bias2 = in->beta2 = 0;
    if (d->idx_delta_v >= 0) {
        in->bias2 = p[d->idx_bias2];
        in->beta2 = p[d->idx_bb2] / in->bias2;
    }
    if (d->idx_beta_bias >= 0) in->bias = p[d->idx_beta_bias] / in->beta;
}

static void velocity_shift(const orc_model *m, const double *p, double *r, double *mu, double z) {
    /* AbsCorrelationModel.cc:146-157 */
    double dv = p[m->d.idx_delta_v];
    double zp1 = 1 + z;
    double om = m->d.omega_matter;
    double dpi = (dv / 100.) * (1 + z) / sqrt(1 - om + om * zp1 * zp1 * zp1);
    double rnew = sqrt((*r) * (*r) + 2 * (*r) * (*mu) * dpi + dpi * dpi);
    double munew = ((*r) * (*mu) + dpi) / rnew;
    *r = rnew;
    *mu = munew;
}

static double norm_factor(const orc_model *m, const double *p, int ell, double z) {
    /* AbsCorrelationModel.cc:163-205 */
    const bf_model_desc *d = &m->d;
    double beta = p[d->idx_beta], bb = p[d->idx_bb];
    double bias = bb / (1 + beta);
    if (d->flags & BF_FLAG_COMBINED_BIAS) bias = p[d->idx_beta_bias] / beta;
    double betaAvg, betaProd, biasSq;
    if (d->flags & BF_FLAG_CROSS_CORRELATION) {
        double bias2 = p[d->idx_bias2], bb2 = p[d->idx_bb2];
        double beta2 = bb2 / bias2;
        betaAvg = (beta + beta2) / 2;
        betaProd = beta * beta2;
        biasSq = bias * bias2;
    } else {
        betaAvg = beta;
        betaProd = beta * beta;
        biasSq = bias * bias;
    }
    double gammaBias = p[d->idx_gamma_bias], gammaBeta = p[d->idx_gamma_beta];
    biasSq = orc_redshift_evolution(biasSq, gammaBias, z, d->zref);
    betaAvg = orc_redshift_evolution(betaAvg, gammaBeta, z, d->zref);
    betaProd = orc_redshift_evolution(betaProd, 2 * gammaBeta, z, d->zref);
    if (ell == 4) return biasSq * betaProd * (8. / 35.);
    if (ell == 2) return biasSq * ((4. / 3.) * betaAvg + (4. / 7.) * betaProd);
    return biasSq * (1 + (2. / 3.) * betaAvg + (1. / 5.) * betaProd);
}

static double broadband_eval(const bf_broadband_spec *s, const double *p, double r, double mu, double z) {
    /* BroadbandModel.cc:197-226 */
    double xi = 0;
    double rr = r / s->r0, rrP = r * mu / s->r0, rrT = r * sqrt(1 - mu * mu) / s->r0;
    double zz = (1 + z) / (1 + s->z0);
    int off = 0, zi, mi, ri, pi, ti;
    for (zi = s->z_min; zi <= s->z_max; zi += s->z_step) {
        double zF = pow(zz, zi);
        for (mi = s->mu_min; mi <= s->mu_max; mi += s->mu_step) {
            double muF = orc_legendre(mi, mu);
            for (ri = s->r_min; ri <= s->r_max; ri += s->r_step) {
                double rF = pow(ri > 0 ? rr - 1 : rr, (double)ri / s->r_denom);
                for (pi = s->rp_min; pi <= s->rp_max; pi += s->rp_step) {
                    double rPF = pow(pi > 0 ? rrP - 1 : rrP, pi);
                    for (ti = s->rt_min; ti <= s->rt_max; ti += s->rt_step) {
                        double rTF = pow(ti > 0 ? rrT - 1 : rrT, ti);
                        double coef = p[s->idx_base + off];
                        off++;
                        xi += coef * rF * muF * rPF * rTF * zF;
                    }
                }
            }
        }
    }
    return xi;
}

/* likely::BiCubicInterpolator restated: bicubic patch from centred finite differences with
 * periodic index wrap == tensor-product Catmull-Rom (UNVERIFIED dep; see oracle.h) */
static double catmull(double pm1, double p0, double p1, double p2, double t) {
    return p0 + 0.5 * t * (p1 - pm1 + t * (2 * pm1 - 5 * p0 + 4 * p1 - p2 + t * (3 * (p0 - p1) + p2 - pm1)));
}
static int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }
static double bicubic_eval(const double *g, int nx, int ny, double x0, double y0, double dx, double dy,
                           double x, double y) {
    double fx = (x - x0) / dx, fy = (y - y0) / dy;
    int ix = (int)floor(fx), iy = (int)floor(fy), a, b;
    double tx = fx - ix, ty = fy - iy, col[4];
    for (b = 0; b < 4; ++b) {
        const double *row = g + (size_t)wrapi(iy - 1 + b, ny) * nx;
        double v[4];
        for (a = 0; a < 4; ++a) v[a] = row[wrapi(ix - 1 + a, nx)];
        col[b] = catmull(v[0], v[1], v[2], v[3], tx);
    }
    return catmull(col[0], col[1], col[2], col[3], ty);
}

static void update_norm_factors(double *n0, double *n2, double *n4, double biasSq, double betaAvg, double betaProd) {
    /* MetalCorrelationModel.cc:359-363 */
    *n0 = biasSq * (1 + (2. / 3.) * betaAvg + (1. / 5.) * betaProd);
    *n2 = biasSq * ((4. / 3.) * betaAvg + (4. / 7.) * betaProd);
    *n4 = biasSq * betaProd * (8. / 35.);
}

static double metal_eval(const orc_model *m, const double *p, const internal_t *in,
                         double r, double mu, double z, int index) {
    /* MetalCorrelationModel.cc:235-351 */
    const bf_model_desc *d = &m->d;
    double xi = 0;
    int i;
    (void)z;
    if (d->metal_kind == BF_METAL_AUTO_TABLE || d->metal_kind == BF_METAL_CROSS_BICUBIC) {
        double betap[8], biasp[8], n0 = 0, n2 = 0, n4 = 0;
        int cross = d->metal_kind == BF_METAL_CROSS_BICUBIC;
        double rperp = 0, rpar = 0;
        if (cross) { /* :269-274 */
            double xmax = d->metal_x0 + (d->metal_nx - 1) * d->metal_dx;
            double ymax = d->metal_y0 + (d->metal_ny - 1) * d->metal_dy;
            rperp = r * sqrt(1 - mu * mu);
            rpar = r * mu;
            if (rperp < d->metal_x0) rperp = d->metal_x0;
            if (rperp > xmax) rperp = xmax;
            if (rpar < d->metal_y0) rpar = d->metal_y0;
            if (rpar > ymax) rpar = ymax;
        }
        for (i = 0; i < d->metal_nmet; ++i) {
            if (i == 0) {
                betap[0] = cross ? in->beta2 : in->beta; /* :247-248, :282-283 */
                biasp[0] = cross ? in->bias2 : in->bias;
            } else {
                betap[i] = p[d->idx_metal + 2 * (i - 1)];
                biasp[i] = p[d->idx_metal + 2 * (i - 1) + 1];
            }
        }
        for (i = 0; i < d->metal_ncomb; ++i) {
            int a = m->metal_p1[i], b = m->metal_p2[i];
            double biasSq = biasp[a] * biasp[b];
            double betaAvg = 0.5 * (betap[a] + betap[b]);
            double betaProd = betap[a] * betap[b];
            double zg = m->metal_zgrid[(size_t)i * d->metal_ngrid + index];
            biasSq = orc_redshift_evolution(biasSq, in->gamma_bias, zg, d->zref);
            betaAvg = orc_redshift_evolution(betaAvg, in->gamma_beta, zg, d->zref);
            betaProd = orc_redshift_evolution(betaProd, 2 * in->gamma_beta, zg, d->zref);
            update_norm_factors(&n0, &n2, &n4, biasSq, betaAvg, betaProd);
            if (cross) {
                size_t sz = (size_t)d->metal_nx * d->metal_ny;
                const double *g = m->metal_cross + (size_t)(3 * i) * sz;
                xi += n0 * bicubic_eval(g, d->metal_nx, d->metal_ny, d->metal_x0, d->metal_y0, d->metal_dx, d->metal_dy, rperp, rpar)
                    + n2 * bicubic_eval(g + sz, d->metal_nx, d->metal_ny, d->metal_x0, d->metal_y0, d->metal_dx, d->metal_dy, rperp, rpar)
                    + n4 * bicubic_eval(g + 2 * sz, d->metal_nx, d->metal_ny, d->metal_x0, d->metal_y0, d->metal_dx, d->metal_dy, rperp, rpar);
            } else {
                const double *t = m->metal_auto + (size_t)(3 * i) * d->metal_ngrid;
                xi += n0 * t[index] + n2 * t[d->metal_ngrid + index] + n4 * t[2 * (size_t)d->metal_ngrid + index];
            }
        }
        return xi;
    }
    if (d->metal_kind == BF_METAL_TOY) { /* :303-346 */
        static const double rpar0[4] = {59.533, 111.344, 134.998, 174.525};
        static const double sigpar[4] = {0, 4.88626, 5.489, 8.48942};
        static const double sigperp[4] = {5.55309, 4.13032, 6.30484, 7.85636};
        double rperp = r * sqrt(1 - mu * mu), rpar = fabs(r * mu);
        for (i = 0; i < 4; ++i) {
            double sp = i == 0 ? p[d->idx_metal + 4] : sigpar[i];
            double exppar = (rpar - rpar0[i]) / sp, expperp = (rperp - 2.0) / sigperp[i];
            if (fabs(exppar) < 5 && fabs(expperp) < 10)
                xi += 1e-4 * p[d->idx_metal + i] * exp(-exppar * exppar - expperp);
        }
        return xi;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* BaoCorrelationModel::_evaluate, BaoCorrelationModel.cc:90-179                         */
/* ------------------------------------------------------------------------------------ */

static double rspace_evaluate(const orc_model *m, const double *p, const internal_t *in,
                              double r, double mu, double z, int index) {
    const bf_model_desc *d = &m->d;
    double ampl = p[d->idx_bao], scale = p[d->idx_bao + 1], scale_parallel = p[d->idx_bao + 2];
    double scale_perp = p[d->idx_bao + 3], gamma_scale = p[d->idx_bao + 4];
    if (d->flags & BF_FLAG_COMBINED_SCALE) scale_perp = p[d->idx_comb_scale] * scale_parallel;
    scale = orc_redshift_evolution(scale, gamma_scale, z, d->zref);
    scale_parallel = orc_redshift_evolution(scale_parallel, gamma_scale, z, d->zref);
    scale_perp = orc_redshift_evolution(scale_perp, gamma_scale, z, d->zref);
    double rBAO, muBAO;
    if (d->flags & BF_FLAG_ANISOTROPIC) {
        double ap1 = scale_parallel, bp1 = scale_perp, musq = mu * mu;
        double rscale = sqrt(ap1 * ap1 * musq + (1 - musq) * bp1 * bp1);
        rBAO = r * rscale;
        muBAO = mu * ap1 / rscale;
    } else {
        rBAO = r * scale;
        muBAO = mu;
    }
    double norm0 = norm_factor(m, p, 0, z), norm2 = norm_factor(m, p, 2, z), norm4 = norm_factor(m, p, 4, z);
    double musq = muBAO * muBAO;
    double L2 = (-1 + 3 * musq) / 2., L4 = (3 + musq * (-30 + 35 * musq)) / 8.;
    double fid = norm0 * spline_eval(&m->fid[0], rBAO) + norm2 * L2 * spline_eval(&m->fid[1], rBAO) + norm4 * L4 * spline_eval(&m->fid[2], rBAO);
    double nw = norm0 * spline_eval(&m->nw[0], rBAO) + norm2 * L2 * spline_eval(&m->nw[1], rBAO) + norm4 * L4 * spline_eval(&m->nw[2], rBAO);
    double peak = ampl * (fid - nw);
    double smooth = nw;
    if (d->flags & BF_FLAG_DECOUPLED) {
        double musq1 = mu * mu;
        double L2a = (-1 + 3 * musq1) / 2., L4a = (3 + musq1 * (-30 + 35 * musq1)) / 8.;
        smooth = norm0 * spline_eval(&m->nw[0], r) + norm2 * L2a * spline_eval(&m->nw[1], r) + norm4 * L4a * spline_eval(&m->nw[2], r);
    }
    double xi = peak + smooth;
    if (d->metal_kind != BF_METAL_NONE) xi += metal_eval(m, p, in, r, mu, z, index);
    if (d->bb[BF_BB_DIST_MUL].enabled) xi *= 1 + broadband_eval(&d->bb[BF_BB_DIST_MUL], p, r, mu, z);
    if (d->bb[BF_BB_DIST_ADD].enabled) {
        double distortion = broadband_eval(&d->bb[BF_BB_DIST_ADD], p, r, mu, z);
        xi += orc_redshift_evolution(distortion, in->gamma_bias, z, d->zref);
    }
    if (d->idx_rad >= 0) { /* :158-176 */
        double rad_strength = p[d->idx_rad], rad_aniso = p[d->idx_rad + 1];
        double mean_free_path = p[d->idx_rad + 2], quasar_lifetime = p[d->idx_rad + 3];
        if (rad_strength > 0 && r > 0.) {
            double rad = rad_strength / (r * r);
            rad *= exp(-r / mean_free_path);
            rad *= (1 - rad_aniso * (1 - mu * mu));
            double ctd = r * (1 - mu) / (1 + z);
            rad *= exp(-ctd / quasar_lifetime);
            xi += rad;
        }
    }
    return xi;
}

/* ------------------------------------------------------------------------------------ */
/* k-space: D(k,mu_k), multipoles, transform                                             */
/* ------------------------------------------------------------------------------------ */

static double lin_interp5(const double *x, const double *y, double xq) {
    /* likely::Interpolator(z,y,"linear") with end-point clamp */
    int i;
    if (xq <= x[0]) return y[0];
    if (xq >= x[4]) return y[4];
    for (i = 0; i < 4; ++i)
        if (xq < x[i + 1]) break;
    return y[i] + (y[i + 1] - y[i]) * (xq - x[i]) / (x[i + 1] - x[i]);
}

static double nl_correction(const orc_model *m, const double *p, double k, double mu_k, double pk, double z) {
    /* NonLinearCorrectionModel.cc:49-86 */
    const bf_model_desc *d = &m->d;
    static const double zA[] = {2.2000, 2.4000, 2.6000, 2.8000, 3.0000};
    static const double qnlA[] = {0.8670, 0.8510, 0.7810, 0.7730, 0.7920};
    static const double kvA[] = {1.1200, 1.1122, 1.2570, 1.2765, 1.2928};
    static const double avA[] = {0.5140, 0.5480, 0.6110, 0.6080, 0.5780};
    static const double bvA[] = {1.6000, 1.6100, 1.6400, 1.6500, 1.6300};
    static const double kpA[] = {19.400, 19.500, 21.100, 19.200, 17.100};
    if (d->flags & (BF_FLAG_NL_CORRECTION | BF_FLAG_FIT_NL_CORRECTION)) {
        double qnl = lin_interp5(zA, qnlA, z), kv = lin_interp5(zA, kvA, z), av = lin_interp5(zA, avA, z);
        double bv = lin_interp5(zA, bvA, z), kp = lin_interp5(zA, kpA, z);
        if (d->flags & BF_FLAG_FIT_NL_CORRECTION) {
            qnl = p[d->idx_nlcorr];
            kp = p[d->idx_nlcorr + 1];
        }
        double sigma8Sim = 0.8338, pi = 4 * atan(1);
        pk = pk * (sigma8Sim / d->sigma8) * (sigma8Sim / d->sigma8) * orc_redshift_evolution(1., -2., z, d->zref);
        double deltaSq = k * k * k * pk / (2 * pi * pi);
        double growth = qnl * deltaSq;
        double pecvelocity = pow(k / kv, av) * pow(fabs(mu_k), bv);
        double pressure = (k / kp) * (k / kp);
        return exp(growth * (1 - pecvelocity) - pressure);
    }
    if (d->flags & BF_FLAG_NL_CORRECTION_ALT) {
        double knl = 6.4, pnl = 0.569, kpp = 15.3, pp = 2.01, kv0 = 1.22, pv = 1.5, kvi = 0.923, pvi = 0.451;
        double kvel = kv0 * pow(1 + k / kvi, pvi);
        double growth = pow(k / knl, pnl);
        double pressure = pow(k / kpp, pp);
        double pecvelocity = pow(fabs(k * mu_k) / kvel, pv);
        return exp(growth - pressure - pecvelocity);
    }
    return 1;
}

static double kspace_distortion(const orc_model *m, const double *p, double k, double mu_k, double pk) {
    /* BaoKSpaceCorrelationModel.cc:215-283, using the mutable members set by _evaluate */
    const bf_model_desc *d = &m->d;
    int cross = (d->flags & BF_FLAG_CROSS_CORRELATION) != 0;
    double mu2 = mu_k * mu_k;
    double kpar = fabs(k * mu_k);
    double tracer1 = 1 + m->betaz * mu2;
    double tracer2 = cross ? 1 + m->beta2z * mu2 : tracer1;
    double linear = tracer1 * tracer2;
    if (d->flags & BF_FLAG_UVFLUCTUATION) {
        double uvRelBias = p[d->idx_uv], uvResp = p[d->idx_uv + 1], uvMfp = p[d->idx_uv + 2];
        double s = uvMfp * k;
        double Ws = atan(s) / s;
        double uvFunc = 1 + uvRelBias * Ws / (1 + uvResp * Ws);
        tracer1 = uvFunc + m->betaz * mu2;
        tracer2 = cross ? 1 + m->beta2z * mu2 : tracer1;
        linear = tracer1 * tracer2;
    }
    if (d->flags & (BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT)) {
        double hcdBeta = p[d->idx_hcd], hcdRelBias = p[d->idx_hcd + 1], hcdScale = p[d->idx_hcd + 2];
        double hcdFunc = 1;
        if (d->flags & BF_FLAG_HCD_MODEL) hcdFunc = sin(hcdScale * kpar) / (hcdScale * kpar);
        else hcdFunc = exp(-(hcdScale * kpar) * (hcdScale * kpar) / 2);
        double tracerHcd = hcdRelBias * (1 + hcdBeta * mu2) * hcdFunc;
        linear = cross ? tracer1 * tracer2 + tracer2 * tracerHcd
                       : tracer1 * tracer1 + 2 * tracer1 * tracerHcd + tracerHcd * tracerHcd;
    }
    double smoothbin = 1;
    if (d->flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) {
        double bsPar = p[d->idx_bs], bsPerp = p[d->idx_bs + 1];
        if (d->flags & BF_FLAG_BIN_SMOOTH) {
            double kperp = k * sqrt(1 - mu2);
            double sPar = sin(0.5 * bsPar * kpar) / (0.5 * bsPar * kpar);
            double sPerp = sin(0.5 * bsPerp * kperp) / (0.5 * bsPerp * kperp);
            smoothbin = sPar * sPerp;
        }
        if (d->flags & BF_FLAG_BIN_SMOOTH_ALT) {
            double bs2 = bsPar * bsPar * mu2 + bsPerp * bsPerp * (1 - mu2);
            smoothbin = exp(-0.5 * bs2 * k * k);
        }
        smoothbin = smoothbin * smoothbin;
    }
    double gaussmooth = 1;
    if (d->flags & BF_FLAG_SMOOTH_GAUSS) {
        double sg = p[d->idx_smgaus];
        gaussmooth = exp(-0.5 * sg * sg * kpar * kpar);
    }
    double lorsmooth = 1;
    if (d->flags & BF_FLAG_SMOOTH_LORENTZ) {
        double sl = p[d->idx_smlor];
        lorsmooth = sqrt(1 / (1 + sl * sl * kpar * kpar));
    }
    double snl2 = m->snl_par2 * mu2 + m->snl_perp2 * (1 - mu2);
    double nonlinear = exp(-0.5 * snl2 * k * k);
    double nonlinearcorr = nl_correction(m, p, k, mu_k, pk, m->zeff);
    if (cross) nonlinearcorr = sqrt(nonlinearcorr);
    return nonlinear * nonlinearcorr * linear * smoothbin * gaussmooth * lorsmooth;
}

/* sets the mutable members of BaoKSpaceCorrelationModel.cc:315-330 for a transform at z */
static void kspace_set_state(orc_model *m, const double *p, double z, int component) {
    const bf_model_desc *d = &m->d;
    double beta = p[d->idx_beta], gammaBeta = p[d->idx_gamma_beta];
    m->zeff = d->zeff > 0 ? d->zeff : z;
    m->betaz = orc_redshift_evolution(beta, gammaBeta, z, d->zref);
    m->beta2z = 0;
    if (d->flags & BF_FLAG_CROSS_CORRELATION) {
        double beta2 = p[d->idx_bb2] / p[d->idx_bias2];
        m->beta2z = orc_redshift_evolution(beta2, gammaBeta, z, d->zref);
    }
    double snlPerp = p[d->idx_nl], snlPar = snlPerp * p[d->idx_nl + 1];
    m->snl_perp2 = snlPerp * snlPerp;
    m->snl_par2 = snlPar * snlPar;
    /* :364-367 the no-wiggles component is only broadened with nl-broadband */
    if (component == 1 && !(d->flags & BF_FLAG_NL_BROADBAND)) m->snl_perp2 = m->snl_par2 = 0;
}

double orc_kspace_distortion(orc_model *m, const double *params, double z, int component, double k, double muk) {
    kspace_set_state(m, params, z, component);
    double lk = log(k);
    double pk = component == 0 ? spline_eval(&m->pk_fid, lk) - spline_eval(&m->pk_nw, lk)
                               : spline_eval(&m->pk_nw, lk);
    return kspace_distortion(m, params, k, muk, pk);
}

/* multipole ell of D at coarse node i: (2 ell + 1) Int_0^1 L_ell(mu) D(k_i, mu) dmu */
static double d_multipole(const orc_model *m, const double *p, int component, int i, int ell) {
    double k = m->kc[i], lk = log(k), s = 0;
    double pk = component == 0 ? spline_eval(&m->pk_fid, lk) - spline_eval(&m->pk_nw, lk)
                               : spline_eval(&m->pk_nw, lk);
    int q;
    for (q = 0; q < NMU_PROJ; ++q)
        s += m->gl_w[q] * orc_legendre(ell, m->gl_x[q]) * kspace_distortion(m, p, k, m->gl_x[q], pk);
    return (2 * ell + 1) * s;
}

/* 4-point Lagrange interpolation of coarse samples f[0..nk) (uniform in ln k) at ln k = u */
static double lagrange4(const orc_model *m, const double *f, double u) {
    double u0 = log(m->kc[0]), h = (log(m->kc[m->nk - 1]) - u0) / (m->nk - 1);
    double x = (u - u0) / h;
    int i = (int)floor(x);
    if (i < 1) i = 1;
    if (i > m->nk - 3) i = m->nk - 3;
    double t = x - i; /* stencil nodes i-1, i, i+1, i+2 at t = -1, 0, 1, 2 */
    double w0 = -t * (t - 1) * (t - 2) / 6, w1 = (t + 1) * (t - 1) * (t - 2) / 2;
    double w2 = -(t + 1) * t * (t - 2) / 2, w3 = (t + 1) * t * (t - 1) / 6;
    return w0 * f[i - 1] + w1 * f[i] + w2 * f[i + 1] + w3 * f[i + 2];
}

static void sph_bessel_024(double x, double *j0, double *j2, double *j4) {
    if (x < 0.5) { /* series, avoids cancellation */
        double x2 = x * x;
        *j0 = 1 - x2 / 6 * (1 - x2 / 20 * (1 - x2 / 42 * (1 - x2 / 72)));
        *j2 = x2 / 15 * (1 - x2 / 14 * (1 - x2 / 36 * (1 - x2 / 66 * (1 - x2 / 104))));
        *j4 = x2 * x2 / 945 * (1 - x2 / 22 * (1 - x2 / 52 * (1 - x2 / 90 * (1 - x2 / 136))));
        return;
    }
    double s = sin(x), c = cos(x), xi = 1 / x, xi2 = xi * xi;
    *j0 = s * xi;
    *j2 = (3 * xi2 - 1) * s * xi - 3 * c * xi2;
    *j4 = (105 * xi2 * xi2 - 45 * xi2 + 1) * s * xi - (105 * xi2 - 10) * c * xi2;
}

/* brute-force Int G_l(k) k^2 j_l(k r) dk, l = 0,2,4, for G_l piecewise linear on the fine nodes:
 * every fine panel is cut into sub-panels of at most 1.5 rad of phase, 8-point Gauss-Legendre each */
static void hankel_brute(int nf, const double *kf, const double *g0, const double *g2, const double *g4, double r,
                         double out[3]) {
    static const double glx[8] = {-0.9602898564975363, -0.7966664774136267, -0.5255324099163290, -0.1834346424956498,
                                  0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    static const double glw[8] = {0.1012285362903763, 0.2223810344533745, 0.3137066458778873, 0.3626837833783620,
                                  0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    const double *g[3] = {g0, g2, g4};
    double acc[3] = {0, 0, 0}, comp[3] = {0, 0, 0};
    int f, s, q, l;
    for (f = 0; f + 1 < nf; ++f) {
        double a = kf[f], b = kf[f + 1];
        int any = 0;
        for (l = 0; l < 3; ++l) any |= (g[l][f] != 0 || g[l][f + 1] != 0);
        if (!any) continue;
        int nsub = (int)ceil((b - a) * r / 1.5) + 1;
        double hs = (b - a) / nsub;
        for (s = 0; s < nsub; ++s) {
            double lo = a + s * hs, part[3] = {0, 0, 0};
            for (q = 0; q < 8; ++q) {
                double k = lo + 0.5 * hs * (1 + glx[q]);
                double t = (k - a) / (b - a);
                double j[3];
                sph_bessel_024(k * r, &j[0], &j[1], &j[2]);
                double w = 0.5 * hs * glw[q] * k * k;
                for (l = 0; l < 3; ++l) part[l] += w * (g[l][f] + (g[l][f + 1] - g[l][f]) * t) * j[l];
            }
            for (l = 0; l < 3; ++l) { /* Kahan */
                double y = part[l] - comp[l], t = acc[l] + y;
                comp[l] = (t - acc[l]) - y;
                acc[l] = t;
            }
        }
    }
    double c = 1.0 / (2 * M_PI * M_PI);
    out[0] = c * acc[0];
    out[1] = -c * acc[1]; /* i^2 */
    out[2] = c * acc[2];  /* i^4 */
}

double orc_hankel_table(int n, const double *k, const double *pk, int ell, double r) {
    /* plain transform of a tabulated P(k), refined to NFINE_SUB sub-nodes by a cubic spline
     * in ln k exactly as kspace_setup does */
    double *lnk = (double *)malloc(sizeof(double) * n), out[3];
    int nf = (n - 1) * NFINE_SUB + 1, i, s;
    double *kf = (double *)malloc(sizeof(double) * nf), *pf = (double *)malloc(sizeof(double) * nf);
    spline_t sp;
    for (i = 0; i < n; ++i) lnk[i] = log(k[i]);
    spline_make(&sp, n, lnk, pk);
    for (i = 0; i < n - 1; ++i)
        for (s = 0; s < NFINE_SUB; ++s) {
            double kk = k[i] + (k[i + 1] - k[i]) * s / NFINE_SUB;
            kf[i * NFINE_SUB + s] = kk;
            pf[i * NFINE_SUB + s] = s == 0 ? pk[i] : spline_eval(&sp, log(kk));
        }
    kf[nf - 1] = k[n - 1];
    pf[nf - 1] = pk[n - 1];
    hankel_brute(nf, kf, pf, pf, pf, r, out);
    spline_free(&sp);
    free(lnk); free(kf); free(pf);
    return out[ell / 2];
}

int orc_transform_nr(const orc_model *m) { return m->nr; }
double orc_transform_r0(const orc_model *m) { return m->r0; }
double orc_transform_dr(const orc_model *m) { return m->dr; }

/* transform of one component at radial nodes 0, jstep, 2 jstep, ...; out[ell*nr + j] (others untouched) */
static void transform_component(orc_model *m, const double *p, double z, int component, int jstep, double *out) {
    int lmaxi = m->d.ell_max / 2, l, i, f, j;
    double *dl = (double *)malloc(sizeof(double) * m->nk);
    double *g = (double *)calloc((size_t)3 * m->nfine, sizeof(double));
    const double *pf = component == 0 ? m->pf_pk : m->pf_nw;
    kspace_set_state(m, p, z, component);
    for (l = 0; l <= lmaxi; ++l) {
        for (i = 0; i < m->nk; ++i) dl[i] = d_multipole(m, p, component, i, 2 * l);
        for (f = 0; f < m->nfine; ++f) g[(size_t)l * m->nfine + f] = pf[f] * lagrange4(m, dl, log(m->kf[f]));
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) private(l)
#endif
    for (j = 0; j < m->nr; j += jstep) {
        double o3[3];
        hankel_brute(m->nfine, m->kf, g, g + m->nfine, g + 2 * (size_t)m->nfine, m->tab_r[j], o3);
        for (l = 0; l < 3; ++l) out[l * m->nr + j] = o3[l];
    }
    free(dl);
    free(g);
}

int orc_transform_sub(orc_model *m, const double *params, double z_trigger, int jstep, double *out) {
    if (m->d.kind != BF_MODEL_KSPACE) { set_err("orc_transform: not a k-space model"); return -1; }
    transform_component(m, params, z_trigger, 0, jstep, out);
    transform_component(m, params, z_trigger, 1, jstep, out + 3 * (size_t)m->nr);
    return 0;
}

int orc_transform(orc_model *m, const double *params, double z_trigger, double *out) {
    /* out[(comp*3 + ell)*nr + j]; BaoKSpaceCorrelationModel.cc:333-380 */
    if (m->d.kind != BF_MODEL_KSPACE) { set_err("orc_transform: not a k-space model"); return -1; }
    transform_component(m, params, z_trigger, 0, 1, out);
    transform_component(m, params, z_trigger, 1, 1, out + 3 * (size_t)m->nr);
    return 0;
}

static void kspace_update_tables(orc_model *m, const double *p, double z) {
    int c, l, n = m->d.npar;
    if (m->have_tables && m->last_ztrig == z && memcmp(m->last_params, p, sizeof(double) * n) == 0) return;
    double *buf = (double *)malloc(sizeof(double) * 6 * m->nr);
    orc_transform(m, p, z, buf);
    for (c = 0; c < 2; ++c)
        for (l = 0; l < 3; ++l) {
            memcpy(m->tab_y[c][l], buf + (size_t)(c * 3 + l) * m->nr, sizeof(double) * m->nr);
            orc_cspline_coeffs(m->nr, m->tab_r, m->tab_y[c][l], m->tab_c[c][l]);
        }
    free(buf);
    memcpy(m->last_params, p, sizeof(double) * n);
    m->last_ztrig = z;
    m->have_tables = 1;
}

static double kspace_correlation(const orc_model *m, int c, double r, double mu) {
    /* cosmo::DistortedPowerCorrelation::getCorrelation: sum_ell L_ell(mu) xi_ell(r) */
    double musq = mu * mu;
    double L2 = (-1 + 3 * musq) / 2., L4 = (3 + musq * (-30 + 35 * musq)) / 8.;
    return orc_cspline_eval(m->nr, m->tab_r, m->tab_y[c][0], m->tab_c[c][0], r)
         + L2 * orc_cspline_eval(m->nr, m->tab_r, m->tab_y[c][1], m->tab_c[c][1], r)
         + L4 * orc_cspline_eval(m->nr, m->tab_r, m->tab_y[c][2], m->tab_c[c][2], r);
}

static double radiation_k(const orc_model *m, const double *p, double r, double mu, double z) {
    /* BaoKSpaceCorrelationModel.cc:441-455 */
    const bf_model_desc *d = &m->d;
    double rad_strength = p[d->idx_rad], rad_aniso = p[d->idx_rad + 1];
    double rad_mean_free = p[d->idx_rad + 2], rad_lifetime = p[d->idx_rad + 3];
    double rad = rad_strength / (r * r);
    if (rad_aniso > 0) rad *= (1 - rad_aniso * (1 - mu * mu));
    if (rad_mean_free > 0) rad *= exp(-r / rad_mean_free);
    if (rad_lifetime > 0) rad *= exp(-(r * (1 - mu) / (1 + z)) / rad_lifetime);
    return rad;
}

/* BaoKSpaceCorrelationModel::_evaluate up to (not including) the distortion matrix, for one
 * point; zero_outside = the range checks of the distortion loop (:466-469,:485-488) */
static double kspace_point(const orc_model *m, const double *p, const internal_t *in, double r, double mu,
                           double z, int index, int grid_point, int *status) {
    const bf_model_desc *d = &m->d;
    int cross = (d->flags & BF_FLAG_CROSS_CORRELATION) != 0;
    double beta = p[d->idx_beta], bb = p[d->idx_bb];
    double bias = bb / (1 + beta);
    if (d->flags & BF_FLAG_COMBINED_BIAS) bias = p[d->idx_beta_bias] / beta;
    double biasSq = cross ? bias * p[d->idx_bias2] : bias * bias;
    double gammaBias = p[d->idx_gamma_bias];
    if (grid_point && (r < m->rlo || r > m->rhi)) return 0;
    double biasSqz = orc_redshift_evolution(biasSq, gammaBias, z, d->zref);
    double ampl = p[d->idx_bao], scale = p[d->idx_bao + 1], scale_parallel = p[d->idx_bao + 2];
    double scale_perp = p[d->idx_bao + 3], gamma_scale = p[d->idx_bao + 4];
    if (d->flags & BF_FLAG_COMBINED_SCALE) scale_perp = p[d->idx_comb_scale] * scale_parallel;
    double rBAO, muBAO, scalez;
    if (d->flags & BF_FLAG_ANISOTROPIC) {
        double apar = orc_redshift_evolution(scale_parallel, gamma_scale, z, d->zref);
        double aperp = orc_redshift_evolution(scale_perp, gamma_scale, z, d->zref);
        double musq = mu * mu;
        scalez = sqrt(apar * apar * musq + aperp * aperp * (1 - musq));
        rBAO = r * scalez;
        muBAO = apar * mu / scalez;
    } else {
        scalez = orc_redshift_evolution(scale, gamma_scale, z, d->zref);
        rBAO = r * scalez;
        muBAO = mu;
    }
    if (grid_point) {
        if (rBAO < m->rlo || rBAO > m->rhi) return 0;
    } else {
        if (scalez < d->dilmin) *status |= BF_STATUS_DILATION_MIN; /* :420-423 */
        else if (scalez > d->dilmax) *status |= BF_STATUS_DILATION_MAX; /* :424-427 */
    }
    double peak = kspace_correlation(m, 0, rBAO, muBAO);
    double smooth = (d->flags & BF_FLAG_DECOUPLED) ? kspace_correlation(m, 1, r, mu) : kspace_correlation(m, 1, rBAO, muBAO);
    double xi = biasSqz * (ampl * peak + smooth);
    if (d->metal_kind != BF_METAL_NONE) xi += metal_eval(m, p, in, r, mu, z, index);
    if ((d->flags & BF_FLAG_RADIATION_MODEL) && r > 0) xi += radiation_k(m, p, r, mu, z);
    if (grid_point) { /* pre-matrix broadband :512-517 */
        if (d->bb[BF_BB_DM_DIST_MUL].enabled) xi *= 1 + broadband_eval(&d->bb[BF_BB_DM_DIST_MUL], p, r, mu, z);
        if (d->bb[BF_BB_DM_DIST_ADD].enabled)
            xi += orc_redshift_evolution(broadband_eval(&d->bb[BF_BB_DM_DIST_ADD], p, r, mu, z), gammaBias, z, d->zref);
    }
    return xi;
}


/* ------------------------------------------------------------------------------------ */
/* BaoKSpaceFftCorrelationModel (baofit/BaoKSpaceFftCorrelationModel.cc)                 */
/* ------------------------------------------------------------------------------------ */

/* tabulated P(k) of one input table (0 fiducial, 1 no-wiggles): cubic spline in ln k inside the
 * table, power law outside */
static double fft_table_power(const orc_model *m, int which, double k) {
    const bf_model_desc *d = &m->d;
    const double *pk = which == 0 ? d->pk_fid : d->pk_nw;
    int n = m->n_pk;
    if (k < d->pk_k[0]) return pk[0] * pow(k / d->pk_k[0], m->pk_slope[which][0]);
    if (k > d->pk_k[n - 1]) return pk[n - 1] * pow(k / d->pk_k[n - 1], m->pk_slope[which][1]);
    return spline_eval(which == 0 ? &m->pk_fid : &m->pk_nw, log(k));
}
static double fft_power(const orc_model *m, int comp, double k) {
    /* :81 peak = fid - nw (createDelta), smooth = nw */
    double pn = fft_table_power(m, 1, k);
    return comp == 0 ? fft_table_power(m, 0, k) - pn : pn;
}

static double fft_distortion(const orc_model *m, const double *p, double k, double mu_k, double pk) {
    /* _evaluateKSpaceDistortion, BaoKSpaceFftCorrelationModel.cc:120-153 */
    const bf_model_desc *d = &m->d;
    int cross = (d->flags & BF_FLAG_CROSS_CORRELATION) != 0;
    double mu2 = mu_k * mu_k;
    double tracer1 = 1 + m->betaz * mu2;
    double tracer2 = cross ? 1 + m->beta2z * mu2 : tracer1;
    double linear = tracer1 * tracer2;
    double snl2 = m->snl_par2 * mu2 + m->snl_perp2 * (1 - mu2);
    double nonlinear = exp(-0.5 * snl2 * k * k);
    double kpar = fabs(k * mu_k);
    double kc = p[d->idx_cont], pc = p[d->idx_cont + 1];
    double contdistortion;
    if (d->flags & BF_FLAG_NO_DISTORTION) contdistortion = 1;
    else if (d->flags & BF_FLAG_DISTORTION_ALT) contdistortion = tanh(pow(kpar / kc, pc));
    else {
        double k1 = pow(kpar / kc + 1, 0.75);
        contdistortion = pow((k1 - 1 / k1) / (k1 + 1 / k1), pc);
    }
    double nonlinearcorr = nl_correction(m, p, k, mu_k, pk, m->zeff);
    if (cross) {
        contdistortion = sqrt(contdistortion);
        nonlinearcorr = sqrt(nonlinearcorr);
    }
    return contdistortion * nonlinear * nonlinearcorr * linear;
}

/* the members _evaluate sets before a transform (:178-188,:200-211), for a transform triggered at z */
static void fft_set_state(orc_model *m, const double *p, double z, int component) {
    const bf_model_desc *d = &m->d;
    double beta = p[d->idx_beta], gammaBeta = p[d->idx_gamma_beta];
    m->zeff = z; /* :169, the FFT model has no zeff option */
    m->betaz = orc_redshift_evolution(beta, gammaBeta, z, d->zref);
    m->beta2z = 0;
    if (d->flags & BF_FLAG_CROSS_CORRELATION)
        m->beta2z = orc_redshift_evolution(p[d->idx_bb2] / p[d->idx_bias2], gammaBeta, z, d->zref);
    double snlPerp = p[d->idx_nl], snlPar = snlPerp * p[d->idx_nl + 1];
    m->snl_perp2 = snlPerp * snlPerp;
    m->snl_par2 = snlPar * snlPar;
    if (component == 1 && !(d->flags & BF_FLAG_NL_BROADBAND)) m->snl_perp2 = m->snl_par2 = 0; /* :207-210 */
}

double orc_fft_distortion_power(orc_model *m, const double *params, double z, int component, double k, double muk) {
    if (k <= 0) return 0;
    fft_set_state(m, params, z, component);
    double pk = fft_power(m, component, k);
    return fft_distortion(m, params, k, muk, pk) * pk;
}

/* cosmo::DistortedPowerCorrelationFft::transform restated (UNVERIFIED dep, see oracle.h): the box has
 * nx x ny x nz cells of side delta, y is the line of sight, the k grid is the FFT grid of the box,
 *   xi(x) = 1/(nx ny nz delta^3) sum_k D(k, k_y/k) P(k) exp(i k.x),     the k = 0 term dropped,
 * and only the z = 0 plane is kept. D P is even in every k component, so the plane is the cosine
 * series over the non-negative frequencies with weight 2 on the interior ones. */
static void fft_transform_component(orc_model *m, const double *p, double z, int comp, double *plane) {
    const int nx = m->fnx, ny = m->fny, nz = m->fnz, Mx = m->fmx, My = m->fmy, Mz = m->fmz;
    const double dkx = 2 * M_PI / (nx * m->fdelta), dky = 2 * M_PI / (ny * m->fdelta), dkz = 2 * M_PI / (nz * m->fdelta);
    double *G = (double *)calloc((size_t)My * Mx, sizeof(double));
    double *X = (double *)calloc((size_t)My * m->fnpx, sizeof(double));
    int my;
    fft_set_state(m, p, z, comp);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (my = 0; my < My; ++my) {
        int mx, mz, ix;
        double ky = my * dky;
        for (mx = 0; mx < Mx; ++mx) {
            double kx = mx * dkx, s = 0;
            for (mz = 0; mz < Mz; ++mz) {
                double kz = mz * dkz, k = sqrt(kx * kx + ky * ky + kz * kz);
                if (k == 0) continue;
                double w = (mz == 0 || 2 * mz == nz) ? 1.0 : 2.0;
                double pk = fft_power(m, comp, k);
                s += w * fft_distortion(m, p, k, ky / k, pk) * pk;
            }
            G[(size_t)my * Mx + mx] = s;
        }
        for (ix = 0; ix < m->fnpx; ++ix) {
            double s = 0;
            for (mx = 0; mx < Mx; ++mx) {
                double w = (mx == 0 || 2 * mx == nx) ? 1.0 : 2.0;
                s += w * cos(2 * M_PI * (double)(((long long)mx * ix) % nx) / nx) * G[(size_t)my * Mx + mx];
            }
            X[(size_t)my * m->fnpx + ix] = s;
        }
    }
    double norm = 1.0 / ((double)nx * ny * nz * m->fdelta * m->fdelta * m->fdelta);
    int iy;
    for (iy = 0; iy < m->fnpy; ++iy) {
        int ix;
        for (ix = 0; ix < m->fnpx; ++ix) {
            double s = 0;
            for (my = 0; my < My; ++my) {
                double w = (my == 0 || 2 * my == ny) ? 1.0 : 2.0;
                s += w * cos(2 * M_PI * (double)(((long long)my * iy) % ny) / ny) * X[(size_t)my * m->fnpx + ix];
            }
            plane[(size_t)iy * m->fnpx + ix] = norm * s;
        }
    }
    free(G);
    free(X);
}

int orc_fft_plane_dims(const orc_model *m, int *npx, int *npy) {
    if (m->d.kind != BF_MODEL_KSPACE_FFT) return -1;
    *npx = m->fnpx;
    *npy = m->fnpy;
    return 0;
}

int orc_fft_planes(orc_model *m, const double *params, double z_trigger, double *out) {
    /* out[(comp*npy + iy)*npx + ix] */
    if (m->d.kind != BF_MODEL_KSPACE_FFT) { set_err("orc_fft_planes: not an FFT model"); return -1; }
    fft_transform_component(m, params, z_trigger, 0, out);
    fft_transform_component(m, params, z_trigger, 1, out + (size_t)m->fnpx * m->fnpy);
    return 0;
}

static void fft_update_tables(orc_model *m, const double *p, double z) {
    /* :191-214: the reference re-transforms when the broadening, continuum, nlcorr parameters or
     * beta changed; here: whenever anything D(k,mu) reads changed (same thing for every BASELINE
     * config, see DESIGN.md "FFT model") */
    const bf_model_desc *d = &m->d;
    double key[16];
    int n = 0, i;
    memset(key, 0, sizeof(key));
    key[n++] = z; key[n++] = p[d->idx_beta]; key[n++] = p[d->idx_gamma_beta];
    key[n++] = p[d->idx_nl]; key[n++] = p[d->idx_nl + 1]; key[n++] = p[d->idx_cont]; key[n++] = p[d->idx_cont + 1];
    if (d->flags & BF_FLAG_CROSS_CORRELATION) { key[n++] = p[d->idx_bias2]; key[n++] = p[d->idx_bb2]; }
    if (d->flags & BF_FLAG_FIT_NL_CORRECTION) { key[n++] = p[d->idx_nlcorr]; key[n++] = p[d->idx_nlcorr + 1]; }
    if (m->fhave && memcmp(key, m->fkey, sizeof(key)) == 0) return;
    for (i = 0; i < 2; ++i) fft_transform_component(m, p, z, i, m->fplane[i]);
    memcpy(m->fkey, key, sizeof(key));
    m->fhave = 1;
}

/* cosmo::DistortedPowerCorrelationFft::getCorrelation(r,mu) restated: bicubic (Catmull-Rom)
 * interpolation on the plane at (r_perp, |r_par|); the plane is even in both coordinates, which
 * supplies the node at index -1 (UNVERIFIED dep). Returns NaN and flags outside the kept plane. */
static double fft_correlation(const orc_model *m, int c, double r, double mu, int *status) {
    double x = r * sqrt(1 - mu * mu) / m->fdelta, y = fabs(r * mu) / m->fdelta;
    int ix = (int)floor(x), iy = (int)floor(y), a, b;
    if (ix + 2 > m->fnpx - 1 || iy + 2 > m->fnpy - 1) { *status |= BF_STATUS_FFT_RANGE; return NAN; }
    double tx = x - ix, ty = y - iy, col[4];
    for (b = 0; b < 4; ++b) {
        int jy = abs(iy - 1 + b);
        const double *row = m->fplane[c] + (size_t)jy * m->fnpx;
        double v[4];
        for (a = 0; a < 4; ++a) v[a] = row[abs(ix - 1 + a)];
        col[b] = catmull(v[0], v[1], v[2], v[3], tx);
    }
    return catmull(col[0], col[1], col[2], col[3], ty);
}

static double fft_zcorr(const orc_model *m, double r, double mu, double z) {
    /* :165-168 */
    const bf_model_desc *d = &m->d;
    if (d->zcorr0 > 0) {
        double rpar = fabs(r * mu) / 100.;
        z = d->zcorr0 + d->zcorr1 * rpar + d->zcorr2 * rpar * rpar;
    }
    return z;
}

/* BaoKSpaceFftCorrelationModel::_evaluate, :155-250; z_trigger = corrected z of the bin whose
 * evaluation triggers the transform */
static double fft_evaluate(orc_model *m, const double *p, double r, double mu, double z, double z_trigger, int *status) {
    const bf_model_desc *d = &m->d;
    int cross = (d->flags & BF_FLAG_CROSS_CORRELATION) != 0;
    double beta = p[d->idx_beta], bb = p[d->idx_bb];
    double bias = bb / (1 + beta);
    double biasSq = cross ? bias * p[d->idx_bias2] : bias * bias;
    double gammaBias = p[d->idx_gamma_bias];
    z = fft_zcorr(m, r, mu, z);
    biasSq = orc_redshift_evolution(biasSq, gammaBias, z, d->zref);
    fft_update_tables(m, p, z_trigger);
    double ampl = p[d->idx_bao], scale = p[d->idx_bao + 1], scale_parallel = p[d->idx_bao + 2];
    double scale_perp = p[d->idx_bao + 3], gamma_scale = p[d->idx_bao + 4];
    double rBAO, muBAO;
    if (d->flags & BF_FLAG_ANISOTROPIC) {
        double apar = orc_redshift_evolution(scale_parallel, gamma_scale, z, d->zref);
        double aperp = orc_redshift_evolution(scale_perp, gamma_scale, z, d->zref);
        double musq = mu * mu;
        scale = sqrt(apar * apar * musq + aperp * aperp * (1 - musq));
        rBAO = r * scale;
        muBAO = apar * mu / scale;
    } else {
        scale = orc_redshift_evolution(scale, gamma_scale, z, d->zref);
        rBAO = r * scale;
        muBAO = mu;
    }
    double peak = fft_correlation(m, 0, rBAO, muBAO, status);
    double smooth = (d->flags & BF_FLAG_DECOUPLED) ? fft_correlation(m, 1, r, mu, status) : fft_correlation(m, 1, rBAO, muBAO, status);
    double xi = biasSq * (ampl * peak + smooth);
    if (d->bb[BF_BB_DIST_MUL].enabled) xi *= 1 + broadband_eval(&d->bb[BF_BB_DIST_MUL], p, r, mu, z);
    if (d->bb[BF_BB_DIST_ADD].enabled)
        xi += orc_redshift_evolution(broadband_eval(&d->bb[BF_BB_DIST_ADD], p, r, mu, z), gammaBias, z, d->zref);
    return xi;
}

/* corrected redshift of the bin that triggers the transform: the first kept bin, after the
 * velocity shift of AbsCorrelationModel::evaluate (AbsCorrelationModel.cc:25) */
static double fft_trigger_z(const orc_model *m, const double *p, double r, double mu, double z) {
    if (m->d.idx_delta_v >= 0) velocity_shift(m, p, &r, &mu, z);
    return fft_zcorr(m, r, mu, z);
}

/* ------------------------------------------------------------------------------------ */
/* AbsCorrelationModel::evaluate, AbsCorrelationModel.cc:21-41                           */
/* ------------------------------------------------------------------------------------ */

static double evaluate_with_ztrig(orc_model *m, const orc_data *data, const double *p, double r, double mu,
                                  double z, int index, double z_trigger, int *status) {
    const bf_model_desc *d = &m->d;
    internal_t in;
    update_internal(m, p, &in);
    if (d->idx_delta_v >= 0) velocity_shift(m, p, &r, &mu, z);
    if (d->kind == BF_MODEL_RSPACE) return rspace_evaluate(m, p, &in, r, mu, z, index);
    if (d->kind == BF_MODEL_KSPACE_FFT) return fft_evaluate(m, p, r, mu, z, z_trigger, status);
    /* k-space */
    kspace_update_tables(m, p, z_trigger);
    double xi = kspace_point(m, p, &in, r, mu, z, index, 0, status);
    if (m->dist_matrix && index >= 0) { /* :457-527 */
        int nb = d->dist_matrix_order, bin;
        xi = 0;
        for (bin = 0; bin < nb; ++bin) {
            double rb = data->grid_r[bin], mub = data->grid_mu[bin], zb = data->grid_z[bin];
            if (d->idx_delta_v >= 0) velocity_shift(m, p, &rb, &mub, zb); /* AbsCorrelationModel.cc:26-37 */
            double xiu = kspace_point(m, p, &in, rb, mub, zb, bin, 1, status);
            xi += m->dist_matrix[(size_t)index * nb + bin] * xiu;
        }
    }
    if (d->bb[BF_BB_DIST_MUL].enabled) xi *= 1 + broadband_eval(&d->bb[BF_BB_DIST_MUL], p, r, mu, z);
    if (d->bb[BF_BB_DIST_ADD].enabled)
        xi += orc_redshift_evolution(broadband_eval(&d->bb[BF_BB_DIST_ADD], p, r, mu, z), p[d->idx_gamma_bias], z, d->zref);
    return xi;
}

double orc_evaluate(orc_model *m, const orc_data *data, const double *params, double r, double mu, double z,
                    int index, int *status) {
    int st = 0;
    double ztrig = m->d.kind == BF_MODEL_KSPACE_FFT ? fft_trigger_z(m, params, r, mu, z) : z;
    double v = evaluate_with_ztrig(m, data, params, r, mu, z, index, ztrig, &st);
    if (status) *status = st;
    return v;
}

double orc_evaluate_multipole(orc_model *m, const double *params, double r, int ell, double z, int nmu) {
    /* AbsCorrelationModel.cc:65-79: (2l+1)/2 Int_-1^1 L_l(mu) xi(r,mu) dmu; delta-v has no effect */
    double *x = (double *)malloc(sizeof(double) * nmu), *w = (double *)malloc(sizeof(double) * nmu), s = 0;
    internal_t in;
    int q, st = 0;
    gauss_legendre_01(nmu, x, w);
    update_internal(m, params, &in);
    for (q = 0; q < nmu; ++q) {
        double mu = 2 * x[q] - 1, v;
        if (m->d.kind == BF_MODEL_RSPACE) v = rspace_evaluate(m, params, &in, r, mu, z, -1);
        else if (m->d.kind == BF_MODEL_KSPACE_FFT) v = fft_evaluate(m, params, r, mu, z, fft_zcorr(m, r, 0, z), &st);
        else { kspace_update_tables(m, params, z); v = kspace_point(m, params, &in, r, mu, z, -1, 0, &st); }
        s += 2 * w[q] * orc_legendre(ell, mu) * v;
    }
    free(x);
    free(w);
    return 0.5 * (2 * ell + 1) * s;
}

int orc_undistorted(orc_model *m, const orc_data *d, const double *params, double *xiu) {
    int bin, st = 0;
    internal_t in;
    if (m->d.kind != BF_MODEL_KSPACE || !m->dist_matrix) { set_err("orc_undistorted: needs a distortion matrix"); return -1; }
    update_internal(m, params, &in);
    kspace_update_tables(m, params, d->z[0]);
    for (bin = 0; bin < m->d.dist_matrix_order; ++bin) {
        double rb = d->grid_r[bin], mub = d->grid_mu[bin], zb = d->grid_z[bin];
        if (m->d.idx_delta_v >= 0) velocity_shift(m, params, &rb, &mub, zb);
        xiu[bin] = kspace_point(m, params, &in, rb, mub, zb, bin, 1, &st);
    }
    return st;
}

int orc_predict(orc_model *m, const orc_data *d, const double *params, double *pred) {
    /* CorrelationFitter.cc:53-72. The k-space transform is triggered by the first bin visited
     * after a parameter change, i.e. with z of the first kept bin. */
    int i, st = 0;
    double ztrig = d->d.n_data > 0 ? d->z[0] : m->d.zref;
    if (m->d.kind == BF_MODEL_KSPACE_FFT && d->d.n_data > 0) ztrig = fft_trigger_z(m, params, d->r[0], d->mu[0], d->z[0]);
    if (m->d.kind == BF_MODEL_KSPACE && m->dist_matrix) {
        /* evaluate the undistorted grid once, then one matrix row per data bin */
        int nb = m->d.dist_matrix_order;
        double *xiu = (double *)malloc(sizeof(double) * nb);
        internal_t in;
        update_internal(m, params, &in);
        kspace_update_tables(m, params, ztrig);
        for (i = 0; i < nb; ++i) {
            double rb = d->grid_r[i], mub = d->grid_mu[i], zb = d->grid_z[i];
            if (m->d.idx_delta_v >= 0) velocity_shift(m, params, &rb, &mub, zb);
            xiu[i] = kspace_point(m, params, &in, rb, mub, zb, i, 1, &st);
        }
        for (i = 0; i < d->d.n_data; ++i) {
            double r = d->r[i], mu = d->mu[i], z = d->z[i], xi = 0;
            int bin, index = d->index[i];
            if (m->d.idx_delta_v >= 0) velocity_shift(m, params, &r, &mu, z);
            /* the dilation check of the data bin itself still applies (:420-427) */
            (void)kspace_point(m, params, &in, r, mu, z, index, 0, &st);
            for (bin = 0; bin < nb; ++bin) xi += m->dist_matrix[(size_t)index * nb + bin] * xiu[bin];
            if (m->d.bb[BF_BB_DIST_MUL].enabled) xi *= 1 + broadband_eval(&m->d.bb[BF_BB_DIST_MUL], params, r, mu, z);
            if (m->d.bb[BF_BB_DIST_ADD].enabled)
                xi += orc_redshift_evolution(broadband_eval(&m->d.bb[BF_BB_DIST_ADD], params, r, mu, z),
                                             params[m->d.idx_gamma_bias], z, m->d.zref);
            pred[i] = xi;
        }
        free(xiu);
        return st;
    }
    if (d->d.binning_type == BF_BINNING_MULTIPOLE) {
        /* CorrelationFitter.cc:66-69 -> AbsCorrelationModel::evaluate(r, multipole, z, ...), :43-49,65-79 */
        int nmu = d->d.nmu_project > 0 ? d->d.nmu_project : 16, q;
        double *x = (double *)malloc(sizeof(double) * nmu), *w = (double *)malloc(sizeof(double) * nmu);
        if (m->dist_matrix || m->d.idx_delta_v >= 0) { /* the multipole overload applies no velocity shift (:43-49) */
            set_err("orc_predict: multipole data with a distortion matrix / cross-correlation is not restated");
            free(x); free(w);
            return -1;
        }
        if (m->d.kind == BF_MODEL_KSPACE_FFT) ztrig = fft_trigger_z(m, params, d->r[0], 0.0, d->z[0]); /* :73, mu = 0 */
        gauss_legendre_01(nmu, x, w);
        for (i = 0; i < d->d.n_data; ++i) {
            int ell = (int)floor(d->mu[i] + 0.5);
            double s = 0;
            for (q = 0; q < nmu; ++q) {
                double mu = 2 * x[q] - 1;
                s += 2 * w[q] * orc_legendre(ell, mu) * evaluate_with_ztrig(m, d, params, d->r[i], mu, d->z[i], d->index[i], ztrig, &st);
            }
            pred[i] = 0.5 * (2 * ell + 1) * s;
        }
        free(x);
        free(w);
        return st;
    }
    for (i = 0; i < d->d.n_data; ++i)
        pred[i] = evaluate_with_ztrig(m, d, params, d->r[i], d->mu[i], d->z[i], d->index[i], ztrig, &st);
    return st;
}

double orc_chi2(const orc_data *d, const double *pred, const double *data_override) {
    /* likely::BinnedData::chiSquare: delta^T C^-1 delta */
    int n = d->d.n_data, i, j;
    const double *dat = data_override ? data_override : d->data;
    double *delta = (double *)malloc(sizeof(double) * n), chi2 = 0;
    for (i = 0; i < n; ++i) delta[i] = dat[i] - pred[i];
    if (d->d.cov_is_inverse) {
        for (i = 0; i < n; ++i) {
            double s = 0;
            for (j = 0; j < n; ++j) s += d->cov[(size_t)i * n + j] * delta[j];
            chi2 += delta[i] * s;
        }
    } else {
        /* y = L^-1 delta; chi2 = y.y */
        for (i = 0; i < n; ++i) {
            double s = delta[i];
            const double *Li = d->L + (size_t)i * n;
            for (j = 0; j < i; ++j) s -= Li[j] * delta[j];
            delta[i] = s / Li[i];
            chi2 += delta[i] * delta[i];
        }
    }
    free(delta);
    return chi2;
}

/* copies the model state that evaluation mutates, so that OpenMP threads do not share it */
static orc_model *model_thread_clone(const orc_model *m) {
    orc_model *c = (orc_model *)malloc(sizeof(orc_model));
    int i, s;
    *c = *m;
    if (m->d.kind == BF_MODEL_KSPACE) {
        for (i = 0; i < 2; ++i)
            for (s = 0; s < 3; ++s) {
                c->tab_y[i][s] = (double *)calloc(m->nr, sizeof(double));
                c->tab_c[i][s] = (double *)calloc(m->nr, sizeof(double));
            }
        c->last_params = (double *)malloc(sizeof(double) * m->d.npar);
        c->have_tables = 0;
    }
    if (m->d.kind == BF_MODEL_KSPACE_FFT) {
        for (i = 0; i < 2; ++i) c->fplane[i] = (double *)calloc((size_t)m->fnpx * m->fnpy, sizeof(double));
        c->fhave = 0;
    }
    return c;
}
static void model_thread_free(orc_model *c) {
    int i, s;
    if (c->d.kind == BF_MODEL_KSPACE) {
        for (i = 0; i < 2; ++i) for (s = 0; s < 3; ++s) { free(c->tab_y[i][s]); free(c->tab_c[i][s]); }
        free(c->last_params);
    }
    if (c->d.kind == BF_MODEL_KSPACE_FFT) { free(c->fplane[0]); free(c->fplane[1]); }
    free(c);
}

int orc_chi2_batch(orc_model *m, const orc_data *d, const double *params, int B, const double *data_override,
                   double *chi2, int *status, double *pred, int ldp, int nthreads) {
    int n = d->d.n_data, npar = m->d.npar;
    if (nthreads < 1) nthreads = 1;
    int outer = nthreads;
#ifdef _OPENMP
    if (m->d.kind == BF_MODEL_KSPACE || m->d.kind == BF_MODEL_KSPACE_FFT) { outer = 1; omp_set_num_threads(nthreads); }
#endif
#ifdef _OPENMP
#pragma omp parallel num_threads(outer)
#endif
    {
        orc_model *mc = model_thread_clone(m);
        double *pb = (double *)malloc(sizeof(double) * n);
        int b, i;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (b = 0; b < B; ++b) {
            int st = orc_predict(mc, d, params + (size_t)b * npar, pb);
            if (status) status[b] = st;
            if (chi2) chi2[b] = st ? NAN : orc_chi2(d, pb, data_override ? data_override + (size_t)b * n : NULL);
            if (pred) for (i = 0; i < n; ++i) pred[(size_t)i * ldp + b] = st ? NAN : pb[i];
        }
        free(pb);
        model_thread_free(mc);
    }
    return 0;
}
