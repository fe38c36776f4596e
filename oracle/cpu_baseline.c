/*
 * cpu_baseline.c -- reference-SHAPED CPU timing baseline of the fit-evaluation hot path (bench.py only).
 *
 * TEST / BENCH INFRASTRUCTURE, never linked into the product. This is NOT the parity oracle: the parity oracle
 * (oracle.c) evaluates every k -> r transform by brute-force quadrature so as to share nothing with the GPU's
 * Bessel-weight GEMM, which makes it ~10^3 times slower per evaluation than the reference binary. A timing baseline
 * has to cost what the reference costs, so this file keeps the reference's control flow and swaps only the
 * transform for one of the reference's order of work:
 *
 *   - per-bin prediction loop, CorrelationFitter::getPrediction (baofit/CorrelationFitter.cc:53-72), one model
 *     evaluation per kept data bin, the undistorted correlation on the whole grid when parameters changed and one
 *     dense matrix row per data bin (baofit/BaoKSpaceCorrelationModel.cc:457-527), chi-square from the factored
 *     covariance (baofit/CorrelationFitter.cc:85) -- all of that IS oracle.c's code, included below verbatim;
 *   - the transform is re-run only when a parameter D(k, mu) reads has changed since the previous evaluation, peak
 *     and no-wiggles components decided separately (baofit/BaoKSpaceCorrelationModel.cc:333-380): beta, gamma-beta,
 *     bias2, beta2*bias2, SigmaNL-perp, 1+f, and the bin-smoothing / HCD / UV / smoothing / nlcorr blocks;
 *   - one transform = the multipoles of D(k, mu) on the reference's coarse k grid with its 20 mu nodes
 *     (n_k x 20 evaluations of _evaluateKSpaceDistortion, :215-283, :347) followed by a mat-vec against
 *     pre-computed Bessel weights [2 components][3 ell][n_r][n_k] -- about 1.1 Mflop, the same order as the
 *     FFT-based multipole transform of the missing cosmo library, and certainly not more (so the baseline errs on
 *     the fast side: the reference's own header comments quote 25-39 s for a ~10^3-evaluation k-space fit).
 * The weights come from the product library's host-side set-up (bf_kspace_transform_weights, no GPU involved);
 * results are checked against the GPU path in bench.py (cpu_baseline.max_rel_chi2_diff_vs_gpu).
 */
#include "oracle.c"

typedef struct cpub {
    orc_model *m;
    const orc_data *d;
    int nr, nk;
    double *W; /* [2][3][nr][nk] */
} cpub;

cpub *cpub_create(orc_model *m, const orc_data *d, int nr, int nk, const double *weights) {
    cpub *c = (cpub *)calloc(1, sizeof(cpub));
    c->m = m;
    c->d = d;
    if (m->d.kind == BF_MODEL_KSPACE) {
        if (!weights || nr != m->nr || nk != m->nk) { set_err("cpub_create: transform weights do not match the model grids"); free(c); return NULL; }
        c->nr = nr;
        c->nk = nk;
        c->W = dupd(weights, (size_t)6 * nr * nk);
    }
    return c;
}

void cpub_destroy(cpub *c) {
    if (!c) return;
    free(c->W);
    free(c);
}

/* cosmo::DistortedPowerCorrelation::transform stand-in: multipoles of D on the coarse grid, then the weight mat-vec */
static void cpub_transform(const cpub *c, orc_model *mc, const double *p, double z, int component, double *dl) {
    int lmaxi = mc->d.ell_max / 2, l, i, j;
    kspace_set_state(mc, p, z, component);
    for (l = 0; l < 3; ++l) {
        double *y = mc->tab_y[component][l];
        if (l > lmaxi) {
            for (j = 0; j < c->nr; ++j) y[j] = 0;
        } else {
            const double *W = c->W + (size_t)(component * 3 + l) * c->nr * c->nk;
            for (i = 0; i < c->nk; ++i) dl[i] = d_multipole(mc, p, component, i, 2 * l);
            for (j = 0; j < c->nr; ++j) {
                const double *w = W + (size_t)j * c->nk;
                double s = 0;
                for (i = 0; i < c->nk; ++i) s += w[i] * dl[i];
                y[j] = s;
            }
        }
        orc_cspline_coeffs(c->nr, mc->tab_r, y, mc->tab_c[component][l]);
    }
}

static int changed(const double *p, const double *q, int idx, int n) {
    int k;
    if (idx < 0) return 0;
    for (k = 0; k < n; ++k)
        if (p[idx + k] != q[idx + k]) return 1;
    return 0;
}

/* chi2[b] of B parameter vectors; every thread walks a contiguous block of vectors in order, like one reference process
 * would, and re-transforms only when a D-relevant parameter differs from the previous vector of its block.
 * transforms (may be NULL) receives the number of component transforms actually run. */
int cpub_chi2_batch(cpub *c, const double *params, int B, double *chi2, int *status, int nthreads, long long *transforms) {
    const orc_data *d = c->d;
    const bf_model_desc *md = &c->m->d;
    int n = d->d.n_data, npar = md->npar;
    long long ntr = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
    omp_set_max_active_levels(1);
#pragma omp parallel num_threads(nthreads) reduction(+ : ntr)
#endif
    {
        int tid = 0, nt = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        orc_model *mc = model_thread_clone(c->m);
        double *pb = (double *)malloc(sizeof(double) * n), *dl = (double *)malloc(sizeof(double) * (c->nk > 0 ? c->nk : 1));
        const double *prev = NULL;
        int b0 = (int)((long long)B * tid / nt), b1 = (int)((long long)B * (tid + 1) / nt), b;
        double ztrig = n > 0 ? d->z[0] : md->zref;
        for (b = b0; b < b1; ++b) {
            const double *p = params + (size_t)b * npar;
            if (md->kind == BF_MODEL_KSPACE) {
                /* baofit/BaoKSpaceCorrelationModel.cc:333-380 */
                int first = prev == NULL;
                int nl = first || changed(p, prev, md->idx_nl, 2);
                int other = first || changed(p, prev, md->idx_beta, 1) || changed(p, prev, md->idx_gamma_beta, 1) ||
                            ((md->flags & BF_FLAG_CROSS_CORRELATION) && (changed(p, prev, md->idx_bias2, 1) || changed(p, prev, md->idx_bb2, 1))) ||
                            ((md->flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) && changed(p, prev, md->idx_bs, 2)) ||
                            ((md->flags & BF_FLAG_FIT_NL_CORRECTION) && changed(p, prev, md->idx_nlcorr, 2)) ||
                            ((md->flags & (BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT)) && changed(p, prev, md->idx_hcd, 3)) ||
                            ((md->flags & BF_FLAG_UVFLUCTUATION) && changed(p, prev, md->idx_uv, 3)) ||
                            ((md->flags & BF_FLAG_SMOOTH_GAUSS) && changed(p, prev, md->idx_smgaus, 1)) ||
                            ((md->flags & BF_FLAG_SMOOTH_LORENTZ) && changed(p, prev, md->idx_smlor, 1)) ||
                            ((md->flags & BF_FLAG_COMBINED_BIAS) && changed(p, prev, md->idx_beta_bias, 1));
                if (nl || other) { cpub_transform(c, mc, p, ztrig, 0, dl); ++ntr; }
                if (((md->flags & BF_FLAG_NL_BROADBAND) && nl) || other) { cpub_transform(c, mc, p, ztrig, 1, dl); ++ntr; }
                /* the tables are current: let the per-bin code below find them in its cache */
                memcpy(mc->last_params, p, sizeof(double) * npar);
                mc->last_ztrig = ztrig;
                mc->have_tables = 1;
            }
            {
                int st = orc_predict(mc, d, p, pb);
                if (status) status[b] = st;
                chi2[b] = st ? NAN : orc_chi2(d, pb, NULL);
            }
            prev = p;
        }
        free(pb);
        free(dl);
        model_thread_free(mc);
    }
    if (transforms) *transforms = ntr;
    return 0;
}
