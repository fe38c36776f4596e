/*
 * oracle.h -- CPU restatement (plain C) of the reference's fit-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU arm. The product (libbaofit_b200.so) never links
 * or calls it and has no CPU fallback.
 *
 * The reference itself cannot be compiled here (its likely / cosmo / Boost / Minuit2 / GSL
 * dependencies are un-vendored and unpinned, configure.ac:25-43), so there is no oracle/_ref.
 * Each function below cites the reference file:line it follows. Pinning status:
 *   PINNED   r-space model, natural cubic spline with end-point clamp, Legendre weights, norm
 *            factors, bin conventions, chi-square: demo/rmu.theory, demo/multi.theory,
 *            chi2(demo/rmu.data | theory) = 761.405865008929, data/BOSSDR11QSOLyaF.scan
 *            (see tests/test_oracle_golden.py).
 *   UNPINNED k-space transforms (cosmo::DistortedPowerCorrelation is not in the tree: the
 *            quadrature here is our own, anchored only on models/DR9LyaMocksLCDM.{0,2,4}.dat
 *            which cosmocalc produced from DR9LyaMocksLCDM_matterpower.dat), bicubic metal
 *            templates (likely::BiCubicInterpolator), distortion-matrix product and metals
 *            (no fixture ships a .dmat or metal template), the 3-D FFT transform and plane
 *            interpolation of cosmo::DistortedPowerCorrelationFft (our definition is pinned
 *            against numpy.fft.ifftn of the same filled box, tests/test_oracle_golden.py).
 *
 * The input descriptions are the PODs of include/baofit_b200.h (types only).
 */
#ifndef BAOFIT_ORACLE_H
#define BAOFIT_ORACLE_H

#include "../include/baofit_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;
typedef struct orc_data orc_data;

orc_model *orc_model_create(const bf_model_desc *desc);
void orc_model_destroy(orc_model *m);
orc_data *orc_data_create(const bf_data_desc *desc);
void orc_data_destroy(orc_data *d);
const char *orc_last_error(void);

/* natural cubic spline = likely::Interpolator(x,y,"cspline") = gsl_interp_cspline, with the
 * end-point value returned outside the table */
void orc_cspline_coeffs(int n, const double *x, const double *y, double *c);
double orc_cspline_eval(int n, const double *x, const double *y, const double *c, double xq);

double orc_legendre(int ell, double mu);                                  /* BroadbandModel.cc:171-195 */
double orc_redshift_evolution(double p0, double gamma, double z, double zref); /* AbsCorrelationModel.cc:159-161 */

/* AbsCorrelationModel::evaluate(r,mu,z,params,index), AbsCorrelationModel.cc:21-41.
 * data may be NULL for models without a distortion matrix. status receives BF_STATUS_*. */
double orc_evaluate(orc_model *m, const orc_data *data, const double *params,
                    double r, double mu, double z, int index, int *status);
/* multipole evaluation AbsCorrelationModel.cc:65-79 with an nmu-point Gauss-Legendre projection */
double orc_evaluate_multipole(orc_model *m, const double *params, double r, int ell, double z, int nmu);

/* CorrelationFitter::getPrediction, CorrelationFitter.cc:53-72; pred[n_data]; returns status */
int orc_predict(orc_model *m, const orc_data *d, const double *params, double *pred);
/* undistorted correlation on the full grid, BaoKSpaceCorrelationModel.cc:460-521; xiu[n_grid] */
int orc_undistorted(orc_model *m, const orc_data *d, const double *params, double *xiu);
/* BinnedData::chiSquare(pred) (call site CorrelationFitter.cc:85); data_override may be NULL */
double orc_chi2(const orc_data *d, const double *pred, const double *data_override);
/* B vectors: chi2[b]; pred (may be NULL) bin-major pred[i*ldp+b]; nthreads OpenMP threads */
int orc_chi2_batch(orc_model *m, const orc_data *d, const double *params, int B,
                   const double *data_override, double *chi2, int *status,
                   double *pred, int ldp, int nthreads);

/* k-space: D(k,mu_k) of BaoKSpaceCorrelationModel.cc:215-283 for the current parameters */
double orc_kspace_distortion(orc_model *m, const double *params, double z, int component,
                             double k, double muk);
/* brute-force P(k) -> xi_ell(r) transform of component comp (0 peak, 1 no-wiggles) at the
 * model's r nodes; out[3][n_r]; also returns n_r, r0, dr via the getters below */
int orc_transform(orc_model *m, const double *params, double z_trigger, double *out);
/* same, only at radial nodes 0, jstep, 2 jstep, ... (other entries of out untouched) */
int orc_transform_sub(orc_model *m, const double *params, double z_trigger, int jstep, double *out);
int orc_transform_nr(const orc_model *m);
double orc_transform_r0(const orc_model *m);
double orc_transform_dr(const orc_model *m);
/* plain Hankel transform xi_ell(r) = i^ell/(2 pi^2) Int k^2 j_ell(kr) P(k) dk of a tabulated
 * P(k) (cubic spline in ln k), used to anchor the quadrature on models/<name>.<ell>.dat */
double orc_hankel_table(int n, const double *k, const double *pk, int ell, double r);

/* 3-D FFT model (BaoKSpaceFftCorrelationModel.cc): D(k,mu_k) P_comp(k) of :120-153 for the current
 * parameters, and the two planes xi_comp(ix*spacing, iy*spacing) of the z = 0 slice of the box,
 * out[(comp*npy + iy)*npx + ix] */
double orc_fft_distortion_power(orc_model *m, const double *params, double z, int component, double k, double muk);
int orc_fft_plane_dims(const orc_model *m, int *npx, int *npy);
int orc_fft_planes(orc_model *m, const double *params, double z_trigger, double *out);

#ifdef __cplusplus
}
#endif
#endif
