// One-time host-side preprocessing done at bf_model_create / bf_data_create.
#pragma once
#include <vector>

namespace bf {

// gsl_interp_cspline natural spline: c[n] (= y''/2)
void cspline_c(int n, const double *x, const double *y, double *c);
// per-interval cubic coefficients [n-1][4] = y_j, b_j, c_j, d_j
void cspline_interval_coefs(int n, const double *x, const double *y, double *coef4);
double cspline_eval_host(int n, const double *x, const double *y, const double *c, double xq);

// Gauss-Legendre nodes/weights on [0,1]
void gauss_legendre_01(int n, double *x, double *w);

struct HankelSetup {
    int nk = 0, nr = 0;
    double r0 = 0, dr = 0, rlo = 0, rhi = 0;
    std::vector<double> kc;      // [nk]
    std::vector<double> pk_kc;   // [2][nk]
    std::vector<double> weights; // [2][3][nr][nk]
};
// Filon weights of the fine-grid hat-function quadrature (DESIGN.md "k-space transform"):
// xi_ell^comp(r_j) = sum_i W[comp][ell][j][i] * D_ell(k_i)
void build_hankel_weights(int n_pk, const double *pk_k, const double *pk_fid, const double *pk_nw,
                          double rmin, double rmax, double dilmin, double dilmax,
                          int samples_per_decade, int fine_sub, HankelSetup &out);

// In-place Cholesky of a dense SPD matrix (row-major, lower triangle returned, upper zeroed).
// Returns false if not positive definite.
bool cholesky_lower(int n, double *a);
// Inverse of a lower-triangular matrix (row-major) into out (lower triangular).
void invert_lower(int n, const double *L, double *out);

}  // namespace bf
