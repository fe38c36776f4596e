/*
 * baofit_b200.h -- C ABI of the B200-native fit-evaluation hot path of igmhub/baofit.
 *
 * This header is the drop-in boundary: plain C, plain pointers and sizes, all arithmetic fp64,
 * no exceptions cross it, no torch / CUDA types appear in any signature. Every entry point
 * names the reference interface (file:line under /root/reference) whose work it replaces.
 * The reference has no FFI of its own (its "plugin API" is C++ virtual dispatch, SURVEY.md
 * section 8b); the host-side C++ classes in baofit_b200/host/ keep the reference's class names
 * (AbsCorrelationModel, BaoCorrelationModel, BaoKSpaceCorrelationModel, BroadbandModel,
 * MetalCorrelationModel, DistortionMatrix, CorrelationFitter ...) and call down through
 * exactly these functions.
 *
 * Conventions
 *   - return value 0 = BF_OK, negative = error class (mirrors the reference's exit codes
 *     -1 options, -2 model initialisation, -3 data, -4 analysis; src/baofit.cc:313-333,468-471,
 *     629-632,816-819); bf_last_error() returns the message of the last failure on this thread.
 *   - parameter vectors are row-major params[b*npar + j], j in the reference's defineParameter
 *     order (baofit/AbsCorrelationModel.cc:95-114, baofit/BaoCorrelationModel.cc:31-48,
 *     baofit/BaoKSpaceCorrelationModel.cc:53-118).
 *   - predictions are bin-major: pred[i*ldp + b] for kept data bin i (iteration order of the
 *     reference's BinnedData index iterator, baofit/CorrelationFitter.cc:57) and vector b.
 *   - status[b] != 0 marks a parameter vector the reference would have thrown on
 *     (baofit/BaoKSpaceCorrelationModel.cc:420-427); its outputs are NaN, the batch continues.
 *   - there is no CPU fallback: every compute entry point fails with BF_ERR_ANALYSIS when no
 *     CUDA device is usable.
 */
#ifndef BAOFIT_B200_H
#define BAOFIT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_ERR_OPTIONS (-1)
#define BF_ERR_MODEL (-2)
#define BF_ERR_DATA (-3)
#define BF_ERR_ANALYSIS (-4)

/* per-vector status bits */
#define BF_STATUS_OK 0
#define BF_STATUS_DILATION_MIN 1 /* "hit min dilation limit", BaoKSpaceCorrelationModel.cc:420 */
#define BF_STATUS_DILATION_MAX 2 /* "hit max dilation limit", BaoKSpaceCorrelationModel.cc:424 */
#define BF_STATUS_FFT_RANGE 4    /* FFT model: a dilated bin falls outside the tabulated (r_perp, r_par) plane (fft_rmax) */
#define BF_STATUS_Z_RETRIGGER 8  /* k-space model, gamma-beta != 0 on data whose consecutive bins jump by more than dzmin in z:
                                  * the reference re-transforms mid-sweep (BaoKSpaceCorrelationModel.cc:317-319), not reproduced */

/* model kinds */
#define BF_MODEL_RSPACE 0     /* baofit/BaoCorrelationModel.cc          */
#define BF_MODEL_KSPACE 1     /* baofit/BaoKSpaceCorrelationModel.cc    */
#define BF_MODEL_KSPACE_FFT 2 /* baofit/BaoKSpaceFftCorrelationModel.cc       */

/* option flags (names follow the reference's command-line options, src/baofit.cc:46-178) */
#define BF_FLAG_CROSS_CORRELATION (1u << 0)
#define BF_FLAG_ANISOTROPIC (1u << 1)
#define BF_FLAG_DECOUPLED (1u << 2)
#define BF_FLAG_COMBINED_BIAS (1u << 3)
#define BF_FLAG_COMBINED_SCALE (1u << 4)
#define BF_FLAG_NL_BROADBAND (1u << 5)
#define BF_FLAG_NL_CORRECTION (1u << 6)
#define BF_FLAG_FIT_NL_CORRECTION (1u << 7)
#define BF_FLAG_NL_CORRECTION_ALT (1u << 8)
#define BF_FLAG_BIN_SMOOTH (1u << 9)
#define BF_FLAG_BIN_SMOOTH_ALT (1u << 10)
#define BF_FLAG_HCD_MODEL (1u << 11)
#define BF_FLAG_HCD_MODEL_ALT (1u << 12)
#define BF_FLAG_UVFLUCTUATION (1u << 13)
#define BF_FLAG_RADIATION_MODEL (1u << 14)
#define BF_FLAG_SMOOTH_GAUSS (1u << 15)
#define BF_FLAG_SMOOTH_LORENTZ (1u << 16)
#define BF_FLAG_DISTORTION_ALT (1u << 17)
#define BF_FLAG_NO_DISTORTION (1u << 18)

/* metal model kinds (baofit/MetalCorrelationModel.cc:27-161) */
#define BF_METAL_NONE 0
#define BF_METAL_AUTO_TABLE 1   /* metal-model: templates tabulated per grid bin       (:238-266) */
#define BF_METAL_CROSS_BICUBIC 2 /* metal-model-interpolate: bicubic in (rperp,rpar)   (:268-301) */
#define BF_METAL_TOY 3          /* toy-metal: four Gaussian bumps                      (:303-346) */

/* broadband slots */
#define BF_BB_DIST_ADD 0      /* "dist add"            post-matrix, additive        */
#define BF_BB_DIST_MUL 1      /* "dist mul"            post-matrix, multiplicative  */
#define BF_BB_DM_DIST_ADD 2   /* "dist Matrix dist add" pre-matrix, additive        */
#define BF_BB_DM_DIST_MUL 3   /* "dist Matrix dist mul" pre-matrix, multiplicative  */

/* One broadband polynomial (baofit/BroadbandModel.cc:93-167 constructor, :197-226 evaluation).
 * Axis ranges are min:max:step as produced by the reference's grammar after finalizeAxis;
 * r exponents are divided by r_denom (negative r step in the spec). Coefficients live in the
 * parent model's parameter vector starting at idx_base, ordered z, mu, r, rP, rT (rT fastest). */
typedef struct bf_broadband_spec {
    int enabled;
    int idx_base;
    int r_min, r_max, r_step, r_denom;
    int mu_min, mu_max, mu_step;
    int rp_min, rp_max, rp_step;
    int rt_min, rt_max, rt_step;
    int z_min, z_max, z_step;
    double r0, z0;
} bf_broadband_spec;

typedef struct bf_model_desc {
    int kind;           /* BF_MODEL_*                                                     */
    unsigned flags;     /* BF_FLAG_*                                                      */
    int npar;           /* total number of parameters (floating + fixed)                  */
    double zref;        /* reference redshift, AbsCorrelationModel.cc:81-84               */
    double omega_matter;/* AbsCorrelationModel.cc:86-89, used by the velocity shift       */

    /* linear-bias block (AbsCorrelationModel.cc:95-128). -1 = parameter not defined.     */
    int idx_beta, idx_bb, idx_gamma_bias, idx_gamma_beta;
    int idx_delta_v, idx_bias2, idx_bb2; /* cross-correlation only                         */
    int idx_beta_bias;                   /* combined-bias only                              */
    /* BAO block: amplitude, alpha-iso, alpha-parallel, alpha-perp, gamma-scale consecutive */
    int idx_bao;
    int idx_comb_scale; /* "BAO aperp/apar", combined-scale only                           */
    /* radiation block: strength, anisotropy, mean free path, lifetime consecutive         */
    int idx_rad;
    /* k-space blocks (BaoKSpaceCorrelationModel.cc:65-105; Fft :51-56)                    */
    int idx_nl;     /* SigmaNL-perp, 1+f                                                   */
    int idx_cont;   /* cont-kc, cont-pc (FFT model)                                        */
    int idx_bs;     /* bin scale-par, bin scale-perp                                       */
    int idx_hcd;    /* HCD beta, rel bias, scale                                           */
    int idx_uv;     /* UV rel bias, abs resp bias, mean free path                          */
    int idx_smgaus; /* smooth scale gauss                                                  */
    int idx_smlor;  /* smooth scale lorentz                                                */
    int idx_nlcorr; /* nlcorr qnl, nlcorr kp                                               */

    /* r-space tabulated multipoles xi_ell(r), ell = 0,2,4 (BaoCorrelationModel.cc:50-71).
     * Interpolation is a natural cubic spline that returns the end-point value outside the
     * table (likely::Interpolator "cspline").                        */
    int n_fid;
    const double *fid_r, *fid_xi0, *fid_xi2, *fid_xi4;
    int n_nw;
    const double *nw_r, *nw_xi0, *nw_xi2, *nw_xi4;

    /* k-space tabulated powers P(k) (BaoKSpaceCorrelationModel.cc:120-147), same k nodes.   */
    int n_pk;
    const double *pk_k, *pk_fid, *pk_nw;
    double rmin, rmax;     /* final-cut radial range before dilation (:156-164)             */
    double dilmin, dilmax; /* allowed radial dilation (:158-162, checked at :420-427)       */
    int samples_per_decade;/* k nodes per decade for the multipoles of D(k,mu) (:140-141)   */
    int ell_max;           /* 0, 2 or 4                                                     */
    double zeff;           /* <=0: take z of the triggering data bin (:315)                 */
    double sigma8;         /* NonLinearCorrectionModel.cc:62-64                             */
    double dzmin;          /* BaoKSpaceCorrelationModel.cc:318; used to flag what is not reproduced */

    bf_broadband_spec bb[4]; /* BF_BB_* slots                                               */

    /* metals (MetalCorrelationModel.cc). idx_metal = first metal parameter.               */
    int metal_kind;
    int idx_metal;
    int metal_ncomb;             /* number of line combinations                             */
    int metal_nmet;              /* number of (beta,bias) pairs incl. the base tracer       */
    const int *metal_p1, *metal_p2; /* [ncomb] indices into the (beta,bias) pair list       */
    int metal_ngrid;             /* length of each zgrid / auto template column             */
    const double *metal_zgrid;   /* [ncomb][ngrid] redshift of each grid bin (.grid files)  */
    const double *metal_auto;    /* [ncomb*3][ngrid] ell=0,2,4 templates per grid bin       */
    /* cross templates on a regular (rperp=x, rpar=y) grid, value at [ (c*3+l)*ny*nx + iy*nx + ix ] */
    int metal_nx, metal_ny;
    double metal_x0, metal_y0, metal_dx, metal_dy;
    const double *metal_cross;

    /* distortion matrix (DistortionMatrix.cc:22-73): dense row-major order x order, or NULL */
    int dist_matrix_order;
    const double *dist_matrix;

    /* 3-D FFT k-space model (BaoKSpaceFftCorrelationModel.cc:21-27,92-97; options gridspacing,
     * ngridx/y/z, zcorr0/1/2 of src/baofit.cc:76-84,102-106). The y axis is the line of sight.
     * xi is tabulated on the (r_perp, r_par) plane through the origin of the periodic box and
     * interpolated bicubically (cosmo::DistortedPowerCorrelationFft::getCorrelation call sites
     * :241-243); the plane is kept out to fft_rmax in both directions, a dilated bin beyond it
     * flags BF_STATUS_FFT_RANGE. fft_rmax <= 0 selects rmax * dilmax.                        */
    double fft_spacing;
    int fft_nx, fft_ny, fft_nz;
    double fft_rmax;
    double zcorr0, zcorr1, zcorr2; /* effective redshift of a bin, :165-168 (zcorr0 <= 0: off) */
} bf_model_desc;

typedef struct bf_data_desc {
    /* kept data bins after final cuts, in BinnedData iteration order
     * (AbsCorrelationData.cc:67-94; CorrelationFitter.cc:57-71)                              */
    int n_data;
    const double *r, *mu, *z; /* getRadius / getCosAngle / getRedshift of each kept bin     */
    const int *index;         /* global grid index of each kept bin                          */
    const double *data;       /* data vector d[n_data]                                       */
    /* dense symmetric n_data x n_data matrix: covariance (is_inverse=0) or its inverse (1)  */
    const double *cov;
    int cov_is_inverse;
    /* full grid coordinates for the distortion matrix (AbsCorrelationModel.cc:51-63);
     * n_grid may be 0 when the model has no distortion matrix                               */
    int n_grid;
    const double *grid_r, *grid_mu, *grid_z;
    /* transverse binning type (AbsCorrelationData.h:19, CorrelationFitter.cc:62-69):
     * 0 = Coordinate: bins are points (r, mu, z);
     * 1 = Multipole (data-format comoving-multipole): mu[i] holds the multipole ell = 0, 2, 4 of bin i
     *     and the prediction is AbsCorrelationModel::evaluate(r, multipole, z, ...), i.e. the projection
     *     (2 ell + 1)/2 Int_-1^1 L_ell(mu) xi(r, mu, z) dmu (AbsCorrelationModel.cc:65-79), evaluated
     *     with an nmu_project-point Gauss-Legendre rule (0 selects 16).                            */
    int binning_type;
    int nmu_project;
} bf_data_desc;

#define BF_BINNING_COORDINATE 0
#define BF_BINNING_MULTIPOLE 1

typedef struct bf_ctx bf_ctx;     /* one CUDA device + stream + workspaces; one per host thread */
typedef struct bf_model bf_model; /* immutable after creation: tables, weights, layout          */
typedef struct bf_data bf_data;   /* immutable after creation: coordinates, d, Cholesky factors */

const char *bf_last_error(void);
const char *bf_version(void);

int bf_ctx_create(int device, bf_ctx **out);
int bf_ctx_destroy(bf_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long bf_ctx_launch_count(const bf_ctx *ctx);
/* Blocks until all work queued on the context's stream has finished. */
int bf_ctx_synchronize(bf_ctx *ctx);
/* Raw cudaStream_t of the context (as void*), so callers can record events on it. */
void *bf_ctx_stream(bf_ctx *ctx);
/* Per-kernel device timing for the roofline report: when enabled, a CUDA event pair brackets
 * every kernel launch on the context's stream. bf_ctx_get_profile synchronises, then returns the
 * number of distinct kernels and fills (up to max_entries) name / total milliseconds / launch
 * count; names stay valid until the context is destroyed. Enabling resets the counters. */
int bf_ctx_set_profiling(bf_ctx *ctx, int enable);
int bf_ctx_get_profile(bf_ctx *ctx, int max_entries, const char **names, double *ms_total,
                       long long *launches);

/* Replaces the constructors of BaoCorrelationModel (BaoCorrelationModel.cc:18-86),
 * BaoKSpaceCorrelationModel (BaoKSpaceCorrelationModel.cc:26-211) and
 * BaoKSpaceFftCorrelationModel (BaoKSpaceFftCorrelationModel.cc:21-113): tables are copied, spline
 * coefficients / Hankel weights are computed once and uploaded. */
int bf_model_create(bf_ctx *ctx, const bf_model_desc *desc, bf_model **out);
int bf_model_destroy(bf_model *model);
int bf_model_npar(const bf_model *model);

/* Replaces AbsCorrelationData finalize (ComovingCorrelationData.cc:28-33) + the lazily
 * inverted likely::CovarianceMatrix behind BinnedData::chiSquare (CorrelationFitter.cc:85):
 * factors C = L L^T once on the host and uploads W = L^-1 (and L for toy noise). */
int bf_data_create(bf_ctx *ctx, const bf_data_desc *desc, bf_data **out);
int bf_data_destroy(bf_data *data);
int bf_data_ndata(const bf_data *data);

/* Replaces CorrelationFitter::getPrediction (CorrelationFitter.cc:53-72) for B parameter
 * vectors at once: pred[i*ldp + b] = model->evaluate(r_i, mu_i, z_i, params_b, index_i)
 * (AbsCorrelationModel.cc:21-41 and the _evaluate bodies). Host buffers. */
int bf_eval_batch(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                  const double *params, int B, double *pred, int ldp, int *status);

/* Replaces the chi-square half of CorrelationFitter::operator() (CorrelationFitter.cc:74-86):
 * chi2[b] = (pred_b - d)^T C^-1 (pred_b - d) = BinnedData::chiSquare(pred_b). Priors, the
 * 0.5*icovScale factor and errorScale are applied by the host-side CorrelationFitter.
 * data_override: NULL, or [B][n_data] alternative data vectors (toy-MC / bootstrap samples,
 * CorrelationAnalyzer.cc:262-299). Host buffers. */
int bf_chi2_batch(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                  const double *params, int B, const double *data_override,
                  double *chi2, int *status);

/* Same two operations on buffers that already live in device memory of the context's GPU
 * (device addresses passed as plain pointers); asynchronous on the context's stream for the r-space and
 * k-space models (which k-space chain runs is decided on the device). The FFT model still reads the key flag of a
 * batch back on the host before it builds its per-key tables, i.e. its _dev calls wait for the GPU once. */
int bf_eval_batch_dev(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                      const double *d_params, int B, double *d_pred, int ldp, int *d_status);
int bf_chi2_batch_dev(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                      const double *d_params, int B, const double *d_data_override,
                      double *d_chi2, int *d_status);

/* bf_chi2_batch in two halves, for callers that evaluate one batch after another (the rounds of a scan or of a
 * toy study, CorrelationAnalyzer.cc:169-190,337-373): submit returns once the work is queued, wait blocks until
 * chi2[] / status[] of that ticket are complete. At most two tickets may be outstanding; the parameter upload of the
 * newer one overlaps the evaluation of the older one. params, chi2 and status must stay valid until the wait
 * returns and should be page-locked (bf_pinned_alloc), or the copies are not asynchronous. */
int bf_chi2_batch_submit(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                         const double *params, int B, double *chi2, int *status, int *ticket);
int bf_chi2_batch_wait(bf_ctx *ctx, int ticket);

/* Undistorted correlation on the full grid (the vector DistortionMatrix::setCorrelation is
 * fed with, BaoKSpaceCorrelationModel.cc:460-521): xiu[g*ldp + b]. Host buffers. Only valid
 * for models with a distortion matrix. Exposed for parity tests of the model kernel alone. */
int bf_eval_undistorted_batch(bf_ctx *ctx, const bf_model *model, const bf_data *data,
                              const double *params, int B, double *xiu, int ldp, int *status);

/* k-space models: multipole tables xi_ell(r_j) of the peak and no-wiggles components after the
 * P(k) -> xi_ell(r) transform of parameter vector b (cosmo::DistortedPowerCorrelation::transform
 * call sites BaoKSpaceCorrelationModel.cc:353-379). out[((comp*3 + l)*n_r + j)*ldb + b],
 * comp 0 = peak, 1 = no-wiggles; r_j = bf_model_transform_r0 + j*bf_model_transform_dr. */
int bf_transform_batch(bf_ctx *ctx, const bf_model *model, const double *params, int B,
                       double z_trigger, double *out, int ldb);
/* FFT k-space model: the two planes xi_comp(ix*spacing, iy*spacing) that the 3-D transform of
 * parameter vector b leaves on the z = 0 plane of the box (cosmo::DistortedPowerCorrelationFft::
 * transform call sites BaoKSpaceFftCorrelationModel.cc:204,212). out[((b*2 + comp)*npy + iy)*npx + ix],
 * comp 0 = peak, 1 = no-wiggles; npx/npy from bf_model_fft_plane_dims. z_trigger is the redshift of
 * the bin that triggers the transform (after the zcorr correction, if any). Host buffers. */
int bf_fft_planes_batch(bf_ctx *ctx, const bf_model *model, const double *params, int B,
                        double z_trigger, double *out);
int bf_model_fft_plane_dims(const bf_model *model, int *npx, int *npy);
int bf_model_transform_nr(const bf_model *model);
double bf_model_transform_r0(const bf_model *model);
double bf_model_transform_dr(const bf_model *model);

/* Host-side set-up of the k-space transform, no GPU involved: the Bessel weights that bf_model_create uploads,
 * xi_ell^comp(r_j) = sum_i weights[((comp*3 + ell/2)*nr + j)*nk + i] * D_ell(k_i) with D_ell the multipoles of
 * D(k, mu) on the coarse k grid (stand-in for the set-up inside cosmo::DistortedPowerCorrelation::initialize,
 * call sites BaoKSpaceCorrelationModel.cc:347-353,371). weights == NULL only reports the sizes. Used by the
 * reference-shaped CPU timing baseline of bench.py (oracle/cpu_baseline.c) and by parity tests. */
int bf_kspace_transform_weights(const bf_model_desc *desc, int *nr, int *nk, double *weights);

/* Toy-MC noise, replaces likely::CovarianceMatrix::sample (call site
 * CorrelationAnalyzer.cc:271): out[t][i] = d_i + scale * (L g_t)_i with g_t ~ N(0,1) drawn
 * from a counter-based generator keyed by (seed, first_toy + t), so that realisations do not
 * depend on how toys are sharded over GPUs. Host output buffer [ntoy][n_data]. */
int bf_sample_toys(bf_ctx *ctx, const bf_data *data, const double *truth, unsigned long long seed,
                   long long first_toy, int ntoy, double scale, double *out);

/* Device-resident toy-MC realisations: the table [ntoy][n_data] of bf_sample_toys, drawn once on the
 * GPU and kept there (the loop of CorrelationAnalyzer.cc:262-299,337-373 refits each realisation many
 * times; only its index travels with every evaluation point). */
typedef struct bf_toys bf_toys;
int bf_toys_create(bf_ctx *ctx, const bf_data *data, const double *truth, unsigned long long seed,
                   long long first_toy, int ntoy, double scale, bf_toys **out);
int bf_toys_destroy(bf_toys *toys);
int bf_toys_download(bf_ctx *ctx, const bf_toys *toys, double *out /* [ntoy][n_data] */);
/* chi2[b] = BinnedData::chiSquare of prediction b against realisation toy_index[b] (host ints, each in
 * [0, ntoy)) of the table. Host buffers for params / chi2 / status. */
int bf_chi2_batch_toys(bf_ctx *ctx, const bf_model *model, const bf_data *data, const double *params, int B,
                       const bf_toys *toys, const int *toy_index, double *chi2, int *status);

/* Gauss-Newton normal equations for many independent fits at once (the minimiser-side consumer of the
 * hot path: CorrelationFitter::fit -> likely::FitModel::findMinimum, CorrelationFitter.cc:88-97, called
 * per scan point / toy by CorrelationAnalyzer.cc:169-190). params holds nfit groups of G consecutive
 * vectors: the centre x of a fit followed by its G-1 probe points. With y(p) = L^-1 (pred(p) - d) the
 * whitened residual (chi2 = |y|^2) and d_i = y(probe i) - y(centre), the call returns per fit the
 * symmetric G x G matrix
 *     gram[0][0] = y0.y0 (= chi2 at the centre),  gram[i][0] = gram[0][i] = d_i.y0,  gram[i][j] = d_i.d_j,
 * from which J^T J and J^T y follow by dividing by the probe steps. toys / toy_index (may be NULL)
 * select, per fit, the device-resident realisation the residuals are taken against.
 * status[nfit*G] as in bf_chi2_batch. Host buffers. */
int bf_gram_batch(bf_ctx *ctx, const bf_model *model, const bf_data *data, const double *params, int nfit, int G,
                  const bf_toys *toys, const int *toy_index, double *gram, int *status);

/* Same normal equations with the probe points generated on the GPU: fit f is described by its centre
 * centres[f*npar ..] and its probe steps steps[f*npar ..] (0 = parameter not probed; every fit probes
 * the same number nprobe of parameters). The group of fit f is
 *     x,  x + h_1 e_1,  x - h_1 e_1,  x + h_2 e_2,  x - h_2 e_2, ...      (G = 2 nprobe + 1 columns)
 * with the probed parameters in index order, i.e. only (2 npar) instead of (G npar) numbers per fit
 * cross the bus. gram[nfit][G][G] as in bf_gram_batch; status[nfit] is the OR over the group. */
int bf_gram_probes(bf_ctx *ctx, const bf_model *model, const bf_data *data, const double *centres, const double *steps,
                   int nfit, int nprobe, const bf_toys *toys, const int *toy_index, double *gram, int *status);
/* The same probes reduced on the device to what a Gauss-Newton / Levenberg-Marquardt step reads (a quarter of the Gram
 * matrix to compute and to copy back). With y0 = y(x), a_i = y(x + h_i) - y(x - h_i), s_i = y(x + h_i) + y(x - h_i) - 2 y0:
 * normal[f] = the (nprobe+1) x (nprobe+1) matrix [y0 a_1 .. a_n]^T [y0 a_1 .. a_n] (row-major; [0][0] = chi2 at the centre,
 * [i][0] = 2 h_i times the chi2 gradient / 2, [i][j] = 4 h_i h_j J_i . J_j) followed by the nprobe values y0 . s_i
 * (h_i^2 times the diagonal second-order term). Replaces the same Minuit2 derivative evaluations as bf_gram_batch. */
int bf_normal_probes(bf_ctx *ctx, const bf_model *model, const bf_data *data, const double *centres, const double *steps,
                     int nfit, int nprobe, const bf_toys *toys, const int *toy_index, double *normal, int *status);
/* Page-locked host memory for the buffers of the host-pointer entry points (results come back at full PCIe speed). */
int bf_pinned_alloc(size_t bytes, void **ptr);
int bf_pinned_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* BAOFIT_B200_H */
