/*
 * baofit_b200_host.h -- C entry points over the host-side C++ mirror of the reference
 * interfaces (baofit_b200/host/: AbsCorrelationModel, BaoCorrelationModel,
 * BaoKSpaceCorrelationModel, AbsCorrelationData, CorrelationFitter, CorrelationAnalyzer).
 *
 * A "session" is what src/baofit.cc builds from an .ini file: a model configured by its
 * model-config scripts, a data set loaded, cut and finalized, and a CorrelationFitter over
 * both. These functions let non-C++ callers (the pytest suite, bench.py) drive it; C++ callers
 * use the classes directly. All compute goes through the CUDA library (include/baofit_b200.h).
 */
#ifndef BAOFIT_B200_HOST_H
#define BAOFIT_B200_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bfh_session bfh_session;

const char *bfh_last_error(void);
/* selects the GPU of this process (one process per GPU); call before the first session */
int bfh_set_device(int device);
/* Host threads of the per-fit linear algebra and of the fit set-up (OpenMP); n <= 0 only queries. Returns the number in use.
 * One process per GPU on a shared host should take cores / processes (launchers such as torchrun export OMP_NUM_THREADS=1). */
int bfh_set_num_threads(int n);

/* Builds model + data + fitter from INI text (same keys as the reference's .ini files,
 * src/baofit.cc:46-300); relative paths are resolved against base_dir. extra_config is appended
 * as a last model-config line (may be NULL). Model construction needs no GPU; the data and the
 * fitter touch the device lazily on the first evaluation. */
int bfh_session_create(const char *ini_text, const char *base_dir, const char *extra_config, bfh_session **out);
int bfh_session_destroy(bfh_session *s);

int bfh_npar(const bfh_session *s);
int bfh_ndata(const bfh_session *s);
const char *bfh_param_name(const bfh_session *s, int i);
/* value, error, floating (0/1) of every parameter after the model-config scripts */
int bfh_param_info(const bfh_session *s, double *values, double *errors, int *floating);
/* kept global bin indices in iteration order (bit-exact mask check) */
int bfh_kept_indices(const bfh_session *s, int *indices);
/* (r, mu, z) of the kept bins in iteration order: AbsCorrelationData::getRadius / getCosAngle /
 * getRedshift (ComovingCorrelationData.cc:46-84, QuasarCorrelationData.cc:139-185) */
int bfh_kept_coordinates(const bfh_session *s, double *r, double *mu, double *z);
/* data vector [ndata] and dense covariance (or inverse covariance, as loaded) [ndata*ndata] of the kept bins */
int bfh_data_vector(const bfh_session *s, double *data, double *cov);
/* replaces the model's distortion matrix by an in-memory dense one (order x order, row-major)
 * and forwards the grid coordinates of the data (AbsCorrelationModel::setCoordinates) */
int bfh_set_distortion_matrix(bfh_session *s, const double *dense, int order);

/* CorrelationFitter::operator() for B vectors: fval[b] = (0.5*icovScale*chi2 + priors)/errScale */
int bfh_evaluate_batch(bfh_session *s, const double *params, int B, double *fval);
/* CorrelationFitter::getPrediction for B vectors: pred[i*B + b] */
int bfh_predict_batch(bfh_session *s, const double *params, int B, double *pred);
/* CorrelationFitter::fit: best-fit values / errors [npar], fval, status (0 OK, 1 WARNING, 2 ERROR),
 * covariance of the floating parameters [nfloat*nfloat] (may be NULL) */
int bfh_fit(bfh_session *s, const char *config, double *values, double *errors, double *fval, int *status,
            int *nevaluations, double *covariance);
/* The result files of src/baofit.cc:661-687 for the latest bfh_fit without a config script: <prefix>fit.chisq ("chisq nbins npar prob",
 * CorrelationAnalyzer::dumpChisquare, baofit/CorrelationAnalyzer.cc:192-206), <prefix>fit.config (the fitted parameters as a model-config
 * script that bfh_session_create's extra_config reads back), <prefix>save.pars ("index value error"), <prefix>save.pcov ("i j cov_ij" over
 * the floating parameters, j <= i; only if the fit has a covariance). */
int bfh_write_fit_outputs(bfh_session *s, const char *prefix);
/* CorrelationAnalyzer::parameterScan over grid points [first, first+count) (count<0: all):
 * chi2[count] = 2*fval, points[count*npar], status[count]; best[npar+1] = unconstrained best fit
 * values followed by its 2*fval. Returns the total grid size through *total. */
int bfh_scan_size(const bfh_session *s, long *total);
int bfh_parameter_scan(bfh_session *s, long first, long count, double *chi2, double *points, int *status,
                       double *best);
/* CorrelationAnalyzer::doToyMCSampling for toys [first, first+count): fitted[count*npar], chi2, status. As in the reference
 * (baofit/CorrelationAnalyzer.cc:361-369) the truth vector is predicted from the parameters of the fit of the combined sample when the
 * session has made one (bfh_fit without a config script), otherwise from the model's initial parameters (no-initial-fit); toymc-config
 * is applied on top. The ini option toymc-scale scales the covariance of the toy prototype. */
int bfh_toymc(bfh_session *s, long first, long count, unsigned long long seed, double *fitted, double *chi2,
              int *status);
/* toymc-save (baofit/CorrelationAnalyzer.cc:281-286): the first realisation of the next toy study that starts at toy 0 is written to
 * `path` in save-data's format; NULL or "" switches it off. */
int bfh_set_toymc_save(bfh_session *s, const char *path);
/* Resampling analyses over the individual observations of a plateroot/platelist session
 * (CorrelationAnalyzer::doJackknifeAnalysis / doBootstrapAnalysis / fitEach, CorrelationAnalyzer.cc:301-335, through
 * likely::BinnedDataResampler). kind = "jackknife" (n = observations dropped per sample), "bootstrap" (n = trials,
 * size = draws per trial with replacement, 0 = number of observations; fix_covariance: account for repeated
 * observations in the sample covariance), "each" (one sample per observation). bfh_resample_data returns the kept
 * data vector [ndata] and covariance [ndata*ndata] of sample seqno (either may be NULL); bfh_resample_fit refits
 * samples [first, first+count) (count < 0: to the end): fitted[count*npar], chi2 = 2*fval, status. */
/* SamplingOutput (CorrelationAnalyzer.cc:375-450), the text file scans, toy-MC, bootstrap and jackknife analyses save:
 * "npar nsave 1" / errors of the input fit / the input fit / one line per sample = npar values, chi2 (2*fval) and, if
 * nsave > 0, nsave x (mono quad hexa) of the sample's model between the rmin and rmax options at redshift zsave. */
int bfh_write_samples(bfh_session *s, const char *path, const double *best_values, const double *best_errors,
                      double best_chi2, long nsamples, const double *fitted /* [nsamples*npar] */, const double *chi2,
                      int nsave, double zsave);
/* The same analyses with the reference's refit pass (refit-config, CorrelationAnalyzer.cc:491-503): every sample is fitted a
 * second time with the script applied to the initial parameters. toymc_config reconfigures the truth (toymc-config). */
int bfh_toymc_refit(bfh_session *s, long first, long count, unsigned long long seed, const char *toymc_config,
                    const char *refit_config, double *fitted, double *chi2, int *status, double *refitted, double *rechi2,
                    int *restatus);
int bfh_write_samples_refit(bfh_session *s, const char *path, const double *best_values, const double *best_errors,
                            double best_chi2, long nsamples, const double *fitted, const double *chi2, int nsave, double zsave,
                            const double *best2_values, const double *best2_errors, double best2_chi2, const double *refitted,
                            const double *rechi2);
/* save-data / save-icov of src/baofit.cc:612-626: the finalized (combined, cut) data set as "<index> <value>" and
 * "<i> <j> <Cinv_ij>" text files that bfh_session_create reads back with load-icov. Either path may be NULL. */
int bfh_save_data(bfh_session *s, const char *data_path, const char *icov_path);
/* CorrelationAnalyzer::compareEach (baofit/CorrelationAnalyzer.cc:77-125; compare-each / compare-each-final, src/baofit.cc:636-645):
 * every observation of a plateroot/platelist session against the combined data before (finalized = 0) or after the final cuts.
 * Per observation: chi-square probability, chi2 over the eigenmodes of C_k - C_combined, log|C_k| / n; the per-mode contributions go
 * to the text file `path` ("<k> <log|C|/n> <chi2> <modes...>", may be NULL). Outputs are [nobservations] each and may be NULL. Host only. */
int bfh_compare_each(bfh_session *s, const char *path, int finalized, double *prob, double *chi2, double *logdet);
/* CorrelationAnalyzer::estimateCombinedCovariance (baofit/CorrelationAnalyzer.cc:733-746; bootstrap-cov-trials, src/baofit.cc:736-756):
 * sample covariance of the combined, unfinalized data vector over `trials` bootstrap draws of the observations (the draws of
 * bfh_bootstrap_counts with size 0). cov [cov_order^2] and bins [cov_order] may be NULL; cov_order = number of bins with data.
 * With icov_path the inverse is written in the loader's format when the estimate is positive definite (*posdef). Host only. */
int bfh_bootstrap_covariance(bfh_session *s, int trials, unsigned long long seed, const char *icov_path, double *cov, int cov_order,
                             int *bins, int *posdef);
/* CorrelationAnalyzer::printScaleZEff (baofit/CorrelationAnalyzer.cc:127-167) as a function of the fit results of a BAO scale parameter
 * and of gamma-scale (values, errors, their covariance): zeff, the scale evolved to zeff and its error. */
int bfh_scale_zeff(double scale, double dscale, double gamma, double dgamma, double cov_scale_gamma, double zref,
                   double *zeff, double *scale_eff, double *dscale_eff);
/* regularised upper incomplete gamma function Q(a, x) used for that probability (boost::math::gamma_p in the reference); -1 on bad arguments */
double bfh_gamma_q(double a, double x);
/* Eigenmodes of a dense symmetric n x n matrix, the decomposition behind project-modes-keep (src/baofit.cc:194,597-603;
 * likely::CovarianceMatrix::getEigenModes in the reference): values ascending, vectors[k*n + i] = component i of mode k. Host only. */
int bfh_symmetric_eigenmodes(int n, const double *matrix, double *values /* [n] */, double *vectors /* [n*n] */);
int bfh_nobservations(const bfh_session *s);
int bfh_resampling_size(bfh_session *s, const char *kind, int n, long *size);
int bfh_bootstrap_counts(bfh_session *s, int size, unsigned long long seed, long trial, int *counts /* [nobservations] */);
int bfh_resample_data(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                      long seqno, double *data, double *cov);
int bfh_resample_fit(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                     long first, long count, double *fitted, double *chi2, int *status);
int bfh_resample_refit(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                       long first, long count, const char *refit_config, double *fitted, double *chi2, int *status,
                       double *refitted, double *rechi2, int *restatus);

/* CorrelationAnalyzer::dumpResiduals (CorrelationAnalyzer.cc:607-676) at the given parameter values / errors
 * into the text file `path`; CorrelationAnalyzer::dumpModel (:678-716): multipoles[ndump*3] = mono, quad, hexa
 * at ndump radii between the rmin and rmax options, at redshift zdump (< 0: first kept bin). */
int bfh_dump_residuals(bfh_session *s, const double *values, const double *errors, const char *path, int gradients);
int bfh_dump_model(bfh_session *s, const double *values, int ndump, double zdump, double *multipoles);

/* CorrelationFitter::fit followed by CorrelationFitter::mcmc (CorrelationFitter.cc:109-123): nchain samples of
 * [npar values, fval], one every `interval` Metropolis steps of nwalkers lock-step chains. */
int bfh_mcmc(bfh_session *s, int nchain, int interval, int nwalkers, unsigned long long seed, double *samples);

/* The minimiser on its own (likely::FitModel::findMinimum "mn2::vmetric" replacement) over a
 * caller-supplied batched objective: fn(points[B*npar], B, f[B], user). Used to run the identical
 * minimiser code over the CPU oracle in the parity tests, and for host-only tests. */
typedef void (*bfh_batch_fn)(const double *points, int B, double *f, void *user);
int bfh_minimize(int npar, const double *values, const double *errors, const int *floating, bfh_batch_fn fn,
                 void *user, double *out_values, double *out_errors, double *fval, int *status, int *nevaluations);

/* The Levenberg-Marquardt driver (likely::LMFit, used for the refits of scans and toy studies) over a
 * caller-supplied batched residual function: fn(points[B*npar], B, resid[B*nres], user), objective
 * 0.5 |resid|^2 + Gaussian priors (prior_sigma[i] > 0: centre prior_centre[i]). The normal equations
 * are formed on the host here (on the GPU in the product path). For host-only tests of the driver. */
typedef void (*bfh_resid_fn)(const double *points, int B, double *resid, void *user);
int bfh_lm_minimize(int npar, const double *values, const double *errors, const int *floating, const double *prior_centre,
                    const double *prior_sigma, int nres, bfh_resid_fn fn, void *user, double *out_values, double *fval,
                    int *status, int *nevaluations, int *niterations);

#ifdef __cplusplus
}
#endif
#endif
