// C entry points over the host-side classes (include/baofit_b200_host.h).
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/baofit_b200_host.h"
#include <omp.h>

#include "baofit_host.h"

using namespace baofit;

struct bfh_session {
    Options options;
    AbsCorrelationModelPtr model;
    AbsCorrelationDataPtr data;
    BinnedDataResamplerPtr resampler;  // platelist sessions only
    std::shared_ptr<CorrelationFitter> fitter;
    std::shared_ptr<CorrelationAnalyzer> analyzer;
    likely::FunctionMinimumPtr lastFit;  // of the latest bfh_fit without a config script (the "initial fit" of src/baofit.cc:646-660)
    std::string toymcSaveName;           // bfh_set_toymc_save
};

static thread_local std::string g_herr;
extern "C" const char *bfh_last_error(void) { return g_herr.c_str(); }

#define BFH_TRY try {
#define BFH_CATCH                          \
    }                                      \
    catch (std::exception const &e) {      \
        g_herr = e.what();                 \
        return -4;                         \
    }                                      \
    return 0;

extern "C" int bfh_set_device(int device) {
    BFH_TRY
    setDevice(device);
    BFH_CATCH
}

extern "C" int bfh_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}

extern "C" int bfh_session_create(const char *ini_text, const char *base_dir, const char *extra_config, bfh_session **out) {
    if (!ini_text || !out) { g_herr = "null argument"; return -1; }
    bfh_session *s = new bfh_session();
    try {
        s->options = parseIni(ini_text);
        if (extra_config && *extra_config) s->options.modelConfig.push_back(extra_config);
    } catch (std::exception const &e) { g_herr = e.what(); delete s; return -1; }
    try {
        s->model = createModel(s->options, base_dir ? base_dir : "");
    } catch (std::exception const &e) { g_herr = e.what(); delete s; return -2; }
    try {
        s->data = createData(s->options, base_dir ? base_dir : "", &s->resampler);
        if (s->options.flag("dist-matrix")) {
            std::vector<double> r, mu, z;
            s->data->gridCoordinates(r, mu, z);
            s->model->setCoordinates(r, mu, z);
        }
    } catch (std::exception const &e) { g_herr = e.what(); delete s; return -3; }
    try {
        int covSampleSize = (int)s->options.num("cov-sample-size", 0);
        s->fitter.reset(new CorrelationFitter(s->data, s->model, covSampleSize));
        s->analyzer.reset(new CorrelationAnalyzer(s->options.str("min-method", "mn2::vmetric"), s->options.num("rmin", 0),
                                                  s->options.num("rmax", 200), covSampleSize, false));
        s->analyzer->addData(s->data);
        s->analyzer->setModel(s->model);
        s->analyzer->setResampler(s->resampler);
        s->analyzer->setToyMCScale(s->options.num("toymc-scale", 1));
    } catch (std::exception const &e) { g_herr = e.what(); delete s; return -4; }
    *out = s;
    return 0;
}

extern "C" int bfh_session_destroy(bfh_session *s) {
    delete s;
    return 0;
}

extern "C" int bfh_npar(const bfh_session *s) { return s ? s->model->getNParameters() : 0; }
extern "C" int bfh_ndata(const bfh_session *s) { return s ? (int)s->data->keptIndices().size() : 0; }
extern "C" const char *bfh_param_name(const bfh_session *s, int i) {
    if (!s || i < 0 || i >= s->model->getNParameters()) return "";
    return s->model->getFitParameters()[i].name.c_str();
}
extern "C" int bfh_param_info(const bfh_session *s, double *values, double *errors, int *floating) {
    BFH_TRY
    auto const &p = s->model->getFitParameters();
    for (size_t i = 0; i < p.size(); ++i) {
        if (values) values[i] = p[i].value;
        if (errors) errors[i] = p[i].error;
        if (floating) floating[i] = p[i].floating ? 1 : 0;
    }
    BFH_CATCH
}
extern "C" int bfh_kept_indices(const bfh_session *s, int *indices) {
    BFH_TRY
    auto const &k = s->data->keptIndices();
    std::copy(k.begin(), k.end(), indices);
    BFH_CATCH
}

extern "C" int bfh_kept_coordinates(const bfh_session *s, double *r, double *mu, double *z) {
    BFH_TRY
    auto const &k = s->data->keptIndices();
    for (size_t i = 0; i < k.size(); ++i) {
        if (r) r[i] = s->data->getRadius(k[i]);
        if (mu) mu[i] = s->data->getTransverseBinningType() == AbsCorrelationData::Coordinate ? s->data->getCosAngle(k[i]) : (double)s->data->getMultipole(k[i]);
        if (z) z[i] = s->data->getRedshift(k[i]);
    }
    BFH_CATCH
}

extern "C" int bfh_data_vector(const bfh_session *s, double *data, double *cov) {
    BFH_TRY
    if (data) { std::vector<double> d = s->data->dataVector(); std::copy(d.begin(), d.end(), data); }
    if (cov) { std::vector<double> c = s->data->covarianceMatrix(); std::copy(c.begin(), c.end(), cov); }
    BFH_CATCH
}

extern "C" int bfh_set_distortion_matrix(bfh_session *s, const double *dense, int order) {
    BFH_TRY
    auto *km = dynamic_cast<BaoKSpaceCorrelationModel *>(s->model.get());
    if (!km) throw RuntimeError("the distortion matrix needs the k-space model");
    if (order != s->data->getGrid().getNBinsTotal()) throw RuntimeError("Distortion matrix order does not match grid size.");
    std::vector<double> d(dense, dense + (size_t)order * order);
    km->setDistortionMatrix(std::make_shared<DistortionMatrix>(d, order));
    std::vector<double> r, mu, z;
    s->data->gridCoordinates(r, mu, z);
    s->model->setCoordinates(r, mu, z);
    BFH_CATCH
}

static std::vector<likely::Parameters> unflatten(const double *params, int B, int npar) {
    std::vector<likely::Parameters> v(B);
    for (int b = 0; b < B; ++b) v[b].assign(params + (size_t)b * npar, params + (size_t)(b + 1) * npar);
    return v;
}

extern "C" int bfh_evaluate_batch(bfh_session *s, const double *params, int B, double *fval) {
    BFH_TRY
    std::vector<double> f;
    s->fitter->evaluateBatch(unflatten(params, B, s->model->getNParameters()), f);
    std::copy(f.begin(), f.end(), fval);
    BFH_CATCH
}

extern "C" int bfh_predict_batch(bfh_session *s, const double *params, int B, double *pred) {
    BFH_TRY
    std::vector<double> p;
    s->fitter->predictBatch(unflatten(params, B, s->model->getNParameters()), p);
    std::copy(p.begin(), p.end(), pred);
    BFH_CATCH
}

extern "C" int bfh_fit(bfh_session *s, const char *config, double *values, double *errors, double *fval, int *status,
                       int *nevaluations, double *covariance) {
    BFH_TRY
    likely::FunctionMinimumPtr fm = s->fitter->fit("mn2::vmetric", config ? config : "");
    if (!config || !*config) s->lastFit = fm;
    for (size_t i = 0; i < fm->parameters.size(); ++i) {
        if (values) values[i] = fm->parameters[i].value;
        if (errors) errors[i] = fm->parameters[i].floating ? fm->parameters[i].error : 0.0;
    }
    if (fval) *fval = fm->minValue;
    if (status) *status = (int)fm->status;
    if (nevaluations) *nevaluations = fm->nEvaluations;
    if (covariance) std::copy(fm->covariance.begin(), fm->covariance.end(), covariance);
    BFH_CATCH
}

// The result files src/baofit.cc:661-687 writes after the fit of the combined sample, with its file names.
extern "C" int bfh_write_fit_outputs(bfh_session *s, const char *prefix) {
    BFH_TRY
    if (!s || !s->lastFit) throw RuntimeError("bfh_write_fit_outputs: call bfh_fit (without a config script) first");
    likely::FunctionMinimum const &fm = *s->lastFit;
    const std::string pre = prefix ? prefix : "";
    auto open = [&](std::string const &name, std::ofstream &out) {
        out.open((pre + name).c_str());
        if (!out.good()) throw RuntimeError("bfh_write_fit_outputs: unable to open " + pre + name);
    };
    char buf[256];
    int nfloat = 0;
    for (auto const &p : fm.parameters) nfloat += p.floating ? 1 : 0;
    {   // CorrelationAnalyzer::dumpChisquare, CorrelationAnalyzer.cc:192-206: "chisq nbins npar prob" (prob = -1 if undefined)
        std::ofstream out;
        open("fit.chisq", out);
        const double chisq = 2 * fm.minValue;
        const int nbins = (int)s->data->keptIndices().size();
        if (chisq > 0 && nbins > nfloat) {
            std::snprintf(buf, sizeof(buf), "%.17g %d %d %.17g", chisq, nbins, nfloat, gammaQ(0.5 * (nbins - nfloat), 0.5 * chisq));
        } else
            std::snprintf(buf, sizeof(buf), "%.17g %d %d -1", chisq, nbins, nfloat);
        out << buf << std::endl;
    }
    {   // likely::fitParametersToScript: the fitted parameters as a model-config script (dependency absent: our own statements,
        // which configureFitParameters reads back)
        std::ofstream out;
        open("fit.config", out);
        for (auto const &p : fm.parameters) {
            if (p.floating) std::snprintf(buf, sizeof(buf), "value[%s]=%.17g; error[%s]=%.17g;", p.name.c_str(), p.value, p.name.c_str(), p.error);
            else std::snprintf(buf, sizeof(buf), "fix[%s]=%.17g;", p.name.c_str(), p.value);
            out << buf << std::endl;
        }
    }
    {   // FunctionMinimum::saveParameters: "<index> <value> <error>", error 0 for fixed parameters (layout ours, dependency absent)
        std::ofstream out;
        open("save.pars", out);
        for (size_t i = 0; i < fm.parameters.size(); ++i) {
            std::snprintf(buf, sizeof(buf), "%d %.17g %.17g", (int)i, fm.parameters[i].value, fm.parameters[i].floating ? fm.parameters[i].error : 0.0);
            out << buf << std::endl;
        }
    }
    if ((int)fm.covariance.size() == nfloat * nfloat && nfloat > 0) {
        // FunctionMinimum::saveFloatingParameterCovariance: "<i> <j> <cov_ij>" over the floating parameters, j <= i, indexed in the
        // full parameter list (layout ours, dependency absent)
        std::ofstream out;
        open("save.pcov", out);
        std::vector<int> index;
        for (size_t i = 0; i < fm.parameters.size(); ++i)
            if (fm.parameters[i].floating) index.push_back((int)i);
        for (int a = 0; a < nfloat; ++a)
            for (int b = 0; b <= a; ++b) {
                std::snprintf(buf, sizeof(buf), "%d %d %.17g", index[a], index[b], fm.covariance[(size_t)a * nfloat + b]);
                out << buf << std::endl;
            }
    }
    BFH_CATCH
}

extern "C" int bfh_dump_residuals(bfh_session *s, const double *values, const double *errors, const char *path, int gradients) {
    BFH_TRY
    likely::FunctionMinimumPtr fm(new likely::FunctionMinimum());
    fm->parameters = s->model->getFitParameters();
    for (size_t i = 0; i < fm->parameters.size(); ++i) {
        fm->parameters[i].value = values[i];
        if (errors) fm->parameters[i].error = errors[i];
    }
    std::ofstream out(path);
    if (!out.good()) throw RuntimeError(std::string("unable to open ") + path);
    s->analyzer->dumpResiduals(out, fm, "", gradients != 0);
    BFH_CATCH
}

extern "C" int bfh_dump_model(bfh_session *s, const double *values, int ndump, double zdump, double *multipoles) {
    BFH_TRY
    likely::FitParameters p = s->model->getFitParameters();
    for (size_t i = 0; i < p.size(); ++i) p[i].value = values[i];
    std::ostringstream os;
    s->analyzer->dumpModel(os, p, ndump, zdump, "", true);
    std::istringstream is(os.str());
    for (int k = 0; k < 3 * ndump; ++k)
        if (!(is >> multipoles[k])) throw RuntimeError("dumpModel produced too few values");
    BFH_CATCH
}

extern "C" int bfh_write_samples(bfh_session *s, const char *path, const double *best_values, const double *best_errors,
                                 double best_chi2, long nsamples, const double *fitted, const double *chi2, int nsave, double zsave) {
    return bfh_write_samples_refit(s, path, best_values, best_errors, best_chi2, nsamples, fitted, chi2, nsave, zsave, nullptr, nullptr, 0,
                                   nullptr, nullptr);
}

extern "C" int bfh_write_samples_refit(bfh_session *s, const char *path, const double *best_values, const double *best_errors,
                                       double best_chi2, long nsamples, const double *fitted, const double *chi2, int nsave, double zsave,
                                       const double *best2_values, const double *best2_errors, double best2_chi2, const double *refitted,
                                       const double *rechi2) {
    BFH_TRY
    if (!path || !best_values || (nsamples > 0 && (!fitted || !chi2))) throw RuntimeError("bfh_write_samples: null argument");
    const int npar = s->model->getNParameters();
    auto minimum = [&](const double *values, const double *errors, double c2) {
        likely::FunctionMinimumPtr fm(new likely::FunctionMinimum());
        fm->parameters = s->model->getFitParameters();
        for (int i = 0; i < npar; ++i) {
            fm->parameters[i].value = values[i];
            if (errors) { fm->parameters[i].error = errors[i]; fm->parameters[i].floating = errors[i] > 0; }
        }
        fm->minValue = 0.5 * c2;
        return fm;
    };
    auto rows = [&](const double *flat) {
        std::vector<std::vector<double>> f((size_t)nsamples);
        for (long k = 0; k < nsamples; ++k) f[k].assign(flat + (size_t)k * npar, flat + (size_t)(k + 1) * npar);
        return f;
    };
    std::vector<std::vector<double>> f = rows(fitted);
    std::vector<double> c(chi2, chi2 + nsamples);
    if (best2_values) {
        if (nsamples > 0 && (!refitted || !rechi2)) throw RuntimeError("bfh_write_samples_refit: null refit arrays");
        std::vector<std::vector<double>> f2 = rows(refitted);
        std::vector<double> c2(rechi2, rechi2 + nsamples);
        s->analyzer->writeSamples(path, minimum(best_values, best_errors, best_chi2), f, c, nsave, zsave,
                                  minimum(best2_values, best2_errors, best2_chi2), &f2, &c2);
    } else
        s->analyzer->writeSamples(path, minimum(best_values, best_errors, best_chi2), f, c, nsave, zsave);
    BFH_CATCH
}

extern "C" int bfh_mcmc(bfh_session *s, int nchain, int interval, int nwalkers, unsigned long long seed, double *samples) {
    BFH_TRY
    likely::FunctionMinimumPtr fm = s->fitter->fit("mn2::vmetric", "");
    std::vector<double> out;
    s->fitter->mcmc(fm, nchain, interval, out, nwalkers, seed);
    std::copy(out.begin(), out.end(), samples);
    BFH_CATCH
}

extern "C" int bfh_scan_size(const bfh_session *s, long *total) {
    BFH_TRY
    *total = s->analyzer->scanSize();
    BFH_CATCH
}

extern "C" int bfh_parameter_scan(bfh_session *s, long first, long count, double *chi2, double *points, int *status,
                                  double *best) {
    BFH_TRY
    ScanResult r = s->analyzer->parameterScan(s->data, first, count);
    int npar = s->model->getNParameters();
    for (size_t k = 0; k < r.chi2.size(); ++k) {
        if (chi2) chi2[k] = r.chi2[k];
        if (status) status[k] = r.status[k];
        if (points) std::copy(r.points[k].begin(), r.points[k].end(), points + k * npar);
    }
    if (best) {
        likely::Parameters v = r.best->values();
        std::copy(v.begin(), v.end(), best);
        best[npar] = 2 * r.best->minValue;
    }
    BFH_CATCH
}

static void copyOut(int npar, long count, std::vector<std::vector<double>> const &fit, std::vector<double> const &c,
                    std::vector<int> const &st, double *fitted, double *chi2, int *status) {
    for (long t = 0; t < count; ++t) {
        if (fitted) std::copy(fit[t].begin(), fit[t].end(), fitted + (size_t)t * npar);
        if (chi2) chi2[t] = c[t];
        if (status) status[t] = st[t];
    }
}

extern "C" int bfh_toymc(bfh_session *s, long first, long count, unsigned long long seed, double *fitted, double *chi2,
                         int *status) {
    return bfh_toymc_refit(s, first, count, seed, "", "", fitted, chi2, status, nullptr, nullptr, nullptr);
}

extern "C" int bfh_toymc_refit(bfh_session *s, long first, long count, unsigned long long seed, const char *toymc_config,
                               const char *refit_config, double *fitted, double *chi2, int *status, double *refitted, double *rechi2,
                               int *restatus) {
    BFH_TRY
    std::vector<std::vector<double>> fit, fit2;
    std::vector<double> c, c2;
    std::vector<int> st, st2;
    std::string refit = refit_config ? refit_config : "";
    // the truth vector starts from the fit of the combined sample when the session has made one (src/baofit.cc:646-660,788-794)
    s->analyzer->doToyMCSampling(first, count, seed, fit, c, st, toymc_config ? toymc_config : "", refit, &fit2, &c2, &st2,
                                 s->lastFit ? &s->lastFit->parameters : nullptr, s->toymcSaveName);
    int npar = s->model->getNParameters();
    copyOut(npar, count, fit, c, st, fitted, chi2, status);
    if (!refit.empty()) copyOut(npar, count, fit2, c2, st2, refitted, rechi2, restatus);
    BFH_CATCH
}

extern "C" int bfh_set_toymc_save(bfh_session *s, const char *path) {
    BFH_TRY
    if (!s) throw RuntimeError("bfh_set_toymc_save: null session");
    s->toymcSaveName = path ? path : "";
    BFH_CATCH
}

extern "C" int bfh_save_data(bfh_session *s, const char *data_path, const char *icov_path) {
    BFH_TRY
    if (data_path && *data_path) {
        std::ofstream out(data_path);
        if (!out.good()) throw RuntimeError(std::string("bfh_save_data: unable to open ") + data_path);
        s->data->saveData(out);
    }
    if (icov_path && *icov_path) {
        std::ofstream out(icov_path);
        if (!out.good()) throw RuntimeError(std::string("bfh_save_data: unable to open ") + icov_path);
        s->data->saveInverseCovariance(out);
    }
    BFH_CATCH
}

extern "C" int bfh_bootstrap_covariance(bfh_session *s, int trials, unsigned long long seed, const char *icov_path, double *cov, int cov_order,
                                        int *bins, int *posdef) {
    BFH_TRY
    if (!s) throw RuntimeError("bfh_bootstrap_covariance: null session");
    std::vector<int> b;
    std::vector<double> c;
    const bool ok = s->analyzer->estimateCombinedCovariance(trials, seed, icov_path ? icov_path : "", b, c);
    if ((cov || bins) && cov_order != (int)b.size())
        throw RuntimeError("bfh_bootstrap_covariance: the combined data have " + std::to_string(b.size()) + " bins with data");
    if (cov) std::copy(c.begin(), c.end(), cov);
    if (bins) std::copy(b.begin(), b.end(), bins);
    if (posdef) *posdef = ok ? 1 : 0;
    BFH_CATCH
}

// CorrelationAnalyzer::printScaleZEff (CorrelationAnalyzer.cc:127-167): for a fit that floats a BAO scale and gamma-scale, the
// redshift at which scale(z) = scale ((1+z)/(1+zref))^gamma is best determined, the scale there and its error.
extern "C" int bfh_scale_zeff(double scale, double dscale, double gamma, double dgamma, double cov_scale_gamma, double zref,
                              double *zeff, double *scale_eff, double *dscale_eff) {
    BFH_TRY
    if (!(dscale > 0) || !(dgamma > 0) || scale == 0 || gamma == 0) throw RuntimeError("bfh_scale_zeff: scale and gamma-scale must both float (non-zero values, positive errors)");
    const double rho = cov_scale_gamma / (dscale * dgamma);
    const double a = dscale / (scale * dgamma), b = 1 / (2 * gamma);
    const double disc = b * b - a * a * (1 - rho * rho);
    if (disc < 0) throw RuntimeError("bfh_scale_zeff: no real solution for zeff");
    const double logz = -b - a * rho + std::sqrt(disc);
    const double zEff = std::exp(logz) * (1 + zref) - 1;
    const double ratio = (1 + zEff) / (1 + zref), evol = std::pow(ratio, gamma), logRatio = std::log(ratio);
    const double tmp = scale * dgamma * logRatio;
    if (zeff) *zeff = zEff;
    if (scale_eff) *scale_eff = scale * evol;
    if (dscale_eff) *dscale_eff = evol * std::sqrt(dscale * dscale + 2 * rho * scale * dscale * dgamma * logRatio + tmp * tmp);
    BFH_CATCH
}

extern "C" double bfh_gamma_q(double a, double x) {
    try { return gammaQ(a, x); } catch (std::exception const &e) { g_herr = e.what(); return -1.0; }
}

extern "C" int bfh_compare_each(bfh_session *s, const char *path, int finalized, double *prob, double *chi2, double *logdet) {
    BFH_TRY
    if (!s) throw RuntimeError("bfh_compare_each: null session");
    std::vector<double> p, c, l;
    s->analyzer->compareEach(path ? path : "", finalized != 0, p, c, l);
    if (prob) std::copy(p.begin(), p.end(), prob);
    if (chi2) std::copy(c.begin(), c.end(), chi2);
    if (logdet) std::copy(l.begin(), l.end(), logdet);
    BFH_CATCH
}

extern "C" int bfh_symmetric_eigenmodes(int n, const double *matrix, double *values, double *vectors) {
    BFH_TRY
    if (n <= 0 || !matrix || !values || !vectors) throw RuntimeError("bfh_symmetric_eigenmodes: bad arguments");
    std::vector<double> vals, vecs;
    symmetricEigenModes(n, std::vector<double>(matrix, matrix + (size_t)n * n), vals, vecs);
    std::copy(vals.begin(), vals.end(), values);
    std::copy(vecs.begin(), vecs.end(), vectors);
    BFH_CATCH
}

extern "C" int bfh_nobservations(const bfh_session *s) { return s && s->resampler ? s->resampler->getNObservations() : 0; }

extern "C" int bfh_resampling_size(bfh_session *s, const char *kind, int n, long *size) {
    BFH_TRY
    *size = s->analyzer->resamplingSize(kind, n);
    BFH_CATCH
}

extern "C" int bfh_bootstrap_counts(bfh_session *s, int size, unsigned long long seed, long trial, int *counts) {
    BFH_TRY
    if (!s->resampler) throw RuntimeError("bfh_bootstrap_counts: the session has no individual observations.");
    std::vector<int> c = s->resampler->bootstrapCounts(size, seed, trial);
    std::copy(c.begin(), c.end(), counts);
    BFH_CATCH
}

extern "C" int bfh_resample_data(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                                 long seqno, double *data, double *cov) {
    BFH_TRY
    AbsCorrelationDataPtr sample = s->analyzer->resample(kind, n, size, fix_covariance != 0, seed, seqno);
    if (sample->keptIndices() != s->data->keptIndices()) throw RuntimeError("bfh_resample_data: the sample keeps different bins.");
    if (data) { std::vector<double> d = sample->dataVector(); std::copy(d.begin(), d.end(), data); }
    if (cov) { std::vector<double> c = sample->covarianceMatrix(); std::copy(c.begin(), c.end(), cov); }
    BFH_CATCH
}

extern "C" int bfh_resample_fit(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                                long first, long count, double *fitted, double *chi2, int *status) {
    return bfh_resample_refit(s, kind, n, size, fix_covariance, seed, first, count, "", fitted, chi2, status, nullptr, nullptr, nullptr);
}

extern "C" int bfh_resample_refit(bfh_session *s, const char *kind, int n, int size, int fix_covariance, unsigned long long seed,
                                  long first, long count, const char *refit_config, double *fitted, double *chi2, int *status,
                                  double *refitted, double *rechi2, int *restatus) {
    BFH_TRY
    std::vector<std::vector<double>> fit, fit2;
    std::vector<double> c, c2;
    std::vector<int> st, st2;
    std::string refit = refit_config ? refit_config : "";
    long done = s->analyzer->doResampling(kind, n, size, fix_covariance != 0, seed, first, count, fit, c, st, refit, &fit2, &c2, &st2);
    int npar = s->model->getNParameters();
    copyOut(npar, done, fit, c, st, fitted, chi2, status);
    if (!refit.empty()) copyOut(npar, done, fit2, c2, st2, refitted, rechi2, restatus);
    BFH_CATCH
}

extern "C" int bfh_minimize(int npar, const double *values, const double *errors, const int *floating, bfh_batch_fn fn,
                            void *user, double *out_values, double *out_errors, double *fval, int *status,
                            int *nevaluations) {
    BFH_TRY
    likely::FitParameters params(npar);
    for (int i = 0; i < npar; ++i) {
        params[i].name = "p" + std::to_string(i);
        params[i].value = values[i];
        params[i].error = errors[i];
        params[i].floating = floating[i] != 0;
    }
    std::vector<likely::NewtonFit> fits(1, likely::NewtonFit(params));
    likely::runLockStep(fits, [&](std::vector<likely::Parameters> const &pts, std::vector<double> &f) {
        std::vector<double> flat(pts.size() * (size_t)npar);
        for (size_t b = 0; b < pts.size(); ++b) std::copy(pts[b].begin(), pts[b].end(), flat.begin() + b * npar);
        fn(flat.data(), (int)pts.size(), f.data(), user);
    });
    likely::FunctionMinimumPtr fm = fits[0].result();
    for (int i = 0; i < npar; ++i) {
        if (out_values) out_values[i] = fm->parameters[i].value;
        if (out_errors) out_errors[i] = fm->parameters[i].floating ? fm->parameters[i].error : 0.0;
    }
    if (fval) *fval = fm->minValue;
    if (status) *status = (int)fm->status;
    if (nevaluations) *nevaluations = fm->nEvaluations;
    BFH_CATCH
}

extern "C" int bfh_lm_minimize(int npar, const double *values, const double *errors, const int *floating,
                               const double *prior_centre, const double *prior_sigma, int nres, bfh_resid_fn fn, void *user,
                               double *out_values, double *fval, int *status, int *nevaluations, int *niterations) {
    BFH_TRY
    likely::FitParameters params(npar);
    for (int i = 0; i < npar; ++i) {
        params[i].name = "p" + std::to_string(i);
        params[i].value = values[i];
        params[i].error = errors[i];
        params[i].floating = floating[i] != 0;
        if (prior_sigma && prior_sigma[i] > 0) {
            params[i].priorType = likely::FitParameter::GaussPrior;
            params[i].priorMin = prior_centre[i] - prior_sigma[i];
            params[i].priorMax = prior_centre[i] + prior_sigma[i];
        }
    }
    likely::LMFit fit(params);
    std::vector<double> flat, resid, gram, chi2;
    while (!fit.done()) {
        flat.clear();
        if (fit.wantsJacobian()) {
            const int G = fit.groupSize();
            fit.requestJacobian(flat);
            resid.assign((size_t)G * nres, 0.0);
            fn(flat.data(), G, resid.data(), user);
            gram.assign((size_t)G * G, 0.0);
            bool bad = false;
            for (int i = 0; i < G; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = 0;
                    for (int r = 0; r < nres; ++r) {
                        double a = i == 0 ? resid[r] : resid[(size_t)i * nres + r] - resid[r];
                        double b = j == 0 ? resid[r] : resid[(size_t)j * nres + r] - resid[r];
                        s += a * b;
                    }
                    gram[i * G + j] = gram[j * G + i] = s;
                    bad |= !std::isfinite(s);
                }
            fit.consumeJacobian(gram.data(), bad);
        } else {
            fit.requestPoints(flat);
            const int B = fit.pending();
            resid.assign((size_t)B * nres, 0.0);
            fn(flat.data(), B, resid.data(), user);
            chi2.assign(B, 0.0);
            for (int b = 0; b < B; ++b)
                for (int r = 0; r < nres; ++r) chi2[b] += resid[(size_t)b * nres + r] * resid[(size_t)b * nres + r];
            fit.consumePoints(chi2.data());
        }
    }
    likely::FunctionMinimumPtr fm = fit.result();
    for (int i = 0; i < npar; ++i)
        if (out_values) out_values[i] = fm->parameters[i].value;
    if (fval) *fval = fm->minValue;
    if (status) *status = (int)fm->status;
    if (nevaluations) *nevaluations = fm->nEvaluations;
    if (niterations) *niterations = fm->nIterations;
    BFH_CATCH
}
