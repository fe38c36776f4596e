// CorrelationFitter (the objective, batched) and CorrelationAnalyzer (fit, scan, toy MC drivers).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <fstream>
#include <random>
#include <iostream>
#include <limits>
#include <sstream>

#include <cstdio>
#include <cstdlib>

#include "baofit_host.h"
#include "../csrc/host_setup.h"

namespace baofit {

namespace {
void check(int rc, const char *what) {
    if (rc != BF_OK) throw RuntimeError(std::string(what) + ": " + bf_last_error());
}
}  // namespace

CorrelationFitter::CorrelationFitter(AbsCorrelationDataPtr data, AbsCorrelationModelPtr model, int covSampleSize)
    : _data(data), _model(model), _errorScale(1) {
    if (!data || 0 == data->getNBinsWithData()) throw RuntimeError("CorrelationFitter: need some data to fit.");
    if (!model) throw RuntimeError("CorrelationFitter: need a model to fit.");
    if (!data->isFinalized()) data->finalize();
    int n = (int)data->keptIndices().size();
    if (0 < covSampleSize) {
        if (covSampleSize <= n + 2) throw RuntimeError("CorrelationFitter: cannot fit with covSampleSize <= nbins+2");
        // eqn (25) of http://arxiv.org/abs/1212.4359, CorrelationFitter.cc:31-37
        _icovScale = (covSampleSize - n - 2.) / (covSampleSize - 1.);
    } else
        _icovScale = 1;
}

void CorrelationFitter::setErrorScale(double scale) {
    if (scale <= 0) throw RuntimeError("CorrelationFitter::setErrorScale: expected scale > 0.");
    _errorScale = scale;
}

static bool needsGrid(AbsCorrelationModelPtr const &model) { return model->description().dist_matrix != nullptr; }

void CorrelationFitter::predictBatch(std::vector<likely::Parameters> const &params, std::vector<double> &pred) const {
    int B = (int)params.size(), npar = _model->getNParameters(), n = (int)_data->keptIndices().size();
    std::vector<double> flat((size_t)B * npar);
    for (int b = 0; b < B; ++b) {
        if ((int)params[b].size() != npar) throw RuntimeError("CorrelationFitter: got unexpected number of parameters.");
        std::copy(params[b].begin(), params[b].end(), flat.begin() + (size_t)b * npar);
    }
    pred.assign((size_t)n * B, 0.0);
    std::vector<int> status(B);
    check(bf_eval_batch(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), flat.data(), B,
                        pred.data(), B, status.data()),
          "bf_eval_batch");
}

void CorrelationFitter::getPrediction(likely::Parameters const &params, std::vector<double> &prediction) const {
    std::vector<likely::Parameters> one(1, params);
    predictBatch(one, prediction);
}

void CorrelationFitter::evaluateBatch(std::vector<likely::Parameters> const &params, std::vector<double> &fval,
                                      std::vector<std::vector<double>> const *dataOverride) const {
    int B = (int)params.size(), npar = _model->getNParameters(), n = (int)_data->keptIndices().size();
    fval.assign(B, 0.0);
    if (B == 0) return;
    std::vector<double> flat((size_t)B * npar), ov;
    for (int b = 0; b < B; ++b) {
        if ((int)params[b].size() != npar) throw RuntimeError("CorrelationFitter: got unexpected number of parameters.");
        std::copy(params[b].begin(), params[b].end(), flat.begin() + (size_t)b * npar);
    }
    if (dataOverride) {
        if ((int)dataOverride->size() != B) throw RuntimeError("CorrelationFitter: one data vector per parameter vector expected.");
        ov.resize((size_t)B * n);
        for (int b = 0; b < B; ++b) {
            if ((int)(*dataOverride)[b].size() != n) throw RuntimeError("CorrelationFitter: data override has the wrong length.");
            std::copy((*dataOverride)[b].begin(), (*dataOverride)[b].end(), ov.begin() + (size_t)b * n);
        }
    }
    std::vector<double> chi2(B);
    std::vector<int> status(B);
    check(bf_chi2_batch(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), flat.data(), B,
                        dataOverride ? ov.data() : nullptr, chi2.data(), status.data()),
          "bf_chi2_batch");
    likely::FitParameters const &fp = _model->getFitParameters();
    for (int b = 0; b < B; ++b) {
        // CorrelationFitter.cc:85; a vector the reference would have thrown on evaluates to NaN
        fval[b] = status[b] ? std::numeric_limits<double>::quiet_NaN()
                            : (0.5 * _icovScale * chi2[b] + likely::evaluatePriors(fp, params[b])) / _errorScale;
    }
}

void CorrelationFitter::evaluateBatchToys(std::vector<likely::Parameters> const &params, std::vector<double> &fval,
                                          bf_toys const *toys, std::vector<int> const &toyIndex) const {
    int B = (int)params.size(), npar = _model->getNParameters();
    fval.assign(B, 0.0);
    if (B == 0) return;
    if ((int)toyIndex.size() != B) throw RuntimeError("CorrelationFitter: one toy index per parameter vector expected.");
    std::vector<double> flat((size_t)B * npar), chi2(B);
    std::vector<int> status(B);
    for (int b = 0; b < B; ++b) {
        if ((int)params[b].size() != npar) throw RuntimeError("CorrelationFitter: got unexpected number of parameters.");
        std::copy(params[b].begin(), params[b].end(), flat.begin() + (size_t)b * npar);
    }
    check(bf_chi2_batch_toys(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), flat.data(), B, toys,
                             toyIndex.data(), chi2.data(), status.data()),
          "bf_chi2_batch_toys");
    likely::FitParameters const &fp = _model->getFitParameters();
    for (int b = 0; b < B; ++b)
        fval[b] = status[b] ? std::numeric_limits<double>::quiet_NaN()
                            : (0.5 * _icovScale * chi2[b] + likely::evaluatePriors(fp, params[b])) / _errorScale;
}

void CorrelationFitter::chiSquareFlat(std::vector<double> const &params, int B, std::vector<double> &chi2, bf_toys const *toys,
                                      const int *toyIndex) const {
    chi2.assign(B, 0.0);
    if (B == 0) return;
    std::vector<int> status(B);
    bf_data *dd = _data->deviceData(needsGrid(_model));
    if (toys)
        check(bf_chi2_batch_toys(deviceContext(), _model->deviceModel(), dd, params.data(), B, toys, toyIndex, chi2.data(), status.data()),
              "bf_chi2_batch_toys");
    else
        check(bf_chi2_batch(deviceContext(), _model->deviceModel(), dd, params.data(), B, nullptr, chi2.data(), status.data()), "bf_chi2_batch");
    for (int b = 0; b < B; ++b)
        if (status[b]) chi2[b] = std::numeric_limits<double>::quiet_NaN();
}

void CorrelationFitter::gramFlat(std::vector<double> const &params, int nfit, int G, std::vector<double> &gram, std::vector<char> &bad,
                                 bf_toys const *toys, const int *toyIndexPerFit) const {
    gram.assign((size_t)nfit * G * G, 0.0);
    bad.assign(nfit, 0);
    if (nfit == 0) return;
    std::vector<int> status((size_t)nfit * G);
    check(bf_gram_batch(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), params.data(), nfit, G, toys,
                        toys ? toyIndexPerFit : nullptr, gram.data(), status.data()),
          "bf_gram_batch");
    for (int f = 0; f < nfit; ++f)
        for (int g = 0; g < G; ++g) bad[f] |= status[(size_t)f * G + g] != 0;
}

namespace {
double nowSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

void CorrelationFitter::gramProbes(std::vector<double> const &centres, std::vector<double> const &steps, int nfit, int nprobe,
                                   std::vector<double> &gram, std::vector<char> &bad, bf_toys const *toys, const int *toyIndexPerFit) const {
    const int G = 2 * nprobe + 1;
    gram.assign((size_t)nfit * G * G, 0.0);
    bad.assign(nfit, 0);
    if (nfit == 0) return;
    std::vector<int> status(nfit);
    check(bf_gram_probes(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), centres.data(), steps.data(), nfit,
                         nprobe, toys, toys ? toyIndexPerFit : nullptr, gram.data(), status.data()),
          "bf_gram_probes");
    for (int f = 0; f < nfit; ++f) bad[f] = status[f] != 0;
}

namespace {
// page-locked result buffer that grows on demand (one per host thread): device-to-host copies of the normal
// equations and chi-squares of a round run at full PCIe speed instead of through the driver's staging buffers
struct PinnedBuffer {
    double *p = nullptr;
    size_t n = 0;
    ~PinnedBuffer() { if (p) bf_pinned_free(p); }
    double *get(size_t count) {
        if (count > n) {
            if (p) bf_pinned_free(p);
            p = nullptr; n = 0;
            void *q = nullptr;
            check(bf_pinned_alloc((count + count / 4) * sizeof(double), &q), "bf_pinned_alloc");
            p = (double *)q; n = count + count / 4;
        }
        return p;
    }
};
}  // namespace

const double *CorrelationFitter::normalProbes(std::vector<double> const &centres, std::vector<double> const &steps, int nfit, int nprobe,
                                              std::vector<char> &bad, bf_toys const *toys, const int *toyIndexPerFit) const {
    static thread_local PinnedBuffer buffer;
    const size_t per = (size_t)(nprobe + 1) * (nprobe + 1) + nprobe;
    bad.assign(nfit, 0);
    if (nfit == 0) return nullptr;
    double *out = buffer.get(per * nfit);
    std::vector<int> status(nfit);
    check(bf_normal_probes(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), centres.data(), steps.data(), nfit,
                           nprobe, toys, toys ? toyIndexPerFit : nullptr, out, status.data()),
          "bf_normal_probes");
    for (int f = 0; f < nfit; ++f) bad[f] = status[f] != 0;
    return out;
}

void CorrelationFitter::runLockStepLM(std::vector<likely::LMFit> &fits, bf_toys const *toys, std::vector<int> const *toyOfFit) const {
    const bool verbose = std::getenv("BAOFIT_VERBOSE_FITS") != nullptr;
    // BAOFIT_VERBOSE_FITS >= 3: per-kernel CUDA-event times of the whole run (perturbs the wall times printed below)
    const bool kernelProfile = verbose && std::atoi(std::getenv("BAOFIT_VERBOSE_FITS")) >= 3;
    if (kernelProfile) bf_ctx_set_profiling(deviceContext(), 1);
    double tGram = 0, tChi2 = 0, tSolve = 0, t0 = nowSeconds();
    long nGram = 0, nChi2 = 0, rounds = 0;
    std::vector<size_t> offsets;
    std::vector<double> flatJ, steps, flatP, gram, chi2;
    std::vector<char> bad;
    std::vector<int> jfits, pfits, jtoy, ptoy;
    for (;;) {
        // fits are bucketed by group size so that one bf_gram_batch call serves all fits of a size
        std::map<int, std::vector<int>> byG;
        pfits.clear();
        for (size_t k = 0; k < fits.size(); ++k) {
            if (fits[k].done()) continue;
            if (fits[k].wantsJacobian()) byG[fits[k].groupSize()].push_back((int)k);
            else pfits.push_back((int)k);
        }
        if (byG.empty() && pfits.empty()) break;
        ++rounds;
        if (verbose && std::atoi(std::getenv("BAOFIT_VERBOSE_FITS")) == 4) {
            size_t nj = 0;
            for (auto const &kv : byG) nj += kv.second.size();
            std::fprintf(stderr, "  round %ld: %zu Jacobian fits, %zu trial fits, t = %.3f s\n", rounds, nj, pfits.size(), nowSeconds() - t0);
        }
        // The trial points of this round go first and asynchronously (bf_chi2_batch_submit): their evaluation is queued ahead
        // of the probe groups below, and the host waits only once per round instead of once per call.
        int ticket = -1, nPoints = 0;
        double *pinnedChi2 = nullptr;
        int *pinnedStatus = nullptr;
        if (!pfits.empty()) {
            flatP.clear();
            ptoy.clear();
            for (int k : pfits) {
                fits[k].requestPoints(flatP);
                if (toys) ptoy.insert(ptoy.end(), fits[k].pending(), (*toyOfFit)[k]);
            }
            nPoints = (int)(flatP.size() / fits[pfits[0]].nPar());
            if (!toys && !byG.empty()) {
                static thread_local PinnedBuffer inBuf, outBuf, stBuf;
                double *pin = inBuf.get(flatP.size());
                std::memcpy(pin, flatP.data(), flatP.size() * sizeof(double));
                pinnedChi2 = outBuf.get(nPoints);
                pinnedStatus = reinterpret_cast<int *>(stBuf.get((nPoints + 1) / 2));
                double ta = nowSeconds();
                check(bf_chi2_batch_submit(deviceContext(), _model->deviceModel(), _data->deviceData(needsGrid(_model)), pin, nPoints, pinnedChi2,
                                           pinnedStatus, &ticket),
                      "bf_chi2_batch_submit");
                tChi2 += nowSeconds() - ta;
            }
        }
        for (auto const &kv : byG) {
            const int G = kv.first;
            if (G > 95) {
                // more floating parameters than one probe group holds (bf_normal_probes: 2 n + 1 <= 96): these fits
                // are left to the Newton driver, which every caller runs for fits that do not end OK
                for (int k : kv.second) fits[k].abandon();
                continue;
            }
            flatJ.clear();
            steps.clear();
            jtoy.clear();
            for (int k : kv.second) {
                fits[k].requestProbes(flatJ, steps);
                if (toys) jtoy.push_back((*toyOfFit)[k]);
            }
            double ta = nowSeconds();
            const int nprobe = (G - 1) / 2;
            const double *normal = normalProbes(flatJ, steps, (int)kv.second.size(), nprobe, bad, toys, toys ? jtoy.data() : nullptr);
            const size_t per = (size_t)(nprobe + 1) * (nprobe + 1) + nprobe;
            tGram += nowSeconds() - ta;
            nGram += (long)kv.second.size() * G;
            // the normal equations of the fits are independent: solved on all host threads
            ta = nowSeconds();
            const long nq = (long)kv.second.size();
            std::vector<int> const &members = kv.second;
#pragma omp parallel for schedule(static) if (nq >= 128)
            for (long q = 0; q < nq; ++q) fits[members[q]].consumeNormal(normal + (size_t)q * per, bad[q] != 0);
            tSolve += nowSeconds() - ta;
        }
        if (!pfits.empty()) {
            double ta = nowSeconds();
            if (ticket >= 0) {
                check(bf_chi2_batch_wait(deviceContext(), ticket), "bf_chi2_batch_wait");
                chi2.assign(pinnedChi2, pinnedChi2 + nPoints);
                for (int b = 0; b < nPoints; ++b)
                    if (pinnedStatus[b]) chi2[b] = std::numeric_limits<double>::quiet_NaN();
            } else
                chiSquareFlat(flatP, nPoints, chi2, toys, toys ? ptoy.data() : nullptr);
            tChi2 += nowSeconds() - ta;
            nChi2 += (long)nPoints;
            ta = nowSeconds();
            offsets.assign(pfits.size(), 0);
            size_t off = 0;
            for (size_t q = 0; q < pfits.size(); ++q) { offsets[q] = off; off += fits[pfits[q]].pending(); }
            const long np = (long)pfits.size();
#pragma omp parallel for schedule(static) if (np >= 128)
            for (long q = 0; q < np; ++q) fits[pfits[q]].consumePoints(chi2.data() + offsets[q]);
            tSolve += nowSeconds() - ta;
        }
    }
    if (verbose)
        std::fprintf(stderr, "runLockStepLM: %zu fits, %ld rounds, %.3f s total: bf_normal_probes %.3f s (%ld points), bf_chi2_batch %.3f s (%ld points), "
                             "per-fit solves %.3f s\n",
                     fits.size(), rounds, nowSeconds() - t0, tGram, nGram, tChi2, nChi2, tSolve);
    if (kernelProfile) {
        const char *names[64];
        double ms[64];
        long long launches[64];
        int n = std::min(64, bf_ctx_get_profile(deviceContext(), 64, names, ms, launches));
        for (int k = 0; k < n; ++k) std::fprintf(stderr, "  %-48s %9.3f ms in %6lld launches\n", names[k], ms[k], launches[k]);
        bf_ctx_set_profiling(deviceContext(), 0);
    }
}

void CorrelationFitter::mcmc(likely::FunctionMinimumPtr fminStart, int nchain, int interval, std::vector<double> &samples, int nwalkers,
                             unsigned long long seed) const {
    if (!fminStart || fminStart->covariance.empty()) throw RuntimeError("CorrelationFitter::mcmc: the starting minimum has no covariance.");
    if (nchain <= 0 || interval <= 0 || nwalkers <= 0) throw RuntimeError("CorrelationFitter::mcmc: expected nchain, interval, nwalkers > 0.");
    likely::FitParameters const &fp = fminStart->parameters;
    const int npar = (int)fp.size();
    std::vector<int> fidx;
    for (int i = 0; i < npar; ++i)
        if (fp[i].floating) fidx.push_back(i);
    const int n = (int)fidx.size();
    // Cholesky factor of the proposal covariance
    std::vector<double> L(fminStart->covariance);
    for (int j = 0; j < n; ++j) {
        double d = L[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0)) throw RuntimeError("CorrelationFitter::mcmc: covariance is not positive definite.");
        L[j * n + j] = std::sqrt(d);
        for (int i = j + 1; i < n; ++i) {
            double t = L[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / L[j * n + j];
        }
    }
    const double scale = 2.38 / std::sqrt((double)std::max(n, 1));
    nwalkers = std::min(nwalkers, nchain);
    const int perWalker = (nchain + nwalkers - 1) / nwalkers;
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> gauss(0.0, 1.0);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    std::vector<likely::Parameters> x(nwalkers, fminStart->values()), trial(nwalkers);
    std::vector<double> fx, ft, g(n);
    evaluateBatch(x, fx);
    samples.clear();
    samples.reserve((size_t)nchain * (npar + 1));
    for (int step = 1; step <= perWalker * interval && (int)samples.size() < nchain * (npar + 1); ++step) {
        for (int w = 0; w < nwalkers; ++w) {
            for (int i = 0; i < n; ++i) g[i] = gauss(rng);
            trial[w] = x[w];
            for (int i = 0; i < n; ++i) {
                double s = 0;
                for (int k = 0; k <= i; ++k) s += L[i * n + k] * g[k];
                trial[w][fidx[i]] += scale * s;
            }
        }
        evaluateBatch(trial, ft);  // one batch on the GPU per Metropolis step
        for (int w = 0; w < nwalkers; ++w) {
            // fval = -log L; a vector outside the model's valid domain (NaN) is always rejected
            if (std::isfinite(ft[w]) && std::log(unif(rng)) < fx[w] - ft[w]) { x[w] = trial[w]; fx[w] = ft[w]; }
        }
        if (step % interval == 0)
            for (int w = 0; w < nwalkers && (int)samples.size() < nchain * (npar + 1); ++w) {
                samples.insert(samples.end(), x[w].begin(), x[w].end());
                samples.push_back(fx[w]);
            }
    }
}

double CorrelationFitter::operator()(likely::Parameters const &params) const {
    if ((int)params.size() != _model->getNParameters()) throw RuntimeError("CorrelationFitter: got unexpected number of parameters.");
    std::vector<likely::Parameters> one(1, params);
    std::vector<double> f;
    evaluateBatch(one, f);
    if (!std::isfinite(f[0])) throw RuntimeError("BaoKSpaceCorrelationModel: hit dilation limit.");
    return f[0];
}

likely::FunctionMinimumPtr CorrelationFitter::fit(std::string const &methodName, std::string const &config) const {
    (void)methodName;
    likely::FitParameters params = _model->getFitParameters();
    if (!config.empty()) likely::modifyFitParameters(params, config);
    std::vector<likely::NewtonFit> fits(1, likely::NewtonFit(params));
    likely::runLockStep(fits, [this](std::vector<likely::Parameters> const &pts, std::vector<double> &f) { evaluateBatch(pts, f); });
    return fits[0].result();
}

// ------------------------------------------------------------------------------------------
// CorrelationAnalyzer
// ------------------------------------------------------------------------------------------
CorrelationAnalyzer::CorrelationAnalyzer(std::string const &method, double rmin, double rmax, int covSampleSize, bool verbose)
    : _method(method), _rmin(rmin), _rmax(rmax), _covSampleSize(covSampleSize), _verbose(verbose) {}

likely::FunctionMinimumPtr CorrelationAnalyzer::fitSample(AbsCorrelationDataPtr sample, std::string const &config) const {
    CorrelationFitter fitter(sample, _model, _covSampleSize);
    return fitter.fit(_method, config);
}

long CorrelationAnalyzer::scanSize() const {
    long n = 1;
    bool any = false;
    for (auto const &p : _model->getFitParameters())
        if (!p.binning.empty()) { n *= (long)p.binning.size(); any = true; }
    return any ? n : 0;
}

ScanResult CorrelationAnalyzer::parameterScan(AbsCorrelationDataPtr sample, long first, long count) const {
    ScanResult out;
    CorrelationFitter fitter(sample, _model, _covSampleSize);
    likely::FitParameters base = _model->getFitParameters();
    // one axis per parameter with a binning, in parameter order; last axis fastest
    // (likely::getFitParametersGrid, CorrelationAnalyzer.cc:578-584)
    std::vector<int> axes;
    for (size_t i = 0; i < base.size(); ++i)
        if (!base[i].binning.empty()) { axes.push_back((int)i); out.axisNames.push_back(base[i].name); }
    if (axes.empty()) throw RuntimeError("CorrelationAnalyzer::parameterScan: no parameter has a binning.");
    long total = scanSize();
    if (count < 0) count = total - first;
    if (first < 0 || first + count > total) throw RuntimeError("CorrelationAnalyzer::parameterScan: bad grid range.");
    // the unconstrained best fit keeps the Newton driver (its HESSE errors go into the scan header);
    // the grid fits, whose covariance nobody reads, run Levenberg-Marquardt on Gauss-Newton normal
    // equations from the GPU
    const double tScan0 = nowSeconds();
    std::vector<likely::NewtonFit> bestFit(1, likely::NewtonFit(base));
    likely::runLockStep(bestFit, [&fitter](std::vector<likely::Parameters> const &pts, std::vector<double> &f) { fitter.evaluateBatch(pts, f); });
    const double tScan1a = nowSeconds();
    const double cs = fitter.icovScale() / fitter.errorScale(), ps = 1.0 / fitter.errorScale();
    // (built on all host threads: a FitParameters copy is 37 heap-allocated names and two binning vectors, 2100 of them
    // were 5 ms of a 40 ms scan)
    std::vector<likely::LMFit> fits((size_t)count);
    std::vector<likely::Parameters> starts((size_t)count);
    likely::FitParameters lean = base;
    for (auto &q : lean) q.binning.clear();  // a grid fit has no use for the axes
    const bool traceFirst = std::getenv("BAOFIT_VERBOSE_FITS") && std::atoi(std::getenv("BAOFIT_VERBOSE_FITS")) > 1;
#pragma omp parallel for schedule(static) if (count >= 256)
    for (long k = 0; k < count; ++k) {
        likely::FitParameters p = lean;
        long rem = first + k;
        for (int a = (int)axes.size() - 1; a >= 0; --a) {
            int nb = (int)base[axes[a]].binning.size();
            int ib = (int)(rem % nb);
            rem /= nb;
            p[axes[a]].value = base[axes[a]].binning[ib];
            p[axes[a]].floating = false;  // "fix[name]=value"
        }
        // each grid fit starts from the model's initial configuration, not from the best fit
        // (CorrelationAnalyzer.cc:591)
        fits[k] = likely::LMFit(p, cs, ps);
        if (k == 0 && traceFirst) fits[k].debug = true;
        likely::Parameters st;
        for (auto const &q : p) st.push_back(q.value);
        starts[k] = st;
    }
    const double tScan1 = nowSeconds();
    fitter.runLockStepLM(fits);
    const double tScan2 = nowSeconds();
    // grid points where Levenberg-Marquardt did not converge cleanly (flat or saddle-like regions,
    // where the Gauss-Newton curvature is a poor model) are refitted with the Newton driver
    std::vector<likely::FunctionMinimumPtr> results(count);
    std::vector<likely::NewtonFit> redo;
    std::vector<long> redoIdx;
#pragma omp parallel for schedule(static) if (count >= 256)
    for (long k = 0; k < count; ++k) results[k] = fits[k].result();
    for (long k = 0; k < count; ++k) {
        if (results[k]->status != likely::FunctionMinimum::OK) {
            // ... from where LM stopped when that point is usable (it is usually close to the minimum, only not
            // certified), from the initial configuration otherwise
            likely::FitParameters p = results[k]->parameters;
            const bool usable = results[k]->status == likely::FunctionMinimum::WARNING && std::isfinite(results[k]->minValue);
            for (size_t i = 0; i < p.size(); ++i) {
                if (!usable) p[i].value = starts[k][i];
                p[i].error = base[i].error;
            }
            redo.push_back(likely::NewtonFit(p));
            // a refit that continues from LM's end point either certifies it within a few iterations or sits in a
            // flat valley that no iteration count certifies (most of the time of a k-space scan went there)
            if (usable) redo.back().maxIterations = 10;
            redoIdx.push_back(k);
        }
    }
    if (std::getenv("BAOFIT_VERBOSE_FITS")) {
        long nev = 0;
        for (auto const &r : results) nev += r->nEvaluations;
        // the slowest grid fits set the number of lock-step rounds: who are they
        std::vector<long> order(count);
        for (long k = 0; k < count; ++k) order[k] = k;
        std::sort(order.begin(), order.end(), [&](long a, long b) { return results[a]->nIterations > results[b]->nIterations; });
        for (long q = 0; q < std::min<long>(4, count); ++q) {
            likely::FunctionMinimumPtr r = results[order[q]];
            std::fprintf(stderr, "parameterScan: slow fit %ld: %d iterations, status %d, f %.6f edm %.3e, x", first + order[q], r->nIterations, (int)r->status,
                         r->minValue, r->edm);
            for (auto const &fp : r->parameters)
                if (fp.floating) std::fprintf(stderr, " %s=%.6g", fp.name.c_str(), fp.value);
            std::fprintf(stderr, "\n");
        }
        std::fprintf(stderr, "parameterScan: %ld grid fits, %zu handed to the Newton driver, %ld LM evaluations; unconstrained fit %.4f s, "
                             "fit set-up %.4f s, lock-step LM %.4f s\n", count, redo.size(), nev, tScan1a - tScan0, tScan1 - tScan1a, tScan2 - tScan1);
    }
    if (!redo.empty()) {
        long rounds = 0, points = 0;
        likely::runLockStep(redo, [&](std::vector<likely::Parameters> const &pts, std::vector<double> &f) { ++rounds; points += (long)pts.size(); fitter.evaluateBatch(pts, f); });
        if (std::getenv("BAOFIT_VERBOSE_FITS"))
            std::fprintf(stderr, "parameterScan: unconstrained fit %.3f s, lock-step LM %.3f s, Newton refits %.3f s (%ld rounds, %ld points)\n",
                         tScan1 - tScan0, tScan2 - tScan1, nowSeconds() - tScan2, rounds, points);
        for (size_t q = 0; q < redo.size(); ++q) {
            likely::FunctionMinimumPtr fm = redo[q].result();
            bool better = fm->status != likely::FunctionMinimum::ERROR &&
                          (results[redoIdx[q]]->status == likely::FunctionMinimum::ERROR || fm->minValue <= results[redoIdx[q]]->minValue);
            if (better) results[redoIdx[q]] = fm;
        }
    }
    out.best = bestFit[0].result();
    for (long k = 0; k < count; ++k) {
        likely::FunctionMinimumPtr fm = results[k];
        bool ok = fm->status != likely::FunctionMinimum::ERROR;
        out.points.push_back(fm->values());
        out.chi2.push_back(ok ? 2 * fm->minValue : 0.0);  // failed fits are saved with fval = 0 (:595-602)
        out.status.push_back((int)fm->status);
    }
    return out;
}

void CorrelationAnalyzer::dumpResiduals(std::ostream &out, likely::FunctionMinimumPtr fmin, std::string const &script,
                                        bool dumpGradients) const {
    if (!_data || !_model) throw RuntimeError("CorrelationAnalyzer::dumpResiduals: no observations have been added.");
    CorrelationFitter fitter(_data, _model, _covSampleSize);
    likely::FitParameters parameters(fmin->parameters);
    if (!script.empty()) likely::modifyFitParameters(parameters, script);
    const int npar = (int)parameters.size();
    likely::Parameters values, errors;
    for (auto const &p : parameters) { values.push_back(p.value); errors.push_back(p.floating ? p.error : 0.0); }
    // one batch: the prediction itself, then +-0.5 dpar for every parameter with an error (:656-670)
    std::vector<likely::Parameters> points(1, values);
    std::vector<int> slot(npar, -1);
    if (dumpGradients)
        for (int j = 0; j < npar; ++j) {
            double dpar = 0.1 * errors[j];
            if (!(dpar > 0)) continue;
            slot[j] = (int)points.size();
            likely::Parameters hi(values), lo(values);
            hi[j] = values[j] + 0.5 * dpar;
            lo[j] = values[j] - 0.5 * dpar;
            points.push_back(hi);
            points.push_back(lo);
        }
    std::vector<double> pred;
    fitter.predictBatch(points, pred);
    const int B = (int)points.size();
    std::vector<int> const &kept = _data->keptIndices();
    const bool coord = _data->getTransverseBinningType() == AbsCorrelationData::Coordinate;
    char buf[64];
    auto num = [&](double v) { std::snprintf(buf, sizeof(buf), "%.12g", v); return std::string(buf); };
    for (size_t i = 0; i < kept.size(); ++i) {
        const int index = kept[i];
        double c[3];
        _data->getBinCenters(index, c);  // custom centres if present, CorrelationAnalyzer.cc:630-635
        out << index << ' ' << num(c[0]) << ' ' << num(c[1]) << ' ' << num(c[2]) << ' ' << num(_data->getRadius(index)) << ' ';
        if (coord) out << num(_data->getCosAngle(index));
        else out << _data->getMultipole(index);
        double cii = _data->getCovarianceElement(index, index);
        double error = _data->isInverse() ? 0.0 : std::sqrt(std::max(cii, 0.0));
        out << ' ' << num(_data->getRedshift(index)) << ' ' << num(pred[i * B]) << ' ' << num(_data->getData(index)) << ' ' << num(error);
        if (dumpGradients)
            for (int j = 0; j < npar; ++j) {
                double g = 0;
                if (slot[j] >= 0) g = (pred[i * B + slot[j]] - pred[i * B + slot[j] + 1]) / (0.1 * errors[j]);
                out << ' ' << num(g);
            }
        out << std::endl;
    }
}

std::vector<double> CorrelationAnalyzer::modelMultipoles(std::vector<likely::Parameters> const &points, int ndump, double zdump) const {
    if (ndump <= 1) throw RuntimeError("CorrelationAnalyzer::dump: expected ndump > 1.");
    if (!_model) throw RuntimeError("CorrelationAnalyzer::dumpModel: no model.");
    if (zdump < 0) {
        if (!_data) throw RuntimeError("CorrelationAnalyzer::dumpModel: no data to take the redshift from.");
        zdump = _data->getRedshift(_data->keptIndices().at(0));
    }
    // a multipole-binned point set (r_k, ell = 0,2,4, zdump) evaluated for all parameter vectors in one batch
    const int n = 3 * ndump, B = (int)points.size(), npar = _model->getNParameters();
    const double dr = (_rmax - _rmin) / (ndump - 1);
    std::vector<double> r(n), ell(n), z(n, zdump), zero(n, 0.0), eye((size_t)n * n, 0.0), pred((size_t)n * B), flat((size_t)B * npar);
    std::vector<int> idx(n), status(B, 0);
    for (int k = 0; k < ndump; ++k)
        for (int l = 0; l < 3; ++l) { r[3 * k + l] = _rmin + dr * k; ell[3 * k + l] = 2 * l; idx[3 * k + l] = -1; }
    for (int i = 0; i < n; ++i) eye[(size_t)i * n + i] = 1.0;
    for (int b = 0; b < B; ++b) {
        if ((int)points[b].size() != npar) throw RuntimeError("CorrelationAnalyzer::dumpModel: wrong number of parameters.");
        std::copy(points[b].begin(), points[b].end(), flat.begin() + (size_t)b * npar);
    }
    bf_data_desc dd;
    std::memset(&dd, 0, sizeof(dd));
    dd.n_data = n; dd.r = r.data(); dd.mu = ell.data(); dd.z = z.data(); dd.index = idx.data(); dd.data = zero.data();
    dd.cov = eye.data(); dd.binning_type = BF_BINNING_MULTIPOLE;
    bf_data *pts = nullptr;
    check(bf_data_create(deviceContext(), &dd, &pts), "bf_data_create");
    int rc = bf_eval_batch(deviceContext(), _model->deviceModel(), pts, flat.data(), B, pred.data(), B, status.data());
    bf_data_destroy(pts);
    check(rc, "bf_eval_batch");
    std::vector<double> out((size_t)B * n);  // [sample][3 * ndump]
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i) out[(size_t)b * n + i] = pred[(size_t)i * B + b];
    return out;
}

void CorrelationAnalyzer::dumpModel(std::ostream &out, likely::FitParameters parameters, int ndump, double zdump,
                                    std::string const &script, bool oneLine) const {
    if (!script.empty()) likely::modifyFitParameters(parameters, script);
    likely::Parameters values;
    for (auto const &p : parameters) values.push_back(p.value);
    std::vector<double> pred = modelMultipoles(std::vector<likely::Parameters>(1, values), ndump, zdump);
    const double dr = (_rmax - _rmin) / (ndump - 1);
    char buf[64];
    auto num = [&](double v) { std::snprintf(buf, sizeof(buf), "%.17g", v); return std::string(buf); };
    for (int k = 0; k < ndump; ++k) {
        if (!oneLine) out << num(_rmin + dr * k);
        out << ' ' << num(pred[3 * k]) << ' ' << num(pred[3 * k + 1]) << ' ' << num(pred[3 * k + 2]);
        if (!oneLine) out << std::endl;
    }
}

// SamplingOutput, CorrelationAnalyzer.cc:375-450: header "npar nsave nfit" (nfit = 2 with a refit pass), the errors of
// the input fit(s), then the input fit(s) and every sample as one line: parameter values, 2*fval, [refit values,
// 2*fval2,] and nsave x (mono, quad, hexa) of the sample's model(s) when nsave > 0.
void CorrelationAnalyzer::writeSamples(std::string const &path, likely::FunctionMinimumPtr fmin,
                                       std::vector<std::vector<double>> const &fitted, std::vector<double> const &chi2, int nsave,
                                       double zsave, likely::FunctionMinimumPtr fmin2, std::vector<std::vector<double>> const *refitted,
                                       std::vector<double> const *rechi2) const {
    if (nsave < 0) throw RuntimeError("CorrelationAnalyzer::doSamplingAnalysis: expected nsave >= 0.");
    if (fitted.size() != chi2.size()) throw RuntimeError("CorrelationAnalyzer::writeSamples: one chi-square per sample expected.");
    const bool refit = (bool)fmin2;
    if (refit != (refitted && rechi2) || (refit && (refitted->size() != fitted.size() || rechi2->size() != fitted.size())))
        throw RuntimeError("CorrelationAnalyzer::doSamplingAnalysis: inconsistent refit parameters.");
    std::ofstream out(path.c_str());
    if (!out.good()) throw RuntimeError("CorrelationAnalyzer: unable to open " + path);
    char buf[64];
    auto num = [&](double v) { std::snprintf(buf, sizeof(buf), "%.10g", v); return std::string(buf); };
    const int npar = (int)fmin->parameters.size();
    // model multipoles of the input fit(s) and of every sample, all in one GPU batch; failed fits (non-finite
    // parameters) dump the input model, their chi2 says so
    std::vector<double> dumps, dumps2;
    auto dumpAll = [&](likely::FunctionMinimumPtr fm, std::vector<std::vector<double>> const &f) {
        std::vector<likely::Parameters> pts;
        pts.push_back(fm->values());
        for (auto const &v : f) {
            bool finite = true;
            for (double x : v) finite &= std::isfinite(x);
            pts.push_back(finite ? v : fm->values());
        }
        return modelMultipoles(pts, nsave, zsave);
    };
    if (nsave > 0) {
        dumps = dumpAll(fmin, fitted);
        if (refit) dumps2 = dumpAll(fmin2, *refitted);
    }
    auto dump = [&](std::vector<double> const &d, size_t row) {
        for (int k = 0; k < 3 * nsave; ++k) out << ' ' << num(d[row * 3 * nsave + k]);
    };
    auto errors = [&](likely::FunctionMinimumPtr fm, bool first) {
        for (int i = 0; i < npar; ++i) out << (i || !first ? " " : "") << num(fm->parameters[i].floating ? fm->parameters[i].error : 0.0);
    };
    auto line = [&](std::vector<double> const &v, double c2, std::vector<double> const *v2, double c22, size_t row) {
        for (int i = 0; i < npar; ++i) out << num(v[i]) << " ";
        out << num(c2);
        if (v2) {
            for (int i = 0; i < npar; ++i) out << " " << num((*v2)[i]);
            out << " " << num(c22);
        }
        if (nsave > 0) {
            dump(dumps, row);
            if (v2) dump(dumps2, row);
        }
        out << std::endl;
    };
    out << npar << ' ' << nsave << ' ' << (refit ? 2 : 1) << std::endl;
    errors(fmin, true);
    if (refit) errors(fmin2, false);
    out << std::endl;
    std::vector<double> v0 = fmin->values(), v02;
    if (refit) v02 = fmin2->values();
    line(v0, 2 * fmin->minValue, refit ? &v02 : nullptr, refit ? 2 * fmin2->minValue : 0.0, 0);
    for (size_t k = 0; k < fitted.size(); ++k) line(fitted[k], chi2[k], refit ? &(*refitted)[k] : nullptr, refit ? (*rechi2)[k] : 0.0, k + 1);
}

void CorrelationAnalyzer::writeScan(std::string const &path, ScanResult const &scan) {
    // three header lines (npar ndump nfit / errors / best fit) then one line per sample: all
    // parameter values followed by 2*fval (CorrelationAnalyzer.cc:379-443); ndump = 0, nfit = 1
    std::ofstream out(path.c_str());
    if (!out.good()) throw RuntimeError("CorrelationAnalyzer: unable to open " + path);
    char buf[64];
    auto num = [&](double v) { std::snprintf(buf, sizeof(buf), "%.10g", v); return std::string(buf); };
    int npar = (int)scan.best->parameters.size();
    out << npar << " 0 1" << std::endl;
    for (int i = 0; i < npar; ++i) out << (i ? " " : "") << num(scan.best->parameters[i].floating ? scan.best->parameters[i].error : 0.0);
    out << std::endl;
    for (int i = 0; i < npar; ++i) out << num(scan.best->parameters[i].value) << " ";
    out << num(2 * scan.best->minValue) << std::endl;
    for (size_t k = 0; k < scan.points.size(); ++k) {
        for (int i = 0; i < npar; ++i) out << num(scan.points[k][i]) << " ";
        out << num(scan.chi2[k]) << std::endl;
    }
}

void CorrelationAnalyzer::setToyMCScale(double varianceScale) {
    if (!(varianceScale > 0)) throw RuntimeError("CorrelationAnalyzer::doMCSampling: expected varianceScale > 0.");
    _toyScale = varianceScale;
}

void CorrelationAnalyzer::doToyMCSampling(long first, long count, unsigned long long seed,
                                          std::vector<std::vector<double>> &fitted, std::vector<double> &chi2,
                                          std::vector<int> &status, std::string const &toymcConfig, std::string const &refitConfig,
                                          std::vector<std::vector<double>> *refitted, std::vector<double> *rechi2,
                                          std::vector<int> *restatus, likely::FitParameters const *truthBase,
                                          std::string const &saveFirstName) const {
    if (!_data || !_model) throw RuntimeError("CorrelationAnalyzer::doToyMCSampling: need data and a model.");
    if (!refitConfig.empty() && !(refitted && rechi2 && restatus))
        throw RuntimeError("CorrelationAnalyzer::doSamplingAnalysis: inconsistent refit parameters.");
    const bool verbose = std::getenv("BAOFIT_VERBOSE_FITS") != nullptr;
    double tStart = nowSeconds();
    CorrelationFitter fitter(_data, _model, _covSampleSize);
    // truth = model prediction at the (optionally reconfigured) parameters of the fit of the combined sample (:361-369)
    likely::FitParameters truthParams = truthBase ? *truthBase : _model->getFitParameters();
    if (!toymcConfig.empty()) likely::modifyFitParameters(truthParams, toymcConfig);
    likely::Parameters tp;
    for (auto const &p : truthParams) tp.push_back(p.value);
    std::vector<double> truth;
    fitter.getPrediction(tp, truth);
    bool grid = _model->description().dist_matrix != nullptr;
    // the realisations are drawn on the GPU and stay there; every evaluation point carries the
    // index of the toy it belongs to
    bf_toys *toys = nullptr;
    check(bf_toys_create(deviceContext(), _data->deviceData(grid), truth.data(), seed, first, (int)count, std::sqrt(_toyScale), &toys), "bf_toys_create");
    std::shared_ptr<bf_toys> guard(toys, [](bf_toys *t) { bf_toys_destroy(t); });
    if (!saveFirstName.empty() && first == 0 && count > 0) {
        // the first realisation of the study in saveData's format ("<index> <value>" over the kept bins)
        const int n = (int)_data->keptIndices().size();
        std::vector<double> table((size_t)count * n);
        check(bf_toys_download(deviceContext(), toys, table.data()), "bf_toys_download");
        std::ofstream out(saveFirstName.c_str());
        if (!out.good()) throw RuntimeError("CorrelationAnalyzer::doToyMCSampling: unable to open " + saveFirstName);
        char buf[64];
        for (int i = 0; i < n; ++i) {
            std::snprintf(buf, sizeof(buf), "%d %.17g", _data->keptIndices()[i], table[i]);
            out << buf << std::endl;
        }
    }
    double tToys = nowSeconds();
    if (_toyScale != 1) fitter.scaleCovariance(_toyScale);  // after the truth prediction; applies to the LM and Newton drivers alike
    const double cs = fitter.icovScale() / fitter.errorScale(), ps = 1.0 / fitter.errorScale();
    std::vector<int> toyOfFit(count);
    for (long t = 0; t < count; ++t) toyOfFit[t] = (int)t;
    // every toy is (re)fitted from the model's initial configuration with `config` applied (what
    // CorrelationFitter::fit(method, config) does, CorrelationAnalyzer.cc:480,494); all fits in lock step
    auto fitAll = [&](std::string const &config, std::vector<std::vector<double>> &outFitted, std::vector<double> &outChi2,
                      std::vector<int> &outStatus) {
        likely::FitParameters start = _model->getFitParameters();
        if (!config.empty()) likely::modifyFitParameters(start, config);
        std::vector<likely::LMFit> fits((size_t)count);
        {
            const likely::LMFit proto(start, cs, ps);
#pragma omp parallel for schedule(static) if (count >= 256)
            for (long t = 0; t < count; ++t) fits[t] = proto;
        }
        double t0 = nowSeconds();
        fitter.runLockStepLM(fits, toys, &toyOfFit);
        double tLM = nowSeconds();
        std::vector<likely::FunctionMinimumPtr> results(count);
        // toys whose LM fit did not converge cleanly are refitted with the Newton driver
        std::vector<likely::NewtonFit> redo;
        std::vector<int> redoToy;
#pragma omp parallel for schedule(static) if (count >= 256)
        for (long t = 0; t < count; ++t) results[t] = fits[t].result();
        for (long t = 0; t < count; ++t)
            if (results[t]->status != likely::FunctionMinimum::OK) { redo.push_back(likely::NewtonFit(start)); redoToy.push_back((int)t); }
        std::vector<likely::Parameters> points;
        std::vector<int> toyIndex;
        std::vector<double> values;
        while (!redo.empty()) {
            points.clear();
            toyIndex.clear();
            bool any = false;
            for (size_t q = 0; q < redo.size(); ++q) {
                if (redo[q].done()) continue;
                size_t before = points.size();
                redo[q].request(points);
                toyIndex.insert(toyIndex.end(), points.size() - before, redoToy[q]);
                any = true;
            }
            if (!any) break;
            fitter.evaluateBatchToys(points, values, toys, toyIndex);
            size_t off = 0;
            for (size_t q = 0; q < redo.size(); ++q) {
                if (redo[q].done()) continue;
                int np = redo[q].pending();
                redo[q].consume(values.data() + off);
                off += np;
            }
        }
        for (size_t q = 0; q < redo.size(); ++q) {
            likely::FunctionMinimumPtr fm = redo[q].result();
            if (fm->status != likely::FunctionMinimum::ERROR &&
                (results[redoToy[q]]->status == likely::FunctionMinimum::ERROR || fm->minValue <= results[redoToy[q]]->minValue))
                results[redoToy[q]] = fm;
        }
        if (verbose)
            std::fprintf(stderr, "doToyMCSampling: %ld toys%s: lock-step LM %.3f s, %zu Newton refits %.3f s\n", count,
                         config.empty() ? "" : " (refit)", tLM - t0, redo.size(), nowSeconds() - tLM);
        outFitted.clear(); outChi2.clear(); outStatus.clear();
        for (long t = 0; t < count; ++t) {
            likely::FunctionMinimumPtr fm = results[t];
            outFitted.push_back(fm->values());
            outChi2.push_back(2 * fm->minValue);
            outStatus.push_back((int)fm->status);
        }
    };
    if (verbose) std::fprintf(stderr, "doToyMCSampling: truth + %ld realisations %.3f s\n", count, tToys - tStart);
    fitAll("", fitted, chi2, status);
    if (!refitConfig.empty()) fitAll(refitConfig, *refitted, *rechi2, *restatus);
}

// ------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------
std::string Options::str(std::string const &k, std::string const &def) const {
    auto it = values.find(k);
    return it == values.end() ? def : it->second;
}
double Options::num(std::string const &k, double def) const {
    auto it = values.find(k);
    if (it == values.end()) return def;
    char *end = nullptr;
    double v = std::strtod(it->second.c_str(), &end);
    if (end == it->second.c_str()) throw RuntimeError("option " + k + " expects a number, got '" + it->second + "'");
    return v;
}

Options parseIni(std::string const &text) {
    static const char *switches[] = {"kspace", "kspace-fft", "kspace-hybrid", "anisotropic", "decoupled", "nl-broadband",
                                     "nl-correction", "fit-nl-correction", "nl-correction-alt", "distortion-alt",
                                     "no-distortion", "bin-smooth", "bin-smooth-alt", "hcd-model", "hcd-model-alt",
                                     "uvfluctuation", "radiation-model", "smooth-gauss", "smooth-lorentz", "dist-matrix",
                                     "metal-model", "metal-model-interpolate", "metal-civ", "toy-metal", "combined-bias",
                                     "combined-scale", "cross-correlation", "load-icov", "load-wdata", "custom-grid", "parameter-scan",
                                     "verbose", "quiet"};
    Options o;
    std::stringstream ss(text);
    std::string line;
    auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    while (std::getline(ss, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        size_t eq = line.find('=');
        std::string key = trim(eq == std::string::npos ? line : line.substr(0, eq));
        std::string val = eq == std::string::npos ? "" : trim(line.substr(eq + 1));
        bool isSwitch = false;
        for (const char *s : switches) isSwitch |= key == s;
        if (isSwitch) {
            // The reference declares these without a value and tests vm.count(name)
            // (src/baofit.cc:76,132-135,349-354): the mere presence of the key switches the option
            // on, whatever follows the '=' ("yes", "true;", even "no").
            o.flags.insert(key);
        } else if (key == "model-config") {
            o.modelConfig.push_back(val);  // composing, order preserved (src/baofit.cc:335-338)
        } else {
            o.values[key] = val;
        }
    }
    checkOptions(o);
    return o;
}

// Every key must be an option src/baofit.cc:46-300 declares (boost::program_options rejects anything else the same way); options
// that would select a model or a data treatment this build does not have are refused here instead of being ignored.
void checkOptions(Options const &o) {
    static const std::set<std::string> declared = {
        "abserr", "abserr-hybrid", "alt-config", "anisotropic", "axis1-bins", "axis2-bins", "axis3-bins", "bin-smooth", "bin-smooth-alt",
        "bootstrap-cov-trials", "bootstrap-size", "bootstrap-trials", "calculate-gradients", "check-posdef", "combined-bias", "combined-scale",
        "compare-each", "compare-each-final", "constrained-multipoles", "cov-sample-size", "cross-correlation", "cspline", "custom-grid", "data",
        "data-format", "decorrelated", "decoupled", "dilmax", "dilmin", "dist-add", "dist-matrix", "dist-matrix-dist-add", "dist-matrix-dist-mul",
        "dist-matrix-name", "dist-matrix-order", "dist-mul", "dist-r0", "distortion-alt", "dll", "dll2", "dsep", "dz", "dzmin", "ell-max", "fiducial",
        "fit-each", "fit-nl-correction", "fix-aln-cov", "fix-mode-scales", "gridscaling", "gridspacing", "hcd-model", "hcd-model-alt", "help",
        "hubble-constant", "ini-file", "jackknife-drop", "khi-spline", "klo-spline", "kspace", "kspace-fft", "kspace-hybrid", "kxmax", "ll-max",
        "ll-min", "lmax", "lmin", "load-icov", "load-wdata", "max-plates", "maxll", "mcmc-interval", "mcmc-save", "metal-civ", "metal-model",
        "metal-model-interpolate", "metal-model-name", "min-method", "minll", "minsep", "minz", "model-config", "modelroot", "mu-max", "mu-min",
        "n-spline", "naive-covariance", "ndump", "ngridx", "ngridy", "ngridz", "nl-broadband", "nl-correction", "nl-correction-alt",
        "no-distortion", "no-initial-fit", "nowiggles", "nsep", "nz", "omega-matter", "order-spline", "output-prefix", "parameter-scan",
        "platelist", "plateroot", "project-modes-keep", "quiet", "radiation-model", "random-seed", "refit-config", "relerr", "relerr-hybrid",
        "reuse-cov", "rmax", "rmin", "rpar-max", "rpar-min", "rperp-max", "rperp-min", "rveto-center", "rveto-width", "samples-per-decade",
        "save-data", "save-icov", "save-icov-scale", "scalar-weights", "sep-max", "sep-min", "sigma8", "smooth-gauss", "smooth-lorentz",
        "toy-metal", "toymc-config", "toymc-samples", "toymc-save", "toymc-scale", "uvfluctuation", "xi-method", "xi-points", "z-max", "z-min",
        "zcorr0", "zcorr1", "zcorr2", "zdump", "zeff", "zref",
        // not reference options: the radiation density of the quasar-coordinate cosmology, and the counterpart of "quiet"
        "omega-radiation", "verbose"};
    for (auto const &kv : o.values)
        if (!declared.count(kv.first)) throw RuntimeError("unrecognised option '" + kv.first + "'");
    for (auto const &f : o.flags)
        if (!declared.count(f)) throw RuntimeError("unrecognised option '" + f + "'");
    // src/baofit.cc:415-423,504,406: models and data treatments outside the accelerated path
    if (o.num("n-spline", 0) > 0) throw RuntimeError("n-spline > 0 selects PkCorrelationModel, which is not part of this build");
    if (!o.str("xi-points").empty()) throw RuntimeError("xi-points selects XiCorrelationModel, which is not part of this build");
    if (o.has("fix-aln-cov")) throw RuntimeError("fix-aln-cov (QuasarCorrelationData::fixCovariance) is not part of this build");
    if (o.has("scalar-weights")) throw RuntimeError("scalar-weights (BinnedDataResampler with scalar weights) is not part of this build");
}

static std::string resolve(std::string const &baseDir, std::string const &path) {
    if (path.empty() || path[0] == '/' || baseDir.empty()) return path;
    return baseDir + (baseDir.back() == '/' ? "" : "/") + path;
}

AbsCorrelationModelPtr createModel(Options const &opt, std::string const &baseDir) {
    // defaults of src/baofit.cc:58-178
    double OmegaMatter = opt.num("omega-matter", 0.27), zref = opt.num("zref", 2.25), distR0 = opt.num("dist-r0", 100);
    std::string modelroot = resolve(baseDir, opt.str("modelroot")), fid = opt.str("fiducial"), nw = opt.str("nowiggles");
    std::string distAdd = opt.str("dist-add"), distMul = opt.str("dist-mul");
    std::string metalName = opt.str("metal-model-name");
    if (!metalName.empty()) metalName = resolve(baseDir, metalName);
    AbsCorrelationModelPtr model;
    if (opt.flag("kspace-hybrid"))
        throw RuntimeError("the hybrid k-space model is not part of the accelerated path (DESIGN.md, out of scope)");
    if (opt.flag("kspace-fft")) {
        // src/baofit.cc:398-399,436-443
        int ngridx = (int)opt.num("ngridx", 400), ngridy = (int)opt.num("ngridy", 0), ngridz = (int)opt.num("ngridz", 0);
        if (ngridy == 0) ngridy = ngridx;
        if (ngridz == 0) ngridz = ngridy;
        model.reset(new BaoKSpaceFftCorrelationModel(
            modelroot, fid, nw, zref, OmegaMatter, opt.num("gridspacing", 4), ngridx, ngridy, ngridz, distAdd, distMul, distR0,
            opt.num("zcorr0", 0), opt.num("zcorr1", 0), opt.num("zcorr2", 0), opt.num("sigma8", 0.8338), opt.flag("anisotropic"),
            opt.flag("decoupled"), opt.flag("nl-broadband"), opt.flag("nl-correction"), opt.flag("fit-nl-correction"),
            opt.flag("nl-correction-alt"), opt.flag("distortion-alt"), opt.flag("no-distortion"), opt.flag("cross-correlation"),
            opt.flag("verbose"), opt.num("rmax", 200) * opt.num("dilmax", 1.5)));
    } else if (opt.flag("kspace")) {
        std::string dmName = opt.str("dist-matrix-name");
        if (dmName.empty() && opt.flag("dist-matrix")) dmName = opt.str("data");  // src/baofit.cc: data name by default
        if (!dmName.empty()) dmName = resolve(baseDir, dmName);
        model.reset(new BaoKSpaceCorrelationModel(
            modelroot, fid, nw, dmName, metalName, zref, OmegaMatter, opt.num("rmin", 0), opt.num("rmax", 200),
            opt.num("dilmin", 0.5), opt.num("dilmax", 1.5), opt.num("relerr", 1e-3), opt.num("abserr", 1e-5),
            (int)opt.num("ell-max", 4), (int)opt.num("samples-per-decade", 50), distAdd, distMul, distR0, opt.num("zeff", 0),
            opt.num("sigma8", 0.8338), opt.num("dzmin", 0.5), (int)opt.num("dist-matrix-order", 2500),
            opt.str("dist-matrix-dist-add"), opt.str("dist-matrix-dist-mul"), opt.flag("anisotropic"), opt.flag("decoupled"),
            opt.flag("nl-broadband"), opt.flag("nl-correction"), opt.flag("fit-nl-correction"), opt.flag("nl-correction-alt"),
            opt.flag("bin-smooth"), opt.flag("bin-smooth-alt"), opt.flag("hcd-model"), opt.flag("hcd-model-alt"),
            opt.flag("uvfluctuation"), opt.flag("radiation-model"), opt.flag("smooth-gauss"), opt.flag("smooth-lorentz"),
            opt.flag("dist-matrix"), opt.flag("metal-model"), opt.flag("metal-model-interpolate"), opt.flag("metal-civ"),
            opt.flag("toy-metal"), opt.flag("combined-bias"), opt.flag("combined-scale"), opt.flag("cross-correlation"),
            opt.flag("verbose")));
    } else {
        model.reset(new BaoCorrelationModel(modelroot, fid, nw, metalName, distAdd, distMul, distR0, zref, OmegaMatter,
                                            opt.flag("anisotropic"), opt.flag("decoupled"), opt.flag("metal-model"),
                                            opt.flag("metal-model-interpolate"), opt.flag("metal-civ"), opt.flag("toy-metal"),
                                            opt.flag("combined-bias"), opt.flag("combined-scale"), opt.flag("cross-correlation")));
    }
    for (auto const &script : opt.modelConfig) model->configureFitParameters(script);  // src/baofit.cc:462-464
    return model;
}

AbsCorrelationDataPtr createData(Options const &opt, std::string const &baseDir, BinnedDataResamplerPtr *resamplerOut) {
    std::string fmt = opt.str("data-format");
    BinnedGrid grid = createCorrelationGrid(opt.str("axis1-bins"), opt.str("axis2-bins"), opt.str("axis3-bins"));
    auto bare = [opt, fmt, grid]() -> AbsCorrelationDataPtr {
        if (fmt == "comoving-cartesian") return AbsCorrelationDataPtr(new ComovingCorrelationData(grid, ComovingCorrelationData::CartesianCoordinates));
        if (fmt == "comoving-polar") return AbsCorrelationDataPtr(new ComovingCorrelationData(grid, ComovingCorrelationData::PolarCoordinates));
        if (fmt == "comoving-multipole") return AbsCorrelationDataPtr(new ComovingCorrelationData(grid, ComovingCorrelationData::MultipoleCoordinates));
        if (fmt == "quasar") {
            // src/baofit.cc:409-414,495-500
            AbsHomogeneousUniversePtr cosmology(new LambdaCdmRadiationUniverse(opt.num("omega-matter", 0.27), 0, opt.num("hubble-constant", 0.7)));
            if (opt.has("omega-radiation")) cosmology->setOmegaRadiation(opt.num("omega-radiation", 0));  // not a reference option
            return AbsCorrelationDataPtr(new QuasarCorrelationData(grid, opt.num("ll-min", 0), opt.num("ll-max", 1), opt.num("sep-min", 0),
                                                                   opt.num("sep-max", 200), false, cosmology));
        }
        throw RuntimeError("data-format '" + fmt + "' is not supported by this build (comoving-cartesian, comoving-polar, comoving-multipole, quasar)");
    };
    // src/baofit.cc:510 with the defaults of :216-247
    auto prototype = [opt, bare]() -> AbsCorrelationDataPtr {
        AbsCorrelationDataPtr p = bare();
        double rVetoWidth = opt.num("rveto-width", 0), rVetoCenter = opt.num("rveto-center", 114);
        p->setFinalCuts(opt.num("rmin", 0), opt.num("rmax", 200), rVetoCenter - 0.5 * rVetoWidth, rVetoCenter + 0.5 * rVetoWidth,
                        opt.num("mu-min", -1), opt.num("mu-max", 1), opt.num("rperp-min", 0), opt.num("rperp-max", 200),
                        opt.num("rpar-min", -200), opt.num("rpar-max", 200), (int)opt.num("lmin", 0), (int)opt.num("lmax", 2),
                        opt.num("z-min", 0), opt.num("z-max", 10));
        return p;
    };
    // src/baofit.cc:514-524: one mode scale per line; applied to every data set right after it has been loaded (:570-572)
    std::vector<double> modeScales;
    if (!opt.str("fix-mode-scales").empty()) {
        std::string scalesName = resolve(baseDir, opt.str("fix-mode-scales"));
        std::ifstream in(scalesName.c_str());
        if (!in.good()) throw RuntimeError("Unable to open mode scales file " + scalesName);
        double scale;
        while (in >> scale) modeScales.push_back(scale);
        if (modeScales.empty()) throw RuntimeError("No mode scales found in " + scalesName);
    }
    AbsCorrelationDataPtr data;
    if (!opt.str("data").empty()) {
        data = prototype();
        loadCorrelationData(resolve(baseDir, opt.str("data")), *data, opt.flag("load-icov"), opt.flag("load-wdata"), opt.flag("custom-grid"));
        if (!modeScales.empty()) data->rescaleEigenvalues(modeScales);
    } else {
        // individual observations named by plateroot + platelist (src/baofit.cc:531-583), combined with
        // inverse-covariance weights like likely::BinnedDataResampler::combined
        std::string root = resolve(baseDir, opt.str("plateroot")), listName = root + opt.str("platelist");
        std::ifstream list(listName.c_str());
        if (!list.good()) throw RuntimeError("Unable to open platelist file " + listName);
        BinnedDataResamplerPtr resampler(new BinnedDataResampler(prototype));
        std::string plateName;
        int maxPlates = (int)opt.num("max-plates", 0);
        while (list >> plateName) {
            AbsCorrelationDataPtr plate = prototype();
            loadCorrelationData(root + plateName, *plate, opt.flag("load-icov"), opt.flag("load-wdata"), opt.flag("custom-grid"));
            if (!modeScales.empty()) plate->rescaleEigenvalues(modeScales);
            resampler->addObservation(plate);
            if (maxPlates > 0 && resampler->getNObservations() == maxPlates) break;
        }
        if (resampler->getNObservations() == 0) throw RuntimeError("combineCorrelationData: no observations.");
        data = resampler->combined();
        if (resamplerOut) *resamplerOut = resampler;
    }
    // src/baofit.cc:597-603: the combined data before final cuts are projected onto eigenmodes, then finalized
    const int projectModesNKeep = (int)opt.num("project-modes-keep", 0);
    if (projectModesNKeep != 0) data->projectOntoModes(projectModesNKeep);
    data->finalize();
    return data;
}

void CorrelationAnalyzer::compareEach(std::string const &saveName, bool finalized, std::vector<double> &prob, std::vector<double> &chi2,
                                      std::vector<double> &logDet) const {
    if (!_resampler || _resampler->getNObservations() <= 1) throw RuntimeError("CorrelationAnalyzer::compareEach: needs > 1 observation (plateroot/platelist).");
    std::ofstream out;
    if (!saveName.empty()) {
        out.open(saveName.c_str());
        if (!out.good()) throw RuntimeError("CorrelationAnalyzer::compareEach: unable to open " + saveName);
    }
    // the combined data as the reference (CorrelationAnalyzer.cc:84-87)
    AbsCorrelationDataPtr refData = _resampler->combined();
    if (finalized) refData->finalize();
    std::vector<double> dref, cref;
    dataAndCovariance(*refData, finalized, dref, cref);
    const int n = (int)dref.size(), nobs = _resampler->getNObservations();
    prob.assign(nobs, 0.0);
    chi2.assign(nobs, 0.0);
    logDet.assign(nobs, 0.0);
    char buf[64];
    for (int k = 0; k < nobs; ++k) {
        AbsCorrelationDataPtr observation = _resampler->getObservationCopy(k);
        if (finalized) observation->finalize();
        std::vector<double> dk, ck;
        dataAndCovariance(*observation, finalized, dk, ck);
        if ((int)dk.size() != n) throw RuntimeError("CorrelationAnalyzer::compareEach: observation is not congruent with the combined data.");
        // log|C_k| / n from the Cholesky factor
        std::vector<double> L = ck;
        if (!bf::cholesky_lower(n, L.data())) throw RuntimeError("CorrelationAnalyzer::compareEach: observation covariance is not positive definite.");
        double ld = 0;
        for (int i = 0; i < n; ++i) ld += 2.0 * std::log(L[(size_t)i * n + i]);
        logDet[k] = ld / n;
        // chi-square of (d_k - d_ref) with covariance C_k - C_ref, mode by mode (:102-115)
        for (size_t q = 0; q < ck.size(); ++q) ck[q] -= cref[q];
        std::vector<double> values, modes;
        symmetricEigenModes(n, ck, values, modes);
        std::vector<double> chi2modes(n);
        double sum = 0;
        for (int mI = 0; mI < n; ++mI) {
            const double *mode = &modes[(size_t)mI * n];
            double dot = 0;
            for (int i = 0; i < n; ++i) dot += mode[i] * (dk[i] - dref[i]);
            chi2modes[mI] = dot * dot / values[mI];
            sum += chi2modes[mI];
        }
        chi2[k] = sum;
        prob[k] = sum > 0 ? gammaQ(0.5 * n, 0.5 * sum) : 1.0;  // 1 - gamma_p(n/2, chi2/2), :117
        if (out.is_open()) {
            out << k << ' ' << logDet[k] << ' ' << chi2[k];
            for (int mI = 0; mI < n; ++mI) {
                std::snprintf(buf, sizeof(buf), " %.4lg", chi2modes[mI]);
                out << buf;
            }
            out << std::endl;
        }
    }
}

bool CorrelationAnalyzer::estimateCombinedCovariance(int nSamples, unsigned long long seed, std::string const &icovName, std::vector<int> &bins,
                                                     std::vector<double> &cov) const {
    if (!_resampler || _resampler->getNObservations() <= 1) throw RuntimeError("CorrelationAnalyzer::estimateCombinedCovariance: needs > 1 observation (plateroot/platelist).");
    if (nSamples < 2) throw RuntimeError("CorrelationAnalyzer::estimateCombinedCovariance: expected at least 2 samples.");
    // Welford accumulation of the combined data vector over the draws (likely::CovarianceAccumulator in the reference; the
    // unbiased 1/(N-1) normalisation is ours, the dependency is absent)
    std::vector<double> mean, m2, d, c;
    int n = 0;
    for (long trial = 0; trial < nSamples; ++trial) {
        AbsCorrelationDataPtr sample = _resampler->bootstrap(0, false, seed, trial);
        dataAndCovariance(*sample, false, d, c);
        if (trial == 0) {
            bins = sample->binsWithData();
            n = (int)d.size();
            mean.assign(n, 0.0);
            m2.assign((size_t)n * n, 0.0);
        }
        std::vector<double> delta(n);
        for (int i = 0; i < n; ++i) { delta[i] = d[i] - mean[i]; mean[i] += delta[i] / (trial + 1); }
        for (int i = 0; i < n; ++i) {
            const double di = delta[i];
            for (int j = 0; j <= i; ++j) m2[(size_t)i * n + j] += di * (d[j] - mean[j]);
        }
    }
    cov.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) cov[(size_t)i * n + j] = cov[(size_t)j * n + i] = m2[(size_t)i * n + j] / (nSamples - 1);
    std::vector<double> L = cov;
    const bool posdef = bf::cholesky_lower(n, L.data());
    if (posdef && !icovName.empty()) {
        std::vector<double> li((size_t)n * n, 0.0);
        bf::invert_lower(n, L.data(), li.data());
        std::ofstream out(icovName.c_str());
        if (!out.good()) throw RuntimeError("CorrelationAnalyzer::estimateCombinedCovariance: unable to open " + icovName);
        char buf[96];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                double sum = 0;
                for (int k = i; k < n; ++k) sum += li[(size_t)k * n + i] * li[(size_t)k * n + j];  // (L^-T L^-1)_ij
                std::snprintf(buf, sizeof(buf), "%d %d %.17g", bins[i], bins[j], sum);
                out << buf << std::endl;
            }
    }
    return posdef;
}

// ---- resampling analyses ----------------------------------------------------------------------
long CorrelationAnalyzer::resamplingSize(std::string const &kind, int n) const {
    if (!_resampler || _resampler->getNObservations() <= 1) throw RuntimeError("CorrelationAnalyzer: resampling needs > 1 observation (plateroot/platelist).");
    if (kind == "jackknife") {
        if (n <= 0) throw RuntimeError("CorrelationAnalyzer::doJackknifeAnalysis: expected jackknifeDrop > 0.");
        return BinnedDataResampler::jackknifeSize(_resampler->getNObservations(), n);
    }
    if (kind == "bootstrap") {
        if (n <= 0) throw RuntimeError("CorrelationAnalyzer::doBootstrapAnalysis: expected bootstrapTrials > 0.");
        return n;
    }
    if (kind == "each") return _resampler->getNObservations();
    throw RuntimeError("CorrelationAnalyzer: unknown resampling kind '" + kind + "' (jackknife, bootstrap, each).");
}

AbsCorrelationDataPtr CorrelationAnalyzer::resample(std::string const &kind, int n, int size, bool fixCovariance, unsigned long long seed,
                                                    long seqno) const {
    long total = resamplingSize(kind, n);
    if (seqno < 0 || seqno >= total) throw RuntimeError("CorrelationAnalyzer: resampling sequence number out of range.");
    if (size < 0) throw RuntimeError("CorrelationAnalyzer::doBootstrapAnalysis: expected bootstrapSize >= 0.");
    AbsCorrelationDataPtr sample;
    if (kind == "jackknife") sample = _resampler->jackknife(n, seqno);  // non-null: seqno < resamplingSize
    else if (kind == "bootstrap") sample = _resampler->bootstrap(size, fixCovariance, seed, seqno);
    else sample = _resampler->getObservationCopy((int)seqno);
    sample->finalize();
    return sample;
}

long CorrelationAnalyzer::doResampling(std::string const &kind, int n, int size, bool fixCovariance, unsigned long long seed, long first,
                                       long count, std::vector<std::vector<double>> &fitted, std::vector<double> &chi2,
                                       std::vector<int> &status, std::string const &refitConfig,
                                       std::vector<std::vector<double>> *refitted, std::vector<double> *rechi2,
                                       std::vector<int> *restatus) const {
    if (!refitConfig.empty() && !(refitted && rechi2 && restatus))
        throw RuntimeError("CorrelationAnalyzer::doSamplingAnalysis: inconsistent refit parameters.");
    long total = resamplingSize(kind, n);
    if (count < 0) count = total - first;
    if (first < 0 || first + count > total) throw RuntimeError("CorrelationAnalyzer: bad resampling range.");
    fitted.clear(); chi2.clear(); status.clear();
    if (refitted) { refitted->clear(); rechi2->clear(); restatus->clear(); }
    // a failed fit is recorded, not fatal (CorrelationAnalyzer.cc:486-489 swallows per-fit errors)
    auto fitOne = [&](AbsCorrelationDataPtr sample, std::string const &config, std::vector<std::vector<double>> &f, std::vector<double> &c,
                      std::vector<int> &st) {
        try {
            likely::FunctionMinimumPtr fm = fitSample(sample, config);
            f.push_back(fm->values());
            c.push_back(2 * fm->minValue);
            st.push_back((int)fm->status);
        } catch (std::runtime_error const &e) {
            f.push_back(std::vector<double>(_model->getNParameters(), std::nan("")));
            c.push_back(std::nan(""));
            st.push_back((int)likely::FunctionMinimum::ERROR);
        }
    };
    for (long q = first; q < first + count; ++q) {
        AbsCorrelationDataPtr sample = resample(kind, n, size, fixCovariance, seed, q);
        fitOne(sample, "", fitted, chi2, status);
        if (!refitConfig.empty()) fitOne(sample, refitConfig, *refitted, *rechi2, *restatus);
    }
    return count;
}

}  // namespace baofit
