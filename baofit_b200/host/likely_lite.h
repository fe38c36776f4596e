// Minimal stand-ins for the pieces of the un-vendored `likely` library that the reference's hot
// path leans on (fit parameters and their configuration scripts, binning strings, the function
// minimum record, and a minimiser). Names follow likely's so that the baofit classes above read
// like the reference. Behaviour recalled from likely's public sources is marked UNVERIFIED (dep).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace baofit {
class RuntimeError : public std::runtime_error {  // baofit/RuntimeError.h:10-20
public:
    explicit RuntimeError(std::string const &reason) : std::runtime_error(reason) {}
};
}  // namespace baofit

namespace likely {

typedef std::vector<double> Parameters;

class RuntimeError : public std::runtime_error {
public:
    explicit RuntimeError(std::string const &reason) : std::runtime_error(reason) {}
};

// One named fit parameter: value, step/error estimate, floating flag, optional prior, optional
// scan binning (binning[name]={lo:hi}*n).
struct FitParameter {
    std::string name;
    double value = 0, error = 0;
    bool floating = true;
    enum PriorType { NoPrior, GaussPrior, BoxPrior } priorType = NoPrior;
    double priorMin = 0, priorMax = 0, priorScale = 1;
    std::vector<double> binning;  // scan grid values (empty: not scanned)
};
typedef std::vector<FitParameter> FitParameters;

// Applies a configuration script (likely's mini-language, e.g. config/BOSSDR9LyaF.ini:42-85):
//   value[pat]=v ; fix[pat]=v ; fix[pat] ; release[pat] ; error[pat]=v ;
//   binning[pat]={lo:hi}*n | [lo:hi]*n | {a,b,c} ; gaussprior[pat] @ (lo,hi) ; boxprior[pat] @ (lo,hi)
// with '*' globs in pat. Throws RuntimeError on syntax errors or patterns that match nothing.
void modifyFitParameters(FitParameters &params, std::string const &script);

// likely::createBinning strings -> bin centres:  [lo:hi]*n (n bins, centres), {lo:hi}*n (n samples
// including both ends), {a,b,c} (explicit centres). Conventions pinned by demo/rmu.theory and
// demo/multi.theory.
std::vector<double> createBinning(std::string const &spec);
// same strings -> bin widths: (hi-lo)/n for "[lo:hi]*n" bins, 0 for samplings ({..} forms, which
// describe points rather than intervals; likely::AbsBinning::getBinWidth UNVERIFIED (dep))
std::vector<double> createBinningWidths(std::string const &spec);

// Gaussian prior: centre (lo+hi)/2, sigma (hi-lo)/2 ("range is -/+ 1 sigma",
// config/BOSSDR9LyaF.ini:81). Box prior: flat inside, quadratic wall outside (exact form of the
// dependency UNVERIFIED). Returns the -log(prior) sum over floating parameters.
double evaluatePriors(FitParameters const &params, Parameters const &values);

struct FunctionMinimum {
    enum Status { OK, WARNING, ERROR } status = ERROR;
    FitParameters parameters;     // values = best fit, errors = sqrt(diag(cov)) for floating ones
    double minValue = 0;          // -log L at the minimum
    std::vector<double> covariance;  // nfloat x nfloat, row-major (empty if HESSE failed)
    int nEvaluations = 0, nIterations = 0;
    double edm = 0;
    Parameters values() const {
        Parameters v;
        for (auto const &p : parameters) v.push_back(p.value);
        return v;
    }
};
typedef std::shared_ptr<FunctionMinimum> FunctionMinimumPtr;

// Objective evaluated for many parameter vectors at once: f[b] = F(points[b]) (full-length vectors).
typedef std::function<void(std::vector<Parameters> const &points, std::vector<double> &f)> BatchFunction;

// One independent minimisation expressed as a request/consume state machine so that many fits
// can run in lock step and share one batched evaluation per round (scan points, toy samples).
// Algorithm: damped Newton on a finite-difference gradient + Hessian (one batch of
// 1 + 2n + n(n-1)/2 points per iteration), then a batch of trial step lengths; final HESSE by
// central mixed differences. It replaces Minuit2's VariableMetric ("mn2::vmetric"), whose
// iteration path cannot be reproduced (Minuit2 is not in the tree).
class NewtonFit {
public:
    NewtonFit(FitParameters const &params, double errorDef = 0.5);
    bool done() const { return _stage == Done; }
    void request(std::vector<Parameters> &out);  // appends the points this fit wants next
    void consume(const double *f);               // values of the requested points, in order
    int pending() const { return _npending; }
    FunctionMinimumPtr result() const;
    double tolerance = 1e-6;  // EDM goal in units of -log L
    int maxIterations = 60;

private:
    enum Stage { Derivs, LineSearch, Hesse, Done } _stage;
    FitParameters _params;
    std::vector<int> _fidx;  // indices of floating parameters
    int _n, _npending = 0, _iter = 0, _neval = 0, _stall = 0;
    double _up, _f0 = 0, _edm = 0, _lambda = 0;
    bool _haveF0 = false, _hesseOk = false, _failed = false;
    Parameters _x;               // current full vector
    std::vector<double> _h;      // probe steps (floating)
    std::vector<double> _g, _H;  // gradient, Hessian (floating)
    std::vector<double> _step;   // Newton step (floating)
    std::vector<double> _alphas;
    std::vector<double> _cov;
    void solveStep();
    Parameters shifted(std::vector<std::pair<int, double>> const &moves) const;
};

// Runs a set of fits in lock step: gathers every unfinished fit's requests into one batch.
void runLockStep(std::vector<NewtonFit> &fits, BatchFunction const &f);

// -log(prior) of one parameter and its first / second derivative (same forms as evaluatePriors)
void priorDerivatives(FitParameter const &p, double x, double &value, double &grad, double &hess);

// Levenberg-Marquardt on the Gauss-Newton normal equations, request/consume form, for objectives
//   F(x) = chi2Scale * 0.5 |y(x)|^2 + priorScale * priors(x)
// whose residual Gram matrices come from the GPU (bf_gram_batch): one Jacobian round needs the centre
// and the 2n central probes x +- h_i (2n+1 points instead of the 1 + 2n + n(n-1)/2 of NewtonFit), one
// trial round three step lengths. The central probes also give the diagonal of the second-order
// term y.d2y/dx_i^2 that Gauss-Newton drops; it is added to the normal matrix, which matters where the
// model is quadratic in a parameter (bias -> 0). Used for the many independent refits of scans and toy studies, where the
// covariance of each fit is not needed; single fits keep NewtonFit (accurate HESSE errors).
class LMFit {
public:
    LMFit(FitParameters const &params, double chi2Scale = 1.0, double priorScale = 1.0);
    LMFit() : _stage(Done), _n(0), _cs(1.0), _ps(1.0) {}  // placeholder to be assigned (fits of a scan are built on all host threads)
    bool done() const { return _stage == Done; }
    bool wantsJacobian() const { return _stage == Jacobian; }
    int groupSize() const { return 2 * _n + 1; }
    int nPar() const { return (int)_x.size(); }
    // Jacobian round: appends groupSize() full-length vectors (centre first) to flat
    void requestJacobian(std::vector<double> &flat) const;
    // the same round described compactly: the centre and the probe step of every parameter (0 = not
    // probed), npar numbers each; the probe columns are generated on the GPU (bf_gram_probes)
    void requestProbes(std::vector<double> &centres, std::vector<double> &steps) const;
    int nFloating() const { return _n; }
    // gram[G*G] as returned by bf_gram_batch; bad = some column had a non-zero status
    void consumeJacobian(const double *gram, bool bad);
    // the same round fed with the device-reduced normal equations of bf_normal_probes: (n+1)^2 + n values
    void consumeNormal(const double *normal, bool bad);
    // gives the fit up (status ERROR) without evaluating anything: the caller's fallback driver takes over
    void abandon() { _failed = true; _stage = Done; }
    int normalSize() const { return (_n + 1) * (_n + 1) + _n; }
    // trial round: appends pending() full-length vectors
    void requestPoints(std::vector<double> &flat);
    int pending() const { return _npending; }
    // fval[k] = chi2 of point k (NaN if invalid); priors are added here
    void consumePoints(const double *chi2);
    FunctionMinimumPtr result() const;
    double tolerance = 1e-6;
    int maxIterations = 40;  // a fit that needs more is handed to NewtonFit by the callers
    bool debug = false;      // trace iterations on stderr

private:
    enum Stage { Jacobian, Trial, Done } _stage;
    FitParameters _params;
    std::vector<int> _fidx;
    int _n, _npending = 0, _iter = 0, _neval = 0, _stall = 0, _shrinks = 0;
    double _cs, _ps, _f0 = 0, _edm = 0, _lambda = 1e-3, _radius = 5.0;
    bool _capped = false;
    int _smallGains = 0;
    bool _failed = false, _converged = false;
    Parameters _x;
    std::vector<double> _h, _g, _A, _step, _alphas, _scale;  // _scale: natural size of each floating parameter (its error estimate)
    std::vector<char> _active;  // floating parameters held on a wall of their box prior in this iteration
    double priorValue(Parameters const &x) const;
    void solveStep();
};

}  // namespace likely
