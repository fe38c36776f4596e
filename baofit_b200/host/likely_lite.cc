#include "likely_lite.h"

#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <sstream>

namespace likely {

namespace {
std::string trim(std::string const &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

bool globMatch(const char *pat, const char *str) {
    if (!*pat) return !*str;
    if (*pat == '*') {
        for (const char *s = str;; ++s) {
            if (globMatch(pat + 1, s)) return true;
            if (!*s) return false;
        }
    }
    return *str && *pat == *str && globMatch(pat + 1, str + 1);
}

double toDouble(std::string const &s, std::string const &context) {
    std::string t = trim(s);
    char *end = nullptr;
    double v = std::strtod(t.c_str(), &end);
    if (t.empty() || end == t.c_str() || *end) throw RuntimeError("cannot parse number '" + t + "' in " + context);
    return v;
}

void parseRange(std::string const &arg, double &lo, double &hi, std::string const &stmt) {
    // "@ (lo,hi)" possibly with ";scale" inside the parentheses
    size_t a = arg.find('('), b = arg.find(')');
    if (arg.find('@') == std::string::npos || a == std::string::npos || b == std::string::npos || b < a)
        throw RuntimeError("badly formatted prior range in: " + stmt);
    std::string in = arg.substr(a + 1, b - a - 1);
    size_t c = in.find(',');
    if (c == std::string::npos) throw RuntimeError("badly formatted prior range in: " + stmt);
    lo = toDouble(in.substr(0, c), stmt);
    hi = toDouble(in.substr(c + 1), stmt);
    if (!(lo < hi)) throw RuntimeError("prior range must have lo < hi in: " + stmt);
}
}  // namespace

std::vector<double> createBinning(std::string const &specIn) {
    std::string s = trim(specIn);
    if (s.size() < 2) throw RuntimeError("createBinning: cannot parse '" + specIn + "'");
    char open = s[0];
    size_t close = s.find(open == '[' ? ']' : '}');
    if ((open != '[' && open != '{') || close == std::string::npos) throw RuntimeError("createBinning: cannot parse '" + specIn + "'");
    std::string body = s.substr(1, close - 1), tail = trim(s.substr(close + 1));
    std::vector<double> out;
    size_t colon = body.find(':');
    if (colon != std::string::npos) {
        if (tail.empty() || tail[0] != '*') throw RuntimeError("createBinning: expected *n in '" + specIn + "'");
        double lo = toDouble(body.substr(0, colon), specIn), hi = toDouble(body.substr(colon + 1), specIn);
        int n = (int)toDouble(tail.substr(1), specIn);
        if (n <= 0) throw RuntimeError("createBinning: expected n > 0 in '" + specIn + "'");
        if (open == '[') {
            double w = (hi - lo) / n;
            for (int i = 0; i < n; ++i) out.push_back(lo + (i + 0.5) * w);
        } else {
            for (int i = 0; i < n; ++i) out.push_back(n > 1 ? lo + (hi - lo) * i / (n - 1) : lo);
        }
        return out;
    }
    if (open != '{' || !tail.empty()) throw RuntimeError("createBinning: cannot parse '" + specIn + "'");
    std::stringstream ss(body);
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back(toDouble(tok, specIn));
    if (out.empty()) throw RuntimeError("createBinning: no bins in '" + specIn + "'");
    return out;
}

std::vector<double> createBinningWidths(std::string const &specIn) {
    std::vector<double> centres = createBinning(specIn);
    std::string s = trim(specIn);
    std::vector<double> w(centres.size(), 0.0);
    size_t colon = s.find(':'), close = s.find(']');
    if (s[0] == '[' && colon != std::string::npos && close != std::string::npos) {
        double lo = toDouble(s.substr(1, colon - 1), specIn), hi = toDouble(s.substr(colon + 1, close - colon - 1), specIn);
        for (auto &x : w) x = (hi - lo) / (double)centres.size();
    }
    return w;
}

void modifyFitParameters(FitParameters &params, std::string const &script) {
    std::stringstream ss(script);
    std::string stmt;
    while (std::getline(ss, stmt, ';')) {
        stmt = trim(stmt);
        if (stmt.empty()) continue;
        size_t lb = stmt.find('['), rb = stmt.find(']');
        if (lb == std::string::npos || rb == std::string::npos || rb < lb)
            throw RuntimeError("badly formatted parameter script: " + stmt);
        std::string op = trim(stmt.substr(0, lb)), pat = stmt.substr(lb + 1, rb - lb - 1), arg = trim(stmt.substr(rb + 1));
        bool haveValue = false;
        double value = 0;
        std::vector<double> bins;
        double lo = 0, hi = 0;
        if (op == "value" || op == "fix" || op == "error" || op == "release") {
            if (!arg.empty()) {
                if (arg[0] != '=') throw RuntimeError("expected '=' in: " + stmt);
                value = toDouble(arg.substr(1), stmt);
                haveValue = true;
            }
            if ((op == "value" || op == "error") && !haveValue) throw RuntimeError("missing value in: " + stmt);
        } else if (op == "binning") {
            if (arg.empty() || arg[0] != '=') throw RuntimeError("expected '=' in: " + stmt);
            bins = createBinning(arg.substr(1));
        } else if (op == "gaussprior" || op == "boxprior") {
            parseRange(arg, lo, hi, stmt);
        } else if (op == "noprior") {
        } else {
            throw RuntimeError("unknown parameter script command: " + stmt);
        }
        int matched = 0;
        for (auto &p : params) {
            if (!globMatch(pat.c_str(), p.name.c_str())) continue;
            ++matched;
            if (op == "value") p.value = value;
            else if (op == "fix") { p.floating = false; if (haveValue) p.value = value; }
            else if (op == "release") { p.floating = true; if (haveValue) p.value = value; }
            else if (op == "error") p.error = value;
            else if (op == "binning") p.binning = bins;
            else if (op == "gaussprior") { p.priorType = FitParameter::GaussPrior; p.priorMin = lo; p.priorMax = hi; }
            else if (op == "boxprior") { p.priorType = FitParameter::BoxPrior; p.priorMin = lo; p.priorMax = hi; }
            else if (op == "noprior") p.priorType = FitParameter::NoPrior;
        }
        if (!matched) throw RuntimeError("no parameter matches '" + pat + "' in: " + stmt);
    }
}

double evaluatePriors(FitParameters const &params, Parameters const &values) {
    double sum = 0;
    for (size_t i = 0; i < params.size(); ++i) {
        FitParameter const &p = params[i];
        if (!p.floating || p.priorType == FitParameter::NoPrior) continue;
        double c = 0.5 * (p.priorMin + p.priorMax), s = 0.5 * (p.priorMax - p.priorMin), x = values[i];
        if (p.priorType == FitParameter::GaussPrior) {
            double t = (x - c) / s;
            sum += 0.5 * t * t;
        } else {
            double d = x < p.priorMin ? p.priorMin - x : (x > p.priorMax ? x - p.priorMax : 0.0);
            if (d > 0) {
                double t = d / (1e-3 * s);
                sum += 0.5 * t * t;
            }
        }
    }
    return sum;
}

// ------------------------------------------------------------------------------------------
// damped Newton with finite-difference derivatives, request/consume form
// ------------------------------------------------------------------------------------------
namespace {
// Cholesky solve of (A) x = b for symmetric positive definite A (n x n, row-major). Returns false
// if A is not positive definite. If inv != nullptr also returns A^-1.
bool cholSolve(int n, std::vector<double> const &A, std::vector<double> const &b, std::vector<double> &x,
               std::vector<double> *inv = nullptr) {
    std::vector<double> L(A);
    for (int j = 0; j < n; ++j) {
        double s = L[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        if (!(s > 0) || !std::isfinite(s)) return false;
        double d = std::sqrt(s);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double t = L[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / d;
        }
    }
    auto solve = [&](std::vector<double> const &rhs, std::vector<double> &out) {
        out = rhs;
        for (int i = 0; i < n; ++i) {
            double t = out[i];
            for (int k = 0; k < i; ++k) t -= L[i * n + k] * out[k];
            out[i] = t / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double t = out[i];
            for (int k = i + 1; k < n; ++k) t -= L[k * n + i] * out[k];
            out[i] = t / L[i * n + i];
        }
    };
    solve(b, x);
    if (inv) {
        inv->assign((size_t)n * n, 0.0);
        std::vector<double> e(n), col;
        for (int j = 0; j < n; ++j) {
            std::fill(e.begin(), e.end(), 0.0);
            e[j] = 1.0;
            solve(e, col);
            for (int i = 0; i < n; ++i) (*inv)[i * n + j] = col[i];
        }
    }
    return true;
}
}  // namespace

NewtonFit::NewtonFit(FitParameters const &params, double errorDef) : _stage(Derivs), _params(params), _up(errorDef) {
    for (size_t i = 0; i < params.size(); ++i) {
        _x.push_back(params[i].value);
        if (params[i].floating) {
            _fidx.push_back((int)i);
            double e = params[i].error > 0 ? params[i].error : std::max(1e-3, 0.1 * std::fabs(params[i].value));
            _h.push_back(0.1 * e);
        }
    }
    _n = (int)_fidx.size();
    _g.assign(_n, 0.0);
    _H.assign((size_t)_n * _n, 0.0);
    _step.assign(_n, 0.0);
    if (_n == 0) _stage = Hesse;  // nothing to minimise: just evaluate once
}

Parameters NewtonFit::shifted(std::vector<std::pair<int, double>> const &moves) const {
    Parameters p(_x);
    for (auto const &m : moves) p[_fidx[m.first]] += m.second;
    return p;
}

void NewtonFit::request(std::vector<Parameters> &out) {
    _npending = 0;
    auto push = [&](Parameters const &p) { out.push_back(p); ++_npending; };
    if (_stage == Derivs) {
        push(_x);
        for (int i = 0; i < _n; ++i) { push(shifted({{i, _h[i]}})); push(shifted({{i, -_h[i]}})); }
        for (int i = 0; i < _n; ++i)
            for (int j = i + 1; j < _n; ++j) push(shifted({{i, _h[i]}, {j, _h[j]}}));
    } else if (_stage == LineSearch) {
        for (double a : _alphas) {
            Parameters p(_x);
            for (int i = 0; i < _n; ++i) p[_fidx[i]] += a * _step[i];
            push(p);
        }
    } else if (_stage == Hesse) {
        push(_x);
        for (int i = 0; i < _n; ++i) { push(shifted({{i, _h[i]}})); push(shifted({{i, -_h[i]}})); }
        for (int i = 0; i < _n; ++i)
            for (int j = i + 1; j < _n; ++j) {
                push(shifted({{i, _h[i]}, {j, _h[j]}}));
                push(shifted({{i, _h[i]}, {j, -_h[j]}}));
                push(shifted({{i, -_h[i]}, {j, _h[j]}}));
                push(shifted({{i, -_h[i]}, {j, -_h[j]}}));
            }
    }
}

void NewtonFit::solveStep() {
    // (H + lambda D) step = -g with D = diag(|H_ii|); lambda grows until positive definite
    std::vector<double> rhs(_n), A, inv;
    for (int i = 0; i < _n; ++i) rhs[i] = -_g[i];
    for (int attempt = 0; attempt < 40; ++attempt) {
        A = _H;
        for (int i = 0; i < _n; ++i) {
            double d = std::fabs(_H[i * _n + i]);
            if (!(d > 0)) d = 1.0 / (_h[i] * _h[i]);
            A[i * _n + i] += _lambda * d;
        }
        if (cholSolve(_n, A, rhs, _step, &inv)) {
            double edm = 0;
            for (int i = 0; i < _n; ++i) edm -= 0.5 * _g[i] * _step[i];
            _edm = edm;
            return;
        }
        _lambda = _lambda > 0 ? _lambda * 10 : 1e-3;
    }
    // hopeless Hessian: steepest descent scaled by the probe steps
    for (int i = 0; i < _n; ++i) _step[i] = -_g[i] * _h[i] * _h[i];
    _edm = std::numeric_limits<double>::infinity();
}

void NewtonFit::consume(const double *f) {
    _neval += _npending;
    const double inf = std::numeric_limits<double>::infinity();
    auto val = [&](int k) { return std::isfinite(f[k]) ? f[k] : inf; };
    if (_stage == Derivs) {
        int k = 0;
        double f0 = val(k++);
        bool bad = !std::isfinite(f0);
        std::vector<double> fp(_n), fm(_n);
        for (int i = 0; i < _n; ++i) { fp[i] = val(k++); fm[i] = val(k++); bad |= !std::isfinite(fp[i]) || !std::isfinite(fm[i]); }
        std::vector<double> fpp((size_t)_n * _n, 0.0);
        for (int i = 0; i < _n; ++i)
            for (int j = i + 1; j < _n; ++j) { fpp[i * _n + j] = val(k++); bad |= !std::isfinite(fpp[i * _n + j]); }
        if (bad) {
            // a probe left the valid domain (e.g. dilation limits): shrink the probe steps and retry
            if (!std::isfinite(f0) || ++_stall > 8) { _failed = true; _f0 = f0; _stage = Done; return; }
            for (auto &h : _h) h *= 0.25;
            return;
        }
        _f0 = f0;
        _haveF0 = true;
        for (int i = 0; i < _n; ++i) {
            _g[i] = (fp[i] - fm[i]) / (2 * _h[i]);
            _H[i * _n + i] = (fp[i] + fm[i] - 2 * f0) / (_h[i] * _h[i]);
        }
        for (int i = 0; i < _n; ++i)
            for (int j = i + 1; j < _n; ++j) {
                double hij = (fpp[i * _n + j] - fp[i] - fp[j] + f0) / (_h[i] * _h[j]);
                _H[i * _n + j] = _H[j * _n + i] = hij;
            }
        solveStep();
        ++_iter;
        if ((_edm < tolerance && _lambda < 1e-6) || _iter >= maxIterations) {
            _stage = Hesse;
            return;
        }
        _alphas = {1.0, 0.5, 0.25, 0.1, 0.03, 2.0, 4.0};
        _stage = LineSearch;
    } else if (_stage == LineSearch) {
        int best = -1;
        double fbest = _f0;
        for (size_t k = 0; k < _alphas.size(); ++k)
            if (val((int)k) < fbest) { fbest = val((int)k); best = (int)k; }
        if (best >= 0) {
            double a = _alphas[best];
            for (int i = 0; i < _n; ++i) _x[_fidx[i]] += a * _step[i];
            double gain = _f0 - fbest;
            _f0 = fbest;
            _stall = 0;
            // full steps that work relax the damping, short ones keep it
            _lambda = a >= 1.0 ? _lambda * 0.1 : _lambda;
            if (_lambda < 1e-9) _lambda = 0;
            // probe steps from the curvature: delta f ~ 1e-3 * errorDef along each axis
            for (int i = 0; i < _n; ++i) {
                double hii = _H[i * _n + i];
                if (hii > 0) {
                    double h = std::sqrt(2 * 1e-3 * _up / hii);
                    _h[i] = std::min(std::max(h, 1e-8 * std::max(1.0, std::fabs(_x[_fidx[i]]))), 10 * _h[i]);
                }
            }
            if (gain < 0.1 * tolerance && _edm < 10 * tolerance) { _stage = Hesse; return; }
            _stage = Derivs;
        } else {
            // no trial improved: more damping, same derivatives
            if (++_stall > 12) { _stage = Hesse; return; }
            _lambda = _lambda > 0 ? _lambda * 10 : 1e-2;
            solveStep();
            _alphas = {1.0, 0.5, 0.25, 0.1, 0.03, 0.01, 0.003};
        }
    } else if (_stage == Hesse) {
        int k = 0;
        double f0 = val(k++);
        _f0 = f0;
        if (_n == 0) { _hesseOk = std::isfinite(f0); _failed = !_hesseOk; _stage = Done; return; }
        std::vector<double> fp(_n), fm(_n), H((size_t)_n * _n, 0.0), g(_n);
        bool bad = !std::isfinite(f0);
        for (int i = 0; i < _n; ++i) {
            fp[i] = val(k++); fm[i] = val(k++);
            bad |= !std::isfinite(fp[i]) || !std::isfinite(fm[i]);
            g[i] = (fp[i] - fm[i]) / (2 * _h[i]);
            H[i * _n + i] = (fp[i] + fm[i] - 2 * f0) / (_h[i] * _h[i]);
        }
        for (int i = 0; i < _n; ++i)
            for (int j = i + 1; j < _n; ++j) {
                double a = val(k++), b = val(k++), c = val(k++), d = val(k++);
                bad |= !std::isfinite(a + b + c + d);
                H[i * _n + j] = H[j * _n + i] = (a - b - c + d) / (4 * _h[i] * _h[j]);
            }
        if (!bad) {
            std::vector<double> rhs(_n), step;
            for (int i = 0; i < _n; ++i) rhs[i] = -g[i];
            _hesseOk = cholSolve(_n, H, rhs, step, &_cov);
            if (_hesseOk) {
                double edm = 0;
                for (int i = 0; i < _n; ++i) edm -= 0.5 * g[i] * step[i];
                _edm = edm;
                // one more Newton polish with the accurate derivatives if it is still worth it
                if (edm > tolerance && _iter < maxIterations + 5) {
                    _g = g; _H = H; _step = step; _lambda = 0;
                    _alphas = {1.0, 0.5, 0.25, 0.1};
                    ++_iter;
                    _stage = LineSearch;
                    return;
                }
            }
        }
        _stage = Done;
    }
}

FunctionMinimumPtr NewtonFit::result() const {
    FunctionMinimumPtr fm(new FunctionMinimum());
    fm->parameters = _params;
    for (size_t i = 0; i < _params.size(); ++i) fm->parameters[i].value = _x[i];
    fm->minValue = _f0;
    fm->nEvaluations = _neval;
    fm->nIterations = _iter;
    fm->edm = _edm;
    if (_failed || !std::isfinite(_f0)) fm->status = FunctionMinimum::ERROR;
    else if (_hesseOk && _edm < 100 * tolerance) fm->status = FunctionMinimum::OK;
    else fm->status = FunctionMinimum::WARNING;
    if (_hesseOk && _n > 0) {
        fm->covariance = _cov;
        for (int i = 0; i < _n; ++i) fm->parameters[_fidx[i]].error = std::sqrt(std::max(0.0, _cov[i * _n + i]));
    }
    return fm;
}

// ------------------------------------------------------------------------------------------
// Levenberg-Marquardt on Gauss-Newton normal equations
// ------------------------------------------------------------------------------------------
void priorDerivatives(FitParameter const &p, double x, double &value, double &grad, double &hess) {
    value = grad = hess = 0;
    if (!p.floating || p.priorType == FitParameter::NoPrior) return;
    double c = 0.5 * (p.priorMin + p.priorMax), s = 0.5 * (p.priorMax - p.priorMin);
    if (p.priorType == FitParameter::GaussPrior) {
        double t = (x - c) / s;
        value = 0.5 * t * t; grad = t / s; hess = 1 / (s * s);
    } else {
        double w = 1e-3 * s;
        if (x < p.priorMin) { double d = p.priorMin - x; value = 0.5 * d * d / (w * w); grad = -d / (w * w); hess = 1 / (w * w); }
        else if (x > p.priorMax) { double d = x - p.priorMax; value = 0.5 * d * d / (w * w); grad = d / (w * w); hess = 1 / (w * w); }
    }
}

LMFit::LMFit(FitParameters const &params, double chi2Scale, double priorScale)
    : _stage(Jacobian), _params(params), _cs(chi2Scale), _ps(priorScale) {
    for (size_t i = 0; i < params.size(); ++i) {
        _x.push_back(params[i].value);
        if (params[i].floating) {
            _fidx.push_back((int)i);
            double e = params[i].error > 0 ? params[i].error : std::max(1e-3, 0.1 * std::fabs(params[i].value));
            _h.push_back(1e-3 * e);
            _scale.push_back(e);
        }
    }
    _n = (int)_fidx.size();
    _g.assign(_n, 0.0);
    _A.assign((size_t)_n * _n, 0.0);
    _step.assign(_n, 0.0);
}

double LMFit::priorValue(Parameters const &x) const { return _ps * evaluatePriors(_params, x); }

void LMFit::requestJacobian(std::vector<double> &flat) const {
    // columns: x, x + h_1, x - h_1, x + h_2, x - h_2, ...
    flat.insert(flat.end(), _x.begin(), _x.end());
    for (int i = 0; i < _n; ++i)
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
            size_t at = flat.size();
            flat.insert(flat.end(), _x.begin(), _x.end());
            flat[at + _fidx[i]] += sgn * _h[i];
        }
}

void LMFit::requestProbes(std::vector<double> &centres, std::vector<double> &steps) const {
    centres.insert(centres.end(), _x.begin(), _x.end());
    size_t at = steps.size();
    steps.resize(at + _x.size(), 0.0);
    for (int i = 0; i < _n; ++i) steps[at + _fidx[i]] = _h[i];
}

void LMFit::solveStep() {
    std::vector<double> rhs(_n), M;
    for (int i = 0; i < _n; ++i) rhs[i] = -_g[i];
    for (int attempt = 0; attempt < 60; ++attempt) {
        M = _A;
        for (int i = 0; i < _n; ++i) {
            // Marquardt scaling with a floor from the parameter's natural size: where the Gauss-Newton
            // curvature vanishes (e.g. bias -> 0, the model is quadratic in it) the damping still bites
            double d = std::max(_A[i * _n + i], 1e-2 / (_scale[i] * _scale[i]));
            M[i * _n + i] += _lambda * d;
        }
        if (cholSolve(_n, M, rhs, _step)) break;
        _lambda = _lambda > 0 ? _lambda * 10 : 1e-6;
        if (attempt == 59)
            for (int i = 0; i < _n; ++i) _step[i] = -_g[i] * _h[i] * _h[i];
    }
    // trust region in units of the parameters' natural sizes: no parameter moves by more than _radius
    // of them; the radius grows while capped full steps keep succeeding (valleys that run to a prior
    // wall, e.g. beta -> large with beta*bias fixed) and shrinks when a trial round fails
    double worst = 0;
    for (int i = 0; i < _n; ++i) worst = std::max(worst, std::fabs(_step[i]) / _scale[i]);
    _capped = worst > _radius;
    if (_capped)
        for (int i = 0; i < _n; ++i) _step[i] *= _radius / worst;
    // projection onto the box priors: a step that would cross a wall ends on it (see the active set in consumeNormal)
    for (int i = 0; i < _n; ++i) {
        FitParameter const &fp = _params[_fidx[i]];
        if (fp.priorType != FitParameter::BoxPrior) continue;
        const double x = _x[_fidx[i]];
        if (!_active.empty() && _active[i]) _step[i] = 0;
        else if (x <= fp.priorMax && x + _step[i] > fp.priorMax) _step[i] = fp.priorMax - x;
        else if (x >= fp.priorMin && x + _step[i] < fp.priorMin) _step[i] = fp.priorMin - x;
    }
}

void LMFit::consumeJacobian(const double *gram, bool bad) {
    // reduce the full Gram matrix of [y0, d+_1, d-_1, ...] (d = y(x +- h_i) - y0) to the normal-equation form
    const int G = 2 * _n + 1, R = _n + 1;
    for (int k = 0; k < G * G && !bad; ++k) bad |= !std::isfinite(gram[k]);
    std::vector<double> red((size_t)R * R + _n, 0.0);
    if (!bad) {
        auto gm = [&](int a, int b) { return gram[a * G + b]; };
        red[0] = gm(0, 0);
        for (int i = 0; i < _n; ++i) {
            const int ip = 2 * i + 1, im = 2 * i + 2;
            red[(i + 1) * R] = red[i + 1] = gm(ip, 0) - gm(im, 0);
            red[(size_t)R * R + i] = gm(ip, 0) + gm(im, 0);
            for (int j = 0; j < _n; ++j) {
                const int jp = 2 * j + 1, jm = 2 * j + 2;
                red[(i + 1) * R + j + 1] = gm(ip, jp) - gm(ip, jm) - gm(im, jp) + gm(im, jm);
            }
        }
    } else
        red[0] = gram[0];
    consumeNormal(red.data(), bad);
}

void LMFit::consumeNormal(const double *normal, bool bad) {
    const int G = 2 * _n + 1, R = _n + 1;
    _neval += G;
    const double inf = std::numeric_limits<double>::infinity();
    for (int k = 0; k < R * R + _n && !bad; ++k) bad |= !std::isfinite(normal[k]);
    if (bad) {
        // a probe (or the centre) left the valid domain: shrink the probe steps; a bad centre is fatal
        if (!std::isfinite(normal[0]) || ++_shrinks > 8) { _failed = true; _f0 = inf; _stage = Done; return; }
        for (auto &h : _h) h *= 0.25;
        return;
    }
    _f0 = _cs * 0.5 * normal[0] + priorValue(_x);
    // with a_i = y(x + h_i) - y(x - h_i) and s_i = y(x + h_i) + y(x - h_i) - 2 y0:
    //   J_i = a_i/(2 h_i),   y0.d2y/dx_i^2 = y0.s_i/h_i^2
    for (int i = 0; i < _n; ++i) {
        double pv, pg, ph;
        priorDerivatives(_params[_fidx[i]], _x[_fidx[i]], pv, pg, ph);
        _g[i] = _cs * normal[(i + 1) * R] / (2 * _h[i]) + _ps * pg;
        for (int j = 0; j < _n; ++j) _A[i * _n + j] = _cs * normal[(i + 1) * R + j + 1] / (4 * _h[i] * _h[j]);
        // diagonal of the second-order term, kept only while the curvature stays positive
        const double second = _cs * normal[(size_t)R * R + i] / (_h[i] * _h[i]);
        if (_A[i * _n + i] + second > 0.1 * _A[i * _n + i]) _A[i * _n + i] += second;
        _A[i * _n + i] += _ps * ph;
    }
    // Active set for box priors: a parameter that sits on a wall of its box while chi-square still falls towards the
    // outside stays on the wall (its row and column leave the normal equations; it is released as soon as the gradient
    // turns inward). Without this a valley that runs into a wall -- beta -> 100 with (1+beta)*bias fixed by the data in
    // BOSSDR11LyaF_k.ini -- is followed across the kink of the penalty in dozens of ever smaller damped steps; these few
    // fits set the number of lock-step rounds of a whole scan.
    _active.assign(_n, 0);
    for (int i = 0; i < _n; ++i) {
        FitParameter const &fp = _params[_fidx[i]];
        if (fp.priorType != FitParameter::BoxPrior) continue;
        const double x = _x[_fidx[i]], gchi = _cs * normal[(i + 1) * R] / (2 * _h[i]), eps = 1e-9 * (fp.priorMax - fp.priorMin);
        if ((x >= fp.priorMax - eps && gchi < 0) || (x <= fp.priorMin + eps && gchi > 0)) {
            _active[i] = 1;
            _g[i] = 0;
            for (int j = 0; j < _n; ++j) _A[i * _n + j] = _A[j * _n + i] = 0;
            _A[i * _n + i] = 1;
        }
    }
    if (_n == 0) { _converged = true; _stage = Done; return; }
    // estimated distance to the minimum with the undamped Gauss-Newton step
    {
        std::vector<double> rhs(_n), st;
        for (int i = 0; i < _n; ++i) rhs[i] = -_g[i];
        if (cholSolve(_n, _A, rhs, st)) {
            double edm = 0;
            for (int i = 0; i < _n; ++i) edm -= 0.5 * _g[i] * st[i];
            _edm = edm;
        } else
            _edm = inf;
    }
    ++_iter;
    if (debug) {
        std::fprintf(stderr, "LM it %d f0 %.6f edm %.3e lambda %.1e x", _iter, _f0, _edm, _lambda);
        for (int i = 0; i < _n; ++i) std::fprintf(stderr, " %.5g", _x[_fidx[i]]);
        std::fprintf(stderr, "\n");
    }
    if (_edm < tolerance) { _converged = true; _stage = Done; return; }
    if (_iter > maxIterations) { _stage = Done; return; }
    // probe steps from the curvature: delta F ~ 1e-4 along each axis, never more than a tenth of the
    // parameter's natural size
    for (int i = 0; i < _n; ++i) {
        double a = _A[i * _n + i];
        if (a > 0) _h[i] = std::min(0.1 * _scale[i], std::max(std::sqrt(2e-4 / a), 1e-9 * std::max(1.0, std::fabs(_x[_fidx[i]]))));
    }
    solveStep();
    _alphas = {1.0, 0.3, 0.1};
    _stage = Trial;
}

void LMFit::requestPoints(std::vector<double> &flat) {
    _npending = 0;
    for (double a : _alphas) {
        size_t at = flat.size();
        flat.insert(flat.end(), _x.begin(), _x.end());
        for (int i = 0; i < _n; ++i) flat[at + _fidx[i]] += a * _step[i];
        ++_npending;
    }
}

void LMFit::consumePoints(const double *chi2) {
    _neval += _npending;
    int best = -1;
    double fbest = _f0;
    Parameters trial(_x);
    for (size_t k = 0; k < _alphas.size(); ++k) {
        if (!std::isfinite(chi2[k])) continue;
        for (int i = 0; i < _n; ++i) trial[_fidx[i]] = _x[_fidx[i]] + _alphas[k] * _step[i];
        double f = _cs * 0.5 * chi2[k] + priorValue(trial);
        if (f < fbest) { fbest = f; best = (int)k; }
    }
    if (debug) std::fprintf(stderr, "   trial best %d f %.6f (f0 %.6f) stall %d lambda %.1e\n", best, fbest, _f0, _stall, _lambda);
    if (best >= 0) {
        double a = _alphas[best];
        for (int i = 0; i < _n; ++i) _x[_fidx[i]] += a * _step[i];
        const double gain = _f0 - fbest;
        _f0 = fbest;
        _stall = 0;
        _lambda = a >= 1.0 ? std::max(_lambda * 0.1, 1e-12) : _lambda;
        if (_capped && a >= 1.0) _radius *= 3.0;
        // a minimum with a flat direction (bias = 0 makes beta irrelevant: singular normal matrix, no
        // finite EDM) is recognised by accepted steps that no longer gain anything
        _smallGains = gain < 0.01 * tolerance ? _smallGains + 1 : 0;
        if (_smallGains >= 2) { _converged = true; _stage = Done; return; }
        _stage = Jacobian;
    } else {
        // no trial improved: damp more, same normal equations
        if (++_stall > 12) { _converged = _smallGains >= 1; _stage = Done; return; }
        _radius = std::max(_radius / 3.0, 1.0);
        _lambda = std::max(_lambda, 1e-4) * (_stall > 3 ? 100 : 10);
        solveStep();
        _alphas = {1.0, 0.3, 0.1};
    }
}

FunctionMinimumPtr LMFit::result() const {
    FunctionMinimumPtr fm(new FunctionMinimum());
    fm->parameters = _params;
    for (size_t i = 0; i < _params.size(); ++i) fm->parameters[i].value = _x[i];
    fm->minValue = _f0;
    fm->nEvaluations = _neval;
    fm->nIterations = _iter;
    fm->edm = _edm;
    if (_failed || !std::isfinite(_f0)) fm->status = FunctionMinimum::ERROR;
    else fm->status = _converged ? FunctionMinimum::OK : FunctionMinimum::WARNING;
    return fm;
}

void runLockStep(std::vector<NewtonFit> &fits, BatchFunction const &f) {
    std::vector<Parameters> points;
    std::vector<double> values;
    for (;;) {
        points.clear();
        bool any = false;
        for (auto &fit : fits) {
            if (fit.done()) continue;
            fit.request(points);
            any = true;
        }
        if (!any) break;
        values.assign(points.size(), 0.0);
        if (!points.empty()) f(points, values);
        size_t off = 0;
        for (auto &fit : fits) {
            if (fit.done()) continue;
            int n = fit.pending();
            fit.consume(values.data() + off);
            off += n;
        }
    }
}

}  // namespace likely
