// Host-side binned data: grid, loaders, final cuts (bit-exact restatement of the reference's
// comparisons), device upload.
#include <cmath>
#include <omp.h>
#include <cstring>
#include <fstream>
#include <sstream>

#include "baofit_host.h"
#include "../csrc/host_setup.h"

namespace baofit {

namespace {
void check(int rc, const char *what) {
    if (rc != BF_OK) throw RuntimeError(std::string(what) + ": " + bf_last_error());
}
}  // namespace

BinnedGrid::BinnedGrid(std::vector<double> a1, std::vector<double> a2, std::vector<double> a3) {
    _axis[0] = a1;
    _axis[1] = a2;
    _axis[2] = a3;
}

void BinnedGrid::getBinCenters(int index, double centers[3]) const {
    if (index < 0 || index >= getNBinsTotal()) throw RuntimeError("BinnedGrid: invalid bin index.");
    int n2 = (int)_axis[1].size(), n3 = (int)_axis[2].size();
    int i3 = index % n3, i2 = (index / n3) % n2, i1 = index / (n3 * n2);  // last axis fastest
    centers[0] = _axis[0][i1];
    centers[1] = _axis[1][i2];
    centers[2] = _axis[2][i3];
}

void BinnedGrid::setWidths(std::vector<double> w1, std::vector<double> w2, std::vector<double> w3) {
    _width[0] = w1;
    _width[1] = w2;
    _width[2] = w3;
}

void BinnedGrid::getBinWidths(int index, double widths[3]) const {
    if (index < 0 || index >= getNBinsTotal()) throw RuntimeError("BinnedGrid: invalid bin index.");
    int n2 = (int)_axis[1].size(), n3 = (int)_axis[2].size();
    int i3 = index % n3, i2 = (index / n3) % n2, i1 = index / (n3 * n2);
    widths[0] = _width[0].empty() ? 0.0 : _width[0][i1];
    widths[1] = _width[1].empty() ? 0.0 : _width[1][i2];
    widths[2] = _width[2].empty() ? 0.0 : _width[2][i3];
}

BinnedGrid createCorrelationGrid(std::string const &axis1Bins, std::string const &axis2Bins, std::string const &axis3Bins) {
    std::vector<double> a[3], w[3];
    const std::string *spec[3] = {&axis1Bins, &axis2Bins, &axis3Bins};
    for (int i = 0; i < 3; ++i) {
        try {
            a[i] = likely::createBinning(*spec[i]);
            w[i] = likely::createBinningWidths(*spec[i]);
        } catch (likely::RuntimeError const &) {
            throw RuntimeError("createCorrelationGrid: error in axis " + std::to_string(i + 1) + " binning.");
        }
    }
    BinnedGrid g(a[0], a[1], a[2]);
    g.setWidths(w[0], w[1], w[2]);
    return g;
}

AbsCorrelationData::AbsCorrelationData(BinnedGrid grid, TransverseBinningType type) : _grid(grid), _type(type) {}

AbsCorrelationData::~AbsCorrelationData() {
    if (_device) bf_data_destroy(_device);
}

void AbsCorrelationData::setData(int index, double value, bool weighted) {
    if (_finalized) throw RuntimeError("AbsCorrelationData: data is finalized.");
    if (index < 0 || index >= _grid.getNBinsTotal()) throw RuntimeError("AbsCorrelationData::setData: invalid bin index.");
    if (!_values.empty() && weighted != _weighted) throw RuntimeError("AbsCorrelationData::setData: cannot mix weighted and unweighted data.");
    _weighted = weighted;
    _values[index] = value;
}

void AbsCorrelationData::setCustomBinCenters(int index, double c1, double c2, double c3) {
    if (_finalized) throw RuntimeError("AbsCorrelationData: data is finalized.");
    if (index < 0 || index >= _grid.getNBinsTotal()) throw RuntimeError("AbsCorrelationData::setCustomBinCenters: invalid bin index.");
    _custom[index] = {{c1, c2, c3}};
}

void AbsCorrelationData::enableCustomGrid() {
    if (getNCustomBins() != _grid.getNBinsTotal())
        throw RuntimeError("loadCorrelationData: number of custom bins must match the number of default bins.");
    _useCustom = true;
}

void AbsCorrelationData::copyCustomGridFrom(AbsCorrelationData const &other) {
    _custom = other._custom;
    _useCustom = other._useCustom;
}

void AbsCorrelationData::getBinCenters(int index, double centers[3]) const {
    if (_useCustom) {
        auto const &c = _custom.at(index);
        centers[0] = c[0]; centers[1] = c[1]; centers[2] = c[2];
    } else
        _grid.getBinCenters(index, centers);
}

double AbsCorrelationData::getData(int index) const {
    auto it = _values.find(index);
    if (it == _values.end()) throw RuntimeError("AbsCorrelationData::getData: bin has no data.");
    return it->second;
}

double AbsCorrelationData::getCovarianceElement(int i, int j) const {
    auto it = _cov.find(std::make_pair(std::min(i, j), std::max(i, j)));
    return it == _cov.end() ? 0.0 : it->second;
}

void AbsCorrelationData::setCovariance(int i, int j, double value) {
    if (_haveCov && _isInverse) throw RuntimeError("AbsCorrelationData: cannot mix covariance and inverse covariance.");
    _haveCov = true;
    _isInverse = false;
    _cov[std::make_pair(std::min(i, j), std::max(i, j))] = value;
}

void AbsCorrelationData::setInverseCovariance(int i, int j, double value) {
    if (_haveCov && !_isInverse) throw RuntimeError("AbsCorrelationData: cannot mix covariance and inverse covariance.");
    _haveCov = true;
    _isInverse = true;
    _cov[std::make_pair(std::min(i, j), std::max(i, j))] = value;
}

void AbsCorrelationData::setFinalCuts(double rMin, double rMax, double rVetoMin, double rVetoMax, double muMin,
                                      double muMax, double rperpMin, double rperpMax, double rparMin, double rparMax,
                                      int lMin, int lMax, double zMin, double zMax) {
    if (rMin > rMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected r-min <= r-max.");
    if (rVetoMin > rVetoMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected rveto-min <= rveto-max.");
    if (muMin > muMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected mu-min <= mu-max.");
    if (rperpMin > rperpMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected rperp-min <= rperp-max.");
    if (rparMin > rparMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected rpar-min <= rpar-max.");
    if (lMin > lMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected lmin <= lmax.");
    if (zMin > zMax) throw RuntimeError("AbsCorrelationData::setFinalCuts: expected z-min <= z-max.");
    _rMin = rMin; _rMax = rMax; _rVetoMin = rVetoMin; _rVetoMax = rVetoMax; _muMin = muMin; _muMax = muMax;
    _rperpMin = rperpMin; _rperpMax = rperpMax; _rparMin = rparMin; _rparMax = rparMax; _lMin = lMin; _lMax = lMax;
    _zMin = zMin; _zMax = zMax;
    _haveFinalCuts = true;
}

namespace {
// inverse of a dense symmetric positive definite matrix through its Cholesky factor
std::vector<double> spdInverse(int n, std::vector<double> a, const char *what) {
    if (!bf::cholesky_lower(n, a.data())) throw RuntimeError(std::string("AbsCorrelationData: ") + what + " is not positive definite");
    std::vector<double> li((size_t)n * n, 0.0), out((size_t)n * n, 0.0);
    bf::invert_lower(n, a.data(), li.data());
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int k = i; k < n; ++k) s += li[(size_t)k * n + i] * li[(size_t)k * n + j];  // (L^-T L^-1)_ij
            out[(size_t)i * n + j] = out[(size_t)j * n + i] = s;
        }
    return out;
}
}  // namespace

void AbsCorrelationData::finalize() {
    if (_finalized) return;
    if (!_haveFinalCuts) throw RuntimeError("AbsCorrelationData: no final cuts specified yet.");
    // same comparisons in the same order as AbsCorrelationData.cc:71-93 (std::map iterates in
    // ascending index order, like the reference's IndexIterator)
    for (auto const &kv : _values) {
        int index = kv.first;
        double r = getRadius(index), z = getRedshift(index);
        if (r < _rMin || r > _rMax) continue;
        if (r > _rVetoMin && r < _rVetoMax) continue;
        if (z < _zMin || z > _zMax) continue;
        if (_type == Coordinate) {
            double mu = getCosAngle(index);
            if (mu < _muMin || mu > _muMax) continue;
            double rpar = r * mu;
            if (rpar < _rparMin || rpar > _rparMax) continue;
            double rperp = r * std::sqrt(1 - mu * mu);
            if (rperp < _rperpMin || rperp > _rperpMax) continue;
        } else {  // _type == Multipole, AbsCorrelationData.cc:86-89
            int ell = getMultipole(index);
            if (ell < _lMin || ell > _lMax) continue;
        }
        if (extraCut(index)) continue;
        _kept.push_back(index);
    }
    if (_kept.empty()) throw RuntimeError("CorrelationFitter: need some data to fit.");
    if (!_haveCov) throw RuntimeError("AbsCorrelationData: no covariance loaded.");
    const bool pruned = _kept.size() < _values.size();
    if (_weighted || (_isInverse && pruned)) {
        // Both steps need the covariance over ALL bins with data. Weighted values are elements of Cinv.d, so
        // d = C.(Cinv.d). Dropping bins from a data set whose matrix was given as an inverse means marginalising
        // over them, i.e. taking the sub-block of the covariance, not of its inverse (likely::BinnedData::prune
        // works on the covariance representation; UNVERIFIED (dep), but the only choice that keeps the
        // likelihood of the kept bins unchanged).
        std::vector<int> bins = binsWithData();
        const int n = (int)bins.size();
        std::vector<double> m = denseOverBinsWithData();
        std::vector<double> cov = _isInverse ? spdInverse(n, m, "the inverse covariance") : m;
        if (_weighted) {
            std::vector<double> w(n);
            for (int i = 0; i < n; ++i) w[i] = _values.at(bins[i]);
            for (int i = 0; i < n; ++i) {
                double sum = 0;
                for (int j = 0; j < n; ++j) sum += cov[(size_t)i * n + j] * w[j];
                _values[bins[i]] = sum;
            }
            _weighted = false;
        }
        if (_isInverse && pruned) {
            _cov.clear();
            for (int i = 0; i < n; ++i)
                for (int j = 0; j <= i; ++j) _cov[std::make_pair(std::min(bins[i], bins[j]), std::max(bins[i], bins[j]))] = cov[(size_t)i * n + j];
            _isInverse = false;
        }
    }
    _finalized = true;
}

// Eigenmodes of a dense symmetric matrix (likely::CovarianceMatrix::getEigenModes calls LAPACK there; the dependency is
// not in the tree): Householder reflections bring the matrix to tridiagonal form, implicit-shift QL sweeps diagonalise it.
// values ascending; vectors[k*n + i] = component i of mode k, unit norm (the layout BinnedData::projectOntoModes reads).
void symmetricEigenModes(int n, std::vector<double> a, std::vector<double> &values, std::vector<double> &vectors) {
    if (n <= 0 || a.size() != (size_t)n * n) throw RuntimeError("symmetricEigenModes: bad dimensions");
    std::vector<double> d(n), e(n, 0.0), beta(n, 0.0), v(n), pv(n);
    auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
    // reflection k zeroes column k below the sub-diagonal; its vector stays in that column (rows k+1..)
    for (int k = 0; k + 2 < n; ++k) {
        double tail = 0;
        for (int i = k + 2; i < n; ++i) tail += A(i, k) * A(i, k);
        const double x0 = A(k + 1, k);
        if (tail == 0.0) { e[k] = x0; continue; }
        const double alpha = x0 > 0 ? -std::sqrt(x0 * x0 + tail) : std::sqrt(x0 * x0 + tail);
        for (int i = k + 1; i < n; ++i) v[i] = A(i, k);
        v[k + 1] -= alpha;
        double vv = 0;
        for (int i = k + 1; i < n; ++i) vv += v[i] * v[i];
        const double b = 2.0 / vv;
        // p = b A22 v, w = p - (b/2)(v.p) v, A22 -= v w^T + w v^T (both triangles are kept: rows stay contiguous)
        const bool par = n - k >= 640;  // rows are independent in both passes; the sums keep their order within a row
#pragma omp parallel for schedule(static) if (par)
        for (int i = k + 1; i < n; ++i) {
            const double *row = &a[(size_t)i * n];
            double sum = 0;
            for (int j = k + 1; j < n; ++j) sum += row[j] * v[j];
            pv[i] = b * sum;
        }
        double vp = 0;
        for (int i = k + 1; i < n; ++i) vp += v[i] * pv[i];
        const double half = 0.5 * b * vp;
        for (int i = k + 1; i < n; ++i) pv[i] -= half * v[i];
#pragma omp parallel for schedule(static) if (par)
        for (int i = k + 1; i < n; ++i) {
            double *row = &a[(size_t)i * n];
            const double vi = v[i], wi = pv[i];
            for (int j = k + 1; j < n; ++j) row[j] -= vi * pv[j] + wi * v[j];
        }
        e[k] = alpha;
        beta[k] = b;
        for (int i = k + 1; i < n; ++i) A(i, k) = v[i];
    }
    if (n >= 2) e[n - 2] = A(n - 1, n - 2);
    for (int i = 0; i < n; ++i) d[i] = A(i, i);
    // Z = Q^T with Q = H_0 H_1 ...: rows of Z are the (future) modes, so a QL rotation mixes two contiguous rows
    std::vector<double> z((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) z[(size_t)i * n + i] = 1.0;
    for (int k = n - 3; k >= 0; --k) {
        if (beta[k] == 0.0) continue;
        // Q <- H_k Q on rows k+1.. of Q, i.e. on columns k+1.. of Z: Z[r][:] -= beta (Z[r][:].v) v for every row r > k
        for (int i = k + 1; i < n; ++i) v[i] = A(i, k);  // contiguous copy of the reflection vector
#pragma omp parallel for schedule(static) if (n - k >= 640)
        for (int r = k + 1; r < n; ++r) {
            double *row = &z[(size_t)r * n];
            double dot = 0;
            for (int i = k + 1; i < n; ++i) dot += row[i] * v[i];
            dot *= beta[k];
            for (int i = k + 1; i < n; ++i) row[i] -= dot * v[i];
        }
    }
    // implicit-shift QL on (d, e)
    const double eps = 2.220446049250313e-16;
    struct Rotation { int i; double c, s; };
    std::vector<Rotation> rot;
    rot.reserve(n);
    for (int l = 0; l < n; ++l) {
        for (int iter = 0;; ++iter) {
            int m = l;
            for (; m + 1 < n; ++m)
                if (std::fabs(e[m]) <= eps * (std::fabs(d[m]) + std::fabs(d[m + 1]))) break;
            if (m == l) break;
            if (iter == 80) throw RuntimeError("symmetricEigenModes: QL iteration did not converge");
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = std::hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + (g >= 0 ? r : -r));
            double sn = 1.0, cs = 1.0, p = 0.0;
            int i = m - 1;
            rot.clear();
            for (; i >= l; --i) {
                double f = sn * e[i], bb = cs * e[i];
                r = std::hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                sn = f / r;
                cs = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * sn + 2.0 * cs * bb;
                p = sn * r;
                d[i + 1] = g + p;
                g = cs * r - bb;
                rot.push_back({i, cs, sn});
            }
            // the rotations of this sweep mix rows (i, i+1) of Z in sequence; column slices are independent of each other
            {
                const int nrot = (int)rot.size();
                const int slices = (n >= 768 && nrot >= 16) ? std::max(1, std::min(omp_get_max_threads(), n / 64)) : 1;
#pragma omp parallel for schedule(static) num_threads(slices) if (slices > 1)
                for (int sl = 0; sl < slices; ++sl) {
                    const int q0 = (int)((long long)n * sl / slices), q1 = (int)((long long)n * (sl + 1) / slices);
                    for (int k = 0; k < nrot; ++k) {
                        double *zi = &z[(size_t)rot[k].i * n], *zi1 = zi + n;
                        const double c = rot[k].c, sgn = rot[k].s;
                        for (int q = q0; q < q1; ++q) {
                            const double t = zi1[q];
                            zi1[q] = sgn * zi[q] + c * t;
                            zi[q] = c * zi[q] - sgn * t;
                        }
                    }
                }
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p;
            e[l] = g;
            e[m] = 0.0;
        }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
    values.resize(n);
    vectors.resize((size_t)n * n);
    for (int k = 0; k < n; ++k) {
        values[k] = d[order[k]];
        std::copy(z.begin() + (size_t)order[k] * n, z.begin() + (size_t)(order[k] + 1) * n, vectors.begin() + (size_t)k * n);
    }
}

// Q(a, x): series for x < a + 1, Lentz continued fraction otherwise (the two standard expansions of the incomplete gamma function)
double gammaQ(double a, double x) {
    if (!(a > 0) || x < 0) throw RuntimeError("gammaQ: invalid arguments");
    if (x == 0) return 1.0;
    const double lg = std::lgamma(a);
    if (x < a + 1.0) {
        double term = 1.0 / a, sum = term, ap = a;
        for (int it = 0; it < 100000; ++it) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (std::fabs(term) < std::fabs(sum) * 1e-16) break;
        }
        return 1.0 - sum * std::exp(-x + a * std::log(x) - lg);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, dd = 1.0 / b, h = dd;
    for (int it = 1; it < 100000; ++it) {
        const double an = -it * (it - a);
        b += 2.0;
        dd = an * dd + b;
        if (std::fabs(dd) < tiny) dd = tiny;
        c = b + an / c;
        if (std::fabs(c) < tiny) c = tiny;
        dd = 1.0 / dd;
        const double delta = dd * c;
        h *= delta;
        if (std::fabs(delta - 1.0) < 1e-16) break;
    }
    return std::exp(-x + a * std::log(x) - lg) * h;
}

void dataAndCovariance(AbsCorrelationData const &data, bool finalized, std::vector<double> &d, std::vector<double> &cov) {
    if (finalized) {
        if (!data.isFinalized()) throw RuntimeError("dataAndCovariance: data set is not finalized.");
        d = data.dataVector();
        cov = data.covarianceMatrix();
    } else {
        std::vector<int> bins = data.binsWithData();
        d.resize(bins.size());
        for (size_t i = 0; i < bins.size(); ++i) d[i] = data.getData(bins[i]);
        cov = data.denseOverBinsWithData();
    }
    if (data.isWeighted()) throw RuntimeError("dataAndCovariance: weighted data (Cinv.d) must be finalized first.");
    if (data.isInverse()) cov = spdInverse((int)d.size(), cov, "the inverse covariance");
}

// likely::BinnedData::projectOntoModes as src/baofit.cc:597-603 uses it (before the final cuts): the data vector is replaced by
// its projection onto the nkeep eigenmodes of the covariance with the largest (nkeep > 0) or smallest (nkeep < 0) variance;
// the covariance stays as it is. UNVERIFIED (dep): restated from the option's documentation (src/baofit.cc:194-195).
void AbsCorrelationData::projectOntoModes(int nkeep) {
    if (_finalized) throw RuntimeError("AbsCorrelationData::projectOntoModes: data set is already finalized.");
    if (!_haveCov) throw RuntimeError("BinnedData::projectOntoModes: no covariance available.");
    std::vector<int> bins = binsWithData();
    const int n = (int)bins.size();
    if (nkeep == 0 || nkeep > n || -nkeep > n) throw RuntimeError("BinnedData::projectOntoModes: invalid nkeep.");
    std::vector<double> m = denseOverBinsWithData();
    std::vector<double> cov = _isInverse ? spdInverse(n, m, "the inverse covariance") : m;
    std::vector<double> data(n);
    for (int i = 0; i < n; ++i) data[i] = _values.at(bins[i]);
    if (_weighted) {  // stored values are elements of Cinv.d
        std::vector<double> w = data;
        for (int i = 0; i < n; ++i) {
            double sum = 0;
            for (int j = 0; j < n; ++j) sum += cov[(size_t)i * n + j] * w[j];
            data[i] = sum;
        }
        _weighted = false;
    }
    std::vector<double> values, modes;
    symmetricEigenModes(n, cov, values, modes);
    const int first = nkeep > 0 ? n - nkeep : 0, last = nkeep > 0 ? n : -nkeep;
    std::vector<double> projected(n, 0.0);
    for (int k = first; k < last; ++k) {
        const double *mode = &modes[(size_t)k * n];
        double dot = 0;
        for (int i = 0; i < n; ++i) dot += mode[i] * data[i];
        for (int i = 0; i < n; ++i) projected[i] += dot * mode[i];
    }
    for (int i = 0; i < n; ++i) _values[bins[i]] = projected[i];
}

// "fix-mode-scales" (src/baofit.cc:192,514-524,570-572): every loaded data set has the eigenvalues of its covariance multiplied
// by the scales read from a file, smallest-variance mode first (likely::CovarianceMatrix::rescaleEigenvalues; UNVERIFIED (dep)).
// The data vector does not change; a matrix that was loaded as an inverse is kept as a covariance from here on.
void AbsCorrelationData::rescaleEigenvalues(std::vector<double> const &modeScales) {
    if (_finalized) throw RuntimeError("AbsCorrelationData::rescaleEigenvalues: data set is already finalized.");
    if (!_haveCov) throw RuntimeError("BinnedData::rescaleEigenvalues: no covariance available.");
    std::vector<int> bins = binsWithData();
    const int n = (int)bins.size();
    if ((int)modeScales.size() != n) throw RuntimeError("CovarianceMatrix::rescaleEigenvalues: unexpected number of mode scales.");
    for (double scale : modeScales)
        if (!(scale > 0)) throw RuntimeError("CovarianceMatrix::rescaleEigenvalues: expected all mode scales > 0.");
    std::vector<double> m = denseOverBinsWithData();
    std::vector<double> cov = _isInverse ? spdInverse(n, m, "the inverse covariance") : m;
    if (_weighted) {  // stored values are elements of Cinv.d with the covariance as loaded
        std::vector<double> w(n);
        for (int i = 0; i < n; ++i) w[i] = _values.at(bins[i]);
        for (int i = 0; i < n; ++i) {
            double sum = 0;
            for (int j = 0; j < n; ++j) sum += cov[(size_t)i * n + j] * w[j];
            _values[bins[i]] = sum;
        }
        _weighted = false;
    }
    std::vector<double> values, modes;
    symmetricEigenModes(n, cov, values, modes);
    std::fill(cov.begin(), cov.end(), 0.0);
    for (int k = 0; k < n; ++k) {
        const double *mode = &modes[(size_t)k * n];
        const double lambda = values[k] * modeScales[k];
        for (int i = 0; i < n; ++i) {
            const double a = lambda * mode[i];
            for (int j = 0; j <= i; ++j) cov[(size_t)i * n + j] += a * mode[j];
        }
    }
    _cov.clear();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) _cov[std::make_pair(std::min(bins[i], bins[j]), std::max(bins[i], bins[j]))] = cov[(size_t)i * n + j];
    _isInverse = false;
}

void AbsCorrelationData::gridCoordinates(std::vector<double> &r, std::vector<double> &mu, std::vector<double> &z) const {
    int n = _grid.getNBinsTotal();
    r.resize(n); mu.resize(n); z.resize(n);
    for (int i = 0; i < n; ++i) { r[i] = getRadius(i); mu[i] = _type == Coordinate ? getCosAngle(i) : (double)getMultipole(i); z[i] = getRedshift(i); }
}

std::vector<int> AbsCorrelationData::binsWithData() const {
    std::vector<int> b;
    for (auto const &kv : _values) b.push_back(kv.first);
    return b;
}

std::vector<double> AbsCorrelationData::denseOverBinsWithData() const {
    std::vector<int> bins = binsWithData();
    int n = (int)bins.size();
    std::map<int, int> pos;
    for (int i = 0; i < n; ++i) pos[bins[i]] = i;
    std::vector<double> c((size_t)n * n, 0.0);
    for (auto const &kv : _cov) {
        int a = pos.at(kv.first.first), b = pos.at(kv.first.second);
        c[(size_t)a * n + b] = c[(size_t)b * n + a] = kv.second;
    }
    return c;
}

std::vector<double> AbsCorrelationData::dataVector() const {
    std::vector<double> d;
    for (int idx : _kept) d.push_back(_values.at(idx));
    return d;
}

std::vector<double> AbsCorrelationData::covarianceMatrix() const {
    int n = (int)_kept.size();
    std::map<int, int> pos;
    for (int i = 0; i < n; ++i) pos[_kept[i]] = i;
    std::vector<double> c((size_t)n * n, 0.0);
    for (auto const &kv : _cov) {
        auto a = pos.find(kv.first.first), b = pos.find(kv.first.second);
        if (a == pos.end() || b == pos.end()) continue;
        c[(size_t)a->second * n + b->second] = kv.second;
        c[(size_t)b->second * n + a->second] = kv.second;
    }
    return c;
}

void AbsCorrelationData::saveData(std::ostream &out) const {
    if (!_finalized) throw RuntimeError("AbsCorrelationData::saveData: data is not finalized.");
    char buf[64];
    for (int idx : _kept) {
        std::snprintf(buf, sizeof(buf), "%d %.17g", idx, _values.at(idx));
        out << buf << std::endl;
    }
}

void AbsCorrelationData::saveInverseCovariance(std::ostream &out) const {
    if (!_finalized) throw RuntimeError("AbsCorrelationData::saveInverseCovariance: data is not finalized.");
    const int n = (int)_kept.size();
    std::vector<double> m = covarianceMatrix();
    std::vector<double> icov = _isInverse ? m : spdInverse(n, m, "the covariance");
    char buf[96];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            std::snprintf(buf, sizeof(buf), "%d %d %.17g", _kept[i], _kept[j], icov[(size_t)i * n + j]);
            out << buf << std::endl;
        }
}

bf_data *AbsCorrelationData::deviceData(bool withGrid) {
    if (!_finalized) finalize();
    if (_device && _deviceHasGrid == withGrid) return _device;
    if (_device) { bf_data_destroy(_device); _device = nullptr; }
    int n = (int)_kept.size();
    std::vector<double> r(n), mu(n), z(n), d = dataVector(), cov = covarianceMatrix(), gr, gmu, gz;
    for (int i = 0; i < n; ++i) {
        r[i] = getRadius(_kept[i]);
        mu[i] = _type == Coordinate ? getCosAngle(_kept[i]) : (double)getMultipole(_kept[i]);  // multipole bins carry ell
        z[i] = getRedshift(_kept[i]);
    }
    bf_data_desc dd;
    std::memset(&dd, 0, sizeof(dd));
    dd.binning_type = _type == Coordinate ? BF_BINNING_COORDINATE : BF_BINNING_MULTIPOLE;
    dd.n_data = n;
    dd.r = r.data(); dd.mu = mu.data(); dd.z = z.data(); dd.index = _kept.data(); dd.data = d.data();
    dd.cov = cov.data();
    dd.cov_is_inverse = _isInverse ? 1 : 0;
    if (withGrid) {
        gridCoordinates(gr, gmu, gz);
        dd.n_grid = (int)gr.size();
        dd.grid_r = gr.data(); dd.grid_mu = gmu.data(); dd.grid_z = gz.data();
    }
    check(bf_data_create(deviceContext(), &dd, &_device), "bf_data_create");
    _deviceHasGrid = withGrid;
    return _device;
}

ComovingCorrelationData::ComovingCorrelationData(BinnedGrid grid, CoordinateSystem coordinateSystem)
    : AbsCorrelationData(grid, coordinateSystem == MultipoleCoordinates ? Multipole : Coordinate),
      _coordinateSystem(coordinateSystem) {}

double ComovingCorrelationData::getRadius(int index) const {
    double c[3];
    getBinCenters(index, c);
    if (_coordinateSystem == CartesianCoordinates) {
        double rpar = c[0], rperp = c[1];
        return std::sqrt(rpar * rpar + rperp * rperp);
    }
    return c[0];
}

double ComovingCorrelationData::getCosAngle(int index) const {
    double c[3];
    getBinCenters(index, c);
    if (_coordinateSystem == PolarCoordinates) return c[1];
    if (_coordinateSystem == CartesianCoordinates) {
        double rpar = c[0], rperp = c[1];
        return rpar / std::sqrt(rpar * rpar + rperp * rperp);
    }
    throw RuntimeError("ComovingCorrelationData::getMultipole: invalid coordinate system.");
}

int ComovingCorrelationData::getMultipole(int index) const {
    double c[3];
    getBinCenters(index, c);
    if (_coordinateSystem == MultipoleCoordinates) return (int)std::floor(c[1] + 0.5);
    throw RuntimeError("ComovingCorrelationData::getMultipole: invalid coordinate system.");
}

double ComovingCorrelationData::getRedshift(int index) const {
    double c[3];
    getBinCenters(index, c);
    return c[2];
}


// ------------------------------------------------------------------------------------------
// BinnedDataResampler
// ------------------------------------------------------------------------------------------
void BinnedDataResampler::addObservation(AbsCorrelationDataPtr obs) {
    if (!obs) throw RuntimeError("BinnedDataResampler: null observation.");
    std::vector<int> bins = obs->binsWithData();
    if (_icov.empty()) { _bins = bins; _first = obs; }
    else if (bins != _bins) throw RuntimeError("combineCorrelationData: observations are not congruent.");
    const int n = (int)bins.size();
    std::vector<double> m = obs->denseOverBinsWithData();
    std::vector<double> icov = obs->isInverse() ? m : spdInverse(n, m, "a covariance");
    std::vector<double> wd(n, 0.0);
    for (int i = 0; i < n; ++i) {
        if (obs->isWeighted()) { wd[i] = obs->getData(bins[i]); continue; }  // ".wdata": already Cinv.d
        double s = 0;
        for (int j = 0; j < n; ++j) s += icov[(size_t)i * n + j] * obs->getData(bins[j]);
        wd[i] = s;
    }
    _icov.push_back(std::move(icov));
    _icovData.push_back(std::move(wd));
}

AbsCorrelationDataPtr BinnedDataResampler::combine(std::vector<int> const &counts, bool fixCovariance) const {
    if ((int)counts.size() != getNObservations()) throw RuntimeError("BinnedDataResampler::combine: one count per observation expected.");
    const int n = (int)_bins.size();
    std::vector<double> icovSum((size_t)n * n, 0.0), icovSq, icovData(n, 0.0);
    bool repeated = false, any = false;
    for (int c : counts) { if (c < 0) throw RuntimeError("BinnedDataResampler::combine: negative count."); repeated |= c > 1; any |= c > 0; }
    if (!any) throw RuntimeError("BinnedDataResampler::combine: no observations.");
    const bool fix = fixCovariance && repeated;
    if (fix) icovSq.assign((size_t)n * n, 0.0);
    for (size_t k = 0; k < counts.size(); ++k) {
        if (counts[k] == 0) continue;
        const double w = counts[k];
        for (size_t q = 0; q < icovSum.size(); ++q) icovSum[q] += w * _icov[k][q];
        if (fix)
            for (size_t q = 0; q < icovSq.size(); ++q) icovSq[q] += w * w * _icov[k][q];
        for (int i = 0; i < n; ++i) icovData[i] += w * _icovData[k][i];
    }
    std::vector<double> cov = spdInverse(n, icovSum, "the summed inverse covariance");
    AbsCorrelationDataPtr out = _prototype();
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += cov[(size_t)i * n + j] * icovData[j];
        out->setData(_bins[i], s);
    }
    if (fix) {  // C (sum n_k^2 Cinv_k) C
        std::vector<double> t((size_t)n * n, 0.0), f((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) {
                const double a = cov[(size_t)i * n + k];
                for (int j = 0; j < n; ++j) t[(size_t)i * n + j] += a * icovSq[(size_t)k * n + j];
            }
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) {
                const double a = t[(size_t)i * n + k];
                for (int j = 0; j < n; ++j) f[(size_t)i * n + j] += a * cov[(size_t)k * n + j];
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) cov[(size_t)i * n + j] = cov[(size_t)j * n + i] = 0.5 * (f[(size_t)i * n + j] + f[(size_t)j * n + i]);
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) out->setCovariance(_bins[i], _bins[j], cov[(size_t)i * n + j]);
    if (_first->useCustomGrid()) out->copyCustomGridFrom(*_first);
    return out;
}

AbsCorrelationDataPtr BinnedDataResampler::combined() const {
    return combine(std::vector<int>(getNObservations(), 1), false);
}

static long binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    double c = 1;
    for (int i = 1; i <= k; ++i) c = c * (n - k + i) / i;
    return c > 9e18 ? (long)9e18 : (long)(c + 0.5);
}

long BinnedDataResampler::jackknifeSize(int nobs, int ndrop) {
    if (ndrop <= 0 || ndrop >= nobs) return 0;  // at least one observation is dropped and one is left
    return binomial(nobs, ndrop);
}

AbsCorrelationDataPtr BinnedDataResampler::jackknife(int ndrop, long seqno) const {
    const int nobs = getNObservations();
    if (seqno < 0 || seqno >= jackknifeSize(nobs, ndrop)) return AbsCorrelationDataPtr();
    // seqno-th ndrop-subset of {0..nobs-1} in lexicographic order
    std::vector<int> counts(nobs, 1);
    long rank = seqno;
    int next = 0;
    for (int slot = 0; slot < ndrop; ++slot)
        for (int c = next; c < nobs; ++c) {
            const long with_c = binomial(nobs - c - 1, ndrop - slot - 1);  // subsets whose slot-th element is c
            if (rank < with_c) { counts[c] = 0; next = c + 1; break; }
            rank -= with_c;
        }
    return combine(counts, false);
}

std::vector<int> BinnedDataResampler::bootstrapCounts(int size, unsigned long long seed, long trial) const {
    const int nobs = getNObservations();
    if (size <= 0) size = nobs;
    std::vector<int> counts(nobs, 0);
    // splitmix64 on (seed, trial, draw): independent of how trials are distributed over ranks
    auto mix = [](unsigned long long z) {
        z += 0x9e3779b97f4a7c15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    };
    const unsigned long long key = mix(mix(seed) ^ (unsigned long long)trial);
    for (int q = 0; q < size; ++q) {
        unsigned long long u = mix(key + (unsigned long long)q * 0x9e3779b97f4a7c15ULL);
        counts[(int)((u >> 11) * (1.0 / 9007199254740992.0) * nobs)]++;
    }
    return counts;
}

AbsCorrelationDataPtr BinnedDataResampler::bootstrap(int size, bool fixCovariance, unsigned long long seed, long trial) const {
    return combine(bootstrapCounts(size, seed, trial), fixCovariance);
}

AbsCorrelationDataPtr BinnedDataResampler::getObservationCopy(int index) const {
    if (index < 0 || index >= getNObservations()) throw RuntimeError("BinnedDataResampler::getObservationCopy: bad index.");
    std::vector<int> counts(getNObservations(), 0);
    counts[index] = 1;
    return combine(counts, false);
}

void loadCorrelationData(std::string const &dataName, AbsCorrelationData &data, bool icov, bool weighted, bool customGrid) {
    std::string paramsName = dataName + (weighted ? ".wdata" : ".data");
    std::ifstream in(paramsName.c_str());
    if (!in.good()) throw RuntimeError("loadCorrelationData: Unable to open " + paramsName);
    std::string line;
    int lines = 0;
    while (std::getline(in, line)) {
        lines++;
        std::istringstream ss(line);
        int index;
        double value;
        // first two tokens only: trailing junk is accepted like the reference's phrase_parse
        if (!(ss >> index >> value))
            throw RuntimeError("loadCorrelationData: error reading line " + std::to_string(lines) + " of " + paramsName);
        data.setData(index, value, weighted);
    }
    in.close();
    int nbins = data.getGrid().getNBinsTotal();
    std::string covName = dataName + (icov ? ".icov" : ".cov");
    std::ifstream cin(covName.c_str());
    if (!cin.good()) throw RuntimeError("loadCorrelationData: Unable to open " + covName);
    lines = 0;
    while (std::getline(cin, line)) {
        lines++;
        std::istringstream ss(line);
        int i1, i2;
        double value;
        if (!(ss >> i1 >> i2 >> value))
            throw RuntimeError("loadCorrelationData: error reading line " + std::to_string(lines) + " of " + covName);
        if (i1 < 0 || i2 < 0 || i1 >= nbins || i2 >= nbins || !data.hasData(i1) || !data.hasData(i2))
            throw RuntimeError("loadCorrelationData: invalid covariance indices on line " + std::to_string(lines) + " of " + covName);
        if (icov) data.setInverseCovariance(i1, i2, value);
        else data.setCovariance(i1, i2, value);
    }
    if (customGrid) {  // AbsCorrelationData.cc:243-274
        std::string gridName = dataName + ".grid";
        std::ifstream gin(gridName.c_str());
        if (!gin.good()) throw RuntimeError("loadCorrelationData: Unable to open " + gridName);
        lines = 0;
        while (std::getline(gin, line)) {
            lines++;
            std::istringstream ss(line);
            int index;
            double c1, c2, c3;
            if (!(ss >> index >> c1 >> c2 >> c3))
                throw RuntimeError("loadCorrelationData: error reading line " + std::to_string(lines) + " of " + gridName);
            if (index < 0 || index >= nbins)
                throw RuntimeError("loadCorrelationData: invalid bin index on line " + std::to_string(lines) + " of " + gridName);
            data.setCustomBinCenters(index, c1, c2, c3);
        }
        data.enableCustomGrid();
    }
}

// ------------------------------------------------------------------------------------------
// cosmology + quasar coordinates
// ------------------------------------------------------------------------------------------
LambdaCdmRadiationUniverse::LambdaCdmRadiationUniverse(double OmegaMatter, double OmegaK, double hubbleConstant,
                                                       double Tcmb, double Nnu)
    : _OmegaMatter(OmegaMatter), _OmegaK(OmegaK), _h(hubbleConstant) {
    if (hubbleConstant <= 0) throw RuntimeError("LambdaCdmRadiationUniverse: expected hubbleConstant > 0.");
    double t = Tcmb / 2.725;
    _OmegaRadiation = 2.469e-5 * t * t * t * t * (1 + 0.2271 * Nnu) / (hubbleConstant * hubbleConstant);
}

double LambdaCdmRadiationUniverse::getHubbleFunction(double z) const {
    double zp1 = 1 + z, ol = 1 - _OmegaMatter - _OmegaK - _OmegaRadiation;
    return std::sqrt(ol + zp1 * zp1 * (_OmegaK + zp1 * (_OmegaMatter + zp1 * _OmegaRadiation)));
}

double LambdaCdmRadiationUniverse::getLineOfSightComovingDistance(double z) const {
    // c/H0 Int_0^z dz'/E(z'), composite 8-point Gauss-Legendre on panels of 0.05 in z: the integrand
    // is smooth, the result is converged to rounding
    static const double x[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    static const double w[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    if (z == 0) return 0;
    int n = (int)std::ceil(std::fabs(z) / 0.05);
    double h = z / n, sum = 0;
    for (int i = 0; i < n; ++i) {
        double mid = (i + 0.5) * h, part = 0;
        for (int q = 0; q < 4; ++q) part += w[q] * (1 / getHubbleFunction(mid + 0.5 * h * x[q]) + 1 / getHubbleFunction(mid - 0.5 * h * x[q]));
        sum += 0.5 * h * part;
    }
    const double hubbleLength = 299792.458 / 100.;  // c/(100 km/s/Mpc) in Mpc/h
    return hubbleLength * sum;
}

double LambdaCdmRadiationUniverse::getTransverseComovingScale(double z) const {
    const double hubbleLength = 299792.458 / 100.;
    double dc = getLineOfSightComovingDistance(z);
    if (_OmegaK == 0) return dc;
    double sk = std::sqrt(std::fabs(_OmegaK)), x = sk * dc / hubbleLength;
    return hubbleLength / sk * (_OmegaK > 0 ? std::sinh(x) : std::sin(x));
}

QuasarCorrelationData::QuasarCorrelationData(BinnedGrid grid, double llMin, double llMax, double sepMin, double sepMax,
                                             bool fixCov, AbsHomogeneousUniversePtr cosmology)
    : AbsCorrelationData(grid, Coordinate), _llMin(llMin), _llMax(llMax), _sepMin(sepMin), _sepMax(sepMax),
      _cosmology(cosmology) {
    // QuasarCorrelationData.cc:25-37
    if (llMin > llMax) throw RuntimeError("QuasarCorrelationData: expected llMin <= llMax.");
    if (sepMin > sepMax) throw RuntimeError("QuasarCorrelationData: expected sepMin <= sepMax.");
    if (fixCov) throw RuntimeError("QuasarCorrelationData: fixCovariance is not part of the accelerated path (src/baofit.cc:496 never sets it).");
    if (!cosmology) throw RuntimeError("QuasarCorrelationData: need a cosmology.");
    _arcminToRad = 4 * std::atan(1) / (60. * 180.);
}

void QuasarCorrelationData::transform(double ll, double sep, double dsep, double z, double &r, double &mu) const {
    // QuasarCorrelationData.cc:139-152
    double ratio(std::exp(0.5 * ll)), zp1(z + 1);
    double z1(zp1 / ratio - 1), z2(zp1 * ratio - 1);
    double drLos = _cosmology->getLineOfSightComovingDistance(z2) - _cosmology->getLineOfSightComovingDistance(z1);
    double swgt = sep + (dsep * dsep / 12) / sep;
    double drPerp = _cosmology->getTransverseComovingScale(z) * (swgt * _arcminToRad);
    double rsq = drLos * drLos + drPerp * drPerp;
    r = std::sqrt(rsq);
    mu = std::abs(drLos) / r;
}

void QuasarCorrelationData::_setIndex(int index) const {
    if (index == _lastIndex) return;
    double c[3], w[3];
    getBinCenters(index, c);
    _grid.getBinWidths(index, w);
    _zLast = c[2];
    transform(c[0], c[1], w[1], _zLast, _rLast, _muLast);
    _lastIndex = index;
}

double QuasarCorrelationData::getRadius(int index) const { _setIndex(index); return _rLast; }
double QuasarCorrelationData::getCosAngle(int index) const { _setIndex(index); return _muLast; }
double QuasarCorrelationData::getRedshift(int index) const { _setIndex(index); return _zLast; }

bool QuasarCorrelationData::extraCut(int index) const {
    double c[3];
    getBinCenters(index, c);
    double ll(c[0]), sep(c[1]);
    return ll < _llMin || ll > _llMax || sep < _sepMin || sep > _sepMax;
}

}  // namespace baofit
