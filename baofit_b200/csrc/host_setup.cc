// Host-side, one-time preprocessing: spline coefficients, Filon weights of the k-space
// transform, Cholesky factorisation of the covariance. None of this runs per evaluation.
#include "host_setup.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <thread>

namespace bf {

// ------------------------------------------------------------------------------------------
// natural cubic spline, the algorithm of gsl_interp_cspline (what likely::Interpolator
// "cspline" wraps; call sites baofit/BaoCorrelationModel.cc:56-67)
// ------------------------------------------------------------------------------------------
void cspline_c(int n, const double *x, const double *y, double *c) {
    c[0] = 0.0;
    c[n - 1] = 0.0;
    int m = n - 2;
    if (m <= 0) return;
    std::vector<double> diag(m), off(m), g(m);
    for (int i = 0; i < m; ++i) {
        double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        double d0 = y[i + 1] - y[i], d1 = y[i + 2] - y[i + 1];
        off[i] = h1;
        diag[i] = 2.0 * (h1 + h0);
        g[i] = 3.0 * (d1 / h1 - d0 / h0);
    }
    for (int i = 1; i < m; ++i) {
        double w = off[i - 1] / diag[i - 1];
        diag[i] -= w * off[i - 1];
        g[i] -= w * g[i - 1];
    }
    c[m] = g[m - 1] / diag[m - 1];
    for (int i = m - 2; i >= 0; --i) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
}

void cspline_interval_coefs(int n, const double *x, const double *y, double *coef4) {
    std::vector<double> c(n);
    cspline_c(n, x, y, c.data());
    for (int j = 0; j + 1 < n; ++j) {
        double dx = x[j + 1] - x[j], dy = y[j + 1] - y[j];
        coef4[4 * j + 0] = y[j];
        coef4[4 * j + 1] = dy / dx - dx * (c[j + 1] + 2.0 * c[j]) / 3.0;
        coef4[4 * j + 2] = c[j];
        coef4[4 * j + 3] = (c[j + 1] - c[j]) / (3.0 * dx);
    }
}

double cspline_eval_host(int n, const double *x, const double *y, const double *c, double xq) {
    if (xq <= x[0]) return y[0];
    if (xq >= x[n - 1]) return y[n - 1];
    int lo = int(std::upper_bound(x, x + n, xq) - x) - 1;
    if (lo > n - 2) lo = n - 2;
    double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo], t = xq - x[lo];
    double b = dy / dx - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
    double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
    return y[lo] + t * (b + t * (c[lo] + t * d));
}

void gauss_legendre_01(int n, double *x, double *w) {
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < n; ++i) {
        long double t = cosl(pi * (i + 0.75L) / (n + 0.5L)), pp = 0, p1 = 0;
        for (int it = 0; it < 100; ++it) {
            long double p2 = 0;
            p1 = 1;
            for (int j = 0; j < n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j + 1.0L) * t * p2 - j * p3) / (j + 1.0L);
            }
            pp = n * (t * p1 - p2) / (t * t - 1.0L);
            long double dt = p1 / pp;
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        x[i] = double(0.5L * (1.0L - t));
        w[i] = double(1.0L / ((1.0L - t * t) * pp * pp));
    }
}

// ------------------------------------------------------------------------------------------
// Filon weights
// ------------------------------------------------------------------------------------------
namespace {
typedef long double ld;
const ld kPi = 3.141592653589793238462643383279502884L;

// sine integral Si(x), x > 0, to ~1e-18
ld sine_integral(ld x) {
    if (x <= 4.0L) {
        ld term = x, sum = x, x2 = x * x;
        for (int k = 1; k < 60; ++k) {
            term *= -x2 / ((2.0L * k) * (2.0L * k + 1.0L));
            ld add = term / (2.0L * k + 1.0L);
            sum += add;
            if (fabsl(add) < 1e-21L * fabsl(sum)) break;
        }
        return sum;
    }
    if (x >= 50.0L) {
        // asymptotic auxiliary functions f, g: Si = pi/2 - f cos x - g sin x
        ld ix2 = 1.0L / (x * x), f = 0, g = 0, tf = 1.0L, tg = 1.0L;
        for (int k = 0; k < 30; ++k) {
            // f ~ (1/x) sum (-1)^k (2k)!/x^2k ; g ~ (1/x^2) sum (-1)^k (2k+1)!/x^2k
            f += tf;
            g += tg;
            ld nf = -tf * (2.0L * k + 1.0L) * (2.0L * k + 2.0L) * ix2;
            ld ng = -tg * (2.0L * k + 2.0L) * (2.0L * k + 3.0L) * ix2;
            if (fabsl(nf) > fabsl(tf) || fabsl(nf) < 1e-22L) break;
            tf = nf;
            tg = ng;
        }
        f /= x;
        g *= ix2;
        return kPi / 2 - f * cosl(x) - g * sinl(x);
    }
    // modified Lentz continued fraction for E1(ix)
    typedef std::complex<ld> cx;
    cx b(1.0L, x), c(1e300L, 0.0L), d = cx(1.0L, 0.0L) / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        ld a = -(ld)i * i;
        b += cx(2.0L, 0.0L);
        d = cx(1.0L, 0.0L) / (a * d + b);
        c = b + a / c;
        cx del = c * d;
        h *= del;
        if (fabsl(del.real() - 1.0L) + fabsl(del.imag()) < 1e-20L) break;
    }
    h = cx(cosl(x), -sinl(x)) * h;
    return kPi / 2 + h.imag();
}

// antiderivatives A_ln(x) of x^n j_l(x), n = 2,3, l = 0,2,4 (verified against mpmath)
struct Anti {
    ld a[3][2];
};
Anti antiderivatives(ld x) {
    Anti r;
    ld s = sinl(x), c = cosl(x), si = sine_integral(x), x2 = x * x;
    r.a[0][0] = s - x * c;
    r.a[0][1] = 2 * x * s - (x2 - 2) * c;
    r.a[1][0] = 3 * si - 4 * s + x * c;
    r.a[1][1] = (x2 - 8) * c - 5 * x * s;
    r.a[2][0] = 7.5L * si + (11 - 52.5L / x2) * s + (52.5L / x - x) * c;
    r.a[2][1] = (12 * x - 105 / x) * s + (57 - x2) * c;
    return r;
}

void sph_bessel_series(ld x, ld j[3]) {  // x < ~2.5
    ld x2 = x * x;
    ld t0 = 1, t2 = x2 / 15, t4 = x2 * x2 / 945, s0 = 0, s2 = 0, s4 = 0;
    for (int k = 0; k < 40; ++k) {
        s0 += t0;
        s2 += t2;
        s4 += t4;
        t0 *= -x2 / ((2.0L * k + 2) * (2.0L * k + 3));
        t2 *= -x2 / ((2.0L * k + 2) * (2.0L * k + 7));
        t4 *= -x2 / ((2.0L * k + 2) * (2.0L * k + 11));
        if (fabsl(t0) < 1e-22L) break;
    }
    j[0] = s0;
    j[1] = s2;
    j[2] = s4;
}

const ld glx8[8] = {-0.9602898564975362316835609L, -0.7966664774136267395915539L, -0.5255324099163289858177390L,
                    -0.1834346424956498049394761L, 0.1834346424956498049394761L,  0.5255324099163289858177390L,
                    0.7966664774136267395915539L,  0.9602898564975362316835609L};
const ld glw8[8] = {0.1012285362903762591525314L, 0.2223810344533744705443560L, 0.3137066458778872873379622L,
                    0.3626837833783619829651504L, 0.3626837833783619829651504L, 0.3137066458778872873379622L,
                    0.2223810344533744705443560L, 0.1012285362903762591525314L};
}  // namespace

void build_hankel_weights(int n_pk, const double *pk_k, const double *pk_fid, const double *pk_nw, double rmin,
                          double rmax, double dilmin, double dilmax, int samples_per_decade, int fine_sub,
                          HankelSetup &out) {
    // native tables -> natural cubic splines in ln k
    std::vector<double> lnk(n_pk), cf(n_pk), cn(n_pk);
    for (int i = 0; i < n_pk; ++i) lnk[i] = std::log(pk_k[i]);
    cspline_c(n_pk, lnk.data(), pk_fid, cf.data());
    cspline_c(n_pk, lnk.data(), pk_nw, cn.data());
    double klo = pk_k[0], khi = pk_k[n_pk - 1];
    // coarse nodes for the multipoles of D (baofit/BaoKSpaceCorrelationModel.cc:140-141)
    int nk = (int)std::ceil(std::log10(khi / klo) * samples_per_decade);
    out.nk = nk;
    out.kc.resize(nk);
    for (int i = 0; i < nk; ++i) out.kc[i] = klo * std::pow(khi / klo, (double)i / (nk - 1));
    out.kc[nk - 1] = khi;
    out.pk_kc.resize(2 * (size_t)nk);
    for (int i = 0; i < nk; ++i) {
        double lk = std::log(out.kc[i]);
        double pn = cspline_eval_host(n_pk, lnk.data(), pk_nw, cn.data(), lk);
        double pf = cspline_eval_host(n_pk, lnk.data(), pk_fid, cf.data(), lk);
        out.pk_kc[i] = pf - pn;
        out.pk_kc[nk + i] = pn;
    }
    // fine nodes
    int nf = (n_pk - 1) * fine_sub + 1;
    std::vector<double> kf(nf), pf_pk(nf), pf_nw(nf);
    for (int i = 0; i < n_pk - 1; ++i)
        for (int s = 0; s < fine_sub; ++s) {
            int f = i * fine_sub + s;
            double k = pk_k[i] + (pk_k[i + 1] - pk_k[i]) * s / fine_sub;
            kf[f] = k;
            if (s == 0) {
                pf_nw[f] = pk_nw[i];
                pf_pk[f] = pk_fid[i] - pk_nw[i];
            } else {
                double lk = std::log(k);
                double pn = cspline_eval_host(n_pk, lnk.data(), pk_nw, cn.data(), lk);
                pf_nw[f] = pn;
                pf_pk[f] = cspline_eval_host(n_pk, lnk.data(), pk_fid, cf.data(), lk) - pn;
            }
        }
    kf[nf - 1] = pk_k[n_pk - 1];
    pf_nw[nf - 1] = pk_nw[n_pk - 1];
    pf_pk[nf - 1] = pk_fid[n_pk - 1] - pk_nw[n_pk - 1];
    // radial nodes (baofit/BaoKSpaceCorrelationModel.cc:160-166)
    out.rlo = rmin * dilmin;
    out.rhi = rmax * dilmax;
    out.nr = (int)std::ceil(out.rhi - out.rlo);
    out.r0 = out.rlo;
    out.dr = (out.rhi - out.rlo) / (out.nr - 1);
    int nr = out.nr;
    // Lagrange stencil of every fine node on the coarse grid (uniform in ln k)
    std::vector<int> st_i(nf);
    std::vector<double> st_w(4 * (size_t)nf);
    {
        double u0 = std::log(out.kc[0]), h = (std::log(out.kc[nk - 1]) - u0) / (nk - 1);
        for (int f = 0; f < nf; ++f) {
            double xx = (std::log(kf[f]) - u0) / h;
            int i = (int)std::floor(xx);
            if (i < 1) i = 1;
            if (i > nk - 3) i = nk - 3;
            double t = xx - i;
            st_i[f] = i - 1;
            st_w[4 * f + 0] = -t * (t - 1) * (t - 2) / 6;
            st_w[4 * f + 1] = (t + 1) * (t - 1) * (t - 2) / 2;
            st_w[4 * f + 2] = -(t + 1) * t * (t - 2) / 2;
            st_w[4 * f + 3] = (t + 1) * t * (t - 1) / 6;
        }
    }
    out.weights.assign((size_t)6 * nr * nk, 0.0);
    const ld pref[3] = {1.0L / (2 * kPi * kPi), -1.0L / (2 * kPi * kPi), 1.0L / (2 * kPi * kPi)};

    auto do_radius = [&](int j) {
        double rj = out.r0 + j * out.dr;  // same radial node value the kernels use
        ld r = (ld)rj;
        // hat-function integrals H[l][f]
        std::vector<ld> H(3 * (size_t)nf, 0.0L);
        std::vector<Anti> A(nf);
        std::vector<char> haveA(nf, 0);
        for (int f = 0; f < nf; ++f) {
            ld x = (ld)kf[f] * r;
            if (x >= 1.0L) {
                A[f] = antiderivatives(x);
                haveA[f] = 1;
            }
        }
        ld r3 = r * r * r, r4 = r3 * r;
        for (int f = 0; f + 1 < nf; ++f) {
            ld a = kf[f], b = kf[f + 1], h = b - a;
            ld I0[3], I1[3];
            if (b * r < 2.0L || !haveA[f]) {
                for (int l = 0; l < 3; ++l) I0[l] = I1[l] = 0;
                for (int q = 0; q < 8; ++q) {
                    ld k = a + 0.5L * h * (1 + glx8[q]), jl[3];
                    sph_bessel_series(k * r, jl);
                    ld w = 0.5L * h * glw8[q] * k * k, t1 = (k - a) / h;
                    for (int l = 0; l < 3; ++l) {
                        I0[l] += w * (1 - t1) * jl[l];
                        I1[l] += w * t1 * jl[l];
                    }
                }
            } else {
                for (int l = 0; l < 3; ++l) {
                    ld M2 = (A[f + 1].a[l][0] - A[f].a[l][0]) / r3;
                    ld M3 = (A[f + 1].a[l][1] - A[f].a[l][1]) / r4;
                    I1[l] = (M3 - a * M2) / h;
                    I0[l] = (b * M2 - M3) / h;
                }
            }
            for (int l = 0; l < 3; ++l) {
                H[(size_t)l * nf + f] += I0[l];
                H[(size_t)l * nf + f + 1] += I1[l];
            }
        }
        for (int c = 0; c < 2; ++c) {
            const std::vector<double> &pf = c == 0 ? pf_pk : pf_nw;
            for (int l = 0; l < 3; ++l) {
                std::vector<ld> acc(nk, 0.0L);
                for (int f = 0; f < nf; ++f) {
                    ld g = pref[l] * H[(size_t)l * nf + f] * (ld)pf[f];
                    if (g == 0) continue;
                    int i0 = st_i[f];
                    for (int s = 0; s < 4; ++s) acc[i0 + s] += g * (ld)st_w[4 * f + s];
                }
                double *dst = &out.weights[(((size_t)c * 3 + l) * nr + j) * nk];
                for (int i = 0; i < nk; ++i) dst[i] = (double)acc[i];
            }
        }
    };
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t] {
            for (int j = (int)t; j < nr; j += (int)nt) do_radius(j);
        });
    for (auto &th : pool) th.join();
}

// ------------------------------------------------------------------------------------------
// dense SPD factorisation
// ------------------------------------------------------------------------------------------
bool cholesky_lower(int n, double *a) {
    // right-looking, row-major; accumulates in long double for a well-rounded factor
    for (int j = 0; j < n; ++j) {
        long double s = a[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) s -= (long double)a[(size_t)j * n + k] * a[(size_t)j * n + k];
        if (!(s > 0)) return false;
        double ljj = (double)sqrtl(s);
        a[(size_t)j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            const double *ri = a + (size_t)i * n, *rj = a + (size_t)j * n;
            long double t = ri[j];
            for (int k = 0; k < j; ++k) t -= (long double)ri[k] * rj[k];
            a[(size_t)i * n + j] = (double)(t / ljj);
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) a[(size_t)i * n + j] = 0.0;
    return true;
}

void invert_lower(int n, const double *L, double *out) {
    // column by column forward substitution: L X = I
    std::fill(out, out + (size_t)n * n, 0.0);
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([=] {
            std::vector<long double> col(n);
            for (int j = (int)t; j < n; j += (int)nt) {
                for (int i = 0; i < j; ++i) col[i] = 0;
                col[j] = 1.0L / L[(size_t)j * n + j];
                for (int i = j + 1; i < n; ++i) {
                    long double s = 0;
                    const double *ri = L + (size_t)i * n;
                    for (int k = j; k < i; ++k) s -= (long double)ri[k] * col[k];
                    col[i] = s / ri[i];
                }
                for (int i = j; i < n; ++i) out[(size_t)i * n + j] = (double)col[i];
            }
        });
    for (auto &th : pool) th.join();
}

}  // namespace bf
