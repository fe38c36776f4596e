"""ctypes binding of include/baofit_b200_host.h (sessions built from .ini text) -- plumbing."""
import ctypes as C

import numpy as np

from . import capi

_BOUND = False
BATCH_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_void_p)
RESID_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_void_p)

HOST_EXPORTS = ["bfh_last_error", "bfh_set_device", "bfh_set_num_threads", "bfh_session_create", "bfh_session_destroy", "bfh_npar", "bfh_ndata",
                "bfh_param_name", "bfh_param_info", "bfh_kept_indices", "bfh_kept_coordinates", "bfh_data_vector", "bfh_set_distortion_matrix", "bfh_evaluate_batch",
                "bfh_predict_batch", "bfh_fit", "bfh_scan_size", "bfh_parameter_scan", "bfh_toymc", "bfh_minimize",
                "bfh_lm_minimize", "bfh_dump_residuals", "bfh_dump_model", "bfh_mcmc", "bfh_nobservations",
                "bfh_resampling_size", "bfh_bootstrap_counts", "bfh_resample_data", "bfh_resample_fit", "bfh_write_samples",
                "bfh_toymc_refit", "bfh_resample_refit", "bfh_write_samples_refit", "bfh_save_data", "bfh_symmetric_eigenmodes", "bfh_compare_each", "bfh_gamma_q", "bfh_write_fit_outputs", "bfh_bootstrap_covariance", "bfh_set_toymc_save", "bfh_scale_zeff"]


def lib():
    global _BOUND
    L = capi.lib()
    if not _BOUND:
        L.bfh_last_error.restype = C.c_char_p
        L.bfh_session_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.bfh_session_destroy.argtypes = [C.c_void_p]
        L.bfh_npar.argtypes = [C.c_void_p]
        L.bfh_ndata.argtypes = [C.c_void_p]
        L.bfh_param_name.restype = C.c_char_p
        L.bfh_param_name.argtypes = [C.c_void_p, C.c_int]
        L.bfh_param_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_kept_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.bfh_kept_coordinates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_data_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_set_distortion_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.bfh_evaluate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.bfh_predict_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.bfh_fit.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                              C.POINTER(C.c_int), C.c_void_p]
        L.bfh_scan_size.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.bfh_parameter_scan.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_toymc.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_minimize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, BATCH_FN, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bfh_dump_residuals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.bfh_dump_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.bfh_mcmc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_void_p]
        L.bfh_lm_minimize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, RESID_FN,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        L.bfh_write_samples.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_double, C.c_long, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_double]
        L.bfh_toymc_refit.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_ulonglong, C.c_char_p, C.c_char_p] + [C.c_void_p] * 6
        L.bfh_resample_refit.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_long, C.c_long,
                                         C.c_char_p] + [C.c_void_p] * 6
        L.bfh_write_samples_refit.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_double, C.c_long, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                              C.c_void_p]
        L.bfh_save_data.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.bfh_write_fit_outputs.argtypes = [C.c_void_p, C.c_char_p]
        L.bfh_bootstrap_covariance.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.bfh_set_toymc_save.argtypes = [C.c_void_p, C.c_char_p]
        L.bfh_compare_each.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_symmetric_eigenmodes.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bfh_nobservations.argtypes = [C.c_void_p]
        L.bfh_resampling_size.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_long)]
        L.bfh_bootstrap_counts.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_long, C.c_void_p]
        L.bfh_resample_data.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_long, C.c_void_p,
                                        C.c_void_p]
        L.bfh_resample_fit.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_long, C.c_long,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        _BOUND = True
    return L


class HostError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[%d] %s" % (code, msg))
        self.code = code


def _check(rc):
    if rc != 0:
        raise HostError(rc, lib().bfh_last_error().decode())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def set_device(device):
    _check(lib().bfh_set_device(device))


def set_num_threads(n=0):
    """Host threads of the per-fit linear algebra (OpenMP); n <= 0 only queries. Returns the number in use."""
    return int(lib().bfh_set_num_threads(int(n)))


def scale_zeff(scale, dscale, gamma, dgamma, cov_scale_gamma, zref):
    """CorrelationAnalyzer::printScaleZEff: (zeff, scale at zeff, its error) from the fit results of a BAO scale and gamma-scale."""
    L = lib()
    L.bfh_scale_zeff.argtypes = [C.c_double] * 6 + [C.POINTER(C.c_double)] * 3
    z, s, ds = C.c_double(), C.c_double(), C.c_double()
    _check(L.bfh_scale_zeff(scale, dscale, gamma, dgamma, cov_scale_gamma, zref, C.byref(z), C.byref(s), C.byref(ds)))
    return z.value, s.value, ds.value


def symmetric_eigenmodes(matrix):
    """Eigenmodes of a dense symmetric matrix as project-modes-keep uses them: (values ascending, modes as rows)."""
    a = np.ascontiguousarray(matrix, dtype=np.float64)
    n = a.shape[0]
    values, vectors = np.empty(n), np.empty((n, n))
    _check(lib().bfh_symmetric_eigenmodes(n, a.ctypes.data, values.ctypes.data, vectors.ctypes.data))
    return values, vectors


def share_host_cores():
    """One process per GPU on one host: give this process its share of the cores. Launchers such as torchrun export
    OMP_NUM_THREADS=1 for every rank, which leaves the per-fit solves of a scan on a single thread."""
    import os
    local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return set_num_threads(max(1, (os.cpu_count() or 1) // max(1, local)))


class Session:
    """Model + data + CorrelationFitter built from INI text, like src/baofit.cc does."""

    def __init__(self, ini_text, base_dir="", extra_config=""):
        h = C.c_void_p()
        _check(lib().bfh_session_create(ini_text.encode(), base_dir.encode(), extra_config.encode(), C.byref(h)))
        self.h = h
        self.npar = lib().bfh_npar(h)
        self.ndata = lib().bfh_ndata(h)
        self.names = [lib().bfh_param_name(h, i).decode() for i in range(self.npar)]
        self.values = np.zeros(self.npar)
        self.errors = np.zeros(self.npar)
        fl = np.zeros(self.npar, dtype=np.int32)
        _check(lib().bfh_param_info(h, _p(self.values), _p(self.errors), _p(fl)))
        self.floating = fl.astype(bool)

    def close(self):
        if getattr(self, "h", None):
            lib().bfh_session_destroy(self.h)
            self.h = None

    __del__ = close

    def kept_indices(self):
        k = np.zeros(self.ndata, dtype=np.int32)
        _check(lib().bfh_kept_indices(self.h, _p(k)))
        return k

    def kept_coordinates(self):
        r, mu, z = np.zeros(self.ndata), np.zeros(self.ndata), np.zeros(self.ndata)
        _check(lib().bfh_kept_coordinates(self.h, _p(r), _p(mu), _p(z)))
        return r, mu, z

    def data_vector(self):
        d = np.zeros(self.ndata)
        _check(lib().bfh_data_vector(self.h, _p(d), None))
        return d

    def covariance(self):
        c = np.zeros((self.ndata, self.ndata))
        _check(lib().bfh_data_vector(self.h, None, _p(c)))
        return c

    def set_distortion_matrix(self, dense):
        d = np.ascontiguousarray(dense, dtype=np.float64)
        _check(lib().bfh_set_distortion_matrix(self.h, _p(d), d.shape[0]))

    def evaluate(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, self.npar)
        f = np.zeros(p.shape[0])
        _check(lib().bfh_evaluate_batch(self.h, _p(p), p.shape[0], _p(f)))
        return f

    def predict(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, self.npar)
        out = np.zeros((self.ndata, p.shape[0]))
        _check(lib().bfh_predict_batch(self.h, _p(p), p.shape[0], _p(out)))
        return out

    def dump_residuals(self, values, errors, path, gradients=False):
        v = np.ascontiguousarray(values, dtype=np.float64)
        e = np.ascontiguousarray(errors, dtype=np.float64)
        _check(lib().bfh_dump_residuals(self.h, _p(v), _p(e), path.encode(), 1 if gradients else 0))
        return np.loadtxt(path, ndmin=2)

    def dump_model(self, values, ndump, zdump=-1.0):
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros((ndump, 3))
        _check(lib().bfh_dump_model(self.h, _p(v), ndump, float(zdump), _p(out)))
        return out

    def mcmc(self, nchain, interval=10, nwalkers=256, seed=1966):
        """Fit, then nchain Metropolis samples [npar values, fval] from lock-step walkers."""
        out = np.zeros((nchain, self.npar + 1))
        _check(lib().bfh_mcmc(self.h, nchain, interval, nwalkers, seed, _p(out)))
        return out

    def fit(self, config=""):
        v, e = np.zeros(self.npar), np.zeros(self.npar)
        nf = int(self.floating.sum())
        cov = np.zeros((self.npar, self.npar))  # large enough for any config
        fval, status, nev = C.c_double(), C.c_int(), C.c_int()
        _check(lib().bfh_fit(self.h, config.encode(), _p(v), _p(e), C.byref(fval), C.byref(status), C.byref(nev), _p(cov)))
        return dict(values=v, errors=e, fval=fval.value, status=status.value, nevaluations=nev.value,
                    covariance=cov.ravel()[:nf * nf].reshape(nf, nf) if config == "" else None)

    def write_fit_outputs(self, prefix):
        """fit.chisq, fit.config, save.pars, save.pcov of the latest fit() without a config (src/baofit.cc:661-687)."""
        _check(lib().bfh_write_fit_outputs(self.h, prefix.encode()))

    def scan_size(self):
        n = C.c_long()
        _check(lib().bfh_scan_size(self.h, C.byref(n)))
        return n.value

    def parameter_scan(self, first=0, count=-1):
        total = self.scan_size()
        n = total - first if count < 0 else count
        chi2, pts = np.zeros(n), np.zeros((n, self.npar))
        st, best = np.zeros(n, dtype=np.int32), np.zeros(self.npar + 1)
        _check(lib().bfh_parameter_scan(self.h, first, count, _p(chi2), _p(pts), _p(st), _p(best)))
        return dict(chi2=chi2, points=pts, status=st, best_values=best[:-1], best_chi2=best[-1])

    def set_toymc_save(self, path):
        """toymc-save: the first realisation of the next study that starts at toy 0 goes to `path` (save-data's format)."""
        _check(lib().bfh_set_toymc_save(self.h, path.encode() if path else None))

    def toymc(self, first, count, seed=1966, toymc_config="", refit_config=""):
        fitted, chi2 = np.zeros((count, self.npar)), np.zeros(count)
        st = np.zeros(count, dtype=np.int32)
        if not toymc_config and not refit_config:
            _check(lib().bfh_toymc(self.h, first, count, seed, _p(fitted), _p(chi2), _p(st)))
            return dict(fitted=fitted, chi2=chi2, status=st)
        f2, c2, st2 = np.zeros((count, self.npar)), np.zeros(count), np.zeros(count, dtype=np.int32)
        _check(lib().bfh_toymc_refit(self.h, first, count, seed, toymc_config.encode(), refit_config.encode(), _p(fitted), _p(chi2),
                                     _p(st), _p(f2), _p(c2), _p(st2)))
        out = dict(fitted=fitted, chi2=chi2, status=st)
        if refit_config:
            out.update(refitted=f2, rechi2=c2, restatus=st2)
        return out


    def write_samples(self, path, best_values, best_errors, best_chi2, fitted, chi2, nsave=0, zsave=-1.0, refit=None):
        """The reference's SamplingOutput file (scan.dat / toymc.dat ...): header, input fit, one line per sample.
        refit = (best2_values, best2_errors, best2_chi2, refitted, rechi2) adds the columns of a refit pass."""
        bv = np.ascontiguousarray(best_values, dtype=np.float64)
        be = np.ascontiguousarray(best_errors, dtype=np.float64)
        f = np.ascontiguousarray(fitted, dtype=np.float64).reshape(-1, self.npar)
        c = np.ascontiguousarray(chi2, dtype=np.float64)
        if refit is None:
            _check(lib().bfh_write_samples(self.h, path.encode(), _p(bv), _p(be), float(best_chi2), len(c), _p(f), _p(c), nsave, zsave))
            return
        bv2, be2 = (np.ascontiguousarray(x, dtype=np.float64) for x in refit[:2])
        f2 = np.ascontiguousarray(refit[3], dtype=np.float64).reshape(-1, self.npar)
        c2 = np.ascontiguousarray(refit[4], dtype=np.float64)
        _check(lib().bfh_write_samples_refit(self.h, path.encode(), _p(bv), _p(be), float(best_chi2), len(c), _p(f), _p(c), nsave, zsave,
                                             _p(bv2), _p(be2), float(refit[2]), _p(f2), _p(c2)))

    def save_data(self, data_path=None, icov_path=None):
        _check(lib().bfh_save_data(self.h, data_path.encode() if data_path else None, icov_path.encode() if icov_path else None))

    def bootstrap_covariance(self, trials, nbins, seed=1966, icov_path=None):
        """CorrelationAnalyzer::estimateCombinedCovariance: (bins, covariance over them, positive definite?)."""
        cov, bins, ok = np.zeros((nbins, nbins)), np.zeros(nbins, dtype=np.int32), C.c_int()
        _check(lib().bfh_bootstrap_covariance(self.h, trials, seed, icov_path.encode() if icov_path else None, _p(cov), nbins, _p(bins),
                                              C.byref(ok)))
        return bins, cov, bool(ok.value)

    def compare_each(self, path=None, finalized=False):
        """CorrelationAnalyzer::compareEach: every observation against the combined data; returns prob, chi2, log|C|/n."""
        n = self.nobservations
        prob, chi2, logdet = np.zeros(n), np.zeros(n), np.zeros(n)
        _check(lib().bfh_compare_each(self.h, path.encode() if path else None, 1 if finalized else 0, _p(prob), _p(chi2), _p(logdet)))
        return prob, chi2, logdet

    # ---- jackknife / bootstrap / fit-each over the observations of a plateroot/platelist session ----
    @property
    def nobservations(self):
        return lib().bfh_nobservations(self.h)

    def resampling_size(self, kind, n=0):
        out = C.c_long()
        _check(lib().bfh_resampling_size(self.h, kind.encode(), n, C.byref(out)))
        return out.value

    def bootstrap_counts(self, trial, size=0, seed=1966):
        counts = np.zeros(self.nobservations, dtype=np.int32)
        _check(lib().bfh_bootstrap_counts(self.h, size, seed, trial, _p(counts)))
        return counts

    def resample_data(self, kind, seqno, n=0, size=0, fix_covariance=False, seed=1966):
        d, c = np.zeros(self.ndata), np.zeros((self.ndata, self.ndata))
        _check(lib().bfh_resample_data(self.h, kind.encode(), n, size, int(fix_covariance), seed, seqno, _p(d), _p(c)))
        return d, c

    def resample_fit(self, kind, n=0, size=0, fix_covariance=False, seed=1966, first=0, count=-1, refit_config=""):
        if count < 0:
            count = self.resampling_size(kind, n) - first
        fitted, chi2 = np.zeros((count, self.npar)), np.zeros(count)
        st = np.zeros(count, dtype=np.int32)
        f2, c2, st2 = np.zeros((count, self.npar)), np.zeros(count), np.zeros(count, dtype=np.int32)
        _check(lib().bfh_resample_refit(self.h, kind.encode(), n, size, int(fix_covariance), seed, first, count, refit_config.encode(),
                                        _p(fitted), _p(chi2), _p(st), _p(f2), _p(c2), _p(st2)))
        out = dict(fitted=fitted, chi2=chi2, status=st)
        if refit_config:
            out.update(refitted=f2, rechi2=c2, restatus=st2)
        return out


def minimize(fn, values, errors, floating):
    """Runs the library's minimiser over a Python batch objective fn(points[B, npar]) -> f[B]."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    errors = np.ascontiguousarray(errors, dtype=np.float64)
    fl = np.ascontiguousarray(np.asarray(floating).astype(np.int32))
    npar = len(values)

    def cb(pts, B, f, user):
        p = np.ctypeslib.as_array(pts, shape=(B, npar))
        out = np.ctypeslib.as_array(f, shape=(B,))
        out[:] = fn(p.copy())

    cfn = BATCH_FN(cb)
    ov, oe = np.zeros(npar), np.zeros(npar)
    fval, status, nev = C.c_double(), C.c_int(), C.c_int()
    _check(lib().bfh_minimize(npar, _p(values), _p(errors), _p(fl), cfn, None, _p(ov), _p(oe), C.byref(fval),
                              C.byref(status), C.byref(nev)))
    return dict(values=ov, errors=oe, fval=fval.value, status=status.value, nevaluations=nev.value)


def lm_minimize(resid_fn, nres, values, errors, floating, prior_centre=None, prior_sigma=None):
    """Runs the library's Levenberg-Marquardt driver over a Python residual function
    resid_fn(points[B, npar]) -> resid[B, nres]; objective 0.5 |resid|^2 + Gaussian priors."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    errors = np.ascontiguousarray(errors, dtype=np.float64)
    fl = np.ascontiguousarray(np.asarray(floating).astype(np.int32))
    npar = len(values)
    pc = np.zeros(npar) if prior_centre is None else np.ascontiguousarray(prior_centre, dtype=np.float64)
    ps = np.zeros(npar) if prior_sigma is None else np.ascontiguousarray(prior_sigma, dtype=np.float64)

    def cb(pts, B, out, user):
        p = np.ctypeslib.as_array(pts, shape=(B, npar))
        o = np.ctypeslib.as_array(out, shape=(B, nres))
        o[:] = resid_fn(p.copy())

    cfn = RESID_FN(cb)
    ov = np.zeros(npar)
    fval, status, nev, nit = C.c_double(), C.c_int(), C.c_int(), C.c_int()
    _check(lib().bfh_lm_minimize(npar, _p(values), _p(errors), _p(fl), _p(pc), _p(ps), nres, cfn, None, _p(ov), C.byref(fval),
                                 C.byref(status), C.byref(nev), C.byref(nit)))
    return dict(values=ov, fval=fval.value, status=status.value, nevaluations=nev.value, niterations=nit.value)
