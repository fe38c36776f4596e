"""ctypes loader of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package baofit_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h")]
    src.append(os.path.join(_HERE, "..", "include", "baofit_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    so2 = os.path.join(_HERE, "libcpubaseline.so")
    src2 = src + [os.path.join(_HERE, "cpu_baseline.c")]
    if force or not os.path.exists(so2) or any(os.path.getmtime(s) > os.path.getmtime(so2) for s in src2):
        subprocess.check_call(["make", "-C", _HERE, "libcpubaseline.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_void_p]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_data_create.restype = C.c_void_p
        L.orc_data_create.argtypes = [C.c_void_p]
        L.orc_data_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_cspline_coeffs.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
        L.orc_cspline_eval.restype = C.c_double
        L.orc_cspline_eval.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, C.c_double]
        L.orc_legendre.restype = C.c_double
        L.orc_legendre.argtypes = [C.c_int, C.c_double]
        L.orc_evaluate.restype = C.c_double
        L.orc_evaluate.argtypes = [C.c_void_p, C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_double,
                                   C.c_int, c_int_p]
        L.orc_evaluate_multipole.restype = C.c_double
        L.orc_evaluate_multipole.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_int, C.c_double, C.c_int]
        L.orc_predict.argtypes = [C.c_void_p, C.c_void_p, c_double_p, c_double_p]
        L.orc_undistorted.argtypes = [C.c_void_p, C.c_void_p, c_double_p, c_double_p]
        L.orc_chi2.restype = C.c_double
        L.orc_chi2.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_chi2_batch.argtypes = [C.c_void_p, C.c_void_p, c_double_p, C.c_int, c_double_p, c_double_p,
                                     c_int_p, c_double_p, C.c_int, C.c_int]
        L.orc_kspace_distortion.restype = C.c_double
        L.orc_kspace_distortion.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_int, C.c_double, C.c_double]
        L.orc_transform.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p]
        L.orc_transform_sub.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_int, c_double_p]
        L.orc_transform_nr.argtypes = [C.c_void_p]
        L.orc_transform_r0.restype = C.c_double
        L.orc_transform_r0.argtypes = [C.c_void_p]
        L.orc_transform_dr.restype = C.c_double
        L.orc_transform_dr.argtypes = [C.c_void_p]
        L.orc_hankel_table.restype = C.c_double
        L.orc_hankel_table.argtypes = [C.c_int, c_double_p, c_double_p, C.c_int, C.c_double]
        L.orc_fft_distortion_power.restype = C.c_double
        L.orc_fft_distortion_power.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_int, C.c_double, C.c_double]
        L.orc_fft_plane_dims.argtypes = [C.c_void_p, c_int_p, c_int_p]
        L.orc_fft_planes.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class OracleModel:
    def __init__(self, model):
        """model: baofit_b200.workloads.Model (or anything with .desc = ModelDesc)."""
        self._model = model
        self.h = lib().orc_model_create(C.addressof(model.desc))
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self.npar = model.desc.npar

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_destroy(self.h)
            self.h = None

    def evaluate(self, params, r, mu, z, index=-1, data=None):
        p = np.ascontiguousarray(params, dtype=np.float64)
        st = C.c_int(0)
        v = lib().orc_evaluate(self.h, data.h if data else None, _dp(p), r, mu, z, index, C.byref(st))
        return v, st.value

    def evaluate_multipole(self, params, r, ell, z, nmu=16):
        p = np.ascontiguousarray(params, dtype=np.float64)
        return lib().orc_evaluate_multipole(self.h, _dp(p), r, ell, z, nmu)

    def transform(self, params, z_trigger, jstep=1):
        """xi_ell(r_j) tables [comp][ell][j]; with jstep > 1 only every jstep-th radius is filled."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        nr = lib().orc_transform_nr(self.h)
        out = np.full((2, 3, nr), np.nan)
        if lib().orc_transform_sub(self.h, _dp(p), z_trigger, jstep, _dp(out)):
            raise RuntimeError(lib().orc_last_error().decode())
        r = lib().orc_transform_r0(self.h) + lib().orc_transform_dr(self.h) * np.arange(nr)
        return r, out

    def fft_distortion_power(self, params, z, component, k, muk):
        """D(k, mu_k) P_comp(k) of the FFT model (BaoKSpaceFftCorrelationModel.cc:120-153)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        return lib().orc_fft_distortion_power(self.h, _dp(p), z, component, k, muk)

    def fft_planes(self, params, z_trigger):
        """xi_comp on the z = 0 plane of the FFT box: array [comp][iy][ix]."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        npx, npy = C.c_int(0), C.c_int(0)
        if lib().orc_fft_plane_dims(self.h, C.byref(npx), C.byref(npy)):
            raise RuntimeError("not an FFT model")
        out = np.zeros((2, npy.value, npx.value))
        if lib().orc_fft_planes(self.h, _dp(p), z_trigger, _dp(out)):
            raise RuntimeError(lib().orc_last_error().decode())
        return out

    def kspace_distortion(self, params, z, component, k, muk):
        p = np.ascontiguousarray(params, dtype=np.float64)
        return lib().orc_kspace_distortion(self.h, _dp(p), z, component, k, muk)


class OracleData:
    def __init__(self, data):
        self._data = data
        self.h = lib().orc_data_create(C.addressof(data.desc))
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self.n_data = data.desc.n_data

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_data_destroy(self.h)
            self.h = None


def predict(model, data, params):
    p = np.ascontiguousarray(params, dtype=np.float64)
    pred = np.zeros(data.n_data)
    st = lib().orc_predict(model.h, data.h, _dp(p), _dp(pred))
    return pred, st


def undistorted(model, data, params, n_grid):
    p = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros(n_grid)
    st = lib().orc_undistorted(model.h, data.h, _dp(p), _dp(out))
    if st < 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return out, st


def chi2(data, pred, data_override=None):
    pr = np.ascontiguousarray(pred, dtype=np.float64)
    ov = None if data_override is None else np.ascontiguousarray(data_override, dtype=np.float64)
    return lib().orc_chi2(data.h, _dp(pr), _dp(ov) if ov is not None else None)


def chi2_batch(model, data, params, data_override=None, want_pred=False, nthreads=1):
    p = np.ascontiguousarray(params, dtype=np.float64)
    B = p.shape[0]
    out = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    pred = np.zeros((data.n_data, B)) if want_pred else None
    ov = None if data_override is None else np.ascontiguousarray(data_override, dtype=np.float64)
    lib().orc_chi2_batch(model.h, data.h, _dp(p), B, _dp(ov) if ov is not None else None, _dp(out),
                         status.ctypes.data_as(c_int_p), _dp(pred) if want_pred else None, B, nthreads)
    return (out, status, pred) if want_pred else (out, status)


def cspline(x, y, xq):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    c = np.zeros_like(x)
    lib().orc_cspline_coeffs(len(x), _dp(x), _dp(y), _dp(c))
    return np.array([lib().orc_cspline_eval(len(x), _dp(x), _dp(y), _dp(c), float(q)) for q in np.atleast_1d(xq)])


def hankel_table(k, pk, ell, r):
    k = np.ascontiguousarray(k, dtype=np.float64)
    pk = np.ascontiguousarray(pk, dtype=np.float64)
    return lib().orc_hankel_table(len(k), _dp(k), _dp(pk), ell, float(r))


# --------------------------------------------------------------------------------------------
# reference-shaped CPU timing baseline (oracle/cpu_baseline.c); bench.py only
# --------------------------------------------------------------------------------------------
_CPUB = None


def cpub_lib():
    global _CPUB
    if _CPUB is None:
        so = os.path.join(_HERE, "libcpubaseline.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_void_p]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_data_create.restype = C.c_void_p
        L.orc_data_create.argtypes = [C.c_void_p]
        L.orc_data_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_transform_nr.argtypes = [C.c_void_p]
        L.cpub_create.restype = C.c_void_p
        L.cpub_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.cpub_destroy.argtypes = [C.c_void_p]
        L.cpub_chi2_batch.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p, c_int_p, C.c_int, C.POINTER(C.c_longlong)]
        _CPUB = L
    return _CPUB


class CpuBaseline:
    """chi2 of parameter batches on host cores with the reference's control flow (per-bin loop, transform only when a
    parameter D(k, mu) reads changed) and a transform of the reference's cost (weights from the product's host set-up,
    passed in as [6][nr][nk]; None for r-space models)."""

    def __init__(self, model, data, weights=None):
        L = cpub_lib()
        self.model, self.data = model, data  # keep the descriptors (and their arrays) alive
        self.m = L.orc_model_create(C.addressof(model.desc))
        self.d = L.orc_data_create(C.addressof(data.desc))
        if not self.m or not self.d:
            raise RuntimeError(L.orc_last_error().decode())
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.float64)
            nr, nk = w.shape[-2], w.shape[-1]
            self.h = L.cpub_create(self.m, self.d, nr, nk, _dp(w))
        else:
            self.h = L.cpub_create(self.m, self.d, 0, 0, None)
        if not self.h:
            raise RuntimeError(L.orc_last_error().decode())
        self.npar = model.desc.npar

    def chi2_batch(self, params, nthreads=1):
        L = cpub_lib()
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, self.npar)
        B = p.shape[0]
        chi2, status, ntr = np.zeros(B), np.zeros(B, dtype=np.int32), C.c_longlong(0)
        L.cpub_chi2_batch(self.h, _dp(p), B, _dp(chi2), status.ctypes.data_as(c_int_p), int(nthreads), C.byref(ntr))
        return chi2, status, int(ntr.value)

    def close(self):
        L = cpub_lib()
        if getattr(self, "h", None):
            L.cpub_destroy(self.h)
            L.orc_data_destroy(self.d)
            L.orc_model_destroy(self.m)
            self.h = None

    __del__ = close
