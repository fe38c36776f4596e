"""ctypes binding of the C ABI (include/baofit_b200.h) -- plumbing for tests and bench.py.

Every call goes through libbaofit_b200.so; there is no Python or CPU implementation behind it.
If the library is missing or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbaofit_b200.so")
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

EXPORTS = [
    "bf_last_error", "bf_version", "bf_ctx_create", "bf_ctx_destroy", "bf_ctx_launch_count",
    "bf_ctx_synchronize", "bf_ctx_stream", "bf_ctx_set_profiling", "bf_ctx_get_profile", "bf_model_create", "bf_model_destroy", "bf_model_npar",
    "bf_data_create", "bf_data_destroy", "bf_data_ndata", "bf_eval_batch", "bf_chi2_batch",
    "bf_eval_batch_dev", "bf_chi2_batch_dev", "bf_chi2_batch_submit", "bf_chi2_batch_wait", "bf_eval_undistorted_batch", "bf_transform_batch",
    "bf_model_transform_nr", "bf_model_transform_r0", "bf_model_transform_dr", "bf_sample_toys",
    "bf_fft_planes_batch", "bf_model_fft_plane_dims", "bf_kspace_transform_weights",
    "bf_toys_create", "bf_toys_destroy", "bf_toys_download", "bf_chi2_batch_toys", "bf_gram_batch", "bf_gram_probes", "bf_normal_probes", "bf_pinned_alloc", "bf_pinned_free",
]


class BaofitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[%d] %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libbaofit_b200.so is not built: run python -c 'import __graft_entry__ as g; g.build()'")
        L = C.CDLL(LIB_PATH)
        L.bf_last_error.restype = C.c_char_p
        L.bf_version.restype = C.c_char_p
        L.bf_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.bf_ctx_destroy.argtypes = [C.c_void_p]
        L.bf_ctx_launch_count.restype = C.c_longlong
        L.bf_ctx_launch_count.argtypes = [C.c_void_p]
        L.bf_ctx_synchronize.argtypes = [C.c_void_p]
        L.bf_ctx_stream.restype = C.c_void_p
        L.bf_ctx_stream.argtypes = [C.c_void_p]
        L.bf_ctx_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.bf_ctx_get_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), c_double_p,
                                         C.POINTER(C.c_longlong)]
        L.bf_model_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bf_model_destroy.argtypes = [C.c_void_p]
        L.bf_model_npar.argtypes = [C.c_void_p]
        L.bf_data_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bf_data_destroy.argtypes = [C.c_void_p]
        L.bf_data_ndata.argtypes = [C.c_void_p]
        L.bf_eval_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.bf_chi2_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bf_eval_batch_dev.argtypes = L.bf_eval_batch.argtypes
        L.bf_chi2_batch_dev.argtypes = L.bf_chi2_batch.argtypes
        L.bf_chi2_batch_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.bf_chi2_batch_wait.argtypes = [C.c_void_p, C.c_int]
        L.bf_eval_undistorted_batch.argtypes = L.bf_eval_batch.argtypes
        L.bf_transform_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
        L.bf_model_transform_nr.argtypes = [C.c_void_p]
        L.bf_model_transform_r0.restype = C.c_double
        L.bf_model_transform_r0.argtypes = [C.c_void_p]
        L.bf_model_transform_dr.restype = C.c_double
        L.bf_model_transform_dr.argtypes = [C.c_void_p]
        L.bf_sample_toys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_longlong, C.c_int,
                                     C.c_double, C.c_void_p]
        L.bf_toys_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_longlong, C.c_int, C.c_double,
                                     C.POINTER(C.c_void_p)]
        L.bf_toys_destroy.argtypes = [C.c_void_p]
        L.bf_toys_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bf_chi2_batch_toys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.bf_gram_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        L.bf_normal_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.bf_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.bf_pinned_free.argtypes = [C.c_void_p]
        L.bf_gram_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.bf_fft_planes_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.bf_model_fft_plane_dims.argtypes = [C.c_void_p, c_int_p, c_int_p]
        L.bf_kspace_transform_weights.argtypes = [C.c_void_p, c_int_p, c_int_p, C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise BaofitError(rc, lib().bf_last_error().decode())


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().bf_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().bf_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def launches(self):
        return lib().bf_ctx_launch_count(self.h)

    def synchronize(self):
        _check(lib().bf_ctx_synchronize(self.h))

    @property
    def stream(self):
        return lib().bf_ctx_stream(self.h)

    def set_profiling(self, enable):
        _check(lib().bf_ctx_set_profiling(self.h, 1 if enable else 0))

    def get_profile(self):
        """{kernel name: (total ms, launches)} since profiling was enabled."""
        n = 64
        names = (C.c_char_p * n)()
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        k = lib().bf_ctx_get_profile(self.h, n, names, ms, cnt)
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(min(k, n))}


class Model:
    def __init__(self, ctx, model):
        """model: workloads.Model (descriptor + bookkeeping)"""
        self.ctx, self.src = ctx, model
        h = C.c_void_p()
        _check(lib().bf_model_create(ctx.h, C.addressof(model.desc), C.byref(h)))
        self.h = h
        self.npar = lib().bf_model_npar(h)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().bf_model_destroy(self.h)
        self.h = None

    __del__ = close

    def transform(self, params, z_trigger):
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, self.npar)
        B = p.shape[0]
        nr = lib().bf_model_transform_nr(self.h)
        out = np.zeros((2, 3, nr, B))
        _check(lib().bf_transform_batch(self.ctx.h, self.h, _ptr(p), B, float(z_trigger), _ptr(out), B))
        r = lib().bf_model_transform_r0(self.h) + lib().bf_model_transform_dr(self.h) * np.arange(nr)
        return r, out


    def fft_planes(self, params, z_trigger):
        """FFT model: xi_comp on the z = 0 plane of the box for every vector, [B][comp][iy][ix]."""
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, self.npar)
        B = p.shape[0]
        npx, npy = C.c_int(0), C.c_int(0)
        _check(lib().bf_model_fft_plane_dims(self.h, C.byref(npx), C.byref(npy)))
        out = np.zeros((B, 2, npy.value, npx.value))
        _check(lib().bf_fft_planes_batch(self.ctx.h, self.h, _ptr(p), B, float(z_trigger), _ptr(out)))
        return out


class Data:
    def __init__(self, ctx, data):
        """data: workloads.Data"""
        self.ctx, self.src = ctx, data
        h = C.c_void_p()
        _check(lib().bf_data_create(ctx.h, C.addressof(data.desc), C.byref(h)))
        self.h = h
        self.n = lib().bf_data_ndata(h)
        self.n_grid = data.desc.n_grid

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().bf_data_destroy(self.h)
        self.h = None

    __del__ = close

    def sample_toys(self, truth, seed, first_toy, ntoy, scale=1.0):
        t = np.ascontiguousarray(truth, dtype=np.float64)
        out = np.zeros((ntoy, self.n))
        _check(lib().bf_sample_toys(self.ctx.h, self.h, _ptr(t), seed, first_toy, ntoy, scale, _ptr(out)))
        return out


class Toys:
    """Device-resident toy-MC realisations of one data set (bf_toys)."""

    def __init__(self, ctx, data, truth, seed, first_toy, ntoy, scale=1.0):
        self.ctx, self.data, self.ntoy = ctx, data, ntoy
        t = np.ascontiguousarray(truth, dtype=np.float64)
        h = C.c_void_p()
        _check(lib().bf_toys_create(ctx.h, data.h, _ptr(t), seed, first_toy, ntoy, scale, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().bf_toys_destroy(self.h)
        self.h = None

    __del__ = close

    def download(self):
        out = np.zeros((self.ntoy, self.data.n))
        _check(lib().bf_toys_download(self.ctx.h, self.h, _ptr(out)))
        return out


def chi2_batch_toys(ctx, model, data, params, toys, toy_index):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, model.npar)
    B = p.shape[0]
    idx = np.ascontiguousarray(toy_index, dtype=np.int32)
    assert idx.shape == (B,)
    chi2 = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    _check(lib().bf_chi2_batch_toys(ctx.h, model.h, data.h, _ptr(p), B, toys.h, _ptr(idx), _ptr(chi2), _ptr(status)))
    return chi2, status


def gram_batch(ctx, model, data, params, G, toys=None, toy_index=None):
    """Gauss-Newton normal equations of nfit = len(params)/G fits: gram[nfit, G, G], status[nfit, G]."""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, model.npar)
    nfit = p.shape[0] // G
    assert nfit * G == p.shape[0]
    gram = np.zeros((nfit, G, G))
    status = np.zeros((nfit, G), dtype=np.int32)
    idx = None if toy_index is None else np.ascontiguousarray(toy_index, dtype=np.int32)
    _check(lib().bf_gram_batch(ctx.h, model.h, data.h, _ptr(p), nfit, G, toys.h if toys is not None else None,
                               _ptr(idx), _ptr(gram), _ptr(status)))
    return gram, status


def gram_probes(ctx, model, data, centres, steps, toys=None, toy_index=None):
    """bf_gram_probes: central-difference probe groups generated on the GPU; gram[nfit, G, G], status[nfit]."""
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, model.npar)
    h = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, model.npar)
    nfit, nprobe = c.shape[0], int(np.count_nonzero(h[0]))
    G = 2 * nprobe + 1
    gram = np.zeros((nfit, G, G))
    status = np.zeros(nfit, dtype=np.int32)
    idx = None if toy_index is None else np.ascontiguousarray(toy_index, dtype=np.int32)
    _check(lib().bf_gram_probes(ctx.h, model.h, data.h, _ptr(c), _ptr(h), nfit, nprobe, toys.h if toys is not None else None,
                                _ptr(idx), _ptr(gram), _ptr(status)))
    return gram, status


def normal_probes(ctx, model, data, centres, steps, toys=None, toy_index=None):
    """bf_normal_probes: the probe groups of gram_probes reduced on the GPU to normal equations:
    M[nfit, n+1, n+1] = [y0 a]^T [y0 a] with a_i = y(x+h_i) - y(x-h_i), t[nfit, n] = y0 . (y(x+h_i) + y(x-h_i) - 2 y0)."""
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, model.npar)
    h = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, model.npar)
    nfit, n = c.shape[0], int(np.count_nonzero(h[0]))
    out = np.zeros((nfit, (n + 1) * (n + 1) + n))
    status = np.zeros(nfit, dtype=np.int32)
    idx = None if toy_index is None else np.ascontiguousarray(toy_index, dtype=np.int32)
    _check(lib().bf_normal_probes(ctx.h, model.h, data.h, _ptr(c), _ptr(h), nfit, n, toys.h if toys is not None else None,
                                  _ptr(idx), _ptr(out), _ptr(status)))
    return out[:, :(n + 1) * (n + 1)].reshape(nfit, n + 1, n + 1), out[:, (n + 1) * (n + 1):], status


def kspace_transform_weights(model):
    """Host-side Bessel weights [2][3][nr][nk] of a k-space model description (no GPU involved)."""
    nr, nk = C.c_int(0), C.c_int(0)
    _check(lib().bf_kspace_transform_weights(C.addressof(model.desc), C.byref(nr), C.byref(nk), None))
    w = np.zeros((2, 3, nr.value, nk.value))
    _check(lib().bf_kspace_transform_weights(C.addressof(model.desc), C.byref(nr), C.byref(nk), _ptr(w)))
    return w


def eval_batch(ctx, model, data, params):
    """pred[n_data, B], status[B] -- CorrelationFitter::getPrediction for B vectors."""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, model.npar)
    B = p.shape[0]
    pred = np.zeros((data.n, B))
    status = np.zeros(B, dtype=np.int32)
    _check(lib().bf_eval_batch(ctx.h, model.h, data.h, _ptr(p), B, _ptr(pred), B, _ptr(status)))
    return pred, status


def eval_undistorted_batch(ctx, model, data, params):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, model.npar)
    B = p.shape[0]
    xiu = np.zeros((data.n_grid, B))
    status = np.zeros(B, dtype=np.int32)
    _check(lib().bf_eval_undistorted_batch(ctx.h, model.h, data.h, _ptr(p), B, _ptr(xiu), B, _ptr(status)))
    return xiu, status


def chi2_batch(ctx, model, data, params, data_override=None):
    """chi2[B], status[B] -- BinnedData::chiSquare(getPrediction(p_b))."""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, model.npar)
    B = p.shape[0]
    ov = None
    if data_override is not None:
        ov = np.ascontiguousarray(data_override, dtype=np.float64)
        assert ov.shape == (B, data.n)
    chi2 = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    _check(lib().bf_chi2_batch(ctx.h, model.h, data.h, _ptr(p), B, _ptr(ov), _ptr(chi2), _ptr(status)))
    return chi2, status


def chi2_batch_dev(ctx, model, data, d_params_ptr, B, d_chi2_ptr, d_status_ptr=None, d_override_ptr=None):
    """Device-pointer variant (raw addresses as ints); asynchronous on the context's stream."""
    _check(lib().bf_chi2_batch_dev(ctx.h, model.h, data.h, C.c_void_p(d_params_ptr), B,
                                   C.c_void_p(d_override_ptr) if d_override_ptr else None,
                                   C.c_void_p(d_chi2_ptr), C.c_void_p(d_status_ptr) if d_status_ptr else None))


def chi2_batch_submit(ctx, model, data, params_ptr, B, chi2_ptr, status_ptr=None):
    """bf_chi2_batch_submit on raw host addresses (page-locked buffers); returns the ticket for chi2_batch_wait."""
    ticket = C.c_int(-1)
    _check(lib().bf_chi2_batch_submit(ctx.h, model.h, data.h, C.c_void_p(params_ptr), B, C.c_void_p(chi2_ptr),
                                      C.c_void_p(status_ptr) if status_ptr else None, C.byref(ticket)))
    return ticket.value


def chi2_batch_wait(ctx, ticket):
    _check(lib().bf_chi2_batch_wait(ctx.h, ticket))


def chi2_batch_raw(ctx, model, data, params_ptr, B, chi2_ptr, status_ptr=None):
    """Host-buffer entry point on raw addresses (e.g. pinned torch tensors)."""
    _check(lib().bf_chi2_batch(ctx.h, model.h, data.h, C.c_void_p(params_ptr), B, None, C.c_void_p(chi2_ptr),
                               C.c_void_p(status_ptr) if status_ptr else None))
