"""ctypes mirror of the plain-C descriptors in include/baofit_b200.h.

Plumbing only: these classes let Python (tests, bench.py, the synthetic-workload builder) fill
``bf_model_desc`` / ``bf_data_desc`` from numpy arrays and hand them to the C ABI. The same PODs
are accepted by the CPU oracle (oracle/oracle.h), which is how the parity tests feed identical
inputs to both sides. Field order must match the header exactly.
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

BF_MODEL_RSPACE, BF_MODEL_KSPACE, BF_MODEL_KSPACE_FFT = 0, 1, 2

FLAGS = {
    "cross_correlation": 1 << 0, "anisotropic": 1 << 1, "decoupled": 1 << 2,
    "combined_bias": 1 << 3, "combined_scale": 1 << 4, "nl_broadband": 1 << 5,
    "nl_correction": 1 << 6, "fit_nl_correction": 1 << 7, "nl_correction_alt": 1 << 8,
    "bin_smooth": 1 << 9, "bin_smooth_alt": 1 << 10, "hcd_model": 1 << 11,
    "hcd_model_alt": 1 << 12, "uvfluctuation": 1 << 13, "radiation_model": 1 << 14,
    "smooth_gauss": 1 << 15, "smooth_lorentz": 1 << 16, "distortion_alt": 1 << 17,
    "no_distortion": 1 << 18,
}
BF_METAL_NONE, BF_METAL_AUTO_TABLE, BF_METAL_CROSS_BICUBIC, BF_METAL_TOY = 0, 1, 2, 3
BF_BB_DIST_ADD, BF_BB_DIST_MUL, BF_BB_DM_DIST_ADD, BF_BB_DM_DIST_MUL = 0, 1, 2, 3
BF_STATUS_DILATION_MIN, BF_STATUS_DILATION_MAX, BF_STATUS_FFT_RANGE, BF_STATUS_Z_RETRIGGER = 1, 2, 4, 8


class BroadbandSpec(C.Structure):
    _fields_ = [
        ("enabled", C.c_int), ("idx_base", C.c_int),
        ("r_min", C.c_int), ("r_max", C.c_int), ("r_step", C.c_int), ("r_denom", C.c_int),
        ("mu_min", C.c_int), ("mu_max", C.c_int), ("mu_step", C.c_int),
        ("rp_min", C.c_int), ("rp_max", C.c_int), ("rp_step", C.c_int),
        ("rt_min", C.c_int), ("rt_max", C.c_int), ("rt_step", C.c_int),
        ("z_min", C.c_int), ("z_max", C.c_int), ("z_step", C.c_int),
        ("r0", C.c_double), ("z0", C.c_double),
    ]

    def ncoef(self):
        n = 1
        for a in ("z", "mu", "r", "rp", "rt"):
            lo, hi, st = (getattr(self, "%s_%s" % (a, s)) for s in ("min", "max", "step"))
            n *= len(range(lo, hi + 1, st))
        return n


class ModelDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("flags", C.c_uint), ("npar", C.c_int),
        ("zref", C.c_double), ("omega_matter", C.c_double),
        ("idx_beta", C.c_int), ("idx_bb", C.c_int), ("idx_gamma_bias", C.c_int),
        ("idx_gamma_beta", C.c_int),
        ("idx_delta_v", C.c_int), ("idx_bias2", C.c_int), ("idx_bb2", C.c_int),
        ("idx_beta_bias", C.c_int),
        ("idx_bao", C.c_int), ("idx_comb_scale", C.c_int), ("idx_rad", C.c_int),
        ("idx_nl", C.c_int), ("idx_cont", C.c_int), ("idx_bs", C.c_int), ("idx_hcd", C.c_int),
        ("idx_uv", C.c_int), ("idx_smgaus", C.c_int), ("idx_smlor", C.c_int),
        ("idx_nlcorr", C.c_int),
        ("n_fid", C.c_int),
        ("fid_r", c_double_p), ("fid_xi0", c_double_p), ("fid_xi2", c_double_p), ("fid_xi4", c_double_p),
        ("n_nw", C.c_int),
        ("nw_r", c_double_p), ("nw_xi0", c_double_p), ("nw_xi2", c_double_p), ("nw_xi4", c_double_p),
        ("n_pk", C.c_int),
        ("pk_k", c_double_p), ("pk_fid", c_double_p), ("pk_nw", c_double_p),
        ("rmin", C.c_double), ("rmax", C.c_double), ("dilmin", C.c_double), ("dilmax", C.c_double),
        ("samples_per_decade", C.c_int), ("ell_max", C.c_int),
        ("zeff", C.c_double), ("sigma8", C.c_double), ("dzmin", C.c_double),
        ("bb", BroadbandSpec * 4),
        ("metal_kind", C.c_int), ("idx_metal", C.c_int), ("metal_ncomb", C.c_int),
        ("metal_nmet", C.c_int),
        ("metal_p1", c_int_p), ("metal_p2", c_int_p),
        ("metal_ngrid", C.c_int),
        ("metal_zgrid", c_double_p), ("metal_auto", c_double_p),
        ("metal_nx", C.c_int), ("metal_ny", C.c_int),
        ("metal_x0", C.c_double), ("metal_y0", C.c_double), ("metal_dx", C.c_double),
        ("metal_dy", C.c_double),
        ("metal_cross", c_double_p),
        ("dist_matrix_order", C.c_int), ("dist_matrix", c_double_p),
        ("fft_spacing", C.c_double),
        ("fft_nx", C.c_int), ("fft_ny", C.c_int), ("fft_nz", C.c_int),
        ("fft_rmax", C.c_double),
        ("zcorr0", C.c_double), ("zcorr1", C.c_double), ("zcorr2", C.c_double),
    ]


class DataDesc(C.Structure):
    _fields_ = [
        ("n_data", C.c_int),
        ("r", c_double_p), ("mu", c_double_p), ("z", c_double_p),
        ("index", c_int_p), ("data", c_double_p),
        ("cov", c_double_p), ("cov_is_inverse", C.c_int),
        ("n_grid", C.c_int),
        ("grid_r", c_double_p), ("grid_mu", c_double_p), ("grid_z", c_double_p),
        ("binning_type", C.c_int), ("nmu_project", C.c_int),
    ]


def _dptr(a):
    return a.ctypes.data_as(c_double_p)


def _iptr(a):
    return a.ctypes.data_as(c_int_p)


class Keep:
    """A descriptor together with the numpy arrays its pointers refer to."""

    def __init__(self, desc):
        self.desc = desc
        self._arrays = []

    def f64(self, field, array):
        a = np.ascontiguousarray(np.asarray(array, dtype=np.float64))
        self._arrays.append(a)
        setattr(self.desc, field, _dptr(a))
        return a

    def i32(self, field, array):
        a = np.ascontiguousarray(np.asarray(array, dtype=np.int32))
        self._arrays.append(a)
        setattr(self.desc, field, _iptr(a))
        return a


def parse_broadband(spec, idx_base, r0, z0):
    """Restates the reference's broadband grammar (baofit/BroadbandModel.cc:21-90, 93-146).

    Accepted forms: ``[r,mu,z=]R,MU,Z`` ``r,mu=R,MU`` ``rP,rT=P,T`` ``rP,rT,z=P,T,Z``
    ``r,mu,rP,rT,z=R,MU,P,T,Z`` where each axis is ``n``, ``n1:n2`` or ``n1:n2:dn``.
    """
    axes = {"r": [0, 0, 1], "mu": [0, 0, 1], "rp": [0, 0, 1], "rt": [0, 0, 1], "z": [0, 0, 1]}
    forms = [("r,mu,rP,rT,z=", ["r", "mu", "rp", "rt", "z"]), ("rP,rT,z=", ["rp", "rt", "z"]),
             ("rP,rT=", ["rp", "rt"]), ("r,mu,z=", ["r", "mu", "z"]), ("r,mu=", ["r", "mu"])]
    body, names = spec, ["r", "mu", "z"]
    for tag, nm in forms:
        if spec.startswith(tag):
            body, names = spec[len(tag):], nm
            break
    parts = body.split(",")
    if len(parts) != len(names):
        raise ValueError("BroadbandModel: badly formatted parameter specification: " + spec)
    for name, part in zip(names, parts):
        v = [int(t) for t in part.split(":")]
        if len(v) == 1:
            v = [v[0], v[0], 1]
        elif len(v) == 2:
            v = [v[0], v[1], 1]
        elif len(v) != 3:
            raise ValueError("BroadbandModel: badly formatted parameter specification: " + spec)
        axes[name] = v
    s = BroadbandSpec()
    s.enabled, s.idx_base, s.r0, s.z0 = 1, idx_base, r0, z0
    rmin, rmax, rstep = axes["r"]
    if rmax < rmin or rstep == 0:
        raise ValueError("BroadbandModel: illegal r-parameter specification.")
    if rstep < 0:
        s.r_denom = -rstep
        rmin, rmax, rstep = rmin * s.r_denom, rmax * s.r_denom, 1
    else:
        s.r_denom = 1
    s.r_min, s.r_max, s.r_step = rmin, rmax, rstep
    for a in ("mu", "rp", "rt", "z"):
        lo, hi, st = axes[a]
        if hi < lo or st <= 0:
            raise ValueError("BroadbandModel: illegal %s-parameter specification." % a)
        setattr(s, a + "_min", lo), setattr(s, a + "_max", hi), setattr(s, a + "_step", st)
    if s.mu_min < 0 or s.mu_max > 8:
        raise ValueError("BroadbandModel: only multipoles 0-8 are allowed in mu-parameter specification.")
    return s


def new_model_desc():
    k = Keep(ModelDesc())
    d = k.desc
    for name, _ in ModelDesc._fields_:
        if name.startswith("idx_"):
            setattr(d, name, -1)
    d.zref, d.omega_matter = 2.25, 0.27
    d.dilmin, d.dilmax, d.samples_per_decade, d.ell_max = 0.5, 1.5, 50, 4
    d.sigma8, d.dzmin, d.zeff = 0.8338, 0.5, 0.0
    return k


def new_data_desc(r, mu, z, index, data, cov, cov_is_inverse=False, grid=None, multipole=False, nmu=0):
    k = Keep(DataDesc())
    d = k.desc
    d.n_data = len(r)
    k.f64("r", r), k.f64("mu", mu), k.f64("z", z), k.i32("index", index), k.f64("data", data)
    cov = np.asarray(cov, dtype=np.float64)
    assert cov.shape == (len(r), len(r))
    k.f64("cov", cov)
    d.cov_is_inverse = 1 if cov_is_inverse else 0
    d.binning_type, d.nmu_project = (1 if multipole else 0), nmu
    if grid is not None:
        gr, gmu, gz = grid
        d.n_grid = len(gr)
        k.f64("grid_r", gr), k.f64("grid_mu", gmu), k.f64("grid_z", gz)
    return k
