"""Python-side builders of model / data descriptors for tests and bench.py (plumbing).

These mirror, in Python, what the host-side C++ classes (baofit_b200/host/) do when they are
configured from an .ini file: lay the parameters out in the reference's defineParameter order,
load the tabulated templates, build the binned grid and apply the final cuts. They exist so
that pytest and bench.py can assemble workloads of the DR9 / DR11 shapes from numpy arrays
(synthetic covariance, distortion matrix, metal templates) without going
through files. No arithmetic of the hot path happens here.
"""
import gzip
import os
import re

import numpy as np

from . import desc as D

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# --------------------------------------------------------------------------------------------
# file formats
# --------------------------------------------------------------------------------------------

def load_table(path):
    """Two-column table ``x y`` (models/<name>.<ell>.dat, <name>_matterpower.dat)."""
    a = np.loadtxt(path)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


def load_multipoles(root, name):
    r0, x0 = load_table(os.path.join(root, "%s.0.dat" % name))
    r2, x2 = load_table(os.path.join(root, "%s.2.dat" % name))
    r4, x4 = load_table(os.path.join(root, "%s.4.dat" % name))
    assert np.array_equal(r0, r2) and np.array_equal(r0, r4)
    return r0, x0, x2, x4


def parse_binning(spec):
    """likely::createBinning strings -> bin centres (baofit/AbsCorrelationData.cc:96-150).

    ``[lo:hi]*n``  n equal bins spanning [lo,hi], centres lo+(i+1/2)w
    ``{lo:hi}*n``  n samples including both end points
    ``{a,b,c}``    explicit centres
    (conventions verified against demo/rmu.theory and demo/multi.theory)
    """
    s = spec.strip()
    m = re.fullmatch(r"\[\s*([^:\]]+):([^\]]+)\]\s*\*\s*(\d+)", s)
    if m:
        lo, hi, n = float(m.group(1)), float(m.group(2)), int(m.group(3))
        w = (hi - lo) / n
        return lo + (np.arange(n) + 0.5) * w
    m = re.fullmatch(r"\{\s*([^:\}]+):([^\}]+)\}\s*\*\s*(\d+)", s)
    if m:
        lo, hi, n = float(m.group(1)), float(m.group(2)), int(m.group(3))
        return lo + (hi - lo) * np.arange(n) / (n - 1) if n > 1 else np.array([lo])
    m = re.fullmatch(r"\{([^\}]*)\}", s)
    if m:
        return np.array([float(t) for t in m.group(1).split(",")])
    raise ValueError("createBinning: cannot parse " + spec)


def grid_coordinates(axis1, axis2, axis3, fmt):
    """(r, mu, z) of every grid bin; global index = axis1 slowest, axis3 fastest.

    comoving-cartesian: (rpar, rperp, z); comoving-polar: (r, mu, z)
    (baofit/ComovingCorrelationData.cc:46-69).
    """
    c1, c2, c3 = np.meshgrid(axis1, axis2, axis3, indexing="ij")
    c1, c2, c3 = c1.ravel(), c2.ravel(), c3.ravel()
    if fmt == "comoving-cartesian":
        r = np.sqrt(c1 * c1 + c2 * c2)
        mu = c1 / np.sqrt(c1 * c1 + c2 * c2)
    elif fmt == "comoving-polar":
        r, mu = c1.copy(), c2.copy()
    else:
        raise ValueError("unsupported data-format " + fmt)
    return r, mu, c3.copy()


class LambdaCdmRadiationUniverse:
    """Stand-in for cosmo::LambdaCdmRadiationUniverse(OmegaMatter, OmegaK = 0, h) (src/baofit.cc:413,
    dependency not in the tree): flat LCDM with photons and massless neutrinos, distances in Mpc/h.
    Omega_r h^2 = 2.469e-5 (1 + 0.2271 * 3.04) -- UNVERIFIED (dep); pinned end to end by
    data/BOSSDR9LyaF.scan (tests/test_quasar_data.py)."""

    def __init__(self, omega_matter=0.27, hubble_constant=0.7, omega_radiation=None):
        self.om = omega_matter
        self.orad = 2.469e-5 * (1 + 0.2271 * 3.04) / hubble_constant ** 2 if omega_radiation is None else omega_radiation

    def efunc(self, z):
        zp1 = 1 + z
        return np.sqrt(1 - self.om - self.orad + zp1 ** 3 * (self.om + zp1 * self.orad))

    def los_distance(self, z):
        from scipy.integrate import quad
        return 2997.92458 * quad(lambda x: 1.0 / self.efunc(x), 0, z, epsabs=0, epsrel=1e-13)[0]


def quasar_grid_coordinates(axis1, axis2, axis2_width, axis3, cosmology):
    """(r, mu, z) of every bin of an observing-coordinate grid (log(lam2/lam1), separation in
    arcmin, z): QuasarCorrelationData::transform (baofit/QuasarCorrelationData.cc:139-152).
    Also returns the (ll, sep) centres for the ll-min / sep cuts of finalize (:123)."""
    c1, c2, c3 = np.meshgrid(axis1, axis2, axis3, indexing="ij")
    ll, sep, z = c1.ravel(), c2.ravel(), c3.ravel()
    arcmin = 4 * np.arctan(1) / (60. * 180.)
    r, mu = np.zeros(len(ll)), np.zeros(len(ll))
    cache = {}

    def dist(x):
        if x not in cache:
            cache[x] = cosmology.los_distance(x)
        return cache[x]

    for i in range(len(ll)):
        ratio, zp1 = np.exp(0.5 * ll[i]), z[i] + 1
        z1, z2 = zp1 / ratio - 1, zp1 * ratio - 1
        dr_los = dist(z2) - dist(z1)
        swgt = sep[i] + (axis2_width * axis2_width / 12) / sep[i]
        dr_perp = dist(z[i]) * (swgt * arcmin)
        r[i] = np.sqrt(dr_los * dr_los + dr_perp * dr_perp)
        mu[i] = abs(dr_los) / r[i]
    return r, mu, z.copy(), ll, sep


def final_cuts(r, mu, z, has_data, rmin=0., rmax=200., rveto=(0., 0.), mumin=-1., mumax=1.,
               rperpmin=0., rperpmax=200., rparmin=-200., rparmax=200., zmin=0., zmax=10.):
    """Same comparisons, in the same order, as AbsCorrelationData::_applyFinalCuts
    (baofit/AbsCorrelationData.cc:67-94). Returns the kept global indices (ascending)."""
    keep = []
    for i in np.nonzero(has_data)[0]:
        ri, zi, mi = r[i], z[i], mu[i]
        if ri < rmin or ri > rmax:
            continue
        if ri > rveto[0] and ri < rveto[1]:
            continue
        if zi < zmin or zi > zmax:
            continue
        if mi < mumin or mi > mumax:
            continue
        rpar = ri * mi
        if rpar < rparmin or rpar > rparmax:
            continue
        rperp = ri * np.sqrt(1 - mi * mi)
        if rperp < rperpmin or rperp > rperpmax:
            continue
        keep.append(i)
    return np.array(keep, dtype=np.int32)


def load_data_vector(path, nbins):
    """``index value`` lines; extra tokens ignored (baofit/AbsCorrelationData.cc:174-195)."""
    val = np.zeros(nbins)
    has = np.zeros(nbins, dtype=bool)
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            i = int(t[0])
            val[i] = float(t[1])
            has[i] = True
    return val, has


def load_cov(path, nbins):
    """``i j value`` lines, any triangle(s) (baofit/AbsCorrelationData.cc:204-236)."""
    op = gzip.open if path.endswith(".gz") else open
    cov = np.zeros((nbins, nbins))
    with op(path, "rt") as f:
        a = np.loadtxt(f)
    i, j = a[:, 0].astype(int), a[:, 1].astype(int)
    cov[i, j] = a[:, 2]
    cov[j, i] = a[:, 2]
    return cov


# --------------------------------------------------------------------------------------------
# parameter layouts
# --------------------------------------------------------------------------------------------

class Model:
    """Descriptor + parameter bookkeeping (names, start values, step sizes, floating flags)."""

    def __init__(self, keep, names, values, errors):
        self.keep = keep
        self.desc = keep.desc
        self.names = list(names)
        self.values = np.array(values, dtype=np.float64)
        self.errors = np.array(errors, dtype=np.float64)
        self.floating = np.ones(len(names), dtype=bool)
        self.desc.npar = len(names)

    def index(self, name):
        return self.names.index(name)

    def configure(self, script):
        """Minimal restatement of likely's parameter script: ``value[name]=v``,
        ``fix[name]=v`` / ``fix[name]``, ``release[name]``, ``error[name]=v``; ``*`` globs;
        ``;`` separated. binning/prior statements are ignored here."""
        for stmt in script.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            m = re.fullmatch(r"(\w+)\[([^\]]+)\]\s*(?:=\s*(\S+))?", stmt)
            if not m:
                continue
            op, pat, val = m.group(1), m.group(2), m.group(3)
            rx = re.compile("^" + re.escape(pat).replace(r"\*", ".*") + "$")
            hits = [i for i, n in enumerate(self.names) if rx.match(n)]
            if not hits and op in ("value", "fix", "release", "error"):
                raise ValueError("no parameter matches " + pat)
            for i in hits:
                if op == "value":
                    self.values[i] = float(val)
                elif op == "fix":
                    self.floating[i] = False
                    if val is not None:
                        self.values[i] = float(val)
                elif op == "release":
                    self.floating[i] = True
                elif op == "error":
                    self.errors[i] = float(val)
        return self

    def random_batch(self, B, rng, scale=0.5, sweep=None):
        """p_b = p0 + scale*err*U(-1,1) on floating parameters; ``sweep`` maps a parameter name to
        a (lo, hi) box drawn uniformly instead."""
        p = np.tile(self.values, (B, 1))
        for j in np.nonzero(self.floating)[0]:
            p[:, j] += scale * self.errors[j] * rng.uniform(-1, 1, size=B)
        for name, (lo, hi) in (sweep or {}).items():
            p[:, self.index(name)] = rng.uniform(lo, hi, size=B)
        return np.ascontiguousarray(p)


def _add_broadband(names, values, errors, spec, tag, r0, z0):
    s = D.parse_broadband(spec, len(names), r0, z0)
    for zi in range(s.z_min, s.z_max + 1, s.z_step):
        for mi in range(s.mu_min, s.mu_max + 1, s.mu_step):
            for ri in range(s.r_min, s.r_max + 1, s.r_step):
                for pi in range(s.rp_min, s.rp_max + 1, s.rp_step):
                    for ti in range(s.rt_min, s.rt_max + 1, s.rt_step):
                        # baofit/BroadbandModel.cc:150 "%s z%d mu%d r%+d rP%d rT%d"
                        names.append("%s z%d mu%d r%+d rP%d rT%d" % (tag, zi, mi, ri, pi, ti))
                        values.append(0.0)
                        errors.append(1e-3)
    return s


def _linear_bias(d, names, values, errors, cross, combined_bias, set_flags=True):
    d.idx_beta = len(names); names.append("beta"); values.append(1.4); errors.append(0.1)
    d.idx_bb = len(names); names.append("(1+beta)*bias"); values.append(-0.336); errors.append(0.03)
    d.idx_gamma_bias = len(names); names.append("gamma-bias"); values.append(3.8); errors.append(0.3)
    d.idx_gamma_beta = len(names); names.append("gamma-beta"); values.append(0.0); errors.append(0.1)
    if cross:
        d.idx_delta_v = len(names); names.append("delta-v"); values.append(0.0); errors.append(10.)
        d.idx_bias2 = len(names); names.append("bias2"); values.append(3.6); errors.append(0.1)
        d.idx_bb2 = len(names); names.append("beta2*bias2"); values.append(1.0); errors.append(0.05)


def _metals(k, names, values, errors, metal):
    """metal = dict(kind=..., civ=bool, zgrid=[ncomb][ngrid], auto=..., cross=..., nx, ny, x0, y0, dx, dy)"""
    d = k.desc
    kind = metal["kind"]
    d.metal_kind = kind
    d.idx_metal = len(names)
    if kind == D.BF_METAL_TOY:
        for i in range(4):
            names.append("metal ampl%d" % i); values.append(1.0); errors.append(0.1)
        names.append("metal width0"); values.append(6.0); errors.append(0.6)
        return
    lines = ["Si2a", "Si2b", "Si2c", "Si3"] + (["CIV"] if metal.get("civ") else [])
    for ln in lines:
        names.append("beta " + ln); values.append(1.0); errors.append(0.1)
        names.append("bias " + ln); values.append(-0.01); errors.append(0.001)
    if kind == D.BF_METAL_AUTO_TABLE:
        # baofit/MetalCorrelationModel.cc:60-65 (+ :80-86 with CIV)
        p1 = [0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 4, 4, 4, 4]
        p2 = [1, 2, 3, 4, 1, 2, 3, 2, 3, 3, 1, 2, 3, 4]
        if metal.get("civ"):
            p1 += [0, 1, 2, 3, 4, 5]
            p2 += [5, 5, 5, 5, 5, 5]
    else:
        # baofit/MetalCorrelationModel.cc:102-107 (+ :128-134)
        p1, p2 = [0, 0, 0, 0], [1, 2, 3, 4]
        if metal.get("civ"):
            p1, p2 = p1 + [0], p2 + [5]
    d.metal_ncomb = len(p1)
    d.metal_nmet = 5 + (1 if metal.get("civ") else 0)
    k.i32("metal_p1", p1), k.i32("metal_p2", p2)
    zg = np.asarray(metal["zgrid"], dtype=np.float64)
    assert zg.shape[0] == d.metal_ncomb
    d.metal_ngrid = zg.shape[1]
    k.f64("metal_zgrid", zg)
    if kind == D.BF_METAL_AUTO_TABLE:
        a = np.asarray(metal["auto"], dtype=np.float64)
        assert a.shape == (d.metal_ncomb * 3, d.metal_ngrid)
        k.f64("metal_auto", a)
    else:
        c = np.asarray(metal["cross"], dtype=np.float64)
        d.metal_nx, d.metal_ny = metal["nx"], metal["ny"]
        assert c.shape == (d.metal_ncomb * 3, d.metal_ny, d.metal_nx)
        d.metal_x0, d.metal_y0, d.metal_dx, d.metal_dy = metal["x0"], metal["y0"], metal["dx"], metal["dy"]
        k.f64("metal_cross", c)


def rspace_model(fid, nw, zref=2.25, omega_matter=0.27, anisotropic=True, decoupled=True,
                 cross=False, combined_bias=False, combined_scale=False, dist_add="", dist_mul="",
                 dist_r0=100., metal=None):
    """BaoCorrelationModel parameter layout (baofit/BaoCorrelationModel.cc:27-86).
    fid / nw = (r, xi0, xi2, xi4)."""
    k = D.new_model_desc()
    d = k.desc
    d.kind, d.zref, d.omega_matter = D.BF_MODEL_RSPACE, zref, omega_matter
    flags = 0
    for name, on in (("anisotropic", anisotropic), ("decoupled", decoupled), ("cross_correlation", cross),
                     ("combined_bias", combined_bias), ("combined_scale", combined_scale)):
        if on:
            flags |= D.FLAGS[name]
    d.flags = flags
    names, values, errors = [], [], []
    _linear_bias(d, names, values, errors, cross, combined_bias)
    if combined_bias:
        d.idx_beta_bias = len(names); names.append("beta*bias"); values.append(-0.196); errors.append(0.02)
    d.idx_bao = len(names)
    for n, v, e in (("BAO amplitude", 1, 0.15), ("BAO alpha-iso", 1, 0.02), ("BAO alpha-parallel", 1, 0.1),
                    ("BAO alpha-perp", 1, 0.1), ("gamma-scale", 0, 0.5)):
        names.append(n); values.append(float(v)); errors.append(e)
    d.idx_rad = len(names)
    for n, v, e in (("Rad strength", 0., 0.1), ("Rad anisotropy", 0., 0.1), ("Rad mean free path", 200., 10.),
                    ("Rad quasar lifetime", 10., 0.1)):
        names.append(n); values.append(v); errors.append(e)
    if combined_scale:
        d.idx_comb_scale = len(names); names.append("BAO aperp/apar"); values.append(1.0); errors.append(0.1)
    k.f64("fid_r", fid[0]), k.f64("fid_xi0", fid[1]), k.f64("fid_xi2", fid[2]), k.f64("fid_xi4", fid[3])
    d.n_fid = len(fid[0])
    k.f64("nw_r", nw[0]), k.f64("nw_xi0", nw[1]), k.f64("nw_xi2", nw[2]), k.f64("nw_xi4", nw[3])
    d.n_nw = len(nw[0])
    if metal:
        _metals(k, names, values, errors, metal)
    if dist_add:
        d.bb[D.BF_BB_DIST_ADD] = _add_broadband(names, values, errors, dist_add, "dist add", dist_r0, zref)
    if dist_mul:
        d.bb[D.BF_BB_DIST_MUL] = _add_broadband(names, values, errors, dist_mul, "dist mul", dist_r0, zref)
    m = Model(k, names, values, errors)
    m.configure("fix[Rad*]")  # baofit/BaoCorrelationModel.cc:42
    if combined_bias:
        m.configure("fix[(1+beta)*bias]")
    if combined_scale:
        m.configure("fix[BAO alpha-perp]")
    return m


def kspace_model(pk_fid, pk_nw, rmin, rmax, zref=2.25, omega_matter=0.27, dilmin=0.5, dilmax=1.5,
                 samples_per_decade=50, ell_max=4, anisotropic=True, decoupled=True, cross=False,
                 nl_broadband=False, nl_correction=False, fit_nl_correction=False, nl_correction_alt=False,
                 bin_smooth=False, bin_smooth_alt=False, hcd_model=False, hcd_model_alt=False,
                 uvfluctuation=False, radiation_model=False, smooth_gauss=False, smooth_lorentz=False,
                 combined_bias=False, combined_scale=False, dist_add="", dist_mul="", dist_r0=100.,
                 zeff=0., sigma8=0.8338, dzmin=0.5, metal=None, dist_matrix=None,
                 dm_dist_add="", dm_dist_mul=""):
    """BaoKSpaceCorrelationModel parameter layout (baofit/BaoKSpaceCorrelationModel.cc:49-211).
    pk_fid / pk_nw = (k, P) on identical k nodes."""
    k = D.new_model_desc()
    d = k.desc
    d.kind, d.zref, d.omega_matter = D.BF_MODEL_KSPACE, zref, omega_matter
    flags = 0
    loc = locals()
    for name in ("anisotropic", "decoupled", "nl_broadband", "nl_correction", "fit_nl_correction",
                 "nl_correction_alt", "bin_smooth", "bin_smooth_alt", "hcd_model", "hcd_model_alt",
                 "uvfluctuation", "radiation_model", "smooth_gauss", "smooth_lorentz", "combined_bias",
                 "combined_scale"):
        if loc[name]:
            flags |= D.FLAGS[name]
    if cross:
        flags |= D.FLAGS["cross_correlation"]
    d.flags = flags
    d.rmin, d.rmax, d.dilmin, d.dilmax = rmin, rmax, dilmin, dilmax
    d.samples_per_decade, d.ell_max, d.zeff, d.sigma8, d.dzmin = samples_per_decade, ell_max, zeff, sigma8, dzmin
    names, values, errors = [], [], []

    def add(n, v, e):
        names.append(n); values.append(float(v)); errors.append(float(e))
        return len(names) - 1

    _linear_bias(d, names, values, errors, cross, combined_bias)
    d.idx_nl = add("SigmaNL-perp", 3.26, 0.3); add("1+f", 2, 0.1)
    d.idx_bao = add("BAO amplitude", 1, 0.15); add("BAO alpha-iso", 1, 0.02)
    add("BAO alpha-parallel", 1, 0.1); add("BAO alpha-perp", 1, 0.1); add("gamma-scale", 0, 0.5)
    if bin_smooth or bin_smooth_alt:
        d.idx_bs = add("bin scale-par", 4, 0.4); add("bin scale-perp", 4, 0.4)
    if hcd_model or hcd_model_alt:
        d.idx_hcd = add("HCD beta", 1, 0.1); add("HCD rel bias", 0.1, 0.01); add("HCD scale", 20, 2)
    if uvfluctuation:
        d.idx_uv = add("UV rel bias", -1, 0.1); add("UV abs resp bias", -0.667, 0.06); add("UV mean free path", 300, 30)
    if radiation_model:
        d.idx_rad = add("Rad strength", 0, 0.1); add("Rad anisotropy", 0, 0.1)
        add("Rad mean free path", 200, 10); add("Rad quasar lifetime", 10, 0.1)
    if smooth_gauss:
        d.idx_smgaus = add("smooth scale gauss", 1, 0.1)
    if smooth_lorentz:
        d.idx_smlor = add("smooth scale lorentz", 1, 0.1)
    if combined_bias:
        d.idx_beta_bias = add("beta*bias", -0.196, 0.02)
    if combined_scale:
        d.idx_comb_scale = add("BAO aperp/apar", 1, 0.1)
    assert np.array_equal(pk_fid[0], pk_nw[0])
    k.f64("pk_k", pk_fid[0]), k.f64("pk_fid", pk_fid[1]), k.f64("pk_nw", pk_nw[1])
    d.n_pk = len(pk_fid[0])
    if fit_nl_correction:  # baofit/NonLinearCorrectionModel.cc:41-44
        d.idx_nlcorr = add("nlcorr qnl", 0.867, 0.08); add("nlcorr kp", 19.4, 1)
    if metal:
        _metals(k, names, values, errors, metal)
    if dist_matrix is not None:
        dm = np.asarray(dist_matrix, dtype=np.float64)
        assert dm.ndim == 2 and dm.shape[0] == dm.shape[1]
        d.dist_matrix_order = dm.shape[0]
        k.f64("dist_matrix", dm)
        if dm_dist_add:
            d.bb[D.BF_BB_DM_DIST_ADD] = _add_broadband(names, values, errors, dm_dist_add, "dist Matrix dist add", dist_r0, zref)
        if dm_dist_mul:
            d.bb[D.BF_BB_DM_DIST_MUL] = _add_broadband(names, values, errors, dm_dist_mul, "dist Matrix dist mul", dist_r0, zref)
    if dist_add:
        d.bb[D.BF_BB_DIST_ADD] = _add_broadband(names, values, errors, dist_add, "dist add", dist_r0, zref)
    if dist_mul:
        d.bb[D.BF_BB_DIST_MUL] = _add_broadband(names, values, errors, dist_mul, "dist mul", dist_r0, zref)
    m = Model(k, names, values, errors)
    if combined_bias:
        m.configure("fix[(1+beta)*bias]")
    if combined_scale:
        m.configure("fix[BAO alpha-perp]")
    return m


def fft_model(pk_fid, pk_nw, spacing=4., nx=400, ny=0, nz=0, rmax=180., dilmax=1.5, fft_rmax=0.,
              zref=2.25, omega_matter=0.27, zcorr=(0., 0., 0.), sigma8=0.8338, anisotropic=True,
              decoupled=True, cross=False, nl_broadband=False, nl_correction=False,
              fit_nl_correction=False, nl_correction_alt=False, distortion_alt=False,
              no_distortion=False, dist_add="", dist_mul="", dist_r0=100.):
    """BaoKSpaceFftCorrelationModel parameter layout (baofit/BaoKSpaceFftCorrelationModel.cc:36-113;
    ngridy / ngridz default to ngridx, src/baofit.cc:398-399)."""
    k = D.new_model_desc()
    d = k.desc
    d.kind, d.zref, d.omega_matter = D.BF_MODEL_KSPACE_FFT, zref, omega_matter
    flags = 0
    loc = locals()
    for name in ("anisotropic", "decoupled", "nl_broadband", "nl_correction", "fit_nl_correction",
                 "nl_correction_alt", "distortion_alt", "no_distortion"):
        if loc[name]:
            flags |= D.FLAGS[name]
    if cross:
        flags |= D.FLAGS["cross_correlation"]
    d.flags = flags
    d.fft_spacing, d.fft_nx, d.fft_ny, d.fft_nz = spacing, nx, ny or nx, nz or ny or nx
    d.rmax, d.dilmax, d.fft_rmax, d.sigma8 = rmax, dilmax, fft_rmax, sigma8
    d.zcorr0, d.zcorr1, d.zcorr2 = zcorr
    names, values, errors = [], [], []

    def add(n, v, e):
        names.append(n); values.append(float(v)); errors.append(float(e))
        return len(names) - 1

    _linear_bias(d, names, values, errors, cross, False)
    d.idx_nl = add("SigmaNL-perp", 3.26, 0.3); add("1+f", 2, 0.1)
    d.idx_cont = add("cont-kc", 0.02, 0.002); add("cont-pc", 1, 0.1)
    d.idx_bao = add("BAO amplitude", 1, 0.15); add("BAO alpha-iso", 1, 0.02)
    add("BAO alpha-parallel", 1, 0.1); add("BAO alpha-perp", 1, 0.1); add("gamma-scale", 0, 0.5)
    assert np.array_equal(pk_fid[0], pk_nw[0])
    k.f64("pk_k", pk_fid[0]), k.f64("pk_fid", pk_fid[1]), k.f64("pk_nw", pk_nw[1])
    d.n_pk = len(pk_fid[0])
    if dist_add:
        d.bb[D.BF_BB_DIST_ADD] = _add_broadband(names, values, errors, dist_add, "dist add", dist_r0, zref)
    if dist_mul:
        d.bb[D.BF_BB_DIST_MUL] = _add_broadband(names, values, errors, dist_mul, "dist mul", dist_r0, zref)
    if fit_nl_correction:  # constructed last, :111-112
        d.idx_nlcorr = add("nlcorr qnl", 0.867, 0.08); add("nlcorr kp", 19.4, 1)
    return Model(k, names, values, errors)


# --------------------------------------------------------------------------------------------
# ready-made workloads
# --------------------------------------------------------------------------------------------

QSO_AXIS1 = ("{-154.997,-144.999,-134.999,-125.001,-115.002,-105.004,-95.0051,-85.0055,-75.0069,-65.009,"
             "-55.0095,-45.011,-35.0132,-25.0134,-15.0126,-5.01466,4.98391,14.9834,24.9824,34.9815,44.9809,"
             "54.9802,64.9791,74.9796,84.9788,94.9768,104.977,114.977,124.976,134.975,144.977,154.976}")
QSO_AXIS2 = ("{6.65322,15.54,25.313,35.2281,45.1747,55.154,65.1229,75.1073,85.0988,95.0845,105.076,115.064,"
             "125.067,135.052,145.053,154.959}")
QSO_AXIS3 = "{2.35717}"


class Data:
    def __init__(self, keep, grid):
        self.keep = keep
        self.desc = keep.desc
        self.grid = grid
        self.n_data = keep.desc.n_data


def make_data(r, mu, z, keep_idx, dvec, cov, cov_is_inverse=False, with_grid=False, multipole=False, nmu=0):
    """multipole=True: ``mu`` holds the multipole ell of every bin (data-format comoving-multipole)."""
    grid = (r, mu, z)
    k = D.new_data_desc(r[keep_idx], mu[keep_idx], z[keep_idx], keep_idx, dvec, cov, cov_is_inverse,
                        grid if with_grid else None, multipole=multipole, nmu=nmu)
    return Data(k, grid)


def synthetic_cov(xi0, seed=1, rel=0.1):
    """C = A A^T / N + diag(sigma^2), sigma_i = rel |xi_i| + 1e-6 (rel = 0.1;
    rel = 1 gives BOSS-like errors on the BAO scale parameters, a few per cent)."""
    n = len(xi0)
    rng = np.random.Generator(np.random.PCG64(seed))
    A = rng.standard_normal((n, n))
    sig = rel * np.abs(xi0) + 1e-6
    C = (A @ A.T) / n
    C = C * np.outer(sig, sig) * 0.2 + np.diag(sig * sig)
    return 0.5 * (C + C.T)


def synthetic_dist_matrix(axis1, axis2, seed=2):
    """D = I - W + noise: W averages over bins with the same r_perp (axis2) bin, i.e. dense in
    r_parallel blocks like a continuum-fitting projection."""
    n1, n2 = len(axis1), len(axis2)
    n = n1 * n2
    rng = np.random.Generator(np.random.PCG64(seed))
    i1, i2 = np.divmod(np.arange(n), n2)
    same = (i2[:, None] == i2[None, :]).astype(np.float64)
    w = np.exp(-0.5 * ((axis1[i1][:, None] - axis1[i1][None, :]) / 60.0) ** 2) * same
    w /= w.sum(axis=1, keepdims=True)
    return np.eye(n) - 0.5 * w + 1e-3 * rng.uniform(-1, 1, size=(n, n))


def synthetic_cross_metals(root=None, nx=41, ny=81, dx=4.0, dy=4.0, x0=0.0, y0=-160.0, ngrid=512, z=2.35717):
    """Four QSO x metal combinations, three multipoles each, on a regular (rperp, rpar) grid:
    shifted, scaled copies of the fiducial xi_ell(r) tables."""
    root = root or os.path.join(GOLDEN, "models")
    r, x0t, x2t, x4t = load_multipoles(root, "DR9LyaMocksLCDM")
    tabs = (x0t, x2t, x4t)
    X = x0 + dx * np.arange(nx)
    Y = y0 + dy * np.arange(ny)
    shifts = [21.0, -52.0, 103.0, -60.0]
    out = np.zeros((12, ny, nx))
    for c, sh in enumerate(shifts):
        rr = np.sqrt(X[None, :] ** 2 + (Y[:, None] - sh) ** 2)
        for l in range(3):
            out[3 * c + l] = 1e-2 * np.interp(np.clip(rr, r[0], r[-1]), r, tabs[l]) * (rr / 50.0) ** 2 / (1 + (rr / 50.0) ** 2)
    zgrid = np.full((4, ngrid), z)
    return dict(kind=D.BF_METAL_CROSS_BICUBIC, zgrid=zgrid, cross=out, nx=nx, ny=ny, x0=x0, y0=y0, dx=dx, dy=dy)


def qso_rspace(root=None, golden=GOLDEN):
    """config/BOSSDR11QSOLyaF.ini: r-space cross-correlation fit with the real data and
    covariance (the golden .scan case)."""
    root = root or os.path.join(golden, "models")
    fid = load_multipoles(root, "DR9LyaMocksNLeffPk")
    nw = load_multipoles(root, "DR9LyaMocksSB")
    m = rspace_model(fid, nw, zref=2.25, cross=True, dist_add="rP,rT=0:2,-3:1")
    m.configure("value[beta]=1.1; fix[(1+beta)*bias]=-0.336; fix[gamma-bias]=0.9; fix[gamma-beta]=0;"
                "value[bias2]=3.64; fix[beta2*bias2]=0.962524; value[delta-v]=0; fix[BAO amplitude]=1;"
                "fix[BAO alpha-iso]; value[BAO alpha-p*]=1; fix[gamma-scale]=0")
    a1, a2, a3 = parse_binning(QSO_AXIS1), parse_binning(QSO_AXIS2), parse_binning(QSO_AXIS3)
    r, mu, z = grid_coordinates(a1, a2, a3, "comoving-cartesian")
    n = len(r)
    val, has = load_data_vector(os.path.join(golden, "data", "BOSSDR11QSOLyaF.data"), n)
    cov = load_cov(os.path.join(golden, "data", "BOSSDR11QSOLyaF.cov.gz"), n)
    keep = final_cuts(r, mu, z, has, rmin=40, rmax=180)
    data = make_data(r, mu, z, keep, val[keep], cov[np.ix_(keep, keep)])
    return m, data


def qso_kspace(root=None, golden=GOLDEN, with_dist_matrix=True, with_metals=True, real_cov=True):
    """config/BOSSDR11QSOLyaF_k.ini plus the BASELINE add-ons (distortion matrix of order 512,
    metal-model-interpolate), which the reference ships no files for: synthetic fixtures."""
    root = root or os.path.join(golden, "models")
    pk_fid = load_table(os.path.join(root, "DR9LyaMocksLCDM_matterpower.dat"))
    pk_nw = load_table(os.path.join(root, "DR9LyaMocksLCDMSB_matterpower.dat"))
    a1, a2, a3 = parse_binning(QSO_AXIS1), parse_binning(QSO_AXIS2), parse_binning(QSO_AXIS3)
    r, mu, z = grid_coordinates(a1, a2, a3, "comoving-cartesian")
    n = len(r)
    dm = synthetic_dist_matrix(a1, a2) if with_dist_matrix else None
    metal = synthetic_cross_metals(root, ngrid=n) if with_metals else None
    m = kspace_model(pk_fid, pk_nw, rmin=40, rmax=180, zref=2.3, cross=True, dist_add="rP,rT=0:2,-3:1",
                     dist_matrix=dm, metal=metal)
    m.configure("value[beta]=1.1; fix[(1+beta)*bias]=-0.351; fix[gamma-bias]=0.9; fix[gamma-beta]=0;"
                "value[bias2]=3.64; fix[beta2*bias2]=0.962524; value[delta-v]=0; fix[BAO amplitude]=1;"
                "fix[BAO alpha-iso]; value[BAO alpha-p*]=1; fix[gamma-scale]=0; fix[1+f]=1.966;"
                "fix[SigmaNL-perp]=3.26")
    if with_metals:
        m.configure("fix[beta Si*]")
    val, has = load_data_vector(os.path.join(golden, "data", "BOSSDR11QSOLyaF.data"), n)
    keep = final_cuts(r, mu, z, has, rmin=40, rmax=180)
    if real_cov:
        cov = load_cov(os.path.join(golden, "data", "BOSSDR11QSOLyaF.cov.gz"), n)[np.ix_(keep, keep)]
    else:
        cov = synthetic_cov(val[keep])
    data = make_data(r, mu, z, keep, val[keep], cov, with_grid=with_dist_matrix)
    return m, data


DR9_AXIS1 = ("{0.,0.001,0.003,0.005,0.007,0.009,0.011,0.013,0.015,0.017,0.019,0.021,0.023,0.025,0.027,0.029,0.031,"
             "0.033,0.035,0.037,0.039,0.041,0.043,0.045,0.047,0.049,0.0591608,0.0828251}")


def dr9_data(golden=GOLDEN, cosmology=None):
    """data/BOSSDR9LyaF.{data,cov} on the 28 x 18 x 3 observing grid of config/BOSSDR9LyaF.ini
    (data-format = quasar), cuts r in [50,190], ll >= 0.003."""
    cosmology = cosmology or LambdaCdmRadiationUniverse()
    a1, a2, a3 = parse_binning(DR9_AXIS1), parse_binning("[0:180]*18"), parse_binning("{2.0,2.5,3.0}")
    r, mu, z, ll, sep = quasar_grid_coordinates(a1, a2, 10.0, a3, cosmology)
    n = len(r)
    val, has = load_data_vector(os.path.join(golden, "data", "BOSSDR9LyaF.data"), n)
    cov = load_cov(os.path.join(golden, "data", "BOSSDR9LyaF.cov.gz"), n)
    keep = final_cuts(r, mu, z, has, rmin=50, rmax=190)
    keep = np.array([i for i in keep if not (ll[i] < 0.003 or ll[i] > 1 or sep[i] < 0 or sep[i] > 200)], dtype=np.int32)
    return make_data(r, mu, z, keep, val[keep], cov[np.ix_(keep, keep)])


def dr9_rspace(root=None, golden=GOLDEN, cosmology=None):
    """config/BOSSDR9LyaF.ini: BaoCorrelationModel, Lya auto-correlation, quasar coordinates, real
    data and covariance (the golden BOSSDR9LyaF.scan case). Priors live in the host fitter."""
    root = root or os.path.join(golden, "models")
    fid = load_multipoles(root, "DR9LyaMocksNLeffPk")
    nw = load_multipoles(root, "DR9LyaMocksSB")
    m = rspace_model(fid, nw, zref=2.25, dist_add="-2:0,0:4:2,0")
    m.configure("value[beta]=1.4; value[(1+beta)*bias]=-0.336; value[gamma-bias]=3.8; fix[gamma-beta]=0;"
                "fix[BAO amplitude]=1; fix[BAO alpha-iso]; value[BAO alpha-p*]=1; fix[gamma-scale]=0")
    return m, dr9_data(golden, cosmology)


def dr9_kspace(root=None, golden=GOLDEN, cosmology=None):
    """config/BOSSDR9LyaF_k.ini: BaoKSpaceCorrelationModel on the same data (BASELINE config 2)."""
    root = root or os.path.join(golden, "models")
    pk_fid = load_table(os.path.join(root, "DR9LyaMocksLCDM_matterpower.dat"))
    pk_nw = load_table(os.path.join(root, "DR9LyaMocksLCDMSB_matterpower.dat"))
    m = kspace_model(pk_fid, pk_nw, rmin=50, rmax=190, zref=2.25, dist_add="-2:0,0:4:2,0")
    m.configure("value[beta]=1.4; value[(1+beta)*bias]=-0.336; value[gamma-bias]=3.8; fix[gamma-beta]=0;"
                "fix[BAO amplitude]=1; fix[BAO alpha-iso]; value[BAO alpha-p*]=1; fix[gamma-scale]=0;"
                "fix[1+f]=1.966; fix[SigmaNL-perp]=3.26")
    return m, dr9_data(golden, cosmology)


def demo_rmu(root=None, golden=GOLDEN):
    """demo/rmu_to_bao.ini: 40x20 (r,mu) grid at z=2.25, diagonal inverse covariance."""
    root = root or os.path.join(golden, "models")
    fid = load_multipoles(root, "DR9LyaMocks")
    nw = load_multipoles(root, "DR9LyaMocksSB")
    m = rspace_model(fid, nw, zref=2.25)
    m.configure("fix[(1+beta)*bias]=-0.336; fix[gamma-bias]=3.8; fix[gamma-beta]=0; fix[gamma-scale]=0;"
                "fix[BAO amplitude]=1; fix[BAO alpha-iso]=1; value[beta]=1.4; value[BAO alpha-p*]=1.0")
    a1, a2, a3 = parse_binning("[40:200]*40"), parse_binning("[0:1]*20"), parse_binning("{2.25}")
    r, mu, z = grid_coordinates(a1, a2, a3, "comoving-polar")
    n = len(r)
    val, has = load_data_vector(os.path.join(golden, "demo", "rmu.data"), n)
    icov = load_cov(os.path.join(golden, "demo", "rmu.icov"), n)
    keep = final_cuts(r, mu, z, has, rmin=40, rmax=200)
    data = make_data(r, mu, z, keep, val[keep], icov[np.ix_(keep, keep)], cov_is_inverse=True)
    return m, data


def demo_multi(root=None, golden=GOLDEN, which=0, lmax=2):
    """demo/multi_to_bao.ini: 29 radii x ell = 0,2,4 multipole bins at z = 2.25 (data-format
    comoving-multipole), one of the three demo realisations demo/multi_<which>.{data,cov}."""
    root = root or os.path.join(golden, "models")
    fid = load_multipoles(root, "DR9LyaMocks")
    nw = load_multipoles(root, "DR9LyaMocksSB")
    m = rspace_model(fid, nw, zref=2.25)
    m.configure("fix[(1+beta)*bias]=-0.336; fix[gamma-bias]=3.8; fix[gamma-beta]=0; fix[gamma-scale]=0;"
                "fix[BAO amplitude]=1; fix[BAO alpha-iso]=1; value[beta]=1.4; value[BAO alpha-p*]=1.0")
    a1, a2, a3 = parse_binning("{40:180}*29"), parse_binning("{0,2,4}"), parse_binning("{2.25}")
    c1, c2, c3 = np.meshgrid(a1, a2, a3, indexing="ij")
    r, ell, z = c1.ravel(), c2.ravel(), c3.ravel()
    n = len(r)
    val, has = load_data_vector(os.path.join(golden, "demo", "multi_%d.data" % which), n)
    cov = load_cov(os.path.join(golden, "demo", "multi_%d.cov" % which), n)
    # final cuts of the Multipole type: r, z and lmin <= ell <= lmax (AbsCorrelationData.cc:75-90; the lmax
    # option defaults to 2, src/baofit.cc:247, so the hexadecapole bins are dropped unless lmax = 4 is given)
    keep = np.array([i for i in np.nonzero(has)[0] if 40 <= r[i] <= 200 and 0 <= ell[i] <= lmax], dtype=np.int32)
    return m, make_data(r, ell, z, keep, val[keep], cov[np.ix_(keep, keep)], multipole=True)


def dr11_auto_kspace(root=None, golden=GOLDEN, seed=1966):
    """config/BOSSDR11LyaF_k.ini shape: 50x50 cartesian grid, 4 Mpc/h bins, z=2.34, r in [40,180]
    (N_data = 1515); covariance blob is missing from the reference, so C and d are synthetic."""
    root = root or os.path.join(golden, "models")
    pk_fid = load_table(os.path.join(root, "DR9LyaMocksLCDM_matterpower.dat"))
    pk_nw = load_table(os.path.join(root, "DR9LyaMocksLCDMSB_matterpower.dat"))
    m = kspace_model(pk_fid, pk_nw, rmin=40, rmax=180, zref=2.3, dist_add="-2:0,0:4:2,0")
    m.configure("value[beta]=1.4; value[(1+beta)*bias]=-0.336; fix[gamma-bias]=3.8; fix[gamma-beta]=0;"
                "fix[BAO amplitude]=1; fix[BAO alpha-iso]; value[BAO alpha-p*]=1; fix[gamma-scale]=0;"
                "fix[1+f]=1.966; fix[SigmaNL-perp]=3.26")
    a1, a2, a3 = parse_binning("[0:200]*50"), parse_binning("[0:200]*50"), parse_binning("{2.34}")
    r, mu, z = grid_coordinates(a1, a2, a3, "comoving-cartesian")
    keep = final_cuts(r, mu, z, np.ones(len(r), dtype=bool), rmin=40, rmax=180)
    return m, (r, mu, z, keep)


def dr11_auto_fft(root=None, golden=GOLDEN, ngrid=400, spacing=4., seed=1966, nbins=50, **kw):
    """config/BOSSDR11LyaF_fft.ini: BaoKSpaceFftCorrelationModel, Lya auto-correlation, 3-D grid of
    ngrid^3 cells of 4 Mpc/h, effective-redshift correction and non-linear correction on; 50x50
    cartesian data grid at z = 2.34, r in [40,180] (N_data = 1515). The covariance blob is missing
    from the reference, so the data vector and covariance are synthetic."""
    root = root or os.path.join(golden, "models")
    pk_fid = load_table(os.path.join(root, "DR9LyaMocksLCDM_matterpower.dat"))
    pk_nw = load_table(os.path.join(root, "DR9LyaMocksLCDMSB_matterpower.dat"))
    args = dict(spacing=spacing, nx=ngrid, rmax=180., dilmax=1.5, zref=2.3,
                zcorr=(2.32454, 5.1141e-3, 8.00906e-3), sigma8=0.788, nl_correction=True)
    args.update(kw)
    m = fft_model(pk_fid, pk_nw, **args)
    m.configure("value[beta]=1.4; value[(1+beta)*bias]=-0.351; fix[BAO amplitude]=1; fix[BAO alpha-iso];"
                "value[BAO alpha-p*]=1; fix[gamma-bias]=3.8; fix[gamma-beta]=0; fix[gamma-scale]=0;"
                "fix[1+f]=1.966; fix[SigmaNL-perp]=3.26; value[cont-kc]=0.005; fix[cont-pc]=3")
    a1 = parse_binning("[0:200]*%d" % nbins)
    a2, a3 = a1.copy(), parse_binning("{2.34}")
    r, mu, z = grid_coordinates(a1, a2, a3, "comoving-cartesian")
    keep = final_cuts(r, mu, z, np.ones(len(r), dtype=bool), rmin=40, rmax=180)
    return m, (r, mu, z, keep)


def synthetic_dataset(model, coords, predict, seed=1966, cov_seed=1, rel=0.1):
    """d = model(p0) + L g with the synthetic covariance of SURVEY.md section 8d; ``predict`` maps
    (Model, Data, params[1][npar]) to the prediction [n_data] (the oracle in tests, the CUDA
    library in bench.py)."""
    r, mu, z, keep = coords
    n = len(keep)
    probe = make_data(r, mu, z, keep, np.zeros(n), np.eye(n))
    xi0 = np.asarray(predict(model, probe, model.values[None, :])).reshape(n)
    cov = synthetic_cov(xi0, seed=cov_seed, rel=rel)
    rng = np.random.Generator(np.random.PCG64(seed))
    d = xi0 + np.linalg.cholesky(cov) @ rng.standard_normal(n)
    return make_data(r, mu, z, keep, d, cov), xi0


def write_dataset_files(path_prefix, nbins, keep, dvec, cov):
    """Writes ``<prefix>.data`` / ``<prefix>.cov`` in the reference's text formats (SURVEY.md
    appendix A) for a data set that lives on the kept bins only; the other grid bins get no data."""
    keep = np.asarray(keep)
    with open(path_prefix + ".data", "w") as f:
        np.savetxt(f, np.column_stack([keep, dvec]), fmt="%d %.17g")
    i, j = np.tril_indices(len(keep))
    with open(path_prefix + ".cov", "w") as f:
        np.savetxt(f, np.column_stack([keep[i], keep[j], cov[i, j]]), fmt="%d %d %.17g")
    return path_prefix


def make_golden_tree(root):
    """Lays out tests/golden like the reference checkout (config/ data/ demo/ models/) under
    `root`, expanding the gzip-compressed covariance, so that the .ini recipes run unmodified."""
    import shutil
    os.makedirs(root, exist_ok=True)
    for sub in ("models", "demo"):
        dst = os.path.join(root, sub)
        if not os.path.exists(dst):
            os.symlink(os.path.join(GOLDEN, sub), dst)
    if not os.path.exists(os.path.join(root, "config")):
        shutil.copytree(os.path.join(GOLDEN, "config"), os.path.join(root, "config"))
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    for f in os.listdir(os.path.join(GOLDEN, "data")):
        src = os.path.join(GOLDEN, "data", f)
        if f.endswith(".gz"):
            dst = os.path.join(root, "data", f[:-3])
            if not os.path.exists(dst):
                with gzip.open(src, "rb") as g, open(dst, "wb") as o:
                    shutil.copyfileobj(g, o)
        else:
            dst = os.path.join(root, "data", f)
            if not os.path.exists(dst):
                os.symlink(src, dst)
    return root
