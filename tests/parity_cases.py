"""Named parity cases shared by tools/gen_oracle_fixtures.py (which runs the slow CPU oracle once and stores
its answers under tests/golden/oracle_<case>.npz) and tests/test_oracle_fixtures.py (which holds the CUDA
path against those answers at 1e-10). Pure descriptor plumbing: no hot-path arithmetic lives here.

Cases
  cfg2 .. cfg5      BASELINE configs 2-5 at full size (SURVEY.md section 8 size table), B = 256
  opt_*             one k-space model per optional term of D(k, mu) and of the linear-bias block that no
                    BASELINE config switches on (baofit/BaoKSpaceCorrelationModel.cc:229-275,441-455,
                    baofit/AbsCorrelationModel.cc:120-128), B = 8: vectors 0-3 share the key parameters
                    (shared-transform path), vectors 4-7 float them (per-vector transform path)
  metal_*           the metal-model branches the headline workload does not reach: templates tabulated per
                    grid bin (baofit/MetalCorrelationModel.cc:238-266), the toy model (:303-346), and the
                    five-combination cross model with CIV (:128-134)
"""
import os

import numpy as np

from baofit_b200 import desc as D
from baofit_b200 import workloads as W

GOLDEN = W.GOLDEN
SWEEP_AP = {"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)}

BIG = {"cfg2": 2002, "cfg3": 2003, "cfg4": 2004, "cfg5": 2005}
OPTIONS = ["opt_uv_cross", "opt_uv_auto", "opt_hcd_cross", "opt_hcd_alt_auto", "opt_binsmooth", "opt_binsmooth_alt",
           "opt_gauss_lorentz", "opt_radiation_dm", "opt_combined", "opt_nlcorr_fit"]
METALS = ["metal_auto_table", "metal_auto_table_civ_dm", "metal_toy", "metal_cross_civ"]
ALL = list(BIG) + OPTIONS + METALS


def fixture_path(name):
    return os.path.join(GOLDEN, "oracle_%s.npz" % name)


def _qso_grid():
    a1, a2, a3 = W.parse_binning(W.QSO_AXIS1), W.parse_binning(W.QSO_AXIS2), W.parse_binning(W.QSO_AXIS3)
    r, mu, z = W.grid_coordinates(a1, a2, a3, "comoving-cartesian")
    return a1, a2, r, mu, z


def _qso_data(with_grid=False):
    """Kept bins, data vector and covariance of data/BOSSDR11QSOLyaF on its 32 x 16 grid (N_d = 440)."""
    a1, a2, r, mu, z = _qso_grid()
    n = len(r)
    val, has = W.load_data_vector(os.path.join(GOLDEN, "data", "BOSSDR11QSOLyaF.data"), n)
    keep = W.final_cuts(r, mu, z, has, rmin=40, rmax=180)
    cov = W.load_cov(os.path.join(GOLDEN, "data", "BOSSDR11QSOLyaF.cov.gz"), n)[np.ix_(keep, keep)]
    return W.make_data(r, mu, z, keep, val[keep], cov, with_grid=with_grid)


def _pk():
    root = os.path.join(GOLDEN, "models")
    return (W.load_table(os.path.join(root, "DR9LyaMocksLCDM_matterpower.dat")),
            W.load_table(os.path.join(root, "DR9LyaMocksLCDMSB_matterpower.dat")))


BASE_CFG = ("value[beta]=1.1; fix[gamma-beta]=0; fix[BAO amplitude]=1; fix[BAO alpha-iso]; value[BAO alpha-p*]=1;"
            "fix[gamma-scale]=0; fix[1+f]=1.966; fix[SigmaNL-perp]=3.26")
CROSS_CFG = "; value[bias2]=3.64; fix[beta2*bias2]=0.962524; value[delta-v]=0"


def _synthetic_auto_metals(ngrid, civ, seed=77):
    """Templates tabulated per grid bin + the .grid redshifts (three distinct values, so that the metal
    normalisation factors need more than one redshift class)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ncomb = 20 if civ else 14
    a1, a2, r, mu, z = _qso_grid()
    base = 1e-3 * np.exp(-0.5 * ((r[None, :] - rng.uniform(40, 160, size=(ncomb * 3, 1))) / 25.0) ** 2)
    auto = base * rng.uniform(-1, 1, size=(ncomb * 3, 1)) * (1 + 0.3 * mu[None, :] ** 2)
    zgrid = 2.30 + 0.05 * rng.integers(0, 3, size=(ncomb, ngrid))
    return dict(kind=D.BF_METAL_AUTO_TABLE, civ=civ, zgrid=zgrid, auto=auto)


def build(name, xi0=None, predict=None):
    """Returns (model, data, sweep, B, xi0). Synthetic data sets (cfg3, cfg5) are rebuilt from `xi0` (stored
    in the fixture) or, when it is None, from `predict` (the oracle, in the generator)."""
    pk_fid, pk_nw = _pk()
    if name == "cfg2":
        m, d = W.dr9_kspace()
        return m, d, {"BAO alpha-parallel": (0.7, 1.4), "BAO alpha-perp": (0.6, 1.3)}, 256, None
    if name == "cfg4":
        m, d = W.qso_kspace()
        return m, d, dict(SWEEP_AP, **{"delta-v": (-300., 300.)}), 256, None
    if name in ("cfg3", "cfg5"):
        m, coords = W.dr11_auto_kspace() if name == "cfg3" else W.dr11_auto_fft()
        pr = (lambda mm, dd, p: xi0) if xi0 is not None else predict
        d, x0 = W.synthetic_dataset(m, coords, pr)
        return m, d, dict(SWEEP_AP), 256, x0
    sweep = dict(SWEEP_AP)
    kw = dict(rmin=40, rmax=180, zref=2.3, dist_add="rP,rT=0:2,-3:1")
    cfg = BASE_CFG
    cross = name in ("opt_uv_cross", "opt_hcd_cross", "opt_radiation_dm", "opt_combined", "opt_gauss_lorentz", "metal_cross_civ")
    with_grid = False
    if name == "opt_uv_cross":
        kw.update(uvfluctuation=True)
    elif name == "opt_uv_auto":
        kw.update(uvfluctuation=True, dist_add="-2:0,0:4:2,0")
    elif name == "opt_hcd_cross":
        kw.update(hcd_model=True)
    elif name == "opt_hcd_alt_auto":
        kw.update(hcd_model_alt=True, dist_add="-2:0,0:4:2,0")
    elif name == "opt_binsmooth":
        kw.update(bin_smooth=True, dist_add="-2:0,0:4:2,0")
    elif name == "opt_binsmooth_alt":
        kw.update(bin_smooth_alt=True, dist_add="-2:0,0:4:2,0", nl_correction=True, zeff=2.4)
    elif name == "opt_gauss_lorentz":
        kw.update(smooth_gauss=True, smooth_lorentz=True)
    elif name == "opt_radiation_dm":
        a1, a2, r, mu, z = _qso_grid()
        kw.update(radiation_model=True, dist_matrix=W.synthetic_dist_matrix(a1, a2), dm_dist_add="0:1,0:2:2,0")
        cfg += "; value[Rad strength]=0.4; value[Rad anisotropy]=0.3; value[Rad mean free path]=150; value[Rad quasar lifetime]=20"
        with_grid = True
    elif name == "opt_combined":
        kw.update(combined_bias=True, combined_scale=True)
        sweep = {"BAO alpha-parallel": (0.8, 1.3), "BAO aperp/apar": (0.8, 1.2)}
    elif name == "opt_nlcorr_fit":
        kw.update(fit_nl_correction=True, nl_broadband=True, dist_add="-2:0,0:4:2,0", dist_mul="0:1,0:2:2,0")
    elif name == "metal_auto_table":
        kw.update(metal=_synthetic_auto_metals(512, civ=False), dist_add="-2:0,0:4:2,0")
    elif name == "metal_auto_table_civ_dm":
        a1, a2, r, mu, z = _qso_grid()
        kw.update(metal=_synthetic_auto_metals(512, civ=True), dist_add="-2:0,0:4:2,0", dist_matrix=W.synthetic_dist_matrix(a1, a2))
        with_grid = True
    elif name == "metal_toy":
        kw.update(metal=dict(kind=D.BF_METAL_TOY), dist_add="-2:0,0:4:2,0")
    elif name == "metal_cross_civ":
        mt = W.synthetic_cross_metals(ngrid=512)
        c5 = np.concatenate([mt["cross"], 0.5 * mt["cross"][3:6][:, ::-1, :]], axis=0)  # a fifth (CIV) combination
        rng = np.random.Generator(np.random.PCG64(78))
        mt.update(civ=True, cross=c5, zgrid=2.30 + 0.05 * rng.integers(0, 3, size=(5, 512)))
        kw.update(metal=mt)
    else:
        raise ValueError("unknown parity case " + name)
    m = W.kspace_model(pk_fid, pk_nw, cross=cross, **kw)
    m.configure(cfg + (CROSS_CFG if cross else ""))
    return m, _qso_data(with_grid), sweep, 8, None


def make_params(name, m, sweep, B):
    """Seeded parameter vectors of a case. Option / metal cases: the second half of the batch also floats the
    parameters the shared-transform key is made of, the first half keeps them at their start values."""
    seed = BIG.get(name, 3000 + ALL.index(name))
    rng = np.random.Generator(np.random.PCG64(seed))
    params = m.random_batch(B, rng, sweep=sweep)
    key_names = ("bin scale", "HCD", "UV", "smooth scale", "nlcorr", "SigmaNL-perp", "1+f")
    for j, n in enumerate(m.names):
        if n.startswith("bias Si") or n.startswith("bias CIV"):
            params[:, j] *= rng.uniform(0.5, 2.0, size=B)
        if name not in BIG and n.startswith(key_names):
            params[: B // 2, j] = m.values[j]
            if n in ("SigmaNL-perp", "1+f"):
                params[B // 2:, j] = m.values[j] * rng.uniform(0.9, 1.1, size=B - B // 2)
    return params
