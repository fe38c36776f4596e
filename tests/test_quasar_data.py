"""data-format = quasar (baofit/QuasarCorrelationData.cc) and BASELINE configs 1-2
(config/BOSSDR9LyaF.ini, config/BOSSDR9LyaF_k.ini): observing coordinates -> (r, mu) through the
cosmology stand-in, ll / separation cuts, and the published chi-square surface data/BOSSDR9LyaF.scan
as the end-to-end pin of that chain."""
import os

import numpy as np
import pytest

import oracle
from baofit_b200 import host_api as H, workloads as W
from util import RTOL, rel_err


def _session(golden_tree, name, extra=""):
    base = os.path.join(golden_tree, "config")
    return H.Session(open(os.path.join(base, name)).read(), base, extra)


def test_quasar_transform_known_values():
    """Flat LCDM: D_C(z) against an independent quadrature; r, mu of a bin from its definition."""
    c = W.LambdaCdmRadiationUniverse(0.27, 0.7, omega_radiation=0.0)
    z = np.linspace(0, 2.5, 200001)
    ref = 2997.92458 * np.trapezoid(1 / np.sqrt(0.73 + 0.27 * (1 + z) ** 3), z)
    assert abs(c.los_distance(2.5) - ref) < 1e-6 * ref
    r, mu, zz, ll, sep = W.quasar_grid_coordinates(np.array([0.02]), np.array([35.0]), 10.0, np.array([2.5]), c)
    dlos = c.los_distance(3.5 * np.exp(0.01) - 1) - c.los_distance(3.5 / np.exp(0.01) - 1)
    dperp = c.los_distance(2.5) * (35.0 + 100. / 12 / 35.0) * np.pi / 10800
    assert abs(r[0] - np.hypot(dlos, dperp)) < 1e-10 and abs(mu[0] - dlos / r[0]) < 1e-13


def test_host_quasar_data_matches_python_restatement(built, golden_tree):
    """Host C++ QuasarCorrelationData vs the numpy restatement: identical kept-bin mask (bit-exact
    indices), coordinates to rounding, parameter layout of SURVEY.md appendix C config 1."""
    s = _session(golden_tree, "BOSSDR9LyaF.ini")
    m, d = W.dr9_rspace()
    assert s.npar == 22 and int(s.floating.sum()) == 14 and s.names == m.names
    assert np.array_equal(s.floating, m.floating) and np.allclose(s.values, m.values)
    assert np.array_equal(s.kept_indices(), d.keep._arrays[3])
    r, mu, z = s.kept_coordinates()
    assert np.max(np.abs(r - d.keep._arrays[0]) / r) < 1e-12
    assert np.max(np.abs(mu - d.keep._arrays[1])) < 1e-12 and np.array_equal(z, d.keep._arrays[2])
    assert r.min() >= 50 and r.max() <= 190
    s2 = _session(golden_tree, "BOSSDR9LyaF_k.ini")
    m2, _ = W.dr9_kspace()
    assert s2.npar == 20 and int(s2.floating.sum()) == 14 and s2.names == m2.names  # appendix C config 2
    assert np.array_equal(s2.kept_indices(), d.keep._arrays[3])


@pytest.mark.gpu
def test_gpu_dr9_rspace_evaluate(built, golden_tree, ctx):
    from baofit_b200 import capi
    m, d = W.dr9_rspace()
    rng = np.random.Generator(np.random.PCG64(2))
    params = m.random_batch(64, rng, sweep={"BAO alpha-parallel": (0.7, 1.4), "BAO alpha-perp": (0.6, 1.3)})
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, params)
    ref, rst = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, nthreads=8)
    assert np.array_equal(st, rst) and rel_err(chi2, ref) < RTOL
    gm.close(), gd.close()
    # the host session over the unmodified .ini adds the Gaussian priors of :80-84 on top of chi2/2
    s = _session(golden_tree, "BOSSDR9LyaF.ini")
    f = s.evaluate(params[:8])
    prior = 0.5 * ((params[:8, m.index("beta")] - 1.4) / 0.4) ** 2 + 0.5 * ((params[:8, m.index("gamma-bias")] - 3.8) / 1.0) ** 2
    assert np.max(np.abs(f - (0.5 * ref[:8] + prior)) / f) < 1e-10


@pytest.mark.gpu
def test_gpu_dr9_kspace_evaluate(built, ctx):
    from baofit_b200 import capi
    m, d = W.dr9_kspace()
    rng = np.random.Generator(np.random.PCG64(4))
    params = m.random_batch(4, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.7, 1.3)})
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, params)
    ref, rst = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, nthreads=16)
    assert np.array_equal(st, rst) and rel_err(chi2[rst == 0], ref[rst == 0]) < RTOL
    gm.close(), gd.close()


@pytest.mark.gpu
def test_gpu_dr9_scan_against_published_surface(built, golden_tree):
    """data/BOSSDR9LyaF.scan (Slosar et al. 2013): 50 x 50 published chi-square values. The whole
    chain is pinned here: quasar coordinates through the cosmology, r-space auto-correlation model,
    broadband, Gaussian priors, dense-covariance chi-square, refit at every grid point."""
    s = _session(golden_tree, "BOSSDR9LyaF.ini")
    scan = s.parameter_scan()
    pub = np.loadtxt(os.path.join(golden_tree, "data", "BOSSDR9LyaF.scan"))
    assert len(scan["chi2"]) == 2500 == len(pub)
    ia, ip = s.names.index("BAO alpha-parallel"), s.names.index("BAO alpha-perp")
    assert np.allclose(scan["points"][:, ip], pub[:, 0], atol=5e-7) and np.allclose(scan["points"][:, ia], pub[:, 1], atol=5e-7)
    dchi2 = scan["chi2"] - scan["best_chi2"]
    diff = dchi2 - pub[:, 2]
    print("DR9 scan: chi2_min %.4f, diff min %.4g max %.4g, within 1.5e-3: %.3f, within 1e-2: %.3f" %
          (scan["best_chi2"], diff.min(), diff.max(), np.mean(np.abs(diff) < 1.5e-3), np.mean(np.abs(diff) < 1e-2)))
    assert np.all(scan["status"] != 2)
    assert diff.max() < 2.5e-3, (diff.max(), int(np.argmax(diff)))
    assert np.mean(np.abs(diff) < 2.5e-3) >= 0.80
