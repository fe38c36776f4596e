"""project-modes-keep (src/baofit.cc:194,597-603): the combined data before final cuts are projected onto the nkeep
largest- or smallest-variance eigenmodes of their covariance (likely::BinnedData::projectOntoModes; dependency absent,
restated from the option's documentation), then finalized. Host only: no GPU needed."""
import os

import numpy as np
import pytest

from baofit_b200 import host_api as H, workloads as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_symmetric_eigenmodes_against_numpy(built):
    rng = np.random.Generator(np.random.PCG64(11))
    for n in (1, 2, 3, 17, 120):
        a = rng.standard_normal((n, n))
        cov = a @ a.T + 0.1 * np.eye(n)
        values, modes = H.symmetric_eigenmodes(cov)
        ref = np.linalg.eigvalsh(cov)
        assert np.all(np.diff(values) >= 0)
        assert np.max(np.abs(values - ref)) < 1e-12 * ref[-1]
        assert np.max(np.abs(modes @ modes.T - np.eye(n))) < 1e-12            # orthonormal rows
        assert np.max(np.abs(modes @ cov @ modes.T - np.diag(values))) < 1e-11 * ref[-1]
    # a matrix that is already tridiagonal / diagonal (reflections skipped), and repeated eigenvalues
    values, modes = H.symmetric_eigenmodes(np.diag([3., 1., 2., 2.]))
    assert np.allclose(values, [1., 2., 2., 3.])
    assert np.max(np.abs(modes @ np.diag([3., 1., 2., 2.]) @ modes.T - np.diag(values))) < 1e-14


def _qso(golden_tree):
    text = open(os.path.join(golden_tree, "config", "BOSSDR11QSOLyaF.ini")).read()
    a1, a2, a3 = W.parse_binning(W.QSO_AXIS1), W.parse_binning(W.QSO_AXIS2), W.parse_binning(W.QSO_AXIS3)
    n = len(a1) * len(a2) * len(a3)
    val, has = W.load_data_vector(os.path.join(W.GOLDEN, "data", "BOSSDR11QSOLyaF.data"), n)
    cov = W.load_cov(os.path.join(W.GOLDEN, "data", "BOSSDR11QSOLyaF.cov.gz"), n)
    assert has.all()
    return text, os.path.join(golden_tree, "config"), val, cov


@pytest.mark.parametrize("nkeep", [60, -100, 512])
def test_project_modes_keep_through_the_ini_option(built, golden_tree, nkeep):
    """config/BOSSDR11QSOLyaF.ini as shipped plus the option: 512 bins with data are projected, then cut to 440."""
    text, base, data, cov = _qso(golden_tree)
    n = len(data)
    plain = H.Session(text, base)
    proj = H.Session(text + "\nproject-modes-keep = %d\n" % nkeep, base)
    kept = plain.kept_indices()
    assert np.array_equal(kept, proj.kept_indices()) and len(kept) == 440    # the final cuts come after the projection
    d0, d1 = plain.data_vector(), proj.data_vector()
    assert np.array_equal(plain.covariance(), proj.covariance())             # the covariance is not touched
    w, v = np.linalg.eigh(cov)                                               # ascending
    sel = v[:, n - nkeep:] if nkeep > 0 else v[:, :-nkeep]
    expected = sel @ (sel.T @ data)
    scale = np.max(np.abs(data))
    assert np.max(np.abs(d1 - expected[kept])) < 1e-10 * scale
    if nkeep == n:
        assert np.max(np.abs(d1 - d0)) < 1e-10 * scale                       # all modes: nothing changes
    else:
        assert np.max(np.abs(d1 - d0)) > 1e-3 * scale
    plain.close(), proj.close()


def test_project_modes_keep_rejects_bad_nkeep(built, golden_tree):
    text, base, _, _ = _qso(golden_tree)
    with pytest.raises(H.HostError):
        H.Session(text + "\nproject-modes-keep = 513\n", base)


def test_fix_mode_scales_option(built, golden_tree, tmp_path):
    """fix-mode-scales (src/baofit.cc:192,514-524,570-572): eigenvalues of the loaded covariance times the scales of a file,
    smallest-variance mode first; the data vector is untouched; the cuts then take the kept block."""
    text, base, data, cov = _qso(golden_tree)
    n = len(data)
    rng = np.random.Generator(np.random.PCG64(3))
    scales = np.exp(0.3 * rng.standard_normal(n))
    np.savetxt(tmp_path / "scales.dat", scales, fmt="%.17g")
    plain = H.Session(text, base)
    fixed = H.Session(text + "\nfix-mode-scales = %s\n" % (tmp_path / "scales.dat"), base)
    kept = plain.kept_indices()
    assert np.array_equal(fixed.data_vector(), plain.data_vector())
    w, v = np.linalg.eigh(cov)
    expected = (v * (w * scales)) @ v.T
    got = fixed.covariance()
    assert np.max(np.abs(got - expected[np.ix_(kept, kept)])) < 1e-10 * np.max(np.abs(cov))
    assert np.max(np.abs(got - plain.covariance())) > 1e-3 * np.max(np.abs(cov))
    np.savetxt(tmp_path / "bad.dat", scales[:-1], fmt="%.17g")
    with pytest.raises(H.HostError):
        H.Session(text + "\nfix-mode-scales = %s\n" % (tmp_path / "bad.dat"), base)
    plain.close(), fixed.close()


def test_scale_zeff_is_where_the_evolved_scale_is_best_determined(built):
    """printScaleZEff (CorrelationAnalyzer.cc:127-167): scale(z) = scale ((1+z)/(1+zref))^gamma has, by error propagation with the
    (scale, gamma) covariance, an error that is stationary at zeff; the returned scale and error are the propagated ones."""
    zref = 2.25
    for scale, ds, gamma, dg, rho in ((1.01, 0.02, 0.6, 0.5, -0.4), (0.98, 0.03, -1.2, 0.8, 0.3), (1.0, 0.015, 2.0, 1.5, 0.0)):
        zeff, se, dse = H.scale_zeff(scale, ds, gamma, dg, rho * ds * dg, zref)

        def propagated(z):
            lr = np.log((1 + z) / (1 + zref))
            ev = ((1 + z) / (1 + zref)) ** gamma
            return scale * ev, ev * np.sqrt(ds ** 2 + 2 * rho * scale * ds * dg * lr + (scale * dg * lr) ** 2)
        s0, e0 = propagated(zeff)
        assert abs(se - s0) < 1e-14 and abs(dse - e0) < 1e-14
        # zeff solves V' + 2 gamma V = 0 for V = (error / evolution)^2 as a function of x = log((1+z)/(1+zref)): the ABSOLUTE error of the
        # evolved scale is stationary there
        err = lambda z: propagated(z)[1]
        h = 1e-4
        slope = (err(zeff + h) - err(zeff - h)) / (2 * h)
        typical = abs(err(zeff + 0.3) - err(zeff)) / 0.3
        assert abs(slope) < 1e-6 * max(typical, 1e-6) + 1e-10, (slope, typical)
    with pytest.raises(H.HostError):
        H.scale_zeff(1.0, 0.0, 0.5, 0.5, 0.0, zref)
