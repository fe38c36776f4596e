"""Host-side C++ mirror of the reference interfaces, driven from the reference's own .ini recipes.
No GPU needed: sessions only touch the device when something is evaluated."""
import os

import numpy as np
import pytest

from baofit_b200 import host_api as H, workloads as W


def ini(golden_tree, name):
    return open(os.path.join(golden_tree, "config", name)).read(), os.path.join(golden_tree, "config")


def test_qso_rspace_session_matches_reference_layout(built, golden_tree):
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base)
    m, d = W.qso_rspace()
    assert s.names == m.names and s.npar == 31
    assert np.array_equal(s.values, m.values) and np.array_equal(s.floating, m.floating)
    assert s.ndata == 440 and np.array_equal(s.kept_indices(), d.keep.desc.n_data and np.asarray(
        [d.keep.desc.index[i] for i in range(440)], dtype=np.int32))
    assert s.scan_size() == 60 * 35
    # floating set of config/BOSSDR11QSOLyaF.ini: beta, delta-v, bias2, both alphas, 15 broadband terms
    assert s.floating.sum() == 20


def test_qso_kspace_session_layout(built, golden_tree):
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF_k.ini")
    s = H.Session(text, base)
    m, _ = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    assert s.names == m.names and s.npar == 29 and s.floating.sum() == 20
    assert np.array_equal(s.values, m.values)
    assert s.names[11:13] == ["BAO alpha-parallel", "BAO alpha-perp"]


def test_demo_rmu_session(built, golden_tree):
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    s = H.Session(text, base)
    assert s.ndata == 800 and [n for n, f in zip(s.names, s.floating) if f] == ["beta", "BAO alpha-parallel", "BAO alpha-perp"]


def test_model_errors_mirror_reference_messages(built, golden_tree):
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    with pytest.raises(H.HostError) as e:
        H.Session(text.replace("fiducial = DR9LyaMocksNLeffPk", "fiducial = nosuchmodel"), base)
    assert e.value.code == -2 and "BaoCorrelationModel: error while reading model interpolation data." in str(e.value)
    with pytest.raises(H.HostError) as e:
        H.Session(text.replace("dist-add = rP,rT=0:2,-3:1", "dist-add = rP,rT=0:2,x"), base)
    assert e.value.code == -2 and "BroadbandModel: badly formatted parameter specification" in str(e.value)
    with pytest.raises(H.HostError) as e:
        H.Session(text.replace("data = ../data/BOSSDR11QSOLyaF", "data = ../data/missing"), base)
    assert e.value.code == -3 and "loadCorrelationData: Unable to open" in str(e.value)
    with pytest.raises(H.HostError) as e:
        H.Session(text, base, extra_config="fix[no such parameter]=1")
    assert e.value.code == -2


def test_parameter_script_language(built, golden_tree):
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base, extra_config="fix[dist*]=0; value[beta]=1.3; error[bias2]=0.5; release[gamma-bias]")
    assert s.floating.sum() == 20 - 15 + 1
    assert s.values[s.names.index("beta")] == 1.3 and s.errors[s.names.index("bias2")] == 0.5


def test_minimiser_on_quadratic_and_rosenbrock(built):
    A = np.array([[4.0, 1.0, 0.5], [1.0, 3.0, 0.2], [0.5, 0.2, 2.0]])
    x0 = np.array([1.0, -2.0, 0.5])

    def quad(p):
        d = p[:, :3] - x0
        return 0.5 * np.einsum("bi,ij,bj->b", d, A, d) + 7.0

    r = H.minimize(quad, [0, 0, 0, 9.0], [1, 1, 1, 1], [1, 1, 1, 0])
    assert r["status"] == 0 and np.allclose(r["values"][:3], x0, atol=1e-6) and r["values"][3] == 9.0
    assert np.allclose(r["errors"][:3], np.sqrt(np.diag(np.linalg.inv(A))), rtol=1e-4)

    def rosen(p):
        return 100 * (p[:, 1] - p[:, 0] ** 2) ** 2 + (1 - p[:, 0]) ** 2

    r = H.minimize(rosen, [-1.2, 1.0], [0.1, 0.1], [1, 1])
    assert r["fval"] < 1e-5 and np.allclose(r["values"], [1, 1], atol=5e-3)


def test_lm_driver_on_cpu_residuals():
    """likely::LMFit (Gauss-Newton normal equations + Marquardt damping + trust region) over a
    host residual function: a 3-parameter exponential fit converges to the Newton driver's minimum
    with far fewer evaluations; a Gaussian prior is honoured."""
    rng = np.random.Generator(np.random.PCG64(4))
    t = np.linspace(0, 4, 60)
    truth = np.array([2.0, 0.7, 0.3])
    data = truth[0] * np.exp(-truth[1] * t) + truth[2] + 0.01 * rng.standard_normal(len(t))

    def resid(p):
        return (p[:, 0:1] * np.exp(-p[:, 1:2] * t[None, :]) + p[:, 2:3] - data[None, :]) / 0.01

    start, err = [1.0, 1.0, 0.0], [0.5, 0.3, 0.2]
    lm = H.lm_minimize(resid, len(t), start, err, [1, 1, 1])
    nw = H.minimize(lambda p: 0.5 * (resid(p) ** 2).sum(axis=1), start, err, [1, 1, 1])
    assert lm["status"] == 0 and nw["status"] == 0
    assert abs(lm["fval"] - nw["fval"]) < 1e-6 and np.allclose(lm["values"], nw["values"], atol=1e-5)
    assert lm["nevaluations"] < nw["nevaluations"]
    pr = H.lm_minimize(resid, len(t), start, err, [1, 1, 1], prior_centre=[0, 0, 0.5], prior_sigma=[0, 0, 1e-3])
    nwp = H.minimize(lambda p: 0.5 * (resid(p) ** 2).sum(axis=1) + 0.5 * ((p[:, 2] - 0.5) / 1e-3) ** 2, start, err, [1, 1, 1])
    assert abs(pr["fval"] - nwp["fval"]) < 1e-5 and np.allclose(pr["values"], nwp["values"], atol=1e-5)
    assert abs(pr["values"][2] - 0.5) < 0.05 and pr["fval"] > lm["fval"]


def _qso_files(tmp_path):
    """BOSSDR11QSOLyaF data/cov over all 512 bins, rewritten under tmp_path in the loader's optional formats."""
    a1, a2, a3 = W.parse_binning(W.QSO_AXIS1), W.parse_binning(W.QSO_AXIS2), W.parse_binning(W.QSO_AXIS3)
    n = len(a1) * len(a2) * len(a3)
    val, has = W.load_data_vector(os.path.join(W.GOLDEN, "data", "BOSSDR11QSOLyaF.data"), n)
    cov = W.load_cov(os.path.join(W.GOLDEN, "data", "BOSSDR11QSOLyaF.cov.gz"), n)
    assert has.all()
    icov = np.linalg.inv(cov)
    prefix = str(tmp_path / "X")
    idx = np.arange(n)
    np.savetxt(prefix + ".data", np.column_stack([idx, val]), fmt="%d %.17g")
    np.savetxt(prefix + ".wdata", np.column_stack([idx, icov @ val]), fmt="%d %.17g")
    i, j = np.tril_indices(n)
    np.savetxt(prefix + ".cov", np.column_stack([i, j, cov[i, j]]), fmt="%d %d %.17g")
    np.savetxt(prefix + ".icov", np.column_stack([i, j, 0.5 * (icov + icov.T)[i, j]]), fmt="%d %d %.17g")
    centres = np.array([[x, y, zz] for x in a1 for y in a2 for zz in a3])
    return prefix, centres, (a1, a2, a3)


def test_weighted_data_inverse_covariance_and_custom_grid(built, golden_tree, tmp_path):
    """loadCorrelationData's optional inputs (baofit/AbsCorrelationData.cc:156-277): .wdata (Cinv.d), .icov with
    bins dropped by the final cuts (pruning marginalises: sub-block of the covariance), custom bin centres."""
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    prefix, centres, axes = _qso_files(tmp_path)
    text = text.replace("data = ../data/BOSSDR11QSOLyaF", "data = " + prefix)
    s0 = H.Session(text, base)
    d0, c0 = s0.data_vector(), s0.covariance()
    assert s0.ndata == 440
    sw = H.Session(text + "\nload-wdata = yes\n", base)
    assert np.allclose(sw.data_vector(), d0, rtol=1e-8, atol=1e-12 * np.abs(d0).max()) and np.array_equal(sw.covariance(), c0)
    si = H.Session(text + "\nload-icov = yes\n", base)
    assert np.allclose(si.covariance(), c0, rtol=1e-6, atol=1e-9 * np.abs(c0).max())
    assert np.array_equal(si.data_vector(), d0)
    swi = H.Session(text + "\nload-icov = yes\nload-wdata = yes\n", base)
    assert np.allclose(swi.data_vector(), d0, rtol=1e-6, atol=1e-9 * np.abs(d0).max())
    # custom centres equal to the grid's: nothing changes; halved separations: the cuts see the new radii
    np.savetxt(prefix + ".grid", np.column_stack([np.arange(len(centres)), centres]), fmt="%d %.17g %.17g %.17g")
    sg = H.Session(text + "\ncustom-grid = yes\n", base)
    assert sg.ndata == 440 and all(np.array_equal(a, b) for a, b in zip(sg.kept_coordinates(), s0.kept_coordinates()))
    half = centres * np.array([0.5, 0.5, 1.0])
    np.savetxt(prefix + ".grid", np.column_stack([np.arange(len(centres)), half]), fmt="%d %.17g %.17g %.17g")
    sh = H.Session(text + "\ncustom-grid = yes\n", base)
    r = np.hypot(half[:, 0], half[:, 1])
    mu = half[:, 0] / r
    keep = W.final_cuts(r, mu, half[:, 2], np.ones(len(r), bool), rmin=40, rmax=180)
    rk, muk, zk = sh.kept_coordinates()
    assert sh.ndata == len(keep) and np.array_equal(sh.kept_indices(), keep)
    assert np.allclose(rk, r[keep], rtol=1e-15) and np.allclose(muk, mu[keep], rtol=1e-15, atol=1e-16)
    # every grid bin needs a centre
    np.savetxt(prefix + ".grid", np.column_stack([np.arange(len(centres) - 1), half[:-1]]), fmt="%d %.17g %.17g %.17g")
    with pytest.raises(H.HostError, match="number of custom bins must match"):
        H.Session(text + "\ncustom-grid = yes\n", base)


def test_options_outside_the_reference_list_are_rejected(built, golden_tree):
    """boost::program_options rejects undeclared keys (src/baofit.cc:46-300, 313-340); options that would select a model or a
    data treatment this build does not have are refused instead of being ignored."""
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    with pytest.raises(H.HostError) as e:
        H.Session(text + "\nrmaxx = 180\n", base)
    assert e.value.code == -1 and "unrecognised option 'rmaxx'" in str(e.value)
    for extra, what in (("xi-points = 50,100,150", "XiCorrelationModel"), ("n-spline = 10", "PkCorrelationModel"),
                        ("fix-aln-cov = yes", "fix-aln-cov"), ("scalar-weights = yes", "scalar-weights")):
        with pytest.raises(H.HostError) as e:
            H.Session(text + "\n" + extra + "\n", base)
        assert e.value.code == -1 and what in str(e.value)
    # analysis-control options of the reference are accepted: they are arguments of the entry points here
    s = H.Session(text + "\nbootstrap-trials = 10\noutput-prefix = x_\nrandom-seed = 7\ncheck-posdef = yes\nndump = 0\n", base)
    assert s.ndata == 440
    s.close()
