"""Fits, scans and toy MC through the host-side classes on the GPU, against the oracle and the
reference's published chi-square surface."""
import os

import numpy as np
import pytest

import oracle
from baofit_b200 import host_api as H, workloads as W

pytestmark = pytest.mark.gpu


def ini(golden_tree, name):
    return open(os.path.join(golden_tree, "config", name)).read(), os.path.join(golden_tree, "config")


def test_session_evaluate_matches_oracle(built, golden_tree):
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base)
    m, d = W.qso_rspace()
    rng = np.random.Generator(np.random.PCG64(3))
    params = m.random_batch(16, rng)
    f = s.evaluate(params)
    ref, _ = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params)
    assert np.max(np.abs(2 * f - ref) / ref) < 1e-10


def test_fit_parity_gpu_vs_oracle_same_minimiser(built, golden_tree):
    """north_star: best-fit parameters within 1e-6 of their errors. The same minimiser code runs
    once over the CUDA objective and once over the CPU oracle objective."""
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    s = H.Session(text, base)
    gpu = s.fit()
    m, d = W.demo_rmu()
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    cpu = H.minimize(lambda p: 0.5 * oracle.chi2_batch(om, od, p, nthreads=8)[0], s.values, s.errors, s.floating)
    assert gpu["status"] == 0 and cpu["status"] == 0
    fl = s.floating
    assert np.max(np.abs(gpu["values"][fl] - cpu["values"][fl]) / cpu["errors"][fl]) < 1e-6
    assert abs(gpu["fval"] - cpu["fval"]) < 1e-9 * abs(cpu["fval"])
    # the demo data were generated at beta = 1.4, alphas = 1: the fit must recover them within errors
    truth = {"beta": 1.4, "BAO alpha-parallel": 1.0, "BAO alpha-perp": 1.0}
    for name, v in truth.items():
        i = s.names.index(name)
        assert abs(gpu["values"][i] - v) < 4 * gpu["errors"][i]


def test_qso_rspace_best_fit_matches_published(built, golden_tree):
    """config/BOSSDR11QSOLyaF.ini: published best fit alpha_par = 1.0418,
    alpha_perp = 0.9302, chi2_min = 426.373."""
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base)
    r = s.fit()
    assert r["status"] == 0
    assert abs(2 * r["fval"] - 426.373) < 2e-3
    assert abs(r["values"][s.names.index("BAO alpha-parallel")] - 1.0418) < 1e-3
    assert abs(r["values"][s.names.index("BAO alpha-perp")] - 0.9302) < 1e-3


def test_qso_rspace_scan_against_published_surface(built, golden_tree):
    """data/BOSSDR11QSOLyaF.scan: 60 x 35 published delta-chi2 values (3 decimals). Acceptance
    (SURVEY.md section 4 iv): never above the published value by more than 1.5e-3 (the reference's
    Minuit stopped in higher local minima at some points, never lower), within 1.5e-3 on >= 80 %."""
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base)
    scan = s.parameter_scan()
    pub = np.loadtxt(os.path.join(golden_tree, "data", "BOSSDR11QSOLyaF.scan"))
    assert len(scan["chi2"]) == 2100 == len(pub)
    ia, ip = s.names.index("BAO alpha-parallel"), s.names.index("BAO alpha-perp")
    # our grid order: parameter order (alpha-parallel outer, alpha-perp fastest) == .scan row order
    assert np.allclose(scan["points"][:, ip], pub[:, 0], atol=5e-7) and np.allclose(scan["points"][:, ia], pub[:, 1], atol=5e-7)
    dchi2 = scan["chi2"] - scan["best_chi2"]
    diff = dchi2 - pub[:, 2]
    assert abs(scan["best_chi2"] - 426.373) < 2e-3
    assert np.all(scan["status"] != 2)
    assert diff.max() < 1.5e-3, (diff.max(), int(np.argmax(diff)))
    assert np.mean(np.abs(diff) < 1.5e-3) >= 0.80, np.mean(np.abs(diff) < 1.5e-3)
    k = int(np.argmin(dchi2))
    assert abs(scan["points"][k, ip] - 0.9356) < 1e-3 and abs(scan["points"][k, ia] - 1.0353) < 1e-3


def test_toymc_is_shard_independent_and_unbiased(built, golden_tree):
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    text = text.replace("load-icov = yes", "")
    # a covariance is needed for toy noise: diagonal 1e-11 (the inverse of rmu.icov)
    tmp = os.path.join(golden_tree, "data", "rmu_cov")
    with open(tmp + ".data", "w") as f:
        f.write(open(os.path.join(golden_tree, "demo", "rmu.data")).read())
    with open(tmp + ".cov", "w") as f:
        for i in range(800):
            f.write("%d %d 1e-11\n" % (i, i))
    text = text.replace("data = ../demo/rmu", "data = ../data/rmu_cov")
    s = H.Session(text, base)
    a = s.toymc(0, 12, seed=7)
    b1, b2 = s.toymc(0, 5, seed=7), s.toymc(5, 7, seed=7)
    assert np.array_equal(a["fitted"], np.vstack([b1["fitted"], b2["fitted"]]))
    assert np.all(a["status"] == 0)
    ib = s.names.index("beta")
    assert abs(a["fitted"][:, ib].mean() - 1.4) < 0.05
    assert 600 < a["chi2"].mean() < 1000  # ~ N_data - n_float
    # SamplingOutput file (CorrelationAnalyzer.cc:375-450): "npar nsave nfit", errors, input fit, one line per toy
    # with the sample's parameters, chi2 and nsave x (mono, quad, hexa) of its model
    best = s.fit()
    path = os.path.join(golden_tree, "data", "toymc_save.dat")
    s.write_samples(path, best["values"], best["errors"], 2 * best["fval"], a["fitted"], a["chi2"], nsave=5, zsave=2.25)
    lines = open(path).read().splitlines()
    assert lines[0].split() == [str(s.npar), "5", "1"] and len(lines) == 3 + 12
    assert len(lines[1].split()) == s.npar
    rows = np.array([[float(t) for t in ln.split()] for ln in lines[2:]])
    assert rows.shape == (13, s.npar + 1 + 15)
    assert np.allclose(rows[0, :s.npar], best["values"], rtol=1e-9, atol=1e-12) and abs(rows[0, s.npar] - 2 * best["fval"]) < 1e-6 * rows[0, s.npar]
    assert np.allclose(rows[1:, :s.npar], a["fitted"], rtol=1e-9, atol=1e-12) and np.allclose(rows[1:, s.npar], a["chi2"], rtol=1e-9)
    for k in (0, 3, 11):
        mp = s.dump_model(a["fitted"][k], 5, 2.25)
        assert np.allclose(rows[1 + k, s.npar + 1:], mp.ravel(), rtol=1e-9, atol=1e-15)
    s.write_samples(path, best["values"], best["errors"], 2 * best["fval"], a["fitted"][:2], a["chi2"][:2])
    assert len(open(path).read().splitlines()) == 5 and open(path).readline().split()[1] == "0"
    # refit pass (refit-config): every toy fitted again with beta fixed at its start value; same toys, so the first
    # pass is unchanged and the constrained chi2 is never below the free one
    # (the session has fitted the combined sample by now, so the truth vector is the best-fit prediction from here on)
    a2 = s.toymc(0, 12, seed=7)
    rf = s.toymc(0, 12, seed=7, refit_config="fix[beta]=1.4")
    assert np.array_equal(rf["fitted"], a2["fitted"]) and np.all(rf["restatus"] == 0)
    assert np.all(rf["refitted"][:, ib] == 1.4) and np.all(rf["rechi2"] >= rf["chi2"] - 1e-6)
    best2 = s.fit("fix[beta]=1.4")
    s.write_samples(path, best["values"], best["errors"], 2 * best["fval"], rf["fitted"], rf["chi2"], nsave=3, zsave=2.25,
                    refit=(best2["values"], best2["errors"], 2 * best2["fval"], rf["refitted"], rf["rechi2"]))
    lines = open(path).read().splitlines()
    assert lines[0].split() == [str(s.npar), "3", "2"] and len(lines[1].split()) == 2 * s.npar
    rows = np.array([[float(t) for t in ln.split()] for ln in lines[2:]])
    assert rows.shape == (13, 2 * (s.npar + 1) + 2 * 9)
    assert np.allclose(rows[1:, s.npar + 1:2 * s.npar + 1], rf["refitted"], rtol=1e-9, atol=1e-12)
    assert np.allclose(rows[1:, 2 * s.npar + 1], rf["rechi2"], rtol=1e-9)


def test_dump_residuals_and_model_against_oracle(built, golden_tree, tmp_path):
    """CorrelationAnalyzer::dumpResiduals (predictions + central-difference gradients, one GPU batch)
    and dumpModel (monopole / quadrupole / hexadecapole of the model) against the oracle; the
    multipoles of the demo model reproduce demo/multi.theory."""
    text, base = ini(golden_tree, "BOSSDR11QSOLyaF.ini")
    s = H.Session(text, base)
    m, d = W.qso_rspace()
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    errors = np.where(s.floating, s.errors, 0.0)
    tab = s.dump_residuals(s.values, errors, str(tmp_path / "resid.dat"), gradients=True)
    assert tab.shape == (s.ndata, 10 + s.npar)
    assert np.array_equal(tab[:, 0].astype(int), s.kept_indices())
    pred, _ = oracle.predict(om, od, s.values)
    assert np.max(np.abs(tab[:, 7] - pred)) < 2e-12 * np.max(np.abs(pred))  # 12 printed digits
    assert np.allclose(tab[:, 8], d.keep._arrays[4], rtol=1e-11) and np.allclose(tab[:, 9], np.sqrt(np.diag(d.keep._arrays[5])), rtol=1e-11)
    for j in np.nonzero(s.floating)[0][:6]:
        hi, lo = s.values.copy(), s.values.copy()
        hi[j] += 0.05 * errors[j]
        lo[j] -= 0.05 * errors[j]
        g = (oracle.predict(om, od, hi)[0] - oracle.predict(om, od, lo)[0]) / (0.1 * errors[j])
        assert np.max(np.abs(tab[:, 10 + j] - g)) < 1e-6 * np.max(np.abs(g)) + 1e-14
    assert np.all(tab[:, 10 + np.nonzero(~s.floating)[0]] == 0)
    # dumpModel on the demo model: the 29 radii of demo/multi.theory are {40:180}*29 = rmin..rmax of the ini
    text2, base2 = ini(golden_tree, "rmu_to_bao.ini")
    s2 = H.Session(text2.replace("rmax = 200", "rmax = 180"), base2)
    mp = s2.dump_model(s2.values, 29, 2.25)
    mt = np.loadtxt(os.path.join(golden_tree, "demo", "multi.theory"))[:, 1].reshape(29, 3)
    assert np.max(np.abs(mp - mt)) < 1e-14


def test_mcmc_walkers_sample_the_fit_posterior(built, golden_tree):
    """CorrelationFitter::mcmc with lock-step walkers: the sample mean and covariance of the floating
    parameters agree with the fit's best values and HESSE covariance (demo data, nearly Gaussian)."""
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    s = H.Session(text, base)
    fit = s.fit()
    chain = s.mcmc(4096, interval=20, nwalkers=512, seed=3)
    fl = np.nonzero(s.floating)[0]
    mean, cov = chain[:, fl].mean(axis=0), np.cov(chain[:, fl].T)
    sig = np.sqrt(np.diag(fit["covariance"]))
    assert np.all(np.abs(mean - fit["values"][fl]) < 0.15 * sig)
    assert np.all(np.abs(np.sqrt(np.diag(cov)) / sig - 1) < 0.15)
    assert np.all(chain[:, -1] >= fit["fval"] - 1e-9)
    assert np.all(chain[:, np.nonzero(~s.floating)[0]] == s.values[~s.floating])


def test_scan_with_more_floating_parameters_than_a_probe_group(built, golden_tree):
    """A probe group of the GPU normal equations holds 2 n + 1 <= 96 columns. Fits with more floating parameters
    are abandoned by the lock-step LM driver and picked up by the Newton driver: the scan still completes."""
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    text += "\ndist-add = -5:4,0:8:2,0\nmodel-config = binning[BAO alpha-perp]={0.98:1.02}*2\nmodel-config = binning[BAO alpha-parallel]={0.98:1.02}*2\n"
    s = H.Session(text, base)
    assert int(s.floating.sum()) - 2 > 47
    res = s.parameter_scan()
    assert len(res["chi2"]) == 4 and np.all(np.isfinite(res["chi2"])) and np.all(res["status"] != 2)
    assert np.all(res["chi2"] >= res["best_chi2"] - 1e-6)


def test_fit_output_files(built, golden_tree, tmp_path):
    """fit.chisq / fit.config / save.pars / save.pcov of src/baofit.cc:661-687 after the fit of the combined sample."""
    from scipy.special import gammaincc
    text = open(os.path.join(golden_tree, "config", "rmu_to_bao.ini")).read()
    base = os.path.join(golden_tree, "config")
    s = H.Session(text, base)
    with pytest.raises(H.HostError, match="bfh_fit"):
        s.write_fit_outputs(str(tmp_path / "early_"))
    r = s.fit()
    prefix = str(tmp_path / "rmu_")
    s.write_fit_outputs(prefix)
    nf = int(s.floating.sum())
    chisq, nbins, npar, prob = open(prefix + "fit.chisq").read().split()
    assert float(chisq) == 2 * r["fval"] and int(nbins) == s.ndata and int(npar) == nf
    assert abs(float(prob) - gammaincc(0.5 * (s.ndata - nf), r["fval"])) < 1e-10
    pars = np.loadtxt(prefix + "save.pars")
    assert np.array_equal(pars[:, 0], np.arange(s.npar)) and np.array_equal(pars[:, 1], r["values"]) and np.array_equal(pars[:, 2], r["errors"])
    pcov = np.loadtxt(prefix + "save.pcov")
    idx = np.nonzero(s.floating)[0]
    assert len(pcov) == nf * (nf + 1) // 2
    for i, j, c in pcov:
        assert c == r["covariance"][list(idx).index(int(i)), list(idx).index(int(j))]
    # the script reproduces the fitted state in a new session
    script = " ".join(l.strip() for l in open(prefix + "fit.config"))
    again = H.Session(text, base, extra_config=script)
    assert np.array_equal(again.values, r["values"]) and np.array_equal(again.floating, s.floating)
    assert np.allclose(again.errors[s.floating], r["errors"][s.floating], rtol=1e-15)
    s.close(), again.close()


def test_toymc_scale_scales_noise_and_chi2(built, golden_tree):
    """toymc-scale (src/baofit.cc:294, CorrelationAnalyzer.cc:349-357): the toy prototype's covariance times the factor. The same
    deviates give sqrt(scale) times the noise, so with no priors the fitted offsets from the truth scale by sqrt(scale) to first
    order and the chi2 values (against the scaled covariance) keep their distribution."""
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    tmp = os.path.join(golden_tree, "data", "rmu_cov2")
    with open(tmp + ".data", "w") as f:
        f.write(open(os.path.join(golden_tree, "demo", "rmu.data")).read())
    with open(tmp + ".cov", "w") as f:
        for i in range(800):
            f.write("%d %d 1e-11\n" % (i, i))
    text = text.replace("load-icov = yes", "").replace("data = ../demo/rmu", "data = ../data/rmu_cov2")
    a = H.Session(text, base).toymc(0, 8, seed=3)
    b = H.Session(text + "\ntoymc-scale = 4\n", base).toymc(0, 8, seed=3)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    assert np.allclose(b["chi2"], a["chi2"], rtol=2e-2)          # same deviates, covariance and noise scaled together
    s = H.Session(text, base)
    fl = s.floating
    da, db = a["fitted"][:, fl] - s.values[fl], b["fitted"][:, fl] - s.values[fl]
    assert np.allclose(db, 2 * da, rtol=0.25, atol=0.1 * np.abs(da).max())  # first order only: the model is not linear
    assert abs(np.median(db[np.abs(da) > 1e-3] / da[np.abs(da) > 1e-3]) - 2) < 0.15
    with pytest.raises(H.HostError, match="varianceScale > 0"):
        H.Session(text + "\ntoymc-scale = 0\n", base)


def test_toymc_truth_follows_the_combined_fit_and_first_toy_is_saved(built, golden_tree, tmp_path):
    """The truth vector of a toy study is the prediction at the parameters of the fit of the combined sample
    (CorrelationAnalyzer.cc:361-369), or at the initial parameters when no fit has been made (no-initial-fit); toymc-save
    writes the first realisation."""
    text, base = ini(golden_tree, "rmu_to_bao.ini")
    tmp = os.path.join(golden_tree, "data", "rmu_cov3")
    rng = np.random.Generator(np.random.PCG64(12))
    rows = [l.split() for l in open(os.path.join(golden_tree, "demo", "rmu.data"))]
    with open(tmp + ".data", "w") as f:   # data moved away from the truth of the recipe so that the fit differs from the start values
        for i, v in rows:
            f.write("%s %.17g\n" % (i, 1.05 * float(v)))
    with open(tmp + ".cov", "w") as f:
        for i in range(800):
            f.write("%d %d 1e-11\n" % (i, i))
    text = text.replace("load-icov = yes", "").replace("data = ../demo/rmu", "data = ../data/rmu_cov3")
    s = H.Session(text, base)
    path = str(tmp_path / "toymcsave.data")
    s.set_toymc_save(path)
    before = s.toymc(0, 6, seed=4)
    saved0 = np.loadtxt(path)
    assert np.array_equal(saved0[:, 0].astype(int), s.kept_indices())
    pred0 = s.predict(s.values[None, :])[:, 0]
    assert np.max(np.abs(saved0[:, 1] - pred0)) < 8 * np.sqrt(1e-11)          # truth at the start values + noise of sigma 1e-5.5
    best = s.fit()
    assert np.max(np.abs(best["values"] - s.values)) > 1e-3
    after = s.toymc(0, 6, seed=4)
    saved1 = np.loadtxt(path)
    pred1 = s.predict(best["values"][None, :])[:, 0]
    assert np.max(np.abs((saved1[:, 1] - pred1) - (saved0[:, 1] - pred0))) < 1e-12   # same deviates on the new truth
    fl = s.floating
    assert np.max(np.abs(after["fitted"][:, fl].mean(axis=0) - best["values"][fl])) < np.max(np.abs(before["fitted"][:, fl].mean(axis=0) - best["values"][fl]))
