"""data-format = comoving-multipole (demo/multi_to_bao.ini): bins are (r, ell, z), predictions are
AbsCorrelationModel::evaluate(r, multipole, z, ...) (baofit/AbsCorrelationModel.cc:43-49,65-79,
baofit/CorrelationFitter.cc:62-69). The oracle is pinned on demo/multi.theory; the CUDA path expands
every bin into Gauss-Legendre points in mu and projects."""
import os

import numpy as np
import pytest

import oracle
from baofit_b200 import host_api as H, workloads as W
from util import RTOL, rel_err, rel_err_bins


def test_oracle_multipole_prediction_is_demo_theory():
    m, d = W.demo_multi(lmax=4)
    pred, st = oracle.predict(oracle.OracleModel(m), oracle.OracleData(d), m.values)
    mt = np.loadtxt(os.path.join(W.GOLDEN, "demo", "multi.theory"))[:, 1]
    assert st == 0 and d.n_data == 87
    assert np.max(np.abs(pred - mt[d.keep._arrays[3]])) < 1e-15


def test_host_multipole_final_cuts(built, golden_tree):
    """lmax defaults to 2 (src/baofit.cc:247): the hexadecapole bins are cut, bit-exact mask."""
    base = os.path.join(golden_tree, "config")
    text = open(os.path.join(base, "multi_to_bao.ini")).read() + "\ndata = ../demo/multi_0\n"
    s = H.Session(text, base)
    m, d = W.demo_multi()
    assert s.ndata == 58 == d.n_data and np.array_equal(s.kept_indices(), d.keep._arrays[3])
    r, ell, z = s.kept_coordinates()
    assert np.array_equal(ell, d.keep._arrays[1]) and np.allclose(r, d.keep._arrays[0], rtol=0, atol=1e-13)
    s4 = H.Session(text + "lmax = 4\n", base)
    assert s4.ndata == 87


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["rspace", "kspace", "fft"])
def test_gpu_multipole_data_against_oracle(ctx, model):
    from baofit_b200 import capi
    m, d = W.demo_multi(lmax=4)
    sweep = {"BAO alpha-parallel": (0.8, 1.25), "BAO alpha-perp": (0.8, 1.25)}
    B = 40
    if model == "kspace":
        pk_fid = W.load_table(os.path.join(W.GOLDEN, "models", "DR9LyaMocksLCDM_matterpower.dat"))
        pk_nw = W.load_table(os.path.join(W.GOLDEN, "models", "DR9LyaMocksLCDMSB_matterpower.dat"))
        m = W.kspace_model(pk_fid, pk_nw, rmin=40, rmax=180, dist_add="-2:0,0:4:2,0")
        m.configure("fix[gamma-beta]=0; fix[gamma-scale]=0; fix[BAO alpha-iso]; fix[1+f]=1.966; fix[SigmaNL-perp]=3.26")
        B = 4
    elif model == "fft":
        m, _ = W.dr11_auto_fft(ngrid=128, spacing=5., nbins=40)
        B = 6
    rng = np.random.Generator(np.random.PCG64(9))
    params = m.random_batch(B, rng, sweep=sweep)
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    pred, st = capi.eval_batch(ctx, gm, gd, params)
    chi2, st2 = capi.chi2_batch(ctx, gm, gd, params)
    ref_chi2, ref_st, ref_pred = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, want_pred=True, nthreads=8)
    assert np.array_equal(st, ref_st) and np.array_equal(st2, ref_st)
    ok = ref_st == 0
    assert ok.sum() >= B // 2
    assert rel_err_bins(pred[:, ok], ref_pred[:, ok]) < RTOL
    assert rel_err(chi2[ok], ref_chi2[ok]) < RTOL
    # the normal-equation path sees the same whitened residuals: gram[f][0][0] = chi2, and toy data vectors work
    gram, gst = capi.gram_batch(ctx, gm, gd, params, 1)
    assert np.array_equal(gst[:, 0], ref_st) and rel_err(gram[ok, 0, 0], ref_chi2[ok]) < RTOL
    toys = capi.Toys(ctx, gd, pred[:, np.argmax(ok)], seed=3, first_toy=0, ntoy=4)
    idx = np.arange(B) % 4
    c_t, _ = capi.chi2_batch_toys(ctx, gm, gd, params, toys, idx)
    ref_t, _ = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, data_override=toys.download()[idx], nthreads=8)
    assert rel_err(c_t[ok], ref_t[ok]) < RTOL
    toys.close()
    gm.close(), gd.close()


@pytest.mark.gpu
def test_gpu_multipole_demo_fit(built, golden_tree):
    """demo/multi_to_bao.ini on demo/multi_0: the fit recovers the generating values within errors."""
    base = os.path.join(golden_tree, "config")
    text = open(os.path.join(base, "multi_to_bao.ini")).read() + "\ndata = ../demo/multi_0\nlmax = 4\n"
    s = H.Session(text, base)
    r = s.fit()
    assert r["status"] == 0
    for name, v in {"beta": 1.4, "BAO alpha-parallel": 1.0, "BAO alpha-perp": 1.0}.items():
        i = s.names.index(name)
        assert abs(r["values"][i] - v) < 4 * r["errors"][i], (name, r["values"][i], r["errors"][i])


def test_host_platelist_combination_is_inverse_covariance_weighted(built, golden_tree):
    """plateroot + platelist (src/baofit.cc:531-583): congruent observations are combined like
    likely::BinnedDataResampler::combined does, C^-1 = sum C_k^-1, C^-1 d = sum C_k^-1 d_k; observations
    with different bins are refused."""
    base = os.path.join(golden_tree, "config")
    m, d0 = W.demo_multi(lmax=4)
    keep, dv, cov = d0.keep._arrays[3], d0.keep._arrays[4], d0.keep._arrays[5]
    rng = np.random.Generator(np.random.PCG64(12))
    obs = []
    for k in range(3):
        ck = cov * (1.0 + 0.5 * k)
        dk = dv + np.linalg.cholesky(ck) @ rng.standard_normal(len(dv))
        W.write_dataset_files(os.path.join(golden_tree, "data", "obs_%d" % k), 87, keep, dk, ck)
        obs.append((dk, ck))
    with open(os.path.join(golden_tree, "data", "obs.list"), "w") as f:
        f.write("obs_0\nobs_1\nobs_2\n")
    text = open(os.path.join(base, "multi_to_bao.ini")).read() + "\nplateroot = ../data/\nplatelist = obs.list\nlmax = 4\n"
    s = H.Session(text, base)
    assert s.ndata == 87
    icov = [np.linalg.inv(c) for _, c in obs]
    csum = np.linalg.inv(sum(icov))
    dcomb = csum @ sum(ic @ dk for ic, (dk, _) in zip(icov, obs))
    assert np.max(np.abs(s.data_vector() - dcomb)) < 1e-11 * np.max(np.abs(dcomb))
    assert np.max(np.abs(s.covariance() - csum)) < 1e-10 * np.max(np.abs(csum))
    assert H.Session(text + "max-plates = 1\n", base).data_vector() == pytest.approx(obs[0][0], rel=1e-12)
    # demo/multi_0, multi_1, multi_2 carry different multipole sets: not congruent
    with open(os.path.join(golden_tree, "data", "bad.list"), "w") as f:
        f.write("../demo/multi_0\n../demo/multi_1\n")
    with pytest.raises(H.HostError):
        H.Session(text.replace("obs.list", "bad.list"), base)
