"""GPU parity tests proper: CUDA path (through the C ABI) versus the CPU oracle on the same
seeded inputs. Tolerance: 1e-10 relative (north_star), see tests/util.py for the floor."""
import os
import numpy as np
import pytest

import oracle
from baofit_b200 import capi, workloads as W
from util import RTOL, rel_err, rel_err_bins

pytestmark = pytest.mark.gpu


def _compare(ctx, m, d, B, seed, sweep=None, scale=0.5):
    rng = np.random.Generator(np.random.PCG64(seed))
    params = m.random_batch(B, rng, scale=scale, sweep=sweep)
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    pred, status = capi.eval_batch(ctx, gm, gd, params)
    chi2, status2 = capi.chi2_batch(ctx, gm, gd, params)
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    ref_chi2, ref_status, ref_pred = oracle.chi2_batch(om, od, params, want_pred=True, nthreads=8)
    assert np.array_equal(status, ref_status)
    assert np.array_equal(status2, ref_status)
    ok = ref_status == 0
    assert ok.any()
    e_pred = rel_err_bins(pred[:, ok], ref_pred[:, ok])
    e_chi2 = rel_err(chi2[ok], ref_chi2[ok])
    assert np.all(np.isnan(chi2[~ok]))
    gm.close(), gd.close()
    return e_pred, e_chi2


def test_rspace_demo_rmu(ctx):
    m, d = W.demo_rmu()
    e_pred, e_chi2 = _compare(ctx, m, d, 200, 11, sweep={"BAO alpha-parallel": (0.7, 1.4), "BAO alpha-perp": (0.6, 1.5)})
    assert e_pred < RTOL and e_chi2 < RTOL, (e_pred, e_chi2)


def test_rspace_demo_rmu_truth_chi2(ctx):
    """chi2 known-answer derived from the reference fixtures: 761.405865008929"""
    m, d = W.demo_rmu()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, _ = capi.chi2_batch(ctx, gm, gd, m.values[None, :])
    assert abs(chi2[0] - 761.405865008929) < 1e-9 * 761.4


def test_rspace_qso_cross(ctx):
    m, d = W.qso_rspace()
    e_pred, e_chi2 = _compare(ctx, m, d, 333, 12, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5),
                                                          "delta-v": (-300, 300)})
    assert e_pred < RTOL and e_chi2 < RTOL, (e_pred, e_chi2)


def test_kspace_transform_tables(ctx):
    """P(k) -> xi_ell(r) tables of the peak / no-wiggles components: Filon weights + device
    projection + DMMA GEMM versus the oracle's brute-force quadrature of the same definition."""
    m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    rng = np.random.Generator(np.random.PCG64(5))
    params = m.random_batch(3, rng)
    gm = capi.Model(ctx, m)
    r, tabs = gm.transform(params, 2.35717)
    om = oracle.OracleModel(m)
    for b in range(params.shape[0]):
        ro, ref = om.transform(params[b], 2.35717, jstep=31)
        sel = ~np.isnan(ref[0, 0])
        assert np.allclose(r, ro, rtol=0, atol=1e-12)
        for c in range(2):
            for l in range(3):
                got, want = tabs[c, l, sel, b], ref[c, l, sel]
                err = np.max(np.abs(got - want)) / np.max(np.abs(want))
                assert err < RTOL, (b, c, l, err)
    gm.close()


def test_kspace_qso_cross_plain(ctx):
    m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    e_pred, e_chi2 = _compare(ctx, m, d, 4, 21, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.4)})
    assert e_pred < RTOL and e_chi2 < RTOL, (e_pred, e_chi2)


def test_kspace_qso_cross_distortion_matrix_metals(ctx):
    """BASELINE config 4: cross-correlation k-space model + distortion matrix + bicubic metals."""
    m, d = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(22))
    params = m.random_batch(4, rng, sweep={"BAO alpha-parallel": (0.85, 1.2), "BAO alpha-perp": (0.7, 1.3),
                                           "delta-v": (-200, 200)})
    for j in range(m.desc.npar):
        if m.names[j].startswith("bias Si"):
            params[:, j] *= rng.uniform(0.5, 2.0, size=params.shape[0])
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    xiu, st = capi.eval_undistorted_batch(ctx, gm, gd, params)
    for b in range(params.shape[0]):
        ref, _ = oracle.undistorted(om, od, params[b], d.desc.n_grid)
        assert rel_err_bins(xiu[:, b], ref) < RTOL
    pred, status = capi.eval_batch(ctx, gm, gd, params)
    chi2, _ = capi.chi2_batch(ctx, gm, gd, params)
    ref_chi2, ref_status, ref_pred = oracle.chi2_batch(om, od, params, want_pred=True, nthreads=8)
    assert np.array_equal(status, ref_status)
    assert rel_err_bins(pred, ref_pred) < RTOL
    assert rel_err(chi2, ref_chi2) < RTOL
    gm.close(), gd.close()


def test_kspace_dilation_limits_flag_not_abort(ctx):
    """The reference throws on dilation limits (BaoKSpaceCorrelationModel.cc:420-427); a batch
    flags the offending vectors instead and the others stay valid."""
    m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    params = np.tile(m.values, (3, 1))
    params[1, m.index("BAO alpha-parallel")] = 1.7
    params[1, m.index("BAO alpha-perp")] = 1.7
    params[2, m.index("BAO alpha-parallel")] = 0.3
    params[2, m.index("BAO alpha-perp")] = 0.3
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, status = capi.chi2_batch(ctx, gm, gd, params)
    assert status[0] == 0 and np.isfinite(chi2[0])
    assert status[1] == 2 and np.isnan(chi2[1])
    assert status[2] == 1 and np.isnan(chi2[2])
    pred, st = capi.eval_batch(ctx, gm, gd, params)
    assert np.all(np.isnan(pred[:, 1])) and np.all(np.isnan(pred[:, 2])) and np.all(np.isfinite(pred[:, 0]))
    gm.close(), gd.close()


def test_kspace_per_vector_transform_path(ctx):
    """Vectors with different non-linear broadening cannot share basis tables: this exercises the
    per-vector projection + Hankel GEMM + spline path end to end."""
    m, d = W.qso_kspace(with_dist_matrix=True, with_metals=False)
    m.configure("release[SigmaNL-perp]; release[1+f]")
    e = []
    rng = np.random.Generator(np.random.PCG64(31))
    params = m.random_batch(3, rng, sweep={"BAO alpha-parallel": (0.9, 1.1), "BAO alpha-perp": (0.9, 1.1)})
    assert len(set(params[:, m.index("SigmaNL-perp")])) == 3
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    pred, status = capi.eval_batch(ctx, gm, gd, params)
    chi2, _ = capi.chi2_batch(ctx, gm, gd, params)
    ref_chi2, ref_status, ref_pred = oracle.chi2_batch(om, od, params, want_pred=True, nthreads=8)
    assert rel_err_bins(pred, ref_pred) < RTOL
    assert rel_err(chi2, ref_chi2) < RTOL
    gm.close(), gd.close()


def test_kspace_shared_and_per_vector_paths_agree(ctx):
    """Same vectors evaluated alone (shared basis tables) and inside a batch whose first vector has
    a different SigmaNL (per-vector transforms) must give the same chi2."""
    m, d = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(32))
    params = m.random_batch(65, rng, sweep={"BAO alpha-parallel": (0.85, 1.2), "BAO alpha-perp": (0.7, 1.3)})
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    ctx.set_profiling(True)
    shared, _ = capi.chi2_batch(ctx, gm, gd, params[1:])
    assert any(k.startswith(("kspace_fast_eval_kernel", "fused_model_chi2_kernel")) for k in ctx.get_profile())
    mixed = params.copy()
    mixed[0, m.index("SigmaNL-perp")] *= 1.1
    ctx.set_profiling(True)
    generic, _ = capi.chi2_batch(ctx, gm, gd, mixed)
    assert any(k.startswith("kspace_project_kernel") for k in ctx.get_profile())
    ctx.set_profiling(False)
    assert rel_err(shared, generic[1:]) < RTOL
    gm.close(), gd.close()


def test_kspace_basis_table_cache_follows_key_parameters(ctx):
    """The shared basis tables are kept across calls while SigmaNL etc. are unchanged and must be
    rebuilt when they change (the reference re-transforms only then, BaoKSpaceCorrelationModel.cc:333-380)."""
    m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    base = m.values.copy()
    for snl in (3.26, 3.26, 4.5, 3.26):
        p = np.tile(base, (3, 1))
        p[:, m.index("SigmaNL-perp")] = snl
        p[:, m.index("beta")] += [0.0, 0.05, -0.05]
        chi2, _ = capi.chi2_batch(ctx, gm, gd, p)
        ref, _ = oracle.chi2_batch(om, od, p, nthreads=8)
        assert rel_err(chi2, ref) < RTOL, snl
    gm.close(), gd.close()


def test_coherence_sort_does_not_change_results(ctx):
    """Batches of >= 2048 vectors are internally sorted on their dilation parameters; every column is
    independent, so chi2 must be identical to the unsorted evaluation of the same vectors."""
    m, d = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(41))
    params = m.random_batch(4096, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5),
                                              "delta-v": (-300, 300)})
    params[7, m.index("BAO alpha-perp")] = 1.9  # one vector beyond the dilation limit
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    big, st_big = capi.chi2_batch(ctx, gm, gd, params)
    parts = [capi.chi2_batch(ctx, gm, gd, params[k:k + 1024]) for k in range(0, 4096, 1024)]
    small = np.concatenate([p[0] for p in parts])
    st_small = np.concatenate([p[1] for p in parts])
    assert np.array_equal(st_big, st_small) and st_big[7] == 2 and st_big.sum() == 2
    ok = st_big == 0
    assert np.array_equal(big[ok], small[ok]) and np.isnan(big[7])
    gm.close(), gd.close()


def test_rspace_large_batch_sorted_path(ctx):
    m, d = W.qso_rspace()
    e_pred, e_chi2 = _compare(ctx, m, d, 2500, 43, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})
    assert e_pred < RTOL and e_chi2 < RTOL, (e_pred, e_chi2)


def test_toy_table_on_device_matches_host_override_path(ctx):
    """bf_toys_create / bf_chi2_batch_toys (realisations stay on the GPU, points carry a toy index)
    against bf_sample_toys + bf_chi2_batch with explicit per-point data vectors, and the oracle."""
    m, d = W.qso_rspace()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    truth, _ = capi.eval_batch(ctx, gm, gd, m.values[None, :])
    truth = np.ascontiguousarray(truth[:, 0])
    toys = capi.Toys(ctx, gd, truth, seed=42, first_toy=100, ntoy=9)
    table = toys.download()
    assert np.array_equal(table, gd.sample_toys(truth, 42, 100, 9))
    rng = np.random.Generator(np.random.PCG64(1))
    params = m.random_batch(50, rng)
    idx = rng.integers(0, 9, size=50)
    a, sa = capi.chi2_batch_toys(ctx, gm, gd, params, toys, idx)
    b, sb = capi.chi2_batch(ctx, gm, gd, params, data_override=table[idx])
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    ref, _ = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, data_override=table[idx], nthreads=8)
    assert rel_err(a, ref) < RTOL
    toys.close(), gm.close(), gd.close()


@pytest.mark.parametrize("case", ["rspace", "kspace_dm", "fft"])
def test_gram_batch_against_oracle_residuals(ctx, case):
    """bf_gram_batch: Gram matrices of the whitened residuals y = L^-1 (pred - d) of probe groups,
    against numpy on the oracle's predictions."""
    import scipy.linalg
    if case == "rspace":
        m, d = W.qso_rspace()
        nfit, G = 7, 19
    elif case == "kspace_dm":
        m, d = W.qso_kspace()
        nfit, G = 2, 3
    else:
        import test_fft_model as T
        m, d, _ = T._gpu_case(ngrid=128, spacing=5., nbins=40)
        nfit, G = 3, 6
    rng = np.random.Generator(np.random.PCG64(31))
    centres = m.random_batch(nfit, rng)
    params = np.repeat(centres, G, axis=0)
    fl = np.nonzero(m.floating)[0]
    for f in range(nfit):
        for g in range(1, G):
            j = fl[(g - 1) % len(fl)]
            params[f * G + g, j] += 0.01 * m.errors[j]
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    gram, st = capi.gram_batch(ctx, gm, gd, params, G)
    chi2, _ = capi.chi2_batch(ctx, gm, gd, params)
    _, ref_st, pred = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), params, want_pred=True, nthreads=8)
    assert np.all(st == 0) and np.all(ref_st == 0)
    cov = d.keep._arrays[5]
    L = np.linalg.cholesky(cov)
    y = scipy.linalg.solve_triangular(L, pred - d.keep._arrays[4][:, None], lower=True)
    for f in range(nfit):
        yf = y[:, f * G:(f + 1) * G]
        D = np.column_stack([yf[:, 0]] + [yf[:, g] - yf[:, 0] for g in range(1, G)])
        ref = D.T @ D
        scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
        assert np.max(np.abs(gram[f] - ref) / scale) < 1e-7  # differences of nearly equal residual vectors
        assert abs(gram[f, 0, 0] - chi2[f * G]) < 1e-10 * chi2[f * G]
    gm.close(), gd.close()


def test_gram_probes_equal_explicit_probe_groups(ctx):
    """bf_gram_probes (probe columns generated on the GPU from centre + steps) returns bit-identical
    Gram matrices to bf_gram_batch fed with the same columns built on the host."""
    m, d = W.qso_rspace()
    rng = np.random.Generator(np.random.PCG64(8))
    nfit = 40
    centres = m.random_batch(nfit, rng)
    steps = np.zeros_like(centres)
    fl = np.nonzero(m.floating)[0]
    steps[:, fl] = 0.01 * m.errors[fl] * rng.uniform(0.5, 1.5, size=(nfit, len(fl)))
    G = 2 * len(fl) + 1
    cols = np.repeat(centres, G, axis=0)
    for f in range(nfit):
        for k, j in enumerate(fl):
            cols[f * G + 1 + 2 * k, j] += steps[f, j]
            cols[f * G + 2 + 2 * k, j] -= steps[f, j]
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    a, sa = capi.gram_probes(ctx, gm, gd, centres, steps)
    b, sb = capi.gram_batch(ctx, gm, gd, cols, G)
    assert np.array_equal(a, b) and np.array_equal(sa, sb.max(axis=1))
    toys = capi.Toys(ctx, gd, np.zeros(gd.n), seed=2, first_toy=0, ntoy=5)
    idx = rng.integers(0, 5, size=nfit)
    a, _ = capi.gram_probes(ctx, gm, gd, centres, steps, toys, idx)
    b, _ = capi.gram_batch(ctx, gm, gd, cols, G, toys, idx)
    assert np.array_equal(a, b)
    # bf_normal_probes: the same groups reduced on the device to [y0 a]^T [y0 a] and y0 . s. The Gram matrix holds
    # d+ = y(x+h) - y0 and d- = y(x-h) - y0, so a = d+ - d- and s = d+ + d-.
    n = len(fl)
    for tt, ii in ((None, None), (toys, idx)):
        g, _ = capi.gram_probes(ctx, gm, gd, centres, steps, tt, ii)
        # (all columns evaluated: the synthesised columns of linear parameters are exact where these differences carry
        # rounding noise, and have their own test, test_linear_probe_columns_are_synthesised_exactly)
        os.environ["BAOFIT_NO_LINEAR_PROBES"] = "1"
        try:
            M, t, st = capi.normal_probes(ctx, gm, gd, centres, steps, tt, ii)
        finally:
            os.environ.pop("BAOFIT_NO_LINEAR_PROBES", None)
        ip, im = 1 + 2 * np.arange(n), 2 + 2 * np.arange(n)
        assert np.all(st == 0)
        assert np.allclose(M[:, 0, 0], g[:, 0, 0], rtol=1e-13)
        ref_a0 = g[:, ip, 0] - g[:, im, 0]
        scale = np.sqrt(g[:, 0, 0])[:, None] * np.sqrt(np.abs(g[:, ip, ip] + g[:, im, im]))  # |y0| |d|: cancellation scale
        assert np.max(np.abs(M[:, 1:, 0] - ref_a0) / scale) < 1e-12 and np.array_equal(M[:, 1:, 0], M[:, 0, 1:])
        assert np.max(np.abs(t - (g[:, ip, 0] + g[:, im, 0])) / scale) < 1e-12
        ref_aa = g[:, ip][:, :, ip] - g[:, ip][:, :, im] - g[:, im][:, :, ip] + g[:, im][:, :, im]
        dn = np.sqrt(np.abs(g[:, ip, ip] + g[:, im, im]))
        assert np.max(np.abs(M[:, 1:, 1:] - ref_aa) / (dn[:, :, None] * dn[:, None, :])) < 1e-12
    toys.close(), gm.close(), gd.close()


def test_fused_data_tail_equals_unfused_path(ctx):
    """k-space models without a distortion matrix run chi2 / whitened residuals as cosmology pass + broadband
    coefficient rows + one GEMM against [W | W.basis | -W.d] (per redshift class for the three z bins of DR9).
    The prediction output still takes the unfused two-pass path: both must agree, also with a data override and
    through the normal-equation probes."""
    import scipy.linalg as sl
    rng = np.random.Generator(np.random.PCG64(41))
    m9, d9 = W.dr9_kspace()
    m11, (r, mu, z, keep) = W.dr11_auto_kspace()
    n11 = len(keep)
    xi0 = 1e-3 * (50.0 / r[keep]) ** 2
    d11 = W.make_data(r, mu, z, keep, xi0 * (1 + 0.1 * rng.standard_normal(n11)), W.synthetic_cov(xi0))
    for m, d in ((m9, d9), (m11, d11), W.qso_kspace()):  # the last one: the fused tail of the distortion-matrix path
        gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
        B = 200
        p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.9, 1.1), "BAO alpha-perp": (0.9, 1.1)})
        jb = [j for j, nm in enumerate(m.names) if nm.startswith("dist add")]
        p[:, jb] += 1e-4 * rng.standard_normal((B, len(jb)))  # broadband coefficients away from zero
        dv, cov = d.keep._arrays[4], d.keep._arrays[5]
        L = np.linalg.cholesky(cov)
        pred, st = capi.eval_batch(ctx, gm, gd, p)
        assert np.all(st == 0)
        chi2, _ = capi.chi2_batch(ctx, gm, gd, p)
        ref = np.sum(sl.solve_triangular(L, pred - dv[:, None], lower=True) ** 2, axis=0)
        assert np.max(np.abs(chi2 - ref) / ref) < 1e-10
        ov = dv[None, :] + (L @ rng.standard_normal((gd.n, B))).T
        chi2o, _ = capi.chi2_batch(ctx, gm, gd, p, data_override=np.ascontiguousarray(ov))
        refo = np.sum(sl.solve_triangular(L, pred - ov.T, lower=True) ** 2, axis=0)
        assert np.max(np.abs(chi2o - refo) / refo) < 1e-10
        steps = np.zeros_like(p[:8])
        fl = np.nonzero(m.floating)[0]
        steps[:, fl] = 0.01 * m.errors[fl]
        M, t, stn = capi.normal_probes(ctx, gm, gd, p[:8], steps)
        assert np.all(stn == 0) and np.max(np.abs(M[:, 0, 0] - chi2[:8]) / chi2[:8]) < 1e-10
        gm.close(), gd.close()


def test_linear_probe_columns_are_synthesised_exactly(ctx):
    """bf_normal_probes on models whose additive broadband enters the whitened residual linearly with a bin-static
    basis (k-space without distortion matrix, no velocity shift): the difference columns of those parameters come
    from W . basis instead of two evaluations each. Same normal equations as the all-evaluated path
    (BAOFIT_NO_LINEAR_PROBES), for one redshift class (DR11 auto) and three (DR9), with and without toys; and far
    fewer evaluations."""
    rng = np.random.Generator(np.random.PCG64(43))
    m9, d9 = W.dr9_kspace()
    m11, (r, mu, z, keep) = W.dr11_auto_kspace()
    n11 = len(keep)
    xi0 = 1e-3 * (50.0 / r[keep]) ** 2
    d11 = W.make_data(r, mu, z, keep, xi0 * (1 + 0.1 * rng.standard_normal(n11)), W.synthetic_cov(xi0))
    # ... and the r-space QSO cross-correlation model, whose velocity shift makes every coefficient (z, p, t) a combination
    # of the static columns (z, m <= p, t) with weights C(p, m) v^(p - m) of the fit's own delta-v
    for m, d in ((m11, d11), (m9, d9), W.qso_rspace()):
        gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
        nfit = 9
        p = m.random_batch(nfit, rng, sweep={"BAO alpha-parallel": (0.9, 1.1), "BAO alpha-perp": (0.9, 1.1)})
        jb = [j for j, nm in enumerate(m.names) if nm.startswith("dist add")]
        p[:, jb] += 1e-4 * rng.standard_normal((nfit, len(jb)))
        if "delta-v" in m.names:
            p[:, m.index("delta-v")] = rng.uniform(-400, 400, nfit)
        steps = np.zeros_like(p)
        fl = [j for j in np.nonzero(m.floating)[0] if not m.names[j].startswith("BAO alpha")]
        assert len(set(fl) & set(jb)) >= 3
        steps[:, fl] = 0.01 * m.errors[fl] * (1 + 0.3 * rng.random((nfit, len(fl))))
        truth = capi.eval_batch(ctx, gm, gd, p[:1])[0][:, 0]
        toys = capi.Toys(ctx, gd, truth, 5, 0, 4)
        ti = rng.integers(0, 4, nfit).astype(np.int32)
        for tt, ii in ((None, None), (toys, ti)):
            out = {}
            for mode in ("lin", "all"):
                if mode == "all":
                    os.environ["BAOFIT_NO_LINEAR_PROBES"] = "1"
                try:
                    ctx.set_profiling(True)
                    out[mode] = capi.normal_probes(ctx, gm, gd, p, steps, tt, ii)
                    prof = ctx.get_profile()
                    ctx.set_profiling(False)
                finally:
                    os.environ.pop("BAOFIT_NO_LINEAR_PROBES", None)
                assert ("normal_probes_lin_kernel" in prof) == (mode == "lin"), prof.keys()
            (Ml, tl, sl_), (Ma, ta, sa) = out["lin"], out["all"]
            assert np.array_equal(sl_, sa) and np.all(sa == 0)
            dn = np.sqrt(np.abs(np.einsum("fii->fi", Ma)))
            assert np.max(np.abs(Ml - Ma) / (dn[:, :, None] * dn[:, None, :])) < 1e-9
            # y0 . s_i: exactly zero for a linear parameter, rounding noise of two evaluations otherwise
            scale = dn[:, :1] * dn[:, :1]
            assert np.max(np.abs(tl - ta) / scale) < 1e-9
        toys.close(), gm.close(), gd.close()


def test_rspace_fast_and_generic_kernels_against_oracle(ctx):
    """The r-space model runs through the basis-table fast pass when both tabulations share one uniform grid and
    through the generic per-vector kernel otherwise; both against the oracle, with the radiation term switched on
    (Rad strength > 0 on half of the vectors), gamma-beta != 0 on three redshift classes, and a dist-mul broadband
    (which keeps chi2 off the fused tail)."""
    root = os.path.join(W.GOLDEN, "models")
    fid, nw = W.load_multipoles(root, "DR9LyaMocksNLeffPk"), W.load_multipoles(root, "DR9LyaMocksSB")
    keepk = np.ones(len(fid[0]), bool)
    keepk[[17, 50, 51, 120]] = False  # non-uniform knots: generic kernel
    thin = lambda t: tuple(np.ascontiguousarray(a[keepk]) for a in t)
    rng = np.random.Generator(np.random.PCG64(52))
    nb = 120
    r = rng.uniform(20., 190., nb)
    mu = rng.uniform(-1., 1., nb)
    z = rng.choice([2.0, 2.4, 2.9], nb)
    keep = np.arange(nb, dtype=np.int32)
    cov = np.diag(rng.uniform(0.5, 2.0, nb) * 1e-8)
    for tabs, kw in (((fid, nw), dict(dist_add="-2:0,0:4:2,0")), ((thin(fid), thin(nw)), dict(dist_add="-2:0,0:4:2,0")),
                     ((fid, nw), dict(dist_add="0:1,0,0", dist_mul="0:1,0,0", cross=True))):
        m = W.rspace_model(tabs[0], tabs[1], **kw)
        d = W.make_data(r, mu, z, keep, 1e-4 * rng.standard_normal(nb), cov)
        B = 64
        p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.2), "BAO alpha-perp": (0.8, 1.2)})
        p[:, m.index("gamma-beta")] = rng.uniform(-0.5, 0.5, B)
        p[::2, m.index("Rad strength")] = rng.uniform(0.01, 0.1, B // 2)
        p[:, m.index("Rad anisotropy")] = rng.uniform(0.0, 0.5, B)
        for j, nm in enumerate(m.names):
            if nm.startswith("dist "):
                p[:, j] = 1e-3 * rng.standard_normal(B)
        gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
        om, od = oracle.OracleModel(m), oracle.OracleData(d)
        pred, st = capi.eval_batch(ctx, gm, gd, p)
        chi2, _ = capi.chi2_batch(ctx, gm, gd, p)
        ref_chi2, ref_st, ref_pred = oracle.chi2_batch(om, od, p, nthreads=8, want_pred=True)
        assert np.array_equal(st, ref_st)
        assert np.max(np.abs(pred - ref_pred)) < RTOL * np.max(np.abs(ref_pred))
        assert rel_err(chi2, ref_chi2) < RTOL
        gm.close(), gd.close()


def test_fused_producer_kernel_equals_unfused_chain(built):
    """BASELINE config 4, 20 000 vectors: fused_model_chi2_kernel (model evaluation as the producer of the chi2 GEMM's
    right-hand operand) against the separate cosmology / metal / GEMM kernels (BAOFIT_UNFUSED=1 in a child process:
    the switch is read once per process), and both against the oracle on a subset."""
    import json, subprocess, sys, tempfile
    code = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
from baofit_b200 import capi, workloads as W
m, d = W.qso_kspace()
rng = np.random.Generator(np.random.PCG64(91))
p = m.random_batch(20000, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5), "delta-v": (-300, 300)})
p[:40, m.index("delta-v")] = np.linspace(-3000, 3000, 40)   # one panel whose velocity shifts do not fit a staged window
ctx = capi.Context(0)
gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
ctx.set_profiling(True)
chi2, st = capi.chi2_batch(ctx, gm, gd, p)
names = sorted(ctx.get_profile())
# ragged batches: a last panel with fewer than 32 vectors, fewer panels than SMs, one panel more than SMs
extra = [capi.chi2_batch(ctx, gm, gd, p[:n]) for n in (1031, 4737, 4801)]
np.save(sys.argv[1], np.vstack([np.concatenate([chi2] + [e[0] for e in extra]), np.concatenate([st] + [e[1] for e in extra])]))
print(json.dumps(names))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ("fused", "unfused"):
            env = dict(os.environ)
            env.pop("BAOFIT_UNFUSED", None)
            if mode == "unfused":
                env["BAOFIT_UNFUSED"] = "1"
            path = os.path.join(tmp, mode + ".npy")
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            out[mode] = (np.load(path), json.loads(r.stdout.strip().splitlines()[-1]))
    assert "fused_model_chi2_kernel" in out["fused"][1] and "fused_model_chi2_kernel" not in out["unfused"][1]
    (cf, sf), (cu, su) = out["fused"][0], out["unfused"][0]
    assert cf.shape == (20000 + 1031 + 4737 + 4801,)
    assert np.array_equal(sf, su)
    ok = sf == 0
    assert ok.sum() > 15000 and np.all(np.isnan(cf[~ok]))
    assert rel_err(cf[ok], cu[ok]) < 1e-11
    # a vector's chi2 does not depend on the batch it is evaluated in
    at = 20000
    for n in (1031, 4737, 4801):
        sub = cf[at:at + n]
        good = ok[:n]
        assert np.array_equal(sub[good], cf[:n][good]), n
        at += n
    m, d = W.qso_kspace()
    rng = np.random.Generator(np.random.PCG64(91))
    p = m.random_batch(20000, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5), "delta-v": (-300, 300)})
    p[:40, m.index("delta-v")] = np.linspace(-3000, 3000, 40)
    pick = np.r_[0:6, 19990:19996]
    ref, rst = oracle.chi2_batch(oracle.OracleModel(m), oracle.OracleData(d), p[pick], nthreads=os.cpu_count() or 8)
    assert np.array_equal(rst, sf[pick].astype(rst.dtype))
    good = rst == 0
    assert rel_err(cf[pick][good], ref[good]) < RTOL
