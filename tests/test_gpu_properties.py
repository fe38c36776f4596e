"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (the oracle is too slow
there): batch-order invariance, linearity in the BAO amplitude and in the broadband coefficients,
chi-square of noise-free data, agreement of the device-pointer and host-buffer entry points."""
import numpy as np
import pytest

from baofit_b200 import capi, workloads as W

pytestmark = pytest.mark.gpu


def _full_batch(m, B, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})


def test_config4_full_batch_properties(ctx):
    """BOSSDR11QSOLyaF_k + distortion matrix + metals, 37 888 vectors (the bench batch)."""
    m, d = W.qso_kspace()
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    B = 37888
    p = _full_batch(m, B, 5)
    chi2, st = capi.chi2_batch(ctx, gm, gd, p)
    assert np.all(st == 0) and np.all(np.isfinite(chi2)) and np.all(chi2 > 0)
    # batch order does not matter (coherence sort, chunking, panel padding): bit-identical results
    perm = np.random.Generator(np.random.PCG64(1)).permutation(B)
    chi2p, _ = capi.chi2_batch(ctx, gm, gd, p[perm])
    assert np.array_equal(chi2p, chi2[perm])
    # a ragged sub-batch gives the same numbers as the full batch
    sub, _ = capi.chi2_batch(ctx, gm, gd, p[:1000 + 37])
    assert np.max(np.abs(sub - chi2[:1037]) / chi2[:1037]) < 1e-12
    # ... also through the small-batch chi2 GEMM (row tiles split over CTAs, partial sums), down to B = 1
    for nb in (1, 65, 9000, 11000):
        sub, _ = capi.chi2_batch(ctx, gm, gd, p[:nb])
        assert np.max(np.abs(sub - chi2[:nb]) / chi2[:nb]) < 1e-12
    # linear in the BAO amplitude: pred(a=2) - pred(a=1) = pred(a=1) - pred(a=0)
    q = np.repeat(p[:256], 3, axis=0)
    q[:, m.index("BAO amplitude")] = np.tile([0.0, 1.0, 2.0], 256)
    pred, _ = capi.eval_batch(ctx, gm, gd, q)
    d1, d2 = pred[:, 1::3] - pred[:, 0::3], pred[:, 2::3] - pred[:, 1::3]
    assert np.max(np.abs(d1 - d2)) < 1e-12 * np.max(np.abs(pred))
    # linear in a broadband coefficient
    jb = m.index("dist add z0 mu0 r+0 rP1 rT-1")
    q = np.repeat(p[:256], 3, axis=0)
    q[:, jb] = np.tile([-1e-3, 0.0, 1e-3], 256)
    pred, _ = capi.eval_batch(ctx, gm, gd, q)
    d1, d2 = pred[:, 1::3] - pred[:, 0::3], pred[:, 2::3] - pred[:, 1::3]
    assert np.max(np.abs(d1 - d2)) < 1e-11 * np.max(np.abs(d1))
    # chi-square against its own noise-free prediction vanishes
    pred0, _ = capi.eval_batch(ctx, gm, gd, p[:64])
    z, _ = capi.chi2_batch(ctx, gm, gd, p[:64], data_override=np.ascontiguousarray(pred0.T))
    assert np.max(z) < 1e-16 * np.min(chi2)
    gm.close(), gd.close()


def test_config5_full_batch_properties(ctx):
    """BOSSDR11LyaF_fft at 400^3, 1515 bins, 18 944 vectors."""
    m, (r, mu, z, keep) = W.dr11_auto_fft()
    n = len(keep)
    rng = np.random.Generator(np.random.PCG64(2))
    d = W.make_data(r, mu, z, keep, 1e-4 * rng.standard_normal(n), W.synthetic_cov(1e-3 * (50. / r[keep]) ** 2))
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    B = 18944
    p = _full_batch(m, B, 6)
    chi2, st = capi.chi2_batch(ctx, gm, gd, p)
    assert np.all(st == 0) and np.all(np.isfinite(chi2))
    perm = np.random.Generator(np.random.PCG64(3)).permutation(B)
    chi2p, _ = capi.chi2_batch(ctx, gm, gd, p[perm])
    assert np.array_equal(chi2p, chi2[perm])
    # the planes are linear in the BAO-independent bias: xi scales with ((1+beta) b)^2 at fixed beta
    q = np.repeat(p[:64], 2, axis=0)
    q[1::2, m.index("(1+beta)*bias")] *= 2.0
    pred, _ = capi.eval_batch(ctx, gm, gd, q)
    assert np.max(np.abs(pred[:, 1::2] - 4.0 * pred[:, 0::2])) < 1e-12 * np.max(np.abs(pred))
    # vectors that differ only in a key parameter are grouped, not mixed up
    q = p[:512].copy()
    q[::3, m.index("SigmaNL-perp")] = 4.0
    mixed, _ = capi.chi2_batch(ctx, gm, gd, q)
    a, _ = capi.chi2_batch(ctx, gm, gd, q[::3])
    idx = np.setdiff1d(np.arange(512), np.arange(0, 512, 3))
    b, _ = capi.chi2_batch(ctx, gm, gd, q[idx])
    assert np.max(np.abs(mixed[::3] - a) / a) < 1e-12 and np.max(np.abs(mixed[idx] - b) / b) < 1e-12
    gm.close(), gd.close()
