"""Independent check of the toy-MC noise (replaces likely::CovarianceMatrix::sample
at its call site baofit/CorrelationAnalyzer.cc:271): the counter-based generator is restated in numpy integer
arithmetic and the realisations are rebuilt as truth + scale * L g with numpy's own Cholesky factor -- nothing
of the CUDA path is reused -- then the sample covariance of many realisations is held against C itself."""
import numpy as np
import pytest

from baofit_b200 import workloads as W

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
        return x ^ (x >> np.uint64(31))


def counter_normals(seed, toy_ids, n):
    """g[t][j] ~ N(0,1) keyed by (seed, toy id, element): Box-Muller on two 53-bit uniforms drawn from
    splitmix64(seed ^ splitmix64(toy * 0x100000001B3 + j)) and its successor (csrc/kernels_gemm.cu)."""
    with np.errstate(over="ignore"):
        t = np.asarray(toy_ids, dtype=np.uint64)[:, None]
        j = np.arange(n, dtype=np.uint64)[None, :]
        k = splitmix64(np.uint64(seed) ^ splitmix64(t * np.uint64(0x100000001B3) + j))
        k2 = splitmix64(k)
    u1 = ((k >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (k2 >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def test_numpy_generator_is_standard_normal():
    g = counter_normals(1966, np.arange(4000), 50)
    assert abs(g.mean()) < 5 / np.sqrt(g.size) and abs(g.var() - 1) < 5 * np.sqrt(2.0 / g.size)
    assert abs(np.mean(g ** 3)) < 5 * np.sqrt(15.0 / g.size) and abs(np.mean(g ** 4) - 3) < 5 * np.sqrt(96.0 / g.size)
    c = np.corrcoef(g[:, :8].T)
    assert np.max(np.abs(c - np.eye(8))) < 5 / np.sqrt(4000)


@pytest.mark.gpu
def test_gpu_toys_equal_truth_plus_L_g(ctx):
    from baofit_b200 import capi
    m, d = W.qso_rspace()
    n = d.n_data
    cov = np.asarray(d.keep._arrays[5], dtype=np.float64).reshape(n, n).copy()
    L = np.linalg.cholesky(cov)
    rng = np.random.Generator(np.random.PCG64(5))
    truth = rng.standard_normal(n) * 1e-3
    gd = capi.Data(ctx, d)
    for seed, first, ntoy, scale in ((1966, 0, 64, 1.0), (7, 123456789012, 33, 0.25)):
        got = gd.sample_toys(truth, seed, first, ntoy, scale)
        g = counter_normals(seed, first + np.arange(ntoy), n)
        want = truth[None, :] + scale * g @ L.T
        sig = scale * np.sqrt(np.diag(cov))[None, :]
        assert np.max(np.abs(got - want) / sig) < 1e-10
        # the device-resident table holds the same realisations
        t = capi.Toys(ctx, gd, truth, seed, first, ntoy, scale)
        assert np.array_equal(t.download(), got)
        t.close()
    gd.close()


@pytest.mark.gpu
def test_gpu_toy_sample_covariance(ctx):
    """40 000 realisations of the 58-bin demo/multi_0 data set: every element of the sample covariance within
    5 sigma of C (Wishart variance (C_ij^2 + C_ii C_jj)/N), the mean within 5 sigma of the truth."""
    from baofit_b200 import capi
    m, d = W.demo_multi(which=0, lmax=2)
    n = d.n_data
    cov = np.asarray(d.keep._arrays[5], dtype=np.float64).reshape(n, n).copy()
    truth = np.zeros(n)
    gd = capi.Data(ctx, d)
    N = 40000
    x = gd.sample_toys(truth, 2024, 0, N, 1.0)
    gd.close()
    mean = x.mean(axis=0)
    assert np.max(np.abs(mean) / np.sqrt(np.diag(cov) / N)) < 5
    S = (x.T @ x) / N
    sd = np.sqrt((cov ** 2 + np.outer(np.diag(cov), np.diag(cov))) / N)
    assert np.max(np.abs(S - cov) / sd) < 5
