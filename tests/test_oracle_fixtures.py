"""Parity at scale against committed oracle answers (tests/golden/oracle_<case>.npz, made by
tools/gen_oracle_fixtures.py from tests/parity_cases.py).

GPU tests: the CUDA path through the C ABI versus the stored oracle predictions / chi-square of every vector
of a file -- 256 vectors for each of BASELINE configs 2-5 at full size, 8 for each optional branch -- at the
north_star tolerance of 1e-10 (tests/util.py for the per-bin floor).
CPU tests: the files are pinned to the oracle as it is now (two vectors of each file re-run), so a change to
oracle/oracle.c cannot silently leave stale answers behind.
"""
import os

import numpy as np
import pytest

import oracle
import parity_cases as PC
from util import RTOL, rel_err, rel_err_bins


def _load(name):
    path = PC.fixture_path(name)
    if not os.path.exists(path):
        pytest.fail("missing fixture %s: run tools/gen_oracle_fixtures.py --cases %s" % (path, name))
    f = np.load(path)
    m, d, sweep, B, _ = PC.build(name, xi0=f["xi0"] if "xi0" in f.files else None)
    assert list(f["names"]) == m.names and int(f["n_data"]) == d.n_data
    assert f["params"].shape == (B, m.desc.npar) and f["pred"].shape == (d.n_data, B)
    return f, m, d


@pytest.mark.parametrize("name", PC.ALL)
def test_fixture_is_what_the_oracle_says_now(built, name):
    f, m, d = _load(name)
    B = f["params"].shape[0]
    pick = [0, B - 1]  # one vector of each half (shared-key and floating-key halves of the option cases)
    params = np.ascontiguousarray(f["params"][pick])
    om, od = oracle.OracleModel(m), oracle.OracleData(d)
    chi2, status, pred = oracle.chi2_batch(om, od, params, want_pred=True, nthreads=os.cpu_count() or 1)
    assert np.array_equal(status, f["status"][pick])
    ok = status == 0
    assert np.allclose(pred[:, ok], f["pred"][:, pick][:, ok], rtol=1e-13, atol=1e-13 * np.abs(f["pred"]).max())
    assert np.allclose(chi2[ok], f["chi2"][pick][ok], rtol=1e-12)
    # the seeded parameter recipe still produces the stored vectors
    assert np.array_equal(PC.make_params(name, m, PC.build(name, xi0=f["xi0"] if "xi0" in f.files else None)[2], B), f["params"])


def _gpu_compare(ctx, m, d, params, ref_pred, ref_chi2, ref_status):
    from baofit_b200 import capi
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    pred, status = capi.eval_batch(ctx, gm, gd, params)
    chi2, status2 = capi.chi2_batch(ctx, gm, gd, params)
    gm.close(), gd.close()
    assert np.array_equal(status, ref_status) and np.array_equal(status2, ref_status)
    ok = ref_status == 0
    assert ok.sum() >= max(1, len(ok) // 2)
    assert np.all(np.isnan(chi2[~ok]))
    return rel_err_bins(pred[:, ok], ref_pred[:, ok]), rel_err(chi2[ok], ref_chi2[ok])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(PC.BIG))
def test_gpu_baseline_configs_at_full_size_256_vectors(ctx, name):
    f, m, d = _load(name)
    e_pred, e_chi2 = _gpu_compare(ctx, m, d, f["params"], f["pred"], f["chi2"], f["status"])
    assert e_pred < RTOL and e_chi2 < RTOL, (name, e_pred, e_chi2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PC.OPTIONS + PC.METALS)
def test_gpu_optional_branches(ctx, name):
    """All vectors in one call (mixed keys: per-vector transforms), then the two halves on their own: the first
    half shares its key parameters and takes the shared-transform path where the model allows it."""
    f, m, d = _load(name)
    B = f["params"].shape[0]
    for sel in (slice(0, B), slice(0, B // 2), slice(B // 2, B)):
        e_pred, e_chi2 = _gpu_compare(ctx, m, d, np.ascontiguousarray(f["params"][sel]), f["pred"][:, sel], f["chi2"][sel],
                                      f["status"][sel])
        assert e_pred < RTOL and e_chi2 < RTOL, (name, sel, e_pred, e_chi2)
