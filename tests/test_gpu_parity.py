"""GPU parity tests proper: CUDA path (through the C ABI) versus the CPU oracle on the same
seeded inputs. Tolerance: 1e-10 relative (north_star), see tests/util.py for the floor."""
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
