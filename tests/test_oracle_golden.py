"""Pins the CPU oracle on every fixture the reference ships for the hot path.
These run without a GPU."""
import os

import numpy as np
import pytest

import oracle
from baofit_b200 import workloads as W

G = W.GOLDEN


@pytest.fixture(scope="module")
def rmu():
    m, d = W.demo_rmu()
    return m, d, oracle.OracleModel(m), oracle.OracleData(d)


def test_rmu_theory_per_bin(rmu):
    """demo/rmu.theory: 800 per-bin known answers of BaoCorrelationModel (demo/rmu_to_bao.ini)."""
    m, d, om, od = rmu
    theory = np.loadtxt(os.path.join(G, "demo", "rmu.theory"))[:, 1]
    pred, st = oracle.predict(om, od, m.values)
    assert st == 0 and len(pred) == 800
    assert np.max(np.abs(pred - theory)) / np.max(np.abs(theory)) < 1e-13


def test_rmu_chi2_known_answer(rmu):
    """chi2(demo/rmu.data | demo/rmu.theory, demo/rmu.icov) = 761.405865008929."""
    m, d, om, od = rmu
    theory = np.loadtxt(os.path.join(G, "demo", "rmu.theory"))[:, 1]
    assert abs(oracle.chi2(od, theory) - 761.405865008929) < 1e-9
    chi2, st = oracle.chi2_batch(om, od, m.values[None, :])
    assert abs(chi2[0] - 761.405865008929) < 1e-9


def test_multi_theory(rmu):
    """demo/multi.theory: 29 radii x ell = 0,2,4 ({40:180}*29 includes both end points)."""
    m, d, om, od = rmu
    mt = np.loadtxt(os.path.join(G, "demo", "multi.theory"))[:, 1]
    rr = W.parse_binning("{40:180}*29")
    vals = np.array([om.evaluate_multipole(m.values, r, l, 2.25, 16) for r in rr for l in (0, 2, 4)])
    assert np.max(np.abs(vals - mt)) / np.max(np.abs(mt)) < 1e-13


def test_cspline_is_natural_and_clamps():
    x = np.linspace(0.0, 10.0, 21)
    y = np.sin(x)
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(x, y, bc_type="natural")
    q = np.linspace(0.0, 10.0, 333)
    assert np.max(np.abs(oracle.cspline(x, y, q) - cs(q))) < 1e-14
    # outside the table: the end-point value (confirmed by data/BOSSDR11QSOLyaF.scan corners)
    assert oracle.cspline(x, y, [-3.0])[0] == y[0]
    assert oracle.cspline(x, y, [42.0])[0] == y[-1]


def test_hankel_quadrature_against_cosmocalc_tables():
    """models/DR9LyaMocksLCDM.{0,2,4}.dat were produced by cosmocalc (same cosmo library the
    k-space model uses) from DR9LyaMocksLCDM_matterpower.dat with relerr ~1e-3: our fine-grid
    Filon quadrature must reproduce them within that tolerance. This is the only anchor the
    reference offers for the k-space transform (parity with cosmo itself is unpinned)."""
    k, p = W.load_table(os.path.join(G, "models", "DR9LyaMocksLCDM_matterpower.dat"))
    r, x0, x2, x4 = W.load_multipoles(os.path.join(G, "models"), "DR9LyaMocksLCDM")
    for rq in (20, 60, 100, 150):
        i = int(np.argmin(np.abs(r - rq)))
        for ell, tab in ((0, x0), (2, x2), (4, x4)):
            h = oracle.hankel_table(k, p, ell, r[i])
            assert abs(h - tab[i]) < 2e-3 * abs(tab[i]) + 2e-7, (rq, ell, h, tab[i])


def test_qso_rspace_mask_and_shapes():
    """config/BOSSDR11QSOLyaF.ini: 32 x 16 grid, r in [40,180] keeps exactly 440 bins."""
    m, d = W.qso_rspace()
    assert d.n_data == 440 and m.desc.npar == 31
    assert m.names[9] == "BAO alpha-parallel" and m.names[10] == "BAO alpha-perp"  # data/README.txt:41
    assert m.names[16] == "dist add z0 mu0 r+0 rP0 rT-3" and m.names[30] == "dist add z0 mu0 r+0 rP2 rT1"


def test_kspace_oracle_plain_vs_kaiser_limit():
    """With SigmaNL = 0 the k-space model must reduce to Kaiser multipoles of the plain Hankel
    transforms: xi_l(r) = C_l(beta) * H_l[P](r). Checks the projection + Lagrange + Filon chain."""
    m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
    m.configure("fix[SigmaNL-perp]=0")
    om = oracle.OracleModel(m)
    p = m.values.copy()
    beta, beta2 = p[m.index("beta")], p[m.index("beta2*bias2")] / p[m.index("bias2")]
    r, tabs = om.transform(p, 2.35717, jstep=83)
    k, pf = W.load_table(os.path.join(G, "models", "DR9LyaMocksLCDM_matterpower.dat"))
    _, pn = W.load_table(os.path.join(G, "models", "DR9LyaMocksLCDMSB_matterpower.dat"))
    bavg, bprod = 0.5 * (beta + beta2), beta * beta2
    C = {0: 1 + 2 / 3 * bavg + bprod / 5, 2: 4 / 3 * bavg + 4 / 7 * bprod, 4: 8 / 35 * bprod}
    sel = np.nonzero(~np.isnan(tabs[0, 0]))[0]
    for j in sel:
        for li, ell in enumerate((0, 2, 4)):
            want = C[ell] * oracle.hankel_table(k, pn, ell, r[j])
            assert abs(tabs[1, li, j] - want) < 1e-9 * abs(want) + 1e-14
