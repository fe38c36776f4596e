"""Host-side plumbing: binning strings, grid indexing, final cuts, broadband grammar, loaders."""
import os

import numpy as np
import pytest

from baofit_b200 import desc as D, workloads as W


def test_binning_conventions():
    c = W.parse_binning("[40:200]*40")
    assert len(c) == 40 and c[0] == 42.0 and c[-1] == 198.0
    s = W.parse_binning("{40:180}*29")
    assert len(s) == 29 and s[0] == 40.0 and s[-1] == 180.0 and abs(s[1] - 45.0) < 1e-12
    e = W.parse_binning("{2.0,2.5,3.0}")
    assert list(e) == [2.0, 2.5, 3.0]
    with pytest.raises(ValueError):
        W.parse_binning("40:200")


def test_grid_index_order_last_axis_fastest():
    r, mu, z = W.grid_coordinates(np.array([1.0, 2.0]), np.array([0.1, 0.2, 0.3]), np.array([2.0, 3.0]), "comoving-polar")
    assert list(r[:6]) == [1.0] * 6 and list(mu[:4]) == [0.1, 0.1, 0.2, 0.2] and list(z[:4]) == [2.0, 3.0, 2.0, 3.0]


def test_final_cuts_counts_match_survey():
    a1, a2, a3 = W.parse_binning(W.QSO_AXIS1), W.parse_binning(W.QSO_AXIS2), W.parse_binning(W.QSO_AXIS3)
    r, mu, z = W.grid_coordinates(a1, a2, a3, "comoving-cartesian")
    keep = W.final_cuts(r, mu, z, np.ones(len(r), bool), rmin=40, rmax=180)
    assert len(r) == 512 and len(keep) == 440
    m, (r, mu, z, keep) = W.dr11_auto_kspace()
    assert len(r) == 2500 and len(keep) == 1515


def test_data_loader_ignores_trailing_tokens(tmp_path):
    p = tmp_path / "x.data"
    p.write_text("0 1.5\n2 0.25 72\n")
    val, has = W.load_data_vector(str(p), 4)
    assert list(has) == [True, False, True, False] and val[2] == 0.25


def test_broadband_grammar():
    s = D.parse_broadband("rP,rT=0:2,-3:1", 14, 100.0, 2.3)
    assert (s.rp_min, s.rp_max, s.rp_step, s.rt_min, s.rt_max, s.rt_step) == (0, 2, 1, -3, 1, 1)
    assert s.ncoef() == 15 and (s.r_min, s.r_max, s.mu_min, s.mu_max) == (0, 0, 0, 0)
    s = D.parse_broadband("-2:0,0:4:2,0", 13, 100.0, 2.25)
    assert (s.r_min, s.r_max, s.r_step, s.mu_min, s.mu_max, s.mu_step, s.z_min, s.z_max) == (-2, 0, 1, 0, 4, 2, 0, 0)
    assert s.ncoef() == 9
    s = D.parse_broadband("r,mu=-4:-1:-2,0", 0, 100.0, 2.25)  # negative step = fractional exponents
    assert (s.r_min, s.r_max, s.r_step, s.r_denom) == (-8, -2, 1, 2)
    with pytest.raises(ValueError):
        D.parse_broadband("r,mu=0:2,0:9", 0, 100.0, 2.25)
    with pytest.raises(ValueError):
        D.parse_broadband("rP,rT=2:0,0", 0, 100.0, 2.25)


def test_parameter_layouts_match_reference_order():
    m, _ = W.qso_kspace()
    n = m.names
    assert n[:7] == ["beta", "(1+beta)*bias", "gamma-bias", "gamma-beta", "delta-v", "bias2", "beta2*bias2"]
    assert n[7:9] == ["SigmaNL-perp", "1+f"] and n[9:14] == ["BAO amplitude", "BAO alpha-iso", "BAO alpha-parallel",
                                                            "BAO alpha-perp", "gamma-scale"]
    assert n[14] == "beta Si2a" and n[21] == "bias Si3" and n[22] == "dist add z0 mu0 r+0 rP0 rT-3"
    assert len(n) == 37  # SURVEY.md appendix C, config 4 with metals
