"""Distortion-matrix and metal-template files through the host-side readers:
<name>.dmat (baofit/DistortionMatrix.cc:46-68), <name>_A_B.{0,2,4}.dat / .grid per line combination
(baofit/MetalCorrelationModel.cc:163-231). The same arrays that tests and bench.py hand to the C ABI directly
are written in the on-disk formats, read back through an unmodified-ini Session with the reference's option
names, and must give the same parameter layout (metals before the broadband block) and
the same chi-square to rounding."""
import os

import numpy as np
import pytest

from baofit_b200 import host_api as H, workloads as W
import parity_cases as PC

pytestmark = pytest.mark.gpu


def _write_dmat(path, dm, drop_below=0.0):
    """Sparse text 'i j value' (missing entries = 0, DistortionMatrix.cc:37,46-68)."""
    i, j = np.nonzero(np.abs(dm) > drop_below)
    with open(path + ".dmat", "w") as f:
        np.savetxt(f, np.column_stack([i, j, dm[i, j]]), fmt="%d %d %.17g")


def _write_cross_metals(prefix, mt, names):
    for c, ln in enumerate(names):
        for l in range(3):
            with open("%s_QSO_%s.%d.dat" % (prefix, ln, 2 * l), "w") as f:
                f.write("%d %d %.17g %.17g %.17g %.17g\n" % (mt["nx"], mt["ny"], mt["x0"], mt["y0"], mt["dx"], mt["dy"]))
                np.savetxt(f, mt["cross"][3 * c + l].reshape(-1, 1), fmt="%.17g")
        z = mt["zgrid"][c]
        with open("%s_QSO_%s.grid" % (prefix, ln), "w") as f:
            np.savetxt(f, np.column_stack([np.arange(len(z)), np.zeros(len(z)), np.zeros(len(z)), z]), fmt="%d %.17g %.17g %.17g")


def _capi_chi2(ctx, m, d, params):
    from baofit_b200 import capi
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    chi2, st = capi.chi2_batch(ctx, gm, gd, params)
    gm.close(), gd.close()
    return chi2, st


def test_dmat_and_cross_metal_files_round_trip(ctx, golden_tree):
    """BASELINE config 4 with its add-ons read from files: config/BOSSDR11QSOLyaF_k.ini + dist-matrix +
    metal-model-interpolate."""
    m, d = W.qso_kspace()
    a1, a2 = W.parse_binning(W.QSO_AXIS1), W.parse_binning(W.QSO_AXIS2)
    dm = W.synthetic_dist_matrix(a1, a2)
    mt = W.synthetic_cross_metals(ngrid=512)
    data_dir = os.path.join(golden_tree, "data")
    _write_dmat(os.path.join(data_dir, "synthQSO"), dm)
    _write_cross_metals(os.path.join(data_dir, "synthmetal"), mt, ["Si2a", "Si2b", "Si2c", "Si3"])
    text = open(os.path.join(golden_tree, "config", "BOSSDR11QSOLyaF_k.ini")).read()
    text += ("\ndist-matrix = yes\ndist-matrix-name = ../data/synthQSO\ndist-matrix-order = 512\n"
             "metal-model-interpolate = yes\nmetal-model-name = ../data/synthmetal\nmodel-config = fix[beta Si*]\n")
    s = H.Session(text, os.path.join(golden_tree, "config"))
    assert s.names == m.names and s.npar == 37
    assert s.names[14:22] == ["beta Si2a", "bias Si2a", "beta Si2b", "bias Si2b", "beta Si2c", "bias Si2c", "beta Si3", "bias Si3"]
    assert s.names[22].startswith("dist add")
    assert np.array_equal(s.floating, m.floating) and np.array_equal(s.values, m.values)
    rng = np.random.Generator(np.random.PCG64(41))
    params = m.random_batch(300, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5), "delta-v": (-300, 300)})
    f = s.evaluate(params)
    chi2, st = _capi_chi2(ctx, m, d, params)
    ok = st == 0
    assert ok.sum() > 250
    assert np.max(np.abs(2 * f[ok] - chi2[ok]) / chi2[ok]) < 1e-12
    s.close()
    # a sparse file: entries that are not listed are zero
    dm2 = np.where(np.abs(dm) > 5e-4, dm, 0.0)
    _write_dmat(os.path.join(data_dir, "synthQSOsparse"), dm, drop_below=5e-4)
    s2 = H.Session(text.replace("synthQSO\n", "synthQSOsparse\n"), os.path.join(golden_tree, "config"))
    m2, d2 = W.qso_kspace()
    m2.keep.f64("dist_matrix", dm2)
    chi2b, stb = _capi_chi2(ctx, m2, d2, params[:64])
    fb = s2.evaluate(params[:64])
    okb = stb == 0
    assert np.max(np.abs(2 * fb[okb] - chi2b[okb]) / chi2b[okb]) < 1e-12
    assert np.max(np.abs(chi2b[okb] - chi2[:64][okb]) / chi2b[okb]) > 1e-9  # the dropped entries do matter
    s2.close()


def test_dmat_errors_mirror_reference_messages(built, golden_tree):
    text = open(os.path.join(golden_tree, "config", "BOSSDR11QSOLyaF_k.ini")).read()
    base = os.path.join(golden_tree, "config")
    with pytest.raises(H.HostError) as e:
        H.Session(text + "\ndist-matrix = yes\ndist-matrix-name = ../data/nosuchmatrix\ndist-matrix-order = 512\n", base)
    assert e.value.code == -2 and "DistortionMatrix: Unable to open" in str(e.value)
    bad = os.path.join(golden_tree, "data", "badmatrix")
    with open(bad + ".dmat", "w") as f:
        f.write("0 0 1.0\n3 700 0.5\n")
    with pytest.raises(H.HostError) as e:
        H.Session(text + "\ndist-matrix = yes\ndist-matrix-name = ../data/badmatrix\ndist-matrix-order = 512\n", base)
    assert e.value.code == -2 and "DistortionMatrix" in str(e.value) and "line 2" in str(e.value)
    with pytest.raises(H.HostError) as e:
        H.Session(text + "\nmetal-model-interpolate = yes\nmetal-model-name = ../data/nosuchmetal\n", base)
    assert e.value.code == -2 and "MetalCorrelationModel" in str(e.value)


def test_auto_metal_table_files_round_trip(ctx, golden_tree):
    """metal-model (templates tabulated per grid bin, 14 line combinations) read from <name>_A_B.ell.dat /
    .grid files on the 512-bin QSO grid with an auto-correlation k-space model."""
    m, d, sweep, B, _ = PC.build("metal_auto_table")
    mt = PC._synthetic_auto_metals(512, civ=False)
    l1 = ["Lya", "Lya", "Lya", "Lya", "Si2a", "Si2a", "Si2a", "Si2b", "Si2b", "Si2c", "Si3", "Si3", "Si3", "Si3"]
    l2 = ["Si2a", "Si2b", "Si2c", "Si3", "Si2a", "Si2b", "Si2c", "Si2b", "Si2c", "Si2c", "Si2a", "Si2b", "Si2c", "Si3"]
    prefix = os.path.join(golden_tree, "data", "synthauto")
    idx = np.arange(512)
    for c in range(14):
        for l in range(3):
            with open("%s_%s_%s.%d.dat" % (prefix, l1[c], l2[c], 2 * l), "w") as f:
                np.savetxt(f, np.column_stack([idx, mt["auto"][3 * c + l]]), fmt="%d %.17g")
        with open("%s_%s_%s.grid" % (prefix, l1[c], l2[c]), "w") as f:
            np.savetxt(f, np.column_stack([idx, np.zeros(512), np.zeros(512), mt["zgrid"][c]]), fmt="%d %.17g %.17g %.17g")
    text = open(os.path.join(golden_tree, "config", "BOSSDR11QSOLyaF_k.ini")).read()
    # the same grid and data as an auto-correlation fit: drop the cross-correlation switch and its parameters
    lines = [ln for ln in text.splitlines() if not ln.strip().startswith("cross-correlation")
             and "bias2" not in ln and "delta-v" not in ln]
    text = "\n".join(lines).replace("dist-add = rP,rT=0:2,-3:1", "dist-add = -2:0,0:4:2,0")
    text += "\nmetal-model = yes\nmetal-model-name = ../data/synthauto\n"
    s = H.Session(text, os.path.join(golden_tree, "config"))
    assert s.names == m.names
    params = PC.make_params("metal_auto_table", m, sweep, 64)
    f = s.evaluate(params)
    chi2, st = _capi_chi2(ctx, m, d, params)
    ok = st == 0
    assert ok.sum() > 40
    assert np.max(np.abs(2 * f[ok] - chi2[ok]) / chi2[ok]) < 1e-12
    s.close()
