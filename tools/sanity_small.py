"""Tiny pass over every kernel family (for compute-sanitizer --tool memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baofit_b200 import capi, workloads as W  # noqa: E402

ctx = capi.Context(0)
rng = np.random.Generator(np.random.PCG64(1))


def run(name, m, d, B, G=3):
    gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
    p = m.random_batch(B, rng)
    chi2, st = capi.chi2_batch(ctx, gm, gd, p)
    pred, _ = capi.eval_batch(ctx, gm, gd, p)
    nf = B // G
    gram, _ = capi.gram_batch(ctx, gm, gd, p[:nf * G], G)
    toys = capi.Toys(ctx, gd, np.nan_to_num(pred[:, 0]), seed=1, first_toy=0, ntoy=3)
    c2, _ = capi.chi2_batch_toys(ctx, gm, gd, p, toys, np.arange(B) % 3)
    print(name, "ok", float(np.nanmax(chi2)), int((st != 0).sum()))
    toys.close(), gm.close(), gd.close()


run("rspace", *W.qso_rspace(), 70)
run("rspace-large-sorted", *W.qso_rspace(), 2100)
run("kspace+dm+metals", *W.qso_kspace(), 70)
run("kspace+dm+metals, fused kernel (TMEM-parked accumulators)", *W.qso_kspace(), 1100)
# repeated small calls: graph capture + replay; probes with synthesised linear columns
m9, d9 = W.dr9_kspace()
gm, gd = capi.Model(ctx, m9), capi.Data(ctx, d9)
p = m9.random_batch(3, rng)
for _ in range(4):
    capi.chi2_batch(ctx, gm, gd, p)
steps = np.zeros_like(p)
fl = [j for j in np.nonzero(m9.floating)[0] if not m9.names[j].startswith("BAO alpha")]
steps[:, fl] = 0.01 * m9.errors[fl]
capi.normal_probes(ctx, gm, gd, p, steps)
print("graphs + linear probes ok")
gm.close(), gd.close()
m, d = W.qso_kspace(with_dist_matrix=False, with_metals=False)
run("kspace plain", m, d, 9)
fm, fc = W.dr11_auto_fft(ngrid=64, spacing=8., nbins=20)
n = len(fc[3])
fd = W.make_data(fc[0], fc[1], fc[2], fc[3], np.zeros(n), np.eye(n))
run("fft", fm, fd, 13)
fm2, _ = W.dr11_auto_fft(ngrid=48, spacing=12., nbins=20, fit_nl_correction=True, fft_rmax=250.)
run("fft mixed keys", fm2, fd, 7)
mm, md = W.demo_multi(lmax=4)
run("multipole", mm, md, 11)
print("done")
