"""Bitwise repeatability of the fused model + chi2 kernel: the same batch through the device-pointer entry point several
times, through the host-buffer entry point (pieces on the copy stream) and in a different column order."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from baofit_b200 import capi, workloads as W  # noqa: E402

ctx = capi.Context(0)
m, d = W.qso_kspace()
gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
B = 37888
rng = np.random.Generator(np.random.PCG64(12))
p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})
dev = torch.device("cuda", 0)
d_p = torch.from_numpy(p).to(dev)
d_chi2 = torch.empty(B, dtype=torch.float64, device=dev)
d_st = torch.empty(B, dtype=torch.int32, device=dev)
outs = []
for i in range(4):
    capi.chi2_batch_dev(ctx, gm, gd, d_p.data_ptr(), B, d_chi2.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    outs.append(d_chi2.cpu().numpy().copy())
host, st = capi.chi2_batch(ctx, gm, gd, p)
perm = rng.permutation(B)
hp, _ = capi.chi2_batch(ctx, gm, gd, p[perm])
ok = st == 0
def cmp(name, a, b):
    bad = ok & ~((a == b) | (np.isnan(a) & np.isnan(b)))
    rel = np.abs(a[ok] - b[ok]) / np.abs(b[ok])
    print("%-28s differing %6d of %d, max rel %.3e, first %s" % (name, bad.sum(), ok.sum(), np.nanmax(rel), np.flatnonzero(bad)[:8]))
for i in range(1, 4):
    cmp("dev run %d vs run 0" % i, outs[i], outs[0])
cmp("host (pieces) vs dev", host, outs[0])
back = np.empty(B); back[perm] = hp
cmp("permuted host vs host", back, host)
small, _ = capi.chi2_batch(ctx, gm, gd, p[:4736])
cmp2 = np.abs(small - host[:4736]) / np.abs(host[:4736])
print("first 4736 alone vs in batch: max rel %.3e, differing %d" % (np.nanmax(cmp2[st[:4736] == 0]), np.sum((small != host[:4736]) & (st[:4736] == 0))))
gm.close(), gd.close(), ctx.close()
