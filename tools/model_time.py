"""Per-kernel CUDA-event timings of one chi2 batch for a named workload: qso_rspace, dr9_rspace, dr9_kspace, dr11_k."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baofit_b200 import capi, workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "qso_rspace"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 37888
ctx = capi.Context(0)
if name == "dr11_k":
    m, (r, mu, z, keep) = W.dr11_auto_kspace()
    n = len(keep)
    d = W.make_data(r, mu, z, keep, np.zeros(n), np.eye(n))
else:
    m, d = getattr(W, name)()
gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
rng = np.random.Generator(np.random.PCG64(1))
p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.7, 1.3)})
for i in range(2):
    capi.chi2_batch(ctx, gm, gd, p)
ctx.set_profiling(True)
for i in range(5):
    capi.chi2_batch(ctx, gm, gd, p)
tot = 0
for k, (ms, cnt) in sorted(ctx.get_profile().items(), key=lambda kv: -kv[1][0]):
    print("%-44s %8.4f ms/launch" % (k, ms / cnt))
    tot += ms / 5
print("%s: N_data %d, B %d, %.3f ms per batch, %.1f M evals/s" % (name, gd.n, B, tot, B / tot / 1e3))
