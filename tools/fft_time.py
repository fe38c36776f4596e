"""Times the kernels of the FFT-model path (per-kernel CUDA events) for one batch size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baofit_b200 import capi, workloads as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
ctx = capi.Context(0)
m, coords = W.dr11_auto_fft()
r, mu, z, keep = coords
n = len(keep)
d = W.make_data(r, mu, z, keep, np.zeros(n), np.eye(n))
gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
rng = np.random.Generator(np.random.PCG64(1))
p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})
for i in range(2):
    capi.chi2_batch(ctx, gm, gd, p)
ctx.set_profiling(True)
for i in range(5):
    capi.chi2_batch(ctx, gm, gd, p)
for k, (ms, cnt) in sorted(ctx.get_profile().items(), key=lambda kv: -kv[1][0]):
    print("%-40s %8.4f ms/launch" % (k, ms / cnt))
