"""Small driver of the FFT-model path (config/BOSSDR11LyaF_fft.ini shape) for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baofit_b200 import capi, workloads as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
m, coords = W.dr11_auto_fft()


def predict(model, data, params):
    gm, gd = capi.Model(ctx, model), capi.Data(ctx, data)
    pred, _ = capi.eval_batch(ctx, gm, gd, params)
    return pred[:, 0]


d, _ = W.synthetic_dataset(m, coords, predict)
gm, gd = capi.Model(ctx, m), capi.Data(ctx, d)
rng = np.random.Generator(np.random.PCG64(1))
for i in range(steps):
    p = m.random_batch(B, rng, sweep={"BAO alpha-parallel": (0.8, 1.3), "BAO alpha-perp": (0.6, 1.5)})
    chi2, st = capi.chi2_batch(ctx, gm, gd, p)
print("chi2[:3]", chi2[:3], "bad", int((st != 0).sum()))
