"""Timing of the FFT-model toy-MC leg (BASELINE configs[4]) with BAOFIT_VERBOSE_FITS."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from baofit_b200 import capi, host_api, workloads as W

ctx = capi.Context(0)
m, coords = W.dr11_auto_fft()
def predict(model, data, params):
    gm, gd = capi.Model(ctx, model), capi.Data(ctx, data)
    pred, _ = capi.eval_batch(ctx, gm, gd, params)
    return pred[:, 0]
d, _ = W.synthetic_dataset(m, coords, predict)
tree = W.make_golden_tree(os.path.join(tempfile.gettempdir(), "probe_fft"))
W.write_dataset_files(os.path.join(tree, "data", "BOSSDR11LyaF_synth"), 2500, coords[3], d.keep._arrays[4], d.keep._arrays[5])
text = open(os.path.join(tree, "config", "BOSSDR11LyaF_fft.ini")).read().replace("data = ../data/BOSSDR11LyaF", "data = ../data/BOSSDR11LyaF_synth")
s = host_api.Session(text, os.path.join(tree, "config"))
s.toymc(0, 4, seed=1)
os.environ["BAOFIT_VERBOSE_FITS"] = "1"
for rep in range(3):
    t0 = time.perf_counter(); res = s.toymc(0, int(sys.argv[1]) if len(sys.argv) > 1 else 1250, seed=1966); print("toys %.3f s, failed %d" % (time.perf_counter() - t0, int((res["status"] == 2).sum())))
