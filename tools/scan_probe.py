"""Timing breakdown of the k-space scan (BASELINE configs[2]) with per-kernel events."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from baofit_b200 import capi, host_api, workloads as W
ctx = capi.Context(0)
m, coords = W.dr11_auto_kspace()
def predict(model, data, params):
    gm, gd = capi.Model(ctx, model), capi.Data(ctx, data)
    pred, _ = capi.eval_batch(ctx, gm, gd, params)
    return pred[:, 0]
d, _ = W.synthetic_dataset(m, coords, predict, rel=float(os.environ.get('REL', '1.0')))
tree = W.make_golden_tree(os.path.join(tempfile.gettempdir(), "probe_k"))
W.write_dataset_files(os.path.join(tree, "data", "BOSSDR11LyaF_synthk"), 2500, coords[3], d.keep._arrays[4], d.keep._arrays[5])
text = open(os.path.join(tree, "config", "BOSSDR11LyaF_k.ini")).read().replace("data = ../data/BOSSDR11LyaF", "data = ../data/BOSSDR11LyaF_synthk")
s = host_api.Session(text, os.path.join(tree, "config"))
s.parameter_scan(0, 2)
os.environ["BAOFIT_VERBOSE_FITS"] = os.environ.get("DBG", "1")
first = int(os.environ.get("FIRST", "0")); cnt = int(os.environ.get("COUNT", "-1"))
t0 = time.perf_counter(); res = s.parameter_scan(first, cnt); dt = time.perf_counter() - t0
print("scan %.3f s, status counts %s" % (dt, np.bincount(res["status"], minlength=3)))
if cnt < 0:
    st = res["status"].reshape(35, 60)
    for row in st:
        print("".join(str(v) for v in row))
