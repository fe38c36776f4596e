"""Timing breakdown of the published 2100-point QSO scan (r-space) with BAOFIT_VERBOSE_FITS."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baofit_b200 import host_api, workloads as W
tree = W.make_golden_tree(os.path.join(tempfile.gettempdir(), "probe_qso"))
s = host_api.Session(open(os.path.join(tree, "config", os.environ.get("INI", "BOSSDR11QSOLyaF.ini"))).read(), os.path.join(tree, "config"))
s.parameter_scan(0, 2)
os.environ["BAOFIT_VERBOSE_FITS"] = os.environ.get("DBG", "1")
for rep in range(2):
    t0 = time.perf_counter(); res = s.parameter_scan(); print("scan %.3f s" % (time.perf_counter() - t0))
