"""Wall time of one full fit (CorrelationFitter::fit, Newton driver + HESSE) for the shipped configurations."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baofit_b200 import host_api, workloads as W
tree = W.make_golden_tree(os.path.join(tempfile.gettempdir(), "probe_fit"))
for ini in ("BOSSDR11QSOLyaF.ini", "BOSSDR11QSOLyaF_k.ini", "BOSSDR9LyaF.ini", "BOSSDR9LyaF_k.ini"):
    s = host_api.Session(open(os.path.join(tree, "config", ini)).read(), os.path.join(tree, "config"))
    s.fit()
    t0 = time.perf_counter(); r = s.fit(); dt = time.perf_counter() - t0
    print("%-24s fit %.1f ms, %d evaluations, status %d, chi2 %.3f / %d bins" % (ini, 1e3 * dt, r["nevaluations"], r["status"], 2 * r["fval"], s.ndata))
