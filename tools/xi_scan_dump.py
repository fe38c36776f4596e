"""Dumps the BOSSDR9LyaFXi scan of the GPU path next to the published surface (diagnostics)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from baofit_b200 import host_api as H, workloads as W
tree = W.make_golden_tree(os.path.join(tempfile.gettempdir(), "xi_tree"))
base = os.path.join(tree, "config")
s = H.Session(open(os.path.join(base, "BOSSDR9LyaFXi.ini")).read(), base)
scan = s.parameter_scan()
pub = np.loadtxt(os.path.join(tree, "data", "BOSSDR9LyaFXi.scan"))
np.savez("gpurun_out/xi_scan.npz", chi2=scan["chi2"], best=scan["best_chi2"], points=scan["points"], status=scan["status"], pub=pub, names=np.array(s.names))
fit = s.fit()
print("fit", fit["status"], 2 * fit["fval"], dict(zip(s.names, fit["values"])))
