import sys, os; sys.path.insert(0,'/root/repo')
import numpy as np
from baofit_b200 import workloads as W, host_api as H
tree=W.make_golden_tree('/tmp/gt_probe')
base=tree+'/config'
text=open(base+'/BOSSDR9LyaF.ini').read()
pub=np.loadtxt(tree+'/data/BOSSDR9LyaF.scan')
for orad in (None, 0.0, 5.05e-5):
    t=text if orad is None else text+"\nomega-radiation = %g\n"%orad
    s=H.Session(t, base)
    scan=s.parameter_scan()
    diff=scan['chi2']-scan['best_chi2']-pub[:,2]
    print('orad',orad,'ndata',s.ndata,'chi2min %.4f'%scan['best_chi2'],'max %.4g min %.4g frac1.5e-3 %.4f rms %.4g'%(diff.max(),diff.min(),np.mean(np.abs(diff)<1.5e-3), np.sqrt(np.mean(np.clip(diff,-0.01,0.01)**2))))
