"""Per-kernel counts of the SASS mnemonics that show which hardware paths a kernel uses (cuobjdump -sass of the built
library): DMMA (FP64 tensor pipe), UBLKCP (TMA bulk copy), LDGSTS (cp.async), LDTM / STTM (tensor-memory load / store),
UTCATOMSWS (tcgen05.alloc), USETMAXREG (setmaxnreg), SYNCS (mbarrier), DFMA. Usage: python tools/sass_summary.py > profiles/sass_summary_r02.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "baofit_b200", "libbaofit_b200.so")
KEYS = ["DMMA", "DFMA", "UBLKCP", "LDGSTS", "LDTM", "STTM", "UTCATOMSWS", "USETMAXREG", "SYNCS", "LDS", "STS", "LDL", "STL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        kernels[cur]["_total"] += 1
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass baofit_b200/libbaofit_b200.so (sm_100a): instruction counts per kernel")
print("# %-86s %6s " % ("kernel", "instr") + " ".join("%6s" % k[:6] for k in KEYS))
for (mangled, c), name in sorted(zip(kernels.items(), names), key=lambda kv: kv[1]):
    name = re.sub(r"\(bf::\w+\)|\(.*\)$", "", name).replace("void ", "").replace("bf::", "")
    print("%-88s %6d " % (name[:88], c["_total"]) + " ".join("%6d" % c[k] for k in KEYS))
