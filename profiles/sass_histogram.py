"""Per-kernel SASS mnemonic histogram of libviwb.so (cuobjdump -sass): the evidence for which pipes a kernel uses -- DMMA (FP64 tensor), DFMA,
UTMALDG (TMA tile loads), IDP (dp2a / dp4a), LDGSTS (cp.async), REDUX, LDL / STL (local-memory spills).
Usage: python profiles/sass_histogram.py [lib] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "viw-fusion_b200", "csrc", "libviwb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = hist.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UTMALDG", "IDP", "IMAD", "LDGSTS", "REDUX", "SHFL", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "SYNCS"]
print("# %s  (cuobjdump -sass, instruction counts per kernel; all = every instruction)" % os.path.relpath(lib, ROOT))
print("%-34s %7s " % ("kernel", "all") + " ".join("%7s" % k for k in KEY))
for name, c in hist.items():
    print("%-34s %7d " % (name[-34:], sum(c.values())) + " ".join("%7d" % c.get(k, 0) for k in KEY))
