#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped libgprb200.so (cuobjdump -sass; runs without a GPU).
Evidence of which hardware paths the kernels use: DMMA (fp64 tensor pipe, mma.sync.m8n8k4.f64), UTMALDG (tensor-map TMA
loads), UBLKCP (1-D bulk copies through the TMA engine), SYNCS (mbarrier), DFMA/MUFU (fp64 vector pipe), and the absence of
UTCMMA/LDTM (tcgen05 has no f64 kind).  Usage: python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpr.jl_b200", "libgprb200.so")
KEY = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UTCMMA",
       "LDTM", "HMMA", "IMMA", "ATOM", "RED", "UCGABAR", "ACQBULK", "MEMBAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
fn, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (sm_100a); columns: " + " ".join(KEY))
tot = collections.Counter()
for mangled, name in zip(hist, names):
    h = hist[mangled]
    short = re.sub(r"\(.*", "", name).replace("gprb::", "")
    cols = " ".join(f"{k}={h[k]}" for k in KEY if h[k])
    print(f"{short:45s} total={sum(h.values()):6d}  {cols}")
    tot.update(h)
print("# library totals: " + " ".join(f"{k}={tot[k]}" for k in KEY))
