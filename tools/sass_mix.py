#!/usr/bin/env python3
"""Instruction mix of the search kernels in libacq_b200.so (static SASS counts).  python tools/sass_mix.py [pattern]"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "k_search"
lib = "flydog_sdr_gps_b200/csrc/libacq_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
        ops[m.group(1).split(".")[0]] += 1
    fp = sum(v for k, v in ops.items() if k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
    print(name[:80], "total", sum(ops.values()), "fp", fp)
    print("   ", ops.most_common(22))
