#!/usr/bin/env python3
"""Static SASS instruction mix of every kernel in libacq_b200.so -> profiles/sass_mix_<kernel>.txt (one per product kernel).
    python tools/sass_mix.py [tag]        (tag defaults to r2)
What to look for: LDTM / STTM (tcgen05.ld / tcgen05.st: tensor memory as thread-private storage), UBLKCP (cp.async.bulk:
TMA bulk copies), SYNCS (mbarrier), UCGABAR (cluster barrier), FFMA2 / FADD2 / FMUL2 (packed f32x2), no UTC*MMA (the
path is not a contraction: tensor cores are not used, as BASELINE.json's north_star says)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flydog_sdr_gps_b200", "csrc", "libacq_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = dict(re.findall(r"Function (\S+):\n\s+(REG:.*)", res))
summary = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = f.split("\n")[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    base = re.sub(r"^void ", "", name)
    base = re.sub(r"\(acq::SearchArgs\)$|\((?:[^()]|\([^()]*\))*\)$", "", base).replace("acq::", "").replace("(bool)", "")
    short = re.sub(r"[^A-Za-z0-9_]+", "_", base).strip("_")
    if not short.startswith("k_"):
        continue   # micro-benchmark kernels (acq_microbench.cu)
    ops, full = collections.Counter(), collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
        ops[m.group(1).split(".")[0]] += 1
        full[m.group(1)] += 1
    total = sum(ops.values())
    packed = sum(v for k, v in ops.items() if k in ("FFMA2", "FADD2", "FMUL2"))
    scalar = sum(v for k, v in ops.items() if k in ("FFMA", "FADD", "FMUL"))
    key = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in ("LDTM", "STTM", "UBLKCP", "SYNCS", "UCGABAR", "LDS",
                                                                            "STS", "LDG", "STG", "BAR", "SHFL", "ATOMG", "MEMBAR")}
    tc = sum(v for k, v in ops.items() if "MMA" in k)   # UTC*MMA / HMMA / ...: tensor-core math (none expected)
    out = os.path.join(ROOT, "profiles", "sass_mix_%s_%s.txt" % (tag, short))
    with open(out, "w") as fh:
        fh.write("%s\n%s\nsm_100a SASS of flydog_sdr_gps_b200/csrc/libacq_b200.so (cuobjdump -sass), static counts\n" % (name, usage.get(mangled, "")))
        fh.write("instructions %d | packed f32x2 (FFMA2+FADD2+FMUL2) %d | scalar fp32 %d | tensor-core MMA %d\n" % (total, packed, scalar, tc))
        fh.write("tensor memory: LDTM %(LDTM)d STTM %(STTM)d | TMA bulk copy UBLKCP %(UBLKCP)d, mbarrier SYNCS %(SYNCS)d | cluster barrier UCGABAR %(UCGABAR)d\n" % key)
        fh.write("shared LDS %(LDS)d STS %(STS)d | global LDG %(LDG)d STG %(STG)d ATOMG %(ATOMG)d MEMBAR %(MEMBAR)d | BAR %(BAR)d SHFL %(SHFL)d\n\n" % key)
        for k, v in full.most_common():
            fh.write("%6d  %s\n" % (v, k))
    summary.append((short, total, packed, key["LDTM"], key["STTM"], key["UBLKCP"], key["UCGABAR"], tc, usage.get(mangled, "")[:7]))
print("%-34s %6s %6s %5s %5s %6s %7s %3s" % ("kernel", "instr", "f32x2", "LDTM", "STTM", "UBLKCP", "UCGABAR", "MMA"))
for row in sorted(summary):
    print("%-34s %6d %6d %5d %5d %6d %7d %3d  %s" % row)
