#!/usr/bin/env python3
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --clock-control none --csv`) of the default bench command:
the kernels of every headline step (front end, forward FFT, search, best-Doppler pick) and the search kernel's share.
    python tools/launch_summary.py profiles/r2_launches_default_v3.csv > profiles/r2_launches_default_v3_summary.txt"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 14]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


L = [(int(r[ix["ID"]]), short(r[ix["Kernel Name"]]), r[ix["Grid Size"]], float(r[ix["Metric Value"]].replace(",", "")) / 1e3) for r in rows]
print("ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cufft` (%s: first %d launches of our\n"
      "kernels + NCCL + fill; --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache)." % (sys.argv[1], len(L)))
print("The headline steps (cfg5, 1024 captures): front end, forward FFT, search, best-Doppler pick")
shares = []
for i, (lid, name, grid, us) in enumerate(L):
    if name.startswith("k_front_end") and grid.startswith("(64, 1024"):
        step = L[i:i + 4]
        if len(step) == 4 and step[2][1].startswith("k_search"):
            tot = sum(s[3] for s in step)
            shares.append(step[2][3] / tot)
            print("step at launch %d: %s | search share %.4f" % (lid, "; ".join("%s %s %.1f us" % (s[1], s[2], s[3]) for s in step), step[2][3] / tot))
if shares:
    print("search kernel share of a headline step: %.4f (mean of %d steps)" % (sum(shares) / len(shares), len(shares)))
tot = {}
for _, name, _, us in L:
    tot[name] = tot.get(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += us
print("all launches by kernel (count, total us):")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %5d %12.1f" % (name, n, us))
