#!/usr/bin/env python3
"""Small workload for ncu captures: a few searches of one BASELINE configuration through the C ABI.
    python tools/ncu_driver.py cfg5 [captures]     (cfg5: receiver farm, default 128 captures)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flydog_sdr_gps_b200 as F
from flydog_sdr_gps_b200 import scenarios, synth

cfg = sys.argv[1]
n_cap = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if cfg == "cfg5" else 1)
table = scenarios.table(cfg)
kw = scenarios.params_kw(cfg)
k = kw.get("k_noncoh", 1)
caps = np.stack([synth.make_capture(900 + c, k, table, scenarios.signals("cfg1" if cfg == "cfg5" else cfg, c)) for c in range(min(n_cap, 8))])
caps = caps[np.arange(n_cap) % len(caps)]
with F.AcqEngine(table, F.default_params(**kw)) as eng:
    for _ in range(5):
        rec = eng.search(caps.reshape(-1))
    print(cfg, "captures", n_cap, "tiles per search", eng.tiles_per_search() * n_cap, "detected", int((rec["snr"] >= kw.get("thr_l1", 16.0)).sum()))
