#!/usr/bin/env python3
"""Host-side latency of acq_search() (C ABI, host buffers) for single-capture searches, per library variant.
    python tools/e2e_latency.py [variant|product ...]     (configs cfg1 cfg4 cfg3, and one satellite = the literal SearchTask step)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flydog_sdr_gps_b200 as F
from flydog_sdr_gps_b200 import scenarios, synth

for variant in sys.argv[1:] or ["product"]:
    for cfg, sel in (("cfg1", None), ("cfg1", np.array([7], np.int32)), ("cfg4", None), ("cfg3", None)):
        table = scenarios.table(cfg)
        kw = scenarios.params_kw(cfg)
        cap = synth.make_capture(1, 1, table, scenarios.signals(cfg, 1))
        with F.AcqEngine(table, F.default_params(**kw), variant=None if variant == "product" else variant) as eng:
            n_sel = len(table) if sel is None else len(sel)
            out = np.zeros(n_sel, F.RECORD_DTYPE)
            for _ in range(50):
                eng.search_ptr(cap.ctypes.data, 1, out.ctypes.data, sel=sel)
            ts = []
            for _ in range(1000):
                t0 = time.perf_counter()
                eng.search_ptr(cap.ctypes.data, 1, out.ctypes.data, sel=sel)
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts) * 1e6
            print("%-8s %s %-9s acq_search median %.1f us  p10 %.1f  min %.1f" % (
                variant, cfg, "all" if sel is None else "1 sat", np.median(ts), np.percentile(ts, 10), ts.min()), flush=True)
