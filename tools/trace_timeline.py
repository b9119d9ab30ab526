#!/usr/bin/env python3
"""Timeline of ONE search from the "trace" variant library (%globaltimer stamps of every CTA: entry, after the wait
for the preceding grid, exit).  Run on the GPU box:
    python tools/trace_timeline.py cfg1 [cfg4 ...]        -> one summary line per kernel, times in us from the first stamp"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flydog_sdr_gps_b200 as F
from flydog_sdr_gps_b200 import _lib, scenarios, synth

NAMES = ["front_end", "fwd_fft", "search_l1", "search_e1b", "pick", "e1b_cluster"]
VARIANT = os.environ.get("TRACE_VARIANT", "trace")
L = _lib.load_variant(VARIANT)
L.acq_trace_read.argtypes = [C.c_void_p, C.c_int]
total = L.acq_trace_read(None, 0)
buf = np.zeros(total, np.uint64)

for cfg in sys.argv[1:] or ["cfg1"]:
    table = scenarios.table(cfg)
    kw = scenarios.params_kw(cfg)
    cap = synth.make_capture(1, kw.get("k_noncoh", 1), table, scenarios.signals(cfg, 1))
    with F.AcqEngine(table, F.default_params(**kw), variant=VARIANT) as eng:
        import time
        for _ in range(20):
            eng.search(cap)
        L.acq_trace_read(buf.ctypes.data, total)
        t0 = time.perf_counter()
        eng.search(cap)
        host_us = (time.perf_counter() - t0) * 1e6
        L.acq_trace_read(buf.ctypes.data, total)
    tr = buf.reshape(len(NAMES), -1, 4).astype(np.int64)
    start = tr[tr > 0].min()
    print("== %s: host acq_search() %.1f us" % (cfg, host_us))
    for k, name in enumerate(NAMES):
        used = tr[k][:, 0] > 0
        if not used.any():
            continue
        e, w, x = [(tr[k][used][:, s] - start) / 1e3 for s in range(3)]
        x = x[tr[k][used][:, 2] > 0] if (tr[k][used][:, 2] > 0).any() else np.array([np.nan])
        w = w[tr[k][used][:, 1] > 0] if (tr[k][used][:, 1] > 0).any() else np.array([np.nan])
        extra = ""
        if (tr[k][used][:, 3] > 0).any():
            extra = "  mark3 %6.1f" % ((tr[k][used][:, 3].max() - start) / 1e3)
        print("  %-12s ctas %4d  entry %6.1f..%6.1f  past-wait %6.1f..%6.1f  exit %6.1f..%6.1f (median %6.1f)%s" % (
            name, used.sum(), e.min(), e.max(), w.min(), w.max(), x.min(), x.max(), np.median(x), extra))
