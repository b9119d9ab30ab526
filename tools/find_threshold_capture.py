#!/usr/bin/env python3
"""Find a noise-only capture whose strongest false E1B peak (50 PRNs x 81 bins x 16368 lags, oracle search) lands
within 3 % of the detection threshold 16 (gps/search.cpp:549) -- the fixture of
tests/test_gpu_parity.py::test_e1b_noise_only_capture_near_the_threshold.  Prints the seeds it tried."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flydog_sdr_gps_b200 import scenarios, synth
from oracle import oracle_py as O

table = scenarios.table("cfg3")
prm = O.default_params(**scenarios.params_kw("cfg3"))
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 61000, 62000):
    cap = synth.make_capture(seed, 1, table, [])
    top = float(O.search(cap, table, params=prm)["snr"].max())
    print(seed, round(top, 3), flush=True)
    if abs(top / 16.0 - 1) < 0.03 and abs(top / 16.0 - 1) > 0.002:
        print("FOUND", seed, top)
        break
