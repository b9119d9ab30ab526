#!/usr/bin/env python3
"""Bring-up script for a GPU box: stage-by-stage comparison of the CUDA engine with the oracle,
plus the on-device micro-benchmarks and a first timing.  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flydog_sdr_gps_b200 as F  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    res = {}
    mb = F.microbench(0)
    print("microbench", json.dumps(mb))
    res["microbench"] = mb
    sats = F.sats.reference_table()
    eng = F.AcqEngine(sats)
    print("device", eng.device_info())
    # code spectra
    worst = 0
    for i in (0, 8, 31, 32, 35, 36, 58):
        c = eng.code_spectrum(i)
        o = O.code_spectrum(sats[i])
        worst = max(worst, rel(c, o))
    print("code spectrum max rel err", worst)
    sig = [(2, 0, 0.0, 50, 0.3), (6, 4000, 4 * F.BIN_HZ, 47, 1.0), (10, 8188, -10 * F.BIN_HZ, 45, 2.0),
           (13, 12000, 19 * F.BIN_HZ, 44, 0.1), (18, 332, -20 * F.BIN_HZ, 43, 0.5),
           (21, 16000, 3 * F.BIN_HZ + 40, 42, 0.7), (27, 5000, 7.5 * F.BIN_HZ, 41, 0.2), (30, 4, 8 * F.BIN_HZ, 40, 0.9),
           (44, 30000, -1500.0, 45, 0.4)]
    cap = O.gen_capture(1234, 1, sats, sig)
    x2, D = eng.capture_spectrum(cap)
    ox2 = O.capture_baseband(cap)
    oD = O.capture_spectrum(cap)
    print("x2 bit-exact", bool(np.array_equal(x2, ox2)), "D rel err", rel(D, oD))
    x2h, Dh = eng.capture_spectrum(cap, 1)
    print("x2 half-rot bit-exact", bool(np.array_equal(x2h, O.capture_baseband(cap, 1))), "D rel err",
          rel(Dh, O.capture_spectrum(cap, 1)))
    # search, reference defaults, whole table
    t = time.time()
    rec, grid = eng.search(cap, want_grid=True)
    print("gpu search (incl. first-call alloc) %.3f s" % (time.time() - t))
    orec, ogrid = O.search(cap, sats, want_grid=True)
    rec = rec[0]
    grid = grid[0]
    print("dop equal", int((rec["dop"] == orec["dop"]).sum()), "/", len(sats), " lag equal",
          int((rec["lag"] == orec["lag"]).sum()))
    print("snr max rel err", float(np.abs(rec["snr"] / orec["snr"] - 1).max()), "peak max rel err",
          float(np.abs(rec["peak"] / orec["peak"] - 1).max()))
    print("grid snr max rel err", float(np.abs(grid["snr"] / ogrid["snr"] - 1).max()), "grid lag equal",
          int((grid["lag"] == ogrid["lag"]).sum()), "/", grid.size)
    for i in range(len(sats)):
        if orec["snr"][i] >= 16 or rec["snr"][i] >= 16 or rec["dop"][i] != orec["dop"][i] or rec["lag"][i] != orec["lag"][i]:
            print("  sat", i, F.sats.label(sats[i]), "gpu", rec["dop"][i], rec["lag"][i], rec["snr"][i], "oracle",
                  orec["dop"][i], orec["lag"][i], orec["snr"][i])
    # timing: cfg1-like (32 Navstar), repeated
    sel = np.arange(32, dtype=np.int32)
    for _ in range(3):
        eng.search(cap, sel=sel)
    n = 20
    t = time.time()
    for _ in range(n):
        eng.search(cap, sel=sel)
    dt = (time.time() - t) / n
    print("cfg1 e2e per search %.1f us -> %.3g cells/s, %.3g tiles/s" % (dt * 1e6, eng.cells_per_search(sel) / dt,
                                                                         eng.tiles_per_search(sel) / dt))
    # batched: 256 captures x 32 sats
    caps = np.tile(cap, 256)
    for _ in range(2):
        eng.search(caps, sel=sel)
    t = time.time()
    r = eng.search(caps, sel=sel)
    dt = time.time() - t
    print("batch 256 e2e %.2f ms -> %.3g cells/s, %.3g tiles/s" % (dt * 1e3, 256 * eng.cells_per_search(sel) / dt,
                                                                   256 * eng.tiles_per_search(sel) / dt))
    print("batch consistent", bool((r["lag"] == r["lag"][0]).all() and (r["snr"] == r["snr"][0]).all()))
    res["batch256_tiles_per_s"] = 256 * eng.tiles_per_search(sel) / dt
    json.dump(res, open("gpurun_out/first_light.json", "w"), indent=1)


if __name__ == "__main__":
    main()
