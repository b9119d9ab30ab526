#!/bin/bash
# compute-sanitizer passes over one small search of every kernel family (smoke-sized inputs).
out=gpurun_out/${1:-san}; mkdir -p $out
cat > /tmp/san_case.py <<'P'
import numpy as np, sys
sys.path.insert(0, '.')
import flydog_sdr_gps_b200 as F
from flydog_sdr_gps_b200 import sats as S, synth
table = S.reference_table()
cap = synth.make_capture(4242, 2, table, [(2, 4000, 4 * F.BIN_HZ, 48, 1.0), (44, 30000, -6 * F.BIN_HZ, 47, 0.4)])
sel = np.array([2, 5, 44, 58], np.int32)
with F.AcqEngine(table) as eng:                                   # K = 1: k_search_l1, k_search_e1b, cluster FFT, k_pick_small
    r = eng.search(cap[:8192], sel=sel); f = eng.refine(r)
    r2 = eng.search(np.concatenate([cap[:8192]] * 40), sel=sel)   # many rows: k_fwd_fft (non-cluster), still k_pick_small
    r5 = eng.search(np.concatenate([cap[:8192]] * 70), sel=sel)   # 280 rows: k_best_dop, records by copy
    assert r5[0].tobytes() == r[0].tobytes() and r5[69].tobytes() == r[0].tobytes()
with F.AcqEngine(table, F.default_params(k_noncoh=2)) as eng:    # k_search_l1_multi, k_search_e1b_multi (82 E1B tiles > SMs/4)
    r6 = eng.search(cap, sel=sel)
with F.AcqEngine(table, F.default_params(k_noncoh=2, half_bin=1, dop_lo=-6, dop_hi=6)) as eng:  # MULTI kernels, E1B cluster
    r3 = eng.search(cap, sel=sel); f3 = eng.refine(r3)
# 2-bit sign/magnitude captures (k_front_end<MAG>) with code-Doppler copies (n_shift > 1) in the MULTI kernels and k_refine
kw = dict(k_noncoh=8, half_bin=1, dop_lo=-80, dop_hi=80, sample_bits=2, code_doppler=1)  # 3 shifted copies
cap2 = synth.make_capture(4243, 8, table, [(2, 4000, 38 * F.BIN_HZ, 48, 1.0), (44, 30000, -39 * F.BIN_HZ, 47, 0.4)],
                          sample_bits=2, code_doppler=True)
with F.AcqEngine(table, F.default_params(**kw)) as eng:
    r4 = eng.search(cap2, sel=sel); f4 = eng.refine(r4)
print("ok", r["snr"], r3["snr"], f["dop_hz"], r4["snr"], f4["dop_hz"], r6["snr"])
P
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $out/$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $out/$tool.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out/$tool.log | tail -1)"
done
