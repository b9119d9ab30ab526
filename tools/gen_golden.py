#!/usr/bin/env python3
"""Generate tests/golden/*.npz: synthetic captures plus the answers of the UNMODIFIED reference
(gps/search.cpp compiled behind stubs, oracle/_ref) for every row of the reference's Sats[] table.

Runs only in the development container (needs /root/reference to build oracle/_ref).  The
fixtures travel with the repository; the GPU box checks the oracle and the CUDA engine against
them without the reference tree.

Fixture content (per file):
  captures  uint8 [n, 8192]   packed 1-bit blocks (the bytes Sample() reads over SPI)
  signals   float64 [n, m, 5] injected (sat, tau, doppler_hz, cn0, phase), NaN padded
  dop, lag  int32  [n, 59]    Correlate()'s *max_snr_dop, *max_snr_i   (gps/search.cpp:495)
  snr       float32 [n, 59]   Correlate()'s return value               (gps/search.cpp:498)
The FFT under the reference here is the oracle FFT (FFTW is not installable): see DESIGN.md.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flydog_sdr_gps_b200 import sats as S, synth  # noqa: E402
from flydog_sdr_gps_b200.engine import BIN_HZ  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    O.build(ref=True)
    assert O.have_ref(), "oracle/_ref could not be built"
    table = S.reference_table()
    assert table == O.ref_sats()
    rng = np.random.default_rng(20261017)
    caps, sigs = [], []
    # three GPS-heavy captures (mix of Navstar, one QZSS, one E1B each), one noise only, one E1B-heavy
    for n in range(5):
        sig = []
        if n < 3:
            for k, sat in enumerate(rng.choice(32, 7, replace=False)):
                f = float(rng.integers(-19, 20)) * BIN_HZ + (0.0 if k % 2 == 0 else float(rng.uniform(-60, 60)))
                sig.append((int(sat), int(rng.integers(0, 16368)), f, float(50 - 1.5 * k), float(rng.uniform(0, 6.28))))
            sig.append((32 + n, int(rng.integers(0, 16368)), float(rng.integers(-19, 20)) * BIN_HZ, 46.0, 1.0))
            sig.append((36 + 3 * n, int(rng.integers(0, 65472)), float(rng.integers(-19, 20)) * BIN_HZ, 46.0, 2.0))
        elif n == 4:
            for sat in 36 + rng.choice(23, 5, replace=False):
                sig.append((int(sat), int(rng.integers(0, 65472)), float(rng.integers(-19, 20)) * BIN_HZ,
                            float(rng.uniform(44, 49)), float(rng.uniform(0, 6.28))))
        caps.append(synth.make_capture(7000 + n, 1, table, sig))
        sigs.append(sig)
    m = max(len(s) for s in sigs)
    sig_arr = np.full((len(caps), m, 5), np.nan)
    for i, s in enumerate(sigs):
        for j, row in enumerate(s):
            sig_arr[i, j] = row
    sel = np.arange(len(table), dtype=np.int32)
    dop = np.zeros((len(caps), len(table)), np.int32)
    lag = np.zeros_like(dop)
    snr = np.zeros(dop.shape, np.float32)
    for i, c in enumerate(caps):
        dop[i], lag[i], snr[i] = O.ref_search(c, sel)
        det = [(S.label(table[k]), int(dop[i, k]), int(lag[i, k]), round(float(snr[i, k]), 1)) for k in sel if snr[i, k] >= 16]
        print("capture", i, "detected", det)
    out = os.path.join(ROOT, "tests", "golden", "ref_search_59sats.npz")
    np.savez_compressed(out, captures=np.stack(caps), signals=sig_arr, dop=dop, lag=lag, snr=snr)
    print("wrote", out, os.path.getsize(out), "bytes")

    # stage fixtures: forward-FFT input (x2) and spectrum checksum for capture 0, code replica checksums
    x2, D = O.ref_sample(caps[0])
    code_sum = np.array([np.abs(O.ref_code_spectrum(k)).astype(np.float64).sum() for k in sel])
    code_x2_sum = np.array([O.ref_code_baseband(k).real.astype(np.float64).sum() for k in sel])
    out2 = os.path.join(ROOT, "tests", "golden", "ref_stages.npz")
    np.savez_compressed(out2, x2=x2, D=D, code_abs_sum=code_sum, code_x2_sum=code_x2_sum,
                        code_x2_sat8=O.ref_code_baseband(8), code_x2_sat40=O.ref_code_baseband(40))
    print("wrote", out2, os.path.getsize(out2), "bytes")

    # literal SearchTask loop (gps/search.cpp:512-604): two passes over Sats[] with 12 free tracking
    # channels; Sample() call k reads capture k % 5.  Event kinds: 1 ChanReset(sat, codegen_init) -> ch,
    # 2 ChanStart(ch, sat, lo_shift, ca_shift, snr), 3 GPSstat(STAT_SAT, snr, ch, sat, weak), 4 GPSstat(STAT_DOP, ch, hz, ca_shift)
    ev = O.ref_search_task(np.stack(caps), passes=2, free_chans=12, min_sig=16)
    ev = ev[ev["kind"] != 5]
    out3 = os.path.join(ROOT, "tests", "golden", "ref_search_task_events.npz")
    np.savez_compressed(out3, events=ev)
    print("wrote", out3, len(ev), "events;", int((ev["kind"] == 2).sum()), "ChanStart calls")


def gen_codes_and_e1b50():
    """Fixtures that pin the code tables to the REFERENCE's own code classes, independently of the product's packed
    tables (csrc/e1b_codes.inc, data/e1b_codes.bin) and of the oracle's copy:
      ref_code_chips_sha256.json  SHA-256 of the chip sequence (one byte per chip, 0/1) of every Galileo E1-B PRN 1..50
                                  (E1BCODE, gps/e1bcode.h:63-92) and of every C/A row of Sats[] (CACODE, gps/cacode.h)
      ref_e1b50.npz               the UNMODIFIED search.cpp over a table of all 50 E1-B codes (oracle/_ref/
                                  libref_search_e1b50.so): code-spectrum fingerprints per PRN and Correlate()'s answers
                                  on a capture whose signals were synthesised from the reference's chips"""
    import hashlib
    import json
    O.build(ref=True)
    assert O.have_ref() and O.ref50() is not None
    sha = {"e1b": {}, "ca": {}}
    for prn in range(1, 51):
        sha["e1b"]["%d" % prn] = hashlib.sha256(O.ref_e1b_chips(prn).tobytes()).hexdigest()
    for row in S.navstar() + S.qzss():
        sha["ca"]["%d" % row[0]] = hashlib.sha256(O.ref_ca_chips(row[1], row[2]).tobytes()).hexdigest()
    sha["_doc"] = "sha256 of the chips (uint8 0/1, one per chip) from the reference's E1BCODE / CACODE; tools/gen_golden.py codes"
    out = os.path.join(ROOT, "tests", "golden", "ref_code_chips_sha256.json")
    json.dump(sha, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)

    table = S.e1b(range(1, 51))
    rng = np.random.default_rng(20261018)
    # PRNs outside the 23 the reference's Sats[] activates get most of the signals
    inactive = [p for p in range(1, 51) if p not in S._E1B_ACTIVE]
    prns = list(rng.choice(inactive, 5, replace=False)) + [11, 36]
    sig = [(int(p - 1), int(rng.integers(0, 65472)), float(rng.integers(-19, 20)) * BIN_HZ, float(rng.uniform(45, 49)),
            float(rng.uniform(0, 6.28))) for p in prns]
    cap = synth.make_capture(7050, 1, table, sig, chip_source=lambda row: (O.ref_e1b_chips(row[0]), True))
    sel = np.arange(50, dtype=np.int32)
    dop, lag, snr = O.ref50_search(cap, sel)
    print("e1b50 detected", [(int(k) + 1, int(dop[k]), int(lag[k]), round(float(snr[k]), 1)) for k in sel if snr[k] >= 16])
    idx = np.sort(rng.choice(16384, 64, replace=False)).astype(np.int32)
    spec = np.stack([O.ref50_code_spectrum(k) for k in range(50)])
    out2 = os.path.join(ROOT, "tests", "golden", "ref_e1b50.npz")
    np.savez_compressed(out2, capture=cap, signals=np.array(sig), dop=dop, lag=lag, snr=snr, spec_idx=idx,
                        spec_bins=spec[:, idx], spec_abs_sum=np.abs(spec).astype(np.float64).sum(axis=1))
    print("wrote", out2, os.path.getsize(out2), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "codes":
        gen_codes_and_e1b50()
    else:
        main()
