"""CPU suite (-m "not gpu"): the oracle against the reference's known answers, the golden vectors
produced by the unmodified reference, and numpy; plus host logic."""
import os

import numpy as np
import pytest

from flydog_sdr_gps_b200 import sats as S, scenarios, synth

# IS-GPS-200 "first 10 chips, octal" column for PRN 1..32 (the reference's own
# L1_PRN_TEST / QZSS_PRN_TEST printers, gps/search.cpp:208-237, print exactly these quantities).
FIRST10_OCTAL = [0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454, 0o1626, 0o1504, 0o1642, 0o1750,
                 0o1764, 0o1772, 0o1775, 0o1776, 0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706,
                 0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712]
QZSS_FIRST10 = {194: 0o0170, 195: 0o0030, 196: 0o0472, 199: 0o1050}


def first_bits(chips, n):
    v = 0
    for i in range(n):
        v = (v << 1) | int(chips[i])
    return v


def test_ca_code_known_answers(oracle):
    for (prn, t1, t2, _), want in zip(S.navstar(), FIRST10_OCTAL):
        assert first_bits(oracle.ca_chips(t1, t2), 10) == want, prn
    # gps/search.cpp:209-219 prints the first 16 chips of PRN 9
    assert first_bits(oracle.ca_chips(3, 10), 16) == 0xE5A9
    for prn, d, init, _ in S.qzss():
        chips = oracle.ca_chips(d, init)
        assert first_bits(chips, 10) == QZSS_FIRST10[prn]
        assert first_bits(chips, 10) == (0o1777 ^ init)  # G1 starts all ones


def test_ca_code_properties(oracle):
    for prn, t1, t2, _ in S.navstar():
        c = oracle.ca_chips(t1, t2).astype(np.int32)
        assert c.sum() in (511, 512)  # balanced Gold code
        b = 1 - 2 * c
        ac = np.array([np.dot(b, np.roll(b, k)) for k in (1, 2, 3, 100, 511)])
        assert set(ac.tolist()) <= {-1, 63, -65}  # three-valued autocorrelation


def test_e1b_known_answers(oracle):
    # expected values are printed by the reference's E1BCODE_TEST (gps/search.cpp:295,302)
    assert first_bits(oracle.e1b_chips(1), 20) == 0xF5D71
    assert first_bits(oracle.e1b_chips(2), 20) == 0x96B85
    for prn in (1, 11, 36, 50):
        assert np.array_equal(oracle.e1b_chips(prn), synth.e1b_chips(prn))


@pytest.mark.skipif(not os.path.exists("/root/reference/gps/e1bcode.h"), reason="reference tree not present")
def test_e1b_table_matches_reference_strings(oracle):
    import re
    strings = re.findall(r'"([0-9A-F]{1023})"', open("/root/reference/gps/e1bcode.h").read())
    assert len(strings) == 50
    for prn in range(1, 51):
        chips = oracle.e1b_chips(prn)
        bits = "".join(format(int(ch, 16), "04b") for ch in strings[prn - 1])
        assert "".join(map(str, chips)) == bits


def test_host_and_oracle_code_generators_agree(oracle):
    for prn, t1, t2, _ in S.navstar() + S.qzss():
        assert np.array_equal(synth.ca_chips(t1, t2), oracle.ca_chips(t1, t2))


def test_oracle_fft_against_numpy(oracle):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(16384) + 1j * rng.standard_normal(16384)).astype(np.complex64)
    f = oracle.fft16384(x, -1)
    g = np.fft.fft(x.astype(np.complex128))
    assert np.abs(f - g).max() / np.abs(g).max() < 1e-6
    b = oracle.fft16384(x, +1)
    h = np.fft.ifft(x.astype(np.complex128)) * 16384  # unnormalised like FFTW_BACKWARD
    assert np.abs(b - h).max() / np.abs(h).max() < 1e-6
    # impulse at k -> pure complex exponential, exact to rounding
    imp = np.zeros(16384, np.complex64)
    imp[5] = 1
    e = oracle.fft16384(imp, +1)
    assert np.abs(e - np.exp(2j * np.pi * 5 * np.arange(16384) / 16384)).max() < 1e-6
    # round trip
    rt = oracle.fft16384(oracle.fft16384(x, -1), +1) / 16384
    assert np.abs(rt - x).max() < 1e-5


def test_half_band_matches_direct_convolution(oracle):
    import ctypes as C
    rng = np.random.default_rng(5)
    n = 4096
    buf = np.zeros(2 * (n + 31), np.float32)
    buf[:2 * n] = rng.standard_normal(2 * n).astype(np.float32)
    x = buf[:2 * n].copy().view(np.complex64).astype(np.complex128)
    oracle.lib().orc_hb_decimate(n, buf.ctypes.data_as(C.POINTER(C.c_float)))
    y = buf[:n].view(np.complex64)
    taps = np.zeros(31)
    even = [-0.010233, 0.010668, -0.016324, 0.024377, -0.036482, 0.056990, -0.101993, 0.316926,
            0.316926, -0.101993, 0.056990, -0.036482, 0.024377, -0.016324, 0.010668, -0.010233]
    taps[0::2] = even
    taps[15] = 0.500009
    xp = np.concatenate([x, np.zeros(31)])
    want = np.array([np.dot(taps, xp[2 * o:2 * o + 31]) for o in range(n // 2)])
    assert np.abs(y - want).max() < 1e-5


def test_golden_search_vectors(oracle, golden_search):
    """The oracle reproduces the unmodified reference (Sample + Correlate) on every golden capture:
    Doppler bin and lag bit-exact, snr to float rounding (same FFT, same operation order)."""
    table = S.reference_table()
    for i, cap in enumerate(golden_search["captures"]):
        rec = oracle.search(cap, table)
        assert np.array_equal(rec["dop"], golden_search["dop"][i])
        assert np.array_equal(rec["lag"], golden_search["lag"][i])
        np.testing.assert_allclose(rec["snr"], golden_search["snr"][i], rtol=2e-6)
        np.testing.assert_allclose(rec["peak"] / rec["noise"], rec["snr"], rtol=1e-6)


def test_golden_injected_signals_are_found(golden_search):
    """Sanity of the fixtures themselves: strong injected signals come back at lag = tau/4 and the nearest bin."""
    table = S.reference_table()
    for i in range(len(golden_search["captures"])):
        for sat, tau, dop_hz, cn0, _ in golden_search["signals"][i]:
            if np.isnan(sat):
                continue
            sat = int(sat)
            L = 16368 if table[sat][3] == S.E1B else 4092
            if cn0 >= 47:
                assert golden_search["snr"][i, sat] >= 16
            if golden_search["snr"][i, sat] < 20:
                continue
            d = (golden_search["lag"][i, sat] - tau / 4.0) % L  # tau is in FS samples, lag in /4 samples
            # E1B: the capture holds one 4 ms period + 64 samples, so for tau beyond half a period the
            # circular correlation peaks on the alignment of the wrapped part, 16 lags later
            assert min(d, L - d) <= 1.0 or (L == 16368 and abs(d - 16) <= 1.0)
            assert golden_search["dop"][i, sat] == int(np.round(dop_hz / 249.755859375))


def test_golden_stage_vectors(oracle, golden_search, golden_stages):
    cap = golden_search["captures"][0]
    assert np.array_equal(oracle.capture_baseband(cap), golden_stages["x2"])       # bit-exact before the FFT
    np.testing.assert_allclose(oracle.capture_spectrum(cap), golden_stages["D"], rtol=0, atol=1e-3)
    table = S.reference_table()
    assert np.array_equal(oracle.code_baseband(table[8]), golden_stages["code_x2_sat8"])
    assert np.array_equal(oracle.code_baseband(table[40]), golden_stages["code_x2_sat40"])
    for k in (0, 8, 33, 40, 58):
        assert abs(np.abs(oracle.code_spectrum(table[k])).astype(np.float64).sum() / golden_stages["code_abs_sum"][k] - 1) < 1e-6
        assert oracle.code_baseband(table[k]).real.astype(np.float64).sum() == pytest.approx(
            golden_stages["code_x2_sum"][k], rel=1e-9, abs=1e-6)


def test_oracle_against_live_reference(oracle):
    """Where oracle/_ref exists (development container, or shipped prebuilt), run the unmodified
    reference side by side on a fresh capture, including the negative-Doppler row overrun."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    table = S.reference_table()
    assert table == oracle.ref_sats()
    cap = synth.make_capture(99, 1, table, [(3, 777, -4000.0, 47, 0.1), (40, 40000, 2000.0, 47, 0.2), (33, 9, -250.0, 46, 0.3)])
    x2, D = oracle.ref_sample(cap)
    assert np.array_equal(oracle.capture_baseband(cap), x2)
    assert np.array_equal(oracle.capture_spectrum(cap), D)
    sel = np.array([0, 3, 31, 32, 33, 35, 36, 40, 58], np.int32)
    dop, lag, snr = oracle.ref_search(cap, sel)
    rec = oracle.search(cap, table, sel=sel)
    assert np.array_equal(rec["dop"], dop) and np.array_equal(rec["lag"], lag) and np.array_equal(rec["snr"], snr)
    # the intended circular wrap differs measurably for negative Doppler: the quirk matters
    circ = oracle.search(cap, table, sel=sel, params=oracle.default_params(wrap_mode=oracle.WRAP_CIRCULAR))
    assert circ["snr"][1] != rec["snr"][1]
    for k in (0, 20, 35, 36, 58):
        assert np.array_equal(oracle.code_spectrum(table[k]), oracle.ref_code_spectrum(k))


def test_extensions_reduce_to_reference(oracle, golden_search):
    """Half-bin indexing and K=1 'multi-block' must give the reference answers on the even half-bins."""
    table = S.navstar()
    cap = golden_search["captures"][1]
    base, bgrid = oracle.search(cap, table, want_grid=True)
    hb, hgrid = oracle.search(cap, table, params=oracle.default_params(dop_lo=-40, dop_hi=40, half_bin=1), want_grid=True)
    assert np.array_equal(hgrid["lag"][:, 0::2], bgrid["lag"])
    assert np.array_equal(hgrid["snr"][:, 0::2], bgrid["snr"])
    # a half-bin Doppler is recovered on the odd index
    f = 7.5 * 249.755859375
    cap2 = synth.make_capture(5, 1, table, [(4, 4000, f, 46, 0.3)])
    r = oracle.search(cap2, table, sel=[4], params=oracle.default_params(dop_lo=-40, dop_hi=40, half_bin=1))
    assert r["dop"][0] == 15 and r["lag"][0] == 1000


def test_noncoherent_accumulation_gain(oracle):
    """K blocks with the 16-lag-per-block code advance removed: a 33 dB-Hz signal invisible at K=1 is found at K=20."""
    table = S.navstar()
    cap = synth.make_capture(11, 20, table, [(7, 8000, 1000.0, 34, 0.5)])
    one = oracle.search(cap[:8192], table, sel=[7])
    many, grid = oracle.search(cap, table, sel=[7], params=oracle.default_params(k_noncoh=20), want_grid=True)
    assert many["lag"][0] == 2000 and many["dop"][0] == 4
    assert not (one["lag"][0] == 2000 and one["dop"][0] == 4 and one["snr"][0] >= 16)
    assert many["snr"][0] > 1.5 * np.median(grid["snr"])


def test_workload_sizes():
    """SURVEY.md 8(d) / BASELINE.md cell and tile counts."""
    def cells(cfg):
        t = scenarios.table(cfg)
        kw = scenarios.params_kw(cfg)
        n_dop = kw.get("dop_hi", 20) - kw.get("dop_lo", -20) + 1
        c = sum(n_dop * (16368 if r[3] == S.E1B else 4092) for r in t) * scenarios.n_captures(cfg)
        tiles = len(t) * n_dop * kw.get("k_noncoh", 1) * scenarios.n_captures(cfg)
        return c, tiles
    assert cells("cfg1") == (5368704, 1312)
    assert cells("cfg2") == (21081984, 103040)
    assert cells("cfg3") == (66290400, 4050)
    assert cells("cfg4") == (38923104, 3362)
    assert cells("cfg5") == (5497552896, 1343488)


def test_shard_partition():
    for n in (1, 7, 82, 1024):
        for w in (1, 2, 3, 4, 8):
            parts = [scenarios.shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_synth_capture_format():
    table = S.navstar()
    cap = synth.make_capture(1, 2, table, scenarios.signals("cfg1", 1))
    assert cap.dtype == np.uint8 and cap.size == 2 * 8192
    assert 0.48 < np.unpackbits(cap).mean() < 0.52
    assert np.array_equal(cap, synth.make_capture(1, 2, table, scenarios.signals("cfg1", 1)))


def _code_err(code_fs, tau, period_fs):
    e = (float(code_fs) - tau) % period_fs
    return e - period_fs if e > period_fs / 2 else e


def test_refinement_recovers_injected_doppler_and_code_phase(oracle):
    """orc_refine: the directly evaluated peak equals the FFT search's peak, the interpolated
    Doppler lands within a quarter bin and the code phase within 2 FS samples of what was injected (the search
    alone is good to half a bin = 125 Hz and 4 FS samples)."""
    table = S.reference_table()
    bin_hz = 16.368e6 / 65536
    sig = [(2, 4001, 3.3 * bin_hz, 50, 1.0), (10, 12346, -7.45 * bin_hz, 48, 2.0), (44, 30003, -6.2 * bin_hz, 50, 0.4)]
    cap = synth.make_capture(77, 1, table, sig)
    sel = np.array([2, 10, 44, 20], np.int32)
    rec = oracle.search(cap, table, sel=sel)
    fine = oracle.refine(cap, table, rec)
    assert np.allclose(fine["peak"], rec["peak"], rtol=1e-4)  # same cell, direct sum against the inverse FFT
    for i, (sat, tau, fd, _, _) in enumerate(sig):
        L = 16368 if table[sat][3] == S.E1B else 4092
        assert rec["snr"][i] >= 16
        assert abs(fine["dop_hz"][i] - fd) < 0.25 * bin_hz, (sat, fine["dop_hz"][i], fd)
        assert abs(_code_err(fine["code_fs"][i], tau, 4 * L)) < 2.0, (sat, fine["code_fs"][i], tau)
        assert abs(fine["dop_hz"][i] - rec["dop"][i] * bin_hz) <= bin_hz          # stays inside the neighbour bins
        assert abs(fine["code_fs"][i] - 4 * rec["lag"][i]) <= 4.0                  # and the neighbour lags
        assert fine["ca_shift"][i] == int(np.rint(fine["code_fs"][i])) % (4 * L)


def test_refinement_with_half_bins_and_noncoherent_blocks(oracle):
    """K = 4 blocks, half-bin Doppler indices: the peak identity and the Doppler estimate hold with the per-block lag
    advance (n + 16 b) and the pre-rotated capture spectrum of odd indices."""
    table = S.navstar()
    bin_hz = 16.368e6 / 65536
    prm = oracle.default_params(k_noncoh=4, half_bin=1, dop_lo=-20, dop_hi=20)
    cap = synth.make_capture(78, 4, table, [(7, 8002, 4.5 * bin_hz, 46, 0.5), (12, 100, -3.1 * bin_hz, 46, 0.1)])
    rec = oracle.search(cap, table, sel=[7, 12], params=prm)
    fine = oracle.refine(cap, table, rec, params=prm)
    assert np.allclose(fine["peak"], rec["peak"], rtol=1e-4)
    assert rec["dop"][0] == 9 and rec["dop"][1] in (-6, -7)
    assert abs(fine["dop_hz"][0] - 4.5 * bin_hz) < 0.25 * bin_hz
    assert abs(fine["dop_hz"][1] + 3.1 * bin_hz) < 0.25 * bin_hz
    assert abs(_code_err(fine["code_fs"][0], 8002, 4 * 4092)) < 2.0


def _half_band_f64(x):
    """Half-band /2 of search.cpp:140-166 in double precision (zero padding past the end, float-narrowed taps)."""
    taps = np.zeros(31)
    taps[0::2] = [-0.010233, 0.010668, -0.016324, 0.024377, -0.036482, 0.056990, -0.101993, 0.316926,
                  0.316926, -0.101993, 0.056990, -0.036482, 0.024377, -0.016324, 0.010668, -0.010233]
    taps[15] = 0.500009
    taps = taps.astype(np.float32).astype(np.float64)
    xp = np.concatenate([x.astype(np.complex128), np.zeros(31, np.complex128)])
    out = np.zeros(len(x) // 2, np.complex128)
    for j in np.flatnonzero(taps):
        out += taps[j] * xp[j:j + len(x):2]
    return out


def test_sign_magnitude_capture_format(oracle):
    """2-bit sign/magnitude captures (extension; include/acq_b200.h ACQ_CAPTURE_BLOCK_BYTES): a block is the reference's
    sign plane followed by a magnitude plane, sample = (sign ? -1 : +1) * (mag ? 3 : 1)."""
    table = S.navstar()
    sig = scenarios.signals("cfg1", 1)
    c1 = synth.make_capture(1, 2, table, sig)
    c2 = synth.make_capture(1, 2, table, sig, sample_bits=2)
    assert c2.size == 2 * 16384
    planes = c2.reshape(2, 2, 8192)
    assert np.array_equal(planes[:, 0].reshape(-1), c1)                  # same sign bits as the 1-bit capture
    assert 0.30 < np.unpackbits(planes[:, 1]).mean() < 0.36              # |s| > 0.98 sigma: about one third
    g1 = oracle.gen_capture(5, 2, table, sig)
    g2 = oracle.gen_capture(5, 2, table, sig, sample_bits=2)
    assert np.array_equal(g2.reshape(2, 2, 8192)[:, 0].reshape(-1), g1)
    # an all-zero magnitude plane reproduces the reference's 1-bit front end bit for bit
    blk = c1[:8192]
    z = np.concatenate([blk, np.zeros(8192, np.uint8)])
    assert np.array_equal(oracle.capture_baseband(z, 0, 2), oracle.capture_baseband(blk))
    assert np.array_equal(oracle.capture_baseband(z, 1, 2), oracle.capture_baseband(blk, 1))
    # a set magnitude bit scales the mixed sample by 3: all-ones plane = 3 x the 1-bit baseband (exact in fp32: every
    # product and partial sum scales by 3 only up to rounding, so compare to 1e-6)
    o = np.concatenate([blk, np.full(8192, 0xFF, np.uint8)])
    b3 = oracle.capture_baseband(o, 0, 2)
    b1 = oracle.capture_baseband(blk)
    assert np.abs(b3 - 3 * b1).max() < 1e-5 * np.abs(b1).max()
    # independent restatement in double precision
    blk2 = c2[:16384]
    bits = np.unpackbits(blk2[:8192], bitorder="little").astype(np.int64)
    mag = np.unpackbits(blk2[8192:], bitorder="little").astype(np.int64)
    i = np.arange(65536)
    w = 1.0 + 2.0 * mag
    x0 = w * (1.0 - 2.0 * (bits ^ np.array([1, 1, 0, 0])[i & 3])) + 1j * w * (1.0 - 2.0 * (bits ^ np.array([1, 0, 0, 1])[i & 3]))
    want = _half_band_f64(_half_band_f64(x0))
    got = oracle.capture_baseband(blk2, 0, 2)
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()


def test_sign_magnitude_search_gains_over_one_bit(oracle):
    """Same signals, same noise: the 4-level capture loses 0.55 dB to quantisation, the sign-only one 1.96 dB, so
    strong satellites come out with a visibly larger peak-to-noise ratio and identical decisions."""
    table = S.navstar()
    sig = [(3, 4000, 4 * 249.755859375, 47, 0.3), (17, 9000, -11 * 249.755859375, 46, 1.3), (25, 120, 0.0, 48, 2.0)]
    gains = []
    for seed in (1, 2, 3):
        c1 = synth.make_capture(seed, 1, table, sig)
        c2 = synth.make_capture(seed, 1, table, sig, sample_bits=2)
        r1 = oracle.search(c1, table, sel=[3, 17, 25])
        r2 = oracle.search(c2, table, sel=[3, 17, 25], params=oracle.default_params(sample_bits=2))
        assert np.array_equal(r1["dop"], r2["dop"]) and np.array_equal(r1["lag"], r2["lag"])
        assert np.array_equal(r2["dop"], [4, -11, 0]) and np.array_equal(r2["lag"], [1000, 2250, 30])
        gains.append(r2["snr"] / r1["snr"])
    g = np.mean(gains)
    assert 1.15 < g < 1.7, g   # 1.41 dB in amplitude^2 terms ~ x1.38 on a peak-dominated max/mean ratio
    # K = 2 blocks of a 2-bit capture: block stride is 16384 bytes
    c2 = synth.make_capture(9, 2, table, sig, sample_bits=2)
    prm = oracle.default_params(sample_bits=2, k_noncoh=2)
    r = oracle.search(c2, table, sel=[3, 17, 25], params=prm)
    assert np.array_equal(r["dop"], [4, -11, 0]) and np.array_equal(r["lag"], [1000, 2250, 30])
    fine = oracle.refine(c2, table, r, params=prm)
    assert np.allclose(fine["peak"], r["peak"], rtol=1e-4)


def test_code_doppler_compensation(oracle):
    """Extension: with the code stretched by the carrier offset, a long
    non-coherent sum at +-9.4 kHz smears the peak over 2 lags; taking block b at lag n + 16 b + s(b, h) restores it.
    s = round(b h / 385) (h/2 for half-bins): FS/4/f_L1 = 1/385 exactly."""
    assert 16.368e6 / 4 / 1575.42e6 == pytest.approx(1 / 385, rel=1e-12)
    table = S.navstar()
    BIN = 16.368e6 / 65536
    K = 20
    kw = dict(dop_lo=-80, dop_hi=80, half_bin=1, k_noncoh=K, thr_l1=2.6)
    sig = [(4, 5000, 38.0 * BIN, 36, 0.2), (20, 16000, -37.5 * BIN, 36, 1.2), (9, 800, 1.0 * BIN, 36, 0.7)]
    cap = synth.make_capture(5, K, table, sig, code_doppler=True)
    sel = [4, 20, 9, 11]
    r0 = oracle.search(cap, table, sel=sel, params=oracle.default_params(**kw))
    r1 = oracle.search(cap, table, sel=sel, params=oracle.default_params(code_doppler=1, **kw))
    assert np.array_equal(r1["dop"][:3], [76, -75, 2]) and np.array_equal(r1["lag"][:3], [1250, 4000, 200])
    assert (r1["snr"][:2] > 1.08 * r0["snr"][:2]).all()           # high Doppler: the smeared peak is recovered
    assert r1[2].tobytes() == r0[2].tobytes()                      # |h| = 2: s(b, 2) = 0 for every b < 20 -- untouched
    # without code Doppler in the signal the compensation would hurt, i.e. it is not a no-op that "always helps"
    cap0 = synth.make_capture(5, K, table, sig)
    q0 = oracle.search(cap0, table, sel=sel, params=oracle.default_params(**kw))
    q1 = oracle.search(cap0, table, sel=sel, params=oracle.default_params(code_doppler=1, **kw))
    assert (q1["snr"][:2] < q0["snr"][:2]).all()
    # refinement follows the same per-block lag
    fine = oracle.refine(cap, table, r1, params=oracle.default_params(code_doppler=1, **kw))
    assert np.allclose(fine["peak"], r1["peak"], rtol=1e-4)
    # K = 1: nothing to compensate
    one = cap[:8192]
    a = oracle.search(one, table, sel=sel)
    b = oracle.search(one, table, sel=sel, params=oracle.default_params(code_doppler=1))
    assert a.tobytes() == b.tobytes()
    # the oracle's generator applies the same code stretch
    g = oracle.gen_capture(3, 4, table, [(4, 5000, 38.0 * BIN, 60, 0.2)], code_doppler=True)
    prm = oracle.default_params(dop_lo=-40, dop_hi=40, k_noncoh=4, code_doppler=1)
    r = oracle.search(g, table, sel=[4], params=prm)
    assert r["dop"][0] == 38 and r["lag"][0] == 1250


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, np.uint8).tobytes()).hexdigest()


def test_code_tables_match_reference_sha256(oracle):
    """Every code table in the repository -- the oracle's, the product's packed ICD table (csrc/e1b_codes.inc, as the
    engine unpacks it) and the generator's copy (data/e1b_codes.bin) -- against SHA-256 digests of the chips the
    REFERENCE's own E1BCODE / CACODE classes produce (tests/golden/ref_code_chips_sha256.json, tools/gen_golden.py
    codes).  Runs without the reference tree: a packing error can no longer be common to oracle and product."""
    import json
    import re
    from conftest import GOLDEN, ROOT
    from flydog_sdr_gps_b200 import sats as S, synth
    want = json.load(open(os.path.join(GOLDEN, "ref_code_chips_sha256.json")))
    assert len(want["e1b"]) == 50 and len(want["ca"]) == 36
    text = open(os.path.join(ROOT, "flydog_sdr_gps_b200", "csrc", "e1b_codes.inc")).read()
    words = np.array([int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]+)", text)], np.uint64).astype(np.uint32)
    assert words.size == 50 * 128
    for prn in range(1, 51):
        w = words[(prn - 1) * 128: prn * 128]
        inc_chips = ((w[np.arange(4092) >> 5] >> (np.arange(4092) & 31).astype(np.uint32)) & 1).astype(np.uint8)
        for name, chips in (("oracle", oracle.e1b_chips(prn)), ("e1b_codes.inc", inc_chips), ("e1b_codes.bin", synth.e1b_chips(prn))):
            assert _sha(chips) == want["e1b"][str(prn)], (name, prn)
    for row in S.navstar() + S.qzss():
        for name, chips in (("oracle", oracle.ca_chips(row[1], row[2])), ("synth", synth.ca_chips(row[1], row[2]))):
            assert _sha(chips) == want["ca"][str(row[0])], (name, row)


def test_all_50_e1b_codes_against_reference_goldens(oracle):
    """The oracle over all 50 Galileo E1-B codes against the unmodified search.cpp run over a 50-row table
    (oracle/_ref/libref_search_e1b50.so; fixture tests/golden/ref_e1b50.npz): code spectra and Correlate() answers
    bit-identical (same FFT underneath)."""
    from conftest import GOLDEN
    from flydog_sdr_gps_b200 import sats as S
    g = np.load(os.path.join(GOLDEN, "ref_e1b50.npz"))
    table = S.e1b(range(1, 51))
    for k in range(50):
        c = oracle.code_spectrum(table[k])
        assert np.array_equal(c[g["spec_idx"]], g["spec_bins"][k]), k
        assert abs(np.abs(c).astype(np.float64).sum() - g["spec_abs_sum"][k]) <= 1e-9 * g["spec_abs_sum"][k]
    rec = oracle.search(g["capture"], table)
    assert np.array_equal(rec["dop"], g["dop"]) and np.array_equal(rec["lag"], g["lag"])
    assert np.array_equal(rec["snr"], g["snr"])
    injected = {int(s[0]) for s in g["signals"]}
    assert injected <= {int(k) for k in np.nonzero(g["snr"] >= 16)[0]}
    if oracle.ref50() is not None:   # live check where the reference build exists
        dop, lag, snr = oracle.ref50_search(g["capture"], np.arange(50, dtype=np.int32))
        assert np.array_equal(dop, g["dop"]) and np.array_equal(lag, g["lag"]) and np.array_equal(snr, g["snr"])


def test_sbas_rows_keep_the_reference_all_zero_spectrum(oracle):
    """SearchInit builds replicas for Navstar / QZSS rows (search.cpp:244) and E1B rows (:306) only: an SBAS row of the
    table keeps the all-zero code[] row of the reference's static array, so Correlate() can never detect it (snr = 0/0
    fails `snr > max_snr`) -- and, in the reference's wrap mode, the satellite BEFORE it reads zeros instead of a next
    row at negative Doppler."""
    from flydog_sdr_gps_b200 import sats as S, scenarios, synth
    nav = S.navstar()
    table = nav[:3] + [(120, 145, 0o1106, S.SBAS)] + nav[3:5]
    assert not oracle.code_spectrum(table[3]).any()
    cap = synth.make_capture(5, 1, nav, scenarios.signals("cfg1", 5))
    rec, grid = oracle.search(cap, table, want_grid=True)
    # reference wrap: at negative Doppler the zero row still reads the first |dop| bins of the NEXT row (search.cpp:471),
    # so its cells are not all zero there -- but it has no peak of its own and stays far below the threshold
    assert (grid[3]["peak"][20:] == 0).all() and rec[3]["snr"] < 8
    circ = oracle.search(cap, table, params=oracle.default_params(wrap_mode=oracle.WRAP_CIRCULAR))
    assert circ[3]["snr"] == 0 and circ[3]["lag"] == 0 and circ[3]["dop"] == 0 and circ[3]["peak"] == 0
    # the row before the SBAS row sees zeros past its end at negative Doppler: same as being last in the table
    alone = oracle.search(cap, nav[:3], want_grid=True)[1]
    assert np.array_equal(grid[2]["snr"], alone[2]["snr"])
    # and it differs from what it would read if a real spectrum followed
    follow = oracle.search(cap, nav[:5], want_grid=True)[1]
    assert not np.array_equal(grid[2]["snr"][:20], follow[2]["snr"][:20])


def test_ofast_reference_build_gives_the_same_decisions(oracle, golden_search):
    """oracle/_ref/libref_search_ofast.so -- the unmodified sources at the reference's own -Ofast (+AVX2/FMA), bench.py's
    timed CPU baseline -- against the strict parity build on the golden captures: identical (dop, lag) for every
    satellite that is not in a numerical tie, snr within 1e-5.  (Its floats are not reproducible bit for bit, which is
    why it is never used for parity.)"""
    if not (os.path.exists(oracle.REF_OFAST_SO) and oracle._cpu_has("avx2", "fma")):
        pytest.skip("no -Ofast reference build usable on this host")
    lib = oracle.ref_bench()
    assert "Ofast" in oracle.REF_BENCH_BUILD
    sel = np.arange(59, dtype=np.int32)
    for i, cap in enumerate(golden_search["captures"][:2]):
        dop, lag, snr = oracle.ref_search(cap, sel, lib=lib)
        np.testing.assert_allclose(snr, golden_search["snr"][i], rtol=1e-5)
        strong = golden_search["snr"][i] >= 16
        assert np.array_equal(dop[strong], golden_search["dop"][i][strong])
        assert np.array_equal(lag[strong], golden_search["lag"][i][strong])
