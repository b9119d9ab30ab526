"""Comparison rules shared by the GPU parity tests (SURVEY.md 7 "Exact decisions", BASELINE north_star):
   - acquisition decisions (detected set, lag, Doppler index) bit-exact,
   - peak / noise / snr within RTOL relative,
   - a (lag, Doppler) disagreement is tolerated ONLY inside a numerical tie zone: the oracle's own table
     must show the engine's choice within TIE relative of the oracle's best (two cells whose fp32 values
     differ by rounding noise); such cases are counted and bounded."""
import numpy as np

RTOL = 1e-3      # north_star tolerance on peak and peak-to-noise metrics
TIE = 2e-5       # relative gap below which two fp32 powers/snrs computed by different FFT orders may swap


class Margins:
    """What a comparison could NOT decide bit-exactly, accumulated over a test and printed (pytest -s / the log):
    records whose snr sits within RTOL of the detection threshold (their decision is not compared), tie-zone swaps of
    undetected satellites, and how close the nearest detected / undetected satellite came to the threshold."""

    def __init__(self, label=""):
        self.label = label
        self.records = self.in_threshold_band = self.tie_swaps = self.detected = self.grid_lag_swaps = 0
        self.min_detected_margin = float("inf")     # min over detected sats of snr/thr - 1
        self.min_undetected_margin = float("inf")   # min over undetected sats of 1 - snr/thr
        self.max_rel_err = 0.0

    def __int__(self):
        return self.tie_swaps

    def report(self):
        msg = ("parity margins [%s]: %d records, %d detected, %d inside +-%.0e of the threshold (decision not compared), "
               "%d tie-zone swaps, %d per-Doppler lag swaps inside the tie zone, nearest detected sat %.2f %% above / "
               "nearest undetected %.2f %% below the threshold, max rel. error of peak/noise/snr %.1e" % (
                   self.label, self.records, self.detected, self.in_threshold_band, RTOL, self.tie_swaps,
                   self.grid_lag_swaps, 100 * self.min_detected_margin, 100 * self.min_undetected_margin, self.max_rel_err))
        print(msg)
        return msg


def compare_records(gpu, orc, ogrid, dop_lo, thr, ggrid=None, max_ties=0, margins=None):
    """gpu/orc: record arrays [n]; ogrid: oracle cells [n, n_dop].  Returns the number of tie-zone swaps; fills
    `margins` (a Margins) when given."""
    gpu = np.asarray(gpu).reshape(-1)
    orc = np.asarray(orc).reshape(-1)
    assert gpu.shape == orc.shape
    assert np.array_equal(gpu["sat"], orc["sat"])
    ties = 0
    m = margins if margins is not None else Margins()
    for i in range(len(orc)):
        g, o = gpu[i], orc[i]
        same = (g["dop"] == o["dop"]) and (g["lag"] == o["lag"])
        if not same:
            # tie zone: the oracle's value at the engine's choice must match the oracle's best
            cell = ogrid[i, g["dop"] - dop_lo]
            near = abs(cell["snr"] / o["snr"] - 1) < TIE
            if g["dop"] == o["dop"]:
                near = near or abs(g["peak"] / o["peak"] - 1) < TIE
            assert near, "sat %d: engine (dop %d, lag %d, snr %g) vs oracle (dop %d, lag %d, snr %g)" % (
                o["sat"], g["dop"], g["lag"], g["snr"], o["dop"], o["lag"], o["snr"])
            # a detected satellite must never be decided differently
            assert o["snr"] < thr * (1 - RTOL), "decision flip on a detected satellite %d" % o["sat"]
            ties += 1
        for f in ("peak", "noise", "snr"):
            assert abs(g[f] / o[f] - 1) < RTOL, "sat %d %s: %g vs %g" % (o["sat"], f, g[f], o[f])
            m.max_rel_err = max(m.max_rel_err, float(abs(g[f] / o[f] - 1)))
        # detection decision, outside a +-RTOL band around the threshold
        m.records += 1
        if abs(o["snr"] / thr - 1) > RTOL:
            assert (g["snr"] >= thr) == (o["snr"] >= thr), "detected-set mismatch on sat %d" % o["sat"]
        else:
            m.in_threshold_band += 1
        if o["snr"] >= thr:
            m.detected += 1
            m.min_detected_margin = min(m.min_detected_margin, float(o["snr"] / thr - 1))
        else:
            m.min_undetected_margin = min(m.min_undetected_margin, float(1 - o["snr"] / thr))
    m.tie_swaps += ties
    assert ties <= max_ties, "%d tie-zone swaps (allowed %d)" % (ties, max_ties)
    if ggrid is not None:
        gg = np.asarray(ggrid).reshape(ogrid.shape)
        rel = np.abs(gg["snr"] / ogrid["snr"] - 1)
        assert rel.max() < RTOL, "grid snr rel err %g" % rel.max()
        lag_ne = gg["lag"] != ogrid["lag"]
        if lag_ne.any():
            # per-Doppler argmax may also swap inside the tie zone
            assert (np.abs(gg["peak"][lag_ne] / ogrid["peak"][lag_ne] - 1) < TIE).all()
            m.grid_lag_swaps += int(lag_ne.sum())
    return ties
