"""The five BASELINE.json configurations as data: satellite table, search parameters, visible
signals.  Used by bench.py and the parity tests so that both exercise the same workloads
(SURVEY.md 8(d) gives the sizes: cells = sum_sats n_dop * lags, tiles = sats * n_dop * K)."""
import numpy as np

from . import sats as S
from .engine import BIN_HZ

# (name, description)
CONFIGS = {
    "cfg1": "GPS L1 C/A cold start: 32 PRNs, reference Doppler span -20..+20 bins, K=1, 8 visible sats",
    "cfg2": "GPS L1 C/A weak signal: 32 PRNs, +-10 kHz at half-bin spacing (161 bins), K=20 non-coherent blocks",
    "cfg3": "Galileo E1B: 50 PRNs, +-10 kHz full bins (81 bins), 16368 lags",
    "cfg4": "GPS+Galileo: 82 PRNs (32 C/A + 50 E1B), reference Doppler span, one capture",
    "cfg5": "receiver farm: 1024 independent captures x 32 GPS PRNs, reference Doppler span",
    # not a BASELINE configuration: the Galileo search of cfg3 with non-coherent sums
    "cfg3_k4": "Galileo E1B, 50 PRNs, 81 bins, K=4 non-coherent blocks (extension; one-CTA k_search_e1b_multi)",
}


# A noise-only capture (synth.make_capture(seed, 1, table("cfg3"), [])) whose strongest false E1B peak over 50 PRNs x 81
# bins x 16368 lags is 16.36: within 3 % of the reference's E1B threshold 16 (gps/search.cpp:549), which sits inside
# the noise distribution of so large a scan (noise-only captures reach 16..19).  tools/find_threshold_capture.py.
E1B_NEAR_THRESHOLD_SEED = 61133


def table(cfg):
    if cfg in ("cfg1", "cfg2", "cfg5"):
        return S.navstar()
    if cfg in ("cfg3", "cfg3_k4"):
        return S.e1b(range(1, 51))
    if cfg == "cfg4":
        return S.all_constellation_table()
    raise KeyError(cfg)


def params_kw(cfg):
    """Keyword overrides on the reference defaults."""
    if cfg == "cfg2":
        # thresholds: after 20 non-coherent sums the noise-only max/mean is ~2, so 16 is meaningless
        return dict(dop_lo=-80, dop_hi=80, half_bin=1, k_noncoh=20, thr_l1=2.6)
    if cfg == "cfg3":
        return dict(dop_lo=-40, dop_hi=40)
    if cfg == "cfg3_k4":
        return dict(dop_lo=-40, dop_hi=40, k_noncoh=4, thr_e1b=6.0)
    return {}


def n_captures(cfg):
    return 1024 if cfg == "cfg5" else 1


def signals(cfg, seed):
    """Visible signals for capture `seed`: list of (sat index, tau [FS samples], doppler [Hz], C/N0 [dB-Hz], phase)."""
    rng = np.random.default_rng(1000 + seed)
    if cfg in ("cfg1", "cfg5"):
        prns = rng.choice(32, 8, replace=False)
        cn0 = [50, 47, 45, 44, 43, 42, 41, 40]
        out = []
        for k, sat in enumerate(prns):
            # a mix of on-bin and off-bin Dopplers inside the +-4.9 kHz reference span
            f = float(rng.integers(-19, 20)) * BIN_HZ + (0.0 if k % 2 == 0 else float(rng.uniform(-60, 60)))
            out.append((int(sat), int(rng.integers(0, 16368)), f, cn0[k], float(rng.uniform(0, 2 * np.pi))))
        return out
    if cfg == "cfg2":
        prns = rng.choice(32, 8, replace=False)
        return [(int(sat), int(rng.integers(0, 16368)), float(rng.uniform(-9500, 9500)), float(rng.uniform(30, 35)),
                 float(rng.uniform(0, 2 * np.pi))) for sat in prns]
    if cfg in ("cfg3", "cfg3_k4"):
        prns = rng.choice(50, 6, replace=False)
        return [(int(sat), int(rng.integers(0, 65472)), float(rng.integers(-38, 39)) * BIN_HZ, float(rng.uniform(43, 48)),
                 float(rng.uniform(0, 2 * np.pi))) for sat in prns]
    if cfg == "cfg4":
        gps = rng.choice(32, 6, replace=False)
        gal = 32 + rng.choice(50, 4, replace=False)
        out = [(int(s), int(rng.integers(0, 16368)), float(rng.integers(-19, 20)) * BIN_HZ, float(rng.uniform(42, 50)),
                float(rng.uniform(0, 2 * np.pi))) for s in gps]
        out += [(int(s), int(rng.integers(0, 65472)), float(rng.integers(-19, 20)) * BIN_HZ, float(rng.uniform(44, 48)),
                 float(rng.uniform(0, 2 * np.pi))) for s in gal]
        return out
    raise KeyError(cfg)


def shard(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` of `world` (captures for cfg5, sat slots for cfg4)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
