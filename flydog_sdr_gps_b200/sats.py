"""Satellite table of the receiver: the (prn, T1, T2, type) rows the host hands to acq_create.

This is the same information as the reference's Sats[] (gps/sats.cpp:25-142): Navstar rows carry
the IS-GPS-200 G2 tap pair, QZSS rows the G2 delay (documentation only) and the G2 preset (octal in
IS-QZSS), Galileo rows just the PRN of the E1-B memory code.  `sat` everywhere is the row index.
A receiver integrating the engine passes its own table; this module provides the reference's
defaults plus the extended tables the BASELINE.json configs ask for.
"""
NAVSTAR, SBAS, QZSS, E1B = 0, 1, 2, 3

# IS-GPS-200 Table 3-Ia: G2 phase-select taps for PRN 1..32
_L1_TAPS = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6),
            (6, 7), (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6),
            (5, 7), (6, 8), (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9)]

# IS-QZSS L1 C/A: (prn, G2 delay, G2 preset) for the satellites the reference searches
_QZSS = [(194, 208, 0o1607), (195, 711, 0o1747), (196, 189, 0o1305), (199, 663, 0o727)]

# Galileo PRNs in active service in the reference's table (gps/sats.cpp:104-139)
_E1B_ACTIVE = [2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 15, 19, 21, 24, 25, 26, 27, 30, 31, 33, 34, 36]


def navstar():
    return [(prn, t1, t2, NAVSTAR) for prn, (t1, t2) in enumerate(_L1_TAPS, start=1)]


def qzss():
    return [(prn, d, init, QZSS) for prn, d, init in _QZSS]


def e1b(prns=None):
    return [(p, 0, 0, E1B) for p in (prns if prns is not None else _E1B_ACTIVE)]


def reference_table():
    """The 59 rows of the reference's Sats[]: 32 Navstar, 4 QZSS, 23 E1B (sat index = row)."""
    return navstar() + qzss() + e1b()


def all_constellation_table():
    """32 Navstar + all 50 E1B memory codes (BASELINE config 4: 82 PRNs)."""
    return navstar() + e1b(range(1, 51))


def label(row):
    """The reference's prn_s label (gps/search.cpp:189-191)."""
    prn, _, _, typ = row
    return {NAVSTAR: "N%02d", SBAS: "S%d", QZSS: "Q%d", E1B: "E%02d"}[typ] % prn
