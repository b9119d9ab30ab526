"""GPU suite, boundary: the COMPILED reference-side adapter (integration/search_gpu.cpp, built against the reference's
gps/gps.h) driving libacq_b200.so.  SearchParams / SearchInit / SearchTask run inside the receiver harness
(integration/harness_gpu.cpp: SPI sampler serving the golden captures, ChanReset / ChanStart / GPSstat recorded), and the
event log is compared with the log of the UNMODIFIED reference's SearchTask (tests/golden/ref_search_task_events.npz,
written by tools/gen_golden.py from oracle/_ref)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from parity import RTOL

pytestmark = pytest.mark.gpu

EV_CHAN_RESET, EV_CHAN_START, EV_STAT_SAT, EV_STAT_DOP, EV_OTHER = 1, 2, 3, 4, 5
STAT_PARAMS, STAT_ACQUIRE = None, None  # positions checked by count only (enum values live in the reference's gps.h)


@pytest.fixture(scope="module")
def adp(gpu_required):
    from flydog_sdr_gps_b200 import adapter
    a = adapter.Adapter()
    assert a.init() == 0, "SearchParams/SearchInit through the adapter failed"
    yield a
    a.L.adp_free()


def _compare(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["kind"] == w["kind"]
        if w["kind"] == EV_CHAN_RESET:      # ChanReset(sat, codegen_init) -> ch
            assert (g["a"], g["b"], g["c"]) == (w["a"], w["b"], w["c"])
        elif w["kind"] == EV_CHAN_START:    # ChanStart(ch, sat, t_sample, lo_shift, ca_shift, (int) snr)
            assert (g["a"], g["b"], g["c"], g["d"]) == (w["a"], w["b"], w["c"], w["d"])
            assert abs(g["e"] - w["e"]) <= 1 and abs(g["e"] - w["e"]) <= RTOL * w["e"] + 1
        elif w["kind"] == EV_STAT_SAT:      # GPSstat(STAT_SAT, snr, ch, sat, snr < min_sig, us)
            assert (g["a"], g["b"]) == (w["a"], w["b"])
            if abs(w["x"] / 16.0 - 1) > RTOL:
                assert g["c"] == w["c"]
            assert abs(g["x"] - w["x"]) <= RTOL * max(w["x"], 1e-9)
        elif w["kind"] == EV_STAT_DOP:      # GPSstat(STAT_DOP, 0, ch, (int)(lo_shift*BIN_SIZE), ca_shift)
            assert (g["a"], g["b"], g["c"]) == (w["a"], w["b"], w["c"])


def test_adapter_search_task_event_log_equals_the_reference(adp, golden_search):
    want = np.load(os.path.join(GOLDEN, "ref_search_task_events.npz"))["events"]
    ev = adp.search_task(golden_search["captures"], passes=2, free_chans=12)
    # SearchTask's prologue: GPSstat(STAT_PARAMS, 0, DECIM, minimum_sig) then GPSstat(STAT_ACQUIRE, 0, 1) (search.cpp:521-522)
    other = ev[ev["kind"] == EV_OTHER]
    assert len(other) == 2
    assert (other[0]["b"], other[0]["c"]) == (4, 16) and other[1]["b"] == 1
    got = ev[ev["kind"] != EV_OTHER]
    _compare(got, want)
    assert int((got["kind"] == EV_CHAN_START).sum()) == 12
    # labels SearchInit gives the satellites (search.cpp:189-191)
    assert adp.L.adp_prn_label(0) == b"N01 " and adp.L.adp_prn_label(32) == b"Q194" and adp.L.adp_prn_label(36) == b"E02 "


def test_adapter_constellation_switches_and_debug_filters(adp, golden_search):
    caps = golden_search["captures"]
    only_gal = adp.search_task(caps[:1], passes=1, free_chans=12, acq=(0, 0, 1))
    sats = {int(e["a"]) for e in only_gal if e["kind"] == EV_CHAN_RESET}
    assert sats and min(sats) >= 36                       # gps.acq_Navstar / acq_QZSS off (search.cpp:533-535)
    dbg = adp.search_task(caps[:1], passes=1, free_chans=12, debug_prn=11)
    assert {int(e["a"]) for e in dbg if e["kind"] == EV_CHAN_RESET} == {10}   # gps_debug: PRN 11, never Galileo (:537-538)
    e1b = adp.search_task(caps[:1], passes=1, free_chans=12, e1b_only=1)
    assert min(int(e["a"]) for e in e1b if e["kind"] == EV_CHAN_RESET) >= 36  # gps_e1b_only (:539)


def test_adapter_search_task_run_policy(adp, golden_search):
    """SearchTaskRun (search.cpp:610-648): sleeps the search task when there are users and enough good satellites,
    wakes it otherwise; never runs during an update/lock."""
    adp.search_task(golden_search["captures"][:1], passes=1)   # SearchTask has run: searchTaskID is set
    s0, w0 = adp.task_run(good=9, users=2, clk_corrections=1)
    assert (s0, w0) == (1, 0)                                   # busy receiver with a fix: stop acquiring
    assert adp.task_run(good=9, users=2, clk_corrections=1) == (1, 0)   # no change, no call
    assert adp.task_run(good=3, users=2, clk_corrections=1) == (1, 1)   # fewer than five good satellites: wake up
    assert adp.task_run(good=3, users=2, clk_corrections=1, locked=1) == (2, 1)
    assert adp.task_run(good=9, users=0, clk_corrections=1) == (2, 2)   # nobody connected: might as well search
    assert adp.task_run(good=9, users=2, clk_corrections=1, always_acq=1) == (2, 2)
