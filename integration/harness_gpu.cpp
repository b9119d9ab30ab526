/*
 * harness_gpu.cpp -- the receiver around integration/search_gpu.cpp, for the tests.
 *
 * TEST INFRASTRUCTURE.  It plays the parts of the reference that the acquisition search talks to: the SPI sampler
 * (serves capture bytes in 512-byte packets, exactly as the SPI half of Sample() consumes them, gps/search.cpp:398-406),
 * the cooperative scheduler (no-ops), and the tracking-side callees ChanReset / ChanStart / GPSstat, which it records
 * into an event log with the layout of oracle/ref_harness/harness.cpp -- so the log of the GPU adapter can be compared
 * entry by entry with the log of the unmodified reference (tests/golden/ref_search_task_events.npz).
 * Nothing here computes a search: the adapter calls libacq_b200.so.
 */
#include <string.h>

#include <vector>

#include "types.h"
#include "kiwi.h"
#include "gps.h"
#include "spi.h"
#include "spi_dev.h"
#include "misc.h"
#include "coroutines.h"

/* ---------------------------------------------------------------- globals the reference's gps code expects */
gps_t gps;
int gps_chans = GPS_MAX_CHANS, gps_debug = 0, gps_e1b_only = 0;
bool update_in_progress, sd_copy_in_progress, backup_in_progress, is_locked;
ref_clk_t clk;
static ref_spi_shmem_t g_shmem;
ref_spi_shmem_t *SPI_SHMEM = &g_shmem;

struct adp_stop { int code; };

void kiwi_exit(int err) { throw adp_stop{err}; }
static int g_users, g_always_acq, g_idle_yields;
int rx_count_server_conns(int) { return g_users; }
bool admcfg_bool(const char *, bool *, int) { return g_always_acq != 0; }
static unsigned g_fake_time_us;
unsigned timer_us(void) { return g_fake_time_us += 1000; }

void NextTask(const char *where)
{
    /* the literal loop spins on NextTask("busy1") once every satellite is busy (search.cpp:551-554) */
    if (where && strcmp(where, "busy1") == 0 && ++g_idle_yields > 4 * MAX_SATS) throw adp_stop{1};
}
void NextTaskP(const char *, int) {}
void TaskSleepUsec(int) {}
static int g_sleep_sec_calls;
void TaskSleepSec(int) { g_sleep_sec_calls++; }
int TaskID(void) { return 7; }
static int g_task_sleeps, g_task_wakeups, g_tasks_created;
void TaskSleepID(int, int) { g_task_sleeps++; }
void TaskWakeup(int) { g_task_wakeups++; }
int CreateTaskF(ref_task_fn, void *, int, int) { g_tasks_created++; return 7; }
void GPSstat_init() {}

/* ---------------------------------------------------------------- SPI sampler */
static const uint8_t *g_capture_list;
static const uint8_t *g_capture;
static size_t g_capture_pos;
static int g_capture_list_n, g_sample_calls;

void spi_set(SPI_CMD cmd, int, int)
{
    if (cmd == CmdSample) {
        g_capture = g_capture_list + (size_t)(g_sample_calls % g_capture_list_n) * 8192;
        g_capture_pos = 0;
        g_sample_calls++;
    }
}

void spi_get(SPI_CMD cmd, SPI_MISO *rx, int bytes, int, int)
{
    if (cmd != CmdGetGPSSamples) return;
    memcpy(rx->byte, g_capture + g_capture_pos, bytes);
    g_capture_pos += bytes;
}

/* ---------------------------------------------------------------- tracking-side callees: event log */
enum { EV_CHAN_RESET = 1, EV_CHAN_START = 2, EV_STAT_SAT = 3, EV_STAT_DOP = 4, EV_STAT_OTHER = 5 };
struct adp_event {
    int32_t kind, a, b, c, d, e;
    double x, y;
};
static std::vector<adp_event> g_events;
static int g_free_chans, g_next_chan, g_last_reset_sat, g_pass_limit, g_passes;

int ChanReset(int sat, int codegen_init)
{
    /* Sats[] is walked in ascending order, so a non-increasing sat index marks a new pass */
    if (sat <= g_last_reset_sat && ++g_passes >= g_pass_limit) throw adp_stop{0};
    g_last_reset_sat = sat;
    g_idle_yields = 0;
    int ch = (g_free_chans > 0) ? g_next_chan : -1;
    g_events.push_back({EV_CHAN_RESET, sat, codegen_init, ch, 0, 0, 0, 0});
    return ch;
}

void ChanStart(int ch, int sat, int t_sample, int lo_shift, int ca_shift, int snr)
{
    (void)t_sample;
    g_events.push_back({EV_CHAN_START, ch, sat, lo_shift, ca_shift, snr, 0, 0});
    g_free_chans--;
    g_next_chan++;
}

void GPSstat(STAT st, double p, int i, int j, int k, int m, double d)
{
    (void)d;
    if (st == STAT_SAT) g_events.push_back({EV_STAT_SAT, i, j, k, 0, 0, p, 0});
    else if (st == STAT_DOP) g_events.push_back({EV_STAT_DOP, i, j, k, 0, 0, p, 0});
    else g_events.push_back({EV_STAT_OTHER, (int)st, i, j, k, m, p, 0});
}

/* ---------------------------------------------------------------- exported driver */
extern "C" {

static bool g_inited;

/* SearchParams + SearchInit once per process (gps_main's order, gps/gps.cpp:49,54).  Returns 0, or -1 if the
 * adapter bailed out through kiwi_exit (e.g. no CUDA device). */
int adp_init(int argc, char **argv)
{
    if (g_inited) return 0;
    try {
        SearchParams(argc, argv);
        SearchInit();
    } catch (adp_stop &) {
        return -1;
    }
    g_inited = true;
    return g_tasks_created == 1 ? 0 : -2;
}

int adp_n_sats(void)
{
    int n = 0;
    for (SATELLITE *sp = Sats; sp->prn != -1; sp++) n++;
    return n;
}

const char *adp_prn_label(int sat) { return Sats[sat].prn_s; }

/* Runs the adapter's SearchTask for `passes` walks over Sats[].  Sample() call k reads capture block k % n_blocks;
 * free_chans idle tracking channels.  Returns the number of events logged. */
int adp_search_task(const uint8_t *blocks, int n_blocks, int passes, int free_chans, int acq_navstar, int acq_qzss,
                    int acq_galileo, int debug_prn, int e1b_only)
{
    g_events.clear();
    g_capture_list = blocks;
    g_capture_list_n = n_blocks;
    g_sample_calls = 0;
    g_free_chans = free_chans;
    g_next_chan = 0;
    g_last_reset_sat = -1;
    g_passes = 0;
    g_idle_yields = 0;
    g_pass_limit = passes;
    g_fake_time_us = 0;
    gps.acq_Navstar = acq_navstar;
    gps.acq_QZSS = acq_qzss;
    gps.acq_Galileo = acq_galileo;
    gps_debug = debug_prn;
    gps_e1b_only = e1b_only;
    for (int s = 0; s < adp_n_sats(); s++) SearchEnable(s);
    try {
        SearchTask(NULL);
    } catch (adp_stop &) {
    }
    return (int)g_events.size();
}

void adp_get_events(adp_event *out, int n) { memcpy(out, g_events.data(), sizeof(adp_event) * (size_t)n); }

/* SearchTaskRun policy probe: sets the receiver state, calls SearchTaskRun, returns (sleeps << 8) | wakeups so far. */
int adp_task_run(int good_sats, int users, int clk_corrections, int always_acq, int locked)
{
    gps.good = good_sats;
    g_users = users;
    clk.adc_gps_clk_corrections = clk_corrections;
    g_always_acq = always_acq;
    is_locked = locked != 0;
    SearchTaskRun();
    return (g_task_sleeps << 8) | g_task_wakeups;
}

void adp_free(void)
{
    if (g_inited) SearchFree();
    g_inited = false;
}

} /* extern "C" */
