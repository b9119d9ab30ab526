// search_gpu.cpp -- what gps/search.cpp becomes in a receiver that runs its acquisition search on a B200.
//
// This translation unit REPLACES the reference's gps/search.cpp: it defines the same six entry points with the
// reference's prototypes (gps/gps.h:140-145) --
//     SearchInit  SearchFree  SearchTask  SearchTaskRun  SearchEnable  SearchParams
// -- keeps the host-side scheduling (one capture per satellite, ChanReset before sampling, the 20 s start-up
// sleep, the load-based sleep/wake policy) and routes everything that was DSP -- the code-spectrum build of
// SearchInit (search.cpp:243-346), Sample()'s decimation and FFT (search.cpp:408-447) and Correlate()
// (search.cpp:453-499) -- to libacq_b200.so through include/search_dropin.h.  FFTW is no longer linked.
//
// It is compiled against the reference's own gps/gps.h (include path: the reference tree), so a drift of any
// prototype it uses (ChanReset, ChanStart, GPSstat with its defaulted arguments, the SATELLITE layout) is a build
// error.  integration/Makefile builds it here behind the stub runtime headers of oracle/ref_harness/stubs together with
// a receiver-side harness (integration/harness_gpu.cpp: SPI, scheduler, tracking callees); tests/ replay the
// reference's SearchTask event log through it on the GPU box.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "types.h"
#include "kiwi.h"
#include "clk.h"
#include "cfg.h"
#include "rx_util.h"
#include "gps.h"
#include "spi.h"
#include "spi_dev.h"
#include "misc.h"
#include "coroutines.h"

#include "search_dropin.h"  // <repo>/include

static acq_dropin *dropin;
static int searchTaskID = -1;
static int g_argc;
static char **g_argv;

// ---- the receiver-side callees, with the reference's argument units (gps/channel.cpp:891-934, gps/stat.cpp) ----
static int h_chan_reset(void *, int sat, int codegen_init) { return ChanReset(sat, codegen_init); }
static void h_chan_start(void *, int ch, int sat, int t_sample, int lo_shift, int ca_shift, int snr)
{
    ChanStart(ch, sat, t_sample, lo_shift, ca_shift, snr);
}
static void h_stat_sat(void *, double snr, int ch, int sat, int too_weak, int us) { GPSstat(STAT_SAT, snr, ch, sat, too_weak, us); }
static void h_stat_dop(void *, int ch, int lo_hz, int ca_shift) { GPSstat(STAT_DOP, 0, ch, lo_hz, ca_shift); }
static unsigned h_timer_us(void *) { return timer_us(); }
static void h_yield(void *, const char *where) { NextTask(where); }

// The SPI half of Sample() (search.cpp:388-406): trigger the sampler, let it fill (one FFT length of signal), read the
// packets.  Everything Sample() did with the bits afterwards now happens on the GPU.
static int h_capture(void *, uint8_t *dst)
{
    const int fill_us = (int)(0.5 + 1000000.0 / BIN_SIZE);
    const int packet = GPS_SAMPS * 2;
    SPI_MISO *rx = &SPI_SHMEM->gps_search_miso;
    spi_set(CmdSample);
    TaskSleepUsec(fill_us);
    for (int off = 0; off < NSAMPLES / 8; off += packet) {
        spi_get(CmdGetGPSSamples, rx, packet);
        memcpy(dst + off, rx->byte, packet);
    }
    return 0;
}

void SearchParams(int argc, char *argv[])
{
    // -gsig N / -gt are parsed by the shim once it exists (SearchInit); gps_main calls SearchParams first (gps/gps.cpp:49)
    g_argc = argc;
    g_argv = argv;
    for (int i = 1; i < argc; i++) {
        const char *v = argv[i];
        if (!strcmp(v, "?") || !strcmp(v, "-?") || !strcmp(v, "--?") || !strcmp(v, "-h") || !strcmp(v, "h") ||
            !strcmp(v, "-help") || !strcmp(v, "--h") || !strcmp(v, "--help")) {
            printf("GPS args:\n\t-gsig signal_threshold\n\t-gt test mode\n");
            kiwi_exit(0);
        }
    }
}

void SearchInit()
{
    static acq_sat table[MAX_SATS];
    int n = 0;
    for (SATELLITE *sp = Sats; sp->prn != -1; sp++, n++) {
        if (n >= MAX_SATS) {
            printf("MAX_SATS=%d not big enough\n", MAX_SATS);
            kiwi_exit(-1);
        }
        sp->sat = n;  // table index = the `sat` of every API on both sides
        const char *fmt = "N%02d ";
        if (sp->type == QZSS) fmt = "Q%d", gps.n_QZSS++;
        else if (sp->type == E1B) fmt = "E%02d ", gps.n_E1B++;
        else gps.n_Navstar++;
        if (asprintf(&sp->prn_s, fmt, sp->prn) < 0) sp->prn_s = NULL;
        table[n].prn = sp->prn;
        table[n].t1 = sp->T1;  // G2_delay for QZSS rows: same storage (gps/gps.h:103-110)
        table[n].t2 = sp->T2;  // G2_init
        table[n].type = (int)sp->type;
    }
    GPSstat_init();
    const acq_host_iface host = {NULL, h_chan_reset, h_chan_start, h_stat_sat, h_stat_dop, h_capture, h_timer_us, h_yield};
    if (acq_dropin_create(&dropin, table, n, &host, /* CUDA device */ 0) != ACQ_OK) {
        printf("GPS: acquisition engine: %s\n", acq_last_error());
        kiwi_exit(-1);
    }
    acq_dropin_params(dropin, g_argc, g_argv);
    CreateTaskF(SearchTask, 0, GPS_ACQ_PRIORITY, CTF_NO_PRIO_INV);
}

void SearchFree()
{
    acq_dropin_destroy(dropin);
    dropin = NULL;
}

void SearchEnable(int sat) { acq_dropin_enable(dropin, sat); }

void SearchTask(void *param)
{
    (void)param;
    TaskSleepSec(20);
    searchTaskID = TaskID();
    GPSstat(STAT_PARAMS, 0, DECIM, acq_dropin_min_sig(dropin));
    GPSstat(STAT_ACQUIRE, 0, 1);
    static uint8_t mask[MAX_SATS];
    for (;;) {
        if (!gps.acq_Navstar && !gps.acq_QZSS && !gps.acq_Galileo) {
            TaskSleepSec(1);  // wait for the UI to enable a constellation
            continue;
        }
        acq_dropin_set_acq(dropin, gps.acq_Navstar, gps.acq_QZSS, gps.acq_Galileo);
        int n = 0;
        for (SATELLITE *sp = Sats; sp->prn != -1; sp++, n++) {  // the debugging filters of search.cpp:537-539
            bool on = true;
            if (gps_debug > 0 && sp->prn != gps_debug) on = false;
            if (gps_debug && sp->type == E1B) on = false;
            if (gps_e1b_only && sp->type != E1B) on = false;
            mask[n] = on;
        }
        acq_dropin_set_mask(dropin, mask, n);
        if (acq_dropin_pass(dropin, ACQ_DROPIN_LITERAL) < 0) {  // one walk over Sats[], search.cpp:530-602
            printf("GPS: acquisition engine: %s\n", acq_last_error());
            TaskSleepSec(1);
        }
    }
}

// Load-based run/sleep policy of the search task (the acquisition used to be the heaviest load on the host CPU).  No DSP
// here; same decisions as before: run while the clock has not been corrected yet, while nobody is connected, while
// fewer than five satellites are good, or when the admin asks for it -- never during an update, SD copy, backup or
// while locked.
void SearchTaskRun()
{
    static int acquiring = 1;
    if (searchTaskID == -1) return;
    const bool blocked = update_in_progress || sd_copy_in_progress || backup_in_progress || is_locked;
    const bool wanted = clk.adc_gps_clk_corrections == 0 || rx_count_server_conns(EXTERNAL_ONLY) == 0 || gps.good < 5 ||
                        admcfg_bool("always_acq_gps", NULL, CFG_REQUIRED);
    const int run = (wanted && !blocked) ? 1 : 0;
    if (run == acquiring) return;
    acquiring = run;
    GPSstat(STAT_ACQUIRE, 0, acquiring);
    if (run) TaskWakeup(searchTaskID);
    else TaskSleepID(searchTaskID, 0);
}
