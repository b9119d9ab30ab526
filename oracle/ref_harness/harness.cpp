/*
 * harness.cpp -- builds the UNMODIFIED reference search (gps/search.cpp) into a test oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repository: the
 * reference translation unit is pulled in below through the include path (-I$(REF) of
 * oracle/Makefile, target `ref`: the same tree sats.cpp and simd.cpp are taken from), at build
 * time, in the development container; outputs go to oracle/_ref/, which is git-ignored.  Including the .cpp (rather than linking it) makes the file-static
 * workers Sample() / Correlate() and the buffers fwd_buf / code[] visible to this harness
 * (reference gps/search.cpp:54,57,382,453).
 *
 * What is stubbed: the cooperative scheduler (no-ops), SPI (serves capture bytes from memory
 * in 512-byte packets exactly as search.cpp:398-406 consumes them), the tracking-side
 * callees ChanReset / ChanStart / GPSstat (recorded into an event log), and FFTW (oracle FFT,
 * see stubs/fftw3.h).
 */
#include "gps/search.cpp"

#include "../orc_fft.h"

#include <vector>

/* ---------------------------------------------------------------- globals the reference expects */
gps_t gps;
int gps_chans = GPS_MAX_CHANS, gps_debug = 0, gps_e1b_only = 0;
bool update_in_progress, sd_copy_in_progress, backup_in_progress, is_locked;
ref_clk_t clk;
static ref_spi_shmem_t g_shmem;
ref_spi_shmem_t *SPI_SHMEM = &g_shmem;

struct ref_stop_exception { int code; };

void kiwi_exit(int err) { throw ref_stop_exception{err}; }
static int g_idle_yields;
int rx_count_server_conns(int) { return 0; }
bool admcfg_bool(const char *, bool *, int) { return true; }

static unsigned g_fake_time_us;
unsigned timer_us(void) { return g_fake_time_us += 1000; }

void NextTask(const char *where)
{
    /* SearchTask spins on NextTask("busy1") when every sat is busy (search.cpp:551-554):
     * stop the literal loop once a whole table's worth of sats has been skipped in a row */
    if (where && strcmp(where, "busy1") == 0 && ++g_idle_yields > 4 * MAX_SATS) throw ref_stop_exception{1};
}
void NextTaskP(const char *, int) {}
void TaskSleepUsec(int) {}
void TaskSleepSec(int) {}
int TaskID(void) { return 7; }
void TaskSleepID(int, int) {}
void TaskWakeup(int) {}
int CreateTaskF(ref_task_fn, void *, int, int) { return 7; }
void GPSstat_init() {}

#ifdef REF_SATS_E1B50
/* Second build of this harness (oracle/_ref/libref_search_e1b50.so): the UNMODIFIED search.cpp over a satellite table
 * holding all 50 Galileo E1-B memory codes (the reference's own Sats[] activates 23 of them, gps/sats.cpp:104-139), in
 * place of the reference's gps/sats.cpp.  Gives reference-made goldens for every PRN of BASELINE configs[2]/[3]. */
#define E1ROW(p) {p, 0, 0, E1B}
SATELLITE Sats[] = {
    E1ROW(1),  E1ROW(2),  E1ROW(3),  E1ROW(4),  E1ROW(5),  E1ROW(6),  E1ROW(7),  E1ROW(8),  E1ROW(9),  E1ROW(10),
    E1ROW(11), E1ROW(12), E1ROW(13), E1ROW(14), E1ROW(15), E1ROW(16), E1ROW(17), E1ROW(18), E1ROW(19), E1ROW(20),
    E1ROW(21), E1ROW(22), E1ROW(23), E1ROW(24), E1ROW(25), E1ROW(26), E1ROW(27), E1ROW(28), E1ROW(29), E1ROW(30),
    E1ROW(31), E1ROW(32), E1ROW(33), E1ROW(34), E1ROW(35), E1ROW(36), E1ROW(37), E1ROW(38), E1ROW(39), E1ROW(40),
    E1ROW(41), E1ROW(42), E1ROW(43), E1ROW(44), E1ROW(45), E1ROW(46), E1ROW(47), E1ROW(48), E1ROW(49), E1ROW(50),
    {-1}
};
#endif

/* ---------------------------------------------------------------- capture source (SPI stub) */
static const uint8_t *g_capture;        /* current 8192-byte block */
static size_t g_capture_pos;
static const uint8_t *g_capture_list;   /* optional list: Sample() call k reads block k % n */
static int g_capture_list_n, g_sample_calls;

void spi_set(SPI_CMD cmd, int, int)
{
    if (cmd == CmdSample) { /* search.cpp:395 trigger sampler */
        if (g_capture_list) g_capture = g_capture_list + (size_t)(g_sample_calls % g_capture_list_n) * 8192;
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

/* ---------------------------------------------------------------- FFTW shim with an input tap */
struct ref_fftwf_plan_s {
    orc_fft_plan *plan;
    float *buf;
    float *scratch;
    int n, sign;
};

static std::vector<float> g_init_inputs;   /* forward-FFT inputs seen during SearchInit (code replicas) */
static std::vector<float> g_sample_inputs; /* forward-FFT input of the last recorded Sample() */
static std::vector<float> *g_record_fwd;   /* where fftwf_execute taps forward inputs, or NULL */

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
    assert(in == out);
    ref_fftwf_plan_s *p = new ref_fftwf_plan_s;
    p->plan = orc_fft_plan_create(n, sign);
    p->buf = (float *)in;
    p->scratch = (float *)malloc(sizeof(float) * 2 * n);
    p->n = n;
    p->sign = sign;
    return p;
}

extern "C" void fftwf_execute(const fftwf_plan p)
{
    if (g_record_fwd && p->sign == FFTW_FORWARD) g_record_fwd->insert(g_record_fwd->end(), p->buf, p->buf + 2 * p->n);
    orc_fft_execute(p->plan, p->buf, p->scratch);
}

extern "C" void fftwf_destroy_plan(fftwf_plan p)
{
    orc_fft_plan_destroy(p->plan);
    free(p->scratch);
    delete p;
}

/* ---------------------------------------------------------------- tracking-side callees: event log */
enum { EV_CHAN_RESET = 1, EV_CHAN_START = 2, EV_STAT_SAT = 3, EV_STAT_DOP = 4, EV_STAT_OTHER = 5 };
struct ref_event {
    int32_t kind, a, b, c, d, e;
    double x, y;
};
static std::vector<ref_event> g_events;
static int g_free_chans, g_next_chan, g_last_reset_sat, g_pass_limit, g_passes;

int ChanReset(int sat, int codegen_init)
{
    /* Sats[] is walked in ascending order, so a non-increasing sat index marks a new pass */
    if (sat <= g_last_reset_sat && ++g_passes >= g_pass_limit) throw ref_stop_exception{0};
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

/* ---------------------------------------------------------------- exported C API */
extern "C" {

static bool g_inited;

int ref_init(void)
{
    if (g_inited) return 0;
    g_init_inputs.clear();
    g_record_fwd = &g_init_inputs;
    FILE *saved = stdout;
    stdout = fopen("/dev/null", "w"); /* SearchInit printf()s */
    SearchInit();                     /* search.cpp:183-350 */
    fclose(stdout);
    stdout = saved;
    g_record_fwd = NULL;
    g_inited = true;
    return 0;
}

int ref_n_sats(void)
{
    int n = 0;
    for (SATELLITE *sp = Sats; sp->prn != -1; sp++) n++;
    return n;
}

void ref_sat(int i, int32_t *prn, int32_t *t1, int32_t *t2, int32_t *type)
{
    *prn = Sats[i].prn;
    *t1 = Sats[i].T1;
    *t2 = Sats[i].T2;
    *type = (int32_t)Sats[i].type;
}

/* chips of Galileo E1-B PRN `prn` (1..50) as the reference's own E1BCODE class expands them from its hex strings
 * (gps/e1bcode.h:63-92): 4092 values 0/1 */
void ref_e1b_chips(int prn, uint8_t *out)
{
    E1BCODE c(prn);
    for (int i = 0; i < E1B_CODELEN; i++) {
        out[i] = (uint8_t)c.Chip();
        c.Clock();
    }
}

/* chips of the C/A code with G2 taps (t1, t2) -- or the G2 preset in t2 when t1 > 10 -- from the reference's CACODE
 * (gps/cacode.h:23-64): 1023 values 0/1 */
void ref_ca_chips(int t1, int t2, uint8_t *out)
{
    CACODE c(t1, t2);
    for (int i = 0; i < L1_CODELEN; i++) {
        out[i] = (uint8_t)c.Chip();
        c.Clock();
    }
}

/* first copy of the code spectrum (search.cpp:283) */
void ref_code_spectrum(int sat, float *out) { memcpy(out, code[sat], sizeof(float) * 2 * FFT_LEN); }

/* input of the forward FFT at search.cpp:280/342 for this sat.  SearchInit transforms all
 * Navstar+QZSS sats first (table order), then all E1B sats (table order). */
int ref_code_baseband(int sat, float *out)
{
    int order = 0, idx = -1;
    for (int pass = 0; pass < 2 && idx < 0; pass++)
        for (SATELLITE *sp = Sats; sp->prn != -1; sp++) {
            const bool e1b = (sp->type == E1B);
            if ((pass == 1) != e1b) continue;
            if (sp->sat == sat) { idx = order; break; }
            order++;
        }
    if (idx < 0 || (size_t)(idx + 1) * 2 * FFT_LEN > g_init_inputs.size()) return -1;
    memcpy(out, &g_init_inputs[(size_t)idx * 2 * FFT_LEN], sizeof(float) * 2 * FFT_LEN);
    return 0;
}

/* Sample() on one 8192-byte block (search.cpp:382-449). x2 (optional) = FFT input, D (optional) = fwd_buf after FFT. */
void ref_sample(const uint8_t *packed, float *x2, float *D)
{
    g_capture_list = NULL;
    g_capture = packed;
    g_sample_inputs.clear();
    g_record_fwd = x2 ? &g_sample_inputs : NULL;
    Sample();
    g_record_fwd = NULL;
    if (x2) memcpy(x2, g_sample_inputs.data(), sizeof(float) * 2 * FFT_LEN);
    if (D) memcpy(D, fwd_buf, sizeof(float) * 2 * FFT_LEN);
}

/* Correlate() against the spectrum left in fwd_buf by the last ref_sample (search.cpp:453-499,574). */
float ref_correlate(int sat, int32_t *dop, int32_t *lag)
{
    int d = 0, l = 0;
    float snr = Correlate(sat, fwd_buf, &d, &l);
    *dop = d;
    *lag = l;
    return snr;
}

/* Sample + Correlate for a list of sats on ONE capture block: what SearchTask does per sat
 * (search.cpp:572-575), except that the same bytes are re-served for every sat. */
void ref_search(const uint8_t *packed, const int32_t *sats, int n, int32_t *dop, int32_t *lag, float *snr)
{
    for (int k = 0; k < n; k++) {
        ref_sample(packed, NULL, NULL);
        snr[k] = ref_correlate(sats[k], &dop[k], &lag[k]);
    }
}

/* Runs the literal SearchTask loop (search.cpp:512-604) for `passes` passes over Sats[].
 * Sample() call k reads capture block k % n_blocks.  free_chans = number of idle tracking
 * channels ChanReset may hand out.  Returns the number of events logged. */
int ref_search_task(const uint8_t *blocks, int n_blocks, int passes, int free_chans, int min_sig, int acq_navstar,
                    int acq_qzss, int acq_galileo)
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
    minimum_sig = min_sig;
    gps.acq_Navstar = acq_navstar;
    gps.acq_QZSS = acq_qzss;
    gps.acq_Galileo = acq_galileo;
    for (SATELLITE *sp = Sats; sp->prn != -1; sp++) sp->busy = false;
    try {
        SearchTask(NULL);
    } catch (ref_stop_exception &) {
    }
    g_capture_list = NULL;
    return (int)g_events.size();
}

void ref_get_events(ref_event *out, int n) { memcpy(out, g_events.data(), sizeof(ref_event) * (size_t)n); }

} /* extern "C" */
