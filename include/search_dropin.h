/*
 * search_dropin.h -- host-side shim that keeps the reference's search scheduling (the body of
 * SearchTask, gps/search.cpp:524-603) and routes Sample()'s DSP and Correlate() to the GPU engine.
 *
 * The reference's six entry points (gps/gps.h:140-145) map onto this shim as follows; the ~40-line
 * adapter that defines those six symbols inside the reference tree is listed in INTEGRATION.md:
 *   SearchParams(argc, argv)  -> acq_dropin_params          (-gsig N, -gt; search.cpp:72-95)
 *   SearchInit()              -> acq_dropin_create          (spectra built on the GPU; search.cpp:183-346)
 *   SearchTask(param)         -> loop { acq_dropin_pass }   (search.cpp:512-604)
 *   SearchEnable(sat)         -> acq_dropin_enable          (search.cpp:504-506)
 *   SearchFree()              -> acq_dropin_destroy         (search.cpp:354-357)
 *   SearchTaskRun()           -> unchanged host logic (search.cpp:610-648); it only sleeps/wakes the task
 *
 * Everything the loop calls on the tracking / UI / hardware side is reached through acq_host_iface,
 * with the reference's own argument units (gps/channel.cpp:891-934, gps/stat.cpp:75-98).
 */
#ifndef ACQ_SEARCH_DROPIN_H
#define ACQ_SEARCH_DROPIN_H

#include "acq_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct acq_host_iface {
    void *user;
    /* int ChanReset(int sat, int codegen_init): free channel index or -1 (gps/channel.cpp:891-921) */
    int (*chan_reset)(void *user, int sat, int codegen_init);
    /* void ChanStart(int ch, int sat, int t_sample, int lo_shift, int ca_shift, int snr) (gps/channel.cpp:925-934)
     * lo_shift in Doppler bins, ca_shift in FS samples (= lag * DECIM), snr truncated to int. */
    void (*chan_start)(void *user, int ch, int sat, int t_sample, int lo_shift, int ca_shift, int snr);
    /* GPSstat(STAT_SAT, snr, ch, sat, snr < min_sig, us) (search.cpp:569,580) */
    void (*stat_sat)(void *user, double snr, int ch, int sat, int too_weak, int us);
    /* GPSstat(STAT_DOP, 0, ch, lo_shift*BIN_SIZE, ca_shift) -- the Hz value is narrowed to int by the
     * reference's prototype (search.cpp:595, gps/gps.h:298) */
    void (*stat_dop)(void *user, int ch, int lo_hz, int ca_shift);
    /* The SPI half of Sample() (search.cpp:395-406): trigger the FPGA sampler, wait 4004 us, read
     * 16 x 512 bytes into dst.  Returns 0 on success. */
    int (*capture)(void *user, uint8_t *dst);
    /* timer_us() */
    unsigned (*timer_us)(void *user);
    /* NextTask(where): cooperative yield; may be NULL */
    void (*yield)(void *user, const char *where);
} acq_host_iface;

typedef struct acq_dropin acq_dropin;

enum {
    /* one capture per satellite, ChanReset before sampling: the reference's exact sequence */
    ACQ_DROPIN_LITERAL = 0,
    /* one capture, every idle enabled satellite searched in ONE GPU call, then
     * ChanReset / ChanStart per detection (see INTEGRATION.md for the hardware caveat) */
    ACQ_DROPIN_BATCH = 1
};

int acq_dropin_create(acq_dropin **out, const acq_sat *sats, int n_sats, const acq_host_iface *host, int device);
int acq_dropin_destroy(acq_dropin *d);
/* SearchParams: parses "-gsig N" (minimum_sig) and "-gt" exactly like search.cpp:72-95. */
int acq_dropin_params(acq_dropin *d, int argc, char *argv[]);
/* gps.acq_Navstar / acq_QZSS / acq_Galileo (search.cpp:525,533-535) */
int acq_dropin_set_acq(acq_dropin *d, int navstar, int qzss, int galileo);
/* Per-satellite search mask for the adapter's debugging filters (gps_debug / gps_e1b_only, search.cpp:537-539):
 * mask[sat] == 0 skips the satellite exactly like the reference's `continue`.  NULL clears the mask. */
int acq_dropin_set_mask(acq_dropin *d, const uint8_t *mask, int n_sats);
/* minimum_sig as SearchParams left it (MIN_SIG or -gsig N): the value SearchTask reports with STAT_PARAMS (search.cpp:521) */
int acq_dropin_min_sig(const acq_dropin *d);
/* Extension, off by default: hand ChanStart the acq_refine values -- ca_shift at FS-sample
 * resolution instead of lag * DECIM, lo_shift = the bin nearest to the interpolated Doppler.  Units and call
 * order are unchanged, so gps/channel.cpp needs no edit. */
int acq_dropin_set_refine(acq_dropin *d, int on);
/* SearchEnable(sat): sat is no longer tracked, search it again (search.cpp:504-506) */
int acq_dropin_enable(acq_dropin *d, int sat);
int acq_dropin_is_busy(const acq_dropin *d, int sat);
/* One pass of `for (sp = Sats; sp->prn != -1; sp++)` (search.cpp:530-602).
 * Returns the number of satellites handed to ChanStart, or a negative acq_status. */
int acq_dropin_pass(acq_dropin *d, int mode);
/* The engine underneath (for acq_last_error-style diagnostics and direct searches). */
acq_engine *acq_dropin_engine(acq_dropin *d);

/* ---- capture sources: the data formats on the input side of the path ----------------
 * The wire format is the FPGA sampler's: 65536 one-bit samples per capture, LSB first (sample i = bit i&7 of
 * byte i>>3, gps/search.cpp:408-411), delivered as 16 SPI packets of GPS_SAMPS*2 = 512 bytes
 * (gps/search.cpp:389,399-406) or, with GPS_SAMPLES_FROM_FILE, read 512 bytes at a time from a raw file of
 * consecutive captures (gps/search.cpp:361-380,400-401). */

/* Concatenate SPI packets into one capture block: dst[k*packet_bytes ..] = packets[k][0 .. packet_bytes).
 * n_packets * packet_bytes must equal ACQ_BLOCK_BYTES (16 x 512 on the reference hardware). */
int acq_capture_from_packets(const uint8_t *const *packets, int n_packets, int packet_bytes, uint8_t *dst);

/* Raw capture file (the GPS_SAMPLES_FROM_FILE format: fs 16.368 MHz, IF 4.092 MHz, 1 bit per sample, packed
 * LSB first).  acq_capture_file_next reads the next n_blocks * 8192 bytes; it returns ACQ_OK, or
 * ACQ_CAPTURE_EOF when fewer bytes remain (the reference prints "end of GPS samples data file" and exits,
 * gps/search.cpp:375-378), or ACQ_ERR_ARG on I/O errors. */
#define ACQ_CAPTURE_EOF 1
typedef struct acq_capture_file acq_capture_file;
int acq_capture_file_open(acq_capture_file **out, const char *path);
int acq_capture_file_next(acq_capture_file *f, uint8_t *dst, int n_blocks);
long long acq_capture_file_remaining(const acq_capture_file *f);  /* whole capture blocks left */
int acq_capture_file_rewind(acq_capture_file *f);
int acq_capture_file_close(acq_capture_file *f);
/* acq_host_iface.capture adaptor: user = acq_capture_file*, one block per call, non-zero at end of file. */
int acq_capture_file_iface(void *user, uint8_t *dst);

#ifdef __cplusplus
}
#endif
#endif
