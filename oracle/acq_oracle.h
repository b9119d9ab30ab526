/*
 * acq_oracle.h -- CPU oracle for the GNSS acquisition search (reference gps/search.cpp).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, link or call anything under oracle/.
 * The product (flydog_sdr_gps_b200/csrc) never includes or links it.
 *
 * PARITY STATUS: every stage up to the FFT input (code generators, sample mixing,
 * half-band decimators) is pinned bit-exactly against the reference's own source
 * compiled from /root/reference (oracle/_ref, see oracle/Makefile) and against the
 * reference's three code known-answer tests.  The FFT arithmetic itself lives in
 * FFTW3f, which is neither vendored nor pinned by the reference and is not
 * installable here: at that boundary parity is UNPINNED ("parity unpinned") and
 * is defined against (reference search.cpp + oracle FFT) -- see DESIGN.md.
 *
 * With default parameters (orc_params_default) this computes exactly what
 * Sample() + Correlate() compute (gps/search.cpp:382-499).  The extra parameters
 * (Doppler span, half-bin spacing, K non-coherent blocks) are the extensions the
 * BASELINE.json configs ask for; their definitions are in SURVEY.md section 8(d).
 */
#ifndef ACQ_ORACLE_H
#define ACQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FFT_LEN 16384   /* gps/gps.h:72  FFT_LEN  */
#define ORC_NSAMPLES 65536  /* gps/gps.h:73  NSAMPLES */
#define ORC_DECIM 4         /* gps/gps.h:62  DECIM    */
#define ORC_BLOCK_BYTES 8192 /* 16 SPI packets x 512 B, gps/search.cpp:389-406 */

/* sat_e values of the reference (gps/gps.h:98) */
enum { ORC_NAVSTAR = 0, ORC_SBAS = 1, ORC_QZSS = 2, ORC_E1B = 3 };

/* Mirrors the leading members of SATELLITE (gps/gps.h:101-112):
 * Navstar {prn, T1, T2}; QZSS {prn, G2_delay, G2_init}; E1B {prn, 0, 0}. */
typedef struct {
    int32_t prn, t1, t2, type;
} orc_sat;

typedef struct {
    int32_t dop_lo, dop_hi; /* inclusive Doppler range, in bins (half_bin=0) or half-bins (half_bin=1) */
    int32_t half_bin;       /* 0 = reference behaviour; 1 = extension, index h <-> (h/2) bins */
    int32_t k_noncoh;       /* 1 = reference behaviour; K>1 sums |r|^2 over K consecutive blocks */
    float thr_l1;           /* detection threshold for Navstar/QZSS (minimum_sig, search.cpp:70,549) */
    float thr_e1b;          /* detection threshold for E1B (16, search.cpp:549) */
    int32_t wrap_mode;      /* ORC_WRAP_REFERENCE (default) or ORC_WRAP_CIRCULAR, see below */
    int32_t sample_bits;    /* 1 = reference capture format (sign only, search.cpp:408-411).  2 = extension: each
                               65536-sample block is its sign plane (that format) followed by a magnitude plane of
                               the same layout; sample value (sign ? -1 : +1) * (mag ? 3 : 1) -- the MAX2769's
                               sign/magnitude levels (dev/gps_fe.cpp:104; the reference FPGA drops I_mag,
                               verilog/gps/gps.v:50).  0 is read as 1. */
    int32_t code_doppler;   /* 0 = off.  1 = extension: code-Doppler compensation
                               of the K-block sum.  A carrier offset f stretches the code by f/f_L1, so block b starts
                               65536 b f/f_L1 FS samples further along the code than 16 b /4-samples: with f = h BIN
                               (BIN = FS/65536, FS/4/f_L1 = 1/385 exactly) that is b h/385 /4-samples (h/2 for
                               half-bin indices).  The power of block b is then taken at lag
                               n + 16 b + s(b,h),  s = round-half-away(b h / (385 hb)), hb = 2 for half-bins else 1. */
} orc_params;

/* s(b, h) above; exact integer arithmetic so that every implementation picks the same lag */
static inline int orc_code_shift(int b, int h, int half_bin)
{
    const int d = 385 * (half_bin ? 2 : 1);
    const int a = b * h, m = a < 0 ? -a : a;
    const int s = (2 * m + d) / (2 * d);
    return a < 0 ? -s : s;
}

/* bytes of one capture block in the format `sample_bits` selects */
#define ORC_CAPTURE_BLOCK_BYTES(sample_bits) ((sample_bits) == 2 ? 2 * ORC_BLOCK_BYTES : ORC_BLOCK_BYTES)

/* How code-spectrum bins beyond the end of a satellite's row are fetched for NEGATIVE Doppler.
 *
 * The reference stores each code spectrum twice back to back (search.cpp:54,283-284) and calls
 *     simd_multiply_conjugate_ccc(FFT_LEN, data, code[sat]+FFT_LEN-dop, prod)      (search.cpp:471)
 * For dop < 0 the last |dop| elements read are code[sat][2*FFT_LEN .. 2*FFT_LEN+|dop|-1], i.e. they
 * run off the end of the row into the FIRST |dop| bins of the NEXT satellite's spectrum
 * (code[sat+1][0..|dop|-1]; all-zero rows after the last satellite, since code[] is a zeroed
 * static with MAX_SATS=64 rows).  The intended behaviour (the #else branch, search.cpp:473-477)
 * wraps to the satellite's own bins.  Because these bins sit next to DC, where the code spectrum
 * is strongest, the two differ by several per cent in snr, so parity with the compiled reference
 * needs the literal behaviour:
 *   ORC_WRAP_REFERENCE: bins k >= N-|dop| use the next table row (zeros after the last row).
 *   ORC_WRAP_CIRCULAR : C[(k-dop) mod N] of the same satellite. */
enum { ORC_WRAP_REFERENCE = 0, ORC_WRAP_CIRCULAR = 1 };

typedef struct {
    int32_t sat;  /* index into the caller's sat table */
    int32_t lag;  /* max_snr_i: code phase in /DECIM samples, [0, L) */
    int32_t dop;  /* max_snr_dop: bins (or half-bins) */
    float peak;   /* max_pwr at the chosen Doppler */
    float noise;  /* ave_pwr at the chosen Doppler */
    float snr;    /* peak / noise */
} orc_record;

/* per (sat, Doppler) entry of the optional full grid */
typedef struct {
    float peak;
    float noise;
    float snr;
    int32_t lag;
} orc_cell;

void orc_params_default(orc_params *p);

/* --- code generators (gps/cacode.h:23-64, gps/e1bcode.h:63-92) --- */
/* 1023 chips (0/1) of the C/A code given the SATELLITE T1/T2 pair. */
void orc_ca_chips(int t1, int t2, uint8_t *chips);
/* 4092 chips (0/1) of E1B PRN prn (1..50). */
void orc_e1b_chips(int prn, uint8_t *chips);

/* --- building blocks, exposed so each can be pinned on its own --- */
/* Half-band decimate-by-2 in place (search.cpp:140-166). buf holds size+31 complex floats. */
int orc_hb_decimate(int size, float *buf);
/* 65536 replica samples -> /4 -> 16384 complex (search.cpp:250-276, 315-338). out: 2*16384 floats. */
void orc_code_baseband(const orc_sat *sat, float *out);
/* code_baseband + forward FFT (search.cpp:280,342). out: 2*16384 floats. */
void orc_code_spectrum(const orc_sat *sat, float *out);
/* 8192 packed bytes -> mixed, /4 decimated baseband (search.cpp:398-442). out: 2*16384 floats.
 * half_rot=1 additionally multiplies sample n by exp(-j*pi*n/16384) (half-bin extension). */
void orc_capture_baseband(const uint8_t *packed, int half_rot, float *out);
/* Same for one block in the `sample_bits` format (1: identical to orc_capture_baseband). */
void orc_capture_baseband_sm(const uint8_t *block, int sample_bits, int half_rot, float *out);
/* capture_baseband + forward FFT (search.cpp:447). */
void orc_capture_spectrum(const uint8_t *packed, int half_rot, float *out);
/* forward / backward 16384-point FFT, in place (sign -1 / +1). */
void orc_fft16384(float *buf, int sign);

/* --- the search (Sample + Correlate for every sat, search.cpp:382-499,574) ---
 * packed: k_noncoh * 8192 bytes (consecutive 65536-sample blocks of one capture).
 * sats/n_sats: table; sel/n_sel: indices to search (sel==NULL -> all).
 * out: n_sel records. grid (optional, may be NULL): n_sel * n_dop cells.
 * nthreads: OpenMP threads over sats (<=0 -> default).
 * Returns 0, or <0 on bad arguments. */
int orc_search(const uint8_t *packed, const orc_sat *sats, int n_sats, const int32_t *sel, int n_sel,
               const orc_params *prm, orc_record *out, orc_cell *grid, int nthreads);

/* Same search with precomputed code spectra (n_sats * 2*16384 floats) -- lets the
 * caller amortise SearchInit's work (search.cpp:243-346) exactly like the reference. */
int orc_search_pre(const uint8_t *packed, const orc_sat *sats, int n_sats, const float *spectra,
                   const int32_t *sel, int n_sel, const orc_params *prm, orc_record *out,
                   orc_cell *grid, int nthreads);

/* --- acquisition refinement (extension; no reference counterpart) ---
 * For each record: the correlation r_d[n] = sum_k conj(D[k]) C[k-d] e^{+j 2 pi k n/N} (search.cpp:471-481 written as
 * a direct sum, double precision here) at (d-1,n) (d,n-1) (d,n) (d,n+1) (d+1,n), summed over the K blocks as
 *   num += Re[(Xm-Xp) conj(2X0-Xm-Xp)], den += |2X0-Xm-Xp|^2, X_d = r_d[n] e^{-j 2 pi d n/N}   -> delta = num/den
 *   e   += Re(r[n-1] conj r[n]),  l += Re(r[n+1] conj r[n])   -> eps = slope (l-e)/(l+e), slope 3 (C/A) | 1/3 (E1B)
 * both clamped to [-1,1].  Block b of a K-block capture is evaluated at lag n + 16 b (its code phase advance). */
typedef struct {
    float dop_hz;     /* (bin + delta) * BIN_SIZE */
    float code_fs;    /* DECIM * (lag + eps), FS samples */
    float peak;       /* sum_b |r_d[n]|^2 : the record's peak power, recomputed directly */
    int32_t ca_shift; /* round(code_fs) mod (L * DECIM) */
} orc_fine;

int orc_refine(const uint8_t *packed, const orc_sat *sats, int n_sats, const orc_params *prm,
               const orc_record *rec, int n_rec, orc_fine *out, int nthreads);

/* --- deterministic synthetic capture generator (SURVEY.md 8(d) "Value distributions") --- */
typedef struct {
    int32_t sat;       /* index into the sat table */
    int32_t tau;       /* code advance at sample 0, in FS samples */
    double doppler_hz; /* carrier offset from FC */
    double cn0_dbhz;
    double phase;      /* carrier phase at sample 0, radians */
    int32_t flip_ms;   /* >0: flip the sign every flip_ms milliseconds (data bits); 0: none */
    int32_t code_doppler; /* 1: the code runs at (1 + doppler_hz/f_L1) chips per nominal chip (as a real signal's does):
                             sample i carries chip floor(i (1 + f/f_L1)) + tau; 0: code Doppler ignored (v1 captures) */
} orc_signal;

/* Fills n_blocks*8192 bytes. */
int orc_gen_capture(uint64_t seed, int n_blocks, const orc_sat *sats, int n_sats, const orc_signal *sig,
                    int n_sig, uint8_t *packed);
/* Same signal and noise (the sign planes equal orc_gen_capture's output for the same seed) quantised to
 * `sample_bits`: with 2, block b is [sign plane 8192 B][magnitude plane 8192 B], mag = (|s| > mag_thr) with s in
 * units of the noise sigma (an AGC holding the magnitude duty cycle near 1/3 corresponds to mag_thr ~ 0.98).
 * Fills n_blocks * ORC_CAPTURE_BLOCK_BYTES(sample_bits) bytes. */
int orc_gen_capture_sm(uint64_t seed, int n_blocks, const orc_sat *sats, int n_sats, const orc_signal *sig,
                       int n_sig, int sample_bits, double mag_thr, uint8_t *packed);

#ifdef __cplusplus
}
#endif
#endif
