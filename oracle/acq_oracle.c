/*
 * acq_oracle.c -- CPU restatement of the reference acquisition search.
 *
 * TEST INFRASTRUCTURE ONLY (see acq_oracle.h).  Plain C, strict IEEE fp32
 * (build with -O2 -ffp-contract=off, never -ffast-math) so that every stage before
 * the FFT reproduces the reference's x86 -O2 build bit for bit.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "acq_oracle.h"
#include "orc_fft.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int orc_threads(int n)
{
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define N ORC_FFT_LEN
#define NS ORC_NSAMPLES
#define NTAPS 31

/* ------------------------------------------------------------------------------------------
 * Galileo E1-B memory codes, packed (tools/gen_e1b_table.py; data from gps/e1bcode.h:10-61).
 * ---------------------------------------------------------------------------------------- */
static const uint32_t E1B_WORDS[50 * 128] = {
#include "../flydog_sdr_gps_b200/csrc/e1b_codes.inc"
};

void orc_params_default(orc_params *p)
{
    /* gps/search.cpp:465: for (dop = -5000/BIN_SIZE; dop <= 5000/BIN_SIZE; dop++) -> -20..+20 */
    p->dop_lo = -20;
    p->dop_hi = 20;
    p->half_bin = 0;
    p->k_noncoh = 1;
    p->thr_l1 = 16.0f;  /* MIN_SIG, gps/gps.h:60, search.cpp:70 */
    p->thr_e1b = 16.0f; /* search.cpp:549 */
    p->wrap_mode = ORC_WRAP_REFERENCE;
    p->sample_bits = 1; /* I_sign only, search.cpp:408-411 */
    p->code_doppler = 0;
}

/* ------------------------------------------------------------------------------------------
 * C/A code: gps/cacode.h:23-64.  G1 and G2 are 10-stage shift registers, stage i in bit i.
 *   feedback  G1: s3 ^ s10          G2: s2 ^ s3 ^ s6 ^ s8 ^ s9 ^ s10        (cacode.h:48-49)
 *   chip      tap mode   : G1.s10 ^ G2.s[t1] ^ G2.s[t2]                     (cacode.h:44)
 *             G2-init mode (t1>10 or t2>10): G1.s10 ^ G2.s10, G2 preset so that bit (i-1)
 *             of t2 lands in stage i                                        (cacode.h:27-32,44)
 * ---------------------------------------------------------------------------------------- */
void orc_ca_chips(int t1, int t2, uint8_t *chips)
{
    const int init_mode = (t1 > 10 || t2 > 10);
    unsigned g1 = 0x7fe; /* stages 1..10 all ones */
    unsigned g2 = 0x7fe;
    if (init_mode) {
        g2 = 0;
        for (int i = 1; i <= 10; i++)
            if ((t2 >> (i - 1)) & 1) g2 |= 1u << i;
    }
#define ST(r, i) (((r) >> (i)) & 1u)
    for (int n = 0; n < 1023; n++) {
        unsigned c = init_mode ? (ST(g1, 10) ^ ST(g2, 10)) : (ST(g1, 10) ^ ST(g2, t1) ^ ST(g2, t2));
        chips[n] = (uint8_t)c;
        unsigned f1 = ST(g1, 3) ^ ST(g1, 10);
        unsigned f2 = ST(g2, 2) ^ ST(g2, 3) ^ ST(g2, 6) ^ ST(g2, 8) ^ ST(g2, 9) ^ ST(g2, 10);
        g1 = ((g1 << 1) & 0x7fc) | (f1 << 1);
        g2 = ((g2 << 1) & 0x7fc) | (f2 << 1);
    }
#undef ST
}

/* gps/e1bcode.h:70-77: chip i = bit (3 - i%4) of hex digit i/4 of the ICD string.  The packed
 * table stores the same chip sequence LSB-first. */
void orc_e1b_chips(int prn, uint8_t *chips)
{
    const uint32_t *w = &E1B_WORDS[(prn - 1) * 128];
    for (int i = 0; i < 4092; i++) chips[i] = (uint8_t)((w[i >> 5] >> (i & 31)) & 1u);
}

/* gps/search.cpp:62-66 */
static inline float bipolar(int bit) { return bit ? -1.0f : 1.0f; }

/* ------------------------------------------------------------------------------------------
 * Half-band decimator: gps/search.cpp:99-166, column FT=0 ("remez") of COEF.
 * The literals are doubles narrowed to float exactly as the reference's initialiser does.
 * ---------------------------------------------------------------------------------------- */
static const float HB[NTAPS] = {
    -0.010233, 0, 0.010668, 0, -0.016324, 0, 0.024377, 0, -0.036482, 0, 0.056990, 0, -0.101993, 0,
    0.316926, 0.500009, 0.316926,
    0, -0.101993, 0, 0.056990, 0, -0.036482, 0, 0.024377, 0, -0.016324, 0, 0.010668, 0, -0.010233,
};

int orc_hb_decimate(int size, float *buf)
{
    const float c0 = HB[0], cm = HB[(NTAPS - 1) / 2];
    memset(buf + 2 * (size_t)size, 0, NTAPS * 2 * sizeof(float)); /* search.cpp:145 */
    for (int i = 0, o = 0; i < size; i += 2, ++o) {
        float accI = buf[2 * i] * c0;
        float accQ = buf[2 * i + 1] * c0;
        for (int j = 2; j < NTAPS; j += 2) { /* search.cpp:151-155 */
            const float c = HB[j];
            accI += buf[2 * (i + j)] * c;
            accQ += buf[2 * (i + j) + 1] * c;
        }
        accI += buf[2 * (i + (NTAPS - 1) / 2)] * cm; /* search.cpp:157-158 */
        accQ += buf[2 * (i + (NTAPS - 1) / 2) + 1] * cm;
        buf[2 * o] = accI;
        buf[2 * o + 1] = accQ;
    }
    return size / 2;
}

/* ------------------------------------------------------------------------------------------
 * FFT plans (one pair per process, like fwd_plan/rev_plan at search.cpp:51,240-241)
 * ---------------------------------------------------------------------------------------- */
static orc_fft_plan *g_fwd, *g_bwd;

static void plans_init(void)
{
#pragma omp critical(orc_plans)
    {
        if (!g_fwd) {
            g_bwd = orc_fft_plan_create(N, +1);
            g_fwd = orc_fft_plan_create(N, -1);
        }
    }
}

void orc_fft16384(float *buf, int sign)
{
    plans_init();
    float *scratch = (float *)malloc(sizeof(float) * 2 * N);
    orc_fft_execute(sign < 0 ? g_fwd : g_bwd, buf, scratch);
    free(scratch);
}

/* ------------------------------------------------------------------------------------------
 * Code replica: gps/search.cpp:205-206,243-276 (C/A, QZSS) and :306-338 (E1B).
 * ca_rate = CPS/FS = 1/16 exactly, so sample i carries chip floor(i/16) and the
 * interpolation at :261-262 multiplies by 1.0 and adds 0.0 (a no-op).
 * E1B adds the BOC(1,1) sub-carrier: chip ^ (phase >= 0.5)  (:317-318).
 * ---------------------------------------------------------------------------------------- */
void orc_code_baseband(const orc_sat *sat, float *out)
{
    float *buf = (float *)malloc(sizeof(float) * 2 * (NS + 2 * NTAPS));
    uint8_t chips[4092];
    int codelen;
    const int e1b = (sat->type == ORC_E1B);
    if (sat->type == ORC_SBAS) {
        /* search.cpp:244 `if (sp->type != Navstar && sp->type != QZSS) continue;` -- and the E1B loop (:306) takes
         * E1B rows only: an SBAS row of Sats[] gets no replica, code[sat] stays all zero (file-static storage). */
        memset(out, 0, sizeof(float) * 2 * N);
        free(buf);
        return;
    }
    if (e1b) {
        orc_e1b_chips(sat->prn, chips);
        codelen = 4092;
    } else {
        orc_ca_chips(sat->t1, sat->t2, chips);
        codelen = 1023;
    }
    for (int i = 0; i < NS; i++) {
        int c = chips[(i >> 4) % codelen];
        if (e1b) c ^= ((i & 15) >= 8);
        buf[2 * i] = bipolar(c);
        buf[2 * i + 1] = 0.0f;
    }
    int n = NS;
    for (int i = ORC_DECIM; i > 1; i >>= 1) n = orc_hb_decimate(n, buf); /* search.cpp:273-275 */
    memcpy(out, buf, sizeof(float) * 2 * N);
    free(buf);
}

void orc_code_spectrum(const orc_sat *sat, float *out)
{
    orc_code_baseband(sat, out);
    orc_fft16384(out, -1); /* search.cpp:280,342 */
}

/* ------------------------------------------------------------------------------------------
 * Capture front end: gps/search.cpp:382-447.
 *   sample i = bit (i&7) of byte i>>3 (:408-411); LO phase = i mod 4 because lo_rate = 4*FC/FS = 1
 *   I = bit ^ {1,1,0,0}[i&3], Q = bit ^ {1,0,0,1}[i&3] (:383-384,419-420)
 *   simd_bit2float + Bipolar(v >= 0) == Bipolar(bit) (support/simd.cpp:138-140, search.cpp:172-175)
 *   then two half-band stages (:437-441).
 * ---------------------------------------------------------------------------------------- */
void orc_capture_baseband(const uint8_t *packed, int half_rot, float *out)
{
    orc_capture_baseband_sm(packed, 1, half_rot, out);
}

/* sample_bits == 2 (extension, no reference counterpart): the magnitude plane follows the sign plane; the mixed
 * +-1 of the reference is scaled by 3 where the magnitude bit is set (x * 1.0f is exact, so sample_bits == 1 and
 * an all-zero magnitude plane give the reference's values bit for bit). */
void orc_capture_baseband_sm(const uint8_t *packed, int sample_bits, int half_rot, float *out)
{
    static const int lo_sin[4] = {1, 1, 0, 0};
    static const int lo_cos[4] = {1, 0, 0, 1};
    float *buf = (float *)malloc(sizeof(float) * 2 * (NS + 2 * NTAPS));
    const uint8_t *mag = (sample_bits == 2) ? packed + ORC_BLOCK_BYTES : NULL;
    for (int i = 0; i < NS; i++) {
        const int bit = (packed[i >> 3] >> (i & 7)) & 1;
        const float w = (mag && ((mag[i >> 3] >> (i & 7)) & 1)) ? 3.0f : 1.0f;
        buf[2 * i] = w * bipolar(bit ^ lo_sin[i & 3]);
        buf[2 * i + 1] = w * bipolar(bit ^ lo_cos[i & 3]);
    }
    int n = NS;
    n = orc_hb_decimate(n, buf);
    for (int i = ORC_DECIM >> 1; i > 1; i >>= 1) n = orc_hb_decimate(n, buf);
    if (half_rot) {
        /* extension: shift the spectrum down by half a bin */
        const double pi = 3.14159265358979323846264338327950288;
        for (int k = 0; k < N; k++) {
            const double a = pi * (double)k / (double)N;
            const float c = (float)cos(a), s = (float)(-sin(a));
            const float xr = buf[2 * k], xi = buf[2 * k + 1];
            buf[2 * k] = xr * c - xi * s;
            buf[2 * k + 1] = xr * s + xi * c;
        }
    }
    memcpy(out, buf, sizeof(float) * 2 * N);
    free(buf);
}

void orc_capture_spectrum(const uint8_t *packed, int half_rot, float *out)
{
    orc_capture_baseband(packed, half_rot, out);
    orc_fft16384(out, -1); /* search.cpp:447 */
}

/* ------------------------------------------------------------------------------------------
 * Correlate: gps/search.cpp:453-499, generalised.
 * ---------------------------------------------------------------------------------------- */
static void correlate_one(const float *Dall, int nvar, int K, const float *C, const float *Cnext, int L,
                          const orc_params *prm,
                          int sat_index, orc_record *rec, orc_cell *grid, float *prod, float *scratch,
                          float *P)
{
    float max_snr = 0;
    rec->sat = sat_index;
    rec->lag = 0;
    rec->dop = 0;
    rec->peak = 0;
    rec->noise = 0;
    rec->snr = 0;
    for (int h = prm->dop_lo; h <= prm->dop_hi; h++) {
        int var = 0, dop = h;
        if (prm->half_bin) {
            var = h & 1;          /* odd half-bin index -> pre-rotated spectrum */
            dop = (h - var) / 2;  /* exact: h - var is even */
        }
        float max_pwr = 0, tot_pwr = 0;
        int max_pwr_i = 0;
        int i;
        for (int b = 0; b < K; b++) {
            const float *data = Dall + (size_t)(b * nvar + var) * 2 * N;
            /* search.cpp:471-477 / support/simd.cpp:38-39: prod = conj(data) * code[(i-dop) mod N] */
            for (i = 0; i < N; i++) {
                int j = i - dop; /* index into the doubled row, relative to its second copy */
                const float *row = C;
                if (j < 0) j += N;
                else if (j >= N) { /* only for dop < 0: past the end of the doubled row */
                    j -= N;
                    if (prm->wrap_mode == ORC_WRAP_REFERENCE) row = Cnext;
                }
                const float dr = data[2 * i], di = data[2 * i + 1];
                const float cr = row ? row[2 * j] : 0.0f, ci = row ? row[2 * j + 1] : 0.0f;
                prod[2 * i] = dr * cr + di * ci;
                prod[2 * i + 1] = dr * ci - di * cr;
            }
            orc_fft_execute(g_bwd, prod, scratch); /* search.cpp:481 */
            if (K == 1) {
                /* search.cpp:486-490 */
                for (i = 0; i < L; i++) {
                    const float pwr = prod[2 * i] * prod[2 * i] + prod[2 * i + 1] * prod[2 * i + 1];
                    if (pwr > max_pwr) max_pwr = pwr, max_pwr_i = i;
                    tot_pwr += pwr;
                }
            } else {
                /* extension: block b starts 16 /4-samples later in code phase, so
                 * lag n of block 0 is lag n + 16 b of block b.  The index wraps modulo the FFT length
                 * (the transform's own circular lag axis), not modulo L: for the 1 ms codes lags
                 * L..L+16b-1 are the same code phases one period later, and this form is a plain
                 * circular time shift of the block (no scatter) -- see DESIGN.md "non-coherent sum". */
                /* code-Doppler compensation (extension, acq_oracle.h): block b's code phase has advanced by a
                 * further s(b, h) /4-samples at Doppler index h */
                const int cs = prm->code_doppler ? orc_code_shift(b, h, prm->half_bin) : 0;
                for (i = 0; i < L; i++) {
                    const int m = (i + 16 * b + cs + N) % N;
                    const float pwr = prod[2 * m] * prod[2 * m] + prod[2 * m + 1] * prod[2 * m + 1];
                    if (b == 0) P[i] = pwr; else P[i] += pwr;
                }
            }
        }
        if (K > 1) {
            for (i = 0; i < L; i++) {
                const float pwr = P[i];
                if (pwr > max_pwr) max_pwr = pwr, max_pwr_i = i;
                tot_pwr += pwr;
            }
        }
        i = L;
        const float ave_pwr = tot_pwr / i; /* search.cpp:493 */
        const float snr = max_pwr / ave_pwr;
        if (grid) {
            orc_cell *g = &grid[h - prm->dop_lo];
            g->peak = max_pwr;
            g->noise = ave_pwr;
            g->snr = snr;
            g->lag = max_pwr_i;
        }
        if (snr > max_snr) { /* search.cpp:495 */
            max_snr = snr;
            rec->dop = h;
            rec->lag = max_pwr_i;
            rec->peak = max_pwr;
            rec->noise = ave_pwr;
            rec->snr = snr;
        }
    }
}

int orc_search_pre(const uint8_t *packed, const orc_sat *sats, int n_sats, const float *spectra,
                   const int32_t *sel, int n_sel, const orc_params *prm, orc_record *out,
                   orc_cell *grid, int nthreads)
{
    if (!packed || !sats || !prm || !out || n_sats <= 0 || prm->k_noncoh < 1 || prm->dop_hi < prm->dop_lo)
        return -1;
    if (!sel) n_sel = n_sats;
    for (int s = 0; s < n_sel; s++) {
        const int sat = sel ? sel[s] : s;
        if (sat < 0 || sat >= n_sats) return -2;
    }
    plans_init();
    const int K = prm->k_noncoh;
    const int nvar = prm->half_bin ? 2 : 1;
    const int n_dop = prm->dop_hi - prm->dop_lo + 1;

    float *D = (float *)malloc(sizeof(float) * 2 * N * (size_t)K * nvar);
    if (!D) return -3;
#pragma omp parallel for schedule(dynamic) num_threads(orc_threads(nthreads))
    for (int bv = 0; bv < K * nvar; bv++) {
        const int b = bv / nvar, v = bv % nvar;
        float *scratch = (float *)malloc(sizeof(float) * 2 * N);
        float *d = D + (size_t)bv * 2 * N;
        orc_capture_baseband_sm(packed + (size_t)b * ORC_CAPTURE_BLOCK_BYTES(prm->sample_bits), prm->sample_bits, v, d);
        orc_fft_execute(g_fwd, d, scratch);
        free(scratch);
    }

#pragma omp parallel num_threads(orc_threads(nthreads))
    {
        float *prod = (float *)malloc(sizeof(float) * 2 * N);
        float *scratch = (float *)malloc(sizeof(float) * 2 * N);
        float *Cown = spectra ? NULL : (float *)malloc(sizeof(float) * 4 * N);
        float *P = (float *)malloc(sizeof(float) * N);
#pragma omp for schedule(dynamic)
        for (int s = 0; s < n_sel; s++) {
            const int sat = sel ? sel[s] : s;
            const float *C, *Cnext = NULL; /* next table row; NULL = the zero rows after the last sat */
            if (spectra) {
                C = spectra + (size_t)sat * 2 * N;
                if (sat + 1 < n_sats) Cnext = spectra + (size_t)(sat + 1) * 2 * N;
            } else {
                orc_code_baseband(&sats[sat], Cown);
                orc_fft_execute(g_fwd, Cown, scratch);
                C = Cown;
                if (sat + 1 < n_sats && prm->wrap_mode == ORC_WRAP_REFERENCE && prm->dop_lo < 0) {
                    orc_code_baseband(&sats[sat + 1], Cown + 2 * N);
                    orc_fft_execute(g_fwd, Cown + 2 * N, scratch);
                    Cnext = Cown + 2 * N;
                }
            }
            /* search.cpp:456,486: lags scanned = SAMPLE_RATE/1000 * code_period_ms */
            const int L = (sats[sat].type == ORC_E1B) ? 16368 : 4092;
            correlate_one(D, nvar, K, C, Cnext, L, prm, sat, &out[s], grid ? grid + (size_t)s * n_dop : NULL, prod,
                          scratch, P);
        }
        free(prod);
        free(scratch);
        free(Cown);
        free(P);
    }
    free(D);
    return 0;
}

int orc_search(const uint8_t *packed, const orc_sat *sats, int n_sats, const int32_t *sel, int n_sel,
               const orc_params *prm, orc_record *out, orc_cell *grid, int nthreads)
{
    return orc_search_pre(packed, sats, n_sats, NULL, sel, n_sel, prm, out, grid, nthreads);
}

/* ------------------------------------------------------------------------------------------
 * Acquisition refinement (extension; see acq_oracle.h).  Direct evaluation of the correlation sum of
 * search.cpp:471-481 at five points around a record's peak, in double precision.
 * ---------------------------------------------------------------------------------------- */
static void code_bin(const float *C, const float *Cnext, int wrap_mode, int i, int dop, double *cr, double *ci)
{
    /* same row addressing as correlate_one (search.cpp:471 with the doubled rows) */
    int j = i - dop;
    const float *row = C;
    if (j < 0) j += N;
    else if (j >= N) {
        j -= N;
        if (wrap_mode == ORC_WRAP_REFERENCE) row = Cnext;
    }
    *cr = row ? row[2 * j] : 0.0;
    *ci = row ? row[2 * j + 1] : 0.0;
}

int orc_refine(const uint8_t *packed, const orc_sat *sats, int n_sats, const orc_params *prm,
               const orc_record *rec, int n_rec, orc_fine *out, int nthreads)
{
    if (!packed || !sats || !prm || !rec || !out || n_sats <= 0 || n_rec < 0) return -1;
    for (int s = 0; s < n_rec; s++)
        if (rec[s].sat < 0 || rec[s].sat >= n_sats) return -2;
    plans_init();
    const int K = prm->k_noncoh;
    const int nvar = prm->half_bin ? 2 : 1;
    float *D = (float *)malloc(sizeof(float) * 2 * N * (size_t)K * nvar);
    double *tw = (double *)malloc(sizeof(double) * 2 * N);
    if (!D || !tw) return -3;
    for (int i = 0; i < N; i++) {
        tw[2 * i] = cos(2.0 * M_PI * i / N);
        tw[2 * i + 1] = sin(2.0 * M_PI * i / N);
    }
#pragma omp parallel for schedule(dynamic) num_threads(orc_threads(nthreads))
    for (int bv = 0; bv < K * nvar; bv++) {
        float *scratch = (float *)malloc(sizeof(float) * 2 * N);
        float *d = D + (size_t)bv * 2 * N;
        orc_capture_baseband_sm(packed + (size_t)(bv / nvar) * ORC_CAPTURE_BLOCK_BYTES(prm->sample_bits), prm->sample_bits,
                                bv % nvar, d);
        orc_fft_execute(g_fwd, d, scratch);
        free(scratch);
    }
#pragma omp parallel for schedule(dynamic) num_threads(orc_threads(nthreads))
    for (int s = 0; s < n_rec; s++) {
        const int sat = rec[s].sat;
        float *C = (float *)malloc(sizeof(float) * 4 * N), *scratch = (float *)malloc(sizeof(float) * 2 * N);
        const float *Cnext = NULL;
        orc_code_baseband(&sats[sat], C);
        orc_fft_execute(g_fwd, C, scratch);
        if (sat + 1 < n_sats && prm->wrap_mode == ORC_WRAP_REFERENCE) {
            orc_code_baseband(&sats[sat + 1], C + 2 * N);
            orc_fft_execute(g_fwd, C + 2 * N, scratch);
            Cnext = C + 2 * N;
        }
        const int e1b = sats[sat].type == ORC_E1B;
        const int L = e1b ? 16368 : 4092;
        const int var = prm->half_bin ? (rec[s].dop & 1) : 0;
        const int dop = prm->half_bin ? (rec[s].dop - var) / 2 : rec[s].dop;
        const int n = rec[s].lag;
        double num = 0, den = 0, early = 0, late = 0, peak = 0;
        for (int b = 0; b < K; b++) {
            const float *data = D + (size_t)(b * nvar + var) * 2 * N;
            /* lag n of block 0 is lag n + 16 b (+ the code-Doppler shift) of block b (see correlate_one) */
            const int nb = (n + 16 * b + (prm->code_doppler ? orc_code_shift(b, rec[s].dop, prm->half_bin) : 0) + N) % N;
            double R[5][2] = {{0}};
            for (int i = 0; i < N; i++) {
                const double dr = data[2 * i], di = data[2 * i + 1];
                double c[3][2];
                for (int j = 0; j < 3; j++) code_bin(C, Cnext, prm->wrap_mode, i, dop - 1 + j, &c[j][0], &c[j][1]);
                double pr[3][2];
                for (int j = 0; j < 3; j++) { /* conj(data) * code, support/simd.cpp:38-39 */
                    pr[j][0] = dr * c[j][0] + di * c[j][1];
                    pr[j][1] = dr * c[j][1] - di * c[j][0];
                }
                const int lag[5] = {nb, (nb + N - 1) % N, nb, (nb + 1) % N, nb};
                const int which[5] = {0, 1, 1, 1, 2};
                for (int j = 0; j < 5; j++) {
                    const int m = (int)(((long long)i * lag[j]) % N);
                    const double wr = tw[2 * m], wi = tw[2 * m + 1];
                    R[j][0] += pr[which[j]][0] * wr - pr[which[j]][1] * wi;
                    R[j][1] += pr[which[j]][0] * wi + pr[which[j]][1] * wr;
                }
            }
            /* X_d = r_d e^{-j 2 pi d n/N}: relative to the centre bin, r_{d-1} gets e^{+j 2 pi n/N}, r_{d+1} its conjugate */
            const double wr = tw[2 * nb], wi = tw[2 * nb + 1];
            const double Xm[2] = {R[0][0] * wr - R[0][1] * wi, R[0][0] * wi + R[0][1] * wr};
            const double Xp[2] = {R[4][0] * wr + R[4][1] * wi, -R[4][0] * wi + R[4][1] * wr};
            const double a[2] = {Xm[0] - Xp[0], Xm[1] - Xp[1]};
            const double g[2] = {2 * R[2][0] - Xm[0] - Xp[0], 2 * R[2][1] - Xm[1] - Xp[1]};
            num += a[0] * g[0] + a[1] * g[1];
            den += g[0] * g[0] + g[1] * g[1];
            early += R[1][0] * R[2][0] + R[1][1] * R[2][1];
            late += R[3][0] * R[2][0] + R[3][1] * R[2][1];
            peak += R[2][0] * R[2][0] + R[2][1] * R[2][1];
        }
        double delta = den > 0 ? num / den : 0.0;
        if (delta > 1) delta = 1;
        if (delta < -1) delta = -1;
        double eps = (early + late) > 0 ? (e1b ? 1.0 / 3.0 : 3.0) * (late - early) / (early + late) : 0.0;
        if (eps > 1) eps = 1;
        if (eps < -1) eps = -1;
        orc_fine *o = &out[s];
        o->dop_hz = (float)(((prm->half_bin ? 0.5 * rec[s].dop : (double)rec[s].dop) + delta) * (16.368e6 / 65536.0));
        o->code_fs = (float)(ORC_DECIM * ((double)n + eps));
        o->peak = (float)peak;
        int cs = (int)lrintf(o->code_fs) % (L * ORC_DECIM);
        if (cs < 0) cs += L * ORC_DECIM;
        o->ca_shift = cs;
        free(C);
        free(scratch);
    }
    free(D);
    free(tw);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic capture generator (test fixture; not a reference function).
 *   s[i] = sum_k A_k * c_k(i + tau_k) * cos(2*pi*(FC + f_k)*i/FS + phi_k) + n[i],  n ~ N(0,1)
 *   A_k  = sqrt(4 * 10^(CN0/10) / FS);  bit = (s < 0);  packed LSB-first (search.cpp:408-411).
 * Counter-based RNG so any sample can be generated independently (OpenMP safe, seed-stable).
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int orc_gen_capture(uint64_t seed, int n_blocks, const orc_sat *sats, int n_sats, const orc_signal *sig,
                    int n_sig, uint8_t *packed)
{
    return orc_gen_capture_sm(seed, n_blocks, sats, n_sats, sig, n_sig, 1, 0.0, packed);
}

int orc_gen_capture_sm(uint64_t seed, int n_blocks, const orc_sat *sats, int n_sats, const orc_signal *sig,
                       int n_sig, int sample_bits, double mag_thr, uint8_t *packed)
{
    if (n_blocks < 1 || n_sig < 0 || !packed || (sample_bits != 1 && sample_bits != 2)) return -1;
    const double FS = 16.368e6, two_pi = 6.283185307179586476925286766559;
    uint8_t *chips = (uint8_t *)malloc((size_t)(n_sig > 0 ? n_sig : 1) * 4092);
    int *codelen = (int *)malloc(sizeof(int) * (n_sig > 0 ? n_sig : 1));
    int *boc = (int *)malloc(sizeof(int) * (n_sig > 0 ? n_sig : 1));
    double *amp = (double *)malloc(sizeof(double) * (n_sig > 0 ? n_sig : 1));
    for (int k = 0; k < n_sig; k++) {
        if (sig[k].sat < 0 || sig[k].sat >= n_sats) { free(chips); free(codelen); free(boc); free(amp); return -2; }
        const orc_sat *sp = &sats[sig[k].sat];
        if (sp->type == ORC_E1B) {
            orc_e1b_chips(sp->prn, chips + (size_t)k * 4092);
            codelen[k] = 4092;
            boc[k] = 1;
        } else {
            orc_ca_chips(sp->t1, sp->t2, chips + (size_t)k * 4092);
            codelen[k] = 1023;
            boc[k] = 0;
        }
        amp[k] = sqrt(4.0 * pow(10.0, sig[k].cn0_dbhz / 10.0) / FS);
    }
    const long total_bytes = (long)n_blocks * ORC_BLOCK_BYTES;
    const uint64_t key = splitmix64(seed ^ 0xA5A5A5A55A5A5A5Aull);
#pragma omp parallel for schedule(static)
    for (long by = 0; by < total_bytes; by++) {
        unsigned byte = 0, mbyte = 0;
        for (int bb = 0; bb < 8; bb++) {
            const long i = by * 8 + bb;
            const uint64_t z1 = splitmix64(key + 2 * (uint64_t)i);
            const uint64_t z2 = splitmix64(key + 2 * (uint64_t)i + 1);
            const double u1 = ((double)(z1 >> 11) + 1.0) * (1.0 / 9007199254740992.0);
            const double u2 = (double)(z2 >> 11) * (1.0 / 9007199254740992.0);
            double s = sqrt(-2.0 * log(u1)) * cos(two_pi * u2);
            for (int k = 0; k < n_sig; k++) {
                const long idx = (sig[k].code_doppler ? (long)floor((double)i * (1.0 + sig[k].doppler_hz / 1575.42e6)) : i)
                                 + sig[k].tau;
                int c = chips[(size_t)k * 4092 + (idx >> 4) % codelen[k]];
                if (boc[k]) c ^= ((idx & 15) >= 8);
                double sgn = c ? -1.0 : 1.0;
                if (sig[k].flip_ms > 0 && ((idx / 16368 / sig[k].flip_ms) & 1)) sgn = -sgn;
                double cyc = sig[k].doppler_hz / FS * (double)i;
                cyc -= floor(cyc);
                cyc += 0.25 * (double)(i & 3); /* FC/FS = 1/4 exactly */
                s += amp[k] * sgn * cos(two_pi * cyc + sig[k].phase);
            }
            byte |= (unsigned)(s < 0.0) << bb;
            mbyte |= (unsigned)(fabs(s) > mag_thr) << bb;
        }
        if (sample_bits == 2) { /* block = [sign plane][magnitude plane] */
            const long blk = by / ORC_BLOCK_BYTES, off = by % ORC_BLOCK_BYTES;
            packed[blk * 2 * ORC_BLOCK_BYTES + off] = (uint8_t)byte;
            packed[blk * 2 * ORC_BLOCK_BYTES + ORC_BLOCK_BYTES + off] = (uint8_t)mbyte;
        } else {
            packed[by] = (uint8_t)byte;
        }
    }
    free(chips);
    free(codelen);
    free(boc);
    free(amp);
    return 0;
}
