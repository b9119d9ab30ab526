/*
 * orc_fft.c -- Stockham autosort radix-4 (+ one radix-2 pass for odd log2 n) fp32 FFT.
 * TEST INFRASTRUCTURE ONLY (see orc_fft.h).  Twiddles are computed in double and
 * rounded once to float; butterflies are plain fp32 (compile with -ffp-contract=off).
 */
#include "orc_fft.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_fft_plan {
    int n;
    int sign;
    float *tw; /* n interleaved complex: exp(sign * 2*pi*i*k/n), k = 0..n-1 */
};

orc_fft_plan *orc_fft_plan_create(int n, int sign)
{
    if (n < 2 || (n & (n - 1)) != 0 || (sign != 1 && sign != -1)) return NULL;
    orc_fft_plan *p = (orc_fft_plan *)malloc(sizeof(*p));
    if (!p) return NULL;
    p->n = n;
    p->sign = sign;
    p->tw = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    if (!p->tw) { free(p); return NULL; }
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < n; k++) {
        double a = two_pi * (double)k / (double)n;
        p->tw[2 * k + 0] = (float)cos(a);
        p->tw[2 * k + 1] = (float)(sign * sin(a));
    }
    return p;
}

void orc_fft_plan_destroy(orc_fft_plan *p)
{
    if (!p) return;
    free(p->tw);
    free(p);
}

int orc_fft_plan_n(const orc_fft_plan *p) { return p->n; }

/* One radix-4 pass: sub-transform length len (a multiple of 4), stride s.
 * x -> y.  jsign = +1 multiplies by +i, -1 by -i in the odd outputs. */
static void pass4(int len, int s, int nfull, const float *tw, int sign,
                  const float *restrict x, float *restrict y)
{
    const int q4 = len / 4;
    const int tstep = nfull / len;
    for (int p = 0; p < q4; p++) {
        const float w1r = tw[2 * (1 * p * tstep) + 0], w1i = tw[2 * (1 * p * tstep) + 1];
        const float w2r = tw[2 * (2 * p * tstep) + 0], w2i = tw[2 * (2 * p * tstep) + 1];
        const float w3r = tw[2 * (3 * p * tstep) + 0], w3i = tw[2 * (3 * p * tstep) + 1];
        const float *xa = x + 2 * (size_t)s * (p + 0 * q4);
        const float *xb = x + 2 * (size_t)s * (p + 1 * q4);
        const float *xc = x + 2 * (size_t)s * (p + 2 * q4);
        const float *xd = x + 2 * (size_t)s * (p + 3 * q4);
        float *y0 = y + 2 * (size_t)s * (4 * p + 0);
        float *y1 = y + 2 * (size_t)s * (4 * p + 1);
        float *y2 = y + 2 * (size_t)s * (4 * p + 2);
        float *y3 = y + 2 * (size_t)s * (4 * p + 3);
        for (int q = 0; q < s; q++) {
            const float ar = xa[2 * q], ai = xa[2 * q + 1];
            const float br = xb[2 * q], bi = xb[2 * q + 1];
            const float cr = xc[2 * q], ci = xc[2 * q + 1];
            const float dr = xd[2 * q], di = xd[2 * q + 1];
            const float apcr = ar + cr, apci = ai + ci;
            const float amcr = ar - cr, amci = ai - ci;
            const float bpdr = br + dr, bpdi = bi + di;
            const float bmdr = br - dr, bmdi = bi - di;
            /* (sign*i)*(b-d) */
            const float jr = (sign > 0) ? -bmdi : bmdi;
            const float ji = (sign > 0) ? bmdr : -bmdr;
            const float t1r = amcr + jr, t1i = amci + ji;
            const float t2r = apcr - bpdr, t2i = apci - bpdi;
            const float t3r = amcr - jr, t3i = amci - ji;
            y0[2 * q] = apcr + bpdr;
            y0[2 * q + 1] = apci + bpdi;
            y1[2 * q] = t1r * w1r - t1i * w1i;
            y1[2 * q + 1] = t1r * w1i + t1i * w1r;
            y2[2 * q] = t2r * w2r - t2i * w2i;
            y2[2 * q + 1] = t2r * w2i + t2i * w2r;
            y3[2 * q] = t3r * w3r - t3i * w3i;
            y3[2 * q + 1] = t3r * w3i + t3i * w3r;
        }
    }
}

/* Final radix-2 pass (len == 2): no twiddles. */
static void pass2(int s, const float *restrict x, float *restrict y)
{
    for (int q = 0; q < s; q++) {
        const float ar = x[2 * q], ai = x[2 * q + 1];
        const float br = x[2 * (q + s)], bi = x[2 * (q + s) + 1];
        y[2 * q] = ar + br;
        y[2 * q + 1] = ai + bi;
        y[2 * (q + s)] = ar - br;
        y[2 * (q + s) + 1] = ai - bi;
    }
}

void orc_fft_execute(const orc_fft_plan *p, float *buf, float *scratch)
{
    float *x = buf, *y = scratch;
    int len = p->n, s = 1;
    while (len >= 4) {
        pass4(len, s, p->n, p->tw, p->sign, x, y);
        float *t = x; x = y; y = t;
        len /= 4;
        s *= 4;
    }
    if (len == 2) {
        pass2(s, x, y);
        float *t = x; x = y; y = t;
    }
    if (x != buf) memcpy(buf, x, sizeof(float) * 2 * (size_t)p->n);
}
