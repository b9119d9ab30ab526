/*
 * orc_fft.h -- in-repo single-precision complex FFT used by the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, link or call it.
 *
 * It stands in for the three FFTW3 (single precision) entry points the reference
 * uses on the search path -- fftwf_plan_dft_1d / fftwf_execute / fftwf_destroy_plan
 * (reference gps/search.cpp:240-241,280,447,481,355-356).  FFTW is an un-vendored,
 * un-pinned third-party dependency of the reference (Makefile:16-25,249,365-366)
 * and is not installed in this image, so the FFT arithmetic is "parity unpinned":
 * this file restates the published definition FFTW documents for those calls --
 *     forward : Y[k] = sum_n X[n] * exp(-2*pi*i*k*n/N)      (FFTW_FORWARD  = -1)
 *     backward: Y[n] = sum_k X[k] * exp(+2*pi*i*k*n/N)      (FFTW_BACKWARD = +1)
 * both unnormalised, complex interleaved float, in place.
 */
#ifndef ORC_FFT_H
#define ORC_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_fft_plan orc_fft_plan;

/* n must be a power of two >= 2.  sign = -1 forward, +1 backward. */
orc_fft_plan *orc_fft_plan_create(int n, int sign);
void orc_fft_plan_destroy(orc_fft_plan *p);
/* In-place transform of n interleaved complex floats.  Thread-safe as long as each
 * thread passes its own scratch (n complex floats). */
void orc_fft_execute(const orc_fft_plan *p, float *buf, float *scratch);
int orc_fft_plan_n(const orc_fft_plan *p);

#ifdef __cplusplus
}
#endif
#endif
