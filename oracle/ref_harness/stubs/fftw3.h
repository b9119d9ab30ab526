/*
 * fftw3.h -- FFTW3 single-precision API shim over the oracle FFT (oracle/orc_fft.c).
 * TEST INFRASTRUCTURE ONLY.  Covers exactly the entry points the reference search path
 * uses (gps/search.cpp:44,51,240-241,280,342,355-356,447,481).  FFTW itself is absent
 * from this image ("parity unpinned" at this boundary, see oracle/acq_oracle.h).
 */
#pragma once
typedef float fftwf_complex[2];
typedef struct ref_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
#ifdef __cplusplus
extern "C" {
#endif
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
