// acq_fft.cuh -- register/shared-memory FFT building blocks for the acquisition kernels (sm_100a).
//
// The 16384-point inverse transform of the correlator (reference gps/search.cpp:241,481:
// fftwf BACKWARD, unnormalised, e^{+j2*pi*k*n/N}) is decomposed as
//
//     k = 1024*a + 64*b + 4*c + k2        n = n0 + 16*n1 + 256*n2 + 4096*m
//     (a,b,c,n0,n1,n2 in 0..15;  k2,m in 0..3)
//
//     W^{kn} = W16^{a n0} . W256^{b n0} W16^{b n1} . W16384^{(4c+k2)(n0+16 n1)} W64^{(4c+k2) n2} . i^{k2 m}
//
// i.e. for each input residue k2 (a "polyphase" quarter of the spectrum, contiguous in the
// polyphase HBM layout) one 4096-point sub-FFT made of three radix-16 passes held in registers
// (stage A over a, stage B over b, stage C over c) with two shared-memory exchanges, followed by
// a radix-4 combine over k2 that lives in registers.  The C/A search only needs the first 4092
// lags (m = 0), so its combine is a plain accumulation and three quarters of the last radix-4
// are never computed.
//
// 256 threads, 16 points per thread.  Thread roles:
//   stage A: t = 16*b + c   holds a = 0..15      -> outputs n0
//   stage B: t = 16*n0 + c  holds b = 0..15      -> outputs n1
//   stage C: t = n0 + 16*n1 holds c = 0..15      -> outputs n2   (lag n = t + 256*n2 + 4096*m)
//
// Twiddles: stage A uses W4096^{t*n0} from a per-CTA shared table (thread-constant, 30 KiB) times
// the small constant W16384^{k2*n0}; stage B uses W1024^{(4c+k2)*n1} (7.5 KiB shared table);
// stage C's W64^{k2*n2} are constants.  All tables are computed on the host in double precision.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "acq_geom.h"

namespace acq {

// Constant twiddles (filled by the host at engine creation; identical on every device).
//   c_cA[k2][n0] = W16384^{k2*n0}     c_cC[k2][n2] = W64^{k2*n2}      (e^{+j...}: inverse transform)
// Defined here (not extern): this header is included by exactly one translation unit
// (acq_kernels.cu), so no relocatable device code is needed.
__constant__ float2 c_cA[4][16];
__constant__ float2 c_cC[4][16];

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
// Complex add/sub as ONE packed f32x2 instruction (sm_100 FADD2 / FFMA2): the FMA pipe does the same work
// as two scalar adds, but the instruction occupies a single issue slot, which is what the add-dominated
// butterflies are short of.  (SASS: FADD2, FFMA2 with a broadcast -1 immediate.)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
// a + j*b and a - j*b: j*b = (-b.y, b.x) is a lane swap (free operand modifier) times the sign pair (-1,+1)
__device__ __forceinline__ float2 cadd_jb(float2 a, float2 b)
{
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.0f, 1.0f), a);
}
__device__ __forceinline__ float2 csub_jb(float2 a, float2 b)
{
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.0f, -1.0f), a);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
// conj(a) * b  (reference support/simd.cpp:12-40: re = ar*br + ai*bi, im = ar*bi - ai*br)
__device__ __forceinline__ float2 cmul_conj_a(float2 a, float2 b)
{
    return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }  // * (+j)
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }  // * (-j)

// Inverse radix-4 butterfly (W4 = +j), in place on four registers.
__device__ __forceinline__ void radix4_inv(float2 &a, float2 &b, float2 &c, float2 &d)
{
    const float2 apc = cadd(a, c), amc = csub(a, c);
    const float2 bpd = cadd(b, d), bmd = csub(b, d);
    a = cadd(apc, bpd);
    b = cadd_jb(amc, bmd);
    c = csub(apc, bpd);
    d = csub_jb(amc, bmd);
}

// Inverse 16-point DFT in registers: X[n] = sum_a x[a] e^{+2*pi*j*a*n/16}.
// Input x[a] in natural order.  Output X[n] is left at register index  r16(n) = (n >> 2) | ((n & 3) << 2).
__host__ __device__ constexpr int r16(int n) { return (n >> 2) | ((n & 3) << 2); }

__device__ __forceinline__ void radix16_inv(float2 (&x)[16])
{
    constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
    constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
    constexpr float R2 = 0.70710678118654752440f;  // sqrt(1/2)
    // layer 1: for each a0, radix-4 over a = a0 + 4*a1  ->  u[a0][nl] at index a0 + 4*nl
#pragma unroll
    for (int a0 = 0; a0 < 4; a0++) radix4_inv(x[a0], x[a0 + 4], x[a0 + 8], x[a0 + 12]);
    // twiddle u[a0][nl] *= W16^{a0*nl}
    x[5] = cmul(x[5], make_float2(C1, S1));                          // W16^1
    x[9] = make_float2((x[9].x - x[9].y) * R2, (x[9].x + x[9].y) * R2);   // W16^2 = (1+j)/sqrt2
    x[13] = cmul(x[13], make_float2(S1, C1));                        // W16^3
    x[6] = make_float2((x[6].x - x[6].y) * R2, (x[6].x + x[6].y) * R2);   // W16^2
    x[10] = mul_pj(x[10]);                                           // W16^4 = j
    x[14] = make_float2(-(x[14].x + x[14].y) * R2, (x[14].x - x[14].y) * R2);  // W16^6 = (-1+j)/sqrt2
    x[7] = cmul(x[7], make_float2(S1, C1));                          // W16^3
    x[11] = make_float2(-(x[11].x + x[11].y) * R2, (x[11].x - x[11].y) * R2);  // W16^6
    x[15] = cmul(x[15], make_float2(-C1, -S1));                      // W16^9
    // layer 2: for each nl, radix-4 over a0  ->  X[nl + 4*nh] at index nh + 4*nl
#pragma unroll
    for (int nl = 0; nl < 4; nl++) radix4_inv(x[4 * nl], x[4 * nl + 1], x[4 * nl + 2], x[4 * nl + 3]);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory workspace of one 256-thread FFT team.
// ---------------------------------------------------------------------------------------------
struct FftSmem {
    float2 *T1;  // [15][256]
    float2 *T2;  // [4][15][16]
    float2 *S1;  // [16][256]
    float2 *S2;  // [16][272]
};

__host__ __device__ constexpr size_t fft_smem_bytes()
{
    return sizeof(float2) * (size_t)(kT1Elems + kT2Elems + kS1Elems + kS2Elems);
}

__device__ __forceinline__ FftSmem fft_smem_carve(unsigned char *base)
{
    FftSmem s;
    s.T1 = reinterpret_cast<float2 *>(base);
    s.T2 = s.T1 + kT1Elems;
    s.S1 = s.T2 + kT2Elems;
    s.S2 = s.S1 + kS1Elems;
    return s;
}

// Copy the twiddle tables (global, [kT1Elems + kT2Elems] float2) into shared memory.
// Ends with a barrier so the tables are visible to every thread of the CTA.
__device__ __forceinline__ void fft_load_tables(const FftSmem &s, const float2 *__restrict__ g_tables, int t,
                                                int nthreads)
{
    const float4 *src = reinterpret_cast<const float4 *>(g_tables);
    float4 *dst = reinterpret_cast<float4 *>(s.T1);
    constexpr int n4 = (kT1Elems + kT2Elems) / 2;
    for (int i = t; i < n4; i += nthreads) dst[i] = __ldg(src + i);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// One 4096-point sub-FFT (input residue k2).
//   in : x[a]  = input element k = 1024*a + 4*t + k2  (stage-A role of thread t)
//   out: x[r16(n2)] = sum over this residue's inputs of  in[k] * W16384^{k*n} / W64^{k2*n2}-free part,
//        i.e. the stage-C DFT output for lag n = t + 256*n2 BEFORE the constant W64^{k2*n2} factor
//        and the i^{k2*m} combine (the caller applies both).
// Contains two __syncthreads(); S1/S2 double buffering makes back-to-back calls safe with no
// further barrier (see DESIGN.md "exchange hazards").
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void subfft4096_inv(float2 (&x)[16], const int k2, const FftSmem &s, const int t)
{
    // ---- stage A: DFT over a, twiddle W4096^{t*n0} * W16384^{k2*n0}, scatter by n0
    // Twiddles are fetched in two batches of registers AHEAD of the stores that use them: the compiler
    // cannot hoist a shared-memory load above a shared-memory store (possible aliasing), and a
    // load -> multiply -> store chain per element would expose the LDS latency 15 times per stage.
    float2 tw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = s.T1[i * 256 + t];  // n0 = 1..8
    radix16_inv(x);
    if (k2 != 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) tw[i] = cmul(tw[i], c_cA[k2][i + 1]);
    }
    s.S1[t] = x[r16(0)];
#pragma unroll
    for (int i = 0; i < 8; i++) s.S1[(i + 1) * 256 + t] = cmul(x[r16(i + 1)], tw[i]);
#pragma unroll
    for (int i = 0; i < 7; i++) tw[i] = s.T1[(i + 8) * 256 + t];  // n0 = 9..15
    if (k2 != 0) {
#pragma unroll
        for (int i = 0; i < 7; i++) tw[i] = cmul(tw[i], c_cA[k2][i + 9]);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) s.S1[(i + 9) * 256 + t] = cmul(x[r16(i + 9)], tw[i]);
    // stage-B twiddles W1024^{(4c+k2)*n1} do not depend on the exchange: first batch before the barrier
    const float2 *twp = s.T2 + k2 * (15 * 16) + (t & 15);
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = twp[i * 16];  // n1 = 1..8
    __syncthreads();
    // ---- stage B: thread (n0, c) = (t >> 4, t & 15) gathers b = 0..15
    {
        const float2 *src = s.S1 + (t & ~15) * 16 + (t & 15);
#pragma unroll
        for (int b = 0; b < 16; b++) x[b] = src[16 * b];
    }
    radix16_inv(x);
    {
        float2 *dst = s.S2 + 17 * (t >> 4) + (t & 15);
        dst[0] = x[r16(0)];
#pragma unroll
        for (int i = 0; i < 8; i++) dst[(i + 1) * kS2Stride] = cmul(x[r16(i + 1)], tw[i]);
#pragma unroll
        for (int i = 0; i < 7; i++) tw[i] = twp[(i + 8) * 16];  // n1 = 9..15
#pragma unroll
        for (int i = 0; i < 7; i++) dst[(i + 9) * kS2Stride] = cmul(x[r16(i + 9)], tw[i]);
    }
    __syncthreads();
    // ---- stage C: thread t = n0 + 16*n1 gathers c = 0..15
    {
        const float2 *src = s.S2 + (t >> 4) * kS2Stride + 17 * (t & 15);
#pragma unroll
        for (int c = 0; c < 16; c++) x[c] = src[c];
    }
    radix16_inv(x);
}

}  // namespace acq
