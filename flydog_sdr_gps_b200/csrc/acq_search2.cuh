// acq_search2.cuh -- two-Doppler-bins-per-thread search kernel on packed f32x2 arithmetic (sm_100a).
//
// Blackwell's FFMA2 / FADD2 / FMUL2 execute two fp32 operations per instruction on an aligned register
// pair, take a scalar register as a BROADCAST operand (R.F32), and carry free negate / lane-swap operand
// modifiers.  The FMA pipe does the same work either way, but each packed instruction costs one issue
// slot instead of two.  This kernel puts two adjacent Doppler bins (dop, dop+1) of the same (capture,
// satellite) into the two lanes:
//   - every butterfly, twiddle and product instruction serves both bins: half the FP issue slots;
//   - the capture spectrum D and all twiddles are loaded ONCE for both bins and enter as scalar
//     broadcast operands: 25 % less L2->SM traffic and half the twiddle reads per tile;
//   - exchanges move float4 (reA, reB, imA, imB): half the shared-memory instructions per tile.
// A complex value is kept as two pairs, re = (re_A, re_B) and im = (im_A, im_B); no lane ever needs data
// from the other, so there is no shuffling.  One 256-thread CTA per SM (the doubled exchange buffers
// and ~200 registers per thread leave room for exactly one).
//
// Same index algebra as acq_fft.cuh (stage A over a, B over b, C over c; residues k2 accumulated).
#pragma once

#include "acq_fft.cuh"

namespace acq {

struct CP {
    float2 re, im;  // (lane A, lane B)
};

__device__ __forceinline__ float2 p_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 p_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ float2 p_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 p_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 p_bc(float s) { return make_float2(s, s); }  // becomes an R.F32 broadcast operand
__device__ __forceinline__ float2 p_neg(float2 a) { return make_float2(-a.x, -a.y); }

__device__ __forceinline__ CP cp_add(CP a, CP b) { return {p_add(a.re, b.re), p_add(a.im, b.im)}; }
__device__ __forceinline__ CP cp_sub(CP a, CP b) { return {p_sub(a.re, b.re), p_sub(a.im, b.im)}; }
// multiply both lanes by the same complex scalar w
__device__ __forceinline__ CP cp_mul_s(CP a, float2 w)
{
    CP r;
    r.re = p_fma(a.im, p_bc(-w.y), p_mul(a.re, p_bc(w.x)));
    r.im = p_fma(a.im, p_bc(w.x), p_mul(a.re, p_bc(w.y)));
    return r;
}

__device__ __forceinline__ void cp_radix4_inv(CP &a, CP &b, CP &c, CP &d)
{
    const CP apc = cp_add(a, c), amc = cp_sub(a, c), bpd = cp_add(b, d), bmd = cp_sub(b, d);
    a = cp_add(apc, bpd);
    c = cp_sub(apc, bpd);
    // +j*bmd = (-bmd.im, bmd.re)
    b.re = p_sub(amc.re, bmd.im);
    b.im = p_add(amc.im, bmd.re);
    d.re = p_add(amc.re, bmd.im);
    d.im = p_sub(amc.im, bmd.re);
}

__device__ __forceinline__ void cp_radix16_inv(CP (&x)[16])
{
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, R2 = 0.70710678118654752440f;
#pragma unroll
    for (int a0 = 0; a0 < 4; a0++) cp_radix4_inv(x[a0], x[a0 + 4], x[a0 + 8], x[a0 + 12]);
    x[5] = cp_mul_s(x[5], make_float2(C1, S1));
    x[13] = cp_mul_s(x[13], make_float2(S1, C1));
    x[7] = cp_mul_s(x[7], make_float2(S1, C1));
    x[15] = cp_mul_s(x[15], make_float2(-C1, -S1));
    {   // W16^2 = (1+j)/sqrt2: ((re - im) R2, (re + im) R2)
        CP t = x[9];
        x[9].re = p_mul(p_sub(t.re, t.im), p_bc(R2));
        x[9].im = p_mul(p_add(t.re, t.im), p_bc(R2));
        t = x[6];
        x[6].re = p_mul(p_sub(t.re, t.im), p_bc(R2));
        x[6].im = p_mul(p_add(t.re, t.im), p_bc(R2));
        // W16^6 = (-1+j)/sqrt2: (-(re + im) R2, (re - im) R2)
        t = x[14];
        x[14].re = p_mul(p_add(t.re, t.im), p_bc(-R2));
        x[14].im = p_mul(p_sub(t.re, t.im), p_bc(R2));
        t = x[11];
        x[11].re = p_mul(p_add(t.re, t.im), p_bc(-R2));
        x[11].im = p_mul(p_sub(t.re, t.im), p_bc(R2));
        // W16^4 = j
        t = x[10];
        x[10].re = p_neg(t.im);
        x[10].im = t.re;
    }
#pragma unroll
    for (int nl = 0; nl < 4; nl++) cp_radix4_inv(x[4 * nl], x[4 * nl + 1], x[4 * nl + 2], x[4 * nl + 3]);
}

constexpr int kS1Elems2 = 4096;            // float4 elements
constexpr int kS2Elems2 = 16 * kS2Stride;  // float4 elements

struct FftSmem2 {
    float2 *T1;  // [15][256]
    float2 *T2;  // [4][15][16]
    float4 *S1;  // [16][256]
    float4 *S2;  // [16][272]
};

__host__ __device__ constexpr size_t fft_smem2_bytes()
{
    return sizeof(float2) * (size_t)(kT1Elems + kT2Elems) + sizeof(float4) * (size_t)(kS1Elems2 + kS2Elems2);
}

__device__ __forceinline__ FftSmem2 fft_smem2_carve(unsigned char *base)
{
    FftSmem2 s;
    s.T1 = reinterpret_cast<float2 *>(base);
    s.T2 = s.T1 + kT1Elems;
    s.S1 = reinterpret_cast<float4 *>(s.T2 + kT2Elems);
    s.S2 = s.S1 + kS1Elems2;
    return s;
}

__device__ __forceinline__ float4 cp_pack(CP v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ CP cp_unpack(float4 v) { return {make_float2(v.x, v.y), make_float2(v.z, v.w)}; }

// One 4096-point sub-FFT for both lanes; see subfft4096_inv for the algebra and the hazard argument.
__device__ __forceinline__ void subfft4096_inv2(CP (&x)[16], const int k2, const FftSmem2 &s, const int t)
{
    float2 tw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = s.T1[i * 256 + t];
    cp_radix16_inv(x);
    if (k2 != 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) tw[i] = cmul(tw[i], c_cA[k2][i + 1]);
    }
    s.S1[t] = cp_pack(x[r16(0)]);
#pragma unroll
    for (int i = 0; i < 8; i++) s.S1[(i + 1) * 256 + t] = cp_pack(cp_mul_s(x[r16(i + 1)], tw[i]));
#pragma unroll
    for (int i = 0; i < 7; i++) tw[i] = s.T1[(i + 8) * 256 + t];
    if (k2 != 0) {
#pragma unroll
        for (int i = 0; i < 7; i++) tw[i] = cmul(tw[i], c_cA[k2][i + 9]);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) s.S1[(i + 9) * 256 + t] = cp_pack(cp_mul_s(x[r16(i + 9)], tw[i]));
    const float2 *twp = s.T2 + k2 * (15 * 16) + (t & 15);
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = twp[i * 16];
    __syncthreads();
    {
        const float4 *src = s.S1 + (t & ~15) * 16 + (t & 15);
#pragma unroll
        for (int b = 0; b < 16; b++) x[b] = cp_unpack(src[16 * b]);
    }
    cp_radix16_inv(x);
    {
        float4 *dst = s.S2 + 17 * (t >> 4) + (t & 15);
        dst[0] = cp_pack(x[r16(0)]);
#pragma unroll
        for (int i = 0; i < 8; i++) dst[(i + 1) * kS2Stride] = cp_pack(cp_mul_s(x[r16(i + 1)], tw[i]));
#pragma unroll
        for (int i = 0; i < 7; i++) tw[i] = twp[(i + 8) * 16];
#pragma unroll
        for (int i = 0; i < 7; i++) dst[(i + 9) * kS2Stride] = cp_pack(cp_mul_s(x[r16(i + 9)], tw[i]));
    }
    __syncthreads();
    {
        const float4 *src = s.S2 + (t >> 4) * kS2Stride + 17 * (t & 15);
#pragma unroll
        for (int c = 0; c < 16; c++) x[c] = cp_unpack(src[c]);
    }
    cp_radix16_inv(x);
}

}  // namespace acq
