// Experiment: throughput of a register radix-16 butterfly + 15 twiddle multiplies,
// scalar float2 (one butterfly per thread) vs packed f32x2 SoA (two butterflies per thread).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }
__device__ __forceinline__ void radix4(float2 &a, float2 &b, float2 &c, float2 &d)
{
    const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), j = mul_pj(csub(b, d));
    a = cadd(apc, bpd); b = cadd(amc, j); c = csub(apc, bpd); d = csub(amc, j);
}
__device__ __forceinline__ void radix16(float2 (&x)[16])
{
    constexpr float C1 = 0.92387953f, S1 = 0.38268343f, R2 = 0.70710678f;
#pragma unroll
    for (int a0 = 0; a0 < 4; a0++) radix4(x[a0], x[a0 + 4], x[a0 + 8], x[a0 + 12]);
    x[5] = cmul(x[5], make_float2(C1, S1));
    x[9] = make_float2((x[9].x - x[9].y) * R2, (x[9].x + x[9].y) * R2);
    x[13] = cmul(x[13], make_float2(S1, C1));
    x[6] = make_float2((x[6].x - x[6].y) * R2, (x[6].x + x[6].y) * R2);
    x[10] = mul_pj(x[10]);
    x[14] = make_float2(-(x[14].x + x[14].y) * R2, (x[14].x - x[14].y) * R2);
    x[7] = cmul(x[7], make_float2(S1, C1));
    x[11] = make_float2(-(x[11].x + x[11].y) * R2, (x[11].x - x[11].y) * R2);
    x[15] = cmul(x[15], make_float2(-C1, -S1));
#pragma unroll
    for (int nl = 0; nl < 4; nl++) radix4(x[4 * nl], x[4 * nl + 1], x[4 * nl + 2], x[4 * nl + 3]);
}

// ---- packed: P = (lane A, lane B); complex value = (re: P, im: P)
typedef float2 P;
__device__ __forceinline__ P padd(P a, P b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ P psub(P a, P b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ P pmul(P a, P b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ P pfma(P a, P b, P c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ P pneg(P a) { return make_float2(-a.x, -a.y); }
struct CP { P re, im; };
__device__ __forceinline__ CP cpadd(CP a, CP b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ CP cpsub(CP a, CP b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
__device__ __forceinline__ CP cpmulc(CP a, float wr, float wi)  // constant twiddle (same for both lanes)
{
    const P WR = make_float2(wr, wr), WI = make_float2(wi, wi), NWI = make_float2(-wi, -wi);
    return {pfma(a.im, NWI, pmul(a.re, WR)), pfma(a.im, WR, pmul(a.re, WI))};
}
__device__ __forceinline__ CP cpmul(CP a, CP w)
{
    return {pfma(pneg(a.im), w.im, pmul(a.re, w.re)), pfma(a.im, w.re, pmul(a.re, w.im))};
}
__device__ __forceinline__ void pradix4(CP &a, CP &b, CP &c, CP &d)
{
    const CP apc = cpadd(a, c), amc = cpsub(a, c), bpd = cpadd(b, d), bmd = cpsub(b, d);
    // j*bmd = (-bmd.im, bmd.re)
    a = cpadd(apc, bpd);
    c = cpsub(apc, bpd);
    b = {psub(amc.re, bmd.im), padd(amc.im, bmd.re)};
    d = {padd(amc.re, bmd.im), psub(amc.im, bmd.re)};
}
__device__ __forceinline__ void pradix16(CP (&x)[16])
{
    constexpr float C1 = 0.92387953f, S1 = 0.38268343f, R2 = 0.70710678f;
#pragma unroll
    for (int a0 = 0; a0 < 4; a0++) pradix4(x[a0], x[a0 + 4], x[a0 + 8], x[a0 + 12]);
    x[5] = cpmulc(x[5], C1, S1);
    x[9] = cpmulc(x[9], R2, R2);
    x[13] = cpmulc(x[13], S1, C1);
    x[6] = cpmulc(x[6], R2, R2);
    x[10] = {pneg(x[10].im), x[10].re};
    x[14] = cpmulc(x[14], -R2, R2);
    x[7] = cpmulc(x[7], S1, C1);
    x[11] = cpmulc(x[11], -R2, R2);
    x[15] = cpmulc(x[15], -C1, -S1);
#pragma unroll
    for (int nl = 0; nl < 4; nl++) pradix4(x[4 * nl], x[4 * nl + 1], x[4 * nl + 2], x[4 * nl + 3]);
}

constexpr int ITERS = 512;

__global__ void __launch_bounds__(256) k_scalar(float2 *out, const float2 *tw)
{
    float2 x[16], w[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 15; i++) w[i] = tw[(threadIdx.x + i) & 255];
    for (int it = 0; it < ITERS; it++) {
        radix16(x);
#pragma unroll
        for (int i = 1; i < 16; i++) x[i] = cmul(x[i], w[i - 1]);
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; i++) s = cadd(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_packed(float2 *out, const float2 *tw)
{
    CP x[16], w[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = {make_float2(threadIdx.x * 0.001f + i, i * 0.25f), make_float2(i * 0.5f, 1.f)};
#pragma unroll
    for (int i = 0; i < 15; i++) {
        const float2 a = tw[(threadIdx.x + i) & 255], b = tw[(threadIdx.x + i + 7) & 255];
        w[i] = {make_float2(a.x, b.x), make_float2(a.y, b.y)};
    }
    for (int it = 0; it < ITERS; it++) {
        pradix16(x);
#pragma unroll
        for (int i = 1; i < 16; i++) x[i] = cpmul(x[i], w[i - 1]);
    }
    P s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; i++) s = padd(s, padd(x[i].re, x[i].im));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float2 *out, *tw; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256); cudaMalloc(&tw, sizeof(float2) * 256);
    float2 h[256]; for (int i = 0; i < 256; i++) h[i] = make_float2(cosf(0.01f * i), sinf(0.01f * i));
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        for (int mode = 0; mode < 2; mode++) {
            float best = 1e30f;
            for (int r = 0; r < 4; r++) {
                cudaEventRecord(e0);
                if (mode == 0) k_scalar<<<sms * ctas, 256>>>(out, tw); else k_packed<<<sms * ctas, 256>>>(out, tw);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double bfly = (double)sms * ctas * 256 * ITERS * (mode ? 2 : 1);
            printf("%s warps/SM=%2d  %.3f ms  %.2f G butterflies(16pt+15tw)/s  -> %.1f M 4096-subFFT-stage-equiv/s\n",
                   mode ? "packed" : "scalar", ctas * 8, best, bfly / best / 1e6, bfly / best / 1e3 / 256);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
