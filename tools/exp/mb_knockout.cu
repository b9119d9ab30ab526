// mb_knockout.cu -- knock-out experiments on the v3 search loop: which resource bounds the kernel?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I flydog_sdr_gps_b200/csrc -o tools/exp/bin/mb_knockout tools/exp/mb_knockout.cu
// Results are wrong by construction for every variant but 0; only the timing matters.
#include <cstdio>
#include <vector>
#include "acq_fft.cuh"
using namespace acq;

enum { NO_LDG = 1, NO_S1 = 2, NO_BAR = 4, NO_S2 = 8, NO_TW = 16, NO_R16 = 32, NO_ACC = 64, LDS_OPS = 128, LDS_D = 256 };

template <int F>
__device__ __forceinline__ void sub3(float2 (&x)[16], const int k2, const float2 b, const int buf, const FftSmem3 &s, const int t)
{
    if (!(F & NO_R16)) radix16_inv(x);
    {
        float2 *dst = s.S1 + buf * kS1Elems + t;
        float2 tw = b;
        if (!(F & NO_S1)) dst[0] = x[r16(0)];
#pragma unroll
        for (int n0 = 1; n0 < 16; n0++) {
            float2 v = (F & NO_TW) ? x[r16(n0)] : cmul(x[r16(n0)], tw);
            if (!(F & NO_S1)) dst[n0 * 256] = v; else x[r16(n0)] = v;
            if (!(F & NO_TW)) if (n0 < 15) tw = cmul(tw, b);
        }
    }
    float2 tw[8];
    const float2 *twp = s.T2 + k2 * (15 * 16) + (t & 15);
    if (!(F & NO_TW)) {
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = twp[i * 16];
    }
    if (!(F & NO_BAR)) __syncthreads();
    if (!(F & NO_S1)) {
        const float2 *src = s.S1 + buf * kS1Elems + (t & ~15) * 16 + (t & 15);
#pragma unroll
        for (int bb = 0; bb < 16; bb++) x[bb] = src[16 * bb];
    }
    if (!(F & NO_R16)) radix16_inv(x);
    float2 *tile = s.S2 + (t >> 4) * kS2TileElems;
    {
        float2 *dst = tile + (t & 15);
        if (!(F & NO_S2)) dst[0] = x[r16(0)];
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 v = (F & NO_TW) ? x[r16(i+1)] : cmul(x[r16(i + 1)], tw[i]); if (!(F & NO_S2)) dst[(i + 1) * 17] = v; else x[r16(i+1)] = v; }
        if (!(F & NO_TW)) {
#pragma unroll
        for (int i = 0; i < 7; i++) tw[i] = twp[(i + 8) * 16];
        }
#pragma unroll
        for (int i = 0; i < 7; i++) { float2 v = (F & NO_TW) ? x[r16(i+9)] : cmul(x[r16(i + 9)], tw[i]); if (!(F & NO_S2)) dst[(i + 9) * 17] = v; else x[r16(i+9)] = v; }
    }
    __syncwarp();
    if (!(F & NO_S2)) {
        const float2 *src = tile + 17 * (t & 15);
#pragma unroll
        for (int c = 0; c < 16; c++) x[c] = src[c];
    }
    if (!(F & NO_R16)) radix16_inv(x);
}

template <int F>
__global__ void __launch_bounds__(256, 2) k3(const float2 *Dp, const float2 *Ep, const float2 *tables, float *out, int n_tiles, int K, int n_dop, int ext_len, int Q)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem3 s = fft_smem3_carve(smem);
    const int t = threadIdx.x;
    for (int i = t; i < kT2Elems; i += 256) s.T2[i] = tables[i];
    __syncthreads();
    const float2 *base = tables + kT2Elems + t;
    int buf = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int d = tile % n_dop, sat = (tile / n_dop) % 32;
        const int dop = d - n_dop / 2;
        float P[16];
        float2 acc[16];
        for (int b = 0; b < K; b++) {
            const float2 *Dblk = Dp + (size_t)b * kN + t;
            float2 x[16];
#pragma unroll 1
            for (int k2 = 0; k2 < 4; k2++) {
                const int r = (k2 - dop) & 3;
                const int q = (k2 - dop - r) >> 2;
                const float2 *Dk = Dblk + k2 * kSub;
                const float2 *Ek = Ep + (size_t)(sat * 4 + r) * ext_len + Q + q + t;
                const float2 bw = __ldg(base + k2 * 256);
                if (F & LDS_OPS) {   // operands as if prefetched into shared memory (garbage data, same access pattern)
                    const float2 *sD = s.S1 + (buf ^ 1) * kS1Elems + t;
                    const float2 *sE = s.S1 + buf * kS1Elems + ((t + q) & 255);
#pragma unroll
                    for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(sD[256 * a], sE[256 * a]);
                } else if (F & LDS_D) {
                    const float2 *sD = s.S1 + (buf ^ 1) * kS1Elems + t;
#pragma unroll
                    for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(sD[256 * a], __ldg(Ek + 256 * a));
                } else if (F & NO_LDG) {
#pragma unroll
                    for (int a = 0; a < 16; a++) x[a] = make_float2(bw.x + a, bw.y - a);
                } else {
#pragma unroll
                    for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(__ldg(Dk + 256 * a), __ldg(Ek + 256 * a));
                }
                sub3<F>(x, k2, bw, buf, s, t);
                buf ^= 1;
                if (F & NO_ACC) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                } else if (k2 == 0) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = cfma(x[r16(n2)], c_cC[k2][n2], acc[n2]);
                }
            }
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) {
                const float2 sq = __fmul2_rn(acc[n2], acc[n2]);
                P[n2] = (b == 0) ? (sq.x + sq.y) : (P[n2] + (sq.x + sq.y));
            }
        }
        float sum = 0;
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) sum += P[n2];
        if (sum == 1.2345f) out[tile] = sum;  // keep the work alive
    }
}

template <int F>
void run(const char *name, const float2 *Dp, const float2 *Ep, const float2 *tab, float *out, int n_tiles, int K, int n_dop, int ext_len, int Q)
{
    const size_t smem = fft_smem3_bytes() + 256;
    cudaFuncSetAttribute(k3<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k3<F><<<296, 256, smem>>>(Dp, Ep, tab, out, n_tiles, K, n_dop, ext_len, Q);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-34s F=%3d  %8.3f ms  %6.2f M tiles/s  %s\n", name, F, best, n_tiles * (double)K / best / 1e3, e ? cudaGetErrorString(e) : "");
}

int main()
{
    const int K = 20, n_dop = 161, n_sats = 32, n_tiles = n_sats * n_dop / 4, Q = 12, ext_len = 4096 + 2 * Q;
    std::vector<float2> hD((size_t)K * kN), hE((size_t)n_sats * 4 * ext_len), hT(kT2Elems + kBaseElems);
    for (auto &v : hD) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : hE) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    for (size_t i = 0; i < hT.size(); i++) { double a = 0.001 * i; hT[i] = make_float2((float)cos(a), (float)sin(a)); }
    float2 *dD, *dE, *dT; float *out;
    cudaMalloc(&dD, hD.size() * 8); cudaMalloc(&dE, hE.size() * 8); cudaMalloc(&dT, hT.size() * 8); cudaMalloc(&out, n_tiles * 4);
    cudaMemcpy(dD, hD.data(), hD.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dE, hE.data(), hE.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dT, hT.data(), hT.size() * 8, cudaMemcpyHostToDevice);
#define RUN(F, name) run<F>(name, dD, dE, dT, out, n_tiles, K, n_dop, ext_len, Q)
    RUN(0, "full");
    RUN(NO_LDG, "no LDG");
    RUN(LDS_OPS, "operands from smem (LDS)");
    RUN(LDS_D, "D from smem, E by LDG");
    RUN(LDS_OPS | NO_BAR, "operands from smem, no barrier");
    RUN(NO_S1, "no S1 exchange");
    RUN(NO_S2, "no S2 exchange");
    RUN(NO_S1 | NO_S2, "no exchanges");
    RUN(NO_BAR, "no barrier");
    RUN(NO_TW, "no twiddles");
    RUN(NO_R16, "no radix16");
    RUN(NO_R16 | NO_TW | NO_ACC, "no FP (LDG+exchange only)");
    RUN(NO_LDG | NO_S1 | NO_S2 | NO_BAR, "FP only");
    RUN(NO_LDG | NO_S1 | NO_S2, "FP + barrier");
    RUN(NO_LDG | NO_BAR, "no LDG no barrier");
    RUN(NO_S1 | NO_S2 | NO_BAR, "LDG + FP");
    return 0;
}
