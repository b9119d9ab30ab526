// acq_microbench.cu -- on-device micro-benchmarks that give the on-SM roofline denominators
// ("shared-memory and FP32 peaks are to be measured by micro-benchmark in the same
// run").  The acquisition kernel is bound by FP32 issue and shared-memory bandwidth, not HBM, so
// MEASURED_PEAKS.json (HBM copy, bf16 GEMM) has no usable denominator for it.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/acq_b200.h"

namespace {

constexpr int kIters = 4096;

// 16 independent FFMA chains per thread (register operands).
__global__ void __launch_bounds__(256) mb_ffma(float *out, float a, float b)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// SM clock: one warp counts clock64 cycles over 200 us of %globaltimer, launched right behind an FFMA run on the
// same stream (the clock the load left behind).  Kept out of mb_ffma: timer reads inside it cost 5 % of its rate.
__global__ void mb_clock(long long *cycles)
{
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    } while (g1 - g0 < 200000ull);
    const long long c1 = clock64();
    cycles[0] = c1 - c0;
    cycles[1] = (long long)(g1 - g0);
}

// 16 independent packed FFMA2 chains per thread (sm_100 fma.rn.f32x2).
__global__ void __launch_bounds__(256) mb_ffma2(float *out, float a, float b)
{
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = __ffma2_rn(acc[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 independent FADD chains (the FFT butterflies are add-dominated).
__global__ void __launch_bounds__(256) mb_fadd(float *out, float b)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = __fadd_rn(acc[i], b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mb_fadd2(float *out, float b)
{
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
    const float2 b2 = make_float2(b, b);
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = __fadd2_rn(acc[i], b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Shared memory: conflict-free 8-byte loads and stores (the exchange pattern of the FFT).
constexpr int kSmemIters = 2048;
__global__ void __launch_bounds__(256) mb_smem(float *out)
{
    extern __shared__ float2 sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 8192; i += 256) sm[i] = make_float2((float)i, 1.0f);
    __syncthreads();
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = make_float2(0.f, 0.f);
    for (int it = 0; it < kSmemIters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 r = sm[((it + j) & 31) * 256 + t];
            v[j].x += r.x;
            v[j].y += r.y;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) sm[((it + j + 7) & 31) * 256 + t] = v[j];
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += v[j].x + v[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// L2 -> SM: grid-stride sweeps over a 16 MiB window (L2 resident after the first sweep), 16-byte loads.
__global__ void __launch_bounds__(256) mb_l2(const float4 *__restrict__ buf, size_t n4, int reps, float *out)
{
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "l"(buf + i));
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F>
float time_ms(F launch, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace

extern "C" int acq_microbench(int device, double *out, int n_out)
{
    if (!out || n_out < 5) return ACQ_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return ACQ_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int sms = prop.multiProcessorCount;
    const int grid = sms * 8;  // 8 CTAs x 256 threads = full occupancy
    float *d_out = nullptr;
    long long *d_cyc = nullptr;
    cudaMalloc(&d_out, sizeof(float) * (size_t)grid * 256);
    cudaMalloc(&d_cyc, 2 * sizeof(long long));
    for (int i = 0; i < n_out; i++) out[i] = 0.0;

    const double n_thr = (double)grid * 256;
    float ms = time_ms([&] { mb_ffma<<<grid, 256>>>(d_out, 1.0001f, 0.5f); }, 5);
    out[0] = n_thr * 16.0 * kIters * 2.0 / (ms * 1e-3) / 1e12;
    mb_ffma<<<grid, 256>>>(d_out, 1.0001f, 0.5f);
    mb_clock<<<1, 32>>>(d_cyc);
    long long cyc[2] = {0, 0};
    cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost);
    // SM clock under load: clock64 cycles over globaltimer nanoseconds of the same loop in CTA 0
    out[4] = cyc[1] > 0 ? (double)cyc[0] / (double)cyc[1] * 1e3 : 0.0;
    ms = time_ms([&] { mb_ffma2<<<grid, 256>>>(d_out, 1.0001f, 0.5f); }, 5);
    out[1] = n_thr * 16.0 * kIters * 4.0 / (ms * 1e-3) / 1e12;
    if (n_out > 5) {
        ms = time_ms([&] { mb_fadd<<<grid, 256>>>(d_out, 0.5f); }, 5);
        out[5] = n_thr * 16.0 * kIters / (ms * 1e-3) / 1e12;  // T adds/s
    }
    if (n_out > 6) {
        ms = time_ms([&] { mb_fadd2<<<grid, 256>>>(d_out, 0.5f); }, 5);
        out[6] = n_thr * 16.0 * kIters * 2.0 / (ms * 1e-3) / 1e12;
    }
    {
        const int g2 = sms * 3;  // 3 x 64 KiB per SM
        cudaFuncSetAttribute(mb_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        ms = time_ms([&] { mb_smem<<<g2, 256, 65536>>>(d_out); }, 5);
        out[2] = (double)g2 * 256 * (double)kSmemIters * 16.0 * 8.0 / (ms * 1e-3) / 1e12;
    }
    {
        const size_t bytes = 16u << 20;
        float4 *buf = nullptr;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 0, bytes);
        const int reps = 256;
        ms = time_ms([&] { mb_l2<<<grid, 256>>>(buf, bytes / 16, reps, d_out); }, 5);
        out[3] = (double)bytes * reps / (ms * 1e-3) / 1e12;
        cudaFree(buf);
    }
    cudaFree(d_out);
    cudaFree(d_cyc);
    const cudaError_t e = cudaDeviceSynchronize();
    if (prev >= 0) cudaSetDevice(prev);
    return e == cudaSuccess ? ACQ_OK : ACQ_ERR_CUDA;
}
