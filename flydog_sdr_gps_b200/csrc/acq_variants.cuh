// acq_variants.cuh -- A/B forms of the search kernels.  NOT part of the product library: this file is compiled only
// into the experiment variants that tools/build_variants.py builds (-DACQ_VARIANT_L1_LDG, -DACQ_VARIANT_L1_X3,
// -DACQ_VARIANT_E1B_LDG); the tests that assert bitwise equality with the product kernels load such a variant next to
// the product library.  Included by acq_kernels.cu inside namespace acq, after the shared helpers.
#pragma once

#ifdef ACQ_VARIANT_L1_LDG
// k_search_l1_ldg: the C/A search with both operands of a sub-FFT read straight from L2 (the load sits at the head of
// every warp's dependent chain: 7 % slower than the TMA-staged product kernel on cfg2).
template <bool MULTI>
__global__ void __launch_bounds__(256, 2) k_search_l1_ldg(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FftSmem3T s = fft_smem3t_carve(smem);
    float *red_f = reinterpret_cast<float *>(smem + fft_smem3t_bytes());  // [2 parities][16]
    int *red_i = reinterpret_cast<int *>(red_f + 32);                      // [2 parities][8]
    const int t = threadIdx.x;
    constexpr int L = ACQ_LAGS_L1;
    // stage-B twiddles live in tensor memory: 128 columns per thread, warps w and w+4 share lanes
    const uint32_t tmem_base = tmem_alloc_cta<2 * kTwCols>(reinterpret_cast<uint32_t *>(red_f + 48), t);
    const uint32_t tw_taddr = tmem_base + tmem_lane_base(t) + (uint32_t)((t >> 7) * kTwCols);
    subfft3_park_twiddles(p.tables, tw_taddr, t);
    const float2 *base = p.tables + kT2Elems + t;  // [k2][256]: W16384^{4t+k2}
    int buf = 0;
    if (p.wait_prior) pdl_wait();
    pdl_trigger_search();  // after the wait (see k_search_l1)
    // The cross-warp merge of a tile's peak is deferred to the next tile: the warp partials are left in a
    // parity slot and thread 0 merges them after the next tile's first sub-FFT barrier, so the reduction
    // costs no CTA barrier of its own (it matters at K = 1, where a tile is only four sub-FFTs).
    int par = 0, pend_cap = -1, pend_slot = 0, pend_d = 0;
    auto flush = [&](int q) {   // thread 0
        store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * q, red_i + 8 * q), L);
    };

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const TileIdx ti(p, tile);
        float P[MULTI ? 16 : 1];
        float2 acc[16];
        for (int b = 0; b < p.K; b++) {
            float2 x[16];
#pragma unroll 1
            for (int k2 = 0; k2 < 4; k2++) {
                load_products(x, p, ti, b, k2, t);
                subfft4096_inv3t(x, k2, __ldg(base + k2 * 256), buf, s, t, tw_taddr);
                buf ^= 1;
                if (t == 0 && b == 0 && k2 == 0 && pend_cap >= 0) flush(par ^ 1);  // previous tile's peak
                if (k2 == 0) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = cfma(x[r16(n2)], c_cC[k2][n2], acc[n2]);
                }
            }
            if (MULTI) {
                // block b was delayed by 16*b samples in the front end (k_front_end), so lag n lines up across blocks
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) P[n2] = (b == 0) ? cpower(acc[n2]) : (P[n2] + cpower(acc[n2]));
            }
        }
        // power, max, first argmax, sum over lags n < 4092   (search.cpp:486-490); a thread's n grows with n2
        Peak best;
        best.p = 0.0f;
        best.n = 0x7fffffff;
        best.sum = 0.0f;
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const int n = lag_of3(t, n2);
            const float pw = MULTI ? P[n2] : cpower(acc[n2]);
            if (n2 < 15 || n < L) {
                if (pw > best.p) best.p = pw, best.n = n;
                best.sum += pw;
            }
        }
        // parity slot `par` was last read (flush) during the previous tile, before >= 3 CTA barriers
        warp_reduce_peak(best, red_f + 16 * par, red_i + 8 * par, t);
        pend_cap = ti.cap;
        pend_slot = ti.slot;
        pend_d = ti.d;
        par ^= 1;
    }
    __syncthreads();
    if (t == 0 && pend_cap >= 0) flush(par ^ 1);
    search_cta_epilogue(p, t);
    tmem_free_cta<2 * kTwCols>(tmem_base, t);
}

#endif  // ACQ_VARIANT_L1_LDG

#ifdef ACQ_VARIANT_L1_X3
// k_search_l1_x3: the C/A search at THREE CTAs per SM (24 warps instead of 16).  What keeps k_search_l1 at two is its
// register file share (126 registers: 16 points + 16 accumulators + 16 block powers per thread) and its 98.6 KiB of
// shared memory.  Here the accumulators over k2 and the block powers live in thread-private TENSOR MEMORY (a CTA
// allocates 128 columns: 64 per thread -- 32 accumulator, 16 block-power and 8 stage-A-base columns), which brings
// the kernel to 80 registers; the operands come straight from L2 (no staging buffers), the stage-B twiddles from a
// 7.5 KiB shared table, and the B->C tiles live inside the exchange rows: 71.9 KiB of shared memory per CTA.
// Same arithmetic in the same order as k_search_l1 (bitwise-equal cells, tested).
constexpr int kX3AccCol = 0, kX3PowCol = 32, kX3BaseCol = 48, kX3Cols = 64;
__host__ __device__ constexpr size_t search_l1_x3_smem()
{
    return sizeof(float2) * (size_t)(2 * kSub + kT2Elems) + 64 * sizeof(float);
}

template <bool MULTI>
__global__ void __launch_bounds__(256, 3) k_search_l1_x3(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float2 *S1 = reinterpret_cast<float2 *>(smem);                 // [2][4096]
    float2 *T2 = S1 + 2 * kSub;                                    // [4][15][16]
    float *red_f = reinterpret_cast<float *>(T2 + kT2Elems);       // [2 parities][16], then the TMEM slot at [48]
    int *red_i = reinterpret_cast<int *>(red_f + 32);              // [2 parities][8]
    const int t = threadIdx.x;
    constexpr int L = ACQ_LAGS_L1;
    const uint32_t tmem_base = tmem_alloc_cta<2 * kX3Cols>(reinterpret_cast<uint32_t *>(red_f + 48), t);
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.tables);
        float4 *dst = reinterpret_cast<float4 *>(T2);
        for (int i = t; i < kT2Elems / 2; i += 256) dst[i] = __ldg(src + i);
    }
    const uint32_t zaddr = tmem_base + tmem_lane_base(t) + (uint32_t)((t >> 7) * kX3Cols);
    float2 bw = __ldg(p.tables + kT2Elems + t);  // W16384^{4t}: base of residue 0
    tmem_st1(zaddr + kX3BaseCol, bw);
#pragma unroll
    for (int k2 = 1; k2 < 4; k2++) tmem_st1(zaddr + kX3BaseCol + 2 * k2, __ldg(p.tables + kT2Elems + k2 * 256 + t));
    tmem_wait_st();
#ifdef ACQ_X3_TMA_D   // capture residue D staged by TMA one sub-FFT ahead (idle half of S1), code run E from L2 issued before the wait
    const uint32_t bar = smem_u32(red_f + 50);
    if (t == 0) mbar_init(bar, 1);
#endif
    __syncthreads();
    if (p.wait_prior) pdl_wait();
    pdl_trigger_search();  // after the wait (see k_search_l1)
    int it = 0;
    int par = 0, pend_cap = -1, pend_slot = 0, pend_d = 0;
    auto flush = [&](int q) {   // thread 0
        store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * q, red_i + 8 * q), L);
    };
#ifdef ACQ_X3_TMA_D
    auto issue = [&](const TileIdx &tn, int bn, int k2n, int half) {
        const float2 *Dk = p.Dp + d_row(p, tn, bn) * kN + k2n * kSub;
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(sizeof(float2) * kSub));
        tma_load_1d(smem_u32(S1 + half * kSub), Dk, (uint32_t)(sizeof(float2) * kSub), bar);
    };
    if (t == 0 && blockIdx.x < p.n_tiles) issue(TileIdx(p, blockIdx.x), 0, 0, 0);
#endif

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const TileIdx ti(p, tile);
        float pw[16];
#ifdef ACQ_X3_TMA_D
        auto products = [&](float2 (&x)[16], int b, int k2) {
            const int r = (k2 - ti.dop) & 3;
            const int q = (k2 - ti.dop - r) >> 2;
            const float2 *Eg = p.Ep + (size_t)(ti.sat * 4 + r) * p.ext_len + p.Q + q + t;
            const float2 *Dk = S1 + (it & 1) * kSub + t;
#pragma unroll
            for (int a = 0; a < 16; a++) x[a] = __ldg(Eg + 256 * a);
            mbar_wait(bar, (uint32_t)(it & 1));
#pragma unroll
            for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(Dk[256 * a], x[a]);
        };
        auto next = [&](int b, int k2) {   // thread 0, right after the CTA barrier of sub-FFT (b, k2)
            if (t != 0) return;
            if (k2 < 3) issue(ti, b, k2 + 1, (it + 1) & 1);
            else if (b + 1 < p.K) issue(ti, b + 1, 0, (it + 1) & 1);
            else if (tile + gridDim.x < p.n_tiles) issue(TileIdx(p, tile + gridDim.x), 0, 0, (it + 1) & 1);
        };
#else
        auto products = [&](float2 (&x)[16], int b, int k2) { load_products(x, p, ti, b, k2, t); };
        auto next = [&](int, int) {};
#endif
        for (int b = 0; b < p.K; b++) {
            float2 x[16];
#pragma unroll 1
            for (int k2 = 0; k2 < 3; k2++) {
                products(x, b, k2);
                subfft4096_inv4s(x, k2, bw, S1 + (it & 1) * kSub, t, T2, BaseFromTmem{zaddr + kX3BaseCol}, [&] { next(b, k2); });
                it++;
                if (t == 0 && b == 0 && k2 == 0 && pend_cap >= 0) flush(par ^ 1);  // previous tile's peak
                float2 z[16];
                if (k2 == 0) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) z[n2] = x[r16(n2)];
                } else {   // acc += x * W64^{k2 n2}, two halves of eight accumulators through tensor memory
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        float2 a[8];
                        tmem_ld8(zaddr + kX3AccCol + 16 * h, a);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; i++) z[8 * h + i] = cfma(x[r16(8 * h + i)], c_cC[k2][8 * h + i], a[i]);
                    }
                }
                tmem_st16(zaddr + kX3AccCol, z);
                tmem_wait_st();
            }
            // last residue: the accumulation ends in the powers
            products(x, b, 3);
            subfft4096_inv4s(x, 3, bw, S1 + (it & 1) * kSub, t, T2, BaseFromTmem{zaddr + kX3BaseCol}, [&] { next(b, 3); });
            it++;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float2 a[8];
                tmem_ld8(zaddr + kX3AccCol + 16 * h, a);
                float2 pb[4];
                if (MULTI && b > 0) tmem_ld4(zaddr + kX3PowCol + 8 * h, pb);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float v = cpower(cfma(x[r16(8 * h + i)], c_cC[3][8 * h + i], a[i]));
                    if (MULTI && b > 0) pw[8 * h + i] = ((i & 1) ? pb[i >> 1].y : pb[i >> 1].x) + v;
                    else pw[8 * h + i] = v;
                }
            }
            if (MULTI && b + 1 < p.K) {
                float2 ps[8];
#pragma unroll
                for (int i = 0; i < 8; i++) ps[i] = make_float2(pw[2 * i], pw[2 * i + 1]);
                tmem_st8(zaddr + kX3PowCol, ps);
                tmem_wait_st();
            }
        }
        Peak best;
        best.p = 0.0f;
        best.n = 0x7fffffff;
        best.sum = 0.0f;
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const int n = lag_of3(t, n2);
            if (n2 < 15 || n < L) {
                if (pw[n2] > best.p) best.p = pw[n2], best.n = n;
                best.sum += pw[n2];
            }
        }
        warp_reduce_peak(best, red_f + 16 * par, red_i + 8 * par, t);
        pend_cap = ti.cap;
        pend_slot = ti.slot;
        pend_d = ti.d;
        par ^= 1;
    }
    __syncthreads();
    if (t == 0 && pend_cap >= 0) flush(par ^ 1);
    search_cta_epilogue(p, t);
    tmem_free_cta<2 * kX3Cols>(tmem_base, t);
}

#endif  // ACQ_VARIANT_L1_X3

#ifdef ACQ_VARIANT_E1B_LDG
// k_search_e1b_ldg: the one-CTA E1B search with operands read straight from L2 and a reduction barrier per tile.
__global__ void __launch_bounds__(256, 2) k_search_e1b_ldg(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FftSmem3 s = fft_smem3_carve(smem);
    float *red_f = reinterpret_cast<float *>(smem + fft_smem3_bytes());
    int *red_i = reinterpret_cast<int *>(red_f + 16);
    const int t = threadIdx.x;
    constexpr int L = ACQ_LAGS_E1B;
    const uint32_t tmem_base = tmem_alloc_cta<kE1bTmemCols>(reinterpret_cast<uint32_t *>(red_f + 32), t);
    load_t2(s, p.tables, t);
    // this thread's scratch: lane 32*(warp%4) + (t%32), columns [96*(warp/4), +96): [k2][n2] complex
    const uint32_t zaddr = tmem_base + tmem_lane_base(t) + (uint32_t)((t >> 7) * 96);
    const float2 *base = p.tables + kT2Elems + t;
    int buf = 0;
    if (p.wait_prior) pdl_wait();
    pdl_trigger_search();  // after the wait (see k_search_l1)

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const TileIdx ti(p, tile);
        float2 x[16];
#pragma unroll 1
        for (int k2 = 0; k2 < 4; k2++) {
            load_products(x, p, ti, 0, k2, t);
            subfft4096_inv3(x, k2, __ldg(base + k2 * 256), buf, s, t);
            buf ^= 1;
            if (k2 < 3) {
                float2 z[16];
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) z[n2] = (k2 == 0) ? x[r16(n2)] : cmul(x[r16(n2)], c_cC[k2][n2]);
                tmem_st16(zaddr + 32 * k2, z);
                tmem_wait_st();
            }
        }
        // radix-4 combine over k2, lags n = lag_of3(t, n2) + 4096 m < 16368.  Lags are not visited in
        // increasing order here, so ties compare the index explicitly (first index wins, search.cpp:488).
        Peak best;
        best.p = 0.0f;
        best.n = 0x7fffffff;
        best.sum = 0.0f;
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            float2 za[4], zb[4], zc[4];
            tmem_ld4(zaddr + 0 * 32 + 8 * c4, za);
            tmem_ld4(zaddr + 1 * 32 + 8 * c4, zb);
            tmem_ld4(zaddr + 2 * 32 + 8 * c4, zc);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int n2 = 4 * c4 + i;
                float2 z0 = za[i], z1 = zb[i], z2 = zc[i];
                float2 z3 = cmul(x[r16(n2)], c_cC[3][n2]);
                radix4_inv(z0, z1, z2, z3);  // z_m = sum_k2 z_k2 * j^{k2*m}
                const float2 zz[4] = {z0, z1, z2, z3};
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int n = lag_of3(t, n2) + 4096 * m;
                    const float pw = cpower(zz[m]);
                    if (n < L) peak_merge(best, pw, n, pw);
                }
            }
        }
        const Peak tot = block_reduce_peak(best, red_f, red_i, t);
        if (t == 0) store_cell(p, ti.cap, ti.slot, ti.d, tot, L);
    }
    __syncthreads();
    search_cta_epilogue(p, t);
    tmem_free_cta<kE1bTmemCols>(tmem_base, t);
}

#endif  // ACQ_VARIANT_E1B_LDG
