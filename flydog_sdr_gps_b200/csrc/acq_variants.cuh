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

// k_search_l1_sp -- the K = 1 C/A search as a SOFTWARE-PIPELINED chain of sub-FFTs, one 256-thread CTA per SM with the
// whole register file (255 registers per thread).  k_search_l1 runs a sub-FFT as "operand loads -> products ->
// stage A | CTA barrier | stage B -> stage C -> accumulate": between two barriers every warp of a CTA is in the same
// phase, the L1/shared data pipe and the FMA pipe take turns, and only the second CTA of the SM fills the gaps (one CTA
// alone needs LSU time + FMA time per sub-FFT: no overlap at all, ncu).  Here the interval between two CTA barriers holds
// the SECOND half of sub-FFT j (stage B, stage C, accumulation) and the FIRST half of sub-FFT j + 1 (products, stage A)
// as one branch-free instruction stream per thread, the loads of both issued at its head, so a warp's own FP work on one
// sub-FFT covers its own shared-memory round trips of the other -- overlap inside every warp instead of between CTAs.
// That needs the operands of j + 1 at the START of interval j, one interval earlier than k_search_l1 stages them: the
// exchange buffer has three slots (D of j + 2 lands in the slot sub-FFT j - 1 has left) and the code run two, 160 KiB of
// shared memory.  Same arithmetic in the same order as k_search_l1<false>: bitwise-equal cells (tested).
#ifdef ACQ_VARIANT_L1_SP
constexpr int kSpSlots = 4;   // exchange slots (a sub-FFT's operands are staged three intervals ahead)
struct SpSmem {
    float2 *S1;  // [kSpSlots][16][256]
    float2 *E;   // [2][4098]
    unsigned long long *bar;  // [2]
    float *red_f;
    int *red_i;
};
__host__ __device__ constexpr size_t sp_smem_bytes()
{
    return sizeof(float2) * (size_t)(kSpSlots * kS1pElems + 2 * kEBufElems) + 16 + 64 * sizeof(float);
}
__device__ __forceinline__ SpSmem sp_smem_carve(unsigned char *smem)
{
    SpSmem m;
    m.S1 = reinterpret_cast<float2 *>(smem);
    m.E = m.S1 + kSpSlots * kS1pElems;
    m.bar = reinterpret_cast<unsigned long long *>(m.E + 2 * kEBufElems);
    m.red_f = reinterpret_cast<float *>(m.bar + 2);
    m.red_i = reinterpret_cast<int *>(m.red_f + 32);
    return m;
}

// Three warpgroups: two of FFT warps (232 registers per thread after setmaxnreg) and one whose first warp stages the
// operands and stores the cells (40 registers; its other three warps only hold the warpgroup together).  A ninth warp
// with the FFT warps' register count would not fit a scheduler's quarter of the register file.
constexpr int kSpThreads = 384, kSpBarThreads = 288;
__device__ __forceinline__ void sp_cta_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSpBarThreads) : "memory"); }

// Interval j (between CTA barriers j and j + 1) of an FFT warp, with x_{j+1} = the products of sub-FFT j + 1 in registers:
//   gather of stage B (j)  |  stage A (j+1): radix-16, twiddles, exchange stores  |  operand loads of j + 2
//   stage B (j): radix-16, twiddles, tile stores  |  tile loads  |  products of j + 2  |  stage C (j), accumulation
// so every shared-memory round trip of one sub-FFT is covered by FP work of another that needs no load.  The operands
// of sub-FFT i are staged three intervals ahead: four exchange slots, two code-run slots.
struct SpCtx {
    SpSmem m;
    uint32_t bar0;
    int sd, sw, sc;
    __device__ __forceinline__ float2 *slot(int i) const { return m.S1 + (i & 3) * kS1pElems; }   // exchange slot of sub-FFT i
};

// The staging warp (lane 0 works): after CTA barrier j it stages the operands of sub-FFT j + 3 (D into the exchange slot
// sub-FFT j - 1 has left, E into the slot sub-FFT j + 1's code run was read from) and, once per tile, merges the warp
// partials of the finished tile and stores its cell.  None of this sits on an FFT warp's path to a barrier.
__device__ __forceinline__ void sp_stage_warp(const SearchArgs &p, const SpCtx &cx, const bool lead)
{
    constexpr int L = ACQ_LAGS_L1;
    float *red_f = cx.m.red_f;
    int *red_i = cx.m.red_i;
    TileIdx ti(p, blockIdx.x);
    auto issue = [&](const TileIdx &tn, int k2n, int i) {
        const int r = (k2n - tn.dop) & 3;
        const int q = (k2n - tn.dop - r) >> 2;
        const float2 *Dk = p.Dp + d_row(p, tn, 0) * kN + k2n * kSub;
        const float2 *Ek = p.Ep + (size_t)(tn.sat * 4 + r) * p.ext_len + ((p.Q + q) & ~1);
        const uint32_t bar = cx.bar0 + 8u * (i & 1);
        fence_proxy_async();  // generic-proxy accesses of these buffers (ordered by the CTA barrier) before the async writes
        mbar_expect_tx(bar, (uint32_t)(sizeof(float2) * (kSub + kEBufElems)));
        tma_load_1d(smem_u32(cx.m.E + (i & 1) * kEBufElems), Ek, (uint32_t)(sizeof(float2) * kEBufElems), bar);
        tma_load_1d(smem_u32(cx.slot(i)), Dk, (uint32_t)(sizeof(float2) * kSub), bar);
    };
    int j = 0, par = 0, pend_cap = -1, pend_slot = 0, pend_d = 0;
    if (lead) {
        issue(ti, 0, 0);
        issue(ti, 1, 1);
    }
    __syncwarp();
    sp_cta_barrier();   // P: the FFT warps have read the operands of sub-FFT 0 (E slot 0 is free)
    if (lead) issue(ti, 2, 2);
    __syncwarp();
    sp_cta_barrier();   // barrier 0
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        TileIdx tn = ti;
        if (tile + gridDim.x < p.n_tiles) tn.step(p, cx.sd, cx.sw, cx.sc);
#pragma unroll 1
        for (int k2 = 0; k2 < 4; k2++) {
            if (lead) {
                issue((k2 < 1) ? ti : tn, (k2 + 3) & 3, j + 3);   // sub-FFT j + 3
                if (k2 == 0 && pend_cap >= 0)                     // the previous tile's peak (its partials precede the barrier)
                    store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
            }
            __syncwarp();
            sp_cta_barrier();
            j++;
        }
        pend_cap = ti.cap;
        pend_slot = ti.slot;
        pend_d = ti.d;
        par ^= 1;
        ti = tn;
    }
    if (lead) {
        if (pend_cap >= 0)
            store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
        mbar_wait(cx.bar0 + 8u * ((j + 2) & 1), (uint32_t)(((j + 2) >> 1) & 1));   // the last staged operands (never used) must have landed
        if (p.ctas_total) {
            __threadfence();  // this CTA's cells (all stored by this thread) before the count
            atomicAdd(p.ctas_done, 1u);
        }
    }
    __syncwarp();
}

// The eight FFT warps.
__device__ __forceinline__ void sp_fft_warps(const SearchArgs &p, const SpCtx &cx, const int t, const uint32_t tw_taddr)
{
    float *red_f = cx.m.red_f;
    int *red_i = cx.m.red_i;
    const uint32_t bar0 = cx.bar0;
    subfft4_park_twiddles(p.tables, tw_taddr, t);
    TileIdx ti(p, blockIdx.x);
    auto e_of = [&](const TileIdx &tn, int k2n, int i) {   // this thread's column of the staged code run of sub-FFT i
        const int r = (k2n - tn.dop) & 3;
        const int q = (k2n - tn.dop - r) >> 2;
        return cx.m.E + (i & 1) * kEBufElems + ((p.Q + q) & 1) + t;
    };
    int j = 0, par = 0;
    float2 xa[16];
    {   // prologue: first half of sub-FFT 0, products of sub-FFT 1
        const float2 bw = __ldg(p.tables + kT2Elems + t);  // W16384^{4t}: base of residue 0; later bases come from TMEM
        const float2 *Dk = cx.slot(0) + t, *Ek = e_of(ti, 0, 0);
        mbar_wait(bar0, 0);
#pragma unroll
        for (int a = 0; a < 16; a++) xa[a] = cmul_conj_a(Dk[kRowElems * a], Ek[256 * a]);
        sp_cta_barrier();   // P
        radix16_inv(xa);
        stage_a_store<kRowElems>(xa, bw, cx.slot(0) + t);
        const float2 *Dk1 = cx.slot(1) + t, *Ek1 = e_of(ti, 1, 1);
        mbar_wait(bar0 + 8, 0);
#pragma unroll
        for (int a = 0; a < 16; a++) xa[a] = cmul_conj_a(Dk1[kRowElems * a], Ek1[256 * a]);
    }
    sp_cta_barrier();   // barrier 0
    float2 acc[16];
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // the tile after this one (a CTA's last tile stages and half-transforms itself once more: branch-free intervals)
        TileIdx tn = ti;
        if (tile + gridDim.x < p.n_tiles) tn.step(p, cx.sd, cx.sw, cx.sc);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) {
            // ---- interval j: second half of sub-FFT j = (ti, k2), stage A of sub-FFT j + 1 (products in xa), products of j + 2
            const TileIdx &t2 = (k2 < 2) ? ti : tn;   // tile of sub-FFT j + 2
            float2 *Sc = cx.slot(j), *Sn = cx.slot(j + 1);
            float2 tw[8], tw2[8];
            tmem_ld8(tw_taddr + 32 * k2, tw);        // stage-B twiddles n1 = 1..8 of residue k2
            tmem_ld8(tw_taddr + 32 * k2 + 16, tw2);  // n1 = 9..15, then the stage-A base of residue k2 + 1
#define TW_AT(i) ((i) < 8 ? tw[(i)] : tw2[(i) - 8])
            mbar_wait(bar0 + 8u * (j & 1), (uint32_t)(((j + 2) >> 1) & 1));   // operands of sub-FFT j + 2
            float2 *row = Sc + (t >> 4) * kRowElems;  // row n0 = t >> 4 of sub-FFT j's exchange
            const int c = t & 15;
            float2 xb[16], d[16], e[16];
            {
                const float2 *src = row + c;
#pragma unroll
                for (int bb = 0; bb < 16; bb++) xb[bb] = src[16 * bb];
            }
            radix16_inv(xa);                                        // stage A of j + 1: needs no load
            tmem_wait_ld();
            stage_a_store<kRowElems>(xa, TW_AT(15), Sn + t);       // sixteenth twiddle slot of residue k2: the base of residue k2 + 1
            {
                const float2 *Dk = cx.slot(j + 2) + t, *Ek = e_of(t2, (k2 + 2) & 3, j + 2);
#pragma unroll
                for (int a = 0; a < 16; a++) d[a] = Dk[kRowElems * a];
#pragma unroll
                for (int a = 0; a < 16; a++) e[a] = Ek[256 * a];
            }
            radix16_inv(xb);                                        // stage B of j
            __syncwarp();  // the half-warp has consumed its row: reuse it as the B->C tile (see subfft4096_inv4)
            {
                const uint32_t wb = smem_u32(row + c);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(wb), "f"(xb[r16(0)].x), "f"(xb[r16(0)].y) : "memory");
#pragma unroll
                for (int i = 0; i < 15; i++) {
                    const int n1 = i + 1;
                    const float2 v = cmul(xb[r16(n1)], TW_AT(i));
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((wb ^ (16u * (n1 & 7))) + 128u * n1), "f"(v.x), "f"(v.y) : "memory");
                }
            }
            __syncwarp();
            {
                const uint32_t rb = smem_u32(row + 16 * c + 2 * (c & 7));
#pragma unroll
                for (int jj = 0; jj < 8; jj++)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(xb[2 * jj].x), "=f"(xb[2 * jj].y), "=f"(xb[2 * jj + 1].x), "=f"(xb[2 * jj + 1].y)
                                 : "r"(rb ^ (16u * jj))
                                 : "memory");
            }
#pragma unroll
            for (int a = 0; a < 16; a++) xa[a] = cmul_conj_a(d[a], e[a]);   // products of j + 2 (search.cpp:471): cover the tile loads
            radix16_inv(xb);                                        // stage C of j
            if (k2 == 0) {
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) acc[n2] = xb[r16(n2)];
            } else {
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) acc[n2] = cfma(xb[r16(n2)], c_cC[k2][n2], acc[n2]);
            }
#undef TW_AT
            if (k2 == 3) {
                float P[16];
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) P[n2] = cpower(acc[n2]);
                warp_reduce_peak_redux(thread_peak_l1(P, t), red_f + 16 * par, red_i + 8 * par, t);
                par ^= 1;
            }
            sp_cta_barrier();
            j++;
        }
        ti = tn;
    }
}

__global__ void __launch_bounds__(kSpThreads, 1) k_search_l1_sp(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    SpCtx cx;
    cx.m = sp_smem_carve(smem);
    const int t = threadIdx.x;
    ACQ_TRACE_STAMP(kTrSearchL1, 0);
    const uint32_t tmem_base = tmem_alloc_cta<2 * kTwCols>(reinterpret_cast<uint32_t *>(cx.m.red_f + 48), t);
    cx.bar0 = smem_u32(cx.m.bar);
    if (t == 0) {
        mbar_init(cx.bar0, 1);
        mbar_init(cx.bar0 + 8, 1);
    }
    __syncthreads();
    if (p.wait_prior) pdl_wait();
    ACQ_TRACE_STAMP(kTrSearchL1, 1);
    pdl_trigger_search();
    cx.sd = (int)(gridDim.x % (unsigned)p.n_dop);
    cx.sw = (int)((gridDim.x / (unsigned)p.n_dop) % (unsigned)p.n_work);
    cx.sc = (int)(gridDim.x / ((unsigned)p.n_dop * (unsigned)p.n_work));
    if (blockIdx.x < p.n_tiles) {   // (the launch never has more CTAs than tiles)
        if (t >= 256) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (t < kSpBarThreads) sp_stage_warp(p, cx, t == 256);
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
            sp_fft_warps(p, cx, t, tmem_base + tmem_lane_base(t) + (uint32_t)((t >> 7) * kTwCols));
        }
    }
    ACQ_TRACE_STAMP(kTrSearchL1, 2);
    tmem_free_cta<2 * kTwCols>(tmem_base, t);
}
#endif  // ACQ_VARIANT_L1_SP

// k_search_l1_st -- k_search_l1<false> with STAGING WARPS.  In k_search_l1 thread 0 of a CTA issues the next sub-FFT's
// bulk copies right after the CTA barrier (address arithmetic, proxy fence, expect_tx, two copies) and once per tile merges
// the warp partials and stores the cell: ncu's warp samples put a third of a sub-FFT period of warp 0 on that path, and the
// other seven warps wait for it at the next barrier (10 % of all warp samples are barrier stalls).  Here one CTA per SM
// holds TWO teams of eight FFT warps -- each team is what a CTA of k_search_l1 is, with its own exchange and code-run
// buffers, mbarrier, tensor-memory columns and named barrier -- plus one staging warp per team that takes part in the
// team's barrier and does all of the above, so no FFT warp ever leaves the common instruction stream.  Five warpgroups:
// the four of FFT warps raise their register count to 120 (setmaxnreg), the staging warpgroup drops to 24.
// Same arithmetic in the same order as k_search_l1<false>: bitwise-equal cells (tested).
#ifdef ACQ_VARIANT_L1_ST
__host__ __device__ constexpr size_t st_team_smem_bytes() { return (fft_smem4_bytes() + 64 * sizeof(float) + 1023) / 1024 * 1024; }
__host__ __device__ constexpr size_t st_smem_bytes() { return 2 * st_team_smem_bytes(); }

__global__ void __launch_bounds__(kStThreads, 1) k_search_l1_st(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x;
    ACQ_TRACE_STAMP(kTrSearchL1, 0);
    constexpr int L = ACQ_LAGS_L1;
    __shared__ uint32_t tmem_slot;
    const uint32_t tmem_base = tmem_alloc_cta<4 * kTwCols>(&tmem_slot, t);
    if (t < 2) mbar_init(smem_u32(l1_smem_carve(smem + t * st_team_smem_bytes()).s.bar), 1);
    __syncthreads();
    if (p.wait_prior) pdl_wait();
    ACQ_TRACE_STAMP(kTrSearchL1, 1);
    pdl_trigger_search();
    // a team strides over the tiles like a CTA of k_search_l1; team 1 takes the tiles after team 0's first ones, so a search of
    // no more tiles than SMs runs one team per SM
    const int team = (t < 512) ? (t >> 8) : ((t >> 5) & 1);
    const unsigned vcta = (unsigned)team * gridDim.x + blockIdx.x, stride = 2u * gridDim.x;
    const L1Smem m = l1_smem_carve(smem + team * st_team_smem_bytes());
    const FftSmem4 &s = m.s;
    const uint32_t bar = smem_u32(s.bar);
    const int sd = (int)(stride % (unsigned)p.n_dop), sw = (int)((stride / (unsigned)p.n_dop) % (unsigned)p.n_work),
              sc = (int)(stride / ((unsigned)p.n_dop * (unsigned)p.n_work));
    if (t >= 512) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (t < 576 && vcta < p.n_tiles) {
            // ---- staging warp of `team` (lane 0 works)
            const bool lead = (t & 31) == 0;
            float *red_f = m.red_f;
            int *red_i = m.red_i;
            auto issue = [&](const TileIdx &tn, int k2n, int half) {
                const int r = (k2n - tn.dop) & 3;
                const int q = (k2n - tn.dop - r) >> 2;
                const float2 *Dk = p.Dp + d_row(p, tn, 0) * kN + k2n * kSub;
                const float2 *Ek = p.Ep + (size_t)(tn.sat * 4 + r) * p.ext_len + ((p.Q + q) & ~1);
                fence_proxy_async();  // generic-proxy accesses of these buffers (ordered by the team barrier) before the async writes
                mbar_expect_tx(bar, (uint32_t)(sizeof(float2) * (kSub + kEBufElems)));
                tma_load_1d(smem_u32(s.E), Ek, (uint32_t)(sizeof(float2) * kEBufElems), bar);
                tma_load_1d(smem_u32(s.S1 + half * kS1pElems), Dk, (uint32_t)(sizeof(float2) * kSub), bar);
            };
            TileIdx ti(p, vcta);
            if (lead) issue(ti, 0, 0);
            int it = 0, par = 0, pend_cap = -1, pend_slot = 0, pend_d = 0;
            for (long long tile = vcta; tile < p.n_tiles; tile += stride) {
                TileIdx tn = ti;
                const bool more = tile + stride < p.n_tiles;
                if (more) tn.step(p, sd, sw, sc);
#pragma unroll 1
                for (int k2 = 0; k2 < 4; k2++) {
                    __syncwarp();
                    st_stage_wait(team);   // every FFT warp is past its operand reads of this sub-FFT and past stage C of the previous one
                    if (lead) {
                        if (k2 < 3) issue(ti, k2 + 1, (it + 1) & 1);
                        else if (more) issue(tn, 0, (it + 1) & 1);
                        if (k2 == 0 && pend_cap >= 0)   // the previous tile's peak: its warp partials precede this barrier
                            store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
                    }
                    it++;
                }
                pend_cap = ti.cap;
                pend_slot = ti.slot;
                pend_d = ti.d;
                par ^= 1;
                ti = tn;
            }
            __syncwarp();
            st_stage_wait_done(team);
            if (lead) {
                store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
                __threadfence();  // this team's cells (all stored by this thread) before the CTA's count
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        if (vcta < p.n_tiles) {
            // ---- the eight FFT warps of `team`
            const int tt = t & 255;
            float *red_f = m.red_f;
            int *red_i = m.red_i;
            const uint32_t tw_taddr = tmem_base + (uint32_t)(team * 2 * kTwCols) + tmem_lane_base(tt) + (uint32_t)((tt >> 7) * kTwCols);
            subfft4_park_twiddles(p.tables, tw_taddr, tt);
            float2 bw = __ldg(p.tables + kT2Elems + tt);  // W16384^{4t}: base of residue 0; later bases come from TMEM
            int it = 0, par = 0;
            int dop = TileIdx(p, vcta).dop, d = TileIdx(p, vcta).d;
            for (long long tile = vcta; tile < p.n_tiles; tile += stride) {
                float P[16];
                float2 acc[16];
                float2 x[16];
#pragma unroll kK2Unroll
                for (int k2 = 0; k2 < 4; k2++) {
                    float2 *S1b = s.S1 + (it & 1) * kS1pElems;
                    {   // x[a] = conj(data[k]) * code[k - dop], k = 1024 a + 4 t + k2   (search.cpp:471), operands from smem
                        const int r = (k2 - dop) & 3;
                        const int q = (k2 - dop - r) >> 2;
                        const float2 *Dk = S1b + tt;
                        const float2 *Ek = s.E + ((p.Q + q) & 1) + tt;
                        mbar_wait(bar, (uint32_t)(it & 1));
#pragma unroll
                        for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(Dk[kRowElems * a], Ek[256 * a]);
                    }
                    subfft4096_inv4<true>(x, k2, bw, S1b, tt, tw_taddr, [&] { st_fft_arrive(team); st_fft_sync(team); }, [] {});
                    it++;
                    if (k2 == 0) {
#pragma unroll
                        for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                    } else {
#pragma unroll
                        for (int n2 = 0; n2 < 16; n2++) acc[n2] = cfma(x[r16(n2)], c_cC[k2][n2], acc[n2]);
                    }
                }
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) P[n2] = cpower(acc[n2]);
                warp_reduce_peak_redux(thread_peak_l1(P, tt), red_f + 16 * par, red_i + 8 * par, tt);
                par ^= 1;
                {   // the Doppler index of this team's next tile (the only tile coordinate an FFT warp needs): d advances by sd mod n_dop
                    d += sd;
                    if (d >= p.n_dop) d -= p.n_dop;
                    const int h = p.dop_lo + d;
                    const int v = p.half_bin ? (h & 1) : 0;
                    dop = p.half_bin ? ((h - v) >> 1) : h;
                }
            }
            st_fft_done(team);   // the last tile's warp partials are in place
        }
    }
    __syncthreads();
    if (t == 0 && p.ctas_total) atomicAdd(p.ctas_done, 1u);   // both teams' cells are stored and fenced (staging warps, above)
    ACQ_TRACE_STAMP(kTrSearchL1, 2);
    tmem_free_cta<4 * kTwCols>(tmem_base, t);
}
#endif  // ACQ_VARIANT_L1_ST

// k_search_l1_mst -- k_search_l1_multi in the two-team form with staging warps (see k_search_l1_dr): a team is what a CTA
// of k_search_l1_multi is (code run parked in tensor memory after block 0, stage-B twiddles from a shared table), its
// staging warp issues the bulk copies -- D for every block, E for block 0 -- and stores the cells.  Teams stride over the
// tiles like the CTAs of k_search_l1_multi.  Same arithmetic in the same order: bitwise-equal cells (tested).
// Measured (cfg2, same box): 3.660 ms against 3.649 ms for k_search_l1_multi -- no gain: at 112 registers the block powers
// P[16] spill, and K = 20 blocks per tile leave thread 0's per-tile work little weight.  Experiment builds only.
#ifdef ACQ_VARIANT_L1_MST
__global__ void __launch_bounds__(kStThreads, 1) k_search_l1_mst(const SearchArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x;
    ACQ_TRACE_STAMP(kTrSearchL1, 0);
    constexpr int L = ACQ_LAGS_L1;
    __shared__ uint32_t tmem_slot;
    const uint32_t tmem_base = tmem_alloc_cta<4 * kTwCols>(&tmem_slot, t);
    if (t < 2) mbar_init(smem_u32(dr_smem_carve(smem + t * dr_team_smem_bytes()).bar), 1);
    __syncthreads();
    if (p.wait_prior) pdl_wait();
    ACQ_TRACE_STAMP(kTrSearchL1, 1);
    pdl_trigger_search();
    const int team = (t < 512) ? (t >> 8) : ((t >> 5) & 1);
    const unsigned vcta = (unsigned)team * gridDim.x + blockIdx.x, stride = 2u * gridDim.x;
    const DrSmem s = dr_smem_carve(smem + team * dr_team_smem_bytes());
    const uint32_t bar = smem_u32(s.bar);
    const int sd = (int)(stride % (unsigned)p.n_dop), sw = (int)((stride / (unsigned)p.n_dop) % (unsigned)p.n_work),
              sc = (int)(stride / ((unsigned)p.n_dop * (unsigned)p.n_work));
    if (t >= 512) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (t < 576 && vcta < p.n_tiles) {
            // ---- staging warp of `team` (lane 0 works)
            const bool lead = (t & 31) == 0;
            float *red_f = s.red_f;
            int *red_i = s.red_i;
            auto issue = [&](const TileIdx &tn, int bn, int k2n, int half) {   // D always, E for block 0
                const float2 *Dk = p.Dp + d_row(p, tn, bn) * kN + k2n * kSub;
                fence_proxy_async();  // generic-proxy accesses of these buffers (ordered by the team barrier) before the async writes
                if (bn == 0) {
                    const int r = (k2n - tn.dop) & 3;
                    const int q = (k2n - tn.dop - r) >> 2;
                    const float2 *Ek = p.Ep + (size_t)(tn.sat * 4 + r) * p.ext_len + ((p.Q + q) & ~1);
                    mbar_expect_tx(bar, (uint32_t)(sizeof(float2) * (kSub + kEBufElems)));
                    tma_load_1d(smem_u32(s.E), Ek, (uint32_t)(sizeof(float2) * kEBufElems), bar);
                } else {
                    mbar_expect_tx(bar, (uint32_t)(sizeof(float2) * kSub));
                }
                tma_load_1d(smem_u32(s.S1 + half * kSub), Dk, (uint32_t)(sizeof(float2) * kSub), bar);
            };
            TileIdx ti(p, vcta);
            if (lead) issue(ti, 0, 0, 0);
            int it = 0, par = 0, pend_cap = -1, pend_slot = 0, pend_d = 0;
            for (long long tile = vcta; tile < p.n_tiles; tile += stride) {
                TileIdx tn = ti;
                const bool more = tile + stride < p.n_tiles;
                if (more) tn.step(p, sd, sw, sc);
                for (int b = 0; b < p.K; b++) {
#pragma unroll 1
                    for (int k2 = 0; k2 < 4; k2++) {
                        __syncwarp();
                        st_stage_wait(team);   // every FFT warp is past its operand reads of this sub-FFT and past stage C of the previous one
                        if (lead) {
                            if (k2 < 3) issue(ti, b, k2 + 1, (it + 1) & 1);
                            else if (b + 1 < p.K) issue(ti, b + 1, 0, (it + 1) & 1);
                            else if (more) issue(tn, 0, 0, (it + 1) & 1);
                            if (b == 0 && k2 == 0 && pend_cap >= 0)   // the previous tile's peak: its warp partials precede this barrier
                                store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
                        }
                        it++;
                    }
                }
                pend_cap = ti.cap;
                pend_slot = ti.slot;
                pend_d = ti.d;
                par ^= 1;
                ti = tn;
            }
            __syncwarp();
            st_stage_wait_done(team);
            if (lead) {
                store_cell(p, pend_cap, pend_slot, pend_d, merge_warp_peaks(red_f + 16 * (par ^ 1), red_i + 8 * (par ^ 1)), L);
                __threadfence();  // this team's cells (all stored by this thread) before the CTA's count
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        if (vcta < p.n_tiles) {
            // ---- the eight FFT warps of `team`
            const int tt = t & 255;
            float *red_f = s.red_f;
            int *red_i = s.red_i;
            const uint32_t e_taddr = tmem_base + (uint32_t)(team * 2 * kTwCols) + tmem_lane_base(tt) + (uint32_t)((tt >> 7) * kTwCols);  // [k2][16 complex]
            {   // stage-B twiddle table into this team's shared memory
                const float4 *src = reinterpret_cast<const float4 *>(p.tables);
                float4 *dst = reinterpret_cast<float4 *>(s.T2);
                for (int i = tt; i < kT2Elems / 2; i += 256) dst[i] = __ldg(src + i);
            }
            const float2 *bases = p.tables + kT2Elems + tt;  // [k2][256]: W16384^{4t+k2}
            float2 bw = __ldg(bases);
            int it = 0, par = 0;
            const TileIdx t0(p, vcta);
            int d = t0.d, dop = t0.dop;
            st_fft_sync(team);   // the twiddle table is in place
            for (long long tile = vcta; tile < p.n_tiles; tile += stride) {
                float P[16];
                float2 acc[16];
                for (int b = 0; b < p.K; b++) {
                    float2 x[16];
#pragma unroll 1
                    for (int k2 = 0; k2 < 4; k2++) {
                        float2 *S1b = s.S1 + (it & 1) * kSub;
                        const float2 *Dk = S1b + tt;
                        if (b == 0) {   // E from the staged run; park this thread's 16 values for the blocks to come
                            const int r = (k2 - dop) & 3;
                            const int q = (k2 - dop - r) >> 2;
                            const float2 *Ek = s.E + ((p.Q + q) & 1) + tt;
                            mbar_wait(bar, (uint32_t)(it & 1));
#pragma unroll
                            for (int a = 0; a < 16; a++) x[a] = Ek[256 * a];
                            tmem_st16(e_taddr + 32 * k2, x);
                            tmem_wait_st();
                        } else {
                            tmem_ld16(e_taddr + 32 * k2, x);
                            mbar_wait(bar, (uint32_t)(it & 1));
                            tmem_wait_ld();
                        }
                        // x[a] = conj(data[k]) * code[k - dop], k = 1024 a + 4 t + k2   (search.cpp:471)
#pragma unroll
                        for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(Dk[256 * a], x[a]);
                        subfft4096_inv4s(x, k2, bw, S1b, tt, s.T2, BaseFromGlobal{bases}, [&] { st_fft_arrive(team); st_fft_sync(team); }, [] {});
                        it++;
                        if (k2 == 0) {
#pragma unroll
                            for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                        } else {
#pragma unroll
                            for (int n2 = 0; n2 < 16; n2++) acc[n2] = cfma(x[r16(n2)], c_cC[k2][n2], acc[n2]);
                        }
                    }
                    // block b was delayed by 16*b samples in the front end (k_front_end), so lag n lines up across blocks
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) P[n2] = (b == 0) ? cpower(acc[n2]) : (P[n2] + cpower(acc[n2]));
                }
                warp_reduce_peak(thread_peak_l1(P, tt), red_f + 16 * par, red_i + 8 * par, tt);
                par ^= 1;
                {   // the Doppler index of this team's next tile (the only tile coordinate an FFT warp needs)
                    d += sd;
                    if (d >= p.n_dop) d -= p.n_dop;
                    const int h = p.dop_lo + d;
                    const int v = p.half_bin ? (h & 1) : 0;
                    dop = p.half_bin ? ((h - v) >> 1) : h;
                }
            }
            st_fft_done(team);   // the last tile's warp partials are in place
        }
    }
    __syncthreads();
    if (t == 0 && p.ctas_total) atomicAdd(p.ctas_done, 1u);   // both teams' cells are stored and fenced (staging warps, above)
    ACQ_TRACE_STAMP(kTrSearchL1, 2);
    tmem_free_cta<4 * kTwCols>(tmem_base, t);
}
#endif  // ACQ_VARIANT_L1_MST
