// acq_kernels.cu -- sm_100a kernels of the acquisition engine.
//
//   K1a k_hb1_bits   : unpack 1-bit capture, fs/4 XOR mix, first half-band /2      (search.cpp:408-437)
//   K6a k_hb1_code   : C/A / E1B(BOC) replica samples, first half-band /2            (search.cpp:250-275,315-337)
//   K1b k_hb2        : second half-band /2 (+ optional half-bin pre-rotation)       (search.cpp:439-441)
//   K2  k_fwd_fft    : 16384-point forward FFT of data or code                       (search.cpp:280,342,447)
//   K6b k_build_ext  : polyphase, margin-extended code-spectrum rows                 (search.cpp:283-284,471)
//   K3-5 k_search<M> : conj(D).C product, 16384-point inverse FFT, |.|^2, non-coherent
//                      sum, max / first-argmax / mean per (capture, sat, Doppler)    (search.cpp:465-494)
//   K5b k_best_dop   : best-snr Doppler per (capture, sat), lowest index on ties    (search.cpp:495)
//
// The half-band stages use explicit round-to-nearest mul/add in the reference's summation order
// (no FMA contraction), so the FFT inputs are bit-identical to the reference's x86 build.
#include "acq_fft.cuh"
#include "acq_search2.cuh"
#include "acq_kernels.cuh"

namespace acq {

// half-band taps in the order the reference applies them (search.cpp:141-158):
//   [0] = COEF[0], [1..15] = COEF[2], COEF[4] .. COEF[30], [16] = COEF[15]
__constant__ float c_hb[17];

int launch_tables_init(const float2 *h_cA, const float2 *h_cC, const float *h_hb)
{
    cudaMemcpyToSymbol(c_cA, h_cA, sizeof(float2) * 64);
    cudaMemcpyToSymbol(c_cC, h_cC, sizeof(float2) * 64);
    cudaMemcpyToSymbol(c_hb, h_hb, sizeof(float) * 17);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// half-band helpers: acc = x[0]*c0; acc += x[j]*c_j (j = 2,4..30); acc += x[15]*c15
// ---------------------------------------------------------------------------------------------
struct HbAcc {
    float re, im;
    __device__ __forceinline__ void first(float xr, float xi, float c)
    {
        re = __fmul_rn(xr, c);
        im = __fmul_rn(xi, c);
    }
    __device__ __forceinline__ void add(float xr, float xi, float c)
    {
        re = __fadd_rn(re, __fmul_rn(xr, c));
        im = __fadd_rn(im, __fmul_rn(xi, c));
    }
};

// K1a.  One thread per output sample o of x1 (32768 per block).  Sample i of the block is bit i&7 of
// byte i>>3 (search.cpp:408-411); LO phase is i&3 (lo_rate == 1, search.cpp:386,422-423);
// I = bit ^ {1,1,0,0}[i&3], Q = bit ^ {1,0,0,1}[i&3]; value = bit ? -1 : +1 (search.cpp:62-66,172-175).
__global__ void __launch_bounds__(256) k_hb1_bits(const uint8_t *__restrict__ packed, float2 *__restrict__ x1)
{
    const int o = blockIdx.x * 256 + threadIdx.x;
    const uint8_t *pk = packed + (size_t)blockIdx.y * ACQ_BLOCK_BYTES;
    const int i0 = 2 * o;
    unsigned long long win = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int idx = (i0 >> 3) + k;
        const unsigned byte = (idx < ACQ_BLOCK_BYTES) ? pk[idx] : 0u;
        win |= (unsigned long long)byte << (8 * k);
    }
    win >>= (i0 & 7);
    auto sample = [&](int j, float &xr, float &xi) {
        const int i = i0 + j;
        if (i < ACQ_NSAMPLES) {
            const unsigned bit = (unsigned)(win >> j) & 1u;
            const unsigned ph = i & 3;
            const unsigned lsin = (ph < 2) ? 1u : 0u;             // {1,1,0,0}
            const unsigned lcos = (ph == 0 || ph == 3) ? 1u : 0u; // {1,0,0,1}
            xr = (bit ^ lsin) ? -1.0f : 1.0f;
            xi = (bit ^ lcos) ? -1.0f : 1.0f;
        } else {  // zero padding past the end of the block (search.cpp:145)
            xr = 0.0f;
            xi = 0.0f;
        }
    };
    HbAcc acc;
    float xr, xi;
    sample(0, xr, xi);
    acc.first(xr, xi, c_hb[0]);
#pragma unroll
    for (int j = 2; j <= 30; j += 2) {
        sample(j, xr, xi);
        acc.add(xr, xi, c_hb[j / 2]);
    }
    sample(15, xr, xi);
    acc.add(xr, xi, c_hb[16]);
    x1[(size_t)blockIdx.y * 32768 + o] = make_float2(acc.re, acc.im);
}

// K6a.  Replica: sample i carries chip (i>>4) mod codelen (ca_rate = 1/16 exactly, search.cpp:205,254-258),
// E1B XORs the BOC(1,1) sub-carrier (i&15) >= 8 (search.cpp:317-318).  Imaginary part is zero (search.cpp:266).
__global__ void __launch_bounds__(256) k_hb1_code(const uint32_t *__restrict__ chips,
                                                  const int *__restrict__ codelen_boc, float2 *__restrict__ x1)
{
    const int o = blockIdx.x * 256 + threadIdx.x;
    const int sat = blockIdx.y;
    const uint32_t *cw = chips + (size_t)sat * 128;
    const int codelen = codelen_boc[2 * sat];
    const int boc = codelen_boc[2 * sat + 1];
    const int i0 = 2 * o;
    auto sample = [&](int j) -> float {
        const int i = i0 + j;
        if (i >= ACQ_NSAMPLES) return 0.0f;
        const int ci = (i >> 4) % codelen;
        unsigned c = (cw[ci >> 5] >> (ci & 31)) & 1u;
        if (boc) c ^= ((i & 15) >= 8) ? 1u : 0u;
        return c ? -1.0f : 1.0f;
    };
    float acc = __fmul_rn(sample(0), c_hb[0]);
#pragma unroll
    for (int j = 2; j <= 30; j += 2) acc = __fadd_rn(acc, __fmul_rn(sample(j), c_hb[j / 2]));
    acc = __fadd_rn(acc, __fmul_rn(sample(15), c_hb[16]));
    x1[(size_t)sat * 32768 + o] = make_float2(acc, 0.0f);
}

// K1b.  Second half-band stage, 32768 -> 16384 per row.  With nvar == 2 also writes the half-bin
// variant x2[n] * exp(-j*pi*n/N) (rot[] is computed on the host in double precision).
// x2 layout: [row][v][16384].
// Non-coherent mode (K > 1): block b = row % K of a capture is stored circularly delayed by 16*b samples,
// x2'[(n + 16 b) mod N] = x2[n], which makes its correlation r'_b[n] = r_b[(n + 16 b) mod N]: the
// 16-lag-per-block code advance (65536 = 4 x 16368 + 64 samples) is removed before the FFT, so the
// search kernel can add block powers lag by lag in registers.
__global__ void __launch_bounds__(256) k_hb2(const float2 *__restrict__ x1, float2 *__restrict__ x2,
                                             const float2 *__restrict__ rot, int nvar, int K, int row0)
{
    const int o = blockIdx.x * 256 + threadIdx.x;
    const float2 *in = x1 + (size_t)blockIdx.y * 32768;
    const int i0 = 2 * o;
    auto sample = [&](int j) -> float2 {
        const int i = i0 + j;
        return (i < 32768) ? in[i] : make_float2(0.0f, 0.0f);
    };
    HbAcc acc;
    float2 v = sample(0);
    acc.first(v.x, v.y, c_hb[0]);
#pragma unroll
    for (int j = 2; j <= 30; j += 2) {
        v = sample(j);
        acc.add(v.x, v.y, c_hb[j / 2]);
    }
    v = sample(15);
    acc.add(v.x, v.y, c_hb[16]);
    float2 *out = x2 + (size_t)blockIdx.y * nvar * kN;
    const int oo = (o + 16 * (int)((row0 + blockIdx.y) % K)) & (kN - 1);
    out[oo] = make_float2(acc.re, acc.im);
    if (nvar == 2) {
        const float2 w = rot[o];
        out[kN + oo] = make_float2(__fsub_rn(__fmul_rn(acc.re, w.x), __fmul_rn(acc.im, w.y)),
                                  __fadd_rn(__fmul_rn(acc.re, w.y), __fmul_rn(acc.im, w.x)));
    }
}

// ---------------------------------------------------------------------------------------------
// K2.  Forward 16384-point FFT of one row per CTA:  Y = conj( IFFT( conj(x) ) ).
// The four sub-FFT outputs of a lag are produced by the same thread, so the radix-4 combine over
// k2 uses a thread-private shared scratch Z (three residues parked, the fourth in registers).
// POLY: write Y in the polyphase layout the search kernel reads ([k2][k1] = Y[4*k1 + k2]).
// ---------------------------------------------------------------------------------------------
constexpr size_t kZBytes = sizeof(float2) * 3 * 16 * 256;

template <bool POLY>
__global__ void __launch_bounds__(256, 1) k_fwd_fft(const float2 *__restrict__ x2, float2 *__restrict__ out,
                                                    const float2 *__restrict__ tables, int n_rows)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem s = fft_smem_carve(smem);
    float2 *Z = reinterpret_cast<float2 *>(smem + fft_smem_bytes());
    const int t = threadIdx.x;
    fft_load_tables(s, tables, t, 256);
    for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const float2 *in = x2 + (size_t)row * kN;
        float2 *o = out + (size_t)row * kN;
        float2 x[16];
        for (int k2 = 0; k2 < 4; k2++) {
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const float2 v = in[1024 * a + 4 * t + k2];
                x[a] = make_float2(v.x, -v.y);
            }
            subfft4096_inv(x, k2, s, t);
            if (k2 < 3) {
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) {
                    const float2 z = (k2 == 0) ? x[r16(n2)] : cmul(x[r16(n2)], c_cC[k2][n2]);
                    Z[(k2 * 16 + n2) * 256 + t] = z;
                }
            }
        }
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            float2 z0 = Z[(0 * 16 + n2) * 256 + t];
            float2 z1 = Z[(1 * 16 + n2) * 256 + t];
            float2 z2 = Z[(2 * 16 + n2) * 256 + t];
            float2 z3 = cmul(x[r16(n2)], c_cC[3][n2]);
            radix4_inv(z0, z1, z2, z3);  // z_m = sum_k2 z_k2 * j^{k2*m}
            const float2 zz[4] = {z0, z1, z2, z3};
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int n = t + 256 * n2 + 4096 * m;
                const float2 y = make_float2(zz[m].x, -zz[m].y);
                if (POLY) o[(n & 3) * 4096 + (n >> 2)] = y;
                else o[n] = y;
            }
        }
        // Z is thread-private and S1/S2 hazards are covered inside subfft4096_inv: no barrier needed.
    }
}

// ---------------------------------------------------------------------------------------------
// K6b.  Extended polyphase rows of the code spectra:
//   Ep[(sat*4 + r)][mm] = E_sat[4*(mm - Q) + r],   E_sat[j] = C_sat[j mod N]          for j <  N
//                                                            = C_{sat+1}[j - N] or 0   for j >= N, reference wrap
//                                                            = C_sat[j - N]            for j >= N, circular wrap
// so that the product for Doppler bin `dop` reads E_sat[k - dop] with no modulo (the reference gets
// the same effect from its doubled rows, search.cpp:54,283-284,471).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_ext(const float2 *__restrict__ C, float2 *__restrict__ Ep, int n_sats,
                                                   int Q, int ext_len, int wrap_mode)
{
    const int mm = blockIdx.x * 256 + threadIdx.x;
    if (mm >= ext_len) return;
    const int sat = blockIdx.y >> 2, r = blockIdx.y & 3;
    const int j = 4 * (mm - Q) + r;
    float2 v;
    if (j < 0) v = C[(size_t)sat * kN + j + kN];
    else if (j < kN) v = C[(size_t)sat * kN + j];
    else if (wrap_mode == ACQ_WRAP_CIRCULAR) v = C[(size_t)sat * kN + j - kN];
    else v = (sat + 1 < n_sats) ? C[(size_t)(sat + 1) * kN + j - kN] : make_float2(0.0f, 0.0f);
    Ep[(size_t)blockIdx.y * ext_len + mm] = v;
}

// ---------------------------------------------------------------------------------------------
// K3-5.  The search kernel.  Persistent CTAs stride over tiles.
//   M = 1: Navstar / QZSS, lags 0..4091 (only m = 0 of the radix-4 combine is formed)
//   M = 4: Galileo E1B,    lags 0..16367
//   MULTI: k_noncoh > 1 (M = 1 only): power summed over blocks, P[n] += |r_b[(n + 16 b) mod N]|^2; the
//          16-lag-per-block code advance is removed in the front end by delaying block b (see k_hb2).
// ---------------------------------------------------------------------------------------------
struct Peak {
    float p;
    int n;
    float sum;
};

__device__ __forceinline__ void peak_merge(Peak &a, float p, int n, float sum)
{
    if (p > a.p || (p == a.p && n < a.n)) {  // first (lowest) index wins ties, search.cpp:488
        a.p = p;
        a.n = n;
    }
    a.sum += sum;
}

__device__ __forceinline__ Peak block_reduce_peak(Peak v, float *red_f, int *red_i, int t)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float p = __shfl_xor_sync(0xffffffffu, v.p, off);
        const int n = __shfl_xor_sync(0xffffffffu, v.n, off);
        const float s = __shfl_xor_sync(0xffffffffu, v.sum, off);
        peak_merge(v, p, n, s);
    }
    const int w = t >> 5;
    if ((t & 31) == 0) {
        red_f[w] = v.p;
        red_f[8 + w] = v.sum;
        red_i[w] = v.n;
    }
    __syncthreads();
    if (t == 0) {
        v.p = red_f[0];
        v.n = red_i[0];
        v.sum = red_f[8];
#pragma unroll
        for (int k = 1; k < 8; k++) peak_merge(v, red_f[k], red_i[k], red_f[8 + k]);
    }
    return v;  // valid in thread 0
}

template <int M, bool MULTI>
__global__ void __launch_bounds__(256, (M == 1) ? 2 : 1) k_search(const SearchArgs p)
{
    static_assert(!(MULTI && M != 1), "non-coherent accumulation is implemented for the 1 ms lag window only");
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem s = fft_smem_carve(smem);
    unsigned char *tail = smem + fft_smem_bytes();
    float2 *Z = reinterpret_cast<float2 *>(tail);  // M == 4 only
    if (M == 4) tail += kZBytes;
    float *red_f = reinterpret_cast<float *>(tail);
    int *red_i = reinterpret_cast<int *>(tail + 16 * sizeof(float));
    const int t = threadIdx.x;
    constexpr int L = (M == 1) ? ACQ_LAGS_L1 : ACQ_LAGS_E1B;

    fft_load_tables(s, p.tables, t, 256);

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int d = (int)(tile % p.n_dop);
        const long long cw = tile / p.n_dop;
        const int wi = (int)(cw % p.n_work);
        const int cap = (int)(cw / p.n_work);
        const int2 wk = p.work[wi];
        const int sat = wk.x, slot = wk.y;
        const int h = p.dop_lo + d;
        const int v = p.half_bin ? (h & 1) : 0;
        const int dop = p.half_bin ? ((h - v) >> 1) : h;

        float P[MULTI ? 16 : 1];
        Peak best;
        best.p = 0.0f;
        best.n = 0x7fffffff;
        best.sum = 0.0f;

        for (int b = 0; b < p.K; b++) {
            const float2 *Dblk = p.Dp + ((size_t)((size_t)cap * p.K + b) * p.nvar + v) * kN + t;
            float2 x[16];
            float2 acc[16];
            for (int k2 = 0; k2 < 4; k2++) {
                const int r = (k2 - dop) & 3;
                const int q = (k2 - dop - r) >> 2;
                const float2 *Dk = Dblk + k2 * kSub;
                const float2 *Ek = p.Ep + (size_t)(sat * 4 + r) * p.ext_len + p.Q + q + t;
                // prod = conj(data) * code[k - dop]   (search.cpp:471, support/simd.cpp:12-40)
#pragma unroll
                for (int a = 0; a < 16; a++) x[a] = cmul_conj_a(__ldg(Dk + 256 * a), __ldg(Ek + 256 * a));
                subfft4096_inv(x, k2, s, t);
                if (M == 1) {
                    if (k2 == 0) {
#pragma unroll
                        for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                    } else {
#pragma unroll
                        for (int n2 = 0; n2 < 16; n2++) acc[n2] = cadd(acc[n2], cmul(x[r16(n2)], c_cC[k2][n2]));
                    }
                } else if (k2 < 3) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) {
                        const float2 z = (k2 == 0) ? x[r16(n2)] : cmul(x[r16(n2)], c_cC[k2][n2]);
                        Z[(k2 * 16 + n2) * 256 + t] = z;
                    }
                }
            }

            if (M == 1 && !MULTI) {
                // power, max, first argmax, sum over lags n = t + 256*n2 < 4092   (search.cpp:486-490)
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) {
                    const int n = t + 256 * n2;
                    const float pw = acc[n2].x * acc[n2].x + acc[n2].y * acc[n2].y;
                    if (n < L) {
                        if (pw > best.p) best.p = pw, best.n = n;
                        best.sum += pw;
                    }
                }
            } else if (M == 1 && MULTI) {
                // block b was delayed by 16*b samples in the front end (k_hb2), so lag n lines up across blocks
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) {
                    const float pw = acc[n2].x * acc[n2].x + acc[n2].y * acc[n2].y;
                    P[n2] = (b == 0) ? pw : P[n2] + pw;
                }
            } else {
                // E1B: radix-4 combine over k2, lags n = t + 256*n2 + 4096*m < 16368
                // (lags are not visited in increasing order here, so ties compare the index explicitly)
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) {
                    float2 z0 = Z[(0 * 16 + n2) * 256 + t];
                    float2 z1 = Z[(1 * 16 + n2) * 256 + t];
                    float2 z2 = Z[(2 * 16 + n2) * 256 + t];
                    float2 z3 = cmul(x[r16(n2)], c_cC[3][n2]);
                    radix4_inv(z0, z1, z2, z3);  // z_m = sum_k2 z_k2 * j^{k2*m}
                    const float2 zz[4] = {z0, z1, z2, z3};
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const int n = t + 256 * n2 + 4096 * m;
                        const float pw = zz[m].x * zz[m].x + zz[m].y * zz[m].y;
                        if (n < L) peak_merge(best, pw, n, pw);
                    }
                }
            }
        }

        if (MULTI) {
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) {
                const int n = t + 256 * n2;
                if (n < L) {
                    if (P[n2] > best.p) best.p = P[n2], best.n = n;
                    best.sum += P[n2];
                }
            }
        }

        const Peak tot = block_reduce_peak(best, red_f, red_i, t);
        if (t == 0) {
            acq_cell c;
            c.peak = tot.p;
            c.noise = __fdiv_rn(tot.sum, (float)L);   // ave_pwr = tot_pwr / i   (search.cpp:493)
            c.snr = __fdiv_rn(tot.p, c.noise);        // snr = max_pwr / ave_pwr (search.cpp:494)
            c.lag = (tot.n == 0x7fffffff) ? 0 : tot.n;
            p.cells[((size_t)cap * p.n_slots + slot) * p.n_dop + d] = c;
        }
        // red_f/red_i are next written after >= 8 barriers of the following tile: no extra barrier.
    }
}

// ---------------------------------------------------------------------------------------------
// K3-5, packed variant for the 1 ms lag window (Navstar / QZSS): two Doppler indices per thread in the two
// lanes of f32x2 instructions (see acq_search2.cuh).  One persistent 256-thread CTA per SM strides over
// (capture, satellite, Doppler pair).  Lane A = index d0, lane B = index d1 = d0 + 1 bin (d1 == d0 for the
// unpaired last index, whose lane-B result is dropped).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_planar(const float2 *__restrict__ Ep, float *__restrict__ ERp,
                                                      float *__restrict__ EIp, size_t n)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const float2 v = Ep[i];
        ERp[i] = v.x;
        EIp[i] = v.y;
    }
}

template <bool MULTI>
__global__ void __launch_bounds__(256, 1) k_search2(const SearchArgs p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem2 s = fft_smem2_carve(smem);
    float *red_f = reinterpret_cast<float *>(smem + fft_smem2_bytes());
    int *red_i = reinterpret_cast<int *>(red_f + 32);
    const int t = threadIdx.x;
    constexpr int L = ACQ_LAGS_L1;
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.tables);
        float4 *dst = reinterpret_cast<float4 *>(s.T1);
        for (int i = t; i < (kT1Elems + kT2Elems) / 2; i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const long long n_pairs = (long long)p.n_cap * p.n_work * p.ppr;
    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int j = (int)(pr % p.ppr);
        const long long cw = pr / p.ppr;
        const int wi = (int)(cw % p.n_work);
        const int cap = (int)(cw / p.n_work);
        const int2 wk = p.work[wi];
        const int sat = wk.x, slot = wk.y;
        const int2 dd = p.pairs[j];
        const int h0 = p.dop_lo + dd.x;
        const int v = p.half_bin ? (h0 & 1) : 0;
        const int dopA = p.half_bin ? ((h0 - v) >> 1) : h0;
        const int dopB = (dd.y != dd.x) ? dopA + 1 : dopA;

        CP acc[16];
        float2 P[MULTI ? 16 : 1];
        for (int b = 0; b < p.K; b++) {
            const float2 *Dblk = p.Dp + ((size_t)((size_t)cap * p.K + b) * p.nvar + v) * kN + t;
            CP x[16];
            for (int k2 = 0; k2 < 4; k2++) {
                const int rA = (k2 - dopA) & 3, qA = (k2 - dopA - rA) >> 2;
                const int rB = (k2 - dopB) & 3, qB = (k2 - dopB - rB) >> 2;
                const size_t oA = (size_t)(sat * 4 + rA) * p.ext_len + p.Q + qA + t;
                const size_t oB = (size_t)(sat * 4 + rB) * p.ext_len + p.Q + qB + t;
                const float2 *Dk = Dblk + k2 * kSub;
                const float *erA = p.ERp + oA, *eiA = p.EIp + oA, *erB = p.ERp + oB, *eiB = p.EIp + oB;
                // prod = conj(D) * E for both lanes; D enters as a broadcast scalar operand
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    const float2 dv = __ldg(Dk + 256 * a);
                    const float2 ER = make_float2(__ldg(erA + 256 * a), __ldg(erB + 256 * a));
                    const float2 EI = make_float2(__ldg(eiA + 256 * a), __ldg(eiB + 256 * a));
                    x[a].re = p_fma(EI, p_bc(dv.y), p_mul(ER, p_bc(dv.x)));
                    x[a].im = p_fma(ER, p_bc(-dv.y), p_mul(EI, p_bc(dv.x)));
                }
                subfft4096_inv2(x, k2, s, t);
                if (k2 == 0) {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) acc[n2] = x[r16(n2)];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < 16; n2++) {
                        const float2 w = c_cC[k2][n2];
                        const CP z = x[r16(n2)];
                        acc[n2].re = p_fma(z.im, p_bc(-w.y), p_fma(z.re, p_bc(w.x), acc[n2].re));
                        acc[n2].im = p_fma(z.im, p_bc(w.x), p_fma(z.re, p_bc(w.y), acc[n2].im));
                    }
                }
            }
            if (MULTI) {
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) {
                    const float2 pw = p_fma(acc[n2].im, acc[n2].im, p_mul(acc[n2].re, acc[n2].re));
                    P[n2] = (b == 0) ? pw : p_add(P[n2], pw);
                }
            }
        }

        Peak bA, bB;
        bA.p = bB.p = 0.0f;
        bA.n = bB.n = 0x7fffffff;
        float2 sum = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const int n = t + 256 * n2;
            float2 pw = MULTI ? P[n2] : p_fma(acc[n2].im, acc[n2].im, p_mul(acc[n2].re, acc[n2].re));
            if (n2 == 15 && n >= L) pw = make_float2(0.0f, 0.0f);  // lags 4092..4095 are not scanned (search.cpp:486)
            if (pw.x > bA.p) bA.p = pw.x, bA.n = n;
            if (pw.y > bB.p) bB.p = pw.y, bB.n = n;
            sum = p_add(sum, pw);
        }
        bA.sum = sum.x;
        bB.sum = sum.y;
        // block reduction of both lanes (shuffles, then one partial per warp)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            peak_merge(bA, __shfl_xor_sync(0xffffffffu, bA.p, off), __shfl_xor_sync(0xffffffffu, bA.n, off),
                       __shfl_xor_sync(0xffffffffu, bA.sum, off));
            peak_merge(bB, __shfl_xor_sync(0xffffffffu, bB.p, off), __shfl_xor_sync(0xffffffffu, bB.n, off),
                       __shfl_xor_sync(0xffffffffu, bB.sum, off));
        }
        const int w = t >> 5;
        if ((t & 31) == 0) {
            red_f[w] = bA.p;
            red_f[8 + w] = bA.sum;
            red_f[16 + w] = bB.p;
            red_f[24 + w] = bB.sum;
            red_i[w] = bA.n;
            red_i[8 + w] = bB.n;
        }
        __syncthreads();
        if (t < 2 && (t == 0 || dd.y != dd.x)) {
            Peak tot;
            const int o = t * 16, oi = t * 8;
            tot.p = red_f[o];
            tot.n = red_i[oi];
            tot.sum = red_f[o + 8];
#pragma unroll
            for (int k = 1; k < 8; k++) peak_merge(tot, red_f[o + k], red_i[oi + k], red_f[o + 8 + k]);
            acq_cell c;
            c.peak = tot.p;
            c.noise = __fdiv_rn(tot.sum, (float)L);
            c.snr = __fdiv_rn(tot.p, c.noise);
            c.lag = (tot.n == 0x7fffffff) ? 0 : tot.n;
            p.cells[((size_t)cap * p.n_slots + slot) * p.n_dop + (t == 0 ? dd.x : dd.y)] = c;
        }
        // red_* are rewritten only after the >= 8 barriers of the next pair
    }
}

// K5b.  max_snr = 0; for dop ascending: if (snr > max_snr) take it   (search.cpp:455,495).
// One warp per (capture, sat) row: lanes stride over the Doppler cells, then a shuffle reduction that
// prefers the larger snr and, on equal snr, the lower Doppler index (what the sequential scan keeps).
// A row whose snr never exceeds 0 (or is NaN) keeps {lag 0, dop 0, zeros}.
__global__ void __launch_bounds__(128) k_best_dop(const acq_cell *__restrict__ cells, const int *__restrict__ slot_sat,
                                                  acq_record *__restrict__ out, int n_rows, int n_slots, int n_dop,
                                                  int dop_lo)
{
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const acq_cell *c = cells + (size_t)row * n_dop;
    float best = 0.0f;
    int best_d = 0x7fffffff;
    for (int d = lane; d < n_dop; d += 32) {
        const float snr = c[d].snr;
        if (snr > best) best = snr, best_d = d;  // ascending d within a lane: first maximum kept
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, best, off);
        const int d2 = __shfl_xor_sync(0xffffffffu, best_d, off);
        if (s2 > best || (s2 == best && d2 < best_d)) best = s2, best_d = d2;
    }
    if (lane == 0) {
        acq_record r;
        r.sat = slot_sat[row % n_slots];
        r.lag = 0;
        r.dop = 0;
        r.peak = 0.0f;
        r.noise = 0.0f;
        r.snr = 0.0f;
        if (best > 0.0f) {
            const acq_cell cc = c[best_d];
            r.lag = cc.lag;
            r.dop = dop_lo + best_d;
            r.peak = cc.peak;
            r.noise = cc.noise;
            r.snr = cc.snr;
        }
        out[row] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static size_t search2_smem_bytes() { return fft_smem2_bytes() + 64 * sizeof(float); }
size_t search_smem_bytes(bool e1b) { return fft_smem_bytes() + (e1b ? kZBytes : 0) + 64 * sizeof(float); }
static size_t fwd_smem_bytes() { return fft_smem_bytes() + kZBytes; }

cudaError_t search_kernels_configure()
{
    cudaError_t e;
    const int l1 = (int)search_smem_bytes(false), e1 = (int)search_smem_bytes(true), fw = (int)fwd_smem_bytes();
    if ((e = cudaFuncSetAttribute(k_search<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1))) return e;
    if ((e = cudaFuncSetAttribute(k_search<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1))) return e;
    if ((e = cudaFuncSetAttribute(k_search<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, e1))) return e;
    const int s2 = (int)search2_smem_bytes();
    if ((e = cudaFuncSetAttribute(k_search2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2))) return e;
    if ((e = cudaFuncSetAttribute(k_search2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2))) return e;
    if ((e = cudaFuncSetAttribute(k_fwd_fft<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw))) return e;
    if ((e = cudaFuncSetAttribute(k_fwd_fft<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw))) return e;
    return cudaSuccess;
}

int launch_hb1_bits(const uint8_t *packed, float2 *x1, int n_blocks, cudaStream_t st)
{
    int launched = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 32768) {  // gridDim.y <= 65535
        const int nb = (n_blocks - b0 < 32768) ? (n_blocks - b0) : 32768;
        k_hb1_bits<<<dim3(128, nb), 256, 0, st>>>(packed + (size_t)b0 * ACQ_BLOCK_BYTES, x1 + (size_t)b0 * 32768);
        launched++;
    }
    return launched;
}

int launch_hb1_code(const uint32_t *chips, const int *codelen_boc, float2 *x1, int n_sats, cudaStream_t st)
{
    k_hb1_code<<<dim3(128, n_sats), 256, 0, st>>>(chips, codelen_boc, x1);
    return 1;
}

int launch_hb2(const float2 *x1, float2 *x2, const float2 *rot, int n_rows, int nvar, int K, cudaStream_t st)
{
    int launched = 0;
    for (int r0 = 0; r0 < n_rows; r0 += 32768) {
        const int nr = (n_rows - r0 < 32768) ? (n_rows - r0) : 32768;
        k_hb2<<<dim3(64, nr), 256, 0, st>>>(x1 + (size_t)r0 * 32768, x2 + (size_t)r0 * nvar * kN, rot, nvar, K, r0);
        launched++;
    }
    return launched;
}

int launch_fwd_fft(const float2 *x2, float2 *out, const float2 *tables, int n_rows, bool polyphase, int sm_count,
                   cudaStream_t st)
{
    const int grid = n_rows < sm_count ? n_rows : sm_count;
    if (polyphase) k_fwd_fft<true><<<grid, 256, fwd_smem_bytes(), st>>>(x2, out, tables, n_rows);
    else k_fwd_fft<false><<<grid, 256, fwd_smem_bytes(), st>>>(x2, out, tables, n_rows);
    return 1;
}

int launch_build_ext(const float2 *C, float2 *Ep, int n_sats, int Q, int ext_len, int wrap_mode, cudaStream_t st)
{
    k_build_ext<<<dim3((ext_len + 255) / 256, n_sats * 4), 256, 0, st>>>(C, Ep, n_sats, Q, ext_len, wrap_mode);
    return 1;
}

int launch_search(const SearchArgs &a, bool e1b, int sm_count, cudaStream_t st)
{
    if (a.n_tiles <= 0) return 0;
    const long long max_ctas = (long long)sm_count * (e1b ? 1 : 2);
    const int grid = (int)(a.n_tiles < max_ctas ? a.n_tiles : max_ctas);
    const size_t smem = search_smem_bytes(e1b);
    if (e1b) k_search<4, false><<<grid, 256, smem, st>>>(a);
    else if (a.K > 1) k_search<1, true><<<grid, 256, smem, st>>>(a);
    else k_search<1, false><<<grid, 256, smem, st>>>(a);
    return 1;
}

int launch_search2(const SearchArgs &a, int sm_count, cudaStream_t st)
{
    const long long n_pairs = (long long)a.n_cap * a.n_work * a.ppr;
    if (n_pairs <= 0) return 0;
    const int grid = (int)(n_pairs < sm_count ? n_pairs : sm_count);
    if (a.K > 1) k_search2<true><<<grid, 256, search2_smem_bytes(), st>>>(a);
    else k_search2<false><<<grid, 256, search2_smem_bytes(), st>>>(a);
    return 1;
}

int launch_build_ext_planar(const float2 *Ep, float *ERp, float *EIp, size_t n, cudaStream_t st)
{
    k_split_planar<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Ep, ERp, EIp, n);
    return 1;
}

int launch_best_dop(const acq_cell *cells, const int *slot_sat, acq_record *out, int n_cap, int n_slots, int n_dop,
                    int dop_lo, cudaStream_t st)
{
    const int n_rows = n_cap * n_slots;
    k_best_dop<<<(n_rows + 3) / 4, 128, 0, st>>>(cells, slot_sat, out, n_rows, n_slots, n_dop, dop_lo);
    return 1;
}

}  // namespace acq
