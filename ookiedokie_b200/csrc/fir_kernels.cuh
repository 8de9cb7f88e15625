// fir_kernels.cuh -- SC16Q11 -> float conversion fused into FIR filter-and-decimate, envelope
// power and threshold decision.  Replaces, per output sample, the reference's
//   sc16q11_to_complexf     src/complexf.h:68-77
//   update / perform_stage  src/fir.c:302-353   (out = sum_i taps[i]*x[n-i], in order, from 0)
//   threshold               src/ookiedokie.c:171-179 (+ src/complexf.h:43-58)
//
// Index conventions (all global, 64-bit): input sample g, stage output j.  Output j of a stage
// with T taps and decimation D is produced when input (j+1)*D-1 arrives and reads inputs
// (j+1)*D-1-i, i = 0..T-1; inputs with a negative index are the zeros fir_reset leaves in the
// delay line (src/fir.c:272-295); raw inputs at or beyond n_valid are the zeros
// sdr_bladerf_file_rx pads the last buffer with (src/sdr/bladeRF_file.c:110-115).
//
// Decisions are written bit-packed, LSB first: decision of global output m lives in bit
// (m - bit_base) of the bit array.
#pragma once

#include <climits>

#include "ookd_common.cuh"

namespace ookd {

// =======================================================================================
// 1. Shape-agnostic stage kernel (any taps/decimation; one launch per stage; intermediates
//    in HBM).  It is the parity dump for fir_filter_and_decimate and the fallback for filter
//    shapes without a tiled kernel.  One thread per output, taps staged in shared memory.
// =======================================================================================
struct GenericStageArgs {
    const void *in;          // int16x2 words (IN_I16) or float2
    i64  in_base;            // global index of in[0]
    i64  in_valid_end;       // inputs >= this index read as zero
    const float *taps;       // device, T floats
    uint32_t T, D;
    i64  out_lo, out_hi;     // global output range [lo, hi)
    float2 *out_cf;          // out_cf[0] <-> out_lo            (may be null)
    uint32_t *out_bits;      // packed decisions, 32-bit words  (may be null)
    i64  bit_base;           // global output index of bit 0; (out_lo - bit_base) % 32 == 0
    float pstar;
};

template <bool IN_I16>
__global__ void __launch_bounds__(256) fir_stage_generic_kernel(const GenericStageArgs a)
{
    extern __shared__ float s_taps[];
    for (uint32_t i = threadIdx.x; i < a.T; i += blockDim.x) {
        s_taps[i] = a.taps[i];
    }
    __syncthreads();

    const i64 j = a.out_lo + (i64) blockIdx.x * blockDim.x + threadIdx.x;
    bool bit = false;
    if (j < a.out_hi) {
        const i64 newest = (j + 1) * (i64) a.D - 1;
        float re = 0.0f, im = 0.0f;
        for (uint32_t i = 0; i < a.T; i++) {
            const i64 g = newest - (i64) i;
            float2 x = make_float2(0.0f, 0.0f);
            if (g >= 0 && g < a.in_valid_end) {
                if (IN_I16) {
                    x = sc16q11_to_float2(((const uint32_t *) a.in)[g - a.in_base]);
                } else {
                    x = ((const float2 *) a.in)[g - a.in_base];
                }
            }
            const float t = s_taps[i];
            re = mac_exact(re, t, x.x);
            im = mac_exact(im, t, x.y);
        }
        if (a.out_cf) {
            a.out_cf[j - a.out_lo] = make_float2(re, im);
        }
        bit = power_exact(re, im) >= a.pstar;
    }
    if (a.out_bits) {
        // (out_lo - bit_base) and blockDim are multiples of 32: a warp owns one whole word
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
        if ((threadIdx.x & 31) == 0 && j < a.out_hi) {
            a.out_bits[(j - a.bit_base) >> 5] = word;
        }
    }
}

// =======================================================================================
// 2. Tiled exact kernel, one stage, decimation 1 (fs32_fs4, fs64_fs8 shapes).
//    CTA = 256 threads; tile = 256*R outputs.  Raw samples are read with 128-bit coalesced
//    loads, converted once and staged in shared memory as float2 with one pad slot per R
//    samples (lane stride R+1 float2: odd => conflict-free 64-bit shared loads).  Each
//    thread keeps a sliding window of R+T-1 samples in registers and runs R independent
//    accumulator pairs through the taps in the reference's order.  Taps are kernel
//    parameters (constant bank 0) so they fold into the FMUL as c[0][..] operands and two
//    handles with different filters never share state.
// =======================================================================================
template <int T>
struct TapsParam {
    float t[T];
};

struct TiledArgs {
    const uint32_t *in;      // int16x2 words
    i64  in_base;            // global index of in[0]; (tile input start - in_base) 16B-aligned or scalar path
    i64  in_valid_end;
    i64  out_lo, out_hi;     // global output range; out_lo % (256*R) == 0 relative to bit_base rule below
    uint8_t *out_bits;       // packed decisions (bytes)
    i64  bit_base;           // (out_lo - bit_base) % 8 == 0
    float pstar;
    const uint32_t *tile_list;   // optional: explicit tile indices (dense-tile pass); null => blockIdx.x
    const uint32_t *tile_count;  // with tile_list: number of entries
};

// (parameter blocks of the screening kernels of section 3; the work list is shared with the FMA form below)
struct ScreenParams {
    uint32_t k0;             // off test: window energy (LSB^2) strictly below this => decision 0
    float g_lo;              // |sum t_i| rounded down
    float t2;                // ||t||_2 rounded up
    float cg;                // gamma * ||t||_2 rounded up
    float theta_hi;          // sqrt(P*) * 2048 grown by 1e-5
    float inv_n;             // 1/48
};

struct ScreenArgs {
    TiledArgs t;
    uint32_t *work_list;     // OUTPUT: indices of 8-output groups (relative to t.bit_base) left to the exact path
    uint32_t *work_count;    // number of groups pushed (may exceed work_cap: then the host redoes the range exactly)
    uint32_t work_cap;
    uint32_t tile_offset;    // first (4096-output) tile of this launch, numbered from t.out_lo
    uint32_t n_tiles;        // tiles of this launch (persistent variant)
};

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // max relative error 2^-23
    return r;
}

// FMA screening (filter-and-refine): the same tiled kernel with ONE fused multiply-add per tap instead of a
// rounded multiply and a rounded add -- half the fp32 instructions -- followed by a rigorous test of whether the reference's
// decision can differ.  With y the exact sum, y_ref the reference's value (2T roundings) and y_f the fused one (T roundings),
//     |y_ref - y_f| <= (g_ref + g_f) sum_i |t_i||x_i| <= (g_ref + g_f) ||t||_2 sqrt(E)       (both components together)
// where E = window energy <= W m^2 (W samples in the window, m^2 = largest |x|^2 staged for the tile), g_* the usual
// recursive-summation factors (host: make_fma_band, in double).  The reference decides 1 iff fl(re^2 + im^2) >= P*, i.e.
// iff its magnitude is at least theta = sqrt(P*) up to 3 ulp; so with p_f the fused value's computed power,
//     p_f >= (theta_hi + D)^2  =>  reference decides 1,        p_f < (theta_lo - D)^2  =>  reference decides 0,
// D = c sqrt(m^2) (c = (g_ref + g_f) ||t||_2 sqrt(W), rounded up).  Anything in between (a few outputs per million on
// real captures) goes to the work list and is recomputed by the exact group kernel: decisions stay bit-identical.
struct FmaBand {
    float c;                 // D = c * sqrt(m^2)
    float theta_hi, theta_lo;    // sqrt(P*) grown / shrunk by the power's own rounding allowance
};

template <bool FMA>
__device__ __forceinline__ float mac_sel(float acc, float tap, float x)
{
    if constexpr (FMA) {
        return fmaf(tap, x, acc);
    } else {
        return mac_exact(acc, tap, x);
    }
}

// classify one output of the fused path: 1 = surely on, 0 = surely off, 2 = the reference could decide either way
__device__ __forceinline__ uint32_t fma_classify(float re, float im, float hi2, float lo2)
{
    const float p = fmaf(re, re, im * im);
    return (p >= hi2) ? 1u : ((p < lo2) ? 0u : 2u);
}

template <int T, int R, bool FMA>
__global__ void __launch_bounds__(256, 2)
fir1_tiled_kernel(const ScreenArgs sa, const TapsParam<T> taps, const FmaBand band)
{
    constexpr int NT = 256;
    constexpr int L = NT * R;                 // outputs per tile
    constexpr int HALO = (T - 1 + 3) & ~3;    // history samples staged in front, multiple of 4
    constexpr int NS = L + HALO;              // staged samples
    constexpr int LOGR = (R == 8) ? 3 : 4;
    static_assert(R == 8 || R == 16, "R");
    static_assert(!FMA || R == 8, "one thread = one 8-output group of the work list");
    __shared__ float2 s_x[NS + (NS >> LOGR) + 1];
    __shared__ float s_m2[NT / 32];

    const TiledArgs &a = sa.t;
    uint32_t n_tiles_here = gridDim.x;
    uint32_t stride = gridDim.x;
    uint32_t tile_it = blockIdx.x;
    if (a.tile_list) {
        n_tiles_here = *a.tile_count;
    }

    for (; tile_it < n_tiles_here; tile_it += stride) {
        const uint32_t tile = a.tile_list ? a.tile_list[tile_it] : tile_it;
        const i64 o0 = a.out_lo + (i64) tile * L;     // first output of the tile
        const i64 g0 = o0 - HALO;                     // first staged input

        // ---- stage inputs: 4 samples (16 B) per thread per step ----
        const bool aligned = (((g0 - a.in_base) & 3) == 0) && ((((uintptr_t) a.in) & 15) == 0);
        float m2 = 0.0f;
        for (int q = threadIdx.x; q < NS / 4; q += NT) {
            const i64 g = g0 + 4 * q;
            uint32_t w[4];
            if (aligned && g >= a.in_base && g >= 0 && g + 4 <= a.in_valid_end) {
                const uint4 v = __ldg((const uint4 *) (a.in + (g - a.in_base)));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const i64 ge = g + e;
                    w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int s = 4 * q + e;
                const float2 x = sc16q11_to_float2(w[e]);
                s_x[s + (s >> LOGR)] = x;
                if constexpr (FMA) m2 = fmaxf(m2, fmaf(x.x, x.x, x.y * x.y));
            }
        }
        if constexpr (FMA) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xFFFFFFFFu, m2, d));
            if ((threadIdx.x & 31) == 0) s_m2[threadIdx.x >> 5] = m2;
        }
        __syncthreads();

        // ---- window into registers: samples s = tid*R + (HALO-(T-1)) + q, q = 0..R+T-2 ----
        float2 win[R + T - 1];
        const int s_first = threadIdx.x * R + (HALO - (T - 1));
#pragma unroll
        for (int q = 0; q < R + T - 1; q++) {
            const int s = s_first + q;
            win[q] = s_x[s + (s >> LOGR)];
        }

        float re[R], im[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            re[j] = 0.0f;
            im[j] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < T; i++) {
#pragma unroll
            for (int j = 0; j < R; j++) {
                re[j] = mac_sel<FMA>(re[j], taps.t[i], win[j + T - 1 - i].x);
                im[j] = mac_sel<FMA>(im[j], taps.t[i], win[j + T - 1 - i].y);
            }
        }

        uint32_t bits = 0;
        bool unsure = false;
        if constexpr (FMA) {
            float mm = s_m2[0];
#pragma unroll
            for (int q = 1; q < NT / 32; q++) mm = fmaxf(mm, s_m2[q]);
            // D rounded up: sqrt.approx is within 2^-22, the products within a few ulp
            const float D = band.c * sqrt_approx(mm) * 1.00001f;
            const float hi = band.theta_hi + D, lo = fmaxf(band.theta_lo - D, 0.0f);
            const float hi2 = hi * hi * 1.000001f, lo2 = lo * lo * 0.999999f;
#pragma unroll
            for (int j = 0; j < R; j++) {
                const uint32_t c = fma_classify(re[j], im[j], hi2, lo2);
                bits |= (c & 1u) << j;
                unsure = unsure || (c == 2u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; j++) {
                bits |= (power_exact(re[j], im[j]) >= a.pstar ? 1u : 0u) << j;
            }
        }
        const i64 o = o0 + (i64) threadIdx.x * R;
        const bool in_range = o < a.out_hi;
        if (in_range) {
            // outputs past out_hi inside the last byte are masked by the consumers
            const i64 byte = (o - a.bit_base) >> 3;
            a.out_bits[byte] = (uint8_t) bits;
            if (R == 16) {
                a.out_bits[byte + 1] = (uint8_t) (bits >> 8);
            }
        }
        if constexpr (FMA) {
            const bool push = unsure && in_range;
            const uint32_t m_push = __ballot_sync(0xFFFFFFFFu, push);
            if (m_push) {
                const int lane = threadIdx.x & 31;
                uint32_t slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(sa.work_count, (uint32_t) __popc(m_push));
                slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
                if (push) {
                    const uint32_t sl = slot0 + __popc(m_push & ((1u << lane) - 1));
                    if (sl < sa.work_cap) sa.work_list[sl] = (uint32_t) ((o - a.bit_base) >> 3);
                }
            }
        }
        __syncthreads();     // s_x is reused by the next tile of this CTA
    }
}


// =======================================================================================
// 3. Screening kernel, one stage, decimation 1, T = 32 taps, 16 outputs per thread.
//
//    The decision the reference takes for output n is  fl(re^2)+fl(im^2) >= P*  on the fp32
//    in-order sums re, im.  Computing those sums costs 4T flops per sample, which on a B200 is
//    ~6x more issue slots than the HBM read of the 4-byte sample allows.  But an OOK capture
//    is mostly "carrier clearly off" or "carrier clearly on", and both can be PROVED from sums
//    of the raw integers, without a single MAC:
//
//    off:  |y[n]| = |sum_i t_i x[n-i]| <= ||t||_2 sqrt(E_n),  E_n = sum over the T samples of the
//          window of |x|^2 (Cauchy-Schwarz).  In integer LSB units E_n is EXACT in 32-bit
//          arithmetic (|I|,|Q| < 4096 is checked), and the reference's rounding moves |y| by at
//          most gamma ||t||_2 sqrt(E_n) and its power by 3 ulp, so  E_n < K0  (K0 computed on the
//          host in double, rounded down) implies the reference decides 0.  E_n is formed for
//          every output from per-thread running prefix sums:
//              E_n = (tot[t-2] - pre[t-2][j]) + tot[t-1] + pre[t][j]        (16 samples per thread)
//    on:   for any complex mu,  y[n] = mu G + sum_i t_i (x[n-i] - mu), hence
//          |y[n]| >= |mu||G| - ||t||_2 sqrt(V),  V = sum over a superset S of the window of
//          |x-mu|^2 = Q - |X|^2/N for mu = mean over S (S = the 48 samples of threads t-2..t).
//          If that lower bound, less the rounding allowance, clears the threshold, all 16
//          outputs of the thread decide 1.
//
//    Groups of 8 outputs that neither test settles are recomputed with the exact in-order MACs
//    (one output per lane, compacted through a shared-memory queue so lanes stay full, samples
//    re-read through L1/L2) by fir1_refine_kernel from a global work list.  Decisions are therefore
//    bit-identical to the exact kernel for EVERY input; only the cost is data dependent.  Spans
//    holding a sample outside the range the 32-bit sums are sized for are simply left undecided.
//    If more groups are undecided than the work list holds (low-SNR captures), the host redoes the
//    range with fir1_exact_tiled_kernel and stops screening on that handle.
// =======================================================================================
#ifndef OOKD_SCREEN_MINB
#define OOKD_SCREEN_MINB 5
#endif
#ifndef OOKD_SCREEN_PERSIST_MINB
#define OOKD_SCREEN_PERSIST_MINB 4
#endif

// prefix sums / sums / range guard of one 16-sample span
__device__ __forceinline__ void screen_span_stats(const uint32_t (&w)[16], uint32_t (&pre)[16], int &sx, int &sy,
                                                  uint32_t &guard)
{
    uint32_t run = 0;
    int xs = 0, ys = 0;
    guard = 0;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int I = (int) (short) (w[e] & 0xFFFFu);
        const int Q = ((int) w[e]) >> 16;
        const uint32_t q = (uint32_t) (I * I) + (uint32_t) (Q * Q);
        guard |= q;
        run += q;
        pre[e] = run;
        xs += I;
        ys += Q;
    }
    sx = xs;
    sy = ys;
}

template <int T>
__global__ void __launch_bounds__(256, OOKD_SCREEN_MINB)
fir1_screen_kernel(const ScreenArgs sa, const ScreenParams sp)
{
    static_assert(T == 32, "window = exactly two 16-sample thread spans");
    constexpr int NT = 256, SPT = 16, L = NT * SPT;    // 4096 outputs per tile
    constexpr int HT = 2;                              // history spans in front of the tile (32 samples)
    __shared__ uint4 s_pre[(NT + HT) * 4];             // per span: 16 running prefix sums of |x|^2 (chunk-rotated)
    __shared__ int2 s_xy[NT + HT];                     // per span: sum I, sum Q
    __shared__ uint8_t s_flag[NT + HT];                // per span: some |x|^2 >= 2^25 (32-bit sums not guaranteed)

    const TiledArgs &a = sa.t;
    const uint32_t tile = blockIdx.x + sa.tile_offset;
    const i64 o0 = a.out_lo + (i64) tile * L;
    const i64 g0 = o0 - HT * SPT;
    const bool fast = ((((uintptr_t) a.in) & 15) == 0) && (((g0 - a.in_base) & 3) == 0) && g0 >= a.in_base && g0 >= 0 &&
                      (g0 + L + HT * SPT) <= a.in_valid_end;

    // ---- pass over the raw words: running prefix of |x|^2, sums of I and Q, range guard ----
    uint32_t pre[SPT], guard = 0;
    int sx = 0, sy = 0;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        int span;
        if (pass == 0) {
            span = threadIdx.x + HT;
        } else {
            if (threadIdx.x >= HT) break;
            span = threadIdx.x;
        }
        const i64 g = g0 + (i64) span * SPT;
        uint32_t w[SPT];
        if (fast) {
            const uint4 *src = (const uint4 *) (a.in + (g - a.in_base));
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 x = __ldg(src + v);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < SPT; e++) {
                const i64 ge = g + e;
                w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
            }
        }
        uint32_t p[SPT], gd;
        int xs, ys;
        screen_span_stats(w, p, xs, ys, gd);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            // rotate the four 16-byte chunks by (span >> 1) so that a warp's stores spread over all banks
            s_pre[span * 4 + ((v + (span >> 1)) & 3)] = make_uint4(p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]);
        }
        s_xy[span] = make_int2(xs, ys);
        s_flag[span] = (gd >> 25) ? 1 : 0;
        if (pass == 0) {
#pragma unroll
            for (int e = 0; e < SPT; e++) pre[e] = p[e];
            sx = xs; sy = ys; guard = gd;
        }
    }
    __syncthreads();

    // ---- decide the 16 outputs of this thread (two groups of 8) ----
    const int span = threadIdx.x + HT;
    uint32_t bits16 = 0;
    bool undecided_lo = false, undecided_hi = false;
    const i64 o = o0 + (i64) threadIdx.x * SPT;
    const bool bad = (guard >> 25) || s_flag[span - 1] || s_flag[span - 2];
    if (!bad) {
        uint32_t p2[SPT];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const uint4 x = s_pre[(span - 2) * 4 + ((v + ((span - 2) >> 1)) & 3)];
            p2[4 * v] = x.x; p2[4 * v + 1] = x.y; p2[4 * v + 2] = x.z; p2[4 * v + 3] = x.w;
        }
        const uint32_t tot1 = s_pre[(span - 1) * 4 + ((3 + ((span - 1) >> 1)) & 3)].w, tot2 = p2[SPT - 1];
        const uint32_t base = tot2 + tot1;
        // window of output j: samples (t-2, j+1) .. (t, j):  E_j = base + pre[j] - p2[j]   (exact in u32)
        int dmax_lo = INT_MIN, dmax_hi = INT_MIN;
#pragma unroll
        for (int j = 0; j < SPT; j++) {
            const int d = (int) (pre[j] - p2[j]);              // all sums < 2^31: signed difference is exact
            if (j < 8) dmax_lo = max(dmax_lo, d); else dmax_hi = max(dmax_hi, d);
        }
        const bool off_lo = (uint32_t) ((int) base + dmax_lo) < sp.k0, off_hi = (uint32_t) ((int) base + dmax_hi) < sp.k0;
        bool on = false;
        if (!(off_lo && off_hi)) {
            const int2 xy1 = s_xy[span - 1], xy2 = s_xy[span - 2];
            const float X = (float) (sx + xy1.x + xy2.x), Y = (float) (sy + xy1.y + xy2.y);
            const float Q = (float) (base + pre[SPT - 1]);
            const float m2 = fmaf(X, X, Y * Y);
            const float mu = sqrt_approx(m2) * sp.inv_n;
            const float V = fmaxf(fmaf(-m2, sp.inv_n, Q), 0.0f) + 1e-5f * Q;
            const float bc = fmaf(sp.t2, sqrt_approx(V), sp.cg * sqrt_approx(Q));
            on = fmaf(mu, sp.g_lo, -bc) * 0.99999f > sp.theta_hi;
        }
        if (on) {
            bits16 = 0xFFFFu;
        } else {
            undecided_lo = !off_lo;
            undecided_hi = !off_hi;
        }
    } else {
        undecided_lo = undecided_hi = true;
    }
    const bool in_lo = o < a.out_hi, in_hi = o + 8 < a.out_hi;
    if (in_lo) {
        // undecided groups are overwritten by fir1_refine_kernel
        const i64 byte = (o - a.bit_base) >> 3;              // even: 16 outputs per thread
        if (in_hi) {
            *(uint16_t *) (a.out_bits + byte) = (uint16_t) bits16;
        } else {
            a.out_bits[byte] = (uint8_t) bits16;
        }
    }
    // ---- undecided groups -> global work list (warp-aggregated reservation, no CTA barrier) ----
    const bool push_lo = undecided_lo && in_lo, push_hi = undecided_hi && in_hi;
    const uint32_t m_lo = __ballot_sync(0xFFFFFFFFu, push_lo), m_hi = __ballot_sync(0xFFFFFFFFu, push_hi);
    const uint32_t n_push = __popc(m_lo) + __popc(m_hi);
    if (n_push) {
        const int lane = threadIdx.x & 31;
        uint32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(sa.work_count, n_push);
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
        const uint32_t below = (1u << lane) - 1;
        const uint32_t grp0 = (uint32_t) ((o - a.bit_base) >> 3);
        if (push_lo) {
            const uint32_t s = slot0 + __popc(m_lo & below);
            if (s < sa.work_cap) sa.work_list[s] = grp0;
        }
        if (push_hi) {
            const uint32_t s = slot0 + __popc(m_lo) + __popc(m_hi & below);
            if (s < sa.work_cap) sa.work_list[s] = grp0 + 1;
        }
    }
}

// Persistent variant of fir1_screen_kernel: each CTA walks a contiguous range of tiles, requests the raw
// words of tile i+1 (4 x LDG.128 per thread) before it processes tile i, and keeps the span statistics in
// a two-tile ring in shared memory so that the last two spans of tile i are the history of tile i+1.
// One CTA barrier per tile; no global-load latency on the critical path.
template <int T>
__global__ void __launch_bounds__(256, OOKD_SCREEN_PERSIST_MINB)
fir1_screen_persist_kernel(const ScreenArgs sa, const ScreenParams sp)
{
    static_assert(T == 32, "window = exactly two 16-sample thread spans");
    constexpr int NT = 256, SPT = 16, L = NT * SPT;
    constexpr int RING = 2 * NT;
    // Rows RING.. : the last two spans of each tile again, in a ring of THREE tiles.  Threads 0 and 1 read
    // them after the barrier of tile i; a copy living in the two-tile ring would be overwritten by the
    // stores of tile i+1, which no barrier separates from those reads.
    constexpr int TAIL = RING, ROWS = RING + 3 * 2;
    __shared__ uint4 s_pre[ROWS * 4];
    __shared__ int2 s_xy[ROWS];
    __shared__ uint8_t s_flag[ROWS];

    const TiledArgs &a = sa.t;
    const uint32_t per = (sa.n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_begin = sa.tile_offset + blockIdx.x * per;
    const uint32_t t_end = min(sa.tile_offset + sa.n_tiles, t_begin + per);
    if (t_begin >= t_end) return;

    const bool ptr_ok = ((((uintptr_t) a.in) & 15) == 0);
    auto load_span = [&](i64 g, bool fast, uint32_t (&w)[16]) {
        if (fast) {
            const uint4 *src = (const uint4 *) (a.in + (g - a.in_base));
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 x = __ldg(src + v);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const i64 ge = g + e;
                w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
            }
        }
    };
    auto tile_fast = [&](uint32_t tile) -> bool {
        const i64 g0 = a.out_lo + (i64) tile * L;
        return ptr_ok && (((g0 - a.in_base) & 3) == 0) && g0 >= a.in_base && g0 >= 0 && (g0 + L) <= a.in_valid_end;
    };
    auto store_span = [&](int rp, const uint32_t (&p)[16], int xs, int ys, uint32_t gd) {
#pragma unroll
        for (int v = 0; v < 4; v++) {
            s_pre[rp * 4 + ((v + (rp >> 1)) & 3)] = make_uint4(p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]);
        }
        s_xy[rp] = make_int2(xs, ys);
        s_flag[rp] = (gd >> 25) ? 1 : 0;
    };

    // history of the first tile: the two spans in front of it go to the end of the "previous" half of the ring
    if (threadIdx.x < 2) {
        const i64 g = a.out_lo + (i64) t_begin * L - 2 * SPT + (i64) threadIdx.x * SPT;
        uint32_t w[16], p[16], gd;
        int xs, ys;
        load_span(g, false, w);
        screen_span_stats(w, p, xs, ys, gd);
        store_span(TAIL + 2 * 2 + threadIdx.x, p, xs, ys, gd);     // slot 2 = "previous" of slot 0
    }

    uint32_t w_cur[16], w_nxt[16];
    load_span(a.out_lo + (i64) t_begin * L + (i64) threadIdx.x * SPT, tile_fast(t_begin), w_cur);

    int slot = 0;                                   // (tile - t_begin) % 3
    for (uint32_t tile = t_begin; tile < t_end; tile++) {
        const int par = (int) (tile & 1);
        const int prev_slot = (slot == 0) ? 2 : slot - 1;
        const i64 o0 = a.out_lo + (i64) tile * L;
        if (tile + 1 < t_end) {
            load_span(o0 + L + (i64) threadIdx.x * SPT, tile_fast(tile + 1), w_nxt);
        }
        uint32_t pre[SPT], guard;
        int sx, sy;
        const int rp = par * NT + threadIdx.x;
        screen_span_stats(w_cur, pre, sx, sy, guard);
        store_span(rp, pre, sx, sy, guard);
        if (threadIdx.x >= NT - 2) store_span(TAIL + 2 * slot + (threadIdx.x - (NT - 2)), pre, sx, sy, guard);
        __syncthreads();     // rows of this half are next written two tiles (= two barriers) from now

        const int r1 = (threadIdx.x >= 1) ? rp - 1 : TAIL + 2 * prev_slot + 1;
        const int r2 = (threadIdx.x >= 2) ? rp - 2 : TAIL + 2 * prev_slot + (int) threadIdx.x;
        uint32_t bits16 = 0;
        bool undecided_lo = false, undecided_hi = false;
        const i64 o = o0 + (i64) threadIdx.x * SPT;
        const bool bad = (guard >> 25) || s_flag[r1] || s_flag[r2];
        if (!bad) {
            uint32_t p2[SPT];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 x = s_pre[r2 * 4 + ((v + (r2 >> 1)) & 3)];
                p2[4 * v] = x.x; p2[4 * v + 1] = x.y; p2[4 * v + 2] = x.z; p2[4 * v + 3] = x.w;
            }
            const uint32_t tot1 = s_pre[r1 * 4 + ((3 + (r1 >> 1)) & 3)].w, tot2 = p2[SPT - 1];
            const uint32_t base = tot2 + tot1;
            int dmax_lo = INT_MIN, dmax_hi = INT_MIN;
#pragma unroll
            for (int j = 0; j < SPT; j++) {
                const int d = (int) (pre[j] - p2[j]);          // all sums < 2^31: signed difference is exact
                if (j < 8) dmax_lo = max(dmax_lo, d); else dmax_hi = max(dmax_hi, d);
            }
            const bool off_lo = (uint32_t) ((int) base + dmax_lo) < sp.k0, off_hi = (uint32_t) ((int) base + dmax_hi) < sp.k0;
            bool on = false;
            if (!(off_lo && off_hi)) {
                const int2 xy1 = s_xy[r1], xy2 = s_xy[r2];
                const float X = (float) (sx + xy1.x + xy2.x), Y = (float) (sy + xy1.y + xy2.y);
                const float Q = (float) (base + pre[SPT - 1]);
                const float m2 = fmaf(X, X, Y * Y);
                const float mu = sqrt_approx(m2) * sp.inv_n;
                const float V = fmaxf(fmaf(-m2, sp.inv_n, Q), 0.0f) + 1e-5f * Q;
                const float bc = fmaf(sp.t2, sqrt_approx(V), sp.cg * sqrt_approx(Q));
                on = fmaf(mu, sp.g_lo, -bc) * 0.99999f > sp.theta_hi;
            }
            if (on) {
                bits16 = 0xFFFFu;
            } else {
                undecided_lo = !off_lo;
                undecided_hi = !off_hi;
            }
        } else {
            undecided_lo = undecided_hi = true;
        }
        const bool in_lo = o < a.out_hi, in_hi = o + 8 < a.out_hi;
        if (in_lo) {
            const i64 byte = (o - a.bit_base) >> 3;
            if (in_hi) {
                *(uint16_t *) (a.out_bits + byte) = (uint16_t) bits16;
            } else {
                a.out_bits[byte] = (uint8_t) bits16;
            }
        }
        const bool push_lo = undecided_lo && in_lo, push_hi = undecided_hi && in_hi;
        const uint32_t m_lo = __ballot_sync(0xFFFFFFFFu, push_lo), m_hi = __ballot_sync(0xFFFFFFFFu, push_hi);
        const uint32_t n_push = __popc(m_lo) + __popc(m_hi);
        if (n_push) {
            const int lane = threadIdx.x & 31;
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(sa.work_count, n_push);
            slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
            const uint32_t below = (1u << lane) - 1;
            const uint32_t grp0 = (uint32_t) ((o - a.bit_base) >> 3);
            if (push_lo) {
                const uint32_t sl = slot0 + __popc(m_lo & below);
                if (sl < sa.work_cap) sa.work_list[sl] = grp0;
            }
            if (push_hi) {
                const uint32_t sl = slot0 + __popc(m_lo) + __popc(m_hi & below);
                if (sl < sa.work_cap) sa.work_list[sl] = grp0 + 1;
            }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) w_cur[e] = w_nxt[e];
        slot = (slot == 2) ? 0 : slot + 1;
    }
}

// Exact recomputation of the groups the screen left undecided: one output per lane (8 lanes per group),
// samples re-read through L1/L2.  Grid-stride over the global work list, so lanes stay full whatever the
// distribution of undecided groups over the capture.
template <int T>
__global__ void __launch_bounds__(256) fir1_refine_kernel(const ScreenArgs sa, const TapsParam<T> taps)
{
    const TiledArgs &a = sa.t;
    const uint32_t n_groups = min(*sa.work_count, sa.work_cap);
    const u64 n_items = ((u64) n_groups * 8 + 31) & ~31ull;
    for (u64 item = (u64) blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (u64) gridDim.x * blockDim.x) {
        const u64 qi = item >> 3;
        const uint32_t j = (uint32_t) (item & 7);
        bool bit = false;
        uint32_t grp = 0;
        if (qi < n_groups) {
            grp = sa.work_list[qi];
            const i64 n = a.bit_base + (i64) grp * 8 + j;    // output index == index of its newest sample
            float re = 0.0f, im = 0.0f;
            if (n - (T - 1) >= a.in_base && n - (T - 1) >= 0 && n < a.in_valid_end) {
                const uint32_t *src = a.in + (n - a.in_base);   // whole window present: no per-tap checks
#pragma unroll
                for (int i = 0; i < T; i++) {
                    const float2 x = sc16q11_to_float2(__ldg(src - i));
                    re = mac_exact(re, taps.t[i], x.x);
                    im = mac_exact(im, taps.t[i], x.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < T; i++) {
                    const i64 g = n - i;
                    const uint32_t w = (g >= 0 && g >= a.in_base && g < a.in_valid_end) ? __ldg(a.in + (g - a.in_base)) : 0u;
                    const float2 x = sc16q11_to_float2(w);
                    re = mac_exact(re, taps.t[i], x.x);
                    im = mac_exact(im, taps.t[i], x.y);
                }
            }
            bit = power_exact(re, im) >= a.pstar;
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
        if (qi < n_groups && j == 0) {
            a.out_bits[grp] = (uint8_t) (ballot >> (threadIdx.x & 24));
        }
    }
}

// Same job, one THREAD per undecided group of 8 outputs: the 39 samples the group's windows cover are read
// once (ten 16-byte loads when the input is 16-byte aligned), converted once and kept in registers; 8
// outputs = 16 independent exact accumulator chains per thread.  ~3x fewer instructions per group than
// the lane-per-output form above and no redundant loads.
template <int T>
__global__ void __launch_bounds__(128) fir1_refine_group_kernel(const ScreenArgs sa, const TapsParam<T> taps)
{
    static_assert(T == 32, "window of 8 outputs = 39 samples, fetched as 40");
    const TiledArgs &a = sa.t;
    const uint32_t n_groups = min(*sa.work_count, sa.work_cap);
    const bool ptr_ok = ((((uintptr_t) a.in) & 15) == 0);
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n_groups; qi += gridDim.x * blockDim.x) {
        const uint32_t grp = sa.work_list[qi];
        const i64 n0 = a.bit_base + (i64) grp * 8;             // first output of the group
        const i64 g0 = n0 - T;                                 // first sample fetched (one more than needed)
        float2 win[T + 8];
        if (ptr_ok && (((g0 - a.in_base) & 3) == 0) && g0 >= a.in_base && g0 >= 0 && g0 + (T + 8) <= a.in_valid_end) {
            const uint4 *src = (const uint4 *) (a.in + (g0 - a.in_base));
#pragma unroll
            for (int v = 0; v < (T + 8) / 4; v++) {
                const uint4 x = __ldg(src + v);
                win[4 * v] = sc16q11_to_float2(x.x);
                win[4 * v + 1] = sc16q11_to_float2(x.y);
                win[4 * v + 2] = sc16q11_to_float2(x.z);
                win[4 * v + 3] = sc16q11_to_float2(x.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < T + 8; q++) {
                const i64 g = g0 + q;
                win[q] = sc16q11_to_float2((g >= 0 && g >= a.in_base && g < a.in_valid_end) ? __ldg(a.in + (g - a.in_base)) : 0u);
            }
        }
        float re[8], im[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            re[j] = 0.0f;
            im[j] = 0.0f;
        }
        // output n0 + j: newest sample is win[T + j], tap i multiplies win[T + j - i]
#pragma unroll
        for (int i = 0; i < T; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                re[j] = mac_exact(re[j], taps.t[i], win[T + j - i].x);
                im[j] = mac_exact(im[j], taps.t[i], win[T + j - i].y);
            }
        }
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            bits |= (power_exact(re[j], im[j]) >= a.pstar ? 1u : 0u) << j;
        }
        a.out_bits[grp] = (uint8_t) bits;
    }
}

// =======================================================================================
// 4. Two-stage shape: 16 taps / decimate 2, then 32 taps / decimate 2 (fs128_fs16_dec4).
//
//    Final output m is produced when input 4m+3 arrives and depends on the 78 inputs 4m+3-77 .. 4m+3
//    through the composite response h = t1 (*) upsample2(t2).  The same two proofs as in section 3
//    apply with ||h||_2, G = sum h and, for the reference's two rounded stages, the absolute composite
//    habs = |t1| (*) upsample2(|t2|) in the rounding allowance:
//        |y_ref - y| <= (g1 + g2 + g1 g2) sum_n habs[n] |x[4m+3-n]| <= gamma ||habs||_2 sqrt(E).
//    Threads own 16 inputs = 4 outputs; a window reaches back into the 5 spans in front, of which it
//    needs only the prefix sums at elements 1, 5, 9, 13 and the totals.
// =======================================================================================
struct Taps2Param {
    float t1[16];
    float t2[32];
    const float *d_t1, *d_t2;       // the same taps in device memory (rolled-loop boundary path)
};

__global__ void __launch_bounds__(256, 5) fir2_screen_kernel(const ScreenArgs sa, const ScreenParams sp)
{
    constexpr int NT = 256, SPT = 16, LIN = NT * SPT;     // 4096 inputs = 1024 outputs per tile
    constexpr int HT = 5;                                 // history spans in front of the tile (80 samples >= 77)
    __shared__ uint4 s_a[NT + HT];                        // pre[1], pre[5], pre[9], pre[13]
    __shared__ uint4 s_b[NT + HT];                        // total, sum I, sum Q, flag

    const TiledArgs &a = sa.t;                            // out_lo / out_hi / bit_base in OUTPUT indices
    const uint32_t tile = blockIdx.x + sa.tile_offset;
    const i64 o0 = a.out_lo + (i64) tile * (LIN / 4);     // first output of the tile
    const i64 g0 = o0 * 4 - HT * SPT;                     // first input of the first history span
    const bool fast = ((((uintptr_t) a.in) & 15) == 0) && (((g0 - a.in_base) & 3) == 0) && g0 >= a.in_base && g0 >= 0 &&
                      (g0 + LIN + HT * SPT) <= a.in_valid_end;

    uint32_t pre[SPT], guard = 0;
    int sx = 0, sy = 0;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        int span;
        if (pass == 0) {
            span = threadIdx.x + HT;
        } else {
            if (threadIdx.x >= HT) break;
            span = threadIdx.x;
        }
        const i64 g = g0 + (i64) span * SPT;
        uint32_t w[SPT];
        if (fast) {
            const uint4 *src = (const uint4 *) (a.in + (g - a.in_base));
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 x = __ldg(src + v);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < SPT; e++) {
                const i64 ge = g + e;
                w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
            }
        }
        uint32_t p[SPT], gd;
        int xs, ys;
        screen_span_stats(w, p, xs, ys, gd);
        s_a[span] = make_uint4(p[1], p[5], p[9], p[13]);
        s_b[span] = make_uint4(p[15], (uint32_t) xs, (uint32_t) ys, (gd >> 25) ? 1u : 0u);
        if (pass == 0) {
#pragma unroll
            for (int e = 0; e < SPT; e++) pre[e] = p[e];
            sx = xs; sy = ys; guard = gd;
        }
    }
    __syncthreads();

    const int span = threadIdx.x + HT;
    const uint4 b1 = s_b[span - 1], b2 = s_b[span - 2], b3 = s_b[span - 3], b4 = s_b[span - 4], b5 = s_b[span - 5];
    const uint4 a4 = s_a[span - 4], a5 = s_a[span - 5];
    uint32_t bits4 = 0, und = 0;                          // decisions / undecided flags of the 4 outputs
    const bool bad = (guard >> 25) || b1.w || b2.w || b3.w || b4.w || b5.w;
    if (!bad) {
        const uint32_t mid3 = b3.x + b2.x + b1.x;         // spans s-3 .. s-1
        const uint32_t mid4 = mid3 + b4.x;                // spans s-4 .. s-1
        // newest sample at element 3, 7, 11: window starts at element 6, 10, 14 of span s-5
        const uint32_t e0 = (b5.x - a5.y) + mid4 + pre[3];
        const uint32_t e1 = (b5.x - a5.z) + mid4 + pre[7];
        const uint32_t e2 = (b5.x - a5.w) + mid4 + pre[11];
        // newest sample at element 15: window starts at element 2 of span s-4
        const uint32_t e3 = (b4.x - a4.x) + mid3 + pre[15];
        und = (e0 < sp.k0 ? 0u : 1u) | (e1 < sp.k0 ? 0u : 2u) | (e2 < sp.k0 ? 0u : 4u) | (e3 < sp.k0 ? 0u : 8u);
        if (und) {
            const float X = (float) (sx + (int) b1.y + (int) b2.y + (int) b3.y + (int) b4.y + (int) b5.y);
            const float Y = (float) (sy + (int) b1.z + (int) b2.z + (int) b3.z + (int) b4.z + (int) b5.z);
            const float Q = (float) (b5.x + mid4 + pre[15]);
            const float m2 = fmaf(X, X, Y * Y);
            const float mu = sqrt_approx(m2) * sp.inv_n;
            const float V = fmaxf(fmaf(-m2, sp.inv_n, Q), 0.0f) + 1e-5f * Q;
            const float bc = fmaf(sp.t2, sqrt_approx(V), sp.cg * sqrt_approx(Q));
            if (fmaf(mu, sp.g_lo, -bc) * 0.99999f > sp.theta_hi) {
                bits4 = 0xF;
                und = 0;
            }
        }
    } else {
        und = 0xF;
    }
    // two threads share a byte of decisions
    const uint32_t other_bits = __shfl_down_sync(0xFFFFFFFFu, bits4, 1);
    const uint32_t other_und = __shfl_down_sync(0xFFFFFFFFu, und, 1);
    const i64 o = o0 + (i64) threadIdx.x * 4;             // first output of this thread
    const bool even = (threadIdx.x & 1) == 0;
    const bool in_range = even && o < a.out_hi;
    const bool push = in_range && ((und | other_und) != 0);
    if (in_range) {
        a.out_bits[(o - a.bit_base) >> 3] = (uint8_t) (bits4 | (other_bits << 4));
    }
    const uint32_t m_push = __ballot_sync(0xFFFFFFFFu, push);
    if (m_push) {
        const int lane = threadIdx.x & 31;
        uint32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(sa.work_count, (uint32_t) __popc(m_push));
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
        if (push) {
            const uint32_t sl = slot0 + __popc(m_push & ((1u << lane) - 1));
            if (sl < sa.work_cap) sa.work_list[sl] = (uint32_t) ((o - a.bit_base) >> 3);
        }
    }
}

// Exact two-stage recomputation of one output per lane (8 lanes per undecided group).
//   y[m]  = sum_{j<32} t2[j] * s[2m+1-j]          accumulated from j = 0   (src/fir.c:313-318, stage 2)
//   s[u]  = sum_{i<16} t1[i] * x[2u+1-i]          accumulated from i = 0   (stage 1); s[u<0] = 0, x[g<0] = 0
// A 16-sample register window slides down by two inputs per stage-1 output (fully unrolled).
__global__ void __launch_bounds__(256) fir2_refine_kernel(const ScreenArgs sa, const Taps2Param taps)
{
    const TiledArgs &a = sa.t;
    const uint32_t n_groups = min(*sa.work_count, sa.work_cap);
    const u64 n_items = ((u64) n_groups * 8 + 31) & ~31ull;
    for (u64 item = (u64) blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (u64) gridDim.x * blockDim.x) {
        const u64 qi = item >> 3;
        const uint32_t jj = (uint32_t) (item & 7);
        bool bit = false;
        uint32_t grp = 0;
        if (qi < n_groups) {
            grp = sa.work_list[qi];
            const i64 m = a.bit_base + (i64) grp * 8 + jj;
            const i64 g_new = 4 * m + 3;                   // newest input of output m
            float re = 0.0f, im = 0.0f;
            if (g_new - 77 >= a.in_base && g_new - 77 >= 0 && g_new < a.in_valid_end) {
                const uint32_t *src = a.in + (g_new - a.in_base);
                float2 w[16];                              // w[k] = x[2u+1-k] for the current u
#pragma unroll
                for (int k = 0; k < 16; k++) w[k] = sc16q11_to_float2(__ldg(src - k));
#pragma unroll
                for (int j = 0; j < 32; j++) {             // u = 2m+1-j
                    float sr = 0.0f, si = 0.0f;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        sr = mac_exact(sr, taps.t1[i], w[i].x);
                        si = mac_exact(si, taps.t1[i], w[i].y);
                    }
                    re = mac_exact(re, taps.t2[j], sr);
                    im = mac_exact(im, taps.t2[j], si);
                    if (j < 31) {
#pragma unroll
                        for (int k = 0; k < 14; k++) w[k] = w[k + 2];
                        w[14] = sc16q11_to_float2(__ldg(src - (2 * j + 16)));
                        w[15] = sc16q11_to_float2(__ldg(src - (2 * j + 17)));
                    }
                }
            } else {
                for (int j = 0; j < 32; j++) {
                    const i64 u = 2 * m + 1 - j;
                    float sr = 0.0f, si = 0.0f;
                    if (u >= 0) {
                        for (int i = 0; i < 16; i++) {
                            const i64 g = 2 * u + 1 - i;
                            const uint32_t wv = (g >= 0 && g >= a.in_base && g < a.in_valid_end) ? __ldg(a.in + (g - a.in_base)) : 0u;
                            const float2 x = sc16q11_to_float2(wv);
                            const float t = __ldg(taps.d_t1 + i);
                            sr = mac_exact(sr, t, x.x);
                            si = mac_exact(si, t, x.y);
                        }
                    }
                    const float t = __ldg(taps.d_t2 + j);
                    re = mac_exact(re, t, sr);
                    im = mac_exact(im, t, si);
                }
            }
            bit = power_exact(re, im) >= a.pstar;
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
        if (qi < n_groups && jj == 0) {
            a.out_bits[grp] = (uint8_t) (ballot >> (threadIdx.x & 24));
        }
    }
}

// Tiled EXACT kernel for the two-stage shape (every output computed with the reference's in-order MACs): the path
// for captures where the screen cannot decide anything (low SNR) and for OOKD_FLAG_NO_SCREEN.
// CTA = 288 threads, tile = 512 final outputs = 2048 inputs.
//   stage inputs  : [4 m0 - 80, 4 m0 + 2048) converted once into shared memory (float2, one pad slot per 8 samples:
//                   thread stride 9 float2 => conflict-free 64-bit loads)
//   phase 1       : thread t < 264 computes the stage-1 outputs s[2 m0 - 30 + 4t .. + 3] from a 22-sample register
//                   window (16 taps, decimation 2) and stores them (one pad slot per 4: thread stride 5 float2)
//   phase 2       : thread t < 256 computes y[m0 + 2t], y[m0 + 2t + 1] from a 34-value window of s (32 taps, decimation 2)
// ~77 instructions per input sample against the 64 the arithmetic itself needs.
constexpr int F2X_NT = 288, F2X_M = 512, F2X_IN = 4 * F2X_M + 80, F2X_NS = 2 * F2X_M + 30;

template <bool FMA>
__global__ void __launch_bounds__(F2X_NT, 2) fir2_tiled_kernel(const ScreenArgs sa, const Taps2Param taps, const FmaBand band)
{
    __shared__ float2 s_x[F2X_IN + F2X_IN / 8 + 8];        // (+8: the last phase-1 thread reads a few slots past its valid window)
    __shared__ float2 s_s[(F2X_NS + 3) / 4 * 5 + 1];
    __shared__ float s_m2[F2X_NT / 32];
    const TiledArgs &a = sa.t;
    const int t = (int) threadIdx.x;
    const i64 m0 = a.out_lo + (i64) blockIdx.x * F2X_M;       // first output of the tile (a.out_lo % 8 == a.bit_base % 8)
    const i64 g0 = 4 * m0 - 80;                               // first staged input

    // ---- stage inputs: 4 samples (16 B) per thread per step ----
    const bool aligned = (((g0 - a.in_base) & 3) == 0) && ((((uintptr_t) a.in) & 15) == 0);
    float m2 = 0.0f;
    for (int q = t; q < F2X_IN / 4; q += F2X_NT) {
        const i64 g = g0 + 4 * q;
        uint32_t w[4];
        if (aligned && g >= a.in_base && g >= 0 && g + 4 <= a.in_valid_end) {
            const uint4 v = __ldg((const uint4 *) (a.in + (g - a.in_base)));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const i64 ge = g + e;
                w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int x = 4 * q + e;
            const float2 v = sc16q11_to_float2(w[e]);
            s_x[x + (x >> 3)] = v;
            if constexpr (FMA) m2 = fmaxf(m2, fmaf(v.x, v.x, v.y * v.y));
        }
    }
    if constexpr (FMA) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xFFFFFFFFu, m2, d));
        if ((t & 31) == 0) s_m2[t >> 5] = m2;
    }
    __syncthreads();

    // ---- phase 1: stage-1 outputs v = 4t .. 4t+3 (u = 2 m0 - 30 + v); s[u] = sum_i t1[i] x[2u+1-i], local x = 2v + 21 - i ----
    if (4 * t < F2X_NS) {
        float2 win[22];                                       // local x 8t + 6 .. 8t + 27
#pragma unroll
        for (int q = 0; q < 22; q++) {
            const int x = 8 * t + 6 + q;
            win[q] = s_x[x + (x >> 3)];
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float sr = 0.0f, si = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                sr = mac_sel<FMA>(sr, taps.t1[i], win[2 * c + 15 - i].x);       // local 2v+21-i = 8t + 6 + (2c + 15 - i)
                si = mac_sel<FMA>(si, taps.t1[i], win[2 * c + 15 - i].y);
            }
            const int v = 4 * t + c;
            if (2 * m0 - 30 + v < 0) {                        // stage-1 outputs before the capture: the zeros of fir_reset
                sr = 0.0f;
                si = 0.0f;
            }
            if (v < F2X_NS) s_s[5 * t + c] = make_float2(sr, si);           // v + (v >> 2) == 5t + c
        }
    }
    __syncthreads();

    // ---- phase 2: y[m0 + 2t + q] = sum_j t2[j] s[2m+1-j], local v = 4t + 2q + 31 - j ----
    uint32_t bits2 = 0, unsure = 0;
    if (t < F2X_M / 2) {
        float2 sw[34];                                        // local v 4t .. 4t + 33
#pragma unroll
        for (int q = 0; q < 34; q++) {
            const int v = 4 * t + q;
            sw[q] = s_s[v + (v >> 2)];
        }
        float hi2 = 0.0f, lo2 = 0.0f;
        if constexpr (FMA) {
            float mm = s_m2[0];
#pragma unroll
            for (int q = 1; q < F2X_NT / 32; q++) mm = fmaxf(mm, s_m2[q]);
            const float D = band.c * sqrt_approx(mm) * 1.00001f;
            const float hi = band.theta_hi + D, lo = fmaxf(band.theta_lo - D, 0.0f);
            hi2 = hi * hi * 1.000001f;
            lo2 = lo * lo * 0.999999f;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            float re = 0.0f, im = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                re = mac_sel<FMA>(re, taps.t2[j], sw[2 * q + 31 - j].x);
                im = mac_sel<FMA>(im, taps.t2[j], sw[2 * q + 31 - j].y);
            }
            if constexpr (FMA) {
                const uint32_t c = fma_classify(re, im, hi2, lo2);
                bits2 |= (c & 1u) << q;
                unsure |= (c >> 1);
            } else {
                bits2 |= (power_exact(re, im) >= a.pstar ? 1u : 0u) << q;
            }
        }
    }
    // four threads share a byte of decisions (warps 0..7 are complete; warp 8 has no outputs)
    if (t < F2X_M / 2) {
        const uint32_t b1 = __shfl_down_sync(0xFFFFFFFFu, bits2, 1);
        const uint32_t b2 = __shfl_down_sync(0xFFFFFFFFu, bits2, 2);
        const uint32_t b3 = __shfl_down_sync(0xFFFFFFFFu, bits2, 3);
        const i64 o = m0 + 2 * t;
        const bool writer = (t & 3) == 0 && o < a.out_hi;
        if (writer) {
            // outputs past out_hi inside the last byte are masked by the consumers
            a.out_bits[(o - a.bit_base) >> 3] = (uint8_t) (bits2 | (b1 << 2) | (b2 << 4) | (b3 << 6));
        }
        if constexpr (FMA) {
            const uint32_t u1 = __shfl_down_sync(0xFFFFFFFFu, unsure, 1);
            const uint32_t u2 = __shfl_down_sync(0xFFFFFFFFu, unsure, 2);
            const uint32_t u3 = __shfl_down_sync(0xFFFFFFFFu, unsure, 3);
            const bool push = writer && ((unsure | u1 | u2 | u3) != 0);
            const uint32_t m_push = __ballot_sync(0xFFFFFFFFu, push);
            if (m_push) {
                const int lane = t & 31;
                uint32_t slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(sa.work_count, (uint32_t) __popc(m_push));
                slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
                if (push) {
                    const uint32_t sl = slot0 + __popc(m_push & ((1u << lane) - 1));
                    if (sl < sa.work_cap) sa.work_list[sl] = (uint32_t) ((o - a.bit_base) >> 3);
                }
            }
        }
    }
}

// Same job, one THREAD per undecided group of 8 outputs m0..m0+7.  The 46 stage-1 outputs s[2 m0 + 15] .. s[2 m0 - 30]
// the group needs are computed ONCE each (newest first, from a 16-sample register window that slides down by two
// inputs per step) and fed to every output y[m] they belong to: walking u downwards visits the taps of each y[m] in
// the reference's order j = 0, 1, ... (src/fir.c:313-318), so all sums stay bit-exact; 4.4x fewer MACs than the
// lane-per-output form.
__global__ void __launch_bounds__(128) fir2_refine_group_kernel(const ScreenArgs sa, const Taps2Param taps)
{
    const TiledArgs &a = sa.t;
    const uint32_t n_groups = min(*sa.work_count, sa.work_cap);
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n_groups; qi += gridDim.x * blockDim.x) {
        const uint32_t grp = sa.work_list[qi];
        const i64 m0 = a.bit_base + (i64) grp * 8;             // first output of the group
        const i64 u_top = 2 * m0 + 15;                         // newest stage-1 output needed (by y[m0+7], j = 0)
        const i64 g_top = 2 * u_top + 1;                       // newest input: 4 m0 + 31
        const i64 g_bot = 2 * (u_top - 45) + 1 - 15;           // oldest input: 4 m0 - 74
        const bool inside = g_bot >= a.in_base && g_bot >= 0 && g_top < a.in_valid_end;
        auto sample = [&](i64 g) -> float2 {
            if (inside) return sc16q11_to_float2(__ldg(a.in + (g - a.in_base)));
            return sc16q11_to_float2((g >= 0 && g >= a.in_base && g < a.in_valid_end) ? __ldg(a.in + (g - a.in_base)) : 0u);
        };
        float2 w[16];                                          // w[k] = x[2u+1-k] for the current u
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = sample(g_top - k);
        float re[8], im[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            re[q] = 0.0f;
            im[q] = 0.0f;
        }
#pragma unroll
        for (int d = 0; d < 46; d++) {                         // u = u_top - d
            float sr = 0.0f, si = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                sr = mac_exact(sr, taps.t1[i], w[i].x);
                si = mac_exact(si, taps.t1[i], w[i].y);
            }
            // stage-1 outputs with a negative index are the zeros fir_reset leaves in the delay line (+0.0 exactly;
            // a sum of products of zero inputs could be -0.0, which the reference never sees there)
            if (u_top - d < 0) {
                sr = 0.0f;
                si = 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                constexpr int dummy = 0;
                (void) dummy;
                const int j = 2 * q + 1 - (15 - d);            // tap of y[m0+q] this s[u] multiplies: j = 2(m0+q)+1-u
                if (j >= 0 && j < 32) {
                    re[q] = mac_exact(re[q], taps.t2[j], sr);
                    im[q] = mac_exact(im[q], taps.t2[j], si);
                }
            }
            if (d < 45) {
#pragma unroll
                for (int k = 0; k < 14; k++) w[k] = w[k + 2];
                w[14] = sample(g_top - (2 * d + 16));
                w[15] = sample(g_top - (2 * d + 17));
            }
        }
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            bits |= (power_exact(re[q], im[q]) >= a.pstar ? 1u : 0u) << q;
        }
        a.out_bits[grp] = (uint8_t) bits;
    }
}

}  // namespace ookd
