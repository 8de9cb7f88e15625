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

template <int T, int R>
__global__ void __launch_bounds__(256, 2)
fir1_exact_tiled_kernel(const TiledArgs a, const TapsParam<T> taps)
{
    constexpr int NT = 256;
    constexpr int L = NT * R;                 // outputs per tile
    constexpr int HALO = (T - 1 + 3) & ~3;    // history samples staged in front, multiple of 4
    constexpr int NS = L + HALO;              // staged samples
    constexpr int LOGR = (R == 8) ? 3 : 4;
    static_assert(R == 8 || R == 16, "R");
    __shared__ float2 s_x[NS + (NS >> LOGR) + 1];

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
                s_x[s + (s >> LOGR)] = sc16q11_to_float2(w[e]);
            }
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
                re[j] = mac_exact(re[j], taps.t[i], win[j + T - 1 - i].x);
                im[j] = mac_exact(im[j], taps.t[i], win[j + T - 1 - i].y);
            }
        }

        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < R; j++) {
            bits |= (power_exact(re[j], im[j]) >= a.pstar ? 1u : 0u) << j;
        }
        const i64 o = o0 + (i64) threadIdx.x * R;
        if (o < a.out_hi) {
            // outputs past out_hi inside the last byte are masked by the consumers
            const i64 byte = (o - a.bit_base) >> 3;
            a.out_bits[byte] = (uint8_t) bits;
            if (R == 16) {
                a.out_bits[byte + 1] = (uint8_t) (bits >> 8);
            }
        }
        __syncthreads();     // s_x is reused by the next tile of this CTA
    }
}


// =======================================================================================
// 3. Screening kernel, one stage, decimation 1, T <= 32 taps, 8 outputs per thread.
//
//    The decision the reference takes for output n is  fl(re^2)+fl(im^2) >= P*  on the fp32
//    in-order sums re, im.  Computing those sums costs 4T flops per sample, which on a B200 is
//    ~6x more issue slots than the HBM read of the 4-byte sample allows.  But an OOK capture
//    is mostly "carrier clearly on" or "carrier clearly off", and that can be PROVED per block
//    of outputs from three sums over the samples S its windows touch (N = |S| = 40):
//        X = sum x,  Q2 = sum |x|^2,   mu = X/N,   V = Q2 - |X|^2/N = sum |x - mu|^2
//        y[n] = mu*G + sum_i t_i (x[n-i] - mu),    |sum_i t_i z_i| <= ||t||_2 sqrt(V)   (Cauchy-Schwarz)
//    so  | |y[n]| - |mu||G| | <= ||t||_2 sqrt(V)  for all 8 outputs of the block, in exact arithmetic.
//    The reference's rounding moves |y| by at most  gamma ||t||_2 sqrt(Q2)  (gamma ~ (T+1) 2^-24) and
//    its power by 3 ulp.  With every slack rounded the safe way (ScreenParams, set on the host
//    in double) the block is decided without a single MAC when the interval misses the
//    threshold; otherwise its 8 outputs are recomputed with the exact in-order MACs (one output
//    per lane, compacted through a shared-memory queue so that lanes stay full).  Decisions are
//    therefore bit-identical to the exact kernel for EVERY input; only the cost is data dependent.
//    Tiles with too many undecided blocks are handed to fir1_exact_tiled_kernel via tile_list.
// =======================================================================================
struct ScreenParams {
    float g_hi, g_lo;        // |sum t_i| rounded up / down
    float t2;                // ||t||_2 rounded up
    float cg;                // gamma * ||t||_2 rounded up
    float theta_lo, theta_hi;// sqrt(P*) shrunk / grown by 1e-5
    float inv_n;             // 1/N
    uint32_t dense_limit;    // undecided blocks per tile above which the tile goes to the dense list
};

struct ScreenArgs {
    TiledArgs t;             // t.tile_list / t.tile_count here are OUTPUT: the dense list + its counter
    uint32_t *dense_list;
    uint32_t *dense_count;
    uint32_t *stat_refined;  // [0] blocks refined in place, [1] tiles sent to the dense list
    uint32_t tile_offset;    // first tile of this launch (tiles are numbered from t.out_lo)
};

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // max relative error 2^-23
    return r;
}

template <int T>
__global__ void __launch_bounds__(256, 4)
fir1_screen_kernel(const ScreenArgs sa, const ScreenParams sp, const TapsParam<T> taps)
{
    constexpr int NT = 256, R = 8, L = NT * R;
    constexpr int HB = (T - 1 + R - 1) / R;           // history blocks in front of the tile (4 for T=32)
    constexpr int HALO = HB * R;
    constexpr int NS = L + HALO;
    constexpr int NB = NT + HB;                        // blocks with statistics
    __shared__ float2 s_x[NS + (NS >> 3) + 1];
    __shared__ float s_sx[NB], s_sy[NB], s_sq[NB];
    __shared__ uint16_t s_queue[NT];
    __shared__ uint32_t s_nq;

    const TiledArgs &a = sa.t;
    const uint32_t tile = blockIdx.x + sa.tile_offset;
    const i64 o0 = a.out_lo + (i64) tile * L;
    const i64 g0 = o0 - HALO;
    if (threadIdx.x == 0) s_nq = 0;

    // ---- load + convert + per-block statistics: thread t owns block t+HB; threads < HB also a halo block ----
    const bool aligned = (((g0 - a.in_base) & 3) == 0) && ((((uintptr_t) a.in) & 15) == 0);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        int blk;
        if (pass == 0) {
            blk = threadIdx.x + HB;
        } else {
            if (threadIdx.x >= HB) break;
            blk = threadIdx.x;
        }
        const i64 g = g0 + (i64) blk * R;
        uint32_t w[R];
        if (aligned && g >= a.in_base && g >= 0 && g + R <= a.in_valid_end) {
            const uint4 v0 = __ldg((const uint4 *) (a.in + (g - a.in_base)));
            const uint4 v1 = __ldg((const uint4 *) (a.in + (g - a.in_base) + 4));
            w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
            w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        } else {
#pragma unroll
            for (int e = 0; e < R; e++) {
                const i64 ge = g + e;
                w[e] = (ge >= 0 && ge >= a.in_base && ge < a.in_valid_end) ? __ldg(a.in + (ge - a.in_base)) : 0u;
            }
        }
        float sx = 0.0f, sy = 0.0f, sq = 0.0f;
        float2 *dst = &s_x[blk * R + blk];             // s + (s >> 3) with s = blk*8 + e
#pragma unroll
        for (int e = 0; e < R; e++) {
            const float2 x = sc16q11_to_float2(w[e]);
            dst[e] = x;
            sx += x.x;                                 // exact: multiples of 2^-11, |sum| < 2^10
            sy += x.y;
            sq = fmaf(x.x, x.x, sq);
            sq = fmaf(x.y, x.y, sq);
        }
        s_sx[blk] = sx;
        s_sy[blk] = sy;
        s_sq[blk] = sq;
    }
    __syncthreads();

    // ---- classify the 8 outputs of block t+HB from the statistics of blocks t .. t+HB ----
    {
        float X = 0.0f, Y = 0.0f, Q2 = 0.0f;
#pragma unroll
        for (int d = 0; d <= HB; d++) {
            X += s_sx[threadIdx.x + d];
            Y += s_sy[threadIdx.x + d];
            Q2 += s_sq[threadIdx.x + d];
        }
        const float m2 = fmaf(X, X, Y * Y);
        const float mu = sqrt_approx(m2) * sp.inv_n;
        const float V = fmaxf(fmaf(-m2, sp.inv_n, Q2), 0.0f) + 2e-5f * Q2;
        const float bc = fmaf(sp.t2, sqrt_approx(V), sp.cg * sqrt_approx(Q2));
        const bool all0 = (fmaf(mu, sp.g_hi, bc)) * 1.00001f < sp.theta_lo;
        const bool all1 = (fmaf(mu, sp.g_lo, -bc)) * 0.99999f > sp.theta_hi;
        const i64 o = o0 + (i64) threadIdx.x * R;
        if (all0 || all1) {
            if (o < a.out_hi) {
                a.out_bits[(o - a.bit_base) >> 3] = all1 ? 0xFF : 0x00;
            }
        } else if (o < a.out_hi) {
            const uint32_t q = atomicAdd(&s_nq, 1u);
            s_queue[q] = (uint16_t) threadIdx.x;
        }
    }
    __syncthreads();

    const uint32_t nq = s_nq;
    if (nq == 0) return;
    if (nq > sp.dense_limit) {
        if (threadIdx.x == 0) {
            const uint32_t slot = atomicAdd(sa.dense_count, 1u);
            sa.dense_list[slot] = tile;
            atomicAdd(&sa.stat_refined[1], 1u);
        }
        return;
    }
    if (threadIdx.x == 0) atomicAdd(&sa.stat_refined[0], nq);

    // ---- exact recomputation of the undecided blocks, one output per lane ----
    for (uint32_t item = threadIdx.x; item < ((nq * R + 31) & ~31u); item += NT) {
        const uint32_t qi = item >> 3, j = item & 7;
        bool bit = false;
        uint32_t blk_t = 0;
        if (qi < nq) {
            blk_t = s_queue[qi];
            const int s_new = (blk_t + HB) * R + j;    // staged index of the newest sample of this output
            float re = 0.0f, im = 0.0f;
#pragma unroll
            for (int i = 0; i < T; i++) {
                const int s = s_new - i;
                const float2 x = s_x[s + (s >> 3)];
                re = mac_exact(re, taps.t[i], x.x);
                im = mac_exact(im, taps.t[i], x.y);
            }
            bit = power_exact(re, im) >= a.pstar;
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
        if (qi < nq && j == 0) {
            const i64 o = o0 + (i64) blk_t * R;
            a.out_bits[(o - a.bit_base) >> 3] = (uint8_t) (ballot >> (threadIdx.x & 24));
        }
    }
}

}  // namespace ookd
