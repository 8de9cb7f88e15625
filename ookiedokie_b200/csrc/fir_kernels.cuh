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

}  // namespace ookd
