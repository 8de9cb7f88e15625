// synth_kernel.cuh -- device-side synthetic SC16Q11 capture (benchmark / test input only).
// Integer-only recipe shared byte for byte with oracle/ookd_oracle.c:ookd_oracle_synth:
//   envelope(n) = parity of #{toggles <= n};  sample = clip(env * (i_on, q_on) + noise, -2048, 2047)
//   noise       = round_half_up(scale * (sum of four (or twelve) 16-bit uniforms - mean) / 2^24)
// with the uniforms drawn from a counter-based 64-bit mixer keyed by (seed, 2n) / (seed, 2n+1).
#pragma once

#include "ookd_common.cuh"

namespace ookd {

__host__ __device__ __forceinline__ u64 synth_mix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// terms = 4: one draw (sum of four 16-bit uniforms, bounded at +-3.46 sigma); terms = 12: three draws (Irwin-Hall of
// twelve uniforms: Gaussian to within a few per cent out to 4 sigma, tails to +-6 sigma).
__device__ __forceinline__ int32_t synth_noise(u64 seed_mixed, u64 ctr, int32_t scale, uint32_t terms)
{
    const u64 base = seed_mixed ^ (ctr * 0xD1342543DE82EF95ull);
    u64 r = synth_mix64(base);
    i64 s = (i64) ((r & 0xFFFF) + ((r >> 16) & 0xFFFF) + ((r >> 32) & 0xFFFF) + ((r >> 48) & 0xFFFF)) - 131070;
    if (terms == 12) {
#pragma unroll
        for (int d = 1; d <= 2; d++) {
            r = synth_mix64(base ^ ((u64) d * 0xA24BAED4963EE407ull));
            s += (i64) ((r & 0xFFFF) + ((r >> 16) & 0xFFFF) + ((r >> 32) & 0xFFFF) + ((r >> 48) & 0xFFFF)) - 131070;
        }
    }
    const i64 v = s * (i64) scale + (1 << 23);
    return (int32_t) (v >> 24);
}

__device__ __forceinline__ int32_t synth_clip(int32_t v)
{
    return max(-2048, min(2047, v));
}

constexpr int SYNTH_SPT = 8;    // samples per thread

__global__ void __launch_bounds__(256) synth_kernel(uint32_t *dst, u64 first_sample, u64 n_samples,
                                                    const u64 *toggles, u64 n_toggles,
                                                    int32_t i_on, int32_t q_on, int32_t scale, u64 seed_mixed, uint32_t terms)
{
    const u64 j0 = ((u64) blockIdx.x * blockDim.x + threadIdx.x) * SYNTH_SPT;
    if (j0 >= n_samples) return;
    const u64 n0 = first_sample + j0;
    u64 lo = 0, hi = n_toggles;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (toggles[mid] <= n0) lo = mid + 1; else hi = mid;
    }
    u64 k = lo;
    u64 next = (k < n_toggles) ? toggles[k] : ~0ull;
#pragma unroll
    for (int q = 0; q < SYNTH_SPT; q++) {
        const u64 j = j0 + q;
        if (j >= n_samples) break;
        const u64 n = first_sample + j;
        while (next <= n) {
            k++;
            next = (k < n_toggles) ? toggles[k] : ~0ull;
        }
        const bool on = (k & 1) != 0;
        int32_t vi = on ? i_on : 0, vq = on ? q_on : 0;
        if (scale != 0) {
            vi += synth_noise(seed_mixed, 2 * n, scale, terms);
            vq += synth_noise(seed_mixed, 2 * n + 1, scale, terms);
        }
        dst[j] = ((uint32_t) (uint16_t) (int16_t) synth_clip(vi)) | (((uint32_t) (uint16_t) (int16_t) synth_clip(vq)) << 16);
    }
}

}  // namespace ookd
