// edge_kernels.cuh -- threshold decisions -> ordered edge list (run-length form).
// Replaces the prev/curr comparison the reference makes per sample in record_dig
// (src/ookiedokie.c:146-169) and in the pulse_start / pulse_end trigger tests
// (src/state_machine.c:441-457): an edge is an output index i >= 1 with bit[i] != bit[i-1].
//
// Three launches over the packed decisions (1/32 of the input bytes for decimation 1):
//   edge_count  : per-CTA popcount of transition masks
//   scan_u32    : single-CTA exclusive scan of the per-CTA counts (+ total)
//   edge_write  : per-CTA recomputation, intra-CTA warp-shuffle scan, ordered scatter
#pragma once

#include "ookd_common.cuh"

namespace ookd {

struct EdgeArgs {
    const u64 *words;        // packed decisions, bit b of the array <-> global output bit_base + b
    i64  bit_base;
    i64  start_bit;          // first array bit that may be an edge (its predecessor is start_bit-1;
                             // 0 => bit 0 has no predecessor and is never an edge)
    i64  n_bits;             // valid bits in the array
    uint32_t *block_counts;  // [gridDim.x] (count) / exclusive offsets (write)
    u64 *edges;              // out (write)
};

constexpr int EDGE_NT = 256;
constexpr int EDGE_WPT = 4;                         // words per thread
constexpr int EDGE_WPB = EDGE_NT * EDGE_WPT;        // words per CTA

__device__ __forceinline__ u64 transition_mask(const EdgeArgs &a, i64 wi)
{
    const i64 lo = wi * 64;
    if (lo >= a.n_bits) {
        return 0;
    }
    const u64 w = a.words[wi];
    u64 prev;
    if (wi == 0) {
        prev = w & 1;                               // bit 0 compared with itself: no edge
    } else {
        prev = a.words[wi - 1] >> 63;
    }
    u64 t = w ^ ((w << 1) | prev);
    if (lo < a.start_bit) {                         // drop array bits before start_bit
        const i64 sh = a.start_bit - lo;
        t = (sh >= 64) ? 0 : (t >> sh) << sh;
    }
    if (lo + 64 > a.n_bits) {                       // drop bits past the end
        const int keep = (int) (a.n_bits - lo);
        t &= (keep >= 64) ? ~0ull : ((1ull << keep) - 1);
    }
    return t;
}

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if (w < warp) base += s_warp[w];
        sum += s_warp[w];
    }
    __syncthreads();
    if (total) *total = sum;
    return base + inc - v;
}

__global__ void __launch_bounds__(EDGE_NT) edge_count_kernel(const EdgeArgs a)
{
    const i64 w0 = ((i64) blockIdx.x * EDGE_NT + threadIdx.x) * EDGE_WPT;
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < EDGE_WPT; q++) {
        c += __popcll(transition_mask(a, w0 + q));
    }
    uint32_t total;
    (void) block_exclusive_scan_256(c, &total);
    if (threadIdx.x == 0) {
        a.block_counts[blockIdx.x] = total;
    }
}

__global__ void __launch_bounds__(EDGE_NT) edge_write_kernel(const EdgeArgs a)
{
    const i64 w0 = ((i64) blockIdx.x * EDGE_NT + threadIdx.x) * EDGE_WPT;
    u64 t[EDGE_WPT];
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < EDGE_WPT; q++) {
        t[q] = transition_mask(a, w0 + q);
        c += __popcll(t[q]);
    }
    u64 dst = (u64) a.block_counts[blockIdx.x] + block_exclusive_scan_256(c, nullptr);
#pragma unroll
    for (int q = 0; q < EDGE_WPT; q++) {
        u64 m = t[q];
        while (m) {
            const int b = __ffsll((long long) m) - 1;
            m &= m - 1;
            a.edges[dst++] = (u64) (a.bit_base + (w0 + q) * 64 + b);
        }
    }
}

// ---------------------------------------------------------------------------------------
// One-pass form: count, scan and ordered write in a single sweep over the decisions (chained scan with
// decoupled look-back).  CTAs take their position from a ticket counter, so a CTA only ever waits for
// CTAs that started before it.  status[i] = flag << 62 | value, flag 1 = CTA i's own count,
// flag 2 = inclusive count of CTAs 0..i.  Edges beyond `cap` are counted but not written (the host then
// grows the list and repeats the pass).
// ---------------------------------------------------------------------------------------
constexpr int EDGE1_WPT = 8;                         // words per thread (64 B)
constexpr int EDGE1_WPB = EDGE_NT * EDGE1_WPT;       // words per CTA (16 KiB of decisions)

struct Edge1Args {
    EdgeArgs e;
    u64 *status;             // [gridDim.x], zeroed before the launch
    uint32_t *ticket;        // zeroed before the launch
    u64 *total;              // out: number of edges
    u64 cap;                 // capacity of e.edges
};

__global__ void __launch_bounds__(EDGE_NT) edge_onepass_kernel(const Edge1Args x)
{
    const EdgeArgs &a = x.e;
    __shared__ uint32_t s_bid;
    __shared__ u64 s_prefix;
    if (threadIdx.x == 0) s_bid = atomicAdd(x.ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const i64 w0 = ((i64) bid * EDGE_NT + threadIdx.x) * EDGE1_WPT;
    const i64 n_words = (a.n_bits + 63) >> 6;

    u64 w[EDGE1_WPT], t[EDGE1_WPT];
    if (w0 + EDGE1_WPT <= n_words) {
        const ulonglong2 *src = (const ulonglong2 *) (a.words + w0);      // bits buffer is 16-byte aligned, w0 % 8 == 0
#pragma unroll
        for (int q = 0; q < EDGE1_WPT / 2; q++) {
            const ulonglong2 v = src[q];
            w[2 * q] = v.x;
            w[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < EDGE1_WPT; q++) w[q] = (w0 + q < n_words) ? a.words[w0 + q] : 0ull;
    }
    u64 prev = (w0 == 0) ? (w[0] & 1) : ((w0 <= n_words) ? (a.words[w0 - 1] >> 63) : 0ull);
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < EDGE1_WPT; q++) {
        const i64 lo = (w0 + q) * 64;
        u64 m = w[q] ^ ((w[q] << 1) | prev);
        prev = w[q] >> 63;
        if (lo < a.start_bit) {
            const i64 sh = a.start_bit - lo;
            m = (sh >= 64) ? 0 : (m >> sh) << sh;
        }
        if (lo + 64 > a.n_bits) {
            const i64 keep = a.n_bits - lo;
            m = (keep <= 0) ? 0 : (m & ((keep >= 64) ? ~0ull : ((1ull << keep) - 1)));
        }
        t[q] = m;
        c += __popcll(m);
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan_256(c, &total);

    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        volatile u64 *st = x.status;
        if (lane == 0) {
            st[bid] = ((bid == 0 ? 2ull : 1ull) << 62) | (u64) total;
            __threadfence();
        }
        u64 run = 0;
        if (bid > 0) {
            i64 top = (i64) bid - 1;                   // nearest predecessor not yet accounted for
            for (;;) {
                const i64 i = top - lane;
                u64 v;
                do {
                    v = (i >= 0) ? st[i] : (2ull << 62);            // before CTA 0: inclusive prefix 0
                } while (__any_sync(0xFFFFFFFFu, (v >> 62) == 0));
                const uint32_t incl = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                const uint32_t upto = incl ? (uint32_t) (__ffs(incl) - 1) : 31u;   // lanes 0..upto contribute
                u64 part = (lane <= upto) ? (v & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
                run += part;
                if (incl) break;
                top -= 32;
            }
            if (lane == 0) {
                st[bid] = (2ull << 62) | (run + total);
                __threadfence();
            }
        }
        if (lane == 0) {
            s_prefix = run;
            if (bid == gridDim.x - 1) *x.total = run + total;
        }
    }
    __syncthreads();
    u64 dst = s_prefix + excl;
#pragma unroll
    for (int q = 0; q < EDGE1_WPT; q++) {
        u64 m = t[q];
        while (m) {
            const int b = __ffsll((long long) m) - 1;
            m &= m - 1;
            if (dst < x.cap) a.edges[dst] = (u64) (a.bit_base + (w0 + q) * 64 + b);
            dst++;
        }
    }
}

// Single-CTA exclusive scan of n uint32 values in place; total to *total (64-bit).
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t *v, uint32_t n, u64 *total)
{
    __shared__ u64 s_part[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per;
    const uint32_t hi = min(n, lo + per);
    u64 sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += v[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (int d = 1; d < 1024; d <<= 1) {
        u64 add = (threadIdx.x >= d) ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    u64 run = s_part[threadIdx.x] - sum;            // exclusive prefix of this thread's segment
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t x = v[i];
        v[i] = (uint32_t) run;                      // per-shard edge counts stay below 2^32
        run += x;
    }
    if (threadIdx.x == 1023) *total = s_part[1023];
}

}  // namespace ookd
