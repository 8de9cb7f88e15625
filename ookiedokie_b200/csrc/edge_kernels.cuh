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
// Production form: tile-local extraction + flatten.
//   edge_local   : persistent CTAs stream the decisions once (coalesced 16-byte loads, next tile prefetched
//                  into registers), keep the transition masks in registers and write each 32 KiB tile's edges,
//                  in order, into that tile's own fixed-size region.  No inter-CTA dependency.
//   scan_u32     : exclusive scan of the per-tile counts (+ total)
//   edge_flatten : copies the regions to their final positions; also leaves the header the state-machine
//                  kernels / the host's single read-back need.
// A tile with more edges than its region holds (1 per 128 decisions) sets the overflow flag; the host then
// uses the count / scan / write kernels above.
// ---------------------------------------------------------------------------------------
#ifndef OOKD_EDGE1_ROWS
#define OOKD_EDGE1_ROWS 4
#endif
constexpr int EDGE1_ROWS = OOKD_EDGE1_ROWS;          // rows per warp; a row = 32 lanes x 2 words (512 B, coalesced)
constexpr int EDGE1_WPW = EDGE1_ROWS * 64;           // words per warp (4 KiB of decisions)
constexpr int EDGE1_WPB = (EDGE_NT / 32) * EDGE1_WPW;   // words per tile (32 KiB)
constexpr int EDGE1_CAP = EDGE1_WPB * 64 / 128;      // edges a tile's region holds (2048)

struct Edge1Args {
    EdgeArgs e;              // e.block_counts: [n_tiles] counts (local) / exclusive offsets (flatten); e.edges: final list
    u64 *tmp;                // [n_tiles * EDGE1_CAP] tile regions
    uint32_t n_tiles;
    uint32_t *overflow;      // set to 1 if a tile overflowed its region
    const u64 *total;        // (flatten) total edge count from the scan
    u64 cap;                 // capacity of e.edges
    u64 *hdr;                // (flatten) [0] = min(total, cap), [1] = decision in front of the shard (low 32 bits),
                             // [2] = first word of decisions, [3] = word holding the shard's first decision
    i64 report_word;         // index of that word
};

__global__ void __launch_bounds__(EDGE_NT) edge_local_kernel(const Edge1Args x)
{
    const EdgeArgs &a = x.e;
    __shared__ uint32_t s_warp_cnt[2][EDGE_NT / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 n_words = (a.n_bits + 63) >> 6;

    auto load_tile = [&](uint32_t tile, u64 (&w)[EDGE1_ROWS][2], u64 &carry) {
        const i64 wbase = ((i64) tile * (EDGE_NT / 32) + warp) * EDGE1_WPW;
        const bool full = (wbase + EDGE1_WPW) <= n_words;
#pragma unroll
        for (int r = 0; r < EDGE1_ROWS; r++) {
            const i64 wi = wbase + r * 64 + 2 * lane;
            if (full) {
                const ulonglong2 v = *(const ulonglong2 *) (a.words + wi);      // 16-byte aligned: wi is even
                w[r][0] = v.x; w[r][1] = v.y;
            } else {
                w[r][0] = (wi < n_words) ? a.words[wi] : 0ull;
                w[r][1] = (wi + 1 < n_words) ? a.words[wi + 1] : 0ull;
            }
        }
        carry = 0;                                        // top bit of the word in front of the warp's first row
        if (wbase < n_words) carry = (wbase == 0) ? (a.words[0] & 1) : (a.words[wbase - 1] >> 63);
    };

    u64 w[EDGE1_ROWS][2], wn[EDGE1_ROWS][2], carry, carry_n = 0;
    uint32_t tile = blockIdx.x;
    if (tile >= x.n_tiles) return;
    load_tile(tile, w, carry);
    for (uint32_t it = 0; tile < x.n_tiles; tile += gridDim.x, it++) {
        const uint32_t nxt = tile + gridDim.x;
        if (nxt < x.n_tiles) load_tile(nxt, wn, carry_n);
        const i64 wbase = ((i64) tile * (EDGE_NT / 32) + warp) * EDGE1_WPW;
        const bool interior = wbase > 0 && (wbase * 64 >= a.start_bit) && ((wbase + EDGE1_WPW) * 64 <= a.n_bits);
        uint32_t c = 0;
#pragma unroll
        for (int r = 0; r < EDGE1_ROWS; r++) {
            const u64 w0 = w[r][0], w1 = w[r][1];
            u64 prev = __shfl_up_sync(0xFFFFFFFFu, w1 >> 63, 1);
            if (lane == 0) prev = carry;
            carry = __shfl_sync(0xFFFFFFFFu, w1 >> 63, 31);
            u64 t0 = w0 ^ ((w0 << 1) | prev);
            u64 t1 = w1 ^ ((w1 << 1) | (w0 >> 63));
            if (!interior) {
                const i64 wi = wbase + r * 64 + 2 * lane;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    u64 &t = h ? t1 : t0;
                    const i64 lo = (wi + h) * 64;
                    if (lo < a.start_bit) {
                        const i64 sh = a.start_bit - lo;
                        t = (sh >= 64) ? 0 : (t >> sh) << sh;
                    }
                    if (lo + 64 > a.n_bits) {
                        const i64 keep = a.n_bits - lo;
                        t = (keep <= 0) ? 0 : (t & ((keep >= 64) ? ~0ull : ((1ull << keep) - 1)));
                    }
                }
            }
            w[r][0] = t0;                                 // the masks replace the words
            w[r][1] = t1;
            c += __popcll(t0) + __popcll(t1);
        }
        const uint32_t warp_cnt = __reduce_add_sync(0xFFFFFFFFu, c);
        uint32_t *cnt = s_warp_cnt[it & 1];               // double buffered: one barrier per tile
        if (lane == 0) cnt[warp] = warp_cnt;
        __syncthreads();
        uint32_t total = 0, warp_off = 0;
#pragma unroll
        for (int q = 0; q < EDGE_NT / 32; q++) {
            if (q < (int) warp) warp_off += cnt[q];
            total += cnt[q];
        }
        if (threadIdx.x == 0) {
            a.block_counts[tile] = total;
            if (total > EDGE1_CAP) atomicExch(x.overflow, 1u);
        }
        if (warp_cnt != 0) {
            // ordered write: rows in order, lanes in order inside a row; rows without a transition cost one ballot
            u64 *out = x.tmp + (u64) tile * EDGE1_CAP;
            uint32_t dst = warp_off;
#pragma unroll
            for (int r = 0; r < EDGE1_ROWS; r++) {
                const uint32_t n = __popcll(w[r][0]) + __popcll(w[r][1]);
                const uint32_t have = __ballot_sync(0xFFFFFFFFu, n != 0);
                if (have == 0) continue;
                uint32_t inc = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= (uint32_t) d) inc += up;
                }
                uint32_t my = dst + (inc - n);
                dst += __shfl_sync(0xFFFFFFFFu, inc, 31);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    u64 t = w[r][h];
                    const i64 lo = (wbase + r * 64 + 2 * lane + h) * 64;
                    while (t) {
                        const int b = __ffsll((long long) t) - 1;
                        t &= t - 1;
                        if (my < EDGE1_CAP) out[my] = (u64) (a.bit_base + lo + b);
                        my++;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < EDGE1_ROWS; r++) {
            w[r][0] = wn[r][0];
            w[r][1] = wn[r][1];
        }
        carry = carry_n;
    }
}

// block b copies tile b's region to its final position (block_counts now holds the exclusive offsets)
__global__ void __launch_bounds__(128) edge_flatten_kernel(const Edge1Args x)
{
    const EdgeArgs &a = x.e;
    const uint32_t tile = blockIdx.x;
    const u64 all = *x.total;
    const u64 off = a.block_counts[tile];
    const u64 nxt = (tile + 1 < x.n_tiles) ? (u64) a.block_counts[tile + 1] : all;
    uint32_t n = (uint32_t) (nxt - off);
    if (n > EDGE1_CAP) n = EDGE1_CAP;
    const u64 *src = x.tmp + (u64) tile * EDGE1_CAP;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (off + i < x.cap) a.edges[off + i] = src[i];
    }
    if (tile == 0 && threadIdx.x == 0 && x.hdr) {
        const u64 word0 = a.words[0];
        x.hdr[0] = all < x.cap ? all : x.cap;
        x.hdr[1] = (word0 >> (a.start_bit ? a.start_bit - 1 : 0)) & 1;
        x.hdr[2] = word0;
        x.hdr[3] = a.words[x.report_word];
    }
}

// Single-CTA exclusive scan of n uint32 values in place; total to *total (64-bit).

constexpr int SCAN_NT = 1024;

__global__ void __launch_bounds__(SCAN_NT) scan_u32_kernel(uint32_t *v, uint32_t n, u64 *total)
{
    __shared__ u64 s_warp[SCAN_NT / 32];
    const uint32_t per = (n + SCAN_NT - 1) / SCAN_NT;
    const uint32_t lo = threadIdx.x * per;
    const uint32_t hi = min(n, lo + per);
    u64 sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += v[i];
    // inclusive scan of the per-thread sums: warp shuffles, then the 8 warp totals
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t) d) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    u64 base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < SCAN_NT / 32; w++) {
        if (w < (int) warp) base += s_warp[w];
        all += s_warp[w];
    }
    u64 run = base + inc - sum;                     // exclusive prefix of this thread's segment
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t x = v[i];
        v[i] = (uint32_t) run;                      // per-shard edge counts stay below 2^32
        run += x;
    }
    if (threadIdx.x == 0) *total = all;
}

}  // namespace ookd
