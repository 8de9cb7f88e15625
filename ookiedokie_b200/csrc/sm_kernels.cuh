// sm_kernels.cuh -- the device state machine run over the edge list instead of per sample.
// Replaces sm_process / process / handle_rx_triggers (src/state_machine.c:421-556) and the
// per-buffer driver device_process (src/device.c:634-658), including its rule that an ERROR
// abandons the rest of the current samples_per_buffer buffer.
//
// Event-driven form.  Between two edges the decision bit is constant, so pulse_start /
// pulse_end cannot fire and the machine can only move through its always / timeout /
// msg_complete triggers, whose firing sample follows from the integer windows of
// sm_compile.c in O(#triggers).  Only edge samples, samples evaluated in RESET (evaluated
// twice, :526-538) and the first sample after a dropped buffer tail (prev_bit is stale there,
// :549-552) take the single-sample path, which mirrors handle_rx_triggers one to one.
//
// Parallel form.  The output stream is cut into chunks of whole buffers, one thread each.
// Round 0 runs every chunk from a guessed entry state (RESET, k=0, true previous bit); every
// later round re-runs exactly the chunks whose entry (= predecessor's exit of the previous
// round) differs from the entry they last ran with.  A round that re-runs nothing is a fixed
// point, and the fixed point is the sequential result because chunk 0's entry is given and
// each chunk is a deterministic function of its entry.
#pragma once

#include "ookd_common.cuh"

namespace ookd {

struct SmArgs {
    const SmTable *tab;
    const u64 *edges;
    u64  n_edges;
    uint32_t base_bit;        // decision at global output out_lo-1 (or of output 0 when out_lo == 0)
    i64  out_lo, out_hi;      // outputs covered by this shard
    u64  spb;                 // input samples per buffer
    u64  dec;                 // total decimation
    u64  first_buffer;        // buffer index whose first output is out_lo
    uint32_t chunk_buffers;
    uint32_t n_chunks;
    uint32_t round;           // 0 => speculative first round
    uint32_t counter_idx;     // slot of n_ran this launch reports into
    SmCarry entry0;
    const SmCarry *exit_prev;
    SmCarry *exit_cur;
    SmCarry *ran_with;
    SmMsg   *slots;           // [n_chunks * slot_cap]
    uint32_t slot_cap;
    uint32_t *slot_count;     // [n_chunks]
    uint32_t *n_ran;          // [rounds]
    uint32_t *overflow;
};

__device__ __forceinline__ i64 first_output_of_buffer(const SmArgs &a, u64 b)
{
    return (i64) ((b * a.spb) / a.dec);          // outputs produced by the first b buffers
}

__device__ __forceinline__ u64 buffer_of_output(const SmArgs &a, i64 m)
{
    return (((u64) m + 1) * a.dec - 1) / a.spb;   // buffer holding the input that emits m
}

__device__ __forceinline__ bool carry_equal(const SmCarry &x, const SmCarry &y)
{
    return x.state == y.state && x.k == y.k && x.num_bits == y.num_bits && x.prev == y.prev &&
           x.data[0] == y.data[0] && x.data[1] == y.data[1] && x.data[2] == y.data[2] &&
           x.data[3] == y.data[3];
}

// handle_actions, src/state_machine.c:388-419 (+ append_data_bit :365-385)
__device__ __forceinline__ int sm_apply(const SmTable &T, SmCarry &s, const ookd_sm_trigger_k &t)
{
    int result = 0;
    if (t.action == OOKD_ACT_APPEND_0 || t.action == OOKD_ACT_APPEND_1) {
        if (s.num_bits <= T.max_bits && s.num_bits < 256) {
            const u64 m = 1ull << (s.num_bits & 63);
            const uint32_t w = s.num_bits >> 6;
            if (t.action == OOKD_ACT_APPEND_1) {
                s.data[w] |= m;
            } else {
                s.data[w] &= ~m;
            }
        }
        s.num_bits++;
    } else if (t.action == OOKD_ACT_OUTPUT_DATA) {
        result = 1;
    }
    s.state = t.next_state;
    return result;
}

// One trigger evaluation on sample value b: handle_rx_triggers, src/state_machine.c:421-519.
__device__ __forceinline__ int sm_eval(const SmTable &T, SmCarry &s, uint32_t b)
{
    const ookd_sm_state_k &st = T.states[s.state];
    int fired = -1;
    bool check = false;
    for (uint32_t i = 0; i < st.num_triggers && fired < 0; i++) {
        const ookd_sm_trigger_k &t = T.triggers[st.first_trigger + i];
        if (s.k < t.kmin || s.k > t.kmax) {
            continue;
        }
        switch (t.cond) {
            case OOKD_COND_ALWAYS:
                fired = (int) i;
                break;
            case OOKD_COND_PULSE_START:
                if (!s.prev && b) { fired = (int) i; check = true; }
                break;
            case OOKD_COND_PULSE_END:
                if (s.prev && !b) { fired = (int) i; check = true; }
                break;
            case OOKD_COND_TIMEOUT:
                if (st.ktimeout != OOKD_K_INF && s.k >= st.ktimeout) { fired = (int) i; }
                break;
            case OOKD_COND_MSG_COMPLETE:
                if (s.num_bits >= T.max_bits) { fired = (int) i; }
                break;
            default:
                break;
        }
    }
    if (fired < 0) {
        s.k = (s.k + 1 < T.k_sat) ? s.k + 1 : T.k_sat;
        return 0;
    }
    int result;
    if (!check || (s.k >= st.dmin && s.k <= st.dmax)) {
        result = sm_apply(T, s, T.triggers[st.first_trigger + fired]);
    } else {
        result = -1;
        s.state = 0;
    }
    s.k = 0;
    return result;
}

// process + the prev_bit update of sm_process, src/state_machine.c:521-556.
__device__ __forceinline__ int sm_step(const SmTable &T, SmCarry &s, uint32_t b)
{
    int r = 0;
    if (s.state == 0) {
        s.num_bits = 0;
        // memset(data, 0, (max_bits + 7) / 8)
        const uint32_t nbytes = (T.max_bits + 7) >> 3;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t lo = 8u * w;
            if (nbytes >= lo + 8) {
                s.data[w] = 0;
            } else if (nbytes > lo) {
                s.data[w] &= ~((1ull << (8 * (nbytes - lo))) - 1);
            }
        }
        r = sm_eval(T, s, b);
    }
    if (r == 0) {
        r = sm_eval(T, s, b);
    }
    s.prev = b;
    return r;
}

// First count k' >= s.k at which a trigger that does not need an edge fires (bit constant,
// prev == bit).  Returns the trigger's index within the state or -1.
__device__ __forceinline__ int sm_next_quiet_fire(const SmTable &T, const SmCarry &s, uint32_t *k_fire)
{
    const ookd_sm_state_k &st = T.states[s.state];
    int best = -1;
    uint32_t best_k = OOKD_K_INF;
    for (uint32_t i = 0; i < st.num_triggers; i++) {
        const ookd_sm_trigger_k &t = T.triggers[st.first_trigger + i];
        uint32_t lo = t.kmin;
        if (t.cond == OOKD_COND_TIMEOUT) {
            if (st.ktimeout == OOKD_K_INF) continue;
            lo = max(lo, st.ktimeout);
        } else if (t.cond == OOKD_COND_MSG_COMPLETE) {
            if (s.num_bits < T.max_bits) continue;
        } else if (t.cond != OOKD_COND_ALWAYS) {
            continue;
        }
        const uint32_t kk = max(lo, s.k);
        if (kk > t.kmax) continue;             // window already closed (finite kmax < k_sat)
        if (kk < best_k) {                      // strict: earlier list position wins ties
            best_k = kk;
            best = (int) i;
        }
    }
    *k_fire = best_k;
    return best;
}

__device__ __forceinline__ u64 edge_lower_bound(const u64 *edges, u64 n, u64 pos)
{
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (edges[mid] < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(32) sm_round_kernel(const SmArgs a)
{
    __shared__ SmTable T;
    {
        const uint32_t *src = (const uint32_t *) a.tab;
        uint32_t *dst = (uint32_t *) &T;
        for (uint32_t i = threadIdx.x; i < sizeof(SmTable) / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_chunks) return;

    const i64 start = (c == 0) ? a.out_lo
                               : first_output_of_buffer(a, a.first_buffer + (u64) c * a.chunk_buffers);
    i64 end = first_output_of_buffer(a, a.first_buffer + (u64) (c + 1) * a.chunk_buffers);
    if (end > a.out_hi || c == a.n_chunks - 1) end = a.out_hi;

    // edges[e] is the first edge at or after `start`; e edges precede it
    u64 e = edge_lower_bound(a.edges, a.n_edges, (u64) start);
    uint32_t tb = a.base_bit ^ (uint32_t) (e & 1);           // true decision at start-1

    SmCarry s;
    if (c == 0) {
        s = a.entry0;
    } else if (a.round == 0) {
        s.state = 0; s.k = 0; s.num_bits = 0; s.prev = tb;
        s.data[0] = s.data[1] = s.data[2] = s.data[3] = 0;
    } else {
        s = a.exit_prev[c - 1];
    }
    if (a.round > 0 && carry_equal(s, a.ran_with[c])) {
        a.exit_cur[c] = a.exit_prev[c];
        return;
    }
    a.ran_with[c] = s;
    atomicAdd(&a.n_ran[a.counter_idx], 1u);

    uint32_t n_msgs = 0;
    SmMsg *slots = a.slots + (u64) c * a.slot_cap;

    const u64 INF = ~0ull;
    u64 next_edge = (e < a.n_edges) ? a.edges[e] : INF;
    u64 after_edge = (e + 1 < a.n_edges) ? a.edges[e + 1] : INF;   // one-ahead prefetch
    i64 pos = start;

    while (pos < end) {
        const bool at_edge = (next_edge == (u64) pos);
        if (s.state == 0 || s.prev != tb || at_edge) {
            const uint32_t b = at_edge ? (tb ^ 1u) : tb;
            const int r = sm_step(T, s, b);
            if (at_edge) {
                tb ^= 1u;
                e++;
                next_edge = after_edge;
                after_edge = (e + 1 < a.n_edges) ? a.edges[e + 1] : INF;
            }
            pos++;
            if (r > 0) {
                if (n_msgs < a.slot_cap) {
                    SmMsg m;
                    m.out_sample = (u64) (pos - 1);
                    m.num_bits = s.num_bits;
                    m.pad = 0;
                    m.data[0] = s.data[0]; m.data[1] = s.data[1]; m.data[2] = s.data[2]; m.data[3] = s.data[3];
                    slots[n_msgs] = m;
                } else {
                    atomicExch(a.overflow, 1u);
                }
                n_msgs++;
            } else if (r < 0) {
                // device_process gives up on this buffer: resume at the next buffer's first output
                i64 nb = first_output_of_buffer(a, buffer_of_output(a, pos - 1) + 1);
                if (nb > end) nb = end;
                if (nb > pos) {
                    if (next_edge < (u64) nb) {
                        const u64 e2 = e + edge_lower_bound(a.edges + e, a.n_edges - e, (u64) nb);
                        tb ^= (uint32_t) ((e2 - e) & 1);
                        e = e2;
                        next_edge = (e < a.n_edges) ? a.edges[e] : INF;
                        after_edge = (e + 1 < a.n_edges) ? a.edges[e + 1] : INF;
                    }
                    pos = nb;
                }
            }
            continue;
        }

        // quiet stretch: samples [pos, limit) carry bit tb == s.prev, machine not in RESET
        const u64 limit = (next_edge < (u64) end) ? next_edge : (u64) end;
        const u64 gap = limit - (u64) pos;
        uint32_t k_fire;
        const int tf = sm_next_quiet_fire(T, s, &k_fire);
        if (tf >= 0 && (u64) (k_fire - s.k) < gap) {
            pos += (i64) (k_fire - s.k);
            const ookd_sm_state_k &st = T.states[s.state];
            const int r = sm_apply(T, s, T.triggers[st.first_trigger + tf]);
            s.k = 0;
            pos++;
            if (r > 0) {
                if (n_msgs < a.slot_cap) {
                    SmMsg m;
                    m.out_sample = (u64) (pos - 1);
                    m.num_bits = s.num_bits;
                    m.pad = 0;
                    m.data[0] = s.data[0]; m.data[1] = s.data[1]; m.data[2] = s.data[2]; m.data[3] = s.data[3];
                    slots[n_msgs] = m;
                } else {
                    atomicExch(a.overflow, 1u);
                }
                n_msgs++;
            }
        } else {
            const u64 kk = (u64) s.k + gap;
            s.k = (kk < (u64) T.k_sat) ? (uint32_t) kk : T.k_sat;
            pos = (i64) limit;
        }
    }

    a.exit_cur[c] = s;
    a.slot_count[c] = (n_msgs < a.slot_cap) ? n_msgs : a.slot_cap;
}

// Ordered compaction of the per-chunk message slots (offsets = exclusive scan of slot_count).
__global__ void sm_gather_kernel(const SmMsg *slots, uint32_t slot_cap, const uint32_t *counts,
                                 const uint32_t *offsets, uint32_t n_chunks, SmMsg *out)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint32_t n = counts[c];
    for (uint32_t i = 0; i < n; i++) {
        out[offsets[c] + i] = slots[(u64) c * slot_cap + i];
    }
}

}  // namespace ookd
