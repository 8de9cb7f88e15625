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
// Parallel form.  The output stream is cut into chunks of whole buffers, and each chunk's boundary is then MOVED to
// the chunk's first "anchor" (first rising edge from which a fresh machine appends a bit, i.e. a plausible message
// start; sm_anchor_kernel, 32 lanes probe 32 candidate edges at once).  The machine's state at a boundary depends
// on everything before it, so each chunk keeps a small TABLE of (entry state -> exit state, messages) pairs, one
// warp per pair (lane j owns trigger j: warp_sm_*):
//   round 0   one speculative seed per chunk: the idle machine (ookd_sm_idle_carry) at the anchor, as a real
//             entry -- the truth is normally idle when a message starts -- or RESET at the chunk's first sample
//             when it has no anchor; chunk 0 runs from the true entry;
//   round r   every exit in chunk c-1's table that is not yet an entry of chunk c's table is run, and the warp
//             runs on into the following chunks while its exit is still unknown there (cascades);
//   link/walk exits are matched to the next chunk's entries and the chain is walked from chunk 0's true entry
//             (links composed over segments in parallel).  If the walk reaches the last chunk the decode is
//             resolved: the chosen pairs ARE the sequential run, because each pair is the deterministic
//             function of its entry.  Otherwise another round adds the missing entries.
// A Jacobi relaxation over single exits (sm_round_kernel, fixed boundaries: re-run exactly the chunks whose entry
// changed until nothing changes) is kept as the always-terminating fallback.
//
// Counts k saturate per state at ksat = 1 + (largest finite bound any predicate of that state
// compares k with): beyond it every predicate of the state is constant, so the saturated machine is
// indistinguishable from the reference's and equal situations compare equal.
#pragma once

#include "ookd_common.cuh"

namespace ookd {

// Edge count and the decision in front of the shard as the edge pass left them in device memory: lets
// the state-machine kernels be enqueued behind the edge pass without a host round trip.
struct SmDevHdr {
    u64 n_edges;
    uint32_t base_bit, pad;
};

struct SmArgs {
    const SmDevHdr *hdr;      // non-null: n_edges / base_bit below are placeholders, read them from here
    const SmTable *tab;
    const u64 *edges;
    u64  n_edges;
    uint32_t base_bit;        // decision at global output out_lo-1 (or of output 0 when out_lo == 0)
    i64  out_lo, out_hi;      // outputs covered by this shard
    u64  spb;                 // input samples per buffer
    u64  dec;                 // total decimation
    uint32_t opb;             // outputs per buffer when spb % dec == 0, else 0
    u64  first_buffer;        // buffer index whose first output is out_lo
    uint32_t chunk_buffers;
    uint32_t n_chunks;
    uint32_t round;           // 0 => speculative first round
    uint32_t counter_idx;     // slot of n_ran this launch reports into
    SmCarry entry0;
    const SmCarry *entry_ptr; // non-null: the shard's true entry lives in device memory (exit of the previous window)
    const SmCarry *exit_prev;
    SmCarry *exit_cur;
    SmCarry *ran_with;
    SmMsg   *slots;           // [n_chunks * slot_cap]
    uint32_t slot_cap;
    uint32_t *slot_count;     // [n_chunks]
    uint32_t *n_ran;          // [rounds]
    uint32_t *overflow;
    // table speculation
    uint32_t tab_k;           // pairs per chunk
    SmCarry *tab_entry;       // [n_chunks * tab_k]
    SmCarry *tab_exit;        // [n_chunks * tab_k]
    uint32_t *tab_nmsg;       // [n_chunks * tab_k]
    const uint32_t *cnt_in;   // [n_chunks] pairs present before this round
    uint32_t *cnt_out;        // [n_chunks] pairs present after it (initialised to cnt_in)
    uint8_t  *link;           // [n_chunks * tab_k]
    uint8_t  *chosen;         // [n_chunks]
    uint32_t *walk_status;    // [0] = chunks resolved, [1] = 1 if complete
    uint32_t *msg_counts;     // [n_chunks] messages of the chosen pair
    SmCarry  *final_exit;     // exit of the last chunk's chosen pair (written by the walk)
    uint32_t *start_slot;     // slot of chunk first_chunk the walk starts from
    uint32_t first_chunk;     // chunk the walk starts at (0)
    uint32_t warm;            // chunk 0 is a warm-up chunk in front of the shard: it has no true entry
    SmCarry  *final_entry;    // warm: state at the shard's first output, as the chosen run of chunk 0 passed it
    i64 report_lo;            // warm: the shard's first output.  Chunk 0 starts one chunk of history earlier and runs THROUGH
                              // it up to chunk 1's anchor; every run of chunk 0 records its state there in mid_carry[slot]
    SmCarry  *mid_carry;      // [tab_k]
    unsigned long long *dbg;  // null, or [n_chunks * 4] per seed warp: cycles, steps, errors << 32 | singles, hops (OOKD_DEBUG)
    uint32_t entry_at_report; // resolve of a warm shard: entry0 is the state AT report_lo; the pair added to chunk 0 starts there
    u64 *chunk_e;             // [n_chunks] index of the first edge at or after each chunk's start (sm_anchor_kernel)
    // Boundaries moved to anchors (sm_anchor_kernel): chunk c covers [bound_pos[c], bound_pos[c+1]).
    i64 *bound_pos;           // [n_chunks]; null => the fixed buffer boundaries (Jacobi fallback)
    i64 *seed_pos;            // [n_chunks] where the round-0 seed run of chunk c starts (>= bound_pos[c])
    u64 *seed_e;              // [n_chunks] index of the first edge at or after seed_pos
    uint8_t *seed_kind;       // [n_chunks] OOKD_SEED_*
    SmCarry canon;            // state an idle machine is in (ookd_sm_idle_carry): the seed assumed at an anchor
};

#define OOKD_SEED_TRUE      0   /* the shard's true entry (chunk 0)                                             */
#define OOKD_SEED_CANON     1   /* idle machine at the chunk's anchor; the chunk's boundary IS the anchor        */
#define OOKD_SEED_IDLE      2   /* no transition in the chunk, carrier off: idle machine at its first sample (silence)   */
#define OOKD_SEED_RESET     3   /* no anchor (chunk 0 of a warm shard; carrier on throughout): RESET at the first sample */
#define OOKD_SEED_NONE      4   /* transitions but no anchor: the chunk continues a burst that began earlier.  A guess at
                                 * its first sample would land mid-message and drag its garbage through the following
                                 * chunks; the run that arrives from the chunk in front adds this chunk's pair instead    */

__device__ __forceinline__ i64 first_output_of_buffer(const SmArgs &a, u64 b)
{
    if (a.opb) return (i64) (b * a.opb);         // spb is a multiple of the decimation
    return (i64) ((b * a.spb) / a.dec);          // outputs produced by the first b buffers
}

__device__ __forceinline__ u64 buffer_of_output(const SmArgs &a, i64 m)
{
    return (((u64) m + 1) * a.dec - 1) / a.spb;   // buffer holding the input that emits m
}

// First output of the buffer after the one holding output m, given that m lies in [lo, lo + 2^32):
// 32-bit arithmetic on the offset from a buffer boundary `lo` when buffers hold a whole number of outputs.
__device__ __forceinline__ i64 next_buffer_start(const SmArgs &a, i64 m, i64 lo)
{
    if (a.opb) {
        const uint32_t rel = (uint32_t) (m - lo);
        return lo + (i64) ((rel / a.opb + 1) * (u64) a.opb);
    }
    return first_output_of_buffer(a, buffer_of_output(a, m) + 1);
}

// Bit 1 of SmCarry::prev: "device_process has given up on the buffer this position lies in" (ERROR before a span
// boundary that is not a buffer boundary); the next span skips to the buffer's end before it evaluates anything.
#define OOKD_CARRY_DROPPING 2u

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
            const u64 set = (t.action == OOKD_ACT_APPEND_1) ? m : 0ull;
            // no dynamic indexing: the carry must stay in registers (a local-memory carry costs an L1 round
            // trip per field access, dozens per step)
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                if (i == w) s.data[i] = (s.data[i] & ~m) | set;
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
        s.k = (s.k + 1 < st.ksat) ? s.k + 1 : st.ksat;
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


// Same, for a position expected to lie only a few edges ahead (the next buffer after an ERROR): gallop, then bisect.
__device__ __forceinline__ u64 edge_lower_bound_near(const u64 *edges, u64 n, u64 pos)
{
    u64 lo = 0, step = 1;
    while (lo + step <= n && edges[lo + step - 1] < pos) {
        lo += step;
        step <<= 1;
    }
    u64 hi = (lo + step - 1 < n) ? lo + step - 1 : n;      // edges[hi] >= pos (or hi == n); edges[lo-1] < pos
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (edges[mid] < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct SpanOut {
    SmMsg   *slots;
    uint32_t cap;
    uint32_t n_msgs;
    uint32_t *overflow;       // null => count only (probe runs)
    uint32_t steps = 0, errs = 0, singles = 0;   // diagnostics (OOKD_DEBUG): loop iterations, ERRORs, single-sample steps
};

__device__ __forceinline__ void sm_emit(SpanOut &o, const SmCarry &s, i64 pos)
{
    if (o.n_msgs < o.cap) {
        SmMsg m;
        m.out_sample = (u64) pos;
        m.num_bits = s.num_bits;
        m.pad = 0;
        m.data[0] = s.data[0]; m.data[1] = s.data[1]; m.data[2] = s.data[2]; m.data[3] = s.data[3];
        o.slots[o.n_msgs] = m;
    } else if (o.overflow) {
        atomicExch(o.overflow, 1u);
    }
    o.n_msgs++;
}

// Run the machine over outputs [pos, end) from state s.  e = index of the first edge at or after
// pos, tb = true decision at pos-1.  PROBE: stop early and report whether a machine started in
// RESET at pos gets as far as appending a bit / emitting a message (1) or falls back to RESET (0).
template <bool PROBE>
__device__ __forceinline__ int sm_run_span(const SmArgs &a, const u64 n_edges, const SmTable &T, SmCarry &s, i64 pos, i64 end,
                                           u64 e, uint32_t tb, SpanOut &o, i64 chunk_lo)
{
    bool left_reset = false;
    const u64 INF = ~0ull;
    u64 next_edge = (e < n_edges) ? a.edges[e] : INF;
    u64 after_edge = (e + 1 < n_edges) ? a.edges[e + 1] : INF;   // one-ahead prefetch

    if (s.prev & OOKD_CARRY_DROPPING) {
        // The previous span ended inside a buffer device_process had given up on (span boundaries sit on
        // anchors, not on buffer boundaries): keep skipping to that buffer's end.
        s.prev &= 1u;
        i64 nb = next_buffer_start(a, pos, chunk_lo);
        const bool more = nb > end;
        if (more) nb = end;
        if (next_edge < (u64) nb) {
            const u64 e2 = e + edge_lower_bound_near(a.edges + e, n_edges - e, (u64) nb);
            tb ^= (uint32_t) ((e2 - e) & 1);
            e = e2;
            next_edge = (e < n_edges) ? a.edges[e] : INF;
            after_edge = (e + 1 < n_edges) ? a.edges[e + 1] : INF;
        }
        pos = nb;
        if (more) s.prev |= OOKD_CARRY_DROPPING;
    }

    while (pos < end) {
        const bool at_edge = (next_edge == (u64) pos);
        if (s.state == 0 || s.prev != tb || at_edge) {
            const uint32_t b = at_edge ? (tb ^ 1u) : tb;
            const int r = sm_step(T, s, b);
            if (at_edge) {
                tb ^= 1u;
                e++;
                next_edge = after_edge;
                after_edge = (e + 1 < n_edges) ? a.edges[e + 1] : INF;
            }
            pos++;
            if (PROBE) {
                if (r > 0 || s.num_bits > 0) return 1;
                if (s.state != 0) left_reset = true;
                if (r < 0 || (left_reset && s.state == 0)) return 0;
            }
            if (r > 0) {
                sm_emit(o, s, pos - 1);
            } else if (r < 0) {
                // device_process gives up on this buffer: resume at the next buffer's first output
                i64 nb = next_buffer_start(a, pos - 1, chunk_lo);
                if (nb > end) {
                    nb = end;
                    s.prev |= OOKD_CARRY_DROPPING;           // the next span goes on skipping
                }
                if (nb > pos) {
                    if (next_edge < (u64) nb) {
                        const u64 e2 = e + edge_lower_bound_near(a.edges + e, n_edges - e, (u64) nb);
                        tb ^= (uint32_t) ((e2 - e) & 1);
                        e = e2;
                        next_edge = (e < n_edges) ? a.edges[e] : INF;
                        after_edge = (e + 1 < n_edges) ? a.edges[e + 1] : INF;
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
            if (PROBE) {
                if (r > 0 || s.num_bits > 0) return 1;
                if (s.state == 0) return 0;
            }
            if (r > 0) {
                sm_emit(o, s, pos - 1);
            }
        } else {
            const uint32_t ksat = T.states[s.state].ksat;
            const u64 kk = (u64) s.k + gap;
            s.k = (kk < (u64) ksat) ? (uint32_t) kk : ksat;
            pos = (i64) limit;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative form of the same machine (machines with <= 32 triggers and <= 32 states).
//
// A run is one chain of dependent steps, so its cost is latency per step.  The per-thread form above
// walks the state's trigger list with shared-memory loads and data-dependent branches (~200 instructions
// and ~2500 cycles per edge).  Here lane j owns trigger j of the compiled machine in REGISTERS, the carry is
// replicated in every lane, and one evaluation is: every lane tests its own trigger, a ballot picks the
// first eligible one in list order (= lowest lane), a shuffle broadcasts its action.  No loop, no
// shared memory, no divergence; the first quiet firing count is a warp min-reduction.  Semantics are
// those of sm_eval / sm_step / sm_next_quiet_fire / sm_run_span<false> line by line.
// ---------------------------------------------------------------------------------------
struct WarpSm {
    uint32_t t_state;         // state owning this lane's trigger (0xFFFFFFFF: lane has no trigger)
    uint32_t t_cond, t_pack;  // condition; action | next_state << 8
    uint32_t t_kmin, t_kmax;
    uint32_t s_dmin, s_dmax, s_ktimeout, s_ksat;   // of that state
    uint32_t ksat_by_state;   // lane s: ksat of state s
    uint32_t max_bits;
};

__device__ __forceinline__ bool warp_sm_supported(const SmTable *tab)
{
    return tab->num_triggers <= 32 && tab->num_states <= 32;
}

__device__ __forceinline__ void warp_sm_load(WarpSm &W, const SmTable *tab, uint32_t lane)
{
    W.max_bits = tab->max_bits;
    W.t_state = 0xFFFFFFFFu;
    W.t_cond = 0; W.t_pack = 0; W.t_kmin = 1; W.t_kmax = 0;
    W.s_dmin = 0; W.s_dmax = 0; W.s_ktimeout = OOKD_K_INF; W.s_ksat = 0;
    W.ksat_by_state = (lane < tab->num_states) ? tab->states[lane].ksat : 0u;
    if (lane < tab->num_triggers) {
        const ookd_sm_trigger_k t = tab->triggers[lane];
        W.t_cond = (uint32_t) t.cond;
        W.t_pack = (uint32_t) t.action | (t.next_state << 8);
        W.t_kmin = t.kmin; W.t_kmax = t.kmax;
        for (uint32_t st = 0; st < tab->num_states; st++) {
            const ookd_sm_state_k x = tab->states[st];
            if (lane >= x.first_trigger && lane < x.first_trigger + x.num_triggers) {
                W.t_state = st;
                W.s_dmin = x.dmin; W.s_dmax = x.dmax; W.s_ktimeout = x.ktimeout; W.s_ksat = x.ksat;
            }
        }
    }
}

// sm_apply with the fired trigger's packed (action, next)
__device__ __forceinline__ int warp_sm_apply(const WarpSm &W, SmCarry &s, uint32_t pack)
{
    const uint32_t action = pack & 0xFFu;
    int result = 0;
    if (action == OOKD_ACT_APPEND_0 || action == OOKD_ACT_APPEND_1) {
        if (s.num_bits <= W.max_bits && s.num_bits < 256) {
            const u64 m = 1ull << (s.num_bits & 63);
            const u64 set = (action == OOKD_ACT_APPEND_1) ? m : 0ull;
            if (s.num_bits < 64) {                              // (uniform) the common case: one word
                s.data[0] = (s.data[0] & ~m) | set;
            } else {
                const uint32_t w = s.num_bits >> 6;
#pragma unroll
                for (uint32_t i = 1; i < 4; i++) {
                    if (i == w) s.data[i] = (s.data[i] & ~m) | set;
                }
            }
        }
        s.num_bits++;
    } else if (action == OOKD_ACT_OUTPUT_DATA) {
        result = 1;
    }
    s.state = pack >> 8;
    return result;
}

// sm_eval: one trigger evaluation on sample value b
__device__ __forceinline__ int warp_sm_eval(const WarpSm &W, SmCarry &s, uint32_t b)
{
    const uint32_t k = s.k;
    const bool rise = !s.prev && b, fall = s.prev && !b;
    const bool cond_ok = (W.t_cond == OOKD_COND_ALWAYS) || (W.t_cond == OOKD_COND_PULSE_START && rise) ||
                         (W.t_cond == OOKD_COND_PULSE_END && fall) ||
                         (W.t_cond == OOKD_COND_TIMEOUT && W.s_ktimeout != OOKD_K_INF && k >= W.s_ktimeout) ||
                         (W.t_cond == OOKD_COND_MSG_COMPLETE && s.num_bits >= W.max_bits);
    const bool elig = (W.t_state == s.state) && k >= W.t_kmin && k <= W.t_kmax && cond_ok;
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, elig);
    if (mask == 0) {
        const uint32_t ksat = __shfl_sync(0xFFFFFFFFu, W.ksat_by_state, (int) s.state);
        s.k = (k + 1 < ksat) ? k + 1 : ksat;
        return 0;
    }
    const int f = __ffs(mask) - 1;
    const bool is_edge = (W.t_cond == OOKD_COND_PULSE_START || W.t_cond == OOKD_COND_PULSE_END);
    const uint32_t dur = __ballot_sync(0xFFFFFFFFu, !is_edge || (k >= W.s_dmin && k <= W.s_dmax));
    const uint32_t pack = __shfl_sync(0xFFFFFFFFu, W.t_pack, f);
    int result;
    if ((dur >> f) & 1u) {
        result = warp_sm_apply(W, s, pack);
    } else {
        result = -1;
        s.state = 0;
    }
    s.k = 0;
    return result;
}

// sm_step: process() + the prev_bit update
__device__ __forceinline__ int warp_sm_step(const WarpSm &W, SmCarry &s, uint32_t b)
{
    int r = 0;
    if (s.state == 0) {
        s.num_bits = 0;
        const uint32_t nbytes = (W.max_bits + 7) >> 3;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t lo = 8u * w;
            if (nbytes >= lo + 8) {
                s.data[w] = 0;
            } else if (nbytes > lo) {
                s.data[w] &= ~((1ull << (8 * (nbytes - lo))) - 1);
            }
        }
        r = warp_sm_eval(W, s, b);
    }
    if (r == 0) {
        r = warp_sm_eval(W, s, b);
    }
    s.prev = b;
    return r;
}

// sm_next_quiet_fire: first count >= s.k at which a trigger that needs no edge fires; lane of that trigger or -1
__device__ __forceinline__ int warp_sm_next_quiet_fire(const WarpSm &W, const SmCarry &s, uint32_t *k_fire)
{
    bool applicable = (W.t_state == s.state);
    uint32_t lo = W.t_kmin;
    if (W.t_cond == OOKD_COND_TIMEOUT) {
        applicable = applicable && (W.s_ktimeout != OOKD_K_INF);
        lo = max(lo, W.s_ktimeout);
    } else if (W.t_cond == OOKD_COND_MSG_COMPLETE) {
        applicable = applicable && (s.num_bits >= W.max_bits);
    } else if (W.t_cond != OOKD_COND_ALWAYS) {
        applicable = false;
    }
    uint32_t kk = max(lo, s.k);
    if (!applicable || kk > W.t_kmax) kk = OOKD_K_INF;
    const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, kk);
    *k_fire = m;
    if (m == OOKD_K_INF) return -1;
    return __ffs(__ballot_sync(0xFFFFFFFFu, kk == m)) - 1;     // earlier list position wins ties
}

// sm_run_span<false>, executed by a whole warp with uniform control flow; lane 0 emits the messages.
// Positions are 32-bit offsets from chunk_lo (the caller guarantees end - chunk_lo < 2^31), and the edge list
// is consumed through a register window: lane j holds edge ebase + j of the current batch of 32 and of the
// next one, so the dependent chain sees a shuffle (~25 cycles) instead of an L2 round trip per edge.
struct WarpEdges {
    const u64 *edges;
    u64 n_edges, ebase, org;      // org = chunk_lo as an edge value
    uint32_t cur, nxt;            // offsets of edges ebase + lane / ebase + 32 + lane (0xFFFFFFFF: none / beyond range)
    uint32_t lane;

    __device__ __forceinline__ uint32_t fetch(u64 idx) const
    {
        if (idx >= n_edges) return 0xFFFFFFFFu;
        const u64 d = edges[idx] - org;
        return d < 0xFFFFFFFFull ? (uint32_t) d : 0xFFFFFFFFu;
    }
    __device__ __forceinline__ void reset(u64 e)
    {
        ebase = e;
        cur = fetch(e + lane);
        nxt = fetch(e + 32 + lane);
    }
    // Index of the first edge at or after e whose offset is >= nb (the next buffer's first output after an ERROR).
    // Dropped buffers come in runs (a broken message raises an ERROR in almost every buffer it touches), so the answer
    // is almost always among the 64 edges already held in registers: one or two ballots instead of a chain of
    // dependent loads per dropped buffer.
    __device__ __forceinline__ u64 first_at_or_after(u64 e, uint32_t nb, u64 nb64)
    {
        if (e - ebase < 64) {
            const bool in1 = (ebase + lane) >= e && cur < nb;            // (0xFFFFFFFF = no such edge: never below nb)
            const bool in2 = (ebase + 32 + lane) >= e && nxt < nb;
            const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, in1), m2 = __ballot_sync(0xFFFFFFFFu, in2);
            const uint32_t last = __shfl_sync(0xFFFFFFFFu, nxt, 31);
            if (last >= nb) return e + (u64) (__popc(m1) + __popc(m2));  // the window reaches past nb (or past the last edge)
            const u64 from = ebase + 64;                                  // everything held is below nb: search on from there
            return from + edge_lower_bound_near(edges + from, n_edges - from, nb64);
        }
        return e + edge_lower_bound_near(edges + e, n_edges - e, nb64);
    }
    // offset of edge e (e >= ebase); slides the window when e has left the current batch
    __device__ __forceinline__ uint32_t at(u64 e)
    {
        if (e - ebase >= 32) {
            if (e - ebase >= 64) {
                reset(e);
            } else {
                cur = nxt;
                ebase += 32;
                nxt = fetch(ebase + 32 + lane);
            }
        }
        return __shfl_sync(0xFFFFFFFFu, cur, (int) (e - ebase));
    }
};

__device__ __forceinline__ void warp_sm_run_span(const SmArgs &a, const u64 n_edges, const WarpSm &W, SmCarry &s, i64 pos64,
                                                 i64 end64, u64 e, uint32_t tb, SpanOut &o, i64 chunk_lo, uint32_t lane)
{
    const uint32_t NONE = 0xFFFFFFFFu;
    uint32_t pos = (uint32_t) (pos64 - chunk_lo);
    const uint32_t end = (uint32_t) (end64 - chunk_lo);
    WarpEdges E;
    E.edges = a.edges; E.n_edges = n_edges; E.org = (u64) chunk_lo; E.lane = lane;
    E.reset(e);
    uint32_t next_edge = E.at(e);                                       // NONE: no further edge in range

    if (s.prev & OOKD_CARRY_DROPPING) {
        // the previous span ended inside a buffer device_process had given up on: keep skipping to its end
        s.prev &= 1u;
        i64 nb64 = next_buffer_start(a, pos64, chunk_lo);
        const bool more = nb64 > end64;
        if (more) nb64 = end64;
        const uint32_t nb = (uint32_t) (nb64 - chunk_lo);
        if (next_edge < nb) {
            const u64 e2 = E.first_at_or_after(e, nb, (u64) nb64);
            tb ^= (uint32_t) ((e2 - e) & 1);
            e = e2;
            next_edge = E.at(e);
        }
        pos = nb;
        if (more) s.prev |= OOKD_CARRY_DROPPING;
    }

    while (pos < end) {
        const bool at_edge = (next_edge == pos);
        o.steps++;
        if (s.state == 0 || s.prev != tb) {
            // single-sample path: RESET (evaluated twice) and the first sample after a dropped buffer tail
            const uint32_t b = at_edge ? (tb ^ 1u) : tb;
            o.singles++;
            const int r = warp_sm_step(W, s, b);
            if (r < 0) o.errs++;
            if (at_edge) {
                tb ^= 1u;
                e++;
                next_edge = E.at(e);
            }
            pos++;
            if (r > 0) {
                if (lane == 0) sm_emit(o, s, chunk_lo + (i64) pos - 1); else o.n_msgs++;
            } else if (r < 0) {
                // device_process gives up on this buffer: resume at the next buffer's first output
                i64 nb64 = next_buffer_start(a, chunk_lo + (i64) pos - 1, chunk_lo);
                if (nb64 > end64) {
                    nb64 = end64;
                    s.prev |= OOKD_CARRY_DROPPING;
                }
                const uint32_t nb = (uint32_t) (nb64 - chunk_lo);
                if (nb > pos) {
                    if (next_edge < nb) {
                        const u64 e2 = E.first_at_or_after(e, nb, (u64) nb64);
                        tb ^= (uint32_t) ((e2 - e) & 1);
                        e = e2;
                        next_edge = E.at(e);
                    }
                    pos = nb;
                }
            }
            continue;
        }

        // ---- fused step: the quiet stretch up to the next edge AND the edge sample itself, in one evaluation ----
        // machine not in RESET, prev_bit == true bit, samples [pos, limit) constant.  Every lane computes the
        // offset (from pos) at which ITS trigger would be the first to fire: a trigger that needs no edge at the
        // first count inside its window, any trigger at the edge sample (offset g) if eligible there with the
        // count the machine has by then; the earliest offset wins, list order (lowest lane) breaks ties --
        // exactly what sm_next_quiet_fire followed by sm_eval on the edge sample compute in two iterations.
        const bool have_edge = next_edge < end;
        const uint32_t limit = have_edge ? next_edge : end;
        const uint32_t g = limit - pos;
        const uint32_t k = s.k;
        const uint32_t b = tb ^ 1u;                                     // value of the edge sample
        const bool in_state = (W.t_state == s.state);
        const uint32_t k_e = min(k + g, W.s_ksat);                      // count at the edge sample (k, g < 2^31)
        const bool mc = s.num_bits >= W.max_bits;
        const bool is_to = W.t_cond == OOKD_COND_TIMEOUT, is_mc = W.t_cond == OOKD_COND_MSG_COMPLETE;
        const bool is_quiet = W.t_cond == OOKD_COND_ALWAYS || is_to || is_mc;
        // (1) first quiet firing count
        uint32_t lo = W.t_kmin;
        if (is_to) lo = max(lo, W.s_ktimeout);
        const uint32_t kk = max(lo, k);
        const bool q_ok = in_state && is_quiet && (!is_mc || mc) && (!is_to || W.s_ktimeout != OOKD_K_INF) && kk <= W.t_kmax;
        const uint32_t tau_q = q_ok ? kk - k : OOKD_K_INF;
        // (2) eligibility at the edge sample
        const bool cond_e = (W.t_cond == OOKD_COND_ALWAYS) || (W.t_cond == OOKD_COND_PULSE_START && b) ||
                            (W.t_cond == OOKD_COND_PULSE_END && !b) ||
                            (is_to && W.s_ktimeout != OOKD_K_INF && k_e >= W.s_ktimeout) || (is_mc && mc);
        const bool e_ok = in_state && have_edge && k_e >= W.t_kmin && k_e <= W.t_kmax && cond_e;
        const uint32_t tau = (tau_q < g) ? tau_q : (e_ok ? g : OOKD_K_INF);
        // earliest offset, ties to the lowest lane (= list order): ONE reduction over (tau << 5 | lane); offsets
        // beyond 2^27 - 2 samples (never in practice) take the two-step form
        uint32_t m;
        int f;
        if (g < (1u << 27) - 1) {
            const uint32_t key = (tau == OOKD_K_INF) ? 0xFFFFFFFFu : ((tau << 5) | lane);
            const uint32_t best = __reduce_min_sync(0xFFFFFFFFu, key);
            m = (best == 0xFFFFFFFFu) ? OOKD_K_INF : (best >> 5);
            f = (int) (best & 31u);
        } else {
            m = __reduce_min_sync(0xFFFFFFFFu, tau);
            f = (m == OOKD_K_INF) ? 0 : __ffs(__ballot_sync(0xFFFFFFFFu, tau == m)) - 1;
        }
        if (m == OOKD_K_INF) {
            // nothing fires up to and including the edge sample
            const uint32_t ksat = __shfl_sync(0xFFFFFFFFu, W.ksat_by_state, (int) s.state);
            const u64 kn = (u64) k + g + (have_edge ? 1u : 0u);
            s.k = (kn < (u64) ksat) ? (uint32_t) kn : ksat;
            pos = limit;
            if (have_edge) {
                s.prev = b;
                tb ^= 1u;
                e++;
                next_edge = E.at(e);
                pos++;
            }
            continue;
        }
        // the winner's (action, next state), and in bit 31 whether its state-duration check passes on the edge sample
        const bool is_edge = (W.t_cond == OOKD_COND_PULSE_START || W.t_cond == OOKD_COND_PULSE_END);
        const uint32_t dur_ok = (!is_edge || (k_e >= W.s_dmin && k_e <= W.s_dmax)) ? 0x80000000u : 0u;
        const uint32_t packd = __shfl_sync(0xFFFFFFFFu, W.t_pack | dur_ok, f);
        const uint32_t pack = packd & 0x7FFFFFFFu;
        if (m < g) {
            // a trigger that needs no edge fires inside the quiet stretch
            pos += m;
            const int r = warp_sm_apply(W, s, pack);
            s.k = 0;
            pos++;
            if (r > 0) {
                if (lane == 0) sm_emit(o, s, chunk_lo + (i64) pos - 1); else o.n_msgs++;
            }
            continue;
        }
        // fires on the edge sample, with count k_e
        int r;
        if (packd >> 31) {
            r = warp_sm_apply(W, s, pack);
        } else {
            r = -1;
            s.state = 0;
            o.errs++;
        }
        s.k = 0;
        s.prev = b;
        tb ^= 1u;
        e++;
        next_edge = E.at(e);
        pos = limit + 1;
        if (r > 0) {
            if (lane == 0) sm_emit(o, s, chunk_lo + (i64) pos - 1); else o.n_msgs++;
        } else if (r < 0) {
            i64 nb64 = next_buffer_start(a, chunk_lo + (i64) pos - 1, chunk_lo);
            if (nb64 > end64) {
                nb64 = end64;
                s.prev |= OOKD_CARRY_DROPPING;
            }
            const uint32_t nb = (uint32_t) (nb64 - chunk_lo);
            if (nb > pos) {
                if (next_edge < nb) {
                    const u64 e2 = E.first_at_or_after(e, nb, (u64) nb64);
                    tb ^= (uint32_t) ((e2 - e) & 1);
                    e = e2;
                    next_edge = E.at(e);
                }
                pos = nb;
            }
        }
    }
}

// Copy the used part of the compiled machine into shared memory (header, states, triggers).
__device__ __forceinline__ void load_table(SmTable &T, const SmTable *src_tab)
{
    const uint32_t ns = src_tab->num_states, nt = src_tab->num_triggers;
    if (threadIdx.x == 0) {
        T.num_states = ns; T.num_triggers = nt; T.max_bits = src_tab->max_bits; T.k_sat = src_tab->k_sat;
    }
    const uint32_t *ss = (const uint32_t *) src_tab->states, *st = (const uint32_t *) src_tab->triggers;
    uint32_t *ds = (uint32_t *) T.states, *dt = (uint32_t *) T.triggers;
    for (uint32_t i = threadIdx.x; i < ns * (sizeof(ookd_sm_state_k) / 4); i += blockDim.x) ds[i] = ss[i];
    for (uint32_t i = threadIdx.x; i < nt * (sizeof(ookd_sm_trigger_k) / 4); i += blockDim.x) dt[i] = st[i];
    __syncthreads();
}

__device__ __forceinline__ void chunk_bounds_fixed(const SmArgs &a, uint32_t c, i64 &start, i64 &end)
{
    start = (c == 0) ? a.out_lo : first_output_of_buffer(a, a.first_buffer + (u64) c * a.chunk_buffers);
    end = first_output_of_buffer(a, a.first_buffer + (u64) (c + 1) * a.chunk_buffers);
    if (end > a.out_hi || c == a.n_chunks - 1) end = a.out_hi;
}

// Span of chunk c: between its boundary and the next chunk's (boundaries sit on anchors where sm_anchor_kernel
// found one).  lo = the chunk's fixed start: a buffer boundary at or before the span, the origin of the
// 32-bit offsets and of next_buffer_start().
__device__ __forceinline__ void chunk_bounds(const SmArgs &a, uint32_t c, i64 &start, i64 &end, i64 &lo)
{
    i64 fixed_end;
    chunk_bounds_fixed(a, c, lo, fixed_end);
    if (a.bound_pos) {
        start = a.bound_pos[c];
        end = (c + 1 < a.n_chunks) ? a.bound_pos[c + 1] : a.out_hi;
    } else {
        start = lo;
        end = fixed_end;
    }
}

__device__ __forceinline__ void carry_reset(SmCarry &s, uint32_t prev)
{
    s.state = 0; s.k = 0; s.num_bits = 0; s.prev = prev;
    s.data[0] = s.data[1] = s.data[2] = s.data[3] = 0;
}

// ---------------------------------------------------------------------------------------
// Table speculation: thread (c, j) = chunk c, pair slot j.
// ---------------------------------------------------------------------------------------
#define OOKD_TAB_INVALID 0xFFFFFFFFu      // entry.state of a seed that matches no real entry

// edge count / decision in front of the shard: by value, or from the header the edge pass wrote
#define OOKD_SM_EDGE_HDR(a)                                                                     \
    const u64 n_edges = (a).hdr ? (a).hdr->n_edges : (a).n_edges;                               \
    const uint32_t base_bit = (a).hdr ? (a).hdr->base_bit : (a).base_bit;

// pairs a chunk has before round 0: its seed's, unless it has none (OOKD_SEED_NONE = 4, below)
__global__ void seed_count_kernel(const uint8_t *kind, uint32_t *cnt, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = (kind[i] == 4) ? 0u : 1u;
}

// Anchors, once per decode.  One warp per chunk: the chunk's first "anchor" is the first rising edge from which
// a freshly reset machine gets as far as appending a bit, i.e. a plausible message start (the 32 lanes probe 32
// candidate edges at once).  The true run, whatever it did before, is normally idle when a message starts, so
// the chunk's boundary is MOVED to its anchor and its table gets the pair "idle machine at the anchor -> ..."
// as a real entry: when the previous chunk's run indeed arrives idle, the chain links up after round 0.
// A chunk without an anchor keeps its fixed boundary and is seeded with RESET at its first sample (a guess
// that usually lands mid-message; later rounds replace it).
// 32-ary lower bound by a whole warp: four dependent loads for 2^20 edges instead of twenty (the chunk's first
// edge is the first thing every anchor warp needs).
__device__ __forceinline__ u64 warp_edge_lower_bound(const u64 *edges, u64 n, u64 pos, uint32_t lane)
{
    u64 lo = 0, hi = n;                                      // answer in [lo, hi]
    while (hi - lo > 32) {
        const u64 step = (hi - lo + 32) / 33;                // 32 pivots lo + step*(lane+1) - 1
        const u64 idx = lo + step * (lane + 1) - 1;
        const bool less = idx < hi && edges[idx] < pos;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, less);
        const uint32_t cnt = (uint32_t) __popc(m);           // pivots are ascending: `less` is a prefix
        const u64 new_lo = cnt ? lo + step * cnt : lo;
        const u64 new_hi = (cnt < 32) ? ((lo + step * (cnt + 1) - 1 < hi) ? lo + step * (cnt + 1) - 1 : hi) : hi;
        lo = new_lo;
        hi = new_hi;
    }
    const u64 idx = lo + lane;
    const bool less = idx < hi && edges[idx] < pos;
    return lo + (u64) __popc(__ballot_sync(0xFFFFFFFFu, less));
}

// One warp, one chunk (T: the compiled machine, in shared or global memory).
__device__ __forceinline__ void sm_anchor_chunk(const SmArgs &a, const SmTable &T, uint32_t c, uint32_t lane, const u64 n_edges,
                                                const uint32_t base_bit)
{
    i64 start, end;
    chunk_bounds_fixed(a, c, start, end);
    const u64 e = warp_edge_lower_bound(a.edges, n_edges, (u64) start, lane);
    const uint32_t tb = base_bit ^ (uint32_t) (e & 1);       // true decision at start-1
    bool have_anchor = false;
    u64 anchor_e = 0;
    if (!(c == 0 && !a.warm)) {
        const u64 er0 = e + (tb == 1 ? 1 : 0);              // edge e falls when tb == 1; the next one rises
        for (int batch = 0; batch < 3 && !have_anchor; batch++) {
            const u64 er = er0 + 2 * (u64) (batch * 32 + lane);
            const bool cand = er < n_edges && a.edges[er] < (u64) end;
            bool alive = false;
            if (cand) {
                SmCarry p;
                carry_reset(p, 0);
                SpanOut po;
                po.slots = nullptr; po.cap = 0; po.n_msgs = 0; po.overflow = nullptr;
                alive = sm_run_span<true>(a, n_edges, T, p, (i64) a.edges[er], a.out_hi, er, 0, po, start) != 0;
            }
            const uint32_t m_alive = __ballot_sync(0xFFFFFFFFu, alive);
            const uint32_t m_cand = __ballot_sync(0xFFFFFFFFu, cand);
            if (m_alive) {
                have_anchor = true;
                anchor_e = er0 + 2 * (u64) (batch * 32 + (__ffs(m_alive) - 1));
            }
            if (m_cand != 0xFFFFFFFFu) break;                // ran out of rising edges in this chunk
        }
    }
    if (lane != 0) return;
    i64 bound = start, seed = start;
    u64 bound_e = e, seed_e = e;
    uint8_t kind;
    if (c == 0 && !a.warm) {
        kind = OOKD_SEED_TRUE;
    } else if (have_anchor) {
        seed = (i64) a.edges[anchor_e];
        seed_e = anchor_e;
        kind = OOKD_SEED_CANON;
        bound = seed;
        bound_e = seed_e;
    } else if (c == 0) {
        kind = OOKD_SEED_RESET;                              // (warm shard: somebody has to start the chain)
    } else if (e < n_edges && a.edges[e] < (u64) end) {
        kind = OOKD_SEED_NONE;
    } else {
        kind = (tb == 0) ? OOKD_SEED_IDLE : OOKD_SEED_RESET;
    }
    a.bound_pos[c] = bound;
    a.chunk_e[c] = bound_e;
    a.seed_pos[c] = seed;
    a.seed_e[c] = seed_e;
    a.seed_kind[c] = kind;
}

__global__ void __launch_bounds__(32) sm_anchor_kernel(const SmArgs a)
{
    OOKD_SM_EDGE_HDR(a)
    __shared__ SmTable T;
    const uint32_t c = blockIdx.x;
    if (c >= a.n_chunks) return;
    load_table(T, a.tab);
    sm_anchor_chunk(a, T, c, threadIdx.x & 31, n_edges, base_bit);
    if (threadIdx.x == 0 && a.cnt_out) a.cnt_out[c] = (a.seed_kind[c] == OOKD_SEED_NONE) ? 0u : 1u;   // the seed's pair (slot 0)
}

// One WARP per (chunk, slot): the work is a chain of dependent steps, so what matters is latency, not lanes; giving
// every run its own warp keeps runs from serialising each other through divergence.
// sm_round_pair is the body for pair (c, j) of round a.round; cnt_in = pairs complete before this round, cnt_out = slots
// handed out (initialised to cnt_in; atomically incremented).  T: compiled machine in shared or global memory.
constexpr int SM_ROUND_WARPS = 4;                            // (chunk, slot) pairs per CTA

__device__ __forceinline__ void sm_round_pair(const SmArgs &a, const SmTable &T, const WarpSm &W, const bool warp_capable,
                                              uint32_t c, uint32_t j, uint32_t lane, const u64 n_edges, const uint32_t base_bit)
{
    const uint32_t K = a.tab_k;
    i64 start, end, lo;
    chunk_bounds(a, c, start, end, lo);
    const bool warp_ok = warp_capable && (end - lo) < (1ll << 31);

    u64 e = a.chunk_e[c];
    uint32_t tb = base_bit ^ (uint32_t) (e & 1);             // true decision at start-1

    SmCarry s, entry;
    i64 pos = start;
    uint32_t slot;
    if (a.round == 0) {
        // one speculative seed per chunk, prepared by sm_anchor_chunk
        if (!warp_ok && lane != 0) return;
        const uint32_t kind = a.seed_kind[c];
        if (kind == OOKD_SEED_NONE) return;
        pos = a.seed_pos[c];
        e = a.seed_e[c];
        tb = base_bit ^ (uint32_t) (e & 1);
        if (kind == OOKD_SEED_TRUE) {
            s = a.entry_ptr ? *a.entry_ptr : a.entry0;
            if (s.state < T.num_states && s.k > T.states[s.state].ksat) s.k = T.states[s.state].ksat;
            entry = s;
        } else if (kind == OOKD_SEED_CANON || kind == OOKD_SEED_IDLE) {
            s = a.canon;
            entry = s;
        } else {
            carry_reset(s, tb);
            entry = s;
        }
        slot = 0;                                            // (every chunk's count starts at 1: the seed's slot)
    } else {
        if (!warp_ok && lane != 0) return;
        if (c == 0) return;
        const uint32_t n_prev = a.cnt_in[c - 1];
        if (j >= n_prev) return;
        s = a.tab_exit[(u64) (c - 1) * K + j];
        for (uint32_t i = 0; i < j; i++) {                 // duplicate among this round's candidates
            if (carry_equal(s, a.tab_exit[(u64) (c - 1) * K + i])) return;
        }
        const uint32_t n_here = a.cnt_in[c];
        for (uint32_t i = 0; i < n_here; i++) {            // already an entry of this chunk
            if (carry_equal(s, a.tab_entry[(u64) c * K + i])) return;
        }
        slot = 0;
        if (lane == 0) slot = atomicAdd(&a.cnt_out[c], 1u);
        if (warp_ok) slot = __shfl_sync(0xFFFFFFFFu, slot, 0);   // (warp_ok: all lanes are here; else only lane 0)
        if (slot >= K) {
            if (lane == 0) atomicExch(a.overflow, 2u);      // table full: host falls back
            return;
        }
        entry = s;
    }
    if (lane == 0) atomicAdd(&a.n_ran[a.counter_idx], 1u);

    // Run the chunk, and keep going: if the exit is not an entry of the NEXT chunk (round 0: not that chunk's seed),
    // the same warp adds that pair too and runs on.  A cascade of consecutive chunks entered in a state no table
    // holds (e.g. a string of messages lost to dropped buffers) is then repaired in ONE round -- its cost is the
    // chain itself -- instead of one round, link and walk per chunk.
    uint32_t cc = c;
    const long long t_dbg = a.dbg ? clock64() : 0;
    for (int hop = 0;; hop++) {
        SpanOut o;
        o.slots = a.slots + ((u64) cc * K + slot) * a.slot_cap;
        o.cap = a.slot_cap;
        o.n_msgs = 0;
        o.overflow = a.overflow;
        i64 upto = end;
        const bool through_report = a.warm && cc == 0 && pos < a.report_lo && a.report_lo <= end;
        if (through_report) upto = a.report_lo;              // warm shard: note the state at the shard's first output
        if (warp_ok) {
            warp_sm_run_span(a, n_edges, W, s, pos, upto, e, tb, o, lo, lane);
        } else {
            sm_run_span<false>(a, n_edges, T, s, pos, upto, e, tb, o, lo);
        }
        if (through_report) {
            if (lane == 0) a.mid_carry[slot] = s;
            if (upto < end) {
                e = warp_ok ? warp_edge_lower_bound(a.edges, n_edges, (u64) upto, lane) : edge_lower_bound(a.edges, n_edges, (u64) upto);
                tb = base_bit ^ (uint32_t) (e & 1);
                if (warp_ok) {
                    warp_sm_run_span(a, n_edges, W, s, upto, end, e, tb, o, lo, lane);
                } else {
                    sm_run_span<false>(a, n_edges, T, s, upto, end, e, tb, o, lo);
                }
            }
        }
        if (lane == 0) {
            a.tab_entry[(u64) cc * K + slot] = entry;
            a.tab_exit[(u64) cc * K + slot] = s;
            a.tab_nmsg[(u64) cc * K + slot] = (o.n_msgs < o.cap) ? o.n_msgs : o.cap;
            if (a.dbg && a.round == 0) {
                a.dbg[(u64) c * 4 + 0] = (unsigned long long) (clock64() - t_dbg);
                a.dbg[(u64) c * 4 + 1] += o.steps;
                a.dbg[(u64) c * 4 + 2] += ((unsigned long long) o.errs << 32) | o.singles;
                a.dbg[(u64) c * 4 + 3] = (unsigned long long) hop + 1;
            }
        }
        if (!warp_ok || hop >= 16 || cc + 1 >= a.n_chunks) return;
        // (all lanes hold the same carry; memory reads below are uniform)
        const uint32_t nn = cc + 1;
        bool known = false;
        if (a.round == 0) {
            // the only entry chunk nn has (or is getting right now, from its own warp) is its seed
            const uint32_t kind = a.seed_kind[nn];
            if (kind == OOKD_SEED_CANON || kind == OOKD_SEED_IDLE) {
                known = carry_equal(s, a.canon);
            } else if (kind == OOKD_SEED_RESET) {
                SmCarry r;
                carry_reset(r, base_bit ^ (uint32_t) (a.seed_e[nn] & 1));
                known = carry_equal(s, r);
            }
        } else {
            const uint32_t n_next = a.cnt_in[nn];            // entries that were complete before this round
            for (uint32_t i = 0; i < n_next; i++) {
                if (carry_equal(s, a.tab_entry[(u64) nn * K + i])) { known = true; break; }
            }
        }
        if (known) return;
        uint32_t nslot = 0;
        if (lane == 0) nslot = atomicAdd(&a.cnt_out[nn], 1u);
        nslot = __shfl_sync(0xFFFFFFFFu, nslot, 0);
        if (nslot >= K) {
            if (lane == 0) atomicExch(a.overflow, 2u);      // table full: host falls back
            return;
        }
        if (lane == 0) atomicAdd(&a.n_ran[a.counter_idx], 1u);
        cc = nn;
        slot = nslot;
        entry = s;
        chunk_bounds(a, cc, start, end, lo);
        if ((end - lo) >= (1ll << 31)) return;               // (cannot happen when the first chunk passed the check)
        pos = start;
        e = a.chunk_e[cc];
        tb = base_bit ^ (uint32_t) (e & 1);
    }
}

__global__ void __launch_bounds__(32 * SM_ROUND_WARPS) sm_table_round_kernel(const SmArgs a)
{
    OOKD_SM_EDGE_HDR(a)
    __shared__ SmTable T;
    __shared__ int s_any;
    const uint32_t K = a.tab_k;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t KR = (a.round == 0) ? 1u : K;             // warps per chunk this round
    const uint32_t gid = blockIdx.x * SM_ROUND_WARPS + (threadIdx.x >> 5);
    const uint32_t c = gid / KR, j = gid % KR;
    if (a.round >= 1 && a.walk_status[1]) return;            // an earlier walk of this burst already resolved the chain
    // cheap rejection before anything is staged: most (chunk, slot) pairs have nothing new to run
    bool live = c < a.n_chunks;
    if (live && a.round != 0) live = (c != 0) && j < a.cnt_in[c - 1];
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    if (live && lane == 0) s_any = 1;
    __syncthreads();
    if (!s_any) return;                                      // (CTA-uniform)
    load_table(T, a.tab);
    if (!live) return;
    const bool capable = warp_sm_supported(a.tab);
    WarpSm W;
    if (capable) warp_sm_load(W, &T, lane);
    sm_round_pair(a, T, W, capable, c, j, lane, n_edges, base_bit);
}

// Resolve: a corrected entry for chunk 0 (the shard's true entry state arrived from the previous
// shard).  Every other pair of every table stays valid -- pairs are functions of their entry only --
// so just chunk 0 gains a pair (unless it already has this entry) and the walk restarts from it.
__global__ void __launch_bounds__(32) sm_table_add_entry_kernel(const SmArgs a)
{
    // One WARP (uniform control flow; lane 0 writes): the span is a chain of dependent steps, and the warp-cooperative
    // evaluator takes a third of the time per step of the per-thread one -- this kernel is the latency of a resolve.
    __shared__ SmTable T;
    load_table(T, a.tab);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t K = a.tab_k;
    const uint32_t c = a.first_chunk;
    const uint32_t n_here = a.cnt_out[c];
    SmCarry s = a.entry0;
    if (!a.entry_at_report) {
        for (uint32_t i = 0; i < n_here; i++) {
            if (carry_equal(s, a.tab_entry[(u64) c * K + i])) { if (lane == 0) *a.start_slot = i; return; }
        }
    } else {
        // warm shard: chunk 0's pairs start in the history; the corrected state applies at the shard's first output
        for (uint32_t i = 0; i < n_here; i++) {
            if (carry_equal(s, a.mid_carry[i])) { if (lane == 0) *a.start_slot = i; return; }
        }
    }
    if (n_here >= K) { if (lane == 0) atomicExch(a.overflow, 2u); return; }
    i64 start, end, lo;
    chunk_bounds(a, c, start, end, lo);
    u64 e = a.chunk_e[c];
    if (a.entry_at_report) {
        start = a.report_lo;
        e = warp_edge_lower_bound(a.edges, a.n_edges, (u64) start, lane);
    }
    const uint32_t tb = a.base_bit ^ (uint32_t) (e & 1);
    SpanOut o;
    o.slots = a.slots + ((u64) c * K + n_here) * a.slot_cap;
    o.cap = a.slot_cap;
    o.n_msgs = 0;
    o.overflow = a.overflow;
    const SmCarry entry = s;
    const bool warp_ok = warp_sm_supported(a.tab) && (end - lo) < (1ll << 31);
    if (start < end) {
        if (warp_ok) {
            WarpSm W;
            warp_sm_load(W, &T, lane);
            warp_sm_run_span(a, a.n_edges, W, s, start, end, e, tb, o, lo, lane);
        } else if (lane == 0) {
            sm_run_span<false>(a, a.n_edges, T, s, start, end, e, tb, o, lo);
        }
    }
    if (lane != 0) return;
    a.tab_entry[(u64) c * K + n_here] = a.entry_at_report ? SmCarry{OOKD_TAB_INVALID, 0, 0, 0, {0, 0, 0, 0}} : entry;
    a.tab_exit[(u64) c * K + n_here] = s;
    a.tab_nmsg[(u64) c * K + n_here] = (o.n_msgs < o.cap) ? o.n_msgs : o.cap;
    if (a.entry_at_report) a.mid_carry[n_here] = entry;
    a.cnt_out[c] = n_here + 1;
    *a.start_slot = n_here;
}

// link[c][i] = slot of chunk c+1 whose entry equals exit[c][i] (0xFF if none)
__device__ __forceinline__ void sm_link_pair(const SmArgs &a, const uint32_t *cnt, uint32_t c, uint32_t i)
{
    const uint32_t K = a.tab_k;
    uint8_t l = 0xFF;
    if (c + 1 < a.n_chunks && i < min(cnt[c], K)) {
        const SmCarry x = a.tab_exit[(u64) c * K + i];
        const uint32_t n_next = min(cnt[c + 1], K);
        for (uint32_t q = 0; q < n_next; q++) {
            if (carry_equal(x, a.tab_entry[(u64) (c + 1) * K + q])) { l = (uint8_t) q; break; }
        }
    }
    a.link[(u64) c * K + i] = l;
}

__global__ void __launch_bounds__(128) sm_link_kernel(const SmArgs a)
{
    const uint32_t K = a.tab_k;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = gid / K, i = gid % K;
    if (c >= a.n_chunks) return;
    if (a.walk_status[1]) return;                            // already resolved: keep the links the walk used
    sm_link_pair(a, a.cnt_in, c, i);
    // the common case -- every seed pair links to the next chunk's seed pair -- is recognised without walking
    if (i == 0 && c >= a.first_chunk && c + 1 < a.n_chunks && a.link[(u64) c * K] != 0) {
        atomicAdd(&a.n_ran[8 + (a.round & 7)], 1u);
    }
}

// One CTA resolves the chain from chunk first_chunk / slot *start_slot.  Chasing 1 link per step through global
// memory would serialise n_chunks L2 latencies, so the link rows of a block of BLK chunks are first staged in
// shared memory (one coalesced pass), composed over segments of SEG chunks (one thread per segment, all K start
// slots at once), the short chain over segments is walked by one thread, and every segment thread then replays
// its own segment from its now-known entry slot.  Needs blockDim.x >= BLK / 32.
constexpr int SM_WALK_NT = 1024;

template <uint32_t BLK>
__device__ __forceinline__ void sm_walk_cta(const SmArgs &a)
{
    constexpr uint32_t SEG = 32, NSEG = BLK / SEG;
    __shared__ uint2 s_link[BLK];                            // link rows of the block (8 slots x 1 byte)
    __shared__ uint8_t s_map[NSEG * 8];                      // composed map of each segment
    __shared__ uint8_t s_in[NSEG];                           // entry slot of each segment (0xFF = unreachable)
    __shared__ uint32_t s_carry, s_max;
    const uint32_t K = a.tab_k;                              // == 8 (one 8-byte link row per chunk)
    if (threadIdx.x == 0) { s_carry = *a.start_slot; s_max = 0; }
    __syncthreads();
    uint32_t done = 0;

    for (uint32_t base = a.first_chunk; base < a.n_chunks; base += BLK) {
        const uint32_t n_here = min(BLK, a.n_chunks - base);
        const uint32_t n_seg = (n_here + SEG - 1) / SEG;
        for (uint32_t i = threadIdx.x; i < n_here; i += blockDim.x) {
            s_link[i] = *(const uint2 *) (a.link + (u64) (base + i) * 8);
        }
        __syncthreads();
        const uint32_t sg = threadIdx.x;
        const uint32_t l_lo = sg * SEG, l_hi = min(l_lo + SEG, n_here);       // block-local chunk range of the segment
        // compose: m[k] = slot of chunk l_hi reached when chunk l_lo is entered in slot k
        if (sg < n_seg) {
            uint32_t m[8];
#pragma unroll
            for (int k = 0; k < 8; k++) m[k] = (uint32_t) k;
            for (uint32_t l = l_lo; l < l_hi; l++) {
                const uint2 lw = s_link[l];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t cur = m[k];
                    const uint32_t word = (cur < 4) ? lw.x : lw.y;
                    m[k] = (cur == 0xFF) ? 0xFFu : ((word >> (8 * (cur & 3))) & 0xFFu);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) s_map[sg * 8 + k] = (uint8_t) m[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cur = s_carry, sgi = 0;
            for (; sgi < n_seg; sgi++) {
                s_in[sgi] = (uint8_t) cur;
                const uint32_t nx = s_map[sgi * 8 + cur];
                if (nx == 0xFF) { sgi++; break; }
                cur = nx;
            }
            for (uint32_t r = sgi; r < n_seg; r++) s_in[r] = 0xFF;
            s_carry = cur;
        }
        __syncthreads();
        // replay the own segment to record the chosen slot of every chunk
        if (sg < n_seg && s_in[sg] != 0xFF) {
            uint32_t cur = s_in[sg], my_done = 0;
            for (uint32_t l = l_lo; l < l_hi; l++) {
                const uint32_t c = base + l;
                a.chosen[c] = (uint8_t) cur;
                my_done = c + 1;
                if (c + 1 == a.n_chunks) break;
                const uint2 lw = s_link[l];
                const uint32_t word = (cur < 4) ? lw.x : lw.y;
                const uint32_t nx = (word >> (8 * (cur & 3))) & 0xFFu;
                if (nx == 0xFF) break;
                cur = nx;
            }
            atomicMax(&s_max, my_done);
        }
        __syncthreads();
        done = s_max;                                        // chunks [0, done) have a chosen pair
        if (done < base + n_here) break;                     // chain broke: another round is needed
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.n_ran[16 + (a.round & 15)] = done;                 // diagnostics: chunks resolved after round a.round
        a.walk_status[0] = done;
        a.walk_status[1] = (done == a.n_chunks) ? 1u : 0u;
        if (done == a.n_chunks) {
            *a.final_exit = a.tab_exit[(u64) (a.n_chunks - 1) * K + a.chosen[a.n_chunks - 1]];
            if (a.warm) *a.final_entry = a.mid_carry[a.chosen[0]];
        }
    }
    for (uint32_t c = threadIdx.x; c < a.n_chunks; c += blockDim.x) {
        a.msg_counts[c] = (c >= a.first_chunk && c < done) ? a.tab_nmsg[(u64) c * K + a.chosen[c]] : 0u;
    }
}

// Exclusive scan of the chosen pairs' message counts by one CTA (offsets[c], total); the copy itself is done by the
// chunks' own warps after the next barrier.
__device__ __forceinline__ void sm_scan_cta(const SmArgs &a, uint32_t *offsets, u64 base, u64 *total_out)
{
    __shared__ u64 s_w[32];
    const uint32_t nt = blockDim.x, nc = a.n_chunks;
    const uint32_t per = (nc + nt - 1) / nt;
    const uint32_t lo = min(nc, threadIdx.x * per), hi = min(nc, lo + per);
    u64 sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += a.msg_counts[i];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t) d) inc += up;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u64 before = 0, all = 0;
    for (uint32_t w = 0; w < (nt + 31) / 32; w++) {
        if (w < warp) before += s_w[w];
        all += s_w[w];
    }
    u64 run = before + inc - sum;
    for (uint32_t c = lo; c < hi; c++) {
        offsets[c] = (uint32_t) run;
        run += a.msg_counts[c];
    }
    if (threadIdx.x == 0) *total_out = base + all;
}

// gather_out != null: a walk that completes the chain also scans the chosen pairs' message counts and writes the
// ordered message list (the single-synchronisation tail: no separate scan / gather launches).
__global__ void __launch_bounds__(SM_WALK_NT) sm_walk_kernel(const SmArgs a, uint32_t *offsets, SmMsg *gather_out, u64 gather_cap,
                                                             u64 *n_msgs_out)
{
    if (a.walk_status[1]) return;                            // resolved by an earlier walk of this burst
    const uint32_t K = a.tab_k;
    if (a.n_ran[8 + (a.round & 7)] == 0 && *a.start_slot == 0 && a.cnt_in[a.first_chunk] >= 1) {
        // every seed pair links to the next one: the chain is the seed pairs
        for (uint32_t c = threadIdx.x; c < a.n_chunks; c += blockDim.x) {
            a.chosen[c] = 0;
            a.msg_counts[c] = (c >= a.first_chunk) ? a.tab_nmsg[(u64) c * K] : 0u;
        }
        if (threadIdx.x == 0) {
            a.n_ran[16 + (a.round & 15)] = a.n_chunks;
            a.walk_status[0] = a.n_chunks;
            a.walk_status[1] = 1u;
            *a.final_exit = a.tab_exit[(u64) (a.n_chunks - 1) * K];
            if (a.warm) *a.final_entry = a.mid_carry[0];
        }
    } else {
        sm_walk_cta<4096>(a);
    }
    if (!gather_out) return;
    __syncthreads();
    if (!a.walk_status[1]) return;                           // (written by this CTA's thread 0)
    sm_scan_cta(a, offsets, 0, n_msgs_out);
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < a.n_chunks; c += blockDim.x) {
        const uint32_t n = a.msg_counts[c];
        const SmMsg *src = a.slots + ((u64) c * K + a.chosen[c]) * a.slot_cap;
        for (uint32_t i = 0; i < n; i++) {
            if ((u64) offsets[c] + i < gather_cap) gather_out[offsets[c] + i] = src[i];
        }
    }
}

__global__ void sm_gather_table_kernel(const SmArgs a, const uint32_t *offsets, SmMsg *out, u64 out_cap)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_chunks) return;
    const uint32_t n = a.msg_counts[c];
    if (n == 0) return;
    const SmMsg *src = a.slots + ((u64) c * a.tab_k + a.chosen[c]) * a.slot_cap;
    for (uint32_t i = 0; i < n; i++) {
        if ((u64) offsets[c] + i < out_cap) out[offsets[c] + i] = src[i];
    }
}

// ---------------------------------------------------------------------------------------
// Fused form of the whole state-machine stage: anchors, seed round, link, walk, repair rounds UNTIL the chain
// resolves, message scan and gather in ONE cooperative launch.  The separate kernels above cost a launch boundary
// each (about a dozen per decode, all on the critical path behind the last sample) and had to enqueue repair rounds
// blindly; here the convergence loop runs on the device and the phases are separated by a grid barrier.
// Warp w of the grid owns chunks w, w + W, ...; the compiled machine is read from global memory (L1 resident), so
// the kernel's only shared memory is the walk's staging area and its CTAs fit beside a running screening kernel.
// ---------------------------------------------------------------------------------------
struct SmFusedArgs {
    SmArgs a;                 // hdr / tables / anchors as for the separate kernels (cnt_in / cnt_out are set here)
    uint32_t *cnt_done;       // [n_chunks] pairs complete as of the last barrier
    uint32_t *cnt_alloc;      // [n_chunks] slots handed out
    uint32_t *bar;            // [2] grid barrier: arrivals, generation (zeroed once, at handle creation)
    uint32_t max_rounds;
    uint32_t *rounds_out;     // rounds run
    uint32_t *offsets;        // [n_chunks] exclusive message offsets of the chosen pairs
    SmMsg *msgs_out;          // ordered message list ...
    u64 msgs_cap;
    u64 *n_msgs_out;          // ... and its length (may exceed msgs_cap: then the host fetches again with room)
    u64 msgs_base;            // messages of earlier windows already in msgs_out (0 for a whole-shard decode)
    long long *stamps;        // null, or [16] phase time stamps (debug)
};

constexpr int SM_FUSED_NT = 128;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (co-resident: cooperative launch) grid arrive; the last one opens the next generation.
__device__ __forceinline__ void sm_grid_barrier(uint32_t *bar, uint32_t n_ctas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t gen = ld_acquire_u32(bar + 1);
        __threadfence();
        if (atomicAdd(bar, 1u) == n_ctas - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
        } else {
            uint32_t ns = 32;
            while (ld_acquire_u32(bar + 1) == gen) {
                __nanosleep(ns);
                if (ns < 512) ns *= 2;                       // (hundreds of pollers on one line: back off)
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// messages of chunk c's chosen pair -> their place in the ordered list (one warp)
__device__ __forceinline__ void sm_gather_chunk(const SmArgs &a, const uint32_t *offsets, SmMsg *out, u64 base, u64 cap,
                                                uint32_t c, uint32_t lane)
{
    const uint32_t n = a.msg_counts[c];
    if (n == 0) return;
    const u64 off = base + offsets[c];
    const u64 *src = (const u64 *) (a.slots + ((u64) c * a.tab_k + a.chosen[c]) * a.slot_cap);
    constexpr uint32_t W8 = sizeof(SmMsg) / 8;               // 8-byte words per message
    for (uint32_t i = lane; i < n * W8; i += 32) {
        if (off + i / W8 < cap) ((u64 *) (out + off))[i] = src[i];
    }
}

// phase stamps of CTA 0 (SM clock), for OOKD_DEBUG: 0 anchors done, 1 past the barrier, 2 round done, 3 links done (past
// the barrier), 4 walk + scan done, 5 past the barrier, 7 end; [8] = start
__device__ __forceinline__ long long global_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// debug: [0] = earliest start, [1 + i] = time the LAST CTA passed point i (global timer, ns)
#define STAMP(i) do { if (f.stamps && threadIdx.x == 0) atomicMax((unsigned long long *) f.stamps + 1 + (i), (unsigned long long) global_ns()); } while (0)

__global__ void __launch_bounds__(SM_FUSED_NT) sm_fused_kernel(const SmFusedArgs f)
{
    if (f.stamps && threadIdx.x == 0) atomicMin((unsigned long long *) f.stamps, (unsigned long long) global_ns());
    SmArgs a = f.a;
    OOKD_SM_EDGE_HDR(a)
    const SmTable &T = *a.tab;
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t WPC = SM_FUSED_NT / 32;
    const uint32_t gw = blockIdx.x * WPC + (threadIdx.x >> 5), n_gw = gridDim.x * WPC;
    const uint32_t nc = a.n_chunks, K = a.tab_k;
    const bool capable = warp_sm_supported(a.tab);
    WarpSm W;
    if (capable) warp_sm_load(W, a.tab, lane);
    a.cnt_in = f.cnt_done;
    a.cnt_out = f.cnt_alloc;

    // ---- anchors; every chunk starts with one pair: its seed ----
    for (uint32_t c = gw; c < nc; c += n_gw) {
        sm_anchor_chunk(a, T, c, lane, n_edges, base_bit);
        if (lane == 0) {
            const uint32_t n0 = (a.seed_kind[c] == OOKD_SEED_NONE) ? 0u : 1u;
            f.cnt_alloc[c] = n0;
            f.cnt_done[c] = n0;
        }
    }
    STAMP(0);
    sm_grid_barrier(f.bar, gridDim.x);
    STAMP(1);

    uint32_t round = 0, complete = 0;
    for (;;) {
        a.round = round;
        a.counter_idx = round & 31;
        if (round == 0) {
            for (uint32_t c = gw; c < nc; c += n_gw) sm_round_pair(a, T, W, capable, c, 0, lane, n_edges, base_bit);
        } else {
            for (u64 g = gw; g < (u64) nc * K; g += n_gw) {
                const uint32_t c = (uint32_t) (g / K), j = (uint32_t) (g % K);
                if (c == 0 || j >= a.cnt_in[c - 1]) continue;           // (warp-uniform) nothing new to run here
                sm_round_pair(a, T, W, capable, c, j, lane, n_edges, base_bit);
            }
        }
        STAMP(2);
        sm_grid_barrier(f.bar, gridDim.x);
        STAMP(8);
        // ---- links over the pairs now complete; publish the counts for the next round ----
        for (u64 g = (u64) blockIdx.x * blockDim.x + threadIdx.x; g < (u64) nc * K; g += (u64) gridDim.x * blockDim.x) {
            const uint32_t c = (uint32_t) (g / K), i = (uint32_t) (g % K);
            sm_link_pair(a, f.cnt_alloc, c, i);
            if (i == 0) f.cnt_done[c] = min(f.cnt_alloc[c], K);
        }
        STAMP(9);
        sm_grid_barrier(f.bar, gridDim.x);
        STAMP(3);
        if (blockIdx.x == 0) {
            sm_walk_cta<1024>(a);
            __syncthreads();
            if (a.walk_status[1]) sm_scan_cta(a, f.offsets, f.msgs_base, f.n_msgs_out);   // (written by this CTA's thread 0)
        }
        STAMP(4);
        sm_grid_barrier(f.bar, gridDim.x);
        STAMP(5);
        round++;
        complete = ld_acquire_u32(a.walk_status + 1);
        if (complete || ld_acquire_u32(a.overflow) != 0 || round >= f.max_rounds) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *f.rounds_out = round;
    if (complete) {
        // (the walk's CTA has already scanned the counts, in front of the last barrier)
        for (uint32_t c = gw; c < nc; c += n_gw) sm_gather_chunk(a, f.offsets, f.msgs_out, f.msgs_base, f.msgs_cap, c, lane);
    }
    STAMP(7);
}

// ---------------------------------------------------------------------------------------
// Fallback: Jacobi relaxation, one exit per chunk (thread = chunk).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) sm_round_kernel(const SmArgs a)
{
    __shared__ SmTable T;
    load_table(T, a.tab);

    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_chunks) return;

    i64 start, end;
    chunk_bounds_fixed(a, c, start, end);
    const u64 e = edge_lower_bound(a.edges, a.n_edges, (u64) start);
    const uint32_t tb = a.base_bit ^ (uint32_t) (e & 1);

    SmCarry s;
    if (c == 0) {
        s = a.entry0;
    } else if (a.round == 0) {
        carry_reset(s, tb);
    } else {
        s = a.exit_prev[c - 1];
    }
    if (a.round > 0 && carry_equal(s, a.ran_with[c])) {
        a.exit_cur[c] = a.exit_prev[c];
        return;
    }
    a.ran_with[c] = s;
    atomicAdd(&a.n_ran[a.counter_idx], 1u);

    SpanOut o;
    o.slots = a.slots + (u64) c * a.slot_cap;
    o.cap = a.slot_cap;
    o.n_msgs = 0;
    o.overflow = a.overflow;
    sm_run_span<false>(a, a.n_edges, T, s, start, end, e, tb, o, start);

    a.exit_cur[c] = s;
    a.slot_count[c] = (o.n_msgs < o.cap) ? o.n_msgs : o.cap;
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
