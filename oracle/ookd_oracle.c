/*
 * TEST INFRASTRUCTURE ONLY -- see ookd_oracle.h.
 *
 * CPU restatement of the OOKiedokie receive path used as the parity oracle
 * for the B200 kernels.  Pinned against the unmodified reference built into
 * oracle/_ref (tools/make_golden.py -> tests/golden/, tests/test_oracle_*.py).
 *
 * Build: gcc -O2 -ffp-contract=off (never -ffast-math / -march=native: the
 * reference's FIR is an in-order fp32 multiply-then-add chain and fusing or
 * reassociating it changes the bits that parity is judged on).
 */
#include "ookd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* SC16Q11 -> complex float            reference: src/complexf.h:68-77 */
/* ------------------------------------------------------------------ */
void ookd_oracle_sc16q11_to_cf(const int16_t *in, float *out, size_t n)
{
    const float scale = 1.0f / 2048.0f;
    for (size_t k = 0; k < 2 * n; k++) {
        out[k] = (float) in[k] * scale;
    }
}

/* ------------------------------------------------------------------ */
/* Multi-stage decimating FIR          reference: src/fir.c            */
/*   stage state (taps, countdown, history)      fir.c:39-57           */
/*   reset: zero history, countdown = decimation fir.c:272-295         */
/*   per input: push, countdown--, at 0 convolve fir.c:302-353         */
/*   stage chaining                              fir.c:355-395         */
/* The reference keeps a 2T delay line with two insertion points; the  */
/* arithmetic it performs is  out = sum_{i=0..T-1} taps[i]*x[n-i],     */
/* accumulated from 0 in index order, real and imaginary separately,   */
/* one rounded multiply then one rounded add per tap (fir.c:313-318).  */
/* Here the history is a plain ring; the arithmetic is the same.       */
/* ------------------------------------------------------------------ */
struct oo_stage {
    uint32_t decimation;
    uint32_t num_taps;
    float   *taps;
    float   *hist;      /* ring of num_taps complex samples */
    uint32_t head;      /* slot of the newest sample */
    uint32_t countdown;
    float   *out;       /* inter-stage scratch, grown on demand */
    size_t   out_cap;
};

struct ookd_oracle_fir {
    uint32_t n_stages;
    uint32_t total_decimation;
    struct oo_stage st[OOKD_ORACLE_MAX_STAGES];
};

ookd_oracle_fir *ookd_oracle_fir_create(uint32_t n_stages, const uint32_t *decimation,
                                        const uint32_t *num_taps, const float *taps)
{
    if (n_stages == 0 || n_stages > OOKD_ORACLE_MAX_STAGES) {
        return NULL;
    }
    ookd_oracle_fir *f = calloc(1, sizeof(*f));
    if (!f) {
        return NULL;
    }
    f->n_stages = n_stages;
    f->total_decimation = 1;
    for (uint32_t s = 0; s < n_stages; s++) {
        struct oo_stage *st = &f->st[s];
        if (decimation[s] == 0 || num_taps[s] == 0) {
            ookd_oracle_fir_destroy(f);
            return NULL;
        }
        st->decimation = decimation[s];
        st->num_taps = num_taps[s];
        st->taps = malloc(sizeof(float) * num_taps[s]);
        st->hist = malloc(sizeof(float) * 2 * num_taps[s]);
        memcpy(st->taps, taps, sizeof(float) * num_taps[s]);
        taps += num_taps[s];
        f->total_decimation *= decimation[s];
    }
    ookd_oracle_fir_reset(f);
    return f;
}

void ookd_oracle_fir_reset(ookd_oracle_fir *f)
{
    for (uint32_t s = 0; s < f->n_stages; s++) {
        struct oo_stage *st = &f->st[s];
        memset(st->hist, 0, sizeof(float) * 2 * st->num_taps);
        st->head = 0;
        st->countdown = st->decimation;
    }
}

uint32_t ookd_oracle_fir_total_decimation(const ookd_oracle_fir *f)
{
    return f->total_decimation;
}

void ookd_oracle_fir_destroy(ookd_oracle_fir *f)
{
    if (!f) {
        return;
    }
    for (uint32_t s = 0; s < f->n_stages; s++) {
        free(f->st[s].taps);
        free(f->st[s].hist);
        free(f->st[s].out);
    }
    free(f);
}

static size_t oo_stage_run(struct oo_stage *st, const float *in, size_t n, float *out)
{
    const uint32_t T = st->num_taps;
    size_t n_out = 0;

    for (size_t k = 0; k < n; k++) {
        st->head = (st->head + 1 == T) ? 0 : st->head + 1;
        st->hist[2 * st->head]     = in[2 * k];
        st->hist[2 * st->head + 1] = in[2 * k + 1];

        if (--st->countdown == 0) {
            float re = 0.0f, im = 0.0f;
            uint32_t pos = st->head;            /* newest sample first */
            for (uint32_t i = 0; i < T; i++) {
                const float pr = st->taps[i] * st->hist[2 * pos];
                const float pi = st->taps[i] * st->hist[2 * pos + 1];
                re = re + pr;
                im = im + pi;
                pos = (pos == 0) ? T - 1 : pos - 1;
            }
            out[2 * n_out]     = re;
            out[2 * n_out + 1] = im;
            n_out++;
            st->countdown = st->decimation;
        }
    }
    return n_out;
}

size_t ookd_oracle_fir_run(ookd_oracle_fir *f, const float *in_iq, size_t n, float *out_iq)
{
    const float *src = in_iq;
    size_t count = n;

    for (uint32_t s = 0; s < f->n_stages; s++) {
        struct oo_stage *st = &f->st[s];
        float *dst;
        if (s == f->n_stages - 1) {
            dst = out_iq;
        } else {
            const size_t need = count / st->decimation + 2;
            if (st->out_cap < need) {
                free(st->out);
                st->out = malloc(sizeof(float) * 2 * need);
                st->out_cap = need;
            }
            dst = st->out;
        }
        count = oo_stage_run(st, src, count, dst);
        src = dst;
    }
    return count;
}

/* ------------------------------------------------------------------ */
/* Envelope threshold   reference: src/ookiedokie.c:171-179,           */
/*                      src/complexf.h:43-58 (power, then sqrtf)       */
/* ------------------------------------------------------------------ */
void ookd_oracle_threshold(const float *iq, size_t n, float thr, uint8_t *bits)
{
    for (size_t k = 0; k < n; k++) {
        const float re = iq[2 * k], im = iq[2 * k + 1];
        const float rr = re * re;
        const float ii = im * im;
        const float p = rr + ii;
        bits[k] = sqrtf(p) >= thr;
    }
}

/* ------------------------------------------------------------------ */
/* Device state machine, RX half   reference: src/state_machine.c      */
/* ------------------------------------------------------------------ */
struct oo_trigger {
    int32_t  cond;
    uint64_t duration_us;
    int32_t  action;
    uint32_t next;
};

struct oo_state {
    uint64_t duration_us;
    uint64_t timeout_us;
    uint32_t first_trigger;
    uint32_t num_triggers;
};

struct ookd_oracle_sm {
    uint32_t num_states;
    struct oo_state *states;
    struct oo_trigger *triggers;
    uint32_t curr;                      /* state index; 0 is RESET (state_machine.c:52) */
    uint8_t  data[OOKD_ORACLE_MSG_BYTES + 1];
    uint32_t max_bits;
    uint32_t num_bits;
    int      prev_bit;
    double   elapsed_us;
    uint32_t k_count;                   /* additions since the last fire (mirror of elapsed_us) */
    uint32_t sample_rate;
};

#define OO_TOLERANCE 0.15               /* state_machine.c:55 */

ookd_oracle_sm *ookd_oracle_sm_create(uint32_t num_states,
                                      const uint64_t *state_duration_us,
                                      const uint64_t *state_timeout_us,
                                      const uint32_t *trig_off,
                                      const int32_t *trig_cond,
                                      const uint64_t *trig_duration_us,
                                      const int32_t *trig_action,
                                      const uint32_t *trig_next,
                                      uint32_t max_bits, uint32_t sample_rate)
{
    if (num_states == 0 || max_bits == 0 || max_bits > 8 * OOKD_ORACLE_MSG_BYTES) {
        return NULL;
    }
    ookd_oracle_sm *sm = calloc(1, sizeof(*sm));
    const uint32_t nt = trig_off[num_states];
    sm->num_states = num_states;
    sm->states = calloc(num_states, sizeof(sm->states[0]));
    sm->triggers = calloc(nt ? nt : 1, sizeof(sm->triggers[0]));
    for (uint32_t s = 0; s < num_states; s++) {
        sm->states[s].duration_us = state_duration_us[s];
        sm->states[s].timeout_us = state_timeout_us[s];
        sm->states[s].first_trigger = trig_off[s];
        sm->states[s].num_triggers = trig_off[s + 1] - trig_off[s];
    }
    for (uint32_t t = 0; t < nt; t++) {
        sm->triggers[t].cond = trig_cond[t];
        sm->triggers[t].duration_us = trig_duration_us[t];
        sm->triggers[t].action = trig_action[t];
        sm->triggers[t].next = trig_next[t];
    }
    sm->max_bits = max_bits;
    sm->sample_rate = sample_rate;
    return sm;
}

void ookd_oracle_sm_destroy(ookd_oracle_sm *sm)
{
    if (sm) {
        free(sm->states);
        free(sm->triggers);
        free(sm);
    }
}

const uint8_t *ookd_oracle_sm_data(const ookd_oracle_sm *sm) { return sm->data; }

/* Snapshot / restore of the streaming state (state_machine.c:57-75), for tests that cut a capture
 * into shards.  elapsed_us is restored by replaying the k additions that produced it. */
void ookd_oracle_sm_get_state(const ookd_oracle_sm *sm, uint32_t *state, uint32_t *k, uint32_t *num_bits,
                              uint32_t *prev_bit, uint8_t *data32)
{
    *state = sm->curr; *k = sm->k_count; *num_bits = sm->num_bits; *prev_bit = (uint32_t) sm->prev_bit;
    memcpy(data32, sm->data, OOKD_ORACLE_MSG_BYTES);
}

void ookd_oracle_sm_set_state(ookd_oracle_sm *sm, uint32_t state, uint32_t k, uint32_t num_bits,
                              uint32_t prev_bit, const uint8_t *data32)
{
    sm->curr = state; sm->num_bits = num_bits; sm->prev_bit = (int) prev_bit;
    memcpy(sm->data, data32, OOKD_ORACLE_MSG_BYTES);
    sm->elapsed_us = 0.0;
    for (uint32_t i = 0; i < k; i++) {
        sm->elapsed_us += ((double) 1 / (double) sm->sample_rate) * 1e6;
    }
    sm->k_count = k;
}
uint32_t ookd_oracle_sm_num_bits(const ookd_oracle_sm *sm) { return sm->num_bits; }

/* state_machine.c:100-133: window is [d-0.15d, d+0.15d] evaluated in double,
 * narrowed to float, then compared against the double elapsed time. 0 = any. */
static int oo_in_window(uint64_t d_us, double elapsed)
{
    if (d_us == 0) {
        return 1;
    }
    const float lo = (float) ((double) d_us - (OO_TOLERANCE * (double) d_us));
    const float hi = (float) ((double) d_us + (OO_TOLERANCE * (double) d_us));
    return elapsed >= lo && elapsed <= hi;
}

/* One trigger evaluation: state_machine.c:421-519 (+ actions :365-419). */
static int oo_eval(ookd_oracle_sm *sm, int b)
{
    const struct oo_state *st = &sm->states[sm->curr];
    const struct oo_trigger *fired = NULL;
    int check_state_duration = 0;
    int result = 0;

    for (uint32_t i = 0; i < st->num_triggers && !fired; i++) {
        const struct oo_trigger *t = &sm->triggers[st->first_trigger + i];
        if (!oo_in_window(t->duration_us, sm->elapsed_us)) {
            continue;
        }
        switch (t->cond) {
            case OO_COND_ALWAYS:
                fired = t;
                break;
            case OO_COND_PULSE_START:
                if (!sm->prev_bit && b) { fired = t; check_state_duration = 1; }
                break;
            case OO_COND_PULSE_END:
                if (sm->prev_bit && !b) { fired = t; check_state_duration = 1; }
                break;
            case OO_COND_TIMEOUT:
                if (st->timeout_us != 0 && sm->elapsed_us >= (double) st->timeout_us) {
                    fired = t;
                }
                break;
            case OO_COND_MSG_COMPLETE:
                if (sm->num_bits >= sm->max_bits) { fired = t; }
                break;
            default:
                return -1;
        }
    }

    if (!fired) {
        sm->elapsed_us += ((double) 1 / (double) sm->sample_rate) * 1e6;   /* :78-82, :514 */
        sm->k_count++;
        return 0;
    }

    if (!check_state_duration || oo_in_window(st->duration_us, sm->elapsed_us)) {
        switch (fired->action) {
            case OO_ACT_NONE:
                break;
            case OO_ACT_APPEND_0:
            case OO_ACT_APPEND_1:
                /* :365-385: LSB-first within bytes, guarded by num_bits <= max_bits */
                if (sm->num_bits <= sm->max_bits) {
                    const uint32_t byte = sm->num_bits / 8, bit = sm->num_bits % 8;
                    if (fired->action == OO_ACT_APPEND_1) {
                        sm->data[byte] |= (uint8_t) (1u << bit);
                    } else {
                        sm->data[byte] &= (uint8_t) ~(1u << bit);
                    }
                }
                sm->num_bits++;
                break;
            case OO_ACT_OUTPUT_DATA:
                result = 1;
                break;
            default:
                result = -1;
        }
        if (result != -1) {
            sm->curr = fired->next;
        }
    } else {
        result = -1;
    }

    if (result == -1) {
        sm->curr = 0;
    }
    sm->elapsed_us = 0;
    sm->k_count = 0;
    return result;
}

/* state_machine.c:521-539: RESET clears the message and is evaluated, then the
 * (possibly new) state is evaluated again on the same sample. */
static int oo_step(ookd_oracle_sm *sm, int b)
{
    if (sm->curr == 0) {
        sm->num_bits = 0;
        memset(sm->data, 0, (sm->max_bits + 7) / 8);
        const int r = oo_eval(sm, b);
        if (r != 0) {
            return r;
        }
    }
    return oo_eval(sm, b);
}

int ookd_oracle_sm_process(ookd_oracle_sm *sm, const uint8_t *bits, uint32_t count,
                           uint32_t *num_proc)
{
    uint32_t i;
    int result = 0;
    for (i = 0; i < count && result == 0; i++) {
        result = oo_step(sm, bits[i] != 0);
        sm->prev_bit = bits[i] != 0;        /* :551, also after an error */
    }
    *num_proc = i;
    return result;
}

/* ------------------------------------------------------------------ */
/* Full receive loop  reference: src/ookiedokie.c:238-290,             */
/*   src/sdr/bladeRF_file.c:97-126 (short read => zero pad, 0 => EOF), */
/*   src/ookiedokie.c:146-169 (edge recorder), src/device.c:634-658    */
/*   (per-buffer driver: an ERROR abandons the rest of the buffer).    */
/* ------------------------------------------------------------------ */
struct oo_vec {
    void  *p;
    size_t n, cap, elem;
};

static void *oo_push(struct oo_vec *v)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 256;
        v->p = realloc(v->p, v->cap * v->elem);
    }
    return (char *) v->p + (v->n++) * v->elem;
}

int ookd_oracle_rx(const int16_t *iq, uint64_t n_samples, ookd_oracle_fir *fir,
                   ookd_oracle_sm *sm, float threshold, uint32_t spb,
                   int want_filtered, int want_bits, ookd_oracle_rx_result *res)
{
    memset(res, 0, sizeof(*res));
    if (spb == 0) {
        return -1;
    }
    const uint64_t n_buffers = (n_samples + spb - 1) / spb;
    const uint32_t dec = fir ? fir->total_decimation : 1;
    const uint64_t max_out = (n_buffers * spb) / dec + 1;

    int16_t *raw = malloc(sizeof(int16_t) * 2 * spb);
    float *cf = malloc(sizeof(float) * 2 * spb);
    float *post = malloc(sizeof(float) * 2 * spb);
    uint8_t *dig = malloc(spb);
    struct oo_vec edges = { NULL, 0, 0, sizeof(uint64_t) };
    struct oo_vec msgs = { NULL, 0, 0, sizeof(ookd_oracle_msg) };

    if (want_filtered) {
        res->filtered = malloc(sizeof(float) * 2 * max_out);
    }
    if (want_bits) {
        res->bits = malloc(max_out);
    }

    uint64_t out_base = 0;
    int dig_prev = 0;

    for (uint64_t b = 0; b < n_buffers; b++) {
        const uint64_t start = b * spb;
        const uint64_t have = (n_samples - start < spb) ? n_samples - start : spb;
        memcpy(raw, iq + 2 * start, sizeof(int16_t) * 2 * have);
        memset(raw + 2 * have, 0, sizeof(int16_t) * 2 * (spb - have));
        ookd_oracle_sc16q11_to_cf(raw, cf, spb);

        const float *to_thr = cf;
        size_t count = spb;
        if (fir) {
            count = ookd_oracle_fir_run(fir, cf, spb, post);
            to_thr = post;
        }
        ookd_oracle_threshold(to_thr, count, threshold, dig);

        if (res->filtered) {
            memcpy(res->filtered + 2 * out_base, to_thr, sizeof(float) * 2 * count);
        }
        if (res->bits) {
            memcpy(res->bits + out_base, dig, count);
        }

        /* ookiedokie.c:146-169 */
        if (out_base == 0 && count > 0) {
            dig_prev = dig[0];
            res->first_bit = dig[0];
        }
        for (size_t i = 0; i < count; i++) {
            if (dig[i] != dig_prev) {
                *(uint64_t *) oo_push(&edges) = out_base + i;
                dig_prev = dig[i];
            }
        }

        /* device.c:634-658 */
        if (sm) {
            uint32_t total = 0, np = 0;
            int r = 0;
            while (total < count && r != -1) {
                r = ookd_oracle_sm_process(sm, dig + total, (uint32_t) count - total, &np);
                total += np;
                if (r == 1) {
                    ookd_oracle_msg *m = oo_push(&msgs);
                    memset(m, 0, sizeof(*m));
                    m->out_sample = out_base + total - 1;
                    m->buffer_idx = b;
                    m->num_bits = sm->num_bits;
                    memcpy(m->data, sm->data, OOKD_ORACLE_MSG_BYTES);
                }
            }
        }
        out_base += count;
    }

    res->n_out = out_base;
    res->n_buffers = n_buffers;
    res->n_edges = edges.n;
    res->edges = edges.p;
    res->n_msgs = msgs.n;
    res->msgs = msgs.p;
    free(raw);
    free(cf);
    free(post);
    free(dig);
    return 0;
}

void ookd_oracle_rx_free(ookd_oracle_rx_result *res)
{
    free(res->filtered);
    free(res->bits);
    free(res->edges);
    free(res->msgs);
    memset(res, 0, sizeof(*res));
}

/* ------------------------------------------------------------------ */
/* Synthetic capture (our recipe, integer-only => identical on GPU)    */
/* ------------------------------------------------------------------ */
static inline uint64_t oo_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline int64_t oo_ih4(uint64_t r)
{
    return (int64_t) ((r & 0xFFFF) + ((r >> 16) & 0xFFFF) + ((r >> 32) & 0xFFFF) + ((r >> 48) & 0xFFFF)) - 131070;
}

static inline int32_t oo_noise(uint64_t seed, uint64_t ctr, int32_t scale, uint32_t terms)
{
    const uint64_t base = oo_mix64(seed) ^ (ctr * 0xD1342543DE82EF95ull);
    int64_t s = oo_ih4(oo_mix64(base));
    if (terms == 12) {
        s += oo_ih4(oo_mix64(base ^ (1ull * 0xA24BAED4963EE407ull)));
        s += oo_ih4(oo_mix64(base ^ (2ull * 0xA24BAED4963EE407ull)));
    }
    const int64_t v = s * (int64_t) scale + (1 << 23);
    return (int32_t) (v >> 24);         /* arithmetic shift: floor => round half up */
}

static inline int16_t oo_clip(int32_t v)
{
    return (int16_t) (v < -2048 ? -2048 : (v > 2047 ? 2047 : v));
}

void ookd_oracle_synth(int16_t *iq, uint64_t first_sample, uint64_t n_samples,
                       const uint64_t *toggles, uint64_t n_toggles,
                       int32_t i_on, int32_t q_on, int32_t noise_scale, uint64_t seed)
{
    ookd_oracle_synth_ex(iq, first_sample, n_samples, toggles, n_toggles, i_on, q_on, noise_scale, seed, 4);
}

void ookd_oracle_synth_ex(int16_t *iq, uint64_t first_sample, uint64_t n_samples,
                          const uint64_t *toggles, uint64_t n_toggles,
                          int32_t i_on, int32_t q_on, int32_t noise_scale, uint64_t seed, uint32_t noise_terms)
{
    /* k = number of toggles <= n */
    uint64_t lo = 0, hi = n_toggles;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        if (toggles[mid] <= first_sample) { lo = mid + 1; } else { hi = mid; }
    }
    uint64_t k = lo;
    for (uint64_t j = 0; j < n_samples; j++) {
        const uint64_t n = first_sample + j;
        while (k < n_toggles && toggles[k] <= n) {
            k++;
        }
        const int on = (int) (k & 1);
        int32_t vi = on ? i_on : 0, vq = on ? q_on : 0;
        if (noise_scale != 0) {
            vi += oo_noise(seed, 2 * n, noise_scale, noise_terms);
            vq += oo_noise(seed, 2 * n + 1, noise_scale, noise_terms);
        }
        iq[2 * j] = oo_clip(vi);
        iq[2 * j + 1] = oo_clip(vq);
    }
}
