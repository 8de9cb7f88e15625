/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the OOKiedokie receive path.
 *
 * This is the parity oracle for the B200 kernels: a plain-C, single-threaded
 * restatement of the reference algorithm, written from the reference's
 * behaviour (each function cites the file:line it follows).  It is pinned
 * against the unmodified reference compiled into oracle/_ref (see Makefile and
 * tools/make_golden.py; the resulting vectors live in tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libookd_gpu.so,
 * the host front end) never links, loads or calls it.
 */
#ifndef OOKD_ORACLE_H
#define OOKD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OOKD_ORACLE_MAX_STAGES 8
#define OOKD_ORACLE_MSG_BYTES  32

/* Trigger condition / action codes: values of enum sm_trigger_cond and
 * enum sm_trigger_action, reference src/state_machine.h:33-50. */
enum {
    OO_COND_ALWAYS = 1, OO_COND_PULSE_START, OO_COND_PULSE_END,
    OO_COND_TIMEOUT, OO_COND_MSG_COMPLETE
};
enum {
    OO_ACT_NONE = 1, OO_ACT_APPEND_0, OO_ACT_APPEND_1, OO_ACT_OUTPUT_DATA
};

typedef struct ookd_oracle_fir ookd_oracle_fir;
typedef struct ookd_oracle_sm ookd_oracle_sm;

typedef struct {
    uint64_t out_sample;   /* global post-decimation index of the sample that completed it */
    uint64_t buffer_idx;   /* index of the samples_per_buffer buffer it was printed with */
    uint32_t num_bits;
    uint8_t  data[OOKD_ORACLE_MSG_BYTES];
} ookd_oracle_msg;

typedef struct {
    /* outputs of ookd_oracle_rx; arrays are malloc'd, free with ookd_oracle_rx_free */
    uint64_t n_out;            /* post-decimation samples produced */
    uint64_t n_buffers;
    float   *filtered;         /* 2*n_out floats (NULL unless requested) */
    uint8_t *bits;             /* n_out threshold decisions (NULL unless requested) */
    uint8_t  first_bit;        /* bits[0] (record_dig's "0, b" row) */
    uint64_t n_edges;
    uint64_t *edges;           /* positions i>=1 with bits[i] != bits[i-1] */
    uint64_t n_msgs;
    ookd_oracle_msg *msgs;
} ookd_oracle_rx_result;

/* complexf.h:68-77 */
void ookd_oracle_sc16q11_to_cf(const int16_t *in, float *out, size_t n);

/* fir.c:39-66, 272-295 (state), 302-395 (filtering).  taps = all stages concatenated. */
ookd_oracle_fir *ookd_oracle_fir_create(uint32_t n_stages, const uint32_t *decimation,
                                        const uint32_t *num_taps, const float *taps);
void   ookd_oracle_fir_reset(ookd_oracle_fir *f);
size_t ookd_oracle_fir_run(ookd_oracle_fir *f, const float *in_iq, size_t n, float *out_iq);
uint32_t ookd_oracle_fir_total_decimation(const ookd_oracle_fir *f);
void   ookd_oracle_fir_destroy(ookd_oracle_fir *f);

/* ookiedokie.c:171-179 with complexf.h:43-58 */
void ookd_oracle_threshold(const float *iq, size_t n, float thr, uint8_t *bits);

/* state_machine.c:33-75, 135-330: states/triggers in the microsecond domain.
 * trig_off has num_states+1 entries (CSR offsets into the trigger arrays). */
ookd_oracle_sm *ookd_oracle_sm_create(uint32_t num_states,
                                      const uint64_t *state_duration_us,
                                      const uint64_t *state_timeout_us,
                                      const uint32_t *trig_off,
                                      const int32_t *trig_cond,
                                      const uint64_t *trig_duration_us,
                                      const int32_t *trig_action,
                                      const uint32_t *trig_next,
                                      uint32_t max_bits, uint32_t sample_rate);
void ookd_oracle_sm_destroy(ookd_oracle_sm *sm);
/* state_machine.c:541-556: returns -1 error / 0 no output / 1 output ready */
int ookd_oracle_sm_process(ookd_oracle_sm *sm, const uint8_t *bits, uint32_t count,
                           uint32_t *num_proc);
const uint8_t *ookd_oracle_sm_data(const ookd_oracle_sm *sm);
void ookd_oracle_sm_get_state(const ookd_oracle_sm *sm, uint32_t *state, uint32_t *k, uint32_t *num_bits,
                              uint32_t *prev_bit, uint8_t *data32);
void ookd_oracle_sm_set_state(ookd_oracle_sm *sm, uint32_t state, uint32_t k, uint32_t num_bits,
                              uint32_t prev_bit, const uint8_t *data32);
uint32_t ookd_oracle_sm_num_bits(const ookd_oracle_sm *sm);

/* The whole loop body of ookiedokie_rx (ookiedokie.c:238-290) + bladeRF_file.c:97-126
 * EOF/zero-pad semantics + device_process (device.c:634-658), over a full capture.
 * fir may be NULL (no filter), sm may be NULL (edges only). */
int ookd_oracle_rx(const int16_t *iq, uint64_t n_samples, ookd_oracle_fir *fir,
                   ookd_oracle_sm *sm, float threshold, uint32_t samples_per_buffer,
                   int want_filtered, int want_bits, ookd_oracle_rx_result *res);
void ookd_oracle_rx_free(ookd_oracle_rx_result *res);

/* Deterministic integer-only synthetic capture (NOT from the reference; the
 * recipe is ours and the GPU generator must reproduce it byte for byte).
 * envelope(n) = parity of #{toggles <= n}; sample = clip(env*(i_on,q_on) + noise).
 * noise_scale: Q24 multiplier applied to a centred sum of four 16-bit uniforms. */
void ookd_oracle_synth(int16_t *iq, uint64_t first_sample, uint64_t n_samples,
                       const uint64_t *toggles, uint64_t n_toggles,
                       int32_t i_on, int32_t q_on, int32_t noise_scale, uint64_t seed);
/* noise_terms = 4 (as above) or 12 (three draws: Irwin-Hall(12), close to Gaussian, tails to +-6 sigma) */
void ookd_oracle_synth_ex(int16_t *iq, uint64_t first_sample, uint64_t n_samples,
                          const uint64_t *toggles, uint64_t n_toggles,
                          int32_t i_on, int32_t q_on, int32_t noise_scale, uint64_t seed, uint32_t noise_terms);

#ifdef __cplusplus
}
#endif
#endif
