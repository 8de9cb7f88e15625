/*
 * ookd_rx.c -- RX / TX drivers of the host front end.
 *
 * ookd_rx() is the drop-in for ookiedokie_rx() (reference src/ookiedokie.c:222-299) with the
 * bladeRF SC16Q11 file source (src/sdr/bladeRF_file.c): instead of converting, filtering,
 * thresholding and stepping the state machine buffer by buffer on the CPU, it hands raw int16
 * windows of the capture to libookd_gpu and prints what comes back, grouped per
 * samples_per_buffer buffer exactly as the reference's one-rx_print-per-buffer loop does.
 */
#include "ookd_host.h"

#include <errno.h>
#include <inttypes.h>
#include <stdlib.h>
#include <string.h>

#define log_error(...) ookd_log(OOKD_LOG_ERROR, __VA_ARGS__)
#define log_warning(...) ookd_log(OOKD_LOG_WARNING, __VA_ARGS__)

void ookd_cfg_init(struct ookd_cfg *c)      /* src/ookiedokie_cfg.c:27-38, :40-80 */
{
    memset(c, 0, sizeof(*c));
    c->rx_fmt = OOKD_RX_FMT_PRETTY;
    c->rx_threshold = 0.1f;
    c->samplerate = 3000000;
    c->samples_per_buffer = 8192;
    c->tx_count = 1;
    c->tx_delay_us = 4000;
    c->gpu_id = -1;
}

void ookd_rx_print(FILE *out, enum ookd_rx_fmt fmt, bool *first_print, const struct ookd_keyval_list *kv)
{
    const size_t len = kv->n;
    switch (fmt) {
        case OOKD_RX_FMT_CSV:
            if (*first_print) {
                for (size_t i = 0; i < len; i++) {
                    fprintf(out, "%s%c", kv->items[i].key, (i < len - 1) ? ',' : '\n');
                }
                *first_print = false;
            }
            for (size_t i = 0; i < len; i++) {
                fprintf(out, "%s%c", kv->items[i].value, (i < len - 1) ? ',' : '\n');
            }
            break;
        case OOKD_RX_FMT_PRETTY:
            for (size_t i = 0; i < len; i++) {
                fprintf(out, "%20s : %s\n", kv->items[i].key, kv->items[i].value);
            }
            fputc('\n', out);
            break;
    }
}

static uint64_t gcd64(uint64_t a, uint64_t b)
{
    while (b) {
        const uint64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

int ookd_rx(const struct ookd_cfg *cfg)
{
    int status = -1;
    struct ookd_fir *fir = NULL;
    struct ookd_device *dev = NULL;
    ookd_gpu *gpu = NULL;
    FILE *in = NULL, *dig = NULL;
    int16_t *buf = NULL;
    FILE *out = cfg->out ? cfg->out : stdout;
    struct ookd_keyval_list kv;
    ookd_keyval_list_init(&kv);

    if (!cfg->sdr_args) {
        log_error("No capture file given (--sdr-args).\n");
        goto out;
    }
    /* filter selection, src/main.c:642-668: explicit name, "none", or the SDR default */
    if (cfg->rx_filter && !strcasecmp(cfg->rx_filter, "none")) {
        fir = NULL;
    } else if (cfg->rx_filter) {
        fir = ookd_fir_init(cfg->rx_filter);
        if (!fir) {
            goto out;
        }
    } else {
        fir = ookd_fir_init("fs128_fs16_dec4");         /* src/sdr/supported_devices.h:65 */
        if (!fir) {
            log_warning("No default filter found for bladerf_file. No filter is being used.\n");
        }
    }
    const unsigned int decimation = fir ? ookd_fir_get_total_decimation(fir) : 1;
    if (cfg->device) {
        dev = ookd_device_init(cfg->device, cfg->samplerate / decimation);    /* src/main.c:674-683 */
        if (!dev) {
            goto out;
        }
    }
    if (cfg->rx_rec_dig) {
        dig = fopen(cfg->rx_rec_dig, "w");
        if (!dig) {
            log_error("Failed to open %s: %s\n", cfg->rx_rec_dig, strerror(errno));
            goto out;
        }
    }
    in = fopen(cfg->sdr_args, "rb");
    if (!in) {
        log_error("Failed to open %s: %s\n", cfg->sdr_args, strerror(errno));
        goto out;
    }

    struct ookd_gpu_config gc;
    memset(&gc, 0, sizeof(gc));
    gc.filter = fir ? ookd_fir_desc(fir) : NULL;
    gc.sm = dev ? ookd_device_sm_desc(dev) : NULL;
    gc.threshold = cfg->rx_threshold;
    gc.samples_per_buffer = cfg->samples_per_buffer;
    gc.device_id = cfg->gpu_id;
    int rc = ookd_gpu_create(&gpu, &gc);
    if (rc != OOKD_OK) {
        log_error("Failed to set up the GPU receive path: %s\n", ookd_gpu_strerror(rc));
        goto out;
    }

    /* window = whole buffers, aligned for sharding; ~1 GiB of samples at a time */
    const uint64_t spb = cfg->samples_per_buffer;
    const uint64_t align = spb / gcd64(spb, decimation) * decimation;
    uint64_t window = (1ull << 28) / align * align;
    if (window == 0) {
        window = align;
    }
    const uint32_t halo = ookd_gpu_halo(gpu);
    buf = ookd_gpu_host_alloc(((size_t) window + halo) * 2 * sizeof(int16_t));
    if (!buf) {
        log_error("Failed to allocate the pinned staging buffer.\n");
        goto out;
    }

    struct ookd_sm_carry carry;
    ookd_gpu_initial_carry(gpu, &carry);
    bool first_print = true, have_first_bit = false;
    uint32_t cur_bit = 0;
    uint64_t first_sample = 0;
    const size_t nbytes = dev ? (ookd_device_num_bits(dev) + 7) / 8 : 0;
    (void) nbytes;

    /* read one window ahead so that the last window is known to be the last */
    size_t have = fread(buf + 2 * (size_t) halo, 4, window, in);
    while (have > 0) {
        int16_t *next = NULL;
        size_t next_have = 0;
        bool last = (have < window);
        if (!last) {
            next = malloc((size_t) window * 4);
            if (!next) {
                log_error("Out of memory.\n");
                goto out;
            }
            next_have = fread(next, 4, window, in);
            if (next_have == 0) {
                last = true;
            }
        }
        const uint64_t halo_avail = first_sample < halo ? first_sample : halo;
        struct ookd_gpu_result res;
        struct ookd_sm_carry exit_carry;
        rc = ookd_gpu_decode_shard(gpu, buf + 2 * ((size_t) halo - halo_avail), 0, first_sample, have, last,
                                   first_sample ? &carry : NULL, &exit_carry, &res);
        if (rc != OOKD_OK) {
            log_error("GPU decode failed: %s (%s)\n", ookd_gpu_strerror(rc), ookd_gpu_last_error(gpu));
            free(next);
            goto out;
        }
        carry = exit_carry;

        if (dig) {                                          /* record_dig, src/ookiedokie.c:146-169 */
            const uint64_t *edges;
            uint64_t n_edges;
            uint32_t fb;
            rc = ookd_gpu_edges(gpu, &edges, &n_edges, &fb);
            if (rc != OOKD_OK) {
                free(next);
                goto out;
            }
            if (!have_first_bit && res.n_out > 0) {
                cur_bit = fb;
                have_first_bit = true;
                fprintf(dig, "0, %c\n", fb ? '1' : '0');
            }
            for (uint64_t i = 0; i < n_edges; i++) {
                fprintf(dig, "%" PRIu64 ", %c\n%" PRIu64 ", %c\n", edges[i] - 1, cur_bit ? '1' : '0', edges[i],
                        cur_bit ? '0' : '1');
                cur_bit ^= 1;
            }
        }
        if (dev) {                                          /* one rx_print per buffer with messages */
            uint64_t i = 0;
            while (i < res.n_msgs) {
                const uint64_t b = res.msgs[i].buffer_idx;
                ookd_keyval_list_clear(&kv);
                for (; i < res.n_msgs && res.msgs[i].buffer_idx == b; i++) {
                    ookd_device_format(dev, res.msgs[i].data, &kv);
                }
                if (kv.n) {
                    ookd_rx_print(out, cfg->rx_fmt, &first_print, &kv);
                }
            }
        }
        if (last) {
            free(next);
            break;
        }
        /* slide: keep the last `halo` samples in front of the next window */
        memmove(buf, buf + 2 * (size_t) have, (size_t) halo * 4);
        memcpy(buf + 2 * (size_t) halo, next, next_have * 4);
        free(next);
        first_sample += have;
        have = next_have;
    }
    fflush(out);
    status = 0;

out:
    if (buf) ookd_gpu_host_free(buf);
    if (in) fclose(in);
    if (dig) fclose(dig);
    ookd_keyval_list_deinit(&kv);
    ookd_gpu_destroy(gpu);
    ookd_device_deinit(dev);
    ookd_fir_deinit(fir);
    return status;
}

/* ookiedokie_tx with the SC16Q11 file sink: src/ookiedokie.c:301-344, src/sdr/bladeRF_file.c:128-155,
 * complexf_to_sc16q11 (src/complexf.h:87-96: truncating cast, 0.95f * 2048.0f -> 1945). */
int ookd_tx(const struct ookd_cfg *cfg)
{
    int status = -1;
    struct ookd_device *dev = NULL;
    uint32_t *runs = NULL;
    size_t n_runs = 0;
    FILE *f = NULL;
    int16_t *chunk = NULL;

    if (!cfg->device || !cfg->sdr_args) {
        log_error("Error: A target device and an output file must be specified.\n");
        return -1;
    }
    dev = ookd_device_init(cfg->device, cfg->samplerate);
    if (!dev) {
        return -1;
    }
    uint8_t data[OOKD_MSG_BYTES];
    if (!ookd_device_message(dev, cfg->device_params, data) || !ookd_device_generate_runs(dev, data, &runs, &n_runs)) {
        goto out;
    }
    f = fopen(cfg->sdr_args, "wb");
    if (!f) {
        log_error("Failed to open %s: %s\n", cfg->sdr_args, strerror(errno));
        goto out;
    }
    const size_t CH = 65536;
    chunk = malloc(CH * 4);
    if (!chunk) {
        goto out;
    }
    const unsigned int delay = (unsigned int) ((uint64_t) cfg->samplerate * cfg->tx_delay_us / 1000000);
    const int16_t on = (int16_t) (0.95f * 2048.0f);
    for (unsigned int c = 0; c < cfg->tx_count; c++) {
        for (size_t r = 0; r <= n_runs; r++) {
            uint64_t count = (r == 0) ? delay : runs[2 * (r - 1) + 1];
            const int16_t v = (r == 0) ? 0 : (runs[2 * (r - 1)] ? on : 0);
            while (count) {
                const size_t n = count < CH ? (size_t) count : CH;
                for (size_t i = 0; i < n; i++) {
                    chunk[2 * i] = v;
                    chunk[2 * i + 1] = 0;
                }
                if (fwrite(chunk, 4, n, f) != n) {
                    log_error("Write failed: %s\n", strerror(errno));
                    goto out;
                }
                count -= n;
            }
        }
    }
    status = 0;

out:
    free(chunk);
    if (f) fclose(f);
    free(runs);
    ookd_device_deinit(dev);
    return status;
}
