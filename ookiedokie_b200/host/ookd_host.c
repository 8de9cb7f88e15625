/*
 * ookd_host.c -- loaders, formatter and transmit-side generator of the host front end.
 * See ookd_host.h for the map onto the reference interface.
 */
#include "ookd_host.h"
#include "ookd_json.h"

#include <errno.h>
#include <float.h>
#include <inttypes.h>
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/time.h>
#include <time.h>

#ifndef OOKD_DEFAULT_DATA_DIR
#define OOKD_DEFAULT_DATA_DIR "/usr/local/share/OOKiedokie/"
#endif

/* ------------------------------------------------------------------------- */
/* logging                                                                    */
/* ------------------------------------------------------------------------- */
static enum ookd_log_level g_verbosity = OOKD_LOG_INFO;

void ookd_log_set_verbosity(enum ookd_log_level level)
{
    g_verbosity = level;
}

void ookd_log(enum ookd_log_level level, const char *fmt, ...)
{
    if (level < g_verbosity) {
        return;
    }
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

#define log_error(...)   ookd_log(OOKD_LOG_ERROR, __VA_ARGS__)
#define log_warning(...) ookd_log(OOKD_LOG_WARNING, __VA_ARGS__)
#define log_debug(...)   ookd_log(OOKD_LOG_DEBUG, __VA_ARGS__)

/* ------------------------------------------------------------------------- */
/* search path: same order as src/find.c:49-59, :185-229                      */
/* ------------------------------------------------------------------------- */
const char *ookd_data_dir(void)
{
    const char *env = getenv("OOKD_DATA_DIR");
    return (env && env[0]) ? env : OOKD_DEFAULT_DATA_DIR;
}

static FILE *try_dirs(const char *prefix, const char *name, const char *ext)
{
    const char *home = getenv("HOME");
    char dirs[4][PATH_MAX];
    int nd = 0;

    dirs[nd++][0] = '\0';                                       /* as given / cwd */
    if (home) {
        snprintf(dirs[nd++], PATH_MAX, "%s/.config/OOKiedokie/", home);
        snprintf(dirs[nd++], PATH_MAX, "%s/.OOKiedokie/", home);
    }
    {
        const char *dd = ookd_data_dir();
        const size_t len = strlen(dd);
        snprintf(dirs[nd++], PATH_MAX, "%s%s", dd, (len && dd[len - 1] != '/') ? "/" : "");
    }
    for (int i = 0; i < nd; i++) {
        char path[4 * PATH_MAX + 64];
        snprintf(path, sizeof(path), "%s%s%s%s", dirs[i], prefix ? prefix : "", name, ext ? ext : "");
        log_debug("Searching for: %s\n", path);
        FILE *f = fopen(path, "r");
        if (f) {
            return f;
        }
    }
    return NULL;
}

static FILE *find_json(const char *subdir, const char *name)
{
    FILE *f = try_dirs(NULL, name, NULL);       /* full path incl. extension */
    if (!f) {
        f = try_dirs(NULL, name, ".json");      /* path without extension */
    }
    if (!f) {
        f = try_dirs(subdir, name, ".json");    /* by name within the search path */
    }
    return f;
}

FILE *ookd_find_device_file(const char *name) { return find_json("devices/", name); }
FILE *ookd_find_filter_file(const char *name) { return find_json("filters/", name); }

/* ------------------------------------------------------------------------- */
/* key/value list                                                             */
/* ------------------------------------------------------------------------- */
void ookd_keyval_list_init(struct ookd_keyval_list *l)
{
    memset(l, 0, sizeof(*l));
}

bool ookd_keyval_list_append(struct ookd_keyval_list *l, const char *key, const char *value)
{
    if (l->n == l->cap) {
        const size_t nc = l->cap ? 2 * l->cap : 16;
        void *tmp = realloc(l->items, nc * sizeof(l->items[0]));
        if (!tmp) {
            return false;
        }
        l->items = tmp;
        l->cap = nc;
    }
    l->items[l->n].key = strdup(key);
    l->items[l->n].value = strdup(value);
    if (!l->items[l->n].key || !l->items[l->n].value) {
        free(l->items[l->n].key);
        free(l->items[l->n].value);
        return false;
    }
    l->n++;
    return true;
}

void ookd_keyval_list_clear(struct ookd_keyval_list *l)
{
    for (size_t i = 0; i < l->n; i++) {
        free(l->items[i].key);
        free(l->items[i].value);
    }
    l->n = 0;
}

void ookd_keyval_list_deinit(struct ookd_keyval_list *l)
{
    ookd_keyval_list_clear(l);
    free(l->items);
    memset(l, 0, sizeof(*l));
}

/* ------------------------------------------------------------------------- */
/* filter loader: what fir_init reads, src/fir.c:68-249                       */
/* ------------------------------------------------------------------------- */
struct ookd_fir {
    struct ookd_filter_desc desc;
    float *taps[OOKD_MAX_STAGES];
    unsigned int total_decimation;
};

void ookd_fir_deinit(struct ookd_fir *f)
{
    if (f) {
        for (int i = 0; i < OOKD_MAX_STAGES; i++) {
            free(f->taps[i]);
        }
        free(f);
    }
}

struct ookd_fir *ookd_fir_init(const char *filter_name)
{
    char err[200];
    struct ookd_fir *fir = NULL;
    struct oj_value *root = NULL;
    int status = -1;

    FILE *in = ookd_find_filter_file(filter_name);
    if (!in) {
        log_error("Unable to find filter file: %s\n", filter_name);
        return NULL;
    }
    root = oj_parse_file(in, err, sizeof(err));
    fclose(in);
    if (!root) {
        log_error("Error in %s (%s)\n", filter_name, err);
        return NULL;
    }
    const struct oj_value *filt = oj_get(root, "filter");
    if (!filt) {
        log_error("Error: Failed to find \"filter\" entry in filter file.\n");
        goto out;
    }
    const struct oj_value *stages = oj_get(filt, "stages");
    if (!stages) {
        log_error("Error: Failed to find \"stages\" entry in filter file.\n");
        goto out;
    }
    if (!oj_is_array(stages)) {
        log_error("Error: \"filter\" entry in filter file is not an array.\n");
        goto out;
    }
    if (stages->n == 0) {
        log_error("Error: Filter must have 1 or more stages.\n");
        goto out;
    }
    if (stages->n > OOKD_MAX_STAGES) {
        log_error("Error: at most %d filter stages are supported.\n", OOKD_MAX_STAGES);
        goto out;
    }
    fir = calloc(1, sizeof(*fir));
    if (!fir) {
        goto out;
    }
    fir->desc.num_stages = (uint32_t) stages->n;
    fir->total_decimation = 1;
    for (size_t i = 0; i < stages->n; i++) {
        const struct oj_value *stage = stages->items[i];
        const struct oj_value *dec = oj_get(stage, "decimation");
        if (dec) {
            if (!oj_is_int(dec)) {
                log_error("Error: Decimation must be an integer.\n");
                goto out;
            }
            if (dec->i <= 0 || dec->i >= (long long) UINT_MAX) {
                log_error("Error: Decimation value is outside of allowed range.\n");
                goto out;
            }
            fir->desc.decimation[i] = (uint32_t) dec->i;
        } else {
            fir->desc.decimation[i] = 1;
        }
        if ((uint64_t) fir->total_decimation * fir->desc.decimation[i] > UINT_MAX) {
            log_error("Error: total decimation is too large.\n");
            goto out;
        }
        fir->total_decimation *= fir->desc.decimation[i];

        const struct oj_value *taps = oj_get(stage, "taps");
        if (!taps) {
            log_error("Error: Filter stage is missing \"taps\" entry.\n");
            goto out;
        }
        if (!oj_is_array(taps)) {
            log_error("Error: Filter \"taps\" must be an array.\n");
            goto out;
        }
        if (taps->n == 0) {
            log_error("Error: Filter stage %zd must have 1 or more taps.\n", i + 1);
            goto out;
        }
        if (taps->n > OOKD_MAX_TAPS) {
            log_error("Error: Filter stage %zd has more than %d taps.\n", i + 1, OOKD_MAX_TAPS);
            goto out;
        }
        fir->taps[i] = malloc(sizeof(float) * taps->n);
        if (!fir->taps[i]) {
            goto out;
        }
        for (size_t t = 0; t < taps->n; t++) {
            if (!oj_is_number(taps->items[t])) {
                log_error("Error: tap %zd in stage %zd is an invalid value.\n", t + 1, i + 1);
                goto out;
            }
            fir->taps[i][t] = (float) taps->items[t]->d;        /* src/fir.c:224 */
        }
        fir->desc.num_taps[i] = (uint32_t) taps->n;
        fir->desc.taps[i] = fir->taps[i];
    }
    status = 0;

out:
    oj_free(root);
    if (status != 0) {
        ookd_fir_deinit(fir);
        fir = NULL;
    }
    return fir;
}

unsigned int ookd_fir_get_total_decimation(const struct ookd_fir *f) { return f->total_decimation; }
const struct ookd_filter_desc *ookd_fir_desc(const struct ookd_fir *f) { return &f->desc; }

/* ------------------------------------------------------------------------- */
/* device: state machine description + fields                                 */
/* ------------------------------------------------------------------------- */
enum field_fmt { FMT_HEX, FMT_UNSIGNED_DEC, FMT_SIGN_MAGNITUDE, FMT_TWOS_COMPLEMENT, FMT_FLOAT, FMT_ENUM };
enum ts_mode { TS_NONE, TS_UNIX_INT, TS_UNIX_FRAC, TS_DATETIME_24, TS_DATETIME_AMPM };

struct enum_def {
    char *str;
    int64_t value;
};

struct field {
    char *name;
    unsigned int start_bit, end_bit;
    enum field_fmt format;
    bool big_endian;
    float scaling, offset;
    int64_t default_value;
    struct enum_def *enums;
    size_t enum_count;
};

struct ookd_device {
    char *name, *description;
    unsigned int num_bits;
    /* state machine, microsecond terms */
    char **state_names;
    struct ookd_sm_state_us *states;
    struct ookd_sm_trigger_us *triggers;
    struct ookd_sm_desc desc;
    /* formatter */
    struct field *fields;
    size_t num_fields;
    enum ts_mode ts_mode;
};

static unsigned int field_width(const struct field *f) { return f->end_bit - f->start_bit + 1; }

static uint64_t field_mask(const struct field *f)
{
    const unsigned int w = field_width(f);
    return (w < 64) ? ((1ull << w) - 1) : ~0ull;
}

static bool parse_u64(const char *s, uint64_t *out)        /* str2uint64, src/conversions.c:115-140 */
{
    char *end;
    errno = 0;
    const unsigned long long v = strtoull(s, &end, 0);
    if (errno != 0 || end == s || *end != '\0') {
        return false;
    }
    *out = v;
    return true;
}

static bool parse_i64(const char *s, int64_t *out)         /* str2int64 (uses strtol), :88-112 */
{
    char *end;
    errno = 0;
    const long v = strtol(s, &end, 0);
    if (errno != 0 || end == s || *end != '\0') {
        return false;
    }
    *out = v;
    return true;
}

/* String -> raw field bits: str_to_spt, src/formatter.c:140-255 */
static bool field_bits_from_str(const struct field *f, const char *str, uint64_t *bits)
{
    const unsigned int w = field_width(f);
    const uint64_t mask = field_mask(f);
    int64_t value = 0;

    switch (f->format) {
        case FMT_HEX:
        case FMT_UNSIGNED_DEC: {
            uint64_t tmp;
            if (!parse_u64(str, &tmp)) goto inval;
            tmp = (uint64_t) (((float) tmp - f->offset) / f->scaling);
            value = (int64_t) tmp;
            break;
        }
        case FMT_TWOS_COMPLEMENT: {
            int64_t tmp;
            if (!parse_i64(str, &tmp)) goto inval;
            tmp = (int64_t) (((float) tmp - f->offset) / f->scaling);
            value = tmp & (int64_t) mask;
            break;
        }
        case FMT_SIGN_MAGNITUDE: {
            int64_t tmp;
            if (!parse_i64(str, &tmp)) goto inval;
            const bool negative = tmp < 0;
            tmp = (int64_t) (((float) tmp - f->offset) / f->scaling);
            tmp &= (1 << (w - 1)) - 1;
            if (negative) {
                tmp |= (1 << (w - 1));
            }
            value = tmp;
            break;
        }
        case FMT_FLOAT: {
            char *end;
            errno = 0;
            const double dv = strtod(str, &end);
            if (errno != 0 || end == str || *end != '\0') goto inval;
            const float tmp = ((float) dv - f->offset) / f->scaling;    /* spt_from_float, src/spt.h:56-60 */
            value = ((int64_t) tmp) & (int64_t) mask;
            break;
        }
        case FMT_ENUM: {
            bool have = false;
            for (size_t i = 0; i < f->enum_count && !have; i++) {
                if (!strcasecmp(str, f->enums[i].str)) {
                    value = f->enums[i].value;
                    have = true;
                }
            }
            if (!have) {
                uint64_t tmp;
                if (!parse_u64(str, &tmp)) goto inval;
                value = (int64_t) tmp;
            }
            break;
        }
    }
    if (((uint64_t) value & mask) != (uint64_t) value) {
        log_error("Value is too large for field \"%s\": %s\n", f->name, str);
        return false;
    }
    *bits = (uint64_t) value;
    return true;

inval:
    log_error("Invalid value for field \"%s\": %s\n", f->name, str);
    return false;
}

/* Field bits -> message bytes: apply_field_bits, src/formatter.c:766-800.  (The reference tests
 * `input_bits & (1 << src_bit)` with an int shift; fields wider than 31 bits are not exercised by
 * any shipped device, and the 64-bit test used here is what that line means.) */
static void apply_field_bits(const struct field *f, uint64_t bits, uint8_t *data)
{
    unsigned int src = f->big_endian ? (f->end_bit - f->start_bit) : 0;
    for (unsigned int i = f->start_bit; i <= f->end_bit; i++) {
        if (bits & (1ull << src)) {
            data[i / 8] |= (uint8_t) (1u << (i % 8));
        } else {
            data[i / 8] &= (uint8_t) ~(1u << (i % 8));
        }
        src += f->big_endian ? -1 : 1;
    }
}

/* Message bytes -> field value: get_field_value, src/formatter.c:425-455 */
static int64_t get_field_value(const struct field *f, const uint8_t *data)
{
    uint64_t tmp = 0;
    unsigned int dest = f->big_endian ? (f->end_bit - f->start_bit) : 0;
    for (unsigned int i = f->start_bit; i <= f->end_bit; i++) {
        const uint64_t v = (data[i / 8] >> (i % 8)) & 1;
        tmp |= v << dest;
        dest += f->big_endian ? -1 : 1;
    }
    return (int64_t) tmp;
}

/* field_data_to_str, src/formatter.c:457-573 (format strings and integer/float conversions kept
 * as they are there, including the 0x%02x used for 9..16-bit fields and the decimal PRIu64 used
 * for fields wider than 32 bits). */
static void field_to_str(char *str, size_t max, int64_t value, const struct field *f)
{
    const unsigned int w = field_width(f);
    const uint64_t mask = field_mask(f);

    switch (f->format) {
        case FMT_HEX: {
            uint64_t tmp = (uint64_t) value;
            tmp = (uint64_t) (((float) tmp * f->scaling) + f->offset);
            if (w <= 8)       snprintf(str, max, "0x%02x", (unsigned) (uint8_t) tmp);
            else if (w <= 16) snprintf(str, max, "0x%02x", (unsigned) (uint16_t) tmp);
            else if (w <= 24) snprintf(str, max, "0x%06x", (uint32_t) tmp);
            else if (w <= 32) snprintf(str, max, "0x%08x", (uint32_t) tmp);
            else if (w <= 40) snprintf(str, max, "0x%010" PRIu64, tmp);
            else if (w <= 48) snprintf(str, max, "0x%012" PRIu64, tmp);
            else if (w <= 56) snprintf(str, max, "0x%014" PRIu64, tmp);
            else              snprintf(str, max, "0x%016" PRIu64, tmp);
            break;
        }
        case FMT_UNSIGNED_DEC: {
            uint64_t tmp = (uint64_t) value;
            tmp = (uint64_t) (((float) tmp * f->scaling) + f->offset);
            snprintf(str, max, "%" PRIu64, tmp);
            break;
        }
        case FMT_TWOS_COMPLEMENT: {
            const bool neg = (value & (1 << (w - 1))) != 0;
            if (neg) {
                value = (int64_t) (((uint64_t) ~value + 1) & mask);
            }
            int64_t tmp = value;
            if (neg) {
                tmp = -tmp;
            }
            tmp = (int64_t) (((float) tmp * f->scaling) + f->offset);
            snprintf(str, max, "%" PRIi64, tmp);
            break;
        }
        case FMT_SIGN_MAGNITUDE: {
            const uint64_t u = (uint64_t) value;
            const bool neg = (u & (1 << (w - 1))) != 0;
            int64_t tmp = (int64_t) (u & ((1 << (w - 1)) - 1));
            if (neg) {
                tmp = -tmp;
            }
            tmp = (int64_t) (((float) tmp * f->scaling) + f->offset);
            snprintf(str, max, "%" PRIi64, tmp);
            break;
        }
        case FMT_FLOAT: {
            float scaling = f->scaling;
            const bool neg = (value & (1 << (w - 1))) != 0;
            if (neg) {
                value = (int64_t) (((uint64_t) ~value + 1) & mask);
                scaling = -f->scaling;
            }
            const float tmp = (float) value * scaling + f->offset;      /* spt_to_float, src/spt.h:82-85 */
            snprintf(str, max, "%1.3f", tmp);
            break;
        }
        case FMT_ENUM: {
            bool have = false;
            for (size_t i = 0; i < f->enum_count && !have; i++) {
                if (f->enums[i].value == value) {
                    snprintf(str, max, "%s", f->enums[i].str);
                    have = true;
                }
            }
            if (!have) {
                snprintf(str, max, "0x%" PRIx64, (uint64_t) value);
            }
            break;
        }
    }
}

/* timestamp, src/formatter.c:618-713.  (The reference's integer "unix" mode prints an
 * uninitialised buffer, formatter.c:636-640; the rounded integer it meant to print is used.) */
static bool append_timestamp(const struct ookd_device *d, struct ookd_keyval_list *out)
{
    char buf[80];
    static const char key[] = "Decode Timestamp";

    switch (d->ts_mode) {
        case TS_NONE:
            return true;
        case TS_UNIX_INT:
        case TS_UNIX_FRAC: {
            struct timeval tv;
            if (gettimeofday(&tv, NULL) != 0) {
                log_error("Failed to get current time: %s\n", strerror(errno));
                return false;
            }
            const double ts = tv.tv_sec + ((double) tv.tv_usec / 1000000.0);
            if (d->ts_mode == TS_UNIX_FRAC) {
                snprintf(buf, sizeof(buf), "%f", ts);
            } else {
                snprintf(buf, sizeof(buf), "%" PRIu64, (uint64_t) (ts + 0.5));
            }
            return ookd_keyval_list_append(out, key, buf);
        }
        case TS_DATETIME_24:
        case TS_DATETIME_AMPM: {
            const time_t t = time(NULL);
            struct tm *tmv = localtime(&t);
            if (!tmv) {
                log_error("Failed to get local time.\n");
                return false;
            }
            const size_t len = strftime(buf, sizeof(buf),
                                        d->ts_mode == TS_DATETIME_AMPM ? "%Y-%m-%d %I:%M:%S %p" : "%Y-%m-%d %H:%M:%S",
                                        tmv);
            if (len == 0) {
                log_error("Failed to format timestamp.\n");
                return false;
            }
            return ookd_keyval_list_append(out, key, buf);
        }
    }
    return false;
}

bool ookd_device_format(const struct ookd_device *d, const uint8_t *data, struct ookd_keyval_list *out)
{
    char buf[80];
    if (!append_timestamp(d, out)) {
        log_error("Failed to timestamp message.\n");
    }
    for (size_t i = 0; i < d->num_fields; i++) {
        memset(buf, 0, sizeof(buf));
        field_to_str(buf, sizeof(buf), get_field_value(&d->fields[i], data), &d->fields[i]);
        if (!ookd_keyval_list_append(out, d->fields[i].name, buf)) {
            return false;
        }
    }
    return true;
}

bool ookd_device_message(const struct ookd_device *d, const struct ookd_keyval_list *params, uint8_t *data)
{
    memset(data, 0, (d->num_bits + 7) / 8);
    for (size_t i = 0; i < d->num_fields; i++) {            /* formatter_default_data */
        apply_field_bits(&d->fields[i], (uint64_t) d->fields[i].default_value, data);
    }
    for (size_t i = 0; params && i < params->n; i++) {       /* formatter_keyval_to_data */
        const struct field *f = NULL;
        for (size_t j = 0; j < d->num_fields && !f; j++) {
            if (!strcasecmp(d->fields[j].name, params->items[i].key)) {
                f = &d->fields[j];
            }
        }
        if (!f) {
            log_error("Invalid parameter name: %s\n", params->items[i].key);
            return false;
        }
        uint64_t bits;
        if (!field_bits_from_str(f, params->items[i].value, &bits)) {
            return false;
        }
        apply_field_bits(f, bits, data);
    }
    return true;
}

/* ---- state slot assignment: get_or_reserve_state, src/state_machine.c:206-247 ---- */
static int slot_of(struct ookd_device *d, const char *name)
{
    const unsigned int n = d->desc.num_states;
    if (!strcasecmp("reset", name) && d->state_names[0] == NULL) {
        d->state_names[0] = strdup(name);
        return d->state_names[0] ? 0 : -1;
    }
    for (unsigned int i = 0; i < n; i++) {
        if (d->state_names[i] == NULL) {
            d->state_names[i] = strdup(name);
            return d->state_names[i] ? (int) i : -1;
        } else if (!strcmp(d->state_names[i], name)) {
            return (int) i;
        }
    }
    log_error("No room left to add state \"%s\"\n", name);
    return -1;
}

static int cond_value(const char *s)        /* sm_trigger_cond_value, src/state_machine.c:332-347 */
{
    if (!strcasecmp(s, "always")) return OOKD_COND_ALWAYS;
    if (!strcasecmp(s, "pulse_start")) return OOKD_COND_PULSE_START;
    if (!strcasecmp(s, "pulse_end")) return OOKD_COND_PULSE_END;
    if (!strcasecmp(s, "timeout")) return OOKD_COND_TIMEOUT;
    if (!strcasecmp(s, "msg_complete")) return OOKD_COND_MSG_COMPLETE;
    return 0;
}

static int action_value(const char *s)      /* sm_trigger_action_value, :349-362 */
{
    if (!strcasecmp(s, "none")) return OOKD_ACT_NONE;
    if (!strcasecmp(s, "append_0")) return OOKD_ACT_APPEND_0;
    if (!strcasecmp(s, "append_1")) return OOKD_ACT_APPEND_1;
    if (!strcasecmp(s, "output_data")) return OOKD_ACT_OUTPUT_DATA;
    return 0;
}

/* create_state_machine / add_state, src/device.c:76-254 */
static bool load_states(struct ookd_device *d, const struct oj_value *dev, unsigned int sample_rate)
{
    const struct oj_value *states = oj_get(dev, "states");
    if (!oj_is_array(states)) {
        log_error("Failed to get states array.\n");
        return false;
    }
    if (states->n == 0) {
        log_error("States array is empty.\n");
        return false;
    }
    size_t total_trig = 0;
    for (size_t i = 0; i < states->n; i++) {
        const struct oj_value *tr = oj_get(states->items[i], "triggers");
        if (oj_is_array(tr)) {
            total_trig += tr->n;
        }
    }
    d->desc.num_states = (uint32_t) states->n;
    d->desc.max_bits = d->num_bits;
    d->desc.sample_rate = sample_rate;
    d->state_names = calloc(states->n, sizeof(char *));
    d->states = calloc(states->n, sizeof(d->states[0]));
    d->triggers = calloc(total_trig ? total_trig : 1, sizeof(d->triggers[0]));
    bool *defined = calloc(states->n, sizeof(bool));
    if (!d->state_names || !d->states || !d->triggers || !defined) {
        free(defined);
        return false;
    }
    uint32_t q = 0;
    bool ok = true;
    for (size_t i = 0; i < states->n && ok; i++) {
        const struct oj_value *st = states->items[i];
        const struct oj_value *tmp = oj_get(st, "name");
        if (!oj_is_string(tmp)) {
            log_error("Failed to get state name.\n");
            ok = false;
            break;
        }
        const char *name = tmp->s;
        uint64_t timeout_us = 0, duration_us = 0;

        tmp = oj_get(st, "timeout_us");
        if (oj_is_int(tmp)) {
            const int v = (int) tmp->i;
            if (v < 0) {
                log_error("Invalid timeout value: %d\n", v);
                ok = false;
                break;
            }
            timeout_us = (uint64_t) v;
        }
        tmp = oj_get(st, "duration_us");
        if (oj_is_int(tmp)) {
            const int v = (int) tmp->i;
            if (v < 0) {
                log_error("Invalid trigger duration.\n");
            } else {
                duration_us = (uint64_t) v;
            }
        }
        const struct oj_value *trigs = oj_get(st, "triggers");
        if (!oj_is_array(trigs)) {
            log_error("Failed to get triggers for state \"%s\"\n", name);
            ok = false;
            break;
        }
        if (trigs->n == 0) {
            log_error("Triggers array is empty for state \"%s\"\n", name);
            ok = false;
            break;
        }
        const int slot = slot_of(d, name);
        if (slot < 0) {
            log_error("Failed to add \"%s\" to state machine.\n", name);
            ok = false;
            break;
        }
        if (defined[slot]) {
            log_warning("State may be getting initialized more than once: %s\n", name);
        }
        defined[slot] = true;
        d->states[slot].duration_us = duration_us;
        d->states[slot].timeout_us = timeout_us;
        d->states[slot].first_trigger = q;
        d->states[slot].num_triggers = (uint32_t) trigs->n;
        for (size_t t = 0; t < trigs->n; t++) {
            const struct oj_value *jt = trigs->items[t];
            struct ookd_sm_trigger_us *o = &d->triggers[q++];
            tmp = oj_get(jt, "condition");
            if (!oj_is_string(tmp)) {
                log_error("Failed to get trigger condition.\n");
                ok = false;
                break;
            }
            o->cond = cond_value(tmp->s);
            if (o->cond == 0) {
                log_error("Got invalid trigger condition: %s\n", tmp->s);
                ok = false;
                break;
            }
            tmp = oj_get(jt, "duration_us");
            o->duration_us = oj_is_int(tmp) ? (uint64_t) tmp->d : 0;
            tmp = oj_get(jt, "state");
            if (!oj_is_string(tmp)) {
                log_error("Failed to get trigger's next state.\n");
                ok = false;
                break;
            }
            const int next = slot_of(d, tmp->s);
            if (next < 0) {
                ok = false;
                break;
            }
            o->next_state = (uint32_t) next;
            tmp = oj_get(jt, "action");
            if (oj_is_string(tmp)) {
                o->action = action_value(tmp->s);
                if (o->action == 0) {
                    log_error("Got invalid trigger action: %s\n", tmp->s);
                    ok = false;
                    break;
                }
            } else {
                o->action = OOKD_ACT_NONE;
            }
        }
    }
    if (ok) {
        for (size_t i = 0; i < states->n; i++) {        /* sm_initialized, src/state_machine.c:176-204 */
            if (!defined[i]) {
                log_error("State machine is missing states or triggers.\n");
                ok = false;
                break;
            }
        }
    }
    free(defined);
    d->desc.num_triggers = q;
    d->desc.states = d->states;
    d->desc.triggers = d->triggers;
    return ok;
}

static bool fmt_value(const char *s, enum field_fmt *out)   /* formatter_fmt_value, src/formatter.c:860-877 */
{
    if (!strcasecmp("hex", s)) *out = FMT_HEX;
    else if (!strcasecmp("unsigned decimal", s)) *out = FMT_UNSIGNED_DEC;
    else if (!strcasecmp("sign-magnitude", s)) *out = FMT_SIGN_MAGNITUDE;
    else if (!strcasecmp("two's complement", s)) *out = FMT_TWOS_COMPLEMENT;
    else if (!strcasecmp("float", s)) *out = FMT_FLOAT;
    else if (!strcasecmp("enumeration", s)) *out = FMT_ENUM;
    else return false;
    return true;
}

/* create_formatter / add_field, src/device.c:256-498 */
static bool load_fields(struct ookd_device *d, const struct oj_value *dev)
{
    const struct oj_value *fields = oj_get(dev, "fields");
    if (!oj_is_array(fields)) {
        log_error("Failed to get fields array.\n");
        return false;
    }
    if (fields->n == 0) {
        log_error("Fields array is empty.\n");
        return false;
    }
    const struct oj_value *ts = oj_get(dev, "ts_mode");
    d->ts_mode = TS_NONE;
    if (ts) {
        if (!oj_is_string(ts)) {
            log_error("'ts_mode' must be a string.\n");
            return false;
        }
        if (!strcasecmp("none", ts->s)) d->ts_mode = TS_NONE;
        else if (!strcasecmp("unix", ts->s)) d->ts_mode = TS_UNIX_INT;
        else if (!strcasecmp("unix-frac", ts->s)) d->ts_mode = TS_UNIX_FRAC;
        else if (!strcasecmp("datetime-24", ts->s)) d->ts_mode = TS_DATETIME_24;
        else if (!strcasecmp("datetime-ampm", ts->s)) d->ts_mode = TS_DATETIME_AMPM;
        else {
            log_error("Invalid 'ts_mode' value: %s\n", ts->s);
            return false;
        }
    }
    d->fields = calloc(fields->n, sizeof(d->fields[0]));
    if (!d->fields) {
        return false;
    }
    d->num_fields = fields->n;
    for (size_t i = 0; i < fields->n; i++) {
        const struct oj_value *jf = fields->items[i];
        struct field *f = &d->fields[i];
        const struct oj_value *tmp;

        tmp = oj_get(jf, "name");
        if (!oj_is_string(tmp)) {
            log_error("Failed to get field name.\n");
            return false;
        }
        f->name = strdup(tmp->s);
        const struct oj_value *def = oj_get(jf, "default");
        if (!oj_is_string(def)) {
            log_error("Failed to get default for \"%s\" field.\n", f->name);
            return false;
        }
        tmp = oj_get(jf, "start_bit");
        if (!oj_is_int(tmp)) {
            log_error("Failed to get start bit for \"%s\" field.\n", f->name);
            return false;
        }
        f->start_bit = (unsigned int) (int) tmp->i;
        tmp = oj_get(jf, "end_bit");
        if (!oj_is_int(tmp)) {
            log_error("Failed to get end bit for \"%s\" field.\n", f->name);
            return false;
        }
        f->end_bit = (unsigned int) (int) tmp->i;
        if (f->end_bit < f->start_bit) {
            log_error("End bit must be >= start bit\n");
            return false;
        }
        if (f->end_bit - f->start_bit + 1 > 64) {
            log_error("Fields larger than 64-bits are not currently supported.\n");
            return false;
        }
        if (f->end_bit >= 8 * ((d->num_bits + 7) / 8)) {
            log_error("Field \"%s\" lies outside the %u-bit message.\n", f->name, d->num_bits);
            return false;
        }
        tmp = oj_get(jf, "endianness");
        if (!oj_is_string(tmp)) {
            log_error("Failed to get endianness for \"%s\" field.\n", f->name);
            return false;
        }
        if (!strcasecmp("big", tmp->s)) f->big_endian = true;
        else if (!strcasecmp("little", tmp->s)) f->big_endian = false;
        else {
            log_error("Invalid endianness specified: %s\n", tmp->s);
            return false;
        }
        tmp = oj_get(jf, "format");
        if (!oj_is_string(tmp)) {
            log_error("Failed to get format for \"%s\" field.\n", f->name);
            return false;
        }
        if (!fmt_value(tmp->s, &f->format)) {
            log_error("Invalid format: %s\n", tmp->s);
            return false;
        }
        float offset = 0, scaling = 0;
        tmp = oj_get(jf, "offset");
        if (oj_is_number(tmp)) offset = (float) tmp->d;
        tmp = oj_get(jf, "scaling");
        if (oj_is_number(tmp)) scaling = (float) tmp->d;
        f->scaling = (scaling == 0) ? 1.0f : scaling;        /* formatter_add_field, src/formatter.c:289 */
        f->offset = offset;

        if (f->format == FMT_ENUM) {
            const struct oj_value *enums = oj_get(jf, "enum_values");
            if (!oj_is_array(enums)) {
                log_error("No \"enum_values\" array found for enumeration: %s\n", f->name);
                return false;
            }
            if (enums->n == 0) {
                log_error("Enumeration format requires 1 or more values to be defined\n");
                return false;
            }
            f->enums = calloc(enums->n, sizeof(f->enums[0]));
            if (!f->enums) {
                return false;
            }
            for (size_t e = 0; e < enums->n; e++) {
                const struct oj_value *s = oj_get(enums->items[e], "string");
                const struct oj_value *v = oj_get(enums->items[e], "value");
                if (!oj_is_string(s)) {
                    log_error("Enumeration value %zd is missing \"string.\"\n", e);
                    return false;
                }
                if (!oj_is_string(v)) {
                    log_error("Enumeration item \"%s\" is missing \"value.\"\n", s->s);
                    return false;
                }
                uint64_t val;
                if (!parse_u64(v->s, &val)) {
                    log_error("Invalid enumeration value: %s\n", v->s);
                    return false;
                }
                for (size_t k = 0; k < f->enum_count; k++) {
                    if (!strcasecmp(f->enums[k].str, s->s)) {
                        log_error("Error: Duplicate enumeration name (%s)\n", s->s);
                        return false;
                    }
                }
                f->enums[f->enum_count].str = strdup(s->s);
                f->enums[f->enum_count].value = (int64_t) val;
                f->enum_count++;
            }
        }
        uint64_t bits;
        if (!field_bits_from_str(f, def->s, &bits)) {
            log_error("Invalid default value for field \"%s\": %s\n", f->name, def->s);
            return false;
        }
        f->default_value = (int64_t) bits;
    }
    return true;
}

void ookd_device_deinit(struct ookd_device *d)
{
    if (!d) {
        return;
    }
    for (size_t i = 0; i < d->num_fields; i++) {
        for (size_t e = 0; e < d->fields[i].enum_count; e++) {
            free(d->fields[i].enums[e].str);
        }
        free(d->fields[i].enums);
        free(d->fields[i].name);
    }
    free(d->fields);
    if (d->state_names) {
        for (uint32_t i = 0; i < d->desc.num_states; i++) {
            free(d->state_names[i]);
        }
    }
    free(d->state_names);
    free(d->states);
    free(d->triggers);
    free(d->name);
    free(d->description);
    free(d);
}

/* device_init / populate_device, src/device.c:500-632 */
struct ookd_device *ookd_device_init(const char *device_name, unsigned int sample_rate)
{
    char err[200];
    FILE *in = ookd_find_device_file(device_name);
    if (!in) {
        log_error("Unable to find device file for: %s\n", device_name);
        return NULL;
    }
    struct oj_value *root = oj_parse_file(in, err, sizeof(err));
    fclose(in);
    if (!root) {
        log_error("Error in %s.json (%s)\n", device_name, err);
        return NULL;
    }
    struct ookd_device *d = calloc(1, sizeof(*d));
    bool ok = false;
    const struct oj_value *dev = oj_get(root, "device");
    if (!d) {
        goto out;
    }
    if (!dev) {
        log_error("Failed to find \"device\" entry in device file\n");
        goto out;
    }
    const struct oj_value *tmp = oj_get(dev, "name");
    if (!oj_is_string(tmp)) {
        log_error("Failed to read device name string.\n");
        goto out;
    }
    d->name = strdup(tmp->s);
    tmp = oj_get(dev, "description");
    if (!oj_is_string(tmp)) {
        log_error("Failed to read device description string.\n");
        goto out;
    }
    d->description = strdup(tmp->s);
    tmp = oj_get(dev, "num_bits");
    if (!oj_is_int(tmp)) {
        log_error("Failed to read \"num_bits\" property.\n");
        goto out;
    }
    if ((int) tmp->i <= 0) {
        log_error("Invalid \"num_bits\" value: %d\n", (int) tmp->i);
        goto out;
    }
    if (tmp->i > 8 * OOKD_MSG_BYTES) {
        log_error("\"num_bits\" above %d is not supported.\n", 8 * OOKD_MSG_BYTES);
        goto out;
    }
    d->num_bits = (unsigned int) tmp->i;
    if (sample_rate == 0) {
        log_error("Invalid sample rate.\n");
        goto out;
    }
    if (!load_states(d, dev, sample_rate)) {
        goto out;
    }
    if (!load_fields(d, dev)) {
        goto out;
    }
    ok = true;

out:
    oj_free(root);
    if (!ok) {
        ookd_device_deinit(d);
        d = NULL;
    }
    return d;
}

const struct ookd_sm_desc *ookd_device_sm_desc(const struct ookd_device *d) { return &d->desc; }
unsigned int ookd_device_num_bits(const struct ookd_device *d) { return d->num_bits; }
const char *ookd_device_name(const struct ookd_device *d) { return d->name; }

/* ------------------------------------------------------------------------- */
/* transmit-side generator: sm_generate, src/state_machine.c:565-873          */
/* ------------------------------------------------------------------------- */
struct gen {
    const struct ookd_device *d;
    uint32_t cur;
    unsigned int num_bits;
    int level;
    uint32_t *runs;
    size_t n, cap;
};

static bool gen_append(struct gen *g, uint64_t duration_us)
{
    /* to_sample_count, src/state_machine.c:88-92 */
    const unsigned int count = (unsigned int) (duration_us * ((double) g->d->desc.sample_rate / 1e6) + 0.5);
    if (count == 0) {
        return true;
    }
    if (g->n + 2 > g->cap) {
        const size_t nc = g->cap ? 2 * g->cap : 256;
        void *tmp = realloc(g->runs, nc * sizeof(uint32_t));
        if (!tmp) {
            return false;
        }
        g->runs = tmp;
        g->cap = nc;
    }
    g->runs[g->n++] = (uint32_t) g->level;
    g->runs[g->n++] = count;
    return true;
}

/* get_tx_trigger, :626-699.  Returns 1 found, 0 none, -1 failure. */
static int gen_pick(struct gen *g, bool bit, bool check_action, const struct ookd_sm_trigger_us **out)
{
    const struct ookd_sm_state_us *st = &g->d->states[g->cur];
    *out = NULL;
    for (uint32_t i = 0; i < st->num_triggers; i++) {
        const struct ookd_sm_trigger_us *t = &g->d->triggers[st->first_trigger + i];
        if (check_action) {
            const bool have = (t->action == OOKD_ACT_APPEND_0 && !bit) || (t->action == OOKD_ACT_APPEND_1 && bit) ||
                              t->action == OOKD_ACT_OUTPUT_DATA;
            if (!have) {
                continue;
            }
        }
        switch (t->cond) {
            case OOKD_COND_MSG_COMPLETE:
                if (g->num_bits == g->d->num_bits) {
                    *out = t;
                    return 1;
                }
                break;
            case OOKD_COND_ALWAYS:
            case OOKD_COND_PULSE_START:
            case OOKD_COND_PULSE_END:
                *out = t;
                return 1;
            case OOKD_COND_TIMEOUT:
                log_error("Encountered SM_TRIGGER_COND_TIMEOUT while generating samples. "
                          "Bug in state machine design?\n");
                return -1;
            default:
                return -1;
        }
    }
    return 0;
}

/* generate / handle_tx_triggers, :701-823 */
static bool gen_bit(struct gen *g, bool bit)
{
    bool done = false;
    while (!done) {
        const struct ookd_sm_trigger_us *t;
        int r = gen_pick(g, bit, true, &t);
        if (r < 0) {
            return false;
        }
        if (r == 0) {
            r = gen_pick(g, bit, false, &t);
            if (r <= 0) {
                return false;
            }
        }
        if (g->d->states[g->cur].duration_us == 0 && t->duration_us != 0) {
            if (!gen_append(g, t->duration_us)) {
                return false;
            }
        }
        if (t->cond == OOKD_COND_PULSE_START) {
            if (g->level) {
                log_error("Bug? Logic value is already 1, got PULSE_START\n");
                return false;
            }
            g->level = 1;
        } else if (t->cond == OOKD_COND_PULSE_END) {
            if (!g->level) {
                log_error("Bug? Logic value is already 0, got PULSE_END\n");
                return false;
            }
            g->level = 0;
        }
        if (t->action == OOKD_ACT_APPEND_0 || t->action == OOKD_ACT_APPEND_1) {
            if (g->num_bits < g->d->num_bits) {
                g->num_bits++;
                done = true;
            } else if (g->num_bits > g->d->num_bits) {
                return false;
            }
        } else if (t->action == OOKD_ACT_OUTPUT_DATA) {
            done = true;
        }
        g->cur = t->next_state;
        if (g->d->states[g->cur].duration_us != 0) {
            if (!gen_append(g, g->d->states[g->cur].duration_us)) {
                return false;
            }
        }
    }
    return true;
}

bool ookd_device_generate_runs(const struct ookd_device *d, const uint8_t *data, uint32_t **runs, size_t *n_runs)
{
    struct gen g;
    memset(&g, 0, sizeof(g));
    g.d = d;
    bool ok = true;
    for (unsigned int i = 0; i < d->num_bits && ok; i++) {
        ok = gen_bit(&g, (data[i / 8] & (1u << (i % 8))) != 0);
    }
    if (ok) {
        ok = gen_bit(&g, false);        /* data-independent remainder of the signal */
    }
    if (!ok) {
        free(g.runs);
        *runs = NULL;
        *n_runs = 0;
        return false;
    }
    *runs = g.runs;
    *n_runs = g.n / 2;
    return true;
}

bool ookd_device_toggles(const struct ookd_device *d, const uint8_t *msgs, size_t n_msgs, uint64_t lead_samples,
                         uint64_t start, uint64_t **toggles, size_t *n_toggles, uint64_t *total_samples)
{
    const size_t nbytes = (d->num_bits + 7) / 8;
    uint64_t pos = start;
    uint64_t *tg = NULL;
    size_t n = 0, cap = 0;
    for (size_t m = 0; m < n_msgs; m++) {
        uint32_t *runs;
        size_t n_runs;
        if (!ookd_device_generate_runs(d, msgs + m * nbytes, &runs, &n_runs)) {
            free(tg);
            return false;
        }
        pos += lead_samples;
        int level = 0;
        for (size_t r = 0; r <= n_runs; r++) {
            const int lvl = (r < n_runs) ? (int) runs[2 * r] : 0;       /* message ends low */
            if (lvl != level) {
                if (n == cap) {
                    cap = cap ? 2 * cap : 1024;
                    void *tmp = realloc(tg, cap * sizeof(uint64_t));
                    if (!tmp) {
                        free(tg);
                        free(runs);
                        return false;
                    }
                    tg = tmp;
                }
                tg[n++] = pos;
                level = lvl;
            }
            if (r < n_runs) {
                pos += runs[2 * r + 1];
            }
        }
        free(runs);
    }
    *toggles = tg;
    *n_toggles = n;
    *total_samples = pos;
    return true;
}
