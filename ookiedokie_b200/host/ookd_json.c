/* ookd_json.c -- see ookd_json.h */
#include "ookd_json.h"

#include <errno.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

struct parser {
    const char *p, *end;
    int line;
    char *err;
    size_t err_len;
    int failed;
};

static void fail(struct parser *ps, const char *fmt, ...)
{
    if (!ps->failed && ps->err && ps->err_len) {
        char msg[160];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(msg, sizeof(msg), fmt, ap);
        va_end(ap);
        snprintf(ps->err, ps->err_len, "line %d: %s", ps->line, msg);
    }
    ps->failed = 1;
}

static void skip_ws(struct parser *ps)
{
    while (ps->p < ps->end) {
        const char c = *ps->p;
        if (c == '\n') {
            ps->line++;
        } else if (c != ' ' && c != '\t' && c != '\r') {
            break;
        }
        ps->p++;
    }
}

static struct oj_value *new_value(enum oj_type t)
{
    struct oj_value *v = calloc(1, sizeof(*v));
    if (v) {
        v->type = t;
    }
    return v;
}

void oj_free(struct oj_value *v)
{
    if (!v) {
        return;
    }
    for (size_t i = 0; i < v->n; i++) {
        if (v->items) {
            oj_free(v->items[i]);
        }
        if (v->keys) {
            free(v->keys[i]);
        }
    }
    free(v->items);
    free(v->keys);
    free(v->s);
    free(v);
}

static int hex4(struct parser *ps, unsigned *out)
{
    unsigned v = 0;
    if (ps->end - ps->p < 4) {
        return 0;
    }
    for (int i = 0; i < 4; i++) {
        const char c = ps->p[i];
        v <<= 4;
        if (c >= '0' && c <= '9') v |= (unsigned) (c - '0');
        else if (c >= 'a' && c <= 'f') v |= (unsigned) (c - 'a' + 10);
        else if (c >= 'A' && c <= 'F') v |= (unsigned) (c - 'A' + 10);
        else return 0;
    }
    ps->p += 4;
    *out = v;
    return 1;
}

static void put_utf8(char **w, unsigned cp)
{
    char *o = *w;
    if (cp < 0x80) {
        *o++ = (char) cp;
    } else if (cp < 0x800) {
        *o++ = (char) (0xC0 | (cp >> 6));
        *o++ = (char) (0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
        *o++ = (char) (0xE0 | (cp >> 12));
        *o++ = (char) (0x80 | ((cp >> 6) & 0x3F));
        *o++ = (char) (0x80 | (cp & 0x3F));
    } else {
        *o++ = (char) (0xF0 | (cp >> 18));
        *o++ = (char) (0x80 | ((cp >> 12) & 0x3F));
        *o++ = (char) (0x80 | ((cp >> 6) & 0x3F));
        *o++ = (char) (0x80 | (cp & 0x3F));
    }
    *w = o;
}

static char *parse_string_raw(struct parser *ps)
{
    if (ps->p >= ps->end || *ps->p != '"') {
        fail(ps, "expected string");
        return NULL;
    }
    ps->p++;
    /* the decoded string is never longer than the encoded one */
    const char *scan = ps->p;
    while (scan < ps->end && *scan != '"') {
        if (*scan == '\\' && scan + 1 < ps->end) {
            scan++;
        }
        scan++;
    }
    if (scan >= ps->end) {
        fail(ps, "unterminated string");
        return NULL;
    }
    char *out = malloc((size_t) (scan - ps->p) + 1);
    if (!out) {
        fail(ps, "out of memory");
        return NULL;
    }
    char *w = out;
    while (*ps->p != '"') {
        char c = *ps->p++;
        if ((unsigned char) c < 0x20) {
            fail(ps, "control character in string");
            free(out);
            return NULL;
        }
        if (c != '\\') {
            *w++ = c;
            continue;
        }
        c = *ps->p++;
        switch (c) {
            case '"': *w++ = '"'; break;
            case '\\': *w++ = '\\'; break;
            case '/': *w++ = '/'; break;
            case 'b': *w++ = '\b'; break;
            case 'f': *w++ = '\f'; break;
            case 'n': *w++ = '\n'; break;
            case 'r': *w++ = '\r'; break;
            case 't': *w++ = '\t'; break;
            case 'u': {
                unsigned cp, lo;
                if (!hex4(ps, &cp)) {
                    fail(ps, "bad \\u escape");
                    free(out);
                    return NULL;
                }
                if (cp >= 0xD800 && cp <= 0xDBFF && ps->end - ps->p >= 6 && ps->p[0] == '\\' && ps->p[1] == 'u') {
                    ps->p += 2;
                    if (!hex4(ps, &lo) || lo < 0xDC00 || lo > 0xDFFF) {
                        fail(ps, "bad surrogate pair");
                        free(out);
                        return NULL;
                    }
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                put_utf8(&w, cp);
                break;
            }
            default:
                fail(ps, "bad escape");
                free(out);
                return NULL;
        }
    }
    ps->p++;
    *w = '\0';
    return out;
}

static struct oj_value *parse_value(struct parser *ps, int depth);

static int push_item(struct oj_value *c, struct oj_value *item, char *key)
{
    struct oj_value **ni = realloc(c->items, sizeof(*ni) * (c->n + 1));
    if (!ni) {
        return 0;
    }
    c->items = ni;
    if (c->type == OJ_OBJECT) {
        char **nk = realloc(c->keys, sizeof(*nk) * (c->n + 1));
        if (!nk) {
            return 0;
        }
        c->keys = nk;
        c->keys[c->n] = key;
    }
    c->items[c->n++] = item;
    return 1;
}

static struct oj_value *parse_number(struct parser *ps)
{
    const char *s = ps->p;
    const char *q = s;
    int is_real = 0;
    if (q < ps->end && *q == '-') q++;
    if (q >= ps->end || *q < '0' || *q > '9') {
        fail(ps, "invalid number");
        return NULL;
    }
    if (*q == '0') {
        q++;
    } else {
        while (q < ps->end && *q >= '0' && *q <= '9') q++;
    }
    if (q < ps->end && *q == '.') {
        is_real = 1;
        q++;
        if (q >= ps->end || *q < '0' || *q > '9') {
            fail(ps, "invalid number");
            return NULL;
        }
        while (q < ps->end && *q >= '0' && *q <= '9') q++;
    }
    if (q < ps->end && (*q == 'e' || *q == 'E')) {
        is_real = 1;
        q++;
        if (q < ps->end && (*q == '+' || *q == '-')) q++;
        if (q >= ps->end || *q < '0' || *q > '9') {
            fail(ps, "invalid number");
            return NULL;
        }
        while (q < ps->end && *q >= '0' && *q <= '9') q++;
    }
    char buf[512];
    const size_t len = (size_t) (q - s);
    if (len >= sizeof(buf)) {
        fail(ps, "number too long");
        return NULL;
    }
    memcpy(buf, s, len);
    buf[len] = '\0';
    struct oj_value *v = new_value(is_real ? OJ_REAL : OJ_INT);
    if (!v) {
        fail(ps, "out of memory");
        return NULL;
    }
    errno = 0;
    if (is_real) {
        v->d = strtod(buf, NULL);       /* decimal -> nearest double, as jansson does */
    } else {
        v->i = strtoll(buf, NULL, 10);
        if (errno == ERANGE) {
            fail(ps, "integer out of range");
            oj_free(v);
            return NULL;
        }
        v->d = (double) v->i;
    }
    ps->p = q;
    return v;
}

static struct oj_value *parse_value(struct parser *ps, int depth)
{
    if (depth > 64) {
        fail(ps, "nesting too deep");
        return NULL;
    }
    skip_ws(ps);
    if (ps->p >= ps->end) {
        fail(ps, "unexpected end of input");
        return NULL;
    }
    const char c = *ps->p;
    if (c == '{' || c == '[') {
        const int is_obj = (c == '{');
        const char close = is_obj ? '}' : ']';
        struct oj_value *v = new_value(is_obj ? OJ_OBJECT : OJ_ARRAY);
        if (!v) {
            fail(ps, "out of memory");
            return NULL;
        }
        ps->p++;
        skip_ws(ps);
        if (ps->p < ps->end && *ps->p == close) {
            ps->p++;
            return v;
        }
        for (;;) {
            char *key = NULL;
            if (is_obj) {
                skip_ws(ps);
                key = parse_string_raw(ps);
                if (!key) {
                    oj_free(v);
                    return NULL;
                }
                for (size_t i = 0; i < v->n; i++) {
                    if (!strcmp(v->keys[i], key)) {
                        fail(ps, "duplicate object key \"%s\"", key);
                        free(key);
                        oj_free(v);
                        return NULL;
                    }
                }
                skip_ws(ps);
                if (ps->p >= ps->end || *ps->p != ':') {
                    fail(ps, "expected ':'");
                    free(key);
                    oj_free(v);
                    return NULL;
                }
                ps->p++;
            }
            struct oj_value *item = parse_value(ps, depth + 1);
            if (!item || !push_item(v, item, key)) {
                if (item) {
                    fail(ps, "out of memory");
                    oj_free(item);
                }
                free(key);
                oj_free(v);
                return NULL;
            }
            skip_ws(ps);
            if (ps->p < ps->end && *ps->p == ',') {
                ps->p++;
                continue;
            }
            if (ps->p < ps->end && *ps->p == close) {
                ps->p++;
                return v;
            }
            fail(ps, "expected ',' or '%c'", close);
            oj_free(v);
            return NULL;
        }
    }
    if (c == '"') {
        char *s = parse_string_raw(ps);
        if (!s) {
            return NULL;
        }
        struct oj_value *v = new_value(OJ_STRING);
        if (!v) {
            free(s);
            fail(ps, "out of memory");
            return NULL;
        }
        v->s = s;
        return v;
    }
    if (c == '-' || (c >= '0' && c <= '9')) {
        return parse_number(ps);
    }
    static const struct { const char *word; enum oj_type t; long long i; } lits[] = {
        { "true", OJ_BOOL, 1 }, { "false", OJ_BOOL, 0 }, { "null", OJ_NULL, 0 } };
    for (size_t k = 0; k < 3; k++) {
        const size_t len = strlen(lits[k].word);
        if ((size_t) (ps->end - ps->p) >= len && !strncmp(ps->p, lits[k].word, len)) {
            struct oj_value *v = new_value(lits[k].t);
            if (!v) {
                fail(ps, "out of memory");
                return NULL;
            }
            v->i = lits[k].i;
            ps->p += len;
            return v;
        }
    }
    fail(ps, "unexpected character '%c'", c);
    return NULL;
}

struct oj_value *oj_parse_text(const char *text, size_t len, char *err, size_t err_len)
{
    struct parser ps = { text, text + len, 1, err, err_len, 0 };
    if (err && err_len) {
        err[0] = '\0';
    }
    struct oj_value *v = parse_value(&ps, 0);
    if (v) {
        skip_ws(&ps);
        if (ps.p != ps.end) {
            fail(&ps, "trailing data after document");
            oj_free(v);
            v = NULL;
        } else if (v->type != OJ_OBJECT && v->type != OJ_ARRAY) {
            fail(&ps, "document must be an object or array");
            oj_free(v);
            v = NULL;
        }
    }
    return v;
}

struct oj_value *oj_parse_file(FILE *f, char *err, size_t err_len)
{
    size_t cap = 1 << 16, len = 0;
    char *buf = malloc(cap);
    if (!buf) {
        return NULL;
    }
    for (;;) {
        const size_t got = fread(buf + len, 1, cap - len, f);
        len += got;
        if (got == 0) {
            break;
        }
        if (len == cap) {
            char *nb = realloc(buf, cap * 2);
            if (!nb) {
                free(buf);
                return NULL;
            }
            buf = nb;
            cap *= 2;
        }
    }
    struct oj_value *v = oj_parse_text(buf, len, err, err_len);
    free(buf);
    return v;
}

const struct oj_value *oj_get(const struct oj_value *obj, const char *key)
{
    if (!obj || obj->type != OJ_OBJECT) {
        return NULL;
    }
    for (size_t i = 0; i < obj->n; i++) {
        if (!strcmp(obj->keys[i], key)) {
            return obj->items[i];
        }
    }
    return NULL;
}
