/*
 * ookd_json.h -- minimal JSON document reader for the filter / device files
 * (filters/README.md, devices/README.md formats).  Keeps the distinctions the
 * reference loaders rely on: integer vs real numbers (json_is_integer,
 * src/device.c:85-140) and rejection of duplicate object keys
 * (JSON_REJECT_DUPLICATES, src/fir.c:87, src/device.c:597).
 */
#ifndef OOKD_JSON_H
#define OOKD_JSON_H

#include <stddef.h>
#include <stdio.h>

enum oj_type { OJ_NULL, OJ_BOOL, OJ_INT, OJ_REAL, OJ_STRING, OJ_ARRAY, OJ_OBJECT };

struct oj_value {
    enum oj_type type;
    long long   i;          /* OJ_INT, OJ_BOOL */
    double      d;          /* OJ_REAL, and (double) i for OJ_INT */
    char       *s;          /* OJ_STRING */
    size_t      n;          /* items (array) / members (object) */
    struct oj_value **items;
    char      **keys;       /* object member names */
};

/* Parse a whole stream.  On failure returns NULL and fills err (line, text). */
struct oj_value *oj_parse_file(FILE *f, char *err, size_t err_len);
struct oj_value *oj_parse_text(const char *text, size_t len, char *err, size_t err_len);
void oj_free(struct oj_value *v);

const struct oj_value *oj_get(const struct oj_value *obj, const char *key);  /* NULL if absent / not object */
static inline int oj_is_int(const struct oj_value *v)    { return v && v->type == OJ_INT; }
static inline int oj_is_number(const struct oj_value *v) { return v && (v->type == OJ_INT || v->type == OJ_REAL); }
static inline int oj_is_string(const struct oj_value *v) { return v && v->type == OJ_STRING; }
static inline int oj_is_array(const struct oj_value *v)  { return v && v->type == OJ_ARRAY; }
static inline int oj_is_object(const struct oj_value *v) { return v && v->type == OJ_OBJECT; }

#endif
