/*
 * TEST INFRASTRUCTURE ONLY -- declarations-only stand-in for <jansson.h>.
 *
 * The image carries the jansson 2.14 runtime (libjansson.so.4) but not its
 * development header, so the unmodified reference sources (which include
 * <jansson.h> from src/fir.c:28 and src/device.c:28) cannot be compiled
 * without one.  This file declares exactly the subset of the public jansson
 * 2.14 ABI those two files use; every definition lives in the system
 * libjansson.so.4 that oracle/Makefile links by full path.  Nothing in the
 * product links or includes this file.
 */
#ifndef OOKD_ORACLE_JANSSON_SHIM_H
#define OOKD_ORACLE_JANSSON_SHIM_H

#include <stdio.h>
#include <stddef.h>

typedef enum {
    JSON_OBJECT, JSON_ARRAY, JSON_STRING, JSON_INTEGER,
    JSON_REAL, JSON_TRUE, JSON_FALSE, JSON_NULL
} json_type;

typedef struct json_t {
    json_type type;
    volatile size_t refcount;
} json_t;

typedef long long json_int_t;

typedef struct json_error_t {
    int line;
    int column;
    int position;
    char source[80];
    char text[160];
} json_error_t;

#define JSON_REJECT_DUPLICATES 0x1

#define json_typeof(j)     ((j)->type)
#define json_is_object(j)  ((j) && json_typeof(j) == JSON_OBJECT)
#define json_is_array(j)   ((j) && json_typeof(j) == JSON_ARRAY)
#define json_is_string(j)  ((j) && json_typeof(j) == JSON_STRING)
#define json_is_integer(j) ((j) && json_typeof(j) == JSON_INTEGER)
#define json_is_real(j)    ((j) && json_typeof(j) == JSON_REAL)
#define json_is_number(j)  (json_is_integer(j) || json_is_real(j))

json_t *json_loadf(FILE *input, size_t flags, json_error_t *error);
json_t *json_object_get(const json_t *object, const char *key);
size_t json_array_size(const json_t *array);
json_t *json_array_get(const json_t *array, size_t index);
const char *json_string_value(const json_t *string);
json_int_t json_integer_value(const json_t *integer);
double json_number_value(const json_t *json);
void json_delete(json_t *json);

static inline void json_decref(json_t *json)
{
    if (json && json->refcount != (size_t) -1 &&
        __atomic_sub_fetch(&json->refcount, 1, __ATOMIC_RELEASE) == 0) {
        json_delete(json);
    }
}

#define json_array_foreach(array, index, value)                               \
    for (index = 0; index < json_array_size(array) &&                         \
                    (value = json_array_get(array, index));                   \
         index++)

#endif
