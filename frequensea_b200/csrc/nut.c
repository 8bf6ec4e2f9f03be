/*
 * nut.c -- host container of the operator API (include/nut.h).
 * Behavioural mirror of the reference's src/nut.c:18-190; written against the
 * header's contract (fresh caller-owned copies, element count = length *
 * channels, u8 <-> f64 by x/256), not transcribed.
 */
#define _POSIX_C_SOURCE 200809L
#include "nut.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* src/nut.c:18-23 */
void nut_sleep_milliseconds(int millis) {
    struct timespec req;
    req.tv_sec = millis / 1000;
    req.tv_nsec = (long)(millis % 1000) * 1000000L;
    nanosleep(&req, NULL);
}

static size_t elem_size(nut_buffer_type type) {
    return type == NUT_BUFFER_U8 ? sizeof(uint8_t) : sizeof(double);
}

/* shared constructor: zero-filled payload, optional copy-in (src/nut.c:25-50) */
static nut_buffer *make_buffer(nut_buffer_type type, int length, int channels, const void *src) {
    nut_buffer *b = (nut_buffer *)calloc(1, sizeof(*b));
    size_t bytes = (size_t)length * (size_t)channels * elem_size(type);
    b->type = type;
    b->length = length;
    b->channels = channels;
    b->size_bytes = (int)bytes;
    b->data.u8 = (uint8_t *)calloc(bytes ? bytes : 1, 1);
    if (src && bytes) memcpy(b->data.u8, src, bytes);
    return b;
}

nut_buffer *nut_buffer_new_u8(int length, int channels, const uint8_t *data) {
    return make_buffer(NUT_BUFFER_U8, length, channels, data);
}

nut_buffer *nut_buffer_new_f64(int length, int channels, const double *data) {
    return make_buffer(NUT_BUFFER_F64, length, channels, data);
}

void nut_buffer_free(nut_buffer *buffer) {
    if (!buffer) return;
    free(buffer->data.u8);     /* union: same pointer either way (src/nut.c:183-190) */
    free(buffer);
}

nut_buffer *nut_buffer_copy(nut_buffer *buffer) {
    assert(buffer != NULL);
    return make_buffer(buffer->type, buffer->length, buffer->channels, buffer->data.u8);
}

/* keep the leading `percentage` of the elements (src/nut.c:62-71) */
nut_buffer *nut_buffer_reduce(nut_buffer *buffer, double percentage) {
    assert(buffer != NULL);
    if (percentage < 0.0) percentage = 0.0;
    if (percentage > 1.0) percentage = 1.0;
    int kept = (int)round(buffer->length * percentage);
    return make_buffer(buffer->type, kept, buffer->channels, buffer->data.u8);
}

/* `offset` advances by payload elements, not by frames -- the reference's
 * behaviour (src/nut.c:73-84), kept because scripts depend on it. */
nut_buffer *nut_buffer_clip(nut_buffer *buffer, int offset, int length) {
    assert(buffer != NULL);
    assert(length < 0 || buffer->length - offset >= length);
    int avail = buffer->length - offset;
    int n = (length < 0 || length > avail) ? avail : length;
    const uint8_t *from = buffer->data.u8 + (size_t)offset * elem_size(buffer->type);
    return make_buffer(buffer->type, n, buffer->channels, from);
}

void nut_buffer_set_data(nut_buffer *dst, nut_buffer *src) {
    assert(dst != NULL && src != NULL);
    assert(dst->type == src->type);
    assert(dst->size_bytes == src->size_bytes);
    memcpy(dst->data.u8, src->data.u8, (size_t)dst->size_bytes);
}

/* grow dst by src's payload (src/nut.c:97-119) */
void nut_buffer_append(nut_buffer *dst, nut_buffer *src) {
    assert(dst != NULL && src != NULL);
    assert(dst->type == src->type);
    size_t es = elem_size(dst->type);
    size_t n_dst = (size_t)dst->length * dst->channels;
    size_t n_src = (size_t)src->length * src->channels;
    uint8_t *grown = (uint8_t *)calloc(n_dst + n_src ? n_dst + n_src : 1, es);
    memcpy(grown, dst->data.u8, (size_t)dst->size_bytes);
    memcpy(grown + n_dst * es, src->data.u8, (size_t)src->size_bytes);
    free(dst->data.u8);
    dst->data.u8 = grown;
    dst->size_bytes = (int)((n_dst + n_src) * es);
    dst->length += src->length;
}

uint8_t nut_buffer_get_u8(nut_buffer *buffer, int offset) {
    return buffer->type == NUT_BUFFER_U8 ? buffer->data.u8[offset]
                                         : (uint8_t)(buffer->data.f64[offset] * 256.0);
}

double nut_buffer_get_f64(nut_buffer *buffer, int offset) {
    return buffer->type == NUT_BUFFER_U8 ? buffer->data.u8[offset] / 256.0
                                         : buffer->data.f64[offset];
}

void nut_buffer_set_u8(nut_buffer *buffer, int offset, uint8_t value) {
    if (buffer->type == NUT_BUFFER_U8) buffer->data.u8[offset] = value;
    else buffer->data.f64[offset] = value / 256.0;
}

void nut_buffer_set_f64(nut_buffer *buffer, int offset, double value) {
    if (buffer->type == NUT_BUFFER_U8) buffer->data.u8[offset] = (uint8_t)(value * 256.0);
    else buffer->data.f64[offset] = value;
}

nut_buffer *nut_buffer_convert(nut_buffer *buffer, nut_buffer_type new_type) {
    assert(buffer != NULL);
    nut_buffer *out = make_buffer(new_type, buffer->length, buffer->channels, NULL);
    int count = buffer->length * buffer->channels;
    for (int i = 0; i < count; i++) {
        if (new_type == NUT_BUFFER_U8) out->data.u8[i] = nut_buffer_get_u8(buffer, i);
        else out->data.f64[i] = nut_buffer_get_f64(buffer, i);
    }
    return out;
}

void nut_buffer_save(nut_buffer *buffer, const char *fname) {
    assert(buffer != NULL);
    FILE *fp = fopen(fname, "wb");
    if (!fp) return;
    fwrite(buffer->data.u8, (size_t)buffer->size_bytes, 1, fp);
    fclose(fp);
    printf("Written %s.\n", fname);
}
