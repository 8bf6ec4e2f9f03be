/*
 * nut.h -- host-side container the operator API speaks (mirror of the
 * reference's src/nut.h:10-45 so that callers written against the reference
 * -- its Lua binding src/main.cpp:117-179, the nrf_* blocks -- compile and
 * link unchanged).  Struct layout and enum values are ABI: keep them.
 *
 * A nut_buffer is a typed flat array: `length` elements of `channels`
 * interleaved values, element count = length * channels, u8 <-> f64
 * conversion convention is x/256 (src/nut.c:121-151).  Every function that
 * returns a nut_buffer* returns a fresh heap object owned by the caller.
 */
#ifndef NUT_H
#define NUT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type tags (src/nut.h:14-17; exported to Lua, src/main.cpp:1181-1183) */
typedef enum {
    NUT_BUFFER_U8 = 1,
    NUT_BUFFER_F64 = 2
} nut_buffer_type;

typedef union nut_buffer_data {
    uint8_t *u8;
    double *f64;
} nut_buffer_data;

/* src/nut.h:24-30 */
typedef struct {
    nut_buffer_type type;
    int length;       /* elements per channel */
    int channels;
    int size_bytes;   /* length * channels * sizeof(element) */
    nut_buffer_data data;
} nut_buffer;

/* construction / destruction (src/nut.c:25-50,183-190); data may be NULL (zero fill) */
nut_buffer *nut_buffer_new_u8(int length, int channels, const uint8_t *data);
nut_buffer *nut_buffer_new_f64(int length, int channels, const double *data);
void nut_buffer_free(nut_buffer *buffer);

/* copies and views-by-copy (src/nut.c:52-82) */
nut_buffer *nut_buffer_copy(nut_buffer *buffer);
nut_buffer *nut_buffer_reduce(nut_buffer *buffer, double percentage);
nut_buffer *nut_buffer_clip(nut_buffer *buffer, int offset, int length);
nut_buffer *nut_buffer_convert(nut_buffer *buffer, nut_buffer_type new_type);

/* in-place (src/nut.c:84-119) */
void nut_buffer_set_data(nut_buffer *dst, nut_buffer *src);
void nut_buffer_append(nut_buffer *dst, nut_buffer *src);

/* element access with the x/256 convention (src/nut.c:121-151) */
uint8_t nut_buffer_get_u8(nut_buffer *buffer, int offset);
double nut_buffer_get_f64(nut_buffer *buffer, int offset);
void nut_buffer_set_u8(nut_buffer *buffer, int offset, uint8_t value);
void nut_buffer_set_f64(nut_buffer *buffer, int offset, double value);

/* raw dump of the payload (src/nut.c:173-181) */
void nut_buffer_save(nut_buffer *buffer, const char *fname);

/* src/nut.c:18-23 */
void nut_sleep_milliseconds(int millis);

#ifdef __cplusplus
}
#endif
#endif /* NUT_H */
