/* hackrf.h -- stand-in for absent libhackrf: every open fails, so the
 * reference falls through to its file-replay ("dummy") device
 * (src/nrf.c:308-318).  TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef ORACLE_STUB_HACKRF_H
#define ORACLE_STUB_HACKRF_H
#include <stdint.h>
typedef struct hackrf_device hackrf_device;
typedef struct {
    hackrf_device *device;
    uint8_t *buffer;
    int buffer_length;
    int valid_length;
    void *rx_ctx;
    void *tx_ctx;
} hackrf_transfer;
typedef int (*hackrf_sample_block_cb_fn)(hackrf_transfer *transfer);
int hackrf_init(void);
int hackrf_exit(void);
int hackrf_open(hackrf_device **device);
int hackrf_close(hackrf_device *device);
int hackrf_set_freq(hackrf_device *device, uint64_t freq_hz);
int hackrf_set_sample_rate(hackrf_device *device, double freq_hz);
int hackrf_set_amp_enable(hackrf_device *device, uint8_t value);
int hackrf_set_lna_gain(hackrf_device *device, uint32_t value);
int hackrf_set_vga_gain(hackrf_device *device, uint32_t value);
int hackrf_start_rx(hackrf_device *device, hackrf_sample_block_cb_fn callback, void *rx_ctx);
int hackrf_stop_rx(hackrf_device *device);
#endif
