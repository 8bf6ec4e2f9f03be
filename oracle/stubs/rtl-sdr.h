/* rtl-sdr.h -- stand-in for absent librtlsdr: open always fails (no hardware).
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef ORACLE_STUB_RTLSDR_H
#define ORACLE_STUB_RTLSDR_H
#include <stdint.h>
typedef struct rtlsdr_dev rtlsdr_dev_t;
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index);
int rtlsdr_close(rtlsdr_dev_t *dev);
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq);
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate);
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual);
int rtlsdr_set_agc_mode(rtlsdr_dev_t *dev, int on);
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev);
int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read);
#endif
