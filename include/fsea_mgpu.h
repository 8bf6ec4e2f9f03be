/*
 * fsea_mgpu.h -- the sweep job on the GPUs of one box, ONE process (libfsea_mgpu.so).
 *
 * The reference sweeps its centre frequencies one after the other on one radio and one core
 * (c/fft-batch-broad.c:183-215: retune, 4096 rows, next frequency) and stitches the PNGs later
 * (c/fft-stitch-broad.c:66-88).  Frequencies are independent: here frequency f
 * belongs to GPU f mod G, every GPU runs a device-resident engine (fsea_sweep_*, fsea.h), and NCCL
 * over NVLink only carries finished u8 images to the GPU that owns the stitched picture -- or,
 * where the images tile the picture (WIDTH_STEP = N), nothing at all: the transform kernels of
 * every GPU store their pixel rows straight into device 0's canvas (peer stores).
 *
 * One process, G devices: ncclCommInitAll, one stream per device, grouped ncclSend / ncclRecv.
 * (bench.py under torchrun runs the same engines one rank per GPU over torch.distributed.)
 */
#ifndef FSEA_MGPU_H
#define FSEA_MGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FSEA_MGPU_GATHER = 0,   /* NCCL: images -> device 0, max-blend stitch there */
       FSEA_MGPU_P2P = 1,      /* peer stores into device 0's canvas (needs width_step == fft_size) */
       FSEA_MGPU_LOCAL = 2 };  /* compute only: images stay where they were made (8(e) number (i)) */

typedef struct {
    int n_gpus;            /* devices 0 .. n_gpus-1 */
    int fft_size, rows;    /* N, HEIGHT */
    int n_freqs;           /* centre frequencies; f -> device f mod n_gpus */
    int gate_rows;         /* c/fft-batch-broad.c:80 */
    double gate_threshold; /* :93 */
    float pix_scale;       /* :112-113 */
    int width_step;        /* WIDTH_STEP of the stitch (c/fft-stitch-broad.c:57) */
    int quiet_every;       /* synthetic source: every k-th frequency is noise-only */
    int mode;              /* FSEA_MGPU_* */
    int reserved[6];
} fsea_mgpu_sweep_config;

typedef struct {
    double stage_s;        /* untimed part: drawing the raw IQ on the devices */
    double gate_s;         /* gate rows + decisions (device time, max over devices) */
    double render_s;       /* pixel rows of the kept frequencies (max over devices) */
    double exchange_s;     /* NCCL gather (GATHER) -- 0 for P2P / LOCAL */
    double stitch_s;       /* max-blend on device 0 (GATHER) */
    double total_s;        /* gate .. stitched picture complete, wall clock around device syncs */
    long rows_rendered;    /* kept frequencies x rows */
    long kept;
    long launches;         /* kernels of ours launched in the timed region */
    double reserved[4];
} fsea_mgpu_sweep_timing;

/* Runs the whole job on synthetic per-frequency IQ drawn on the devices.  means / keep: n_freqs
 * entries, frequency order (may be NULL).  h_stitched: rows x (N + (n_freqs-1) width_step) bytes on the
 * host, or NULL to leave the picture on device 0 (it is freed before return).  Returns FSEA_OK or an
 * FSEA_ERR_* code (fsea.h); NCCL failures are reported as FSEA_ERR_CUDA with the text in fsea_mgpu_last_error(). */
int fsea_mgpu_sweep(const fsea_mgpu_sweep_config *cfg, uint8_t *h_stitched, double *means, uint8_t *keep,
                    fsea_mgpu_sweep_timing *timing);

/* NVLink ingest of one root (SURVEY 8(e) number (iii)): every device 1..G-1 sends bytes_per_gpu to
 * device 0 with grouped ncclSend / ncclRecv; seconds of device time on the root. */
int fsea_mgpu_gather_bench(int n_gpus, long bytes_per_gpu, int repeats, double *seconds_out);

const char *fsea_mgpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
