/*
 * fsea.h -- C ABI of the B200 IQ -> spectrum engine (libfsea.so).
 *
 * This is the drop-in boundary below the reference's operator API: the
 * reference's nrf_fft_* block (src/nrf.h:128-142, src/nrf.c:557-642) and its
 * stand-alone sweep tools (c/fft-batch.c:54-102, c/fft-batch-broad.c:56-129,
 * c/fft.c:45-56,181-197) call FFTW3 once per frame on the CPU
 * (fftw_plan_dft_1d src/nrf.c:564, fftw_execute src/nrf.c:615); here the same
 * per-frame work -- 8-bit IQ unpack, optional Hann window, power-of-two
 * forward C2C FFT, fftshift, magnitude / log-magnitude / u8 pixel -- runs for
 * a whole batch of frames in hand-written sm_100a kernels.  include/nrf.h
 * (the reference's own operator names) is a one-frame client of this API.
 *
 * Threading: a plan is a single-threaded object (one host thread at a time; its
 * launches may go to any streams); different plans are independent.
 *
 * Plain C: pointers and sizes only.  Every function returns FSEA_OK (0) or a
 * negative FSEA_ERR_* code; fsea_strerror() names it.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with
 * FSEA_ERR_NO_DEVICE.
 *
 * Per-frame definition (frame f starts at complex sample f*hop; b = raw bytes):
 *   t(b)   = b                     FSEA_IN_U8  (offset-binary: RTL-SDR, nut_buffer u8)
 *          = (b + 128) mod 256     FSEA_IN_S8  (two's-complement HackRF / rfdata raw files;
 *                                               fuses src/nrf.c:102-105)
 *   x[n]   = (-1)^n * (t(b[2n]) + i t(b[2n+1])) / 256           window NONE  (src/nrf.c:605-612)
 *          = (-1)^n * w[n] * ((t(b[2n])-128) + i (t(b[2n+1])-128)) / 256     window HANN,
 *            w[n] = 0.5 (1 - cos(2 pi n / N))   (extension; no reference pin)
 *          = (-1)^n * w[n] * sample[n]                          FSEA_IN_F32C / FSEA_IN_F64C
 *   X[k]   = sum_n x[n] exp(-2 pi i n k / N)                    (k = N/2 is DC)
 *   out[k] = |X[k]|                                             FSEA_OUT_MAG    (src/nrf.c:624-628)
 *          = 10 log10(|X[k]|^2 + 1e-20)                         FSEA_OUT_LOGMAG (c/fft-batch.c:88-89)
 *            generally log_mult * log10(power_gain |X[k]|^2 + 1e-20): c/fft.c:181-197's scope row is
 *            FSEA_IN_U8, log_mult = intensity, power_gain = (256 / (255 N))^2
 *          = clamp_u8(trunc(pix_scale * 10 log10(|X[k]|^2 + 1e-20)))   FSEA_OUT_PIX (c/fft-batch.c:89-92)
 *   FSEA_FLAG_DC_PATCH:  out[N/2] = out[N/2-1]   (src/nrf.c:621-622, c/fft-batch-broad.c:115-117)
 * Output element type: float32 (MAG/LOGMAG), uint8 (PIX), float64 for
 * FSEA_IN_F64C and for FSEA_ARITH_F64 (computed in fp64).  Output row f holds N bins.
 * The fp32 kernels reproduce the reference's float64 spectrum to ~1e-6 of the row's rms; a pixel whose
 * scaled dB value lies that close to an integer can therefore come out 1 LSB off (measured: < 2e-4 of all
 * pixels, tests/test_gpu_parity.py).  FSEA_ARITH_F64 is the mode without that caveat.
 */
#ifndef FSEA_H
#define FSEA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSEA_VERSION 1

enum { FSEA_IN_U8 = 0, FSEA_IN_S8 = 1, FSEA_IN_F32C = 2, FSEA_IN_F64C = 3 };
enum { FSEA_WIN_NONE = 0, FSEA_WIN_HANN = 1 };
enum { FSEA_OUT_MAG = 0, FSEA_OUT_LOGMAG = 1, FSEA_OUT_PIX = 2 };
enum { FSEA_FLAG_DC_PATCH = 1 };
/* arithmetic variant of the fp32 kernels (results agree to fp32 rounding) */
enum { FSEA_ARITH_AUTO = 0, FSEA_ARITH_F32 = 1, FSEA_ARITH_F32X2 = 2,
       FSEA_ARITH_F32X2_FRAME = 3 /* packed lanes = even / odd samples of ONE frame (N = 2048, 4096, 8192; 8-bit input) */,
       FSEA_ARITH_F32X2_RADIX64 = 4 /* N = 4096 = 64 x 64, 64 points per thread, one exchange (8-bit input, no window) */,
       FSEA_ARITH_F64 = 5 /* 8-bit input transformed in float64 like the reference (src/nrf.c:599-615): samples widened
                             (and windowed) in double on the device, MAG / LOGMAG rows are float64, FSEA_OUT_PIX pixels
                             equal the reference's byte for byte.  Any N; several times slower than the fp32 kernels. */ };

enum {
    FSEA_OK = 0,
    FSEA_ERR_INVALID = -1,      /* bad argument / unsupported combination */
    FSEA_ERR_UNSUPPORTED = -2,  /* fft_size not a power of two in [16, 65536] */
    FSEA_ERR_NO_DEVICE = -3,    /* no CUDA device: there is no CPU fallback */
    FSEA_ERR_CUDA = -4,         /* CUDA runtime error, see fsea_last_cuda_error() */
    FSEA_ERR_NOMEM = -5
};

typedef struct {
    int fft_size;      /* N: power of two, 16..65536 */
    int hop;           /* complex samples between frame starts; 0 = N (no overlap); must be a multiple of 8 for 8-bit input */
    int in_format;     /* FSEA_IN_* */
    int window;        /* FSEA_WIN_* */
    int out_mode;      /* FSEA_OUT_* */
    int flags;         /* FSEA_FLAG_* */
    float pix_scale;   /* S for FSEA_OUT_PIX (c/fft-batch.c: 10, c/fft-batch-broad.c: 5) */
    int device;        /* CUDA device ordinal */
    int arith;         /* FSEA_ARITH_* */
    float log_mult;    /* FSEA_OUT_LOGMAG / PIX: out = log_mult * log10(power_gain * |X|^2 + 1e-20); 0 = 10 */
    float power_gain;  /* 0 = 1.  c/fft.c:50-51,191-196 (the 512-point scope): FSEA_IN_U8, log_mult = intensity,
                          power_gain = (256 / (255 N))^2  (samples / 255 instead of / 256, spectrum / N) */
    int reserved[5];
} fsea_config;

typedef struct fsea_plan fsea_plan;

/* number of CUDA devices visible (0 when none: nothing else will work) */
int fsea_device_count(void);
const char *fsea_strerror(int code);
/* text of the last CUDA error seen by this thread's calls ("" if none) */
const char *fsea_last_cuda_error(void);

void fsea_config_default(fsea_config *cfg, int fft_size);
int fsea_plan_create(const fsea_config *cfg, fsea_plan **plan);
int fsea_plan_destroy(fsea_plan *plan);

/* bytes of input needed for n_frames, and bytes of output produced */
size_t fsea_input_bytes(const fsea_plan *plan, long n_frames);
size_t fsea_output_bytes(const fsea_plan *plan, long n_frames);
/* frames that fit in `n_samples` complex samples */
long fsea_frames_in(const fsea_plan *plan, size_t n_samples);

/* Device-resident batch: d_in / d_out are device pointers on plan's device,
 * d_in 16-byte aligned.  Asynchronous on `stream` (a cudaStream_t, NULL =
 * default stream).  This is what bench.py's `value` times. */
int fsea_exec_device(fsea_plan *plan, const void *d_in, void *d_out, long n_frames, void *stream);

/* The same, with the output rows out_row_stride_bytes apart (>= N * element size, a multiple of 4):
 * rows that land inside a wider picture -- frame f of a sweep frequency goes to row f of the
 * stitched canvas (c/fft-stitch-broad.c:28-36), possibly in a PEER device's memory (the stores of
 * the transform kernel then cross NVLink and no gather or stitch pass is left). */
int fsea_exec_device_strided(fsea_plan *plan, const void *d_in, void *d_out, long out_row_stride_bytes,
                             long n_frames, void *stream);

/* Host buffers: chunked, double-buffered H2D -> kernel -> D2H through pinned
 * staging on three streams (page-locked caller memory is used directly); returns when h_out is complete.  This is the call
 * the nrf_fft shim and bench.py's `e2e` use. */
int fsea_exec_host(fsea_plan *plan, const void *h_in, void *h_out, long n_frames);

/* fsea_exec_host's chunks over the same streams with the kernel left out: what the box's PCIe
 * links allow for this job (bench.py reports e2e as a fraction of it).  h_out receives stale bytes. */
int fsea_copy_only_host(fsea_plan *plan, const void *h_in, void *h_out, long n_frames);

/* Pin the calling thread (and so the first-touch placement of the pinned buffers it allocates
 * next) to the CPUs of the NUMA node `device` hangs off (sysfs local_cpulist of its PCI function).
 * Returns the number of CPUs bound to, 0 if the topology is unknown (nothing changed), < 0 on error. */
int fsea_host_bind_near_device(int device);

/* The same pipeline fed from a capture file: the raw bytes an SDR delivered, block after block,
 * no header (c/rfcap.c:26-41, c/reader.c:25-49; replayed by src/nrf.c:256-284).  Frames
 * [0, n_frames) start at byte_offset (a multiple of 16) + f * hop * 2; pread() fills the pinned
 * slots while earlier slots cross PCIe and compute.  8-bit input formats only.
 * FSEA_ERR_INVALID when the file is missing or shorter than the frames asked for. */
int fsea_exec_file(fsea_plan *plan, const char *path, long byte_offset, long n_frames, void *h_out);

/* Number of kernel launches issued by this plan so far (bench gpu_launches). */
long fsea_plan_launch_count(const fsea_plan *plan);
/* Name of the kernel variant the plan dispatches to, e.g. "fft<1024:32x32x1,f32x2>". */
const char *fsea_plan_kernel_name(const fsea_plan *plan);

/* ---- sweep post-processing -------------------------- */

/* Mean of |X| over the first `rows` frames of a PIX/MAG batch -- the
 * "interesting?" gate of c/fft-batch-broad.c:80-98 -- computed on the device
 * from raw input.  Writes one double. */
int fsea_gate_mean_host(fsea_plan *plan, const void *h_in, long rows, double *mean_out);

/* Max-blend the per-frequency pixel images of a sweep side by side
 * (c/fft-stitch-broad.c:28-36,66-88; c/fft-stitch.c:46-54,167-189): image i
 * (rows x width, uint8, row-major, images contiguous) lands at column
 * i * width_step of an output of width  width + (n_images-1) * width_step,
 * out = max(out, image), zero where no image lands. */
int fsea_stitch_device(const uint8_t *d_images, int n_images, int rows, int width, int width_step,
                       uint8_t *d_out, void *stream);
/* The same from the layout a sharded sweep gathers (frequency f -> rank f mod world):
 * d_images = [world][per_rank][rows][width], frequency i = image i / world of rank i % world. */
int fsea_stitch_sharded_device(const uint8_t *d_images, int n_images, int world, int per_rank, int rows, int width,
                               int width_step, uint8_t *d_out, void *stream);
int fsea_stitch_host(int device, const uint8_t *h_images, int n_images, int rows, int width, int width_step,
                     uint8_t *h_out);

/* ---- the multi-centre-frequency sweep as a device-resident engine (BASELINE cfg4) ----------
 * What c/fft-batch-broad.c:56-129,189-209 does per centre frequency -- rows of an N-point
 * spectrum; after gate_rows rows the mean |X| decides whether anything is there (:80-98); if so
 * `rows` rows of u8 pixels (S = pix_scale, DC patch, :108-118), one image per frequency -- for the
 * `capacity` frequencies of one GPU, with nothing crossing the host between raw IQ and images:
 *   submit   raw IQ of a frequency (rows * N complex int8 samples) becomes resident, once
 *   gate     the gate rows of everything submitted are transformed and summed (deterministically,
 *            as fsea_gate_mean_host); means and keep flags in one small D2H, the job's only sync
 *   render   pixel rows of the kept frequencies, zeros for the others, into the engine's
 *            images [slot][rows][N] or straight into a stitched canvas (row stride = canvas width,
 *            possibly peer memory: c/fft-stitch-broad.c:28-36,66-88 with WIDTH_STEP = N)
 * No allocation after create. */
typedef struct fsea_sweep fsea_sweep;
typedef struct {
    int fft_size;          /* N */
    int rows;              /* HEIGHT, c/fft-batch-broad.c:15 */
    int gate_rows;         /* 100, :80 (clamped to rows) */
    double gate_threshold; /* 1.1, :93 */
    float pix_scale;       /* 5, :112-113 */
    int in_format;         /* FSEA_IN_S8 (HackRF captures) or FSEA_IN_U8 */
    int device;
    int capacity;          /* frequencies held at once */
    int own_raw;           /* 1: the engine keeps the raw IQ (needed by submit_host / submit_synth) */
    int own_images;        /* 1: the engine keeps images [capacity][rows][N] */
    int pixel_arith;       /* arithmetic of the pixel pass: FSEA_ARITH_AUTO (fp32 kernels; < 2e-4 of the pixels 1 LSB off the
                              reference's) or FSEA_ARITH_F64 (the reference's picture byte for byte, ~10 x slower) */
    int reserved[3];
} fsea_sweep_config;
int fsea_sweep_create(const fsea_sweep_config *cfg, fsea_sweep **out);
int fsea_sweep_destroy(fsea_sweep *s);
/* raw IQ of slot `slot`: from host memory (H2D on the engine's stream), from device memory (used in
 * place: must stay valid until render), or the SURVEY 8(d) cfg4 synthetic source of centre
 * frequency freq_idx drawn on the device (noise + one tone whose bin depends on freq_idx; every
 * quiet_every-th frequency noise-only below the gate).  Asynchronous. */
int fsea_sweep_submit_host(fsea_sweep *s, int slot, const void *h_raw);
int fsea_sweep_submit_device(fsea_sweep *s, int slot, const void *d_raw);
int fsea_sweep_submit_synth(fsea_sweep *s, int slot, int freq_idx, int quiet_every);
/* gate decisions of slots [0, n_slots): means[n_slots], keep[n_slots] (host); synchronises the engine's stream */
int fsea_sweep_gate(fsea_sweep *s, int n_slots, double *means, uint8_t *keep);
/* the same in two halves (one host thread driving several devices): begin queues the work, end waits for it */
int fsea_sweep_gate_begin(fsea_sweep *s, int n_slots);
int fsea_sweep_gate_end(fsea_sweep *s, int n_slots, double *means, uint8_t *keep);
/* pixel rows of slots [0, n_slots).  d_canvas == NULL: into the engine's images.  Otherwise image of
 * slot j goes to d_canvas + canvas_col0 + j * canvas_col_step, rows canvas_stride bytes apart.  Asynchronous. */
int fsea_sweep_render(fsea_sweep *s, int n_slots, uint8_t *d_canvas, long canvas_stride, long canvas_col0, long canvas_col_step);
int fsea_sweep_sync(fsea_sweep *s);
/* back on the host (tests, small jobs): images of slots [first, first + count) (own_images); raw IQ of a slot */
int fsea_sweep_read_images(fsea_sweep *s, int first, int count, uint8_t *h_out);
int fsea_sweep_read_raw(fsea_sweep *s, int slot, uint8_t *h_out);
uint8_t *fsea_sweep_images(fsea_sweep *s);        /* device pointer, capacity x rows x N (own_images) */
const void *fsea_sweep_raw(fsea_sweep *s, int slot);   /* device pointer of the slot's raw IQ */
void *fsea_sweep_stream(fsea_sweep *s);           /* the engine's cudaStream_t */
long fsea_sweep_launch_count(const fsea_sweep *s);

/* ---- the operators in front of the path, fused with it ----------------
 * One device block (block_samples complex u8 samples, as nut buffers hold them) -> [nrf_freq_shifter, src/nrf.c:843-866]
 * -> [nrf_iq_filter, src/nrf.c:738-768] -> nrf_fft's row (first fft_size samples, |X|, DC patch, float64) in ONE call on
 * the device: no float64 nut_buffer of the whole block crosses the host between the operators, and the filter computes
 * the fft_size outputs the transform reads instead of all block_samples.  Rotator phase and the filter's taps - 1
 * carried-over samples persist from block to block as in the reference.  The filter's sums are bit-identical to the
 * reference's (fsea_fir_iq_device); coefficients come from the caller (nrf_fir_get_low_pass_coefficients). */
typedef struct fsea_chain fsea_chain;
typedef struct {
    int fft_size;                 /* N */
    int block_samples;            /* complex samples per device block: NRF_SAMPLES_LENGTH = 131072 */
    int shifter;                  /* 1: rotator in front */
    int freq_offset;              /* nrf_freq_shifter_new(freq_offset, sample_rate) */
    int shifter_sample_rate;
    int filter;                   /* 1: FIR low-pass behind the rotator */
    int taps;                     /* coefficients actually used (nrf_fir_filter::length) */
    const double *coefficients;   /* taps doubles (copied) */
    int device;
    int reserved[4];
} fsea_chain_config;
int fsea_chain_create(const fsea_chain_config *cfg, fsea_chain **out);
int fsea_chain_destroy(fsea_chain *chain);
/* h_block: 2 * block_samples bytes (interleaved I, Q, offset-binary u8); h_row: fft_size doubles.  Synchronous. */
int fsea_chain_process_u8(fsea_chain *chain, const uint8_t *h_block, double *h_row);
long fsea_chain_launch_count(const fsea_chain *chain);

/* ---- the history image on the device ----------------
 * nrf_fft's N x H image (src/nrf.c:617,633-635) as a ring of float64 rows in device memory: put = one new row
 * (a float32 or float64 spectrum row already on the device) written in place of the oldest; shift = the retune
 * scroll of every row by s bins, zero fill (src/nrf.c:569-596), from d_src into d_dst (two buffers, swapped by the
 * caller).  include/nrf.h's nrf_fft_* keep their history this way and copy to the host only in nrf_fft_get_buffer. */
int fsea_history_put_device(const void *d_row, int row_is_f64, double *d_ring_row, int n, void *stream);
int fsea_history_shift_device(const double *d_src, double *d_dst, int n, int rows, int s, void *stream);

/* Device memory the other processes of the box can store into (one process per GPU: the stitched canvas of a
 * sweep lives on one rank, the others' transform kernels write their pixel rows into it over NVLink).  alloc:
 * cudaMalloc + the 64-byte CUDA IPC handle to pass around; open: a pointer valid for kernels of `device` in the
 * calling process; close / free when done.  FSEA_ERR_UNSUPPORTED where there is no CUDA IPC. */
int fsea_ipc_alloc(int device, size_t bytes, void **d_ptr, unsigned char handle[64]);
int fsea_ipc_open(int device, const unsigned char handle[64], void **d_ptr);
int fsea_ipc_close(void *d_ptr);
int fsea_ipc_free(void *d_ptr);

/* ---- IQ low-pass in front of the path ---------------- */

/* The FIR evaluation of nrf_iq_filter_get_buffer (src/nrf.c:718-724,759-768):
 *   out[2i] = sum_k coef[k] si[i+k],  out[2i+1] = sum_k coef[k] sq[i+k],  k = 0 .. taps-1,
 * si / sq = the I and Q sample lines of length + taps - 1 doubles (taps-1
 * carried-over samples followed by the new block), out = length interleaved
 * (I, Q) doubles.  float64 with the reference's operation order and no
 * contraction: bit-identical to the reference.  d_out 16-byte aligned. */
int fsea_fir_iq_device(const double *d_coef, int taps, const double *d_si, const double *d_sq, long length,
                       double *d_out, void *stream);
int fsea_fir_iq_host(int device, const double *coef, int taps, const double *si, const double *sq, long length,
                     double *out);

#ifdef __cplusplus
}
#endif
#endif /* FSEA_H */
