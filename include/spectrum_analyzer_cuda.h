/*
 * spectrum_analyzer_cuda.h -- C ABI of libspectrum_analyzer_cuda.so
 *
 * B200 (sm_100a) implementation of the one data-parallel hot path of
 * phip1611/spectrum-analyzer v1.8.0:
 *
 *     window -> real FFT -> magnitude -> FrequencyLimit -> statistics -> scaling
 *
 * i.e. `samples_fft_to_spectrum` (reference src/lib.rs:157-209) with the caller-side
 * window step (src/windows.rs) fused in, plus the batched `[frames x N]` form that
 * actually feeds a GPU.  Every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only; no exceptions cross this boundary; there
 * is no CPU fallback: compute entry points return SA_ERR_CUDA when no usable device
 * is present.
 *
 * Conventions
 *  - "device pointer" arguments are CUDA device addresses owned by the caller.
 *  - a context (sa_ctx) binds one GPU + one stream; it is not internally
 *    synchronised: use one context per host thread.  Calls that take device
 *    pointers are asynchronous on the context's stream (sa_ctx_sync to wait).
 *  - return value: SA_OK (0) or one of the SA_ERR_* codes.  Codes 1..11 mirror
 *    `SpectrumAnalyzerError` (src/error.rs:37-54) and the reference's two
 *    deliberate panics; codes >= 100 are ABI/CUDA errors.
 */
#ifndef SPECTRUM_ANALYZER_CUDA_H
#define SPECTRUM_ANALYZER_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_ABI_VERSION 1u

/* ---- enums --------------------------------------------------------------- */

/* windows::{hann_window, hamming_window, blackman_harris_4term, blackman_harris_7term}
 * (src/windows.rs:40,58,79,97).  NONE = caller already applied a window. */
typedef enum {
  SA_WINDOW_NONE = 0, SA_WINDOW_HANN = 1, SA_WINDOW_HAMMING = 2,
  SA_WINDOW_BLACKMAN_HARRIS_4TERM = 3, SA_WINDOW_BLACKMAN_HARRIS_7TERM = 4
} sa_window_kind;

/* FrequencyLimit::{All, Min(f), Max(f), Range(lo,hi)} (src/limit.rs:38-53).
 * Min uses `limit_lo`, Max uses `limit_hi`, Range uses both. */
typedef enum { SA_LIMIT_ALL = 0, SA_LIMIT_MIN = 1, SA_LIMIT_MAX = 2, SA_LIMIT_RANGE = 3 } sa_limit_kind;

/* Option<&SpectrumScalingFunction> restricted to the crate's built-ins
 * (src/scaling.rs:104-162).  Arbitrary closures cannot run on the device and are
 * rejected by construction: there is no function-pointer argument. */
typedef enum {
  SA_SCALING_NONE = 0,             /* None */
  SA_SCALING_DIVIDE_BY_N = 1,      /* scaling::divide_by_N        src/scaling.rs:134 */
  SA_SCALING_DIVIDE_BY_N_SQRT = 2, /* scaling::divide_by_N_sqrt   src/scaling.rs:152 */
  SA_SCALING_ZERO_TO_ONE = 3,      /* scaling::scale_to_zero_to_one  src/scaling.rs:119 */
  SA_SCALING_20_TIMES_LOG10 = 4    /* scaling::scale_20_times_log10  src/scaling.rs:104 */
} sa_scaling_kind;

typedef enum {
  SA_OK = 0,
  /* SpectrumAnalyzerError (src/error.rs:37-54), precedence of src/lib.rs:164-181,317 */
  SA_ERR_TOO_FEW_SAMPLES = 1,
  SA_ERR_NAN_VALUES = 2,            /* also a per-frame status */
  SA_ERR_INFINITY_VALUES = 3,       /* also a per-frame status */
  SA_ERR_NOT_POWER_OF_TWO = 4,
  SA_ERR_LIMIT_BELOW_MINIMUM = 5,   /* InvalidFrequencyLimit(ValueBelowMinimum) */
  SA_ERR_LIMIT_ABOVE_NYQUIST = 6,   /* InvalidFrequencyLimit(ValueAboveNyquist) */
  SA_ERR_LIMIT_INVALID_RANGE = 7,   /* InvalidFrequencyLimit(InvalidRange) */
  SA_ERR_LIMIT_TOO_NARROW = 8,      /* FrequencyLimitTooNarrow */
  SA_ERR_SCALING = 9,               /* ScalingError (src/spectrum.rs:150-160): per-frame status of sa_apply_scaling_*;
                                     * unreachable in the fused spectrum calls, whose inputs are magnitudes >= 0 */
  SA_ERR_UNSUPPORTED_LENGTH = 10,   /* reference panics (src/fft.rs:52): N > 32768 */
  SA_ERR_NON_FINITE_RESULT = 11,    /* reference panics (src/frequency.rs:52-53); per-frame status */
  /* ABI / CUDA */
  SA_ERR_INVALID_ARGUMENT = 100,
  SA_ERR_CUDA = 101,                /* see sa_last_error_string */
  SA_ERR_OUT_OF_MEMORY = 102
} sa_status;

/* Which per-frame statistics to compute (FrequencySpectrum::{min,max,average,median},
 * src/spectrum.rs:173-208).  Statistics describe the SCALED values, as the getters of
 * the reference do.  Unrequested fields are written as 0. */
#define SA_STATS_MINMAX  1u
#define SA_STATS_AVERAGE 2u
#define SA_STATS_MEDIAN  4u
#define SA_STATS_ALL     7u

/* One 32-byte record per frame.  Replaces the scalar fields of `FrequencySpectrum`
 * (src/spectrum.rs:48-75): min/max are (bin, value) pairs -- the frequency of a bin
 * is `bin as f32 * resolution` (sa_bin_frequencies).  Tie rules are those of the
 * reference's stable sort: min -> lowest bin, max -> highest bin among equals. */
typedef struct {
  float    min_val;
  uint32_t min_bin;   /* absolute FFT bin index (0..N/2) */
  float    max_val;
  uint32_t max_bin;
  float    average;
  float    median;
  uint32_t status;    /* SA_OK, SA_ERR_NAN_VALUES, SA_ERR_INFINITY_VALUES, SA_ERR_NON_FINITE_RESULT, SA_ERR_SCALING */
  uint32_t reserved;  /* 0, except with SA_ERR_SCALING: row index of the offending value | SA_SCALING_ERR_* class */
} sa_frame_stats;

/* `reserved` of a frame whose status is SA_ERR_SCALING: the low 16 bits hold the row index i of the first value
 * whose scaled image was not finite (the bin is bin_lo + i); one class bit says what the image was, so that a
 * caller can rebuild `ScalingError(orig, new)` (src/error.rs:50-53): orig = vals[i] (left as it was), new = NaN,
 * +Inf or -Inf. */
#define SA_SCALING_ERR_INDEX_MASK 0x0000ffffu
#define SA_SCALING_ERR_NAN        0x80000000u
#define SA_SCALING_ERR_POS_INF    0x40000000u
#define SA_SCALING_ERR_NEG_INF    0x20000000u

typedef struct sa_ctx sa_ctx;

/* ---- library / context ---------------------------------------------------- */

uint32_t    sa_abi_version(void);
const char *sa_status_string(int status);          /* Display of src/error.rs:56-78 */
const char *sa_last_error_string(void);            /* detail of the last SA_ERR_CUDA on this thread */

/* Binds `device`; creates an owned non-blocking stream. */
int sa_ctx_create(int device, sa_ctx **out_ctx);
/* Same, but launches on a caller-owned `cudaStream_t` (e.g. torch's current stream). */
int sa_ctx_create_on_stream(int device, void *cuda_stream, sa_ctx **out_ctx);
int sa_ctx_destroy(sa_ctx *ctx);
int sa_ctx_sync(sa_ctx *ctx);

/* Opt-in (default off): the fused 20*log10 scaling of the tuned kernels (n >= 64, suitably aligned input) uses
 * the hardware log2 approximation instead of the libm `log10f` operation sequence.  Values differ from
 * `20.0 * libm::log10f(v)` (src/scaling.rs:104-114) by a few ulp of the f32 dB value (measured <= 4; lg2.approx has a relative error of about 2^-22)
 * and are no longer bit-exact functions of the magnitude; 1.5x-1.7x faster in that mode.  The statistics stay exact order statistics of the values written. */
int sa_ctx_set_fast_log10(sa_ctx *ctx, int enable);
/* Number of kernels this context has launched so far (for bench accounting). */
uint64_t sa_ctx_launch_count(const sa_ctx *ctx);
/* Name of the spectrum kernel the last sa_spectrum_* call on this context launched (template arguments included),
 * e.g. "sa::spectrum_kernel_v2<V2Cfg<12,5,0>,0,1>": what profiles and bench lines cite.  No reference counterpart. */
const char *sa_ctx_last_kernel(const sa_ctx *ctx);

/* ---- host-side pieces of the path (no GPU needed) ------------------------- */

/* The window multipliers the reference computes per sample (src/windows.rs:43-47,
 * 66-69, 138-144), bit-exact f32, n in [1, 32768]. */
int sa_window_coeffs(int window_kind, uint32_t n, float *host_out);

/* FrequencyLimit::verify (src/limit.rs:96-118) + the bin filter of
 * fft_result_to_spectrum (src/lib.rs:237-304, 317): the inclusive kept range is
 * bins [bin_lo, bin_lo + n_bins).  Returns the reference's error for this
 * (n, sampling_rate, limit) or SA_OK; n_bins < 2 -> SA_ERR_LIMIT_TOO_NARROW. */
int sa_limit_to_bins(uint32_t n, uint32_t sampling_rate, int limit_kind, float limit_lo,
                     float limit_hi, uint32_t *bin_lo, uint32_t *n_bins);

/* `fft_index as f32 * frequency_resolution` (src/lib.rs:278, 356-358). */
int sa_bin_frequencies(uint32_t n, uint32_t sampling_rate, uint32_t bin_lo, uint32_t n_bins,
                       float *host_out);

/* ---- the hot path ---------------------------------------------------------- */

/*
 * Batched `samples_fft_to_spectrum` (src/lib.rs:157-209) for `n_frames` frames of `n`
 * samples; frame f starts at d_samples + f*frame_stride (frame_stride = n for dense
 * frames, = hop for STFT-style overlap).  Per frame, fused in one kernel:
 *   window (src/windows.rs) -> NaN/Inf validation of the windowed samples
 *   (src/lib.rs:168-173) -> real FFT with microfft's N/2+1-bin convention
 *   (src/fft.rs:72-133) -> magnitude (src/lib.rs:366-372) -> FrequencyLimit slice
 *   (src/lib.rs:286-304) -> scaling (src/spectrum.rs:128-168, src/scaling.rs) ->
 *   statistics of the scaled values (src/spectrum.rs:481-539).
 * Outputs (device memory):
 *   d_out_vals[f*out_stride + i], i in [0, n_bins): value of bin (bin_lo + i)
 *   d_stats[f]: statistics + per-frame status (may be NULL when stats_mask == 0;
 *               then per-frame validation results are not reported)
 * Host outputs: *bin_lo, *n_bins (may be NULL).
 * n must be a power of two in [2, 32768]; out_stride >= n_bins.
 * Alignment: the element type's for d_samples (checked: SA_ERR_INVALID_ARGUMENT otherwise), 4 bytes for
 * d_out_vals, 16 bytes for d_stats.  No byte outside [d_samples, d_samples + (n_frames-1)*frame_stride + n)
 * is read: frames of n >= 1024 samples that do not start on a 16-byte boundary are staged by bulk copies of
 * their 16-byte aligned superset, clamped at the two ends of the batch (the few bytes there are fetched by
 * ordinary loads).  16-byte aligned frames (base pointer and hop) take the fastest path.
 * Fused window (window_kind != SA_WINDOW_NONE): for n >= 1024 the multiply by the window is contracted into the first
 * butterfly additions (fused multiply-add), so the result is not bit-identical to windowing the samples first
 * (sa_window_batched, or the caller's own `hann_window`, as the reference's callers do) and passing SA_WINDOW_NONE;
 * the two differ by a few ulp of the frame maximum, far inside the 1e-4 x frame-maximum tolerance of the path.
 */
int sa_spectrum_batched(sa_ctx *ctx, const float *d_samples, uint64_t n_frames, uint32_t n,
                        uint64_t frame_stride, uint32_t sampling_rate, int limit_kind,
                        float limit_lo, float limit_hi, int window_kind, int scaling_kind,
                        uint32_t stats_mask, float *d_out_vals, uint64_t out_stride,
                        sa_frame_stats *d_stats, uint32_t *bin_lo, uint32_t *n_bins);

/* Same call on int16 PCM: each sample is converted exactly (`x as f32`, as the reference's callers do,
 * src/tests/mod.rs:70-73) when it is loaded, so the conversion never touches memory.  frame_stride counts
 * samples. */
int sa_spectrum_batched_i16(sa_ctx *ctx, const int16_t *d_samples, uint64_t n_frames, uint32_t n,
                            uint64_t frame_stride, uint32_t sampling_rate, int limit_kind,
                            float limit_lo, float limit_hi, int window_kind, int scaling_kind,
                            uint32_t stats_mask, float *d_out_vals, uint64_t out_stride,
                            sa_frame_stats *d_stats, uint32_t *bin_lo, uint32_t *n_bins);

/* Same call with HOST buffers: pinned staging, chunked H2D -> kernel -> D2H pipeline
 * on the context's copy streams; synchronous.  This is the end-to-end form a drop-in
 * caller with host data uses. */
int sa_spectrum_batched_host(sa_ctx *ctx, const float *h_samples, uint64_t n_frames, uint32_t n,
                             uint64_t frame_stride, uint32_t sampling_rate, int limit_kind,
                             float limit_lo, float limit_hi, int window_kind, int scaling_kind,
                             uint32_t stats_mask, float *h_out_vals, uint64_t out_stride,
                             sa_frame_stats *h_stats, uint32_t *bin_lo, uint32_t *n_bins);

int sa_spectrum_batched_host_i16(sa_ctx *ctx, const int16_t *h_samples, uint64_t n_frames, uint32_t n,
                                 uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                                 float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask,
                                 float *h_out_vals, uint64_t out_stride, sa_frame_stats *h_stats,
                                 uint32_t *bin_lo, uint32_t *n_bins);

/* Single-frame drop-in for `samples_fft_to_spectrum(samples, sampling_rate, limit,
 * scaling_fn)` (src/lib.rs:157-162): host in, host out, full error taxonomy with the
 * reference's precedence (n_samples < 2, NaN, Inf, power of two, limit, narrow).
 * h_freqs / h_vals need room for n_samples/2+1 floats; *n_bins receives the count. */
int sa_spectrum_single(sa_ctx *ctx, const float *h_samples, uint64_t n_samples,
                       uint32_t sampling_rate, int limit_kind, float limit_lo, float limit_hi,
                       int window_kind, int scaling_kind, float *h_freqs, float *h_vals,
                       uint32_t *n_bins, sa_frame_stats *h_stats);

/* `FrequencySpectrum::apply_scaling_fn` (src/spectrum.rs:128-168) on a device-resident
 * batch: scales d_vals[f*stride + i] in place with the statistics of the spectrum as it is
 * (samples_len = n; the maximum that scale_to_zero_to_one divides by is taken from the values
 * themselves, so d_stats[f] need not hold min/max), then recomputes d_stats[f] (min/max always,
 * average / median per stats_mask).  This is the repeated scaling of
 * benches/fft_spectrum_bench.rs:21-36.  Error behaviour of the reference, per frame
 * (src/spectrum.rs:150-160): values are replaced in bin order and the FIRST one whose scaled image
 * is NaN or +-Inf stops the frame -- the values before it stay scaled, it and the rest stay as
 * they were, the statistics are not recomputed, d_stats[f].status = SA_ERR_SCALING and
 * d_stats[f].reserved describes the offender (SA_SCALING_ERR_*).  Reachable with the built-ins:
 * scale_20_times_log10 of a negative (already dB) value is NaN.  Frames whose status is not SA_OK
 * on entry hold no spectrum in the reference's terms and are left untouched.  The call itself
 * returns SA_OK; callers check the per-frame status as after sa_spectrum_batched. */
int sa_apply_scaling_batched(sa_ctx *ctx, float *d_vals, uint64_t n_frames, uint32_t n_bins,
                             uint64_t stride, uint32_t bin_lo, uint32_t n, int scaling_kind,
                             uint32_t stats_mask, sa_frame_stats *d_stats);

/* `scaling::combined(&[f1, f2, ...])` (src/scaling.rs:174-182) applied through apply_scaling_fn on a
 * device-resident batch: every function of the chain (at most 8 built-in kinds) sees the SAME
 * statistics -- those of the spectrum before the call -- then d_stats[f] is recomputed once. */
int sa_apply_scaling_chain_batched(sa_ctx *ctx, float *d_vals, uint64_t n_frames, uint32_t n_bins,
                                   uint64_t stride, uint32_t bin_lo, uint32_t n, const int *scaling_kinds,
                                   uint32_t n_kinds, uint32_t stats_mask, sa_frame_stats *d_stats);

/* The same on HOST buffers (in place; one spectrum or a batch): what `FrequencySpectrum::apply_scaling_fn`
 * (src/spectrum.rs:128-168) of a host-side spectrum object calls.  Synchronous. */
int sa_apply_scaling_chain_host(sa_ctx *ctx, float *h_vals, uint64_t n_frames, uint32_t n_bins,
                                uint64_t stride, uint32_t bin_lo, uint32_t n, const int *scaling_kinds,
                                uint32_t n_kinds, uint32_t stats_mask, sa_frame_stats *h_stats);

/* windows::*_window on a device-resident batch: d_out[f*n + i] = w_i * d_in[f*frame_stride + i]. */
int sa_window_batched(sa_ctx *ctx, int window_kind, const float *d_in, uint64_t n_frames, uint32_t n,
                      uint64_t frame_stride, float *d_out);

/* Same with HOST buffers (drop-in for `windows::hann_window(&[f32]) -> Vec<f32>` etc.): H2D, the window
 * kernel, D2H; synchronous.  h_out is dense [n_frames x n]. */
int sa_window_host(sa_ctx *ctx, int window_kind, const float *h_in, uint64_t n_frames, uint32_t n,
                   uint64_t frame_stride, float *h_out);

/* Synthetic input: counter-based uniform noise in [-1,1) written on the device;
 * value i depends only on (seed, first_index + i) so the CPU checker can regenerate it. */
int sa_generate_noise(sa_ctx *ctx, float *d_out, uint64_t first_index, uint64_t count, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRUM_ANALYZER_CUDA_H */
