/*
 * sa_oracle.h -- CPU oracle for the `samples_fft_to_spectrum` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call it, and there only as the
 * checker / the timed CPU baseline.  The CUDA library never links or calls it.
 *
 * What it restates (all paths relative to /root/reference, crate v1.8.0):
 *   src/windows.rs:40-150   hann / hamming / blackman_harris_{4,7}term
 *   src/lib.rs:157-372      samples_fft_to_spectrum, fft_result_to_spectrum
 *   src/fft.rs:72-133       FftImpl::calc (microfft::real adapter, Nyquist unpack)
 *   src/spectrum.rs:91-168, 481-539   FrequencySpectrum::new / apply_scaling_fn /
 *                           calc_statistics (stable sort statistics)
 *   src/scaling.rs:104-162  the four built-in scaling functions
 *   src/limit.rs:96-118     FrequencyLimit::verify
 *   src/frequency.rs:49-56  OrderableF32 finite assertion
 *   src/tests/sine.rs:36-104  the synthetic sine fixture
 * Third-party arithmetic that is NOT under /root/reference and is restated from
 * the published algorithms (see sa_oracle.c for details):
 *   microfft 0.6.0 (Cargo.lock:1041-1050)  real::rfft_N  -- packed real FFT
 *   libm 0.2.16    (Cargo.lock:987-990)    cosf, log10f  -- musl/FreeBSD forms
 *
 * PARITY STATUS: the in-repo arithmetic (windows, bin frequencies, limits,
 * statistics, scaling, error taxonomy) is pinned against every golden vector the
 * reference's own tests hold (tests/test_oracle_golden.py).  The per-bin FFT
 * values are "parity unpinned": the reference stores no per-bin golden spectra
 * and microfft cannot be built here (no Rust toolchain, crate source absent), so
 * the FFT restatement is anchored on the reference's threshold tests and on an
 * independent f64 DFT, not on microfft's bits.
 */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as include/spectrum_analyzer_cuda.h (kept separate on purpose). */
enum { SAO_WINDOW_NONE = 0, SAO_WINDOW_HANN = 1, SAO_WINDOW_HAMMING = 2,
       SAO_WINDOW_BH4 = 3, SAO_WINDOW_BH7 = 4 };
enum { SAO_LIMIT_ALL = 0, SAO_LIMIT_MIN = 1, SAO_LIMIT_MAX = 2, SAO_LIMIT_RANGE = 3 };
enum { SAO_SCALING_NONE = 0, SAO_SCALING_DIVIDE_BY_N = 1, SAO_SCALING_DIVIDE_BY_N_SQRT = 2,
       SAO_SCALING_ZERO_TO_ONE = 3, SAO_SCALING_20_TIMES_LOG10 = 4 };
/* src/error.rs:37-54 (+ the two deliberate panics of the reference). */
enum { SAO_OK = 0,
       SAO_ERR_TOO_FEW_SAMPLES = 1,
       SAO_ERR_NAN_VALUES = 2,
       SAO_ERR_INFINITY_VALUES = 3,
       SAO_ERR_NOT_POWER_OF_TWO = 4,
       SAO_ERR_LIMIT_BELOW_MINIMUM = 5,
       SAO_ERR_LIMIT_ABOVE_NYQUIST = 6,
       SAO_ERR_LIMIT_INVALID_RANGE = 7,
       SAO_ERR_LIMIT_TOO_NARROW = 8,
       SAO_ERR_SCALING = 9,
       SAO_ERR_UNSUPPORTED_LENGTH = 10,   /* src/fft.rs:52 unimplemented!() panic */
       SAO_ERR_NON_FINITE_RESULT = 11 };  /* src/frequency.rs:52-53 assert panic */

typedef struct {
  float    min_val;  uint32_t min_bin;   /* absolute FFT bin index of the min pair */
  float    max_val;  uint32_t max_bin;
  float    average;
  float    median;
} sao_stats;

/* libm restatements */
float sao_cosf(float x);
float sao_log10f(float x);

/* src/windows.rs: out[i] = w_i * in[i].  Returns 0, or -1 for an unknown kind. */
int sao_window_apply(int window_kind, const float *in, uint32_t n, float *out);
/* window multipliers themselves (apply to a vector of ones is NOT the same for
 * rounding purposes, so this returns the multiplier the reference computes). */
int sao_window_coeffs(int window_kind, uint32_t n, float *out);

/* src/fft.rs:72-133.  out_bins holds (n/2+1) interleaved (re,im) pairs. */
int sao_fft_calc(const float *samples, uint32_t n, float *out_bins);

/* src/limit.rs:96-118 */
int sao_limit_verify(int limit_kind, float lo, float hi, float max_detectable_frequency);

/* src/lib.rs:237-304: resolution and the inclusive kept-bin range.  n_bins may be 0. */
float sao_frequency_resolution(uint32_t sampling_rate, uint32_t n);
void  sao_limit_to_bins(uint32_t n, uint32_t sampling_rate, int limit_kind, float lo, float hi,
                        uint32_t *bin_lo, uint32_t *n_bins);

/* src/spectrum.rs:481-539 on a plain value vector; bin0 is added to the arg-indices. */
void sao_calc_statistics(const float *vals, uint32_t m, uint32_t bin0, sao_stats *out);

/* src/scaling.rs:104-162 */
float sao_scale(int scaling_kind, float v, float stats_min, float stats_max, float stats_avg,
                float stats_median, float stats_n);

/* src/lib.rs:157-209: `samples` are what the caller passes (already windowed).
 * out_freqs/out_vals need room for n/2+1 entries.  scale_err receives (orig,new)
 * when SAO_ERR_SCALING is returned (may be NULL). */
int sao_samples_fft_to_spectrum(const float *samples, uint64_t n, uint32_t sampling_rate,
                                int limit_kind, float lo, float hi, int scaling_kind,
                                float *out_freqs, float *out_vals, uint32_t *out_n_bins,
                                sao_stats *out_stats, float *scale_err);

/* FrequencySpectrum::apply_scaling_fn (src/spectrum.rs:128-168) with scaling::combined(kinds)
 * (src/scaling.rs:174-185) on n_frames spectra of m values each, in place; stats[f] in/out; status[f]
 * in/out (SAO_ERR_SCALING with err_index[f] / err_pair[2f..2f+1] = (orig, new) at the first offender). */
int sao_apply_scaling_chain_batched(float *vals, uint64_t n_frames, uint32_t m, uint64_t stride,
                                    uint32_t bin0, uint32_t n, const int *kinds, uint32_t n_kinds,
                                    sao_stats *stats, uint32_t *status, uint32_t *err_index,
                                    float *err_pair, int threads);

/* Batched driver = window (recomputed per frame, as the reference's callers do)
 * + sao_samples_fft_to_spectrum per frame, frames split statically over
 * `threads` OpenMP threads (the stand-in for "rayon over frames").
 * status[f] = SAO_* code of frame f; out rows are out_stride floats apart.
 * Returns a call-level SAO_* code (limit / length errors) or SAO_OK. */
int sao_spectrum_batched(const float *samples, uint64_t n_frames, uint32_t n, uint64_t frame_stride,
                         uint32_t sampling_rate, int limit_kind, float lo, float hi,
                         int window_kind, int scaling_kind,
                         float *out_vals, uint64_t out_stride, sao_stats *out_stats,
                         uint32_t *status, int threads);

/* src/tests/sine.rs:54-104; returns the number of samples written (<= cap). */
uint64_t sao_sine_wave_audio_data_multiple(const float *frequencies, uint32_t n_freq,
                                           uint32_t sampling_rate, uint32_t duration_ms,
                                           int16_t *out, uint64_t cap);

/* Counter-based synthetic noise, bit-identical twin of the device generator
 * (spectrum_analyzer_b200/csrc/generate.cu): uniform in [-1,1). */
void sao_generate_noise(float *out, uint64_t first_index, uint64_t count, uint64_t seed);

int sao_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
