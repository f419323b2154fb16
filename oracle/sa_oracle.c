/*
 * sa_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see sa_oracle.h).
 *
 * Plain C restatement of the reference's f32 CPU path.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (see oracle/Makefile)
 * so that every f32 operation is a single IEEE-754 binary32 operation, exactly
 * as rustc emits them (Rust never contracts a*b+c into an FMA).
 *
 * Each function cites the reference file:line it follows.  The two third-party
 * crates on the path are not available on this machine and are restated from
 * their published algorithms:
 *
 *  - microfft 0.6.0 `real::rfft_N`: an N-point real FFT computed as an N/2-point
 *    complex FFT of z[n] = x[2n] + i*x[2n+1] (in-place radix-2 decimation in
 *    time: bit-reversal permutation, then log2(N/2) butterfly levels with f32
 *    twiddles taken from a quarter-wave sine table generated in f64 and rounded
 *    to f32), followed by a recombination pass; DC real part in bin[0].re and
 *    the Nyquist real part packed into bin[0].im (the contract the reference
 *    relies on at src/fft.rs:126-131).  The exact operation order inside
 *    microfft cannot be checked here -> per-bin FFT values: PARITY UNPINNED.
 *
 *  - libm 0.2.16 `cosf` / `log10f`: Rust ports of musl's (FreeBSD msun)
 *    s_cosf.c / k_cosf.c / k_sinf.c / e_rem_pio2f.c (medium path) and
 *    e_log10f.c.  `sqrtf` and `/` are IEEE-exact on both sides.
 */
#include "sa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* libm 0.2.16 restatements (musl forms)                                      */
/* ------------------------------------------------------------------------- */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* musl src/math/__cosdf.c: cos on [-pi/4, pi/4] evaluated in double, rounded once. */
static float k_cosdf(double x) {
  static const double C0 = -0x1ffffffd0c5e81.0p-54, C1 = 0x155553e1053a42.0p-57,
                      C2 = -0x16c087e80f1e27.0p-62, C3 = 0x199342e0ee5069.0p-68;
  double z = x * x;
  double w = z * z;
  double r = C2 + z * C3;
  return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}

/* musl src/math/__sindf.c */
static float k_sindf(double x) {
  static const double S1 = -0x15555554cbac77.0p-55, S2 = 0x111110896efbb2.0p-59,
                      S3 = -0x1a00f9e2cae774.0p-65, S4 = 0x16cd878c3b46a7.0p-71;
  double z = x * x;
  double w = z * z;
  double r = S3 + z * S4;
  double s = z * x;
  return (float)((x + s * (S1 + z * S2)) + s * w * r);
}

/* musl src/math/__rem_pio2f.c, medium-size branch (|x| < 2^28*pi/2).  Window
 * arguments never exceed 12*pi, so the large-argument branch is unreachable on
 * this path; it falls back to the C library there. */
static int k_rem_pio2f(float x, double *y) {
  static const double toint = 1.5 / 2.22044604925031308085e-16, /* 1.5/DBL_EPSILON */
      pio4 = 0x1.921fb6p-1, invpio2 = 6.36619772367581382433e-01,
      pio2_1 = 1.57079631090164184570e+00, pio2_1t = 1.58932547735281966916e-08;
  uint32_t ix = f2u(x) & 0x7fffffffu;
  if (ix < 0x4dc90fdbu) {
    volatile double fnv = (double)x * invpio2 + toint; /* keep the rint() trick unfolded */
    double fn = fnv - toint;
    int n = (int)fn;
    *y = x - fn * pio2_1 - fn * pio2_1t;
    if (*y < -pio4) { n--; fn--; *y = x - fn * pio2_1 - fn * pio2_1t; }
    else if (*y > pio4) { n++; fn++; *y = x - fn * pio2_1 - fn * pio2_1t; }
    return n;
  }
  /* unreachable for window arguments */
  double r = remainder((double)x, 1.57079632679489661923);
  *y = r;
  return (int)llrint(((double)x - r) / 1.57079632679489661923) & 3;
}

/* musl src/math/cosf.c (libm 0.2.16 `cosf`; call sites src/windows.rs:45,67,142). */
float sao_cosf(float x) {
  static const double c1pio2 = 1 * 1.57079632679489661923, c2pio2 = 2 * 1.57079632679489661923,
                      c3pio2 = 3 * 1.57079632679489661923, c4pio2 = 4 * 1.57079632679489661923;
  double y;
  uint32_t ix = f2u(x);
  unsigned sign = ix >> 31;
  ix &= 0x7fffffffu;

  if (ix <= 0x3f490fdau) {           /* |x| ~<= pi/4 */
    if (ix < 0x39800000u) return 1.0f; /* |x| < 2**-12 */
    return k_cosdf(x);
  }
  if (ix <= 0x407b53d1u) {           /* |x| ~<= 5*pi/4 */
    if (ix > 0x4016cbe3u)            /* |x|  ~> 3*pi/4 */
      return -k_cosdf(sign ? x + c2pio2 : x - c2pio2);
    return sign ? k_sindf(x + c1pio2) : k_sindf(c1pio2 - x);
  }
  if (ix <= 0x40e231d5u) {           /* |x| ~<= 9*pi/4 */
    if (ix > 0x40afeddfu)            /* |x| ~> 7*pi/4 */
      return k_cosdf(sign ? x + c4pio2 : x - c4pio2);
    return sign ? k_sindf(-x - c3pio2) : k_sindf(x - c3pio2);
  }
  if (ix >= 0x7f800000u) return x - x; /* cos(Inf or NaN) is NaN */
  unsigned n = (unsigned)k_rem_pio2f(x, &y);
  switch (n & 3) {
    case 0: return k_cosdf(y);
    case 1: return k_sindf(-y);
    case 2: return -k_cosdf(y);
    default: return k_sindf(y);
  }
}

/* musl src/math/log10f.c (libm 0.2.16 `log10f`; call site src/scaling.rs:111). */
float sao_log10f(float x) {
  static const float ivln10hi = 4.3432617188e-01f, ivln10lo = -3.1689971365e-05f,
                     log10_2hi = 3.0102920532e-01f, log10_2lo = 7.9034151668e-07f,
                     Lg1 = 0xaaaaaa.0p-24f, Lg2 = 0xccce13.0p-25f, Lg3 = 0x91e9ee.0p-25f,
                     Lg4 = 0xf89e26.0p-26f;
  float hfsq, f, s, z, R, w, t1, t2, dk, hi, lo;
  uint32_t ix = f2u(x);
  int k = 0;
  if (ix < 0x00800000u || (ix >> 31)) { /* x < 2**-126 */
    if ((ix << 1) == 0) return -1 / (x * x); /* log(+-0) = -inf */
    if (ix >> 31) return (x - x) / 0.0f;     /* log(-#) = NaN */
    k -= 25;                                 /* subnormal: scale up */
    x *= 0x1p25f;
    ix = f2u(x);
  } else if (ix >= 0x7f800000u) {
    return x;
  } else if (ix == 0x3f800000u) {
    return 0;
  }
  /* reduce x into [sqrt(2)/2, sqrt(2)] */
  ix += 0x3f800000u - 0x3f3504f3u;
  k += (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  x = u2f(ix);

  f = x - 1.0f;
  s = f / (2.0f + f);
  z = s * s;
  w = z * z;
  t1 = w * (Lg2 + w * Lg4);
  t2 = z * (Lg1 + w * Lg3);
  R = t2 + t1;
  hfsq = 0.5f * f * f;

  hi = f - hfsq;
  hi = u2f(f2u(hi) & 0xfffff000u);
  lo = f - hi - hfsq + s * (hfsq + R);
  dk = (float)k;
  return dk * log10_2lo + (lo + hi) * ivln10lo + lo * ivln10hi + hi * ivln10hi + dk * log10_2hi;
}

/* ------------------------------------------------------------------------- */
/* src/windows.rs                                                             */
/* ------------------------------------------------------------------------- */

static const float SAO_PI = 3.14159274101257324219f; /* core::f32::consts::PI */
static const float BH4_ALPHA[4] = {0.35875f, -0.48829f, 0.14128f, -0.01168f}; /* windows.rs:82 */
static const float BH7_ALPHA[7] = {0.2710514f,    -0.43329793f,     0.218123f, -0.06592545f,
                                   0.010811742f,  -0.00077658484f,  0.000013887217f}; /* :100-108 */

/* Multiplier of sample i; `n` = samples.len().  n <= 1 is a passthrough for every
 * window except Hann (windows.rs:60-63, 123-126). */
static float window_multiplier(int kind, uint32_t i, uint32_t n) {
  const float len = (float)n;
  switch (kind) {
    case SAO_WINDOW_HANN: { /* windows.rs:43-47 */
      float two_pi_i = 2.0f * SAO_PI * (float)i;
      float c = sao_cosf(two_pi_i / len);
      return 0.5f * (1.0f - c);
    }
    case SAO_WINDOW_HAMMING: /* windows.rs:67 */
      return 0.54f - (0.46f * sao_cosf(2.0f * SAO_PI * (float)i / (len - 1.0f)));
    case SAO_WINDOW_BH4:
    case SAO_WINDOW_BH7: { /* windows.rs:138-144 */
      const float *alpha = kind == SAO_WINDOW_BH4 ? BH4_ALPHA : BH7_ALPHA;
      const int terms = kind == SAO_WINDOW_BH4 ? 4 : 7;
      float acc = 0.0f;
      for (int a = 0; a < terms; a++) {
        float two_pi_iteration = 2.0f * (float)a * SAO_PI;
        float c = sao_cosf((two_pi_iteration * (float)i) / (len - 1.0f));
        acc += alpha[a] * c;
      }
      return acc;
    }
    default:
      return 1.0f;
  }
}

int sao_window_coeffs(int kind, uint32_t n, float *out) {
  if (kind < SAO_WINDOW_NONE || kind > SAO_WINDOW_BH7) return -1;
  const int passthrough = kind == SAO_WINDOW_NONE || (n <= 1 && kind != SAO_WINDOW_HANN);
  for (uint32_t i = 0; i < n; i++) out[i] = passthrough ? 1.0f : window_multiplier(kind, i, n);
  return 0;
}

int sao_window_apply(int kind, const float *in, uint32_t n, float *out) {
  if (kind < SAO_WINDOW_NONE || kind > SAO_WINDOW_BH7) return -1;
  if (kind == SAO_WINDOW_NONE || (n <= 1 && kind != SAO_WINDOW_HANN)) {
    memmove(out, in, (size_t)n * sizeof(float));
    return 0;
  }
  for (uint32_t i = 0; i < n; i++) out[i] = window_multiplier(kind, i, n) * in[i];
  return 0;
}

/* ------------------------------------------------------------------------- */
/* microfft 0.6.0 real::rfft_N restatement                                    */
/* ------------------------------------------------------------------------- */

typedef struct { float re, im; } cpx;

#define SAO_MAX_LOG2 15 /* src/fft.rs:91-105: 2 .. 32768 */

/* Quarter-wave sine table per size, sin(2*pi*i/size) for i in 0..=size/4,
 * computed in f64 and rounded to f32 (microfft's build-time table). */
static float *g_sine[SAO_MAX_LOG2 + 1];

static const float *sine_table(int log2size) {
  float *t = g_sine[log2size];
  if (t) return t;
#ifdef _OPENMP
#pragma omp critical(sao_sine_table)
#endif
  {
    if (!g_sine[log2size]) {
      const uint32_t size = 1u << log2size, q = size / 4;
      float *nt = (float *)malloc(sizeof(float) * (q + 1));
      for (uint32_t i = 0; i <= q; i++)
        nt[i] = (float)sin(2.0 * 3.14159265358979323846 * (double)i / (double)size);
      nt[q] = 1.0f;
      g_sine[log2size] = nt;
    }
  }
  return g_sine[log2size];
}

/* exp(-2*pi*i*k/size) for 0 <= k < size/2, size >= 4, from the quarter table. */
static inline cpx twiddle_from_table(const float *tab, uint32_t size, uint32_t k) {
  const uint32_t q = size / 4;
  cpx w;
  if (k <= q) { w.re = tab[q - k]; w.im = -tab[k]; }
  else        { w.re = -tab[k - q]; w.im = -tab[size / 2 - k]; }
  return w;
}

/* In-place complex FFT of length m = 2^log2m (radix-2 DIT, bit-reversal first). */
static void cfft_inplace(cpx *x, int log2m) {
  const uint32_t m = 1u << log2m;
  for (uint32_t i = 0; i < m; i++) { /* bit_reverse_reorder */
    uint32_t j = 0;
    for (int b = 0; b < log2m; b++) j |= ((i >> b) & 1u) << (log2m - 1 - b);
    if (j > i) { cpx t = x[i]; x[i] = x[j]; x[j] = t; }
  }
  if (log2m == 0) return;
  for (uint32_t blk = 0; blk < m; blk += 2) { /* size-2 level: twiddle == 1 */
    cpx a = x[blk], b = x[blk + 1];
    x[blk].re = a.re + b.re;     x[blk].im = a.im + b.im;
    x[blk + 1].re = a.re - b.re; x[blk + 1].im = a.im - b.im;
  }
  for (int ls = 2; ls <= log2m; ls++) { /* compute_butterflies, bottom-up */
    const uint32_t size = 1u << ls, half = size / 2;
    const float *tab = sine_table(ls);
    for (uint32_t blk = 0; blk < m; blk += size) {
      for (uint32_t k = 0; k < half; k++) {
        cpx w = twiddle_from_table(tab, size, k);
        cpx v = x[blk + half + k], u = x[blk + k], y;
        y.re = w.re * v.re - w.im * v.im; /* num_complex Mul: (ac-bd, ad+bc) */
        y.im = w.re * v.im + w.im * v.re;
        x[blk + half + k].re = u.re - y.re; x[blk + half + k].im = u.im - y.im;
        x[blk + k].re = u.re + y.re;        x[blk + k].im = u.im + y.im;
      }
    }
  }
}

/* real::rfft_N on buf[0..n): in place; on return buf holds n/2 complex bins with
 * the Nyquist real part in bin[0].im. */
static void rfft_packed_inplace(float *buf, int log2n) {
  const uint32_t n = 1u << log2n;
  cpx *x = (cpx *)buf;
  if (log2n == 1) { /* 2-point: DC = a+b, Nyquist = a-b */
    float a = buf[0], b = buf[1];
    buf[0] = a + b; buf[1] = a - b;
    return;
  }
  const uint32_t m = n / 2; /* complex length */
  cfft_inplace(x, log2n - 1);
  { cpx x0 = x[0]; x[0].re = x0.re + x0.im; x[0].im = x0.re - x0.im; }
  const uint32_t h = m / 2;
  const float *tab = sine_table(log2n);
  for (uint32_t k = 1; k < h; k++) { /* recombine: X[k] = E[k] + e^{-2 pi i k/n} O[k] */
    const uint32_t nk = m - k;
    cpx w = twiddle_from_table(tab, n, k);
    cpx a = x[k], b = x[nk];
    float sum_re = 0.5f * (a.re + b.re), sum_im = 0.5f * (a.im + b.im);
    float diff_re = 0.5f * (a.re - b.re), diff_im = 0.5f * (a.im - b.im);
    x[k].re = sum_re + w.re * sum_im + w.im * diff_re;
    x[k].im = diff_im + w.im * sum_im - w.re * diff_re;
    x[nk].re = sum_re - w.re * sum_im - w.im * diff_re;
    x[nk].im = -diff_im + w.im * sum_im - w.re * diff_re;
  }
  x[h].im = -x[h].im;
}

static int log2_exact(uint64_t n) {
  int l = 0;
  while ((1ull << l) < n) l++;
  return l;
}

/* src/fft.rs:72-133 */
int sao_fft_calc(const float *samples, uint32_t n, float *out_bins) {
  if (n < 2 || (n & (n - 1)) || n > (1u << SAO_MAX_LOG2)) return SAO_ERR_UNSUPPORTED_LENGTH;
  memcpy(out_bins, samples, (size_t)n * sizeof(float)); /* vec_buffer.extend_from_slice */
  rfft_packed_inplace(out_bins, log2_exact(n));
  const float nyq = out_bins[1];                        /* buffer[0].im */
  out_bins[1] = 0.0f;
  out_bins[n] = nyq; out_bins[n + 1] = 0.0f;            /* push(Complex32::new(nyq, 0.0)) */
  return SAO_OK;
}

/* ------------------------------------------------------------------------- */
/* src/limit.rs, src/lib.rs bin mapping                                       */
/* ------------------------------------------------------------------------- */

static int verify_one(float x, float nyq) { /* limit.rs:99-107 */
  if (x < 0.0f) return SAO_ERR_LIMIT_BELOW_MINIMUM;
  if (x > nyq) return SAO_ERR_LIMIT_ABOVE_NYQUIST;
  return SAO_OK;
}

int sao_limit_verify(int kind, float lo, float hi, float nyq) {
  int e;
  switch (kind) {
    case SAO_LIMIT_ALL: return SAO_OK;
    case SAO_LIMIT_MIN: return verify_one(lo, nyq);
    case SAO_LIMIT_MAX: return verify_one(hi, nyq);
    case SAO_LIMIT_RANGE: /* limit.rs:108-116 */
      if ((e = verify_one(lo, nyq))) return e;
      if ((e = verify_one(hi, nyq))) return e;
      if (lo > hi) return SAO_ERR_LIMIT_INVALID_RANGE;
      return SAO_OK;
    default: return SAO_ERR_LIMIT_INVALID_RANGE;
  }
}

float sao_frequency_resolution(uint32_t sampling_rate, uint32_t n) { /* lib.rs:356-358 */
  return (float)sampling_rate / (float)n;
}

void sao_limit_to_bins(uint32_t n, uint32_t sampling_rate, int kind, float lo, float hi,
                       uint32_t *bin_lo, uint32_t *n_bins) { /* lib.rs:252-304 */
  const float res = sao_frequency_resolution(sampling_rate, n);
  const int has_min = kind == SAO_LIMIT_MIN || kind == SAO_LIMIT_RANGE;
  const int has_max = kind == SAO_LIMIT_MAX || kind == SAO_LIMIT_RANGE;
  uint32_t first = 0, count = 0;
  for (uint32_t i = 0; i <= n / 2; i++) {
    const float fr = (float)i * res;
    if (has_min && !(fr >= lo)) continue;
    if (has_max && !(fr <= hi)) continue;
    if (count == 0) first = i;
    count++;
  }
  *bin_lo = first; *n_bins = count;
}

/* ------------------------------------------------------------------------- */
/* src/spectrum.rs calc_statistics (stable sort by value)                     */
/* ------------------------------------------------------------------------- */

typedef struct { float val; uint32_t idx; } vi_pair;

/* Stable merge sort by val (ascending); any stable sort gives the order of
 * Rust's slice::sort_by (spectrum.rs:496-499). */
static void merge_sort_pairs(vi_pair *a, vi_pair *tmp, uint32_t n) {
  if (n <= 16) {
    for (uint32_t i = 1; i < n; i++) {
      vi_pair key = a[i];
      uint32_t j = i;
      while (j > 0 && a[j - 1].val > key.val) { a[j] = a[j - 1]; j--; }
      a[j] = key;
    }
    return;
  }
  const uint32_t h = n / 2;
  merge_sort_pairs(a, tmp, h);
  merge_sort_pairs(a + h, tmp, n - h);
  uint32_t i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = (a[j].val < a[i].val) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(vi_pair));
}

static void calc_statistics_ws(const float *vals, uint32_t m, uint32_t bin0, sao_stats *out,
                               vi_pair *work, vi_pair *tmp) {
  for (uint32_t i = 0; i < m; i++) { work[i].val = vals[i]; work[i].idx = i; }
  merge_sort_pairs(work, tmp, m);
  float sum = 0.0f; /* spectrum.rs:505-508: fold over the SORTED copy */
  for (uint32_t i = 0; i < m; i++) sum = sum + work[i].val;
  out->average = sum / (float)m;
  const uint32_t mid = m / 2; /* spectrum.rs:515-524 */
  out->median = (m % 2 == 0) ? (work[mid - 1].val + work[mid].val) / 2.0f : work[mid].val;
  out->min_val = work[0].val;     out->min_bin = bin0 + work[0].idx;     /* :529 */
  out->max_val = work[m - 1].val; out->max_bin = bin0 + work[m - 1].idx; /* :530 */
}

void sao_calc_statistics(const float *vals, uint32_t m, uint32_t bin0, sao_stats *out) {
  vi_pair *work = (vi_pair *)malloc(sizeof(vi_pair) * 2 * (size_t)(m ? m : 1));
  calc_statistics_ws(vals, m, bin0, out, work, work + m);
  free(work);
}

/* ------------------------------------------------------------------------- */
/* src/scaling.rs                                                             */
/* ------------------------------------------------------------------------- */

float sao_scale(int kind, float v, float smin, float smax, float savg, float smedian, float n) {
  (void)smin; (void)savg; (void)smedian;
  switch (kind) {
    case SAO_SCALING_DIVIDE_BY_N: return n == 0.0f ? v : v / n;                 /* :134-143 */
    case SAO_SCALING_DIVIDE_BY_N_SQRT: return n == 0.0f ? v : v / sqrtf(n);     /* :152-162 */
    case SAO_SCALING_ZERO_TO_ONE: return smax != 0.0f ? v / smax : 0.0f;        /* :119-128 */
    case SAO_SCALING_20_TIMES_LOG10: return v == 0.0f ? 0.0f : 20.0f * sao_log10f(v); /* :104-113 */
    default: return v;
  }
}

/* ------------------------------------------------------------------------- */
/* src/lib.rs:157-337                                                         */
/* ------------------------------------------------------------------------- */

typedef struct { float *bins; vi_pair *work; } frame_ws;

static int spectrum_core(const float *samples, uint64_t n64, uint32_t sampling_rate, int limit_kind,
                         float lo, float hi, int scaling_kind, float *out_freqs, float *out_vals,
                         uint32_t *out_n_bins, sao_stats *out_stats, float *scale_err, frame_ws *ws) {
  if (n64 < 2) return SAO_ERR_TOO_FEW_SAMPLES;                                   /* :164 */
  for (uint64_t i = 0; i < n64; i++) if (isnan(samples[i])) return SAO_ERR_NAN_VALUES;       /* :168 */
  for (uint64_t i = 0; i < n64; i++) if (isinf(samples[i])) return SAO_ERR_INFINITY_VALUES;  /* :171 */
  if (n64 & (n64 - 1)) return SAO_ERR_NOT_POWER_OF_TWO;                          /* :174 */
  const float nyq = (float)sampling_rate / 2.0f;                                 /* :177 */
  int e = sao_limit_verify(limit_kind, lo, hi, nyq);                             /* :179 */
  if (e) return e;
  if (n64 > (1u << SAO_MAX_LOG2)) return SAO_ERR_UNSUPPORTED_LENGTH;             /* fft.rs:52 */
  const uint32_t n = (uint32_t)n64;

  e = sao_fft_calc(samples, n, ws->bins);                                        /* :194 */
  if (e) return e;
  const float res = sao_frequency_resolution(sampling_rate, n);                  /* :237 */
  const int has_min = limit_kind == SAO_LIMIT_MIN || limit_kind == SAO_LIMIT_RANGE;
  const int has_max = limit_kind == SAO_LIMIT_MAX || limit_kind == SAO_LIMIT_RANGE;
  uint32_t m = 0, bin0 = 0;
  for (uint32_t i = 0; i <= n / 2; i++) {                                        /* :252-313 */
    const float fr = (float)i * res;
    if (has_min && !(fr >= lo)) continue;
    if (has_max && !(fr <= hi)) continue;
    const float re = ws->bins[2 * i], im = ws->bins[2 * i + 1];
    const float sum = re * re + im * im;                                         /* :368 */
    const float mag = sqrtf(sum);
    if (isnan(fr) || isinf(fr) || isnan(mag) || isinf(mag)) return SAO_ERR_NON_FINITE_RESULT;
    if (m == 0) bin0 = i;
    if (out_freqs) out_freqs[m] = fr;
    out_vals[m] = mag;
    m++;
  }
  if (out_n_bins) *out_n_bins = m;
  if (m < 2) return SAO_ERR_LIMIT_TOO_NARROW;                                    /* :317 */

  sao_stats st;
  calc_statistics_ws(out_vals, m, bin0, &st, ws->work, ws->work + m);            /* :324-329 */
  if (scaling_kind != SAO_SCALING_NONE) {                                        /* :332-334 */
    const float nf = (float)n;                                                   /* spectrum.rs:144 */
    for (uint32_t i = 0; i < m; i++) {                                           /* spectrum.rs:150-164 */
      const float v = out_vals[i];
      const float s = sao_scale(scaling_kind, v, st.min_val, st.max_val, st.average, st.median, nf);
      if (isnan(s) || isinf(s)) {
        if (scale_err) { scale_err[0] = v; scale_err[1] = s; }
        return SAO_ERR_SCALING;
      }
      out_vals[i] = s;
    }
    calc_statistics_ws(out_vals, m, bin0, &st, ws->work, ws->work + m);          /* spectrum.rs:166 */
  }
  if (out_stats) *out_stats = st;
  return SAO_OK;
}

static int ws_alloc(frame_ws *ws, uint64_t n) {
  const uint64_t cap = n < 2 ? 4 : n + 2;
  ws->bins = (float *)malloc(sizeof(float) * cap);
  ws->work = (vi_pair *)malloc(sizeof(vi_pair) * 2 * (cap / 2 + 1));
  return (ws->bins && ws->work) ? 0 : -1;
}
static void ws_free(frame_ws *ws) { free(ws->bins); free(ws->work); }

int sao_samples_fft_to_spectrum(const float *samples, uint64_t n, uint32_t sampling_rate,
                                int limit_kind, float lo, float hi, int scaling_kind,
                                float *out_freqs, float *out_vals, uint32_t *out_n_bins,
                                sao_stats *out_stats, float *scale_err) {
  frame_ws ws;
  const uint64_t cap = (n > (1u << SAO_MAX_LOG2)) ? 4 : n;
  if (ws_alloc(&ws, cap)) return -1;
  int e = spectrum_core(samples, n, sampling_rate, limit_kind, lo, hi, scaling_kind, out_freqs,
                        out_vals, out_n_bins, out_stats, scale_err, &ws);
  ws_free(&ws);
  return e;
}

int sao_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* FrequencySpectrum::apply_scaling_fn on a batch of spectra, src/spectrum.rs:128-168, with */
/* scaling::combined (src/scaling.rs:174-185) as the scaling function: every function of the  */
/* chain sees the statistics of the spectrum BEFORE the call.  Per frame: values are replaced  */
/* in order; the first scaled value that is NaN or +-Inf returns ScalingError(orig, new)       */
/* (:150-160) -- the values in front of it stay replaced, the rest and the statistics stay.    */
/* Frames whose status is not OK on entry are skipped (no spectrum object exists for them).    */
/* ------------------------------------------------------------------------- */
int sao_apply_scaling_chain_batched(float *vals, uint64_t n_frames, uint32_t m, uint64_t stride,
                                    uint32_t bin0, uint32_t n, const int *kinds, uint32_t n_kinds,
                                    sao_stats *stats, uint32_t *status, uint32_t *err_index,
                                    float *err_pair, int threads) {
  if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    vi_pair *work = (vi_pair *)malloc(sizeof(vi_pair) * 2 * (size_t)(m ? m : 1));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t f = 0; f < (int64_t)n_frames; f++) {
      if (status[f] != SAO_OK) continue;
      float *v = vals + (uint64_t)f * stride;
      const sao_stats st = stats[f];                       /* :138-145, n = samples_len as f32 */
      const float nf = (float)n;
      int failed = 0;
      for (uint32_t i = 0; i < m; i++) {                   /* :150 */
        float s = v[i];
        for (uint32_t c = 0; c < n_kinds; c++)             /* scaling.rs:176-182 */
          s = sao_scale(kinds[c], s, st.min_val, st.max_val, st.average, st.median, nf);
        if (isnan(s) || isinf(s)) {                        /* :155-160 */
          status[f] = SAO_ERR_SCALING;
          if (err_index) err_index[f] = i;
          if (err_pair) { err_pair[2 * f] = v[i]; err_pair[2 * f + 1] = s; }
          failed = 1;
          break;
        }
        v[i] = s;                                          /* :163 */
      }
      if (!failed) calc_statistics_ws(v, m, bin0, &stats[f], work, work + m);   /* :166 */
    }
    free(work);
  }
  return SAO_OK;
}

int sao_spectrum_batched(const float *samples, uint64_t n_frames, uint32_t n, uint64_t frame_stride,
                         uint32_t sampling_rate, int limit_kind, float lo, float hi,
                         int window_kind, int scaling_kind, float *out_vals, uint64_t out_stride,
                         sao_stats *out_stats, uint32_t *status, int threads) {
  /* call-level checks, same precedence as lib.rs:164-181 minus the per-frame scans */
  if (n < 2) return SAO_ERR_TOO_FEW_SAMPLES;
  if (n & (n - 1)) return SAO_ERR_NOT_POWER_OF_TWO;
  int e = sao_limit_verify(limit_kind, lo, hi, (float)sampling_rate / 2.0f);
  if (e) return e;
  if (n > (1u << SAO_MAX_LOG2)) return SAO_ERR_UNSUPPORTED_LENGTH;
  uint32_t bin_lo, n_bins;
  sao_limit_to_bins(n, sampling_rate, limit_kind, lo, hi, &bin_lo, &n_bins);
  if (n_bins < 2) return SAO_ERR_LIMIT_TOO_NARROW;
  for (int l = 2; l <= log2_exact(n); l++) (void)sine_table(l); /* tables before the parallel region */
  if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    frame_ws ws;
    float *windowed = (float *)malloc(sizeof(float) * n);
    ws_alloc(&ws, n);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t f = 0; f < (int64_t)n_frames; f++) {
      const float *x = samples + (uint64_t)f * frame_stride;
      sao_window_apply(window_kind, x, n, windowed); /* the caller-side window step */
      uint32_t m = 0;
      sao_stats st;
      memset(&st, 0, sizeof st);
      int code = spectrum_core(windowed, n, sampling_rate, limit_kind, lo, hi, scaling_kind, NULL,
                               out_vals + (uint64_t)f * out_stride, &m, &st, NULL, &ws);
      if (out_stats) out_stats[f] = st;
      if (status) status[f] = (uint32_t)code;
    }
    ws_free(&ws);
    free(windowed);
  }
  return SAO_OK;
}

/* ------------------------------------------------------------------------- */
/* src/tests/sine.rs:54-104                                                   */
/* ------------------------------------------------------------------------- */

uint64_t sao_sine_wave_audio_data_multiple(const float *frequencies, uint32_t n_freq,
                                           uint32_t sampling_rate, uint32_t duration_ms,
                                           int16_t *out, uint64_t cap) {
  if (n_freq == 0) return 0;
  const uint64_t sample_count =
      (uint64_t)((float)sampling_rate * ((float)duration_ms / 1000.0f)); /* sine.rs:70 */
  uint64_t written = 0;
  for (uint64_t i = 0; i < sample_count && written < cap; i++) {
    const float t = (1.0f / (float)sampling_rate) * (float)i;            /* sine.rs:76 */
    float acc = 0.0f;
    for (uint32_t w = 0; w < n_freq; w++) acc += sinf(t * frequencies[w] * 2.0f * SAO_PI); /* :37 */
    acc = acc * 32767.0f * 0.1f;                                         /* :87 */
    int16_t s;
    if (acc > 32767.0f) s = 32767;
    else if (acc < -32768.0f) s = -32768;
    else s = (int16_t)acc;                                               /* truncating `as i16` */
    out[written++] = s;
  }
  return written;
}

/* ------------------------------------------------------------------------- */
/* synthetic noise (twin of the device generator)                             */
/* ------------------------------------------------------------------------- */

void sao_generate_noise(float *out, uint64_t first_index, uint64_t count, uint64_t seed) {
  for (uint64_t i = 0; i < count; i++) {
    uint64_t z = seed + (first_index + i) * 0x9E3779B97F4A7C15ull; /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const uint32_t u = (uint32_t)(z >> 40); /* 24 bits */
    out[i] = (float)u * 0x1p-23f - 1.0f;    /* exact: [-1, 1) on a 2^-23 grid */
  }
}
