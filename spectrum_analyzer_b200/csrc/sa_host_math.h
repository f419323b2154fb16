// sa_host_math.h -- host-side arithmetic of the product library (no CUDA):
//   * libm-0.2.16 `cosf` restated (musl s_cosf.c / k_cosf.c / k_sinf.c / e_rem_pio2f.c medium
//     path), used to build the window tables exactly as the reference computes each multiplier
//     (reference src/windows.rs:43-47, 66-69, 138-144);
//   * FrequencyLimit::verify and the inclusive bin filter (src/limit.rs:96-118, src/lib.rs:237-304);
//   * the twiddle table of the device FFT.
// This file is part of the PRODUCT (it is not the test oracle and shares no code with oracle/).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/spectrum_analyzer_cuda.h"

namespace sa_host {

inline uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// |x| <= pi/4 kernels, evaluated in double and rounded once to float
inline float cos_kernel(double x) {
  const double C0 = -0x1ffffffd0c5e81.0p-54, C1 = 0x155553e1053a42.0p-57,
               C2 = -0x16c087e80f1e27.0p-62, C3 = 0x199342e0ee5069.0p-68;
  const double z = x * x, w = z * z, r = C2 + z * C3;
  return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}
inline float sin_kernel(double x) {
  const double S1 = -0x15555554cbac77.0p-55, S2 = 0x111110896efbb2.0p-59,
               S3 = -0x1a00f9e2cae774.0p-65, S4 = 0x16cd878c3b46a7.0p-71;
  const double z = x * x, w = z * z, r = S3 + z * S4, s = z * x;
  return (float)((x + s * (S1 + z * S2)) + s * w * r);
}

// cosf of libm 0.2.16 (the crate the reference links, Cargo.lock:987-990) for the finite,
// moderate arguments a window produces (|x| <= 12*pi).
inline float cosf_libm(float x) {
  const double pio2 = 1.57079632679489661923;
  uint32_t ix = bits(x);
  const bool sign = ix >> 31;
  ix &= 0x7fffffffu;
  if (ix <= 0x3f490fdau) return ix < 0x39800000u ? 1.0f : cos_kernel(x);
  if (ix <= 0x407b53d1u) {
    if (ix > 0x4016cbe3u) return -cos_kernel(sign ? x + 2 * pio2 : x - 2 * pio2);
    return sign ? sin_kernel(x + pio2) : sin_kernel(pio2 - x);
  }
  if (ix <= 0x40e231d5u) {
    if (ix > 0x40afeddfu) return cos_kernel(sign ? x + 4 * pio2 : x - 4 * pio2);
    return sign ? sin_kernel(-x - 3 * pio2) : sin_kernel(x - 3 * pio2);
  }
  if (ix >= 0x7f800000u) return x - x;
  // medium-size argument reduction: 25+53 bit pi
  const double toint = 1.5 / 2.22044604925031308085e-16, pio4 = 0x1.921fb6p-1,
               invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079631090164184570e+00,
               pio2_1t = 1.58932547735281966916e-08;
  volatile double t = (double)x * invpio2 + toint;
  double fn = t - toint;
  int n = (int)fn;
  double y = x - fn * pio2_1 - fn * pio2_1t;
  if (y < -pio4) { n--; fn--; y = x - fn * pio2_1 - fn * pio2_1t; }
  else if (y > pio4) { n++; fn++; y = x - fn * pio2_1 - fn * pio2_1t; }
  switch (n & 3) {
    case 0: return cos_kernel(y);
    case 1: return sin_kernel(-y);
    case 2: return -cos_kernel(y);
    default: return sin_kernel(y);
  }
}

// The multiplier the reference applies to sample i of an n-sample frame.
inline float window_multiplier(int kind, uint32_t i, uint32_t n) {
  const float PI = 3.14159274101257324219f;  // core::f32::consts::PI
  const float len = (float)n;
  if (kind != SA_WINDOW_NONE && kind != SA_WINDOW_HANN && n <= 1) return 1.0f;  // windows.rs:60-63,123-126
  switch (kind) {
    case SA_WINDOW_HANN: {
      const float two_pi_i = 2.0f * PI * (float)i;
      return 0.5f * (1.0f - cosf_libm(two_pi_i / len));
    }
    case SA_WINDOW_HAMMING:
      return 0.54f - (0.46f * cosf_libm(2.0f * PI * (float)i / (len - 1.0f)));
    case SA_WINDOW_BLACKMAN_HARRIS_4TERM:
    case SA_WINDOW_BLACKMAN_HARRIS_7TERM: {
      static const float A4[4] = {0.35875f, -0.48829f, 0.14128f, -0.01168f};
      static const float A7[7] = {0.2710514f, -0.43329793f, 0.218123f, -0.06592545f,
                                  0.010811742f, -0.00077658484f, 0.000013887217f};
      const bool four = kind == SA_WINDOW_BLACKMAN_HARRIS_4TERM;
      const float *alpha = four ? A4 : A7;
      const int terms = four ? 4 : 7;
      float acc = 0.0f;
      for (int a = 0; a < terms; a++) {
        const float two_pi_iteration = 2.0f * (float)a * PI;
        acc += alpha[a] * cosf_libm((two_pi_iteration * (float)i) / (len - 1.0f));
      }
      return acc;
    }
    default:
      return 1.0f;
  }
}

inline int verify_one(float x, float nyquist) {
  if (x < 0.0f) return SA_ERR_LIMIT_BELOW_MINIMUM;
  if (x > nyquist) return SA_ERR_LIMIT_ABOVE_NYQUIST;
  return SA_OK;
}

// FrequencyLimit::verify (src/limit.rs:96-118)
inline int limit_verify(int kind, float lo, float hi, float nyquist) {
  int e;
  switch (kind) {
    case SA_LIMIT_ALL: return SA_OK;
    case SA_LIMIT_MIN: return verify_one(lo, nyquist);
    case SA_LIMIT_MAX: return verify_one(hi, nyquist);
    case SA_LIMIT_RANGE:
      if ((e = verify_one(lo, nyquist)) != SA_OK) return e;
      if ((e = verify_one(hi, nyquist)) != SA_OK) return e;
      return lo > hi ? SA_ERR_LIMIT_INVALID_RANGE : SA_OK;
    default: return SA_ERR_INVALID_ARGUMENT;
  }
}

// Kept bins of fft_result_to_spectrum (src/lib.rs:252-304): fr_i = i as f32 * (sr as f32 / n as f32),
// kept iff fr_i >= min (if any) and fr_i <= max (if any), all in f32.  fr_i is non-decreasing in i,
// so the kept set is one contiguous range found by bisection on the same f32 expression.
inline void limit_bins(uint32_t n, uint32_t sampling_rate, int kind, float lo, float hi,
                       uint32_t *bin_lo, uint32_t *n_bins) {
  const volatile float res = (float)sampling_rate / (float)n;
  const uint32_t last = n / 2;
  auto freq = [&](uint32_t i) { volatile float f = (float)i * res; return (float)f; };
  uint32_t first = 0, end = last + 1;  // [first, end)
  if (kind == SA_LIMIT_MIN || kind == SA_LIMIT_RANGE) {
    uint32_t a = 0, b = last + 1;  // smallest i with freq(i) >= lo
    while (a < b) { const uint32_t m = a + (b - a) / 2; if (freq(m) >= lo) b = m; else a = m + 1; }
    first = a;
  }
  if (kind == SA_LIMIT_MAX || kind == SA_LIMIT_RANGE) {
    uint32_t a = 0, b = last + 1;  // smallest i with freq(i) > hi
    while (a < b) { const uint32_t m = a + (b - a) / 2; if (freq(m) > hi) b = m; else a = m + 1; }
    end = a;
  }
  *n_bins = end > first ? end - first : 0;
  *bin_lo = *n_bins ? first : 0;
}

// W_N[i] = exp(-2*pi*i*i/N), i in [0, N): quarter-wave sine table in double, rounded to f32,
// mirrored so that exact 0 / +-1 appear where they should.
inline std::vector<float> twiddle_table(uint32_t n) {
  std::vector<float> w(2 * (size_t)n);
  if (n < 4) {
    for (uint32_t i = 0; i < n; i++) {
      w[2 * i] = (n == 2 && i == 1) ? -1.0f : 1.0f;
      w[2 * i + 1] = 0.0f;
    }
    return w;
  }
  const uint32_t q = n / 4;
  std::vector<float> s(q + 1);
  for (uint32_t i = 0; i <= q; i++) s[i] = (float)std::sin(2.0 * 3.14159265358979323846 * (double)i / (double)n);
  s[q] = 1.0f;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t r = i % q, quad = i / q;  // angle = quad*pi/2 + r*2pi/n
    const float sn = s[r], cs = s[q - r];
    float c, sv;
    switch (quad) {
      case 0: c = cs; sv = sn; break;
      case 1: c = -sn; sv = cs; break;
      case 2: c = -cs; sv = -sn; break;
      default: c = sn; sv = -cs; break;
    }
    w[2 * i] = c + 0.0f;
    w[2 * i + 1] = -sv + 0.0f;
  }
  return w;
}

}  // namespace sa_host
