// sa_common.cuh -- shared device helpers: ordered keys, group reductions, scaling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spectrum_analyzer_cuda.h"

namespace sa {

typedef unsigned long long u64;

// Kernel arguments of the fused spectrum kernel (POD, passed by value).
struct SpectrumParams {
  const float *samples;       // device, frame f at samples + f*frame_stride
  const float *window;        // device, N multipliers, or nullptr (SA_WINDOW_NONE)
  const float2 *twiddle;      // device, W_N[i] = exp(-2*pi*i*i/N), i in [0, N)
  float *out_vals;            // device, row f at out_vals + f*out_stride
  sa_frame_stats *stats;      // device, may be nullptr
  u64 n_frames;
  u64 frame_stride;
  u64 out_stride;
  uint32_t bin_lo, bin_hi;    // inclusive kept range (src/lib.rs:286-304)
  int scaling_kind;
  uint32_t stats_mask;
  float scale_div;            // stats.n (divide_by_N) or sqrtf(stats.n) (divide_by_N_sqrt)
  int aligned8;               // samples base and frame_stride allow pair loads (8 B for f32, 4 B for i16)
  int in_i16;                 // samples are int16 PCM (converted exactly, `x as f32`, when loaded)
  int aligned16;              // every frame starts on a 16-byte boundary (fast path of the staged loads)
  int fast_log10;             // opt-in (sa_ctx_set_fast_log10): 20*log10 through lg2.approx instead of the libm sequence
};

// ---- sample loads: f32, or int16 PCM converted on the fly (the reference's callers do `x as f32`,
// src/tests/mod.rs:70-73, examples/mp3-samples.rs:94) --------------------------------------------------
__device__ __forceinline__ float load_sample(const SpectrumParams &p, u64 idx) {
  if (p.in_i16) return (float)__ldg(reinterpret_cast<const short *>(p.samples) + idx);
  return __ldg(p.samples + idx);
}
// E complex pairs (samples 2n, 2n+1 with n = j + s*T) of the frame starting at element `base`;
// requires pair alignment.  The element-type branch is hoisted out of the unrolled loops.
template <int E, int T>
__device__ __forceinline__ void load_frame_pairs(float2 (&x)[E], const SpectrumParams &p, u64 base, int j) {
  if (p.in_i16) {
    const short2 *src = reinterpret_cast<const short2 *>(reinterpret_cast<const short *>(p.samples) + base);
#pragma unroll
    for (int s = 0; s < E; s++) {
      const short2 v = __ldg(src + (j + s * T));
      x[s] = make_float2((float)v.x, (float)v.y);
    }
  } else {
    const float2 *src = reinterpret_cast<const float2 *>(p.samples + base);
#pragma unroll
    for (int s = 0; s < E; s++) x[s] = __ldg(src + (j + s * T));
  }
}

// ---- order-preserving float <-> u32 key (total order of OrderableF32, src/frequency.rs:64-75;
// -0.0 == +0.0 there, so -0.0 is canonicalised first) -------------------------------------------
__device__ __forceinline__ uint32_t ordered_key(float v) {
  const uint32_t b = __float_as_uint(v + 0.0f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Opt-in approximation of 20*log10(v) for v > 0: 20*log10(2) * lg2.approx(v).  lg2.approx.f32 has an absolute
// error of about 2^-22, i.e. 1.4e-6 dB -- below the f32 rounding of any |dB| value above 12 dB and far inside the
// path's tolerance, but not the libm bit pattern.  3 instructions instead of ~45.
__device__ __forceinline__ float db_fast(float v) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
  return __fmul_rn(6.02059991327962390427f, l);
}

// ---- built-in scaling functions (src/scaling.rs:104-162), IEEE ops, no contraction ------------

// musl/libm-0.2 log10f restated with explicit round-to-nearest intrinsics so that nvcc cannot
// contract a*b+c; bit-identical to the f32 sequence of the Rust port, including libm's special cases
// (x < 0 -> NaN, +-0 -> -Inf, +Inf / NaN -> x): a resident spectrum that already holds dB values feeds
// negative numbers in when scale_20_times_log10 is applied again, and the NaN is what makes the reference
// return ScalingError (src/spectrum.rs:150-160).
__device__ __forceinline__ float log10f_musl(float x) {
  const float ivln10hi = 4.3432617188e-01f, ivln10lo = -3.1689971365e-05f,
              log10_2hi = 3.0102920532e-01f, log10_2lo = 7.9034151668e-07f,
              Lg1 = 0xaaaaaa.0p-24f, Lg2 = 0xccce13.0p-25f, Lg3 = 0x91e9ee.0p-25f, Lg4 = 0xf89e26.0p-26f;
  uint32_t ix = __float_as_uint(x);
  int k = 0;
  if (ix < 0x00800000u || (ix >> 31)) {  // x < 2**-126
    if ((ix << 1) == 0u) return -INFINITY;                 // log(+-0) = -inf
    if (ix >> 31) return __uint_as_float(0x7fc00000u);      // log(-#) = NaN (also -NaN, -Inf)
    k -= 25;                                                 // subnormal: scale up
    x = __fmul_rn(x, 33554432.0f);
    ix = __float_as_uint(x);
  } else if (ix >= 0x7f800000u) {
    return x;                                                // +Inf, NaN
  } else if (ix == 0x3f800000u) {
    return 0.0f;
  }
  ix += 0x3f800000u - 0x3f3504f3u;
  k += (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  x = __uint_as_float(ix);
  const float f = __fsub_rn(x, 1.0f);
  const float s = __fdiv_rn(f, __fadd_rn(2.0f, f));
  const float z = __fmul_rn(s, s);
  const float w = __fmul_rn(z, z);
  const float t1 = __fmul_rn(w, __fadd_rn(Lg2, __fmul_rn(w, Lg4)));
  const float t2 = __fmul_rn(z, __fadd_rn(Lg1, __fmul_rn(w, Lg3)));
  const float R = __fadd_rn(t2, t1);
  const float hfsq = __fmul_rn(__fmul_rn(0.5f, f), f);
  float hi = __fsub_rn(f, hfsq);
  hi = __uint_as_float(__float_as_uint(hi) & 0xfffff000u);
  const float lo = __fadd_rn(__fsub_rn(__fsub_rn(f, hi), hfsq), __fmul_rn(s, __fadd_rn(hfsq, R)));
  const float dk = (float)k;
  float r = __fmul_rn(dk, log10_2lo);
  r = __fadd_rn(r, __fmul_rn(__fadd_rn(lo, hi), ivln10lo));
  r = __fadd_rn(r, __fmul_rn(lo, ivln10hi));
  r = __fadd_rn(r, __fmul_rn(hi, ivln10hi));
  r = __fadd_rn(r, __fmul_rn(dk, log10_2hi));
  return r;
}

// The same sequence for positive NORMAL finite x only (every magnitude the tuned kernels produce through
// sqrt.approx.ftz: >= 1e-19 or exactly 0, which the caller maps to 0 as src/scaling.rs:108-112 does).  Bit-identical
// to log10f_musl on that domain with fewer instructions and no branch:
//  * no subnormal rescaling;
//  * no early return for x == 1: the general path yields +0 there (f = 0 => every term is +-0, the sum +0);
//  * s = f / (2 + f) by rcp.approx + one Newton step + two FMA corrections of the quotient instead of the IEEE
//    division routine -- equal to __fdiv_rn for ALL 2^23 possible (f, 2+f) of this call site, checked
//    exhaustively by profiles/microbench/log10_div_check.cu.
__device__ __forceinline__ float log10f_musl_normal(float x) {
  const float ivln10hi = 4.3432617188e-01f, ivln10lo = -3.1689971365e-05f,
              log10_2hi = 3.0102920532e-01f, log10_2lo = 7.9034151668e-07f,
              Lg1 = 0xaaaaaa.0p-24f, Lg2 = 0xccce13.0p-25f, Lg3 = 0x91e9ee.0p-25f, Lg4 = 0xf89e26.0p-26f;
  uint32_t ix = __float_as_uint(x);
  ix += 0x3f800000u - 0x3f3504f3u;
  const int k = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  x = __uint_as_float(ix);
  const float f = __fsub_rn(x, 1.0f);
  const float d = __fadd_rn(2.0f, f);
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
  y = __fmaf_rn(y, __fmaf_rn(-d, y, 1.0f), y);
  float s = __fmul_rn(f, y);
  s = __fmaf_rn(__fmaf_rn(-s, d, f), y, s);
  s = __fmaf_rn(__fmaf_rn(-s, d, f), y, s);
  const float z = __fmul_rn(s, s);
  const float w = __fmul_rn(z, z);
  const float t1 = __fmul_rn(w, __fadd_rn(Lg2, __fmul_rn(w, Lg4)));
  const float t2 = __fmul_rn(z, __fadd_rn(Lg1, __fmul_rn(w, Lg3)));
  const float R = __fadd_rn(t2, t1);
  const float hfsq = __fmul_rn(__fmul_rn(0.5f, f), f);
  float hi = __fsub_rn(f, hfsq);
  hi = __uint_as_float(__float_as_uint(hi) & 0xfffff000u);
  const float lo = __fadd_rn(__fsub_rn(__fsub_rn(f, hi), hfsq), __fmul_rn(s, __fadd_rn(hfsq, R)));
  const float dk = (float)k;
  float r = __fmul_rn(dk, log10_2lo);
  r = __fadd_rn(r, __fmul_rn(__fadd_rn(lo, hi), ivln10lo));
  r = __fadd_rn(r, __fmul_rn(lo, ivln10hi));
  r = __fadd_rn(r, __fmul_rn(hi, ivln10hi));
  r = __fadd_rn(r, __fmul_rn(dk, log10_2hi));
  return r;
}

// v: unscaled magnitude; div: stats.n or sqrtf(stats.n); vmax: pre-scaling maximum.
__device__ __forceinline__ float apply_scaling(int kind, float v, float div, float vmax) {
  switch (kind) {
    case SA_SCALING_DIVIDE_BY_N:       // src/scaling.rs:138-142 (n != 0 always: n >= 2)
    case SA_SCALING_DIVIDE_BY_N_SQRT:  // src/scaling.rs:156-161
      return __fdiv_rn(v, div);
    case SA_SCALING_ZERO_TO_ONE:       // src/scaling.rs:123-127
      return vmax != 0.0f ? __fdiv_rn(v, vmax) : 0.0f;
    case SA_SCALING_20_TIMES_LOG10:    // src/scaling.rs:108-112
      return v == 0.0f ? 0.0f : __fmul_rn(20.0f, log10f_musl(v));
    default:
      return v;
  }
}

__device__ __forceinline__ bool is_finite(float v) {
  return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
}

// ---- reductions over a group of T consecutive threads (T a power of two) ----------------------
// T <= 32: xor shuffles inside the warp segment.  T > 32: warp shuffle, then a per-frame smem
// scratch of T/32 entries; every thread of the group gets the result.  `scratch` must be private
// to the group; calls are separated by __syncthreads internally (block-uniform control flow).

struct StatsPartial {
  u64 min_key;   // (ordered_key << 32) | bin  -> min picks the lowest bin among equal values
  u64 max_key;   // (ordered_key << 32) | bin  -> max picks the highest bin among equal values
  float sum;
  uint32_t flags;
};

__device__ __forceinline__ u64 u64min(u64 a, u64 b) { return a < b ? a : b; }
__device__ __forceinline__ u64 u64max(u64 a, u64 b) { return a > b ? a : b; }

template <int T>
__device__ __forceinline__ StatsPartial group_reduce_stats(StatsPartial v, StatsPartial *scratch, int j) {
  constexpr int W = T < 32 ? T : 32;
#pragma unroll
  for (int off = W / 2; off > 0; off >>= 1) {
    v.min_key = u64min(v.min_key, __shfl_xor_sync(0xffffffffu, v.min_key, off));
    v.max_key = u64max(v.max_key, __shfl_xor_sync(0xffffffffu, v.max_key, off));
    v.sum += __shfl_xor_sync(0xffffffffu, v.sum, off);
    v.flags |= __shfl_xor_sync(0xffffffffu, v.flags, off);
  }
  if constexpr (T > 32) {
    constexpr int NW = T / 32;
    __syncthreads();  // scratch may still be read from a previous round
    if ((j & 31) == 0) scratch[j >> 5] = v;
    __syncthreads();
    StatsPartial r = scratch[0];
#pragma unroll 4
    for (int w = 1; w < NW; w++) {
      const StatsPartial o = scratch[w];
      r.min_key = u64min(r.min_key, o.min_key);
      r.max_key = u64max(r.max_key, o.max_key);
      r.sum += o.sum;
      r.flags |= o.flags;
    }
    v = r;
  }
  return v;
}

// max over the group of a non-negative-float bit pattern / generic u32, plus OR of flags
template <int T>
__device__ __forceinline__ uint2 group_reduce_max_or(uint32_t mx, uint32_t flags, uint2 *scratch, int j) {
  constexpr int W = T < 32 ? T : 32;
#pragma unroll
  for (int off = W / 2; off > 0; off >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    flags |= __shfl_xor_sync(0xffffffffu, flags, off);
  }
  if constexpr (T > 32) {
    constexpr int NW = T / 32;
    __syncthreads();
    if ((j & 31) == 0) scratch[j >> 5] = make_uint2(mx, flags);
    __syncthreads();
    uint2 r = scratch[0];
#pragma unroll 4
    for (int w = 1; w < NW; w++) {
      const uint2 o = scratch[w];
      r.x = max(r.x, o.x);
      r.y |= o.y;
    }
    mx = r.x;
    flags = r.y;
  }
  return make_uint2(mx, flags);
}

}  // namespace sa
