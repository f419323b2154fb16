// sa_aux_kernels.cuh -- kernels beside the fused hot path:
//   spectrum_tiny_kernel   N in {2..32}: one thread per frame (size coverage of src/fft.rs:91-105)
//   window_kernel          windows::*_window on a resident batch (src/windows.rs:40-150)
//   rescale_kernel         FrequencySpectrum::apply_scaling_fn on a resident batch
//                          (src/spectrum.rs:128-168; benches/fft_spectrum_bench.rs:21-36)
//   noise_kernel           counter-based synthetic input (twin: oracle sao_generate_noise)
#pragma once
#include "sa_common.cuh"
#include "sa_dft.cuh"

namespace sa {

// ------------------------------------------------------------------------------------------
// N <= 32: one thread per frame, packed real FFT with in-thread radix-2 DIT.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) spectrum_tiny_kernel(const SpectrumParams p, const int log2n) {
  const u64 frame = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= p.n_frames) return;
  const int n = 1 << log2n, m = n >> 1;
  const u64 base = frame * p.frame_stride;
  float2 z[16];
  uint32_t flags = 0;
  for (int i = 0; i < m; i++) {
    float a = load_sample(p, base + 2 * i), b = load_sample(p, base + 2 * i + 1);
    if (p.window) { a = __fmul_rn(__ldg(p.window + 2 * i), a); b = __fmul_rn(__ldg(p.window + 2 * i + 1), b); }
    flags |= (a != a || b != b) ? 1u : 0u;
    flags |= (fabsf(a) == INFINITY || fabsf(b) == INFINITY) ? 2u : 0u;
    z[i] = make_float2(a, b);
  }
  // bit reversal + radix-2 butterflies (w_size^k = W_N[k * N/size])
  const int lm = log2n - 1;
  if (lm > 0)
    for (int i = 0; i < m; i++) {
      const int r = (int)(__brev((unsigned)i) >> (32 - lm));
      if (r > i) { const float2 t = z[i]; z[i] = z[r]; z[r] = t; }
    }
  for (int size = 2; size <= m; size <<= 1) {
    const int half = size >> 1, step = n / size;
    for (int blk = 0; blk < m; blk += size)
      for (int k = 0; k < half; k++) {
        const float2 w = __ldg(&p.twiddle[k * step]);
        const float2 y = cmul(z[blk + half + k], w);
        const float2 u = z[blk + k];
        z[blk + half + k] = csub(u, y);
        z[blk + k] = cadd(u, y);
      }
  }
  // bins 0..m  (src/fft.rs:126-132 convention: DC, ..., Nyquist)
  float mag[17];
  mag[0] = sqrtf((z[0].x + z[0].y) * (z[0].x + z[0].y));
  mag[m] = sqrtf((z[0].x - z[0].y) * (z[0].x - z[0].y));
  for (int k = 1; k < m; k++) {
    const float2 a = z[k], b = z[m - k];
    const float2 w = __ldg(&p.twiddle[k]);
    const float pr = a.x + b.x, pi = a.y - b.y, qr = a.x - b.x, qi = a.y + b.y;
    const float tr = fmaf(w.x, qi, w.y * qr), ti = fmaf(w.y, qi, -(w.x * qr));
    const float ur = pr + tr, ui = pi + ti;
    mag[k] = 0.5f * sqrtf(fmaf(ur, ur, ui * ui));
  }
  const int lo = (int)p.bin_lo, hi = (int)p.bin_hi, count = hi - lo + 1;
  float vmax = 0.f;
  for (int k = lo; k <= hi; k++) {
    if (!is_finite(mag[k])) flags |= 4u;
    vmax = fmaxf(vmax, mag[k]);
  }
  float *dst = p.out_vals + frame * p.out_stride;
  float sorted[17];
  for (int k = lo; k <= hi; k++) {
    const float v = apply_scaling(p.scaling_kind, mag[k], p.scale_div, vmax) + 0.0f;
    dst[k - lo] = v;
    // stable insertion sort by value (src/spectrum.rs:496-499)
    int i = k - lo;
    while (i > 0 && sorted[i - 1] > v) { sorted[i] = sorted[i - 1]; i--; }
    sorted[i] = v;
    mag[k] = v;
  }
  if (p.stats == nullptr) return;
  int imin = lo, imax = lo;
  float sum = 0.f;
  for (int k = lo; k <= hi; k++) {
    if (mag[k] < mag[imin]) imin = k;       // lowest bin among equal minima
    if (mag[k] >= mag[imax]) imax = k;      // highest bin among equal maxima
  }
  for (int i = 0; i < count; i++) sum = __fadd_rn(sum, sorted[i]);  // ascending order, as the reference
  const int mid = count / 2;
  const float median = (count & 1) ? sorted[mid] : __fdiv_rn(__fadd_rn(sorted[mid - 1], sorted[mid]), 2.0f);
  sa_frame_stats st;
  const bool mm = p.stats_mask & SA_STATS_MINMAX;
  st.min_val = mm ? mag[imin] : 0.f; st.min_bin = mm ? imin : 0;
  st.max_val = mm ? mag[imax] : 0.f; st.max_bin = mm ? imax : 0;
  st.average = (p.stats_mask & SA_STATS_AVERAGE) ? __fdiv_rn(sum, (float)count) : 0.f;
  st.median = (p.stats_mask & SA_STATS_MEDIAN) ? median : 0.f;
  st.status = (flags & 1u) ? SA_ERR_NAN_VALUES : (flags & 2u) ? SA_ERR_INFINITY_VALUES
            : (flags & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
  st.reserved = 0;
  p.stats[frame] = st;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_kernel(const float *__restrict__ in, const float *__restrict__ w,
                                                     float *__restrict__ out, u64 n_frames, uint32_t n,
                                                     u64 frame_stride) {
  const u64 total = n_frames * n;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
    const u64 f = i / n;
    const uint32_t k = (uint32_t)(i - f * n);
    const float x = in[f * frame_stride + k];
    out[i] = w ? __fmul_rn(__ldg(w + k), x) : x;
  }
}

__global__ void __launch_bounds__(256) noise_kernel(float *__restrict__ out, u64 first_index, u64 count, u64 seed) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
    u64 z = seed + (first_index + i) * 0x9E3779B97F4A7C15ull;  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const uint32_t u = (uint32_t)(z >> 40);
    out[i] = __fadd_rn(__fmul_rn((float)u, 0x1p-23f), -1.0f);
  }
}

// ------------------------------------------------------------------------------------------
// apply_scaling_fn on a resident batch: one CTA (256 threads) per frame.
// ------------------------------------------------------------------------------------------
constexpr int RESCALE_THREADS = 256;

__device__ __forceinline__ u64 block_reduce_u64(u64 v, bool is_min, u64 *scratch) {
  for (int off = 16; off > 0; off >>= 1) {
    const u64 o = __shfl_xor_sync(0xffffffffu, v, off);
    v = is_min ? u64min(v, o) : u64max(v, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 r = scratch[0];
  for (int w = 1; w < RESCALE_THREADS / 32; w++) r = is_min ? u64min(r, scratch[w]) : u64max(r, scratch[w]);
  return r;
}

__device__ __forceinline__ float block_reduce_sum(float v, u64 *scratch) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) reinterpret_cast<float *>(scratch)[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < RESCALE_THREADS / 32; w++) r += reinterpret_cast<float *>(scratch)[w];
  return r;
}

// k-th smallest key of vals[0..m) (re-read from global each pass); returns key, below, eq.
__device__ __forceinline__ uint32_t block_select(const float *vals, uint32_t m, uint32_t k, uint32_t *hist,
                                                 uint32_t *sel, uint32_t *below, uint32_t *eq) {
  uint32_t prefix = 0, prefmask = 0, below_total = 0, eq_count = 0;
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    for (int b = threadIdx.x; b < 256; b += RESCALE_THREADS) hist[b] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += RESCALE_THREADS) {
      const uint32_t key = ordered_key(vals[i]);
      if ((key & prefmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t c = 0;
      for (int b = 0; b < 256; b++) {
        const uint32_t h = hist[b];
        if (k < c + h) { sel[0] = b; sel[1] = k - c; sel[2] = h; sel[3] = c; break; }
        c += h;
      }
    }
    __syncthreads();
    k = sel[1]; eq_count = sel[2]; below_total += sel[3];
    prefix |= sel[0] << shift; prefmask |= 255u << shift;
  }
  *below = below_total; *eq = eq_count;
  return prefix;
}

struct ScalingChain {
  int kind[8];
  float div[8];   // stats.n or sqrtf(stats.n) per step
  int count;
};

// scaling::combined (src/scaling.rs:174-185): every step sees the same statistics.  Unlike the fused kernels
// (whose inputs are magnitudes >= 0), a resident spectrum may already hold dB values: log10f_musl keeps libm's
// x < 0 -> NaN and +Inf -> +Inf branches, which is what makes ScalingError reachable with the built-ins.
__device__ __forceinline__ float apply_chain(const ScalingChain &chain, float s, float vmax) {
  for (int c = 0; c < chain.count; c++) s = apply_scaling(chain.kind[c], s, chain.div[c], vmax);
  return s;
}

// One CTA per frame.  Semantics of FrequencySpectrum::apply_scaling_fn (src/spectrum.rs:128-168):
//  * the statistics handed to the scaling function are those of the spectrum BEFORE the call; the only field the
//    built-ins read besides n is max (scale_to_zero_to_one), which is recomputed here from the values themselves
//    (stats.max IS the maximum of the data, src/spectrum.rs:529-530), so the call does not depend on which fields
//    the producing call was asked to fill;
//  * values are replaced in bin order; the FIRST value whose scaled image is NaN or +-Inf stops the frame
//    (src/spectrum.rs:150-160): the values before it stay scaled, it and everything after it stay as they were,
//    the statistics are not recomputed, status becomes SA_ERR_SCALING and `reserved` describes the offender
//    (row index | SA_SCALING_ERR_* class bits) so that the caller can rebuild ScalingError(orig, new);
//  * frames whose status is not SA_OK on entry hold no spectrum in the reference's terms and are left untouched.
__global__ void __launch_bounds__(RESCALE_THREADS) rescale_kernel(float *vals, u64 n_frames, uint32_t m, u64 stride,
                                                                   uint32_t bin_lo, const ScalingChain chain,
                                                                   uint32_t stats_mask, sa_frame_stats *stats) {
  __shared__ u64 scratch[RESCALE_THREADS / 32];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel[4];
  bool needs_max = false;
  for (int c = 0; c < chain.count; c++) needs_max |= (chain.kind[c] == SA_SCALING_ZERO_TO_ONE);
  for (u64 f = blockIdx.x; f < n_frames; f += gridDim.x) {
    float *v = vals + f * stride;
    if (stats[f].status != SA_OK) continue;   // block-uniform
    __syncthreads();
    float vmax = 0.f;
    if (needs_max) {
      u64 km = 0ull;
      for (uint32_t i = threadIdx.x; i < m; i += RESCALE_THREADS) km = u64max(km, (u64)ordered_key(v[i]));
      vmax = key_to_float((uint32_t)block_reduce_u64(km, false, scratch));
    }
    // pass 1: first value (in bin order) whose scaled image is not finite
    u64 bad = ~0ull;
    for (uint32_t i = threadIdx.x; i < m; i += RESCALE_THREADS) {
      const float s = apply_chain(chain, v[i], vmax);
      if (!is_finite(s)) {
        const uint32_t cls = (s != s) ? SA_SCALING_ERR_NAN : (s > 0.f ? SA_SCALING_ERR_POS_INF : SA_SCALING_ERR_NEG_INF);
        bad = u64min(bad, ((u64)i << 32) | cls);
        break;   // a thread's indices ascend
      }
    }
    bad = block_reduce_u64(bad, true, scratch);
    const uint32_t stop = bad == ~0ull ? m : (uint32_t)(bad >> 32);
    // pass 2: replace the values in front of the offender, statistics of the scaled spectrum
    u64 kmin = ~0ull, kmax = 0ull;
    float sum = 0.f;
    for (uint32_t i = threadIdx.x; i < stop; i += RESCALE_THREADS) {
      const float s = apply_chain(chain, v[i], vmax) + 0.0f;
      v[i] = s;
      const u64 kk = ((u64)ordered_key(s) << 32) | (bin_lo + i);
      kmin = u64min(kmin, kk); kmax = u64max(kmax, kk);
      sum += s;
    }
    if (stop != m) {   // block-uniform: ScalingError, statistics stay those of the spectrum before the call
      if (threadIdx.x == 0) { stats[f].status = SA_ERR_SCALING; stats[f].reserved = stop | (uint32_t)bad; }
      continue;
    }
    kmin = block_reduce_u64(kmin, true, scratch);
    kmax = block_reduce_u64(kmax, false, scratch);
    sum = block_reduce_sum(sum, scratch);
    __syncthreads();  // scaled values visible to the whole block (same-block global writes + barrier)
    float median = 0.f;
    if (stats_mask & SA_STATS_MEDIAN) {
      uint32_t below, eq;
      const uint32_t mid = m / 2;
      if (m & 1u) {
        median = key_to_float(block_select(v, m, mid, hist, sel, &below, &eq));
      } else {
        const uint32_t ka = block_select(v, m, mid - 1, hist, sel, &below, &eq);
        u64 nxt = ~0ull;
        for (uint32_t i = threadIdx.x; i < m; i += RESCALE_THREADS) {
          const uint32_t key = ordered_key(v[i]);
          if (key > ka) nxt = u64min(nxt, (u64)key);
        }
        nxt = block_reduce_u64(nxt, true, scratch);
        const uint32_t kb = (below + eq <= mid) ? (uint32_t)nxt : ka;
        median = __fdiv_rn(__fadd_rn(key_to_float(ka), key_to_float(kb)), 2.0f);
      }
    }
    if (threadIdx.x == 0) {
      sa_frame_stats st;
      // min/max are always kept (they are free here and later calls / readers may want them)
      st.min_val = key_to_float((uint32_t)(kmin >> 32)); st.min_bin = (uint32_t)kmin;
      st.max_val = key_to_float((uint32_t)(kmax >> 32)); st.max_bin = (uint32_t)kmax;
      st.average = (stats_mask & SA_STATS_AVERAGE) ? __fdiv_rn(sum, (float)m) : 0.f;
      st.median = median;
      st.status = SA_OK;
      st.reserved = 0;
      stats[f] = st;
    }
    __syncthreads();
  }
}

}  // namespace sa
