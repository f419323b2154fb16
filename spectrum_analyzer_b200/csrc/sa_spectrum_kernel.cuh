// sa_spectrum_kernel.cuh -- the fused per-frame kernel family (one instantiation per N).
//
// One kernel does, per frame, everything `samples_fft_to_spectrum` does on the CPU
// (reference src/lib.rs:157-337) plus the caller-side window (src/windows.rs):
//
//   global --(8-byte coalesced loads, window multiply fused, NaN/Inf scan fused)--> registers
//   Stockham autosort radix-16/8/4/2 passes on the packed complex sequence z[n]=x[2n]+i*x[2n+1]
//     (M = N/2 points, E points per thread, T = M/E threads per frame), shared-memory exchange
//     between passes, final pass leaves Z[j + s*T] in slot s of thread j
//   real-FFT recombination with microfft's bin convention (src/fft.rs:126-132): thread j pairs
//     Z[k] with Z[M-k] (upper half exchanged through shared memory) and produces |X[k]|, |X[M-k]|
//   FrequencyLimit slice [bin_lo, bin_hi] (src/lib.rs:286-304), scaling (src/scaling.rs),
//   statistics of the scaled values with the stable-sort tie rules (src/spectrum.rs:481-539):
//     min/argmin, max/argmax, sum by shuffles; median by an exact 4x8-bit radix select
//   coalesced stores of the surviving bins + one 32-byte stats/status record.
//
// Layout: FPB frames per CTA, T threads per frame (small N pack many frames per CTA).
#pragma once
#include "sa_common.cuh"
#include "sa_dft.cuh"

namespace sa {

template <int LOG2N, int E_, int FPB_>
struct SpectrumCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int M = N / 2;          // complex points
  static constexpr int E = E_;             // complex points per thread
  static constexpr int T = M / E;          // threads per frame
  static constexpr int FPB = FPB_;         // frames per block
  static constexpr int NT = T * FPB;       // threads per block
  static constexpr int BUF = M + (M >> 4) + 1;  // padded float2 elements of one exchange buffer
  static constexpr int NV = E + 1;         // output values a thread can own (E pairs + middle bin)
  static_assert(T >= 1 && T * E == M, "bad E");
  static_assert(NT <= 1024, "block too large");
  // shared memory: exchange buffers | per-frame reduction scratch | histograms | select words
  static constexpr size_t SMEM_BUF = sizeof(float2) * (size_t)BUF * FPB;
  static constexpr size_t SMEM_RED = 32 * sizeof(StatsPartial) * FPB;
  static constexpr size_t SMEM_HIST = 256 * sizeof(uint32_t) * FPB;
  static constexpr size_t SMEM_SEL = 4 * sizeof(uint32_t) * FPB;
  static constexpr size_t SMEM_TOTAL = SMEM_BUF + SMEM_RED + SMEM_HIST + SMEM_SEL;
};

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }

// One Stockham pass chain, starting at sub-transform length NS (product of earlier radices).
// On entry x[s] = data[j + s*T]; on exit (after the last pass) x[s] = Z[j + s*T].
template <class C, int NS>
__device__ __forceinline__ void fft_passes(float2 (&x)[C::E], float2 *buf, const int j,
                                           const float2 *__restrict__ twiddle) {
  constexpr int M = C::M, E = C::E, T = C::T;
  constexpr int R = (M / NS) < E ? (M / NS) : E;   // radix of this pass
  constexpr int NB = E / R;                        // butterflies per thread
  constexpr bool LAST = (NS * R == M);
#pragma unroll
  for (int b = 0; b < NB; b++) {
    float2 v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = x[b + k * NB];
    const int jb = j + b * T;
    if constexpr (NS > 1) {
      // twiddle w_{NS*R}^{k*p} = W_N[k*p * N/(NS*R)],  p = jb mod NS
      const int p = jb & (NS - 1);
      constexpr int STEP = C::N / (NS * R);
#pragma unroll
      for (int k = 1; k < R; k++) v[k] = cmul(v[k], __ldg(&twiddle[k * p * STEP]));
    }
    dft<R>(v);
    if constexpr (LAST) {
#pragma unroll
      for (int k = 0; k < R; k++) x[b + k * NB] = v[k];
    } else {
      const int base = (jb / NS) * (NS * R) + (jb & (NS - 1));
#pragma unroll
      for (int k = 0; k < R; k++) buf[pad_idx(base + k * NS)] = v[k];
    }
  }
  if constexpr (!LAST) {
    __syncthreads();
#pragma unroll
    for (int s = 0; s < E; s++) x[s] = buf[pad_idx(j + s * T)];
    __syncthreads();
    fft_passes<C, NS * R>(x, buf, j, twiddle);
  }
}

// Exact k-th smallest (0-based) of the group's valid keys: 4 passes of an 8-bit radix select.
// Returns the key; *below = number of keys < result, *eq = number of keys == result.
template <class C>
__device__ __forceinline__ uint32_t group_select(const uint32_t (&key)[C::NV], const uint32_t valid,
                                                 uint32_t k, uint32_t *hist, uint32_t *sel, const int j,
                                                 uint32_t *below, uint32_t *eq) {
  constexpr int T = C::T;
  constexpr int G = T < 32 ? T : 32;   // scanning threads per group
  constexpr int BPT = 256 / G;         // histogram bins per scanning thread
  uint32_t prefix = 0, prefmask = 0, below_total = 0, eq_count = 0;
#pragma unroll 1
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    for (int b = j; b < 256; b += T) hist[b] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C::NV; i++)
      if (((valid >> i) & 1u) && ((key[i] & prefmask) == prefix)) atomicAdd(&hist[(key[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (T <= 32 || j < 32) {  // warp-uniform
      uint32_t local = 0;
#pragma unroll
      for (int b = 0; b < BPT; b++) local += hist[j * BPT + b];
      uint32_t incl = local;
#pragma unroll
      for (int off = 1; off < G; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off, G);
        if ((j & (G - 1)) >= off) incl += o;
      }
      const uint32_t excl = incl - local;
      if (excl <= k && k < incl) {
        uint32_t c = excl;
#pragma unroll 1
        for (int b = 0; b < BPT; b++) {
          const uint32_t h = hist[j * BPT + b];
          if (k < c + h) {
            sel[0] = j * BPT + b; sel[1] = k - c; sel[2] = h; sel[3] = c;
            break;
          }
          c += h;
        }
      }
    }
    __syncthreads();
    const uint32_t d = sel[0];
    k = sel[1];
    eq_count = sel[2];
    below_total += sel[3];
    prefix |= d << shift;
    prefmask |= 255u << shift;
  }
  *below = below_total;
  *eq = eq_count;
  return prefix;
}

template <class C>
__global__ void __launch_bounds__(C::NT) spectrum_kernel(const SpectrumParams p) {
  constexpr int N = C::N, M = C::M, E = C::E, T = C::T, FPB = C::FPB, NV = C::NV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int fl = tid / T;        // frame slot inside the block
  const int j = tid - fl * T;    // thread index inside the frame group
  float2 *buf = reinterpret_cast<float2 *>(smem_raw) + (size_t)fl * C::BUF;
  StatsPartial *red = reinterpret_cast<StatsPartial *>(smem_raw + C::SMEM_BUF) + fl * 32;
  uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw + C::SMEM_BUF + C::SMEM_RED) + fl * 256;
  uint32_t *sel = reinterpret_cast<uint32_t *>(smem_raw + C::SMEM_BUF + C::SMEM_RED + C::SMEM_HIST) + fl * 4;

  const uint32_t lo = p.bin_lo, hi = p.bin_hi;
  const uint32_t count = hi - lo + 1;
  const bool want_stats = p.stats != nullptr;

#pragma unroll 1
  for (u64 fb = (u64)blockIdx.x * FPB; fb < p.n_frames; fb += (u64)gridDim.x * FPB) {
    const u64 frame = fb + fl;
    const bool active = frame < p.n_frames;

    // ---- load + window + validation (src/windows.rs, src/lib.rs:168-173) ----------------
    float2 x[E];
    uint32_t flags = 0;  // bit0: NaN, bit1: Inf in the windowed samples, bit2: non-finite result
    {
      const u64 base = (active ? frame : 0) * p.frame_stride;
      if (p.aligned8) load_frame_pairs<E, T>(x, p, base, j);
#pragma unroll
      for (int s = 0; s < E; s++) {
        const int n = j + s * T;
        float2 v;
        if (p.aligned8) v = x[s];
        else { v.x = load_sample(p, base + 2 * n); v.y = load_sample(p, base + 2 * n + 1); }
        if (p.window != nullptr) {
          const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window) + n);
          v.x = __fmul_rn(w.x, v.x);
          v.y = __fmul_rn(w.y, v.y);
        }
        if (!active) v = make_float2(0.f, 0.f);
        flags |= (v.x != v.x || v.y != v.y) ? 1u : 0u;
        flags |= (fabsf(v.x) == INFINITY || fabsf(v.y) == INFINITY) ? 2u : 0u;
        x[s] = v;
      }
    }

    // ---- N/2-point complex FFT of the packed sequence ------------------------------------
    fft_passes<C, 1>(x, buf, j, p.twiddle);

    // ---- real-FFT recombination + magnitude (src/fft.rs:126-132, src/lib.rs:366-372) ------
    // thread j owns k = j + s*T for s < E/2 and the mirrored bins M - k; upper half via smem.
    float val[NV];
    uint32_t bin[NV];
    if constexpr (M == 1) {
      // N == 2 never reaches this kernel (handled by the tiny kernel)
    }
#pragma unroll
    for (int s = E / 2; s < E; s++) buf[pad_idx(j + s * T - M / 2)] = x[s];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < E / 2; s++) {
      const int k = j + s * T;
      const float2 a = x[s];
      if (k == 0) {
        // DC and Nyquist (microfft packs Nyquist into bin[0].im): X[0] = Zr+Zi, X[M] = Zr-Zi
        const float dc = a.x + a.y, ny = a.x - a.y;
        val[0] = sqrtf(dc * dc);
        val[1] = sqrtf(ny * ny);
        bin[0] = 0;
        bin[1] = M;
      } else {
        const float2 b = buf[pad_idx(M / 2 - k)];       // Z[M-k]
        const float2 w = __ldg(&p.twiddle[k]);          // (cos, -sin)(2*pi*k/N)
        const float pr = a.x + b.x, pi = a.y - b.y;     // P = A + conj(B)
        const float qr = a.x - b.x, qi = a.y + b.y;     // Q = A - conj(B)
        // T = -i * w * Q,  w = c - i*s  ->  Tr = c*qi - s*qr,  Ti = -(c*qr + s*qi); here w.y = -s
        const float tr = fmaf(w.x, qi, w.y * qr);
        const float ti = fmaf(w.y, qi, -(w.x * qr));
        const float ur = pr + tr, ui = pi + ti;         // 2*X[k]
        const float vr = pr - tr, vi = pi - ti;         // 2*conj(X[M-k])
        val[2 * s] = 0.5f * sqrtf(fmaf(ur, ur, ui * ui));
        val[2 * s + 1] = 0.5f * sqrtf(fmaf(vr, vr, vi * vi));
        bin[2 * s] = k;
        bin[2 * s + 1] = M - k;
      }
    }
    // middle bin k = M/2: X[M/2] = conj(Z[M/2]); owned by thread 0
    uint32_t valid = 0;
    {
      float mid = 0.f;
      if (j == 0) {
        const float2 z = buf[pad_idx(0)];
        mid = sqrtf(fmaf(z.x, z.x, z.y * z.y));
      }
      val[E] = mid;
      bin[E] = M / 2;
    }
    __syncthreads();  // buf is reused by the next frame / nothing else reads it below

    // ---- FrequencyLimit slice (src/lib.rs:286-304) ------------------------------------------
#pragma unroll
    for (int i = 0; i < NV; i++) {
      bool ok = active && bin[i] >= lo && bin[i] <= hi;
      if (i == E) ok = ok && (j == 0);
      if (M == 1 && i == E) ok = false;
      valid |= ok ? (1u << i) : 0u;
    }

    // ---- pre-scaling maximum for scale_to_zero_to_one + result finiteness -------------------
    float vmax = 0.f;
    {
      uint32_t mxbits = 0;
#pragma unroll
      for (int i = 0; i < NV; i++)
        if ((valid >> i) & 1u) {
          mxbits = max(mxbits, __float_as_uint(val[i]));
          flags |= is_finite(val[i]) ? 0u : 4u;
        }
      if (p.scaling_kind == SA_SCALING_ZERO_TO_ONE || want_stats) {
        const uint2 r = group_reduce_max_or<T>(mxbits, flags, reinterpret_cast<uint2 *>(red), j);
        vmax = __uint_as_float(r.x);
        flags = r.y;
      }
    }

    // ---- scaling (src/spectrum.rs:150-164) + store -------------------------------------------
    float *dst = p.out_vals + (active ? frame : 0) * p.out_stride;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      val[i] = apply_scaling(p.scaling_kind, val[i], p.scale_div, vmax);
      if ((valid >> i) & 1u) dst[bin[i] - lo] = val[i];
    }

    // ---- statistics of the scaled values (src/spectrum.rs:481-539) ---------------------------
    if (want_stats) {
      StatsPartial sp;
      sp.min_key = ~0ull;
      sp.max_key = 0ull;
      sp.sum = 0.f;
      sp.flags = flags;
      uint32_t key[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) {
        key[i] = ordered_key(val[i]);
        if ((valid >> i) & 1u) {
          const u64 kk = ((u64)key[i] << 32) | bin[i];
          sp.min_key = u64min(sp.min_key, kk);
          sp.max_key = u64max(sp.max_key, kk);
          sp.sum += val[i];
        }
      }
      sp = group_reduce_stats<T>(sp, red, j);

      float median = 0.f;
      if (p.stats_mask & SA_STATS_MEDIAN) {
        const uint32_t mid = count / 2;
        uint32_t below, eq;
        if (count & 1u) {
          median = key_to_float(group_select<C>(key, valid, mid, hist, sel, j, &below, &eq));
        } else {
          // s[mid-1] and s[mid]: the second is the same value if duplicates reach rank mid,
          // otherwise the smallest value above it.
          const uint32_t ka = group_select<C>(key, valid, mid - 1, hist, sel, j, &below, &eq);
          // "next value above ka" is reduced unconditionally: the reduction contains block-wide
          // barriers / full-mask shuffles, so it must not sit under a per-frame condition.
          uint32_t nxt = 0xffffffffu;
#pragma unroll
          for (int i = 0; i < NV; i++)
            if (((valid >> i) & 1u) && key[i] > ka) nxt = min(nxt, key[i]);
          const uint2 r = group_reduce_max_or<T>(~nxt, 0u, reinterpret_cast<uint2 *>(red), j);  // min == ~max(~x)
          const uint32_t kb = (below + eq <= mid) ? ~r.x : ka;
          // (a + b) / 2.0 in f32 (src/spectrum.rs:517-520)
          median = __fdiv_rn(__fadd_rn(key_to_float(ka), key_to_float(kb)), 2.0f);
        }
      }

      if (active && j == 0) {
        const bool mm = p.stats_mask & SA_STATS_MINMAX;
        uint4 r0, r1;
        r0.x = mm ? __float_as_uint(key_to_float((uint32_t)(sp.min_key >> 32))) : 0u;  // min_val
        r0.y = mm ? (uint32_t)sp.min_key : 0u;                                          // min_bin
        r0.z = mm ? __float_as_uint(key_to_float((uint32_t)(sp.max_key >> 32))) : 0u;  // max_val
        r0.w = mm ? (uint32_t)sp.max_key : 0u;                                          // max_bin
        r1.x = (p.stats_mask & SA_STATS_AVERAGE) ? __float_as_uint(__fdiv_rn(sp.sum, (float)count)) : 0u;
        r1.y = __float_as_uint(median);
        r1.z = (sp.flags & 1u) ? SA_ERR_NAN_VALUES
             : (sp.flags & 2u) ? SA_ERR_INFINITY_VALUES
             : (sp.flags & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
        r1.w = 0u;
        uint4 *rec = reinterpret_cast<uint4 *>(&p.stats[frame]);
        rec[0] = r0;
        rec[1] = r1;
      }
    }
  }
}

}  // namespace sa
