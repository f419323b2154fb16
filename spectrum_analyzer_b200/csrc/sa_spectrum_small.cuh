// sa_spectrum_small.cuh -- fused kernel family for small frames, N in {64, 128, 256, 512}
// ("many frames per CTA", BASELINE.json configs[4]).
//
// A frame is processed by T = N/32 lanes (2, 4, 8 or 16) holding 16 complex points each, so one warp
// carries FPW = 32/T frames in lockstep and a 256-thread CTA 8*FPW frames.  Nothing in the main loop
// synchronises wider than a warp:
//   * two Stockham passes (radix 16, then radix T) with one warp-private shared-memory exchange,
//   * real-FFT recombination (src/fft.rs:126-132 convention) by width-T shuffles (partner lane T-j),
//   * magnitude / scaling / stores as in the tuned family (sa_spectrum_v2.cuh),
//   * statistics by redux.sync over the group's lane mask,
//   * exact median by a per-frame histogram (8 buckets per lane over [min,max] of the value bit
//     patterns), a width-T shuffle scan, compaction of the selected bucket and a rank count; crowded
//     buckets iterate on the narrowed range.  Control flow is warp-uniform (per-group predicates).
#pragma once
#include "sa_spectrum_v2.cuh"

namespace sa {

// Minimum / maximum over the T lanes of a frame (T = 2..16 lanes, 32/T frames per warp).  A redux.sync with one member
// mask PER FRAME diverges -- the frames of a warp take turns (ncu, cfg5: 13 % of the stall samples sat on these three
// reductions) -- whereas xor-shuffles below T stay inside every aligned group of T lanes at once with the full mask.
template <int T>
__device__ __forceinline__ uint32_t group_min_u32(uint32_t v) {
#pragma unroll
  for (int off = T / 2; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
template <int T>
__device__ __forceinline__ uint32_t group_max_u32(uint32_t v) {
#pragma unroll
  for (int off = T / 2; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

template <int LOG2N_>
struct SmallCfg {
  static constexpr int LOG2N = LOG2N_, N = 1 << LOG2N_, M = N / 2, E = 16, T = M / 16;
  static constexpr int FPW = 32 / T, WARPS = 8, NT = 32 * WARPS;
  static constexpr int R2 = T, NB2 = 16 / R2;           // second (last) pass
  static constexpr int FBUF = M + M / 16;               // padded float2 per frame
  static constexpr int NB = 8 * T, LOGNB = LOG2N_ - 2;  // histogram buckets per frame (8 per lane)
  static constexpr int HS = NB + 8;                     // [below | NB buckets | above | pad]
  static constexpr int NTW2 = (R2 - 1) * 16, NTWR = M / 2;
  static_assert(T >= 2 && T <= 16, "small family covers N = 64..512");
  // per-thread constants (window multipliers of the 32 samples a lane owns, last-pass twiddles, recombination
  // twiddles) live in tensor memory, as in the tuned family (sa_tmem.cuh): columns [0,32) window |
  // [32,62) pass-2 twiddles, entry b*(R2-1) + k-1 | [64,80) recombination.  They depend on lane % T only.
  static constexpr int TM_COLS = 128, TM_WIN = 0, TM_TW2 = 32, TM_TWR = 64;
  static constexpr size_t OFF_TMSLOT = 0;
  static constexpr size_t OFF_WARPS = 16;
  static constexpr size_t W_BUF = sizeof(float2) * FBUF * FPW;
  static constexpr size_t W_HIST = sizeof(uint32_t) * HS * FPW;
  static constexpr size_t W_MISC = sizeof(uint32_t) * 16 * FPW;
  static constexpr size_t W_LIST = sizeof(uint32_t) * 32 * FPW;
  static constexpr size_t W_BYTES = W_BUF + W_HIST + W_MISC + W_LIST;
  static constexpr size_t SMEM_TOTAL = OFF_WARPS + W_BYTES * WARPS;
};

enum { SM_ARGMIN = 0, SM_ARGMAX = 1, SM_LCOUNT = 2, SM_FLAGS = 3, SM_RESA = 4, SM_RESB = 5 };

template <class C, int MODE, bool FULL>
__device__ __forceinline__ void spectrum_body_small(const V2Params &P, unsigned char *smem, const uint32_t tm_base) {
  constexpr int N = C::N, M = C::M, E = C::E, T = C::T, FPW = C::FPW, R2 = C::R2, NB2 = C::NB2;
  constexpr bool ODD = FULL;                         // FrequencyLimit::All keeps N/2+1 bins: one middle rank
  constexpr uint32_t LIST_MAX = (2 * T > 8) ? 2 * T : 8;   // larger buckets are narrowed by another pass
  const SpectrumParams &p = P.b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fl = lane / T, j = lane - fl * T;                       // frame slot in the warp, lane in the group
  const uint32_t gmask = (T == 16 ? 0xffffu : ((1u << T) - 1u)) << (fl * T);
  unsigned char *wbase = smem + C::OFF_WARPS + (size_t)warp * C::W_BYTES;
  float2 *buf = reinterpret_cast<float2 *>(wbase) + fl * C::FBUF;
  uint32_t *hist = reinterpret_cast<uint32_t *>(wbase + C::W_BUF) + fl * C::HS;
  uint32_t *misc = reinterpret_cast<uint32_t *>(wbase + C::W_BUF + C::W_HIST) + fl * 16;
  uint32_t *list = reinterpret_cast<uint32_t *>(wbase + C::W_BUF + C::W_HIST + C::W_MISC) + fl * 32;

  const bool has_win = p.window != nullptr;
  const uint32_t ta = tmem_addr(tm_base, warp, 0);   // this thread's table row (lane quadrant warp % 4)
  if (tid < 128) {   // warps 0..3 fill the four lane quadrants; every warp of the CTA reads the same columns
    float v[16];
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const float2 w = has_win ? __ldg(reinterpret_cast<const float2 *>(p.window) + j + (8 * h + s) * T) : make_float2(1.f, 1.f);
        v[2 * s] = w.x; v[2 * s + 1] = w.y;
      }
      tmem_st16(ta + C::TM_WIN + 16 * h, v);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const int e = 8 * h + s;
        float2 w = make_float2(0.f, 0.f);
        if (e < NB2 * (R2 - 1)) {
          const int b = e / (R2 - 1), k = e % (R2 - 1) + 1;
          w = __ldg(P.tw2 + (k - 1) * 16 + j + b * T);
        }
        v[2 * s] = w.x; v[2 * s + 1] = w.y;
      }
      tmem_st16(ta + C::TM_TW2 + 16 * h, v);
    }
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const float2 w = __ldg(P.twr + j + s * T);
      v[2 * s] = w.x; v[2 * s + 1] = w.y;
    }
    tmem_st16(ta + C::TM_TWR, v);
  }
  tmem_fence_before_sync();
  for (int i = j; i < C::HS; i += T) hist[i] = 0;
  if (j == 0) { misc[SM_ARGMIN] = 0xffffffffu; misc[SM_ARGMAX] = 0u; misc[SM_LCOUNT] = 0u; misc[SM_FLAGS] = 0u; }
  __syncthreads();
  tmem_fence_after_sync();

  const uint32_t lo = p.bin_lo, hi = p.bin_hi;
  const uint32_t count = hi - lo + 1;
  const bool want_stats = p.stats != nullptr;
  const bool want_median = want_stats && (p.stats_mask & SA_STATS_MEDIAN);
  const float cscale = P.cmul;

  u64 frame = ((u64)blockIdx.x * C::WARPS + warp) * FPW + fl;
  const u64 fstep = (u64)gridDim.x * C::WARPS * FPW;
  bool active = frame < p.n_frames;
  if (!__any_sync(0xffffffffu, active)) return;   // warp-uniform; no CTA-wide barrier below

  float2 x[E];
#pragma unroll
  for (int s = 0; s < E; s++) x[s] = make_float2(0.f, 0.f);
  if (active) load_frame_pairs<E, T>(x, p, frame * p.frame_stride, j);

  auto binof = [&](int i) -> uint32_t {
    const uint32_t k = (uint32_t)(j + (i >> 1) * T);
    return (i & 1) ? (uint32_t)M - k : k;
  };

#pragma unroll 1
  for (;;) {
    // ---- window (src/windows.rs), exact f32 multiply ---------------------------------------------
    {
      float w[32];
      tmem_ld32(ta + C::TM_WIN, w);
#pragma unroll
      for (int s = 0; s < E; s++) {
        x[s].x = __fmul_rn(w[2 * s], x[s].x);
        x[s].y = __fmul_rn(w[2 * s + 1], x[s].y);
      }
    }
    // ---- pass 1: radix 16 ---------------------------------------------------------------------------
    dft16(x);
    {
      float2 *w = buf + 17 * j;
#pragma unroll
      for (int k = 0; k < 16; k++) w[k] = x[k];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int idx = j + s * T;
      x[s] = buf[idx + (idx >> 4)];
    }
    __syncwarp();
    // ---- pass 2 (last): radix T, NS = 16; x[s] = Z[j + s*T] afterwards ------------------------------
    float t2[32];
    tmem_ld32(ta + C::TM_TW2, t2);   // w_M^(k (j + b T)), entry b*(R2-1) + k-1
#pragma unroll
    for (int b = 0; b < NB2; b++) {
      float2 v[R2];
#pragma unroll
      for (int k = 0; k < R2; k++) v[k] = x[b + k * NB2];
#pragma unroll
      for (int k = 1; k < R2; k++) {
        const int e = b * (R2 - 1) + k - 1;
        v[k] = cmul(v[k], make_float2(t2[2 * e], t2[2 * e + 1]));
      }
      dft<R2>(v);
#pragma unroll
      for (int k = 0; k < R2; k++) x[b + k * NB2] = v[k];
    }
    // ---- recombination by shuffles: partner of k = j + s*T is M - k = lane (T-j)%T, slot 15-s --------
    float val[E];
    float vmid = 0.f;
    {
      const int pl = (T - j) & (T - 1);
      float tr[16];
      tmem_ld16(ta + C::TM_TWR, tr);   // W_N^(j + s T)
#pragma unroll
      for (int s = 0; s < E / 2; s++) {
        float2 bq;
        bq.x = __shfl_sync(0xffffffffu, x[15 - s].x, pl, T);
        bq.y = __shfl_sync(0xffffffffu, x[15 - s].y, pl, T);
        if (j == 0 && s > 0) bq = x[16 - s];   // lane 0 pairs k = s*T with (16-s)*T, its own slot
        const float2 a = x[s];
        const float2 w = make_float2(tr[2 * s], tr[2 * s + 1]);
        const float2 cb = make_float2(bq.x, -bq.y);
        const float2 pp = add2(a, cb), qq = sub2(a, cb);
        const float2 tt = fma2(bcast2(w.x), make_float2(qq.y, -qq.x), mul2(bcast2(w.y), qq));
        const float2 uu = add2(pp, tt), vv = sub2(pp, tt);
        val[2 * s] = sqrt_approx(fmaf(uu.x, uu.x, uu.y * uu.y));
        val[2 * s + 1] = sqrt_approx(fmaf(vv.x, vv.x, vv.y * vv.y));
      }
      if (j == 0) {
        const float2 z0 = x[0], zm = x[8];
        val[0] = 2.0f * fabsf(z0.x + z0.y);                       // X[0]
        val[1] = 2.0f * fabsf(z0.x - z0.y);                       // X[N/2]  (packed Nyquist)
        vmid = 2.0f * sqrt_approx(fmaf(zm.x, zm.x, zm.y * zm.y)); // X[N/4] = conj(Z[M/2])
      }
    }
    // ---- request the next frame ---------------------------------------------------------------------
    const u64 cur = frame;
    const bool cur_active = active;
    frame += fstep;
    active = frame < p.n_frames;
#pragma unroll
    for (int s = 0; s < E; s++) x[s] = make_float2(0.f, 0.f);
    if (active) load_frame_pairs<E, T>(x, p, frame * p.frame_stride, j);
    const bool more = __any_sync(0xffffffffu, active);

    // ---- FrequencyLimit slice ---------------------------------------------------------------------------
    uint32_t valid = 0xffffffffu;
    bool valid_mid = (j == 0);
    if constexpr (!FULL) {
      valid = 0;
#pragma unroll
      for (int i = 0; i < E; i++) {
        const uint32_t b = binof(i);
        valid |= (b >= lo && b <= hi) ? (1u << i) : 0u;
      }
      valid_mid = valid_mid && ((uint32_t)(M / 2) >= lo && (uint32_t)(M / 2) <= hi);
    }
#define SA_VALID(i) (FULL || ((valid >> (i)) & 1u))

    float umax = 0.f;
    if constexpr (MODE == SCALE_ZERO_ONE || MODE == SCALE_LOG10) {
      uint32_t ub = 0;
#pragma unroll
      for (int i = 0; i < E; i++) {
        val[i] *= 0.5f;
        if (SA_VALID(i)) ub = max(ub, __float_as_uint(val[i]) & 0x7fffffffu);
      }
      vmid *= 0.5f;
      if (valid_mid) ub = max(ub, __float_as_uint(vmid) & 0x7fffffffu);
      umax = __uint_as_float(group_max_u32<T>(ub));
    }
    auto scale1 = [&](float v) -> float {
      if constexpr (MODE == SCALE_MUL) return v * cscale;
      else if constexpr (MODE == SCALE_DIV) return div_const(v * 0.5f, P.div, P.rdiv);
      else if constexpr (MODE == SCALE_ZERO_ONE) return umax != 0.0f ? __fdiv_rn(v, umax) : 0.0f;
      else return (v == 0.0f) ? 0.0f : __fmul_rn(20.0f, log10f_musl(v));
    };
    // zero-to-one: one reciprocal per frame + exact FMA correction instead of 17 IEEE divisions (sa_spectrum_v2.cuh);
    // lanes of different frames may take different branches, each branch is the same exact function
    if (MODE == SCALE_ZERO_ONE && frame_max_is_fast_divisor(umax)) {
      const float rmax = __frcp_rn(umax);
#pragma unroll
      for (int i = 0; i < E; i++) val[i] = div_by_frame_max(val[i], umax, rmax);
      vmid = div_by_frame_max(vmid, umax, rmax);
    } else if (MODE == SCALE_LOG10 && p.fast_log10) {   // launch-uniform opt-in
#pragma unroll
      for (int i = 0; i < E; i++) val[i] = (val[i] == 0.0f) ? 0.0f : db_fast(val[i]);
      vmid = (vmid == 0.0f) ? 0.0f : db_fast(vmid);
    } else if constexpr (MODE == SCALE_LOG10) {
      // bins 0 and N/2 of thread 0 are |Zr +- Zi| (they may be subnormal): full libm sequence; every other value
      // comes out of sqrt.approx.ftz (normal or exactly zero): the branch-free variant, same bits
      const float dc0 = val[0], dc1 = val[1];
#pragma unroll
      for (int i = 0; i < E; i++) val[i] = (val[i] == 0.0f) ? 0.0f : __fmul_rn(20.0f, log10f_musl_normal(val[i]));
      vmid = (vmid == 0.0f) ? 0.0f : __fmul_rn(20.0f, log10f_musl_normal(vmid));
      if (j == 0) { val[0] = scale1(dc0); val[1] = scale1(dc1); }
    } else {
#pragma unroll
      for (int i = 0; i < E; i++) val[i] = scale1(val[i]);
      vmid = scale1(vmid);
    }

    // ---- statistics --------------------------------------------------------------------------------------
    if (want_stats) {
      const uint32_t ra = (count & 1u) ? count / 2 : count / 2 - 1, rb = count / 2;
      uint32_t kmn = 0xffffffffu, kmx = 0u;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < E; i++) {
        if (SA_VALID(i)) {
          const uint32_t k = v2_key<MODE>(val[i]);
          kmn = min(kmn, k);
          kmx = max(kmx, k);
          sum += val[i];
        }
      }
      if (valid_mid) {
        const uint32_t k = v2_key<MODE>(vmid);
        kmn = min(kmn, k);
        kmx = max(kmx, k);
        sum += vmid;
      }
      const uint32_t my_kmn = kmn, my_kmx = kmx;
      kmn = group_min_u32<T>(kmn);
      kmx = group_max_u32<T>(kmx);
#pragma unroll
      for (int off = T / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off, T);
      if (my_kmn == kmn) {
        uint32_t best = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < E; i++)
          if (SA_VALID(i) && v2_key<MODE>(val[i]) == kmn) best = min(best, binof(i));
        if (valid_mid && v2_key<MODE>(vmid) == kmn) best = min(best, (uint32_t)(M / 2));
        atomicMin(&misc[SM_ARGMIN], best);
      }
      if (my_kmx == kmx) {
        uint32_t best = 0u;
#pragma unroll
        for (int i = 0; i < E; i++)
          if (SA_VALID(i) && v2_key<MODE>(val[i]) == kmx) best = max(best, binof(i));
        if (valid_mid && v2_key<MODE>(vmid) == kmx) best = max(best, (uint32_t)(M / 2));
        atomicMax(&misc[SM_ARGMAX], best);
      }
      bool nonfinite;
      if constexpr (MODE == SCALE_LOG10 || MODE == SCALE_ZERO_ONE) nonfinite = (__float_as_uint(umax) >= 0x7f800000u);
      else nonfinite = (kmx >= 0x7f800000u);
      nonfinite = nonfinite && cur_active;
      if (__any_sync(0xffffffffu, nonfinite)) {   // rare: classify NaN / Inf in the windowed input
        if (nonfinite) {
          uint32_t f = 4u;
          for (int n = j; n < N; n += T) {
            const float v = __fmul_rn(has_win ? __ldg(p.window + n) : 1.0f, load_sample(p, cur * p.frame_stride + n));
            if (v != v) f |= 1u;
            if (fabsf(v) == INFINITY) f |= 2u;
          }
          atomicOr(&misc[SM_FLAGS], f);
        }
      }

      // ---- median: exact selection, warp-uniform control flow with per-group predicates ----------------
      uint32_t ka = kmn, kb = kmn;
      if (want_median) {
        uint32_t klo = kmn, width = kmx - kmn, rka = ra, rkb = rb;
        uint32_t cand_lo = 0u, cand_w = 0xffffffffu;
        bool done = (kmx == kmn) || !cur_active;
        // first pass: a window around the mean (mean +- (max-min)/8 in value space) instead of [min,max];
        // the "below"/"above" slots keep the ranks exact, and a window that misses the middle rank is
        // detected below (no owner bucket) and replaced by the full range
        bool windowed = false;
        if (!done) {
          const float vmin = v2_unkey<MODE>(kmn), vmax = v2_unkey<MODE>(kmx);
          const float mean = sum / (float)count, d = (vmax - vmin) * 0.125f;
          const uint32_t wlo = v2_key<MODE>(fmaxf(mean - d, vmin)), whi = v2_key<MODE>(fminf(mean + d, vmax));
          if (whi > wlo && wlo >= kmn && whi <= kmx) { klo = wlo; width = whi - wlo; windowed = true; }
        }
#pragma unroll 1
        while (__any_sync(0xffffffffu, !done)) {
          const int sh = max(0, (32 - __clz(width | 1u)) - C::LOGNB);
          // first pass of every frame in the warp (the usual, only pass): every kept key is a candidate -- no range test
          // in front of the atomics (a predicated shared-memory atomic is a branch per value)
          const bool everything = __all_sync(0xffffffffu, done || cand_w == 0xffffffffu);
          if (everything) {
            if (!done) {
#pragma unroll
              for (int i = 0; i < E; i++)
                if (SA_VALID(i)) atomicAdd(&hist[hist_slot<MODE, C::NB>(v2_key<MODE>(val[i]), klo, sh)], 1u);
              if (valid_mid) atomicAdd(&hist[hist_slot<MODE, C::NB>(v2_key<MODE>(vmid), klo, sh)], 1u);
            }
          } else if (!done) {
#pragma unroll
            for (int i = 0; i < E; i++) {
              const uint32_t k = v2_key<MODE>(val[i]);
              if (SA_VALID(i) && (k - cand_lo) <= cand_w) atomicAdd(&hist[hist_slot<MODE, C::NB>(k, klo, sh)], 1u);
            }
            if (valid_mid) {
              const uint32_t k = v2_key<MODE>(vmid);
              if ((k - cand_lo) <= cand_w) atomicAdd(&hist[hist_slot<MODE, C::NB>(k, klo, sh)], 1u);
            }
          }
          __syncwarp();
          // scan: 8 buckets per lane, width-T shuffle prefix; buckets are cleared for the next use
          uint32_t hb[8];
          uint32_t local = 0;
#pragma unroll
          for (int b = 0; b < 8; b++) { hb[b] = hist[1 + j * 8 + b]; local += hb[b]; }
          const uint32_t nbelow = hist[0];
          __syncwarp();
#pragma unroll
          for (int b = 0; b < 8; b++) hist[1 + j * 8 + b] = 0;
          if (j == 0) { hist[0] = 0; hist[C::NB + 1] = 0; }
          uint32_t incl = local;
#pragma unroll
          for (int off = 1; off < T; off <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off, T);
            if (j >= off) incl += o;
          }
          const uint32_t c0 = nbelow + incl - local;
          uint32_t sa_b = 0, sa_below = 0, sa_cnt = 0, sb_b = 0;
          bool own_a = false, own_b = false;
          {
            uint32_t c = c0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
              if (rka >= c && rka < c + hb[b]) { own_a = true; sa_b = j * 8 + b; sa_below = c; sa_cnt = hb[b]; }
              if (!ODD && rkb >= c && rkb < c + hb[b]) { own_b = true; sb_b = j * 8 + b; }
              c += hb[b];
            }
          }
          const uint32_t bal_a = (__ballot_sync(0xffffffffu, own_a) >> (fl * T)) & (T == 16 ? 0xffffu : ((1u << T) - 1u));
          uint32_t bal_b = bal_a;
          if constexpr (!ODD) bal_b = (__ballot_sync(0xffffffffu, own_b) >> (fl * T)) & (T == 16 ? 0xffffu : ((1u << T) - 1u));
          const int oa = bal_a ? __ffs(bal_a) - 1 : 0, ob = bal_b ? __ffs(bal_b) - 1 : 0;
          const uint32_t ba = __shfl_sync(0xffffffffu, sa_b, oa, T);
          const uint32_t below_a = __shfl_sync(0xffffffffu, sa_below, oa, T);
          const uint32_t cnt_a = __shfl_sync(0xffffffffu, sa_cnt, oa, T);
          uint32_t bb = ba;
          if constexpr (!ODD) bb = __shfl_sync(0xffffffffu, sb_b, ob, T);
          const uint32_t lo_a = klo + (ba << sh), lo_b = klo + (bb << sh), bw = (sh ? (1u << sh) : 1u) - 1u;
          // the window missed a middle rank (it fell into the below/above slot): redo with the full range
          const bool miss = !done && windowed && (bal_a == 0u || bal_b == 0u);
          windowed = false;
          const bool single = !done && !miss && sh == 0;
          const bool split = !done && !miss && sh != 0 && ba != bb;
          const bool uselist = !done && !miss && sh != 0 && ba == bb && cnt_a <= LIST_MAX;
          const bool refine = !done && !miss && sh != 0 && ba == bb && cnt_a > LIST_MAX;
          if (single) { ka = klo + ba; kb = klo + bb; }
          if (!ODD && __any_sync(0xffffffffu, split)) {
            uint32_t ma = 0u, mb = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < E; i++) {
              const uint32_t k = v2_key<MODE>(val[i]);
              if (SA_VALID(i) && (k - lo_a) <= bw) ma = max(ma, k);
              if (SA_VALID(i) && (k - lo_b) <= bw) mb = min(mb, k);
            }
            if (valid_mid) {
              const uint32_t k = v2_key<MODE>(vmid);
              if ((k - lo_a) <= bw) ma = max(ma, k);
              if ((k - lo_b) <= bw) mb = min(mb, k);
            }
            ma = group_max_u32<T>(ma);
            mb = group_min_u32<T>(mb);
            if (split) { ka = ma; kb = mb; }
          }
          uint32_t lane_np = 0u, lane_ck = 0u;
          if (uselist) {
            // branch-free count of this thread's candidates (plus the last one); almost always 0 or 1
            uint32_t np = 0, ck = 0;
#pragma unroll
            for (int i = 0; i < E; i++) {
              const uint32_t k = v2_key<MODE>(val[i]);
              const bool c = SA_VALID(i) && (k - lo_a) <= bw;
              ck = c ? k : ck;
              np += c ? 1u : 0u;
            }
            if (valid_mid) {
              const uint32_t k = v2_key<MODE>(vmid);
              const bool c = (k - lo_a) <= bw;
              ck = c ? k : ck;
              np += c ? 1u : 0u;
            }
            lane_np = np; lane_ck = ck;
          }
          // No lane of the warp holds more than one candidate (the usual case): the candidates stay in registers and the
          // wanted ranks are found by stripping group minima with xor-shuffles -- no list, no atomics, no second
          // __syncwarp round trip through shared memory.  Otherwise (warp-uniform decision) the list below.
          // (one middle rank only: with two, the list measured faster -- 256-point Range spectra 0.336 vs 0.323)
          const bool inreg = ODD && !__any_sync(0xffffffffu, uselist && lane_np > 1u);
          if (inreg) {
            constexpr uint32_t NONE = 0xffffffffu;          // never the key of a finite value
            constexpr uint32_t GBITS = (T == 16 ? 0xffffu : ((1u << T) - 1u));
            uint32_t kk = (uselist && lane_np == 1u) ? lane_ck : NONE;
            const uint32_t wa = rka - below_a, wb = rkb - below_a;
            bool needa = uselist, needb = uselist;
            uint32_t base = 0u;
#pragma unroll 1
            while (__any_sync(0xffffffffu, needa || needb)) {
              const uint32_t mn = group_min_u32<T>(kk);
              const bool eq = (kk == mn) && (kk != NONE);
              const uint32_t c = (uint32_t)__popc((__ballot_sync(0xffffffffu, eq) >> (fl * T)) & GBITS);
              if (needa && (wa < base + c || c == 0u)) { ka = mn; needa = false; }
              if (needb && (wb < base + c || c == 0u)) { kb = mn; needb = false; }
              base += c;
              kk = eq ? NONE : kk;
            }
          }
          if (uselist && !inreg) {
            const uint32_t np = lane_np, ck = lane_ck;
            if (np == 1u) {
              list[atomicAdd(&misc[SM_LCOUNT], 1u) & 31u] = ck;
            } else if (np > 1u) {
#pragma unroll
              for (int i = 0; i < E; i++) {
                const uint32_t k = v2_key<MODE>(val[i]);
                if (SA_VALID(i) && (k - lo_a) <= bw) list[atomicAdd(&misc[SM_LCOUNT], 1u) & 31u] = k;
              }
              if (valid_mid) {
                const uint32_t k = v2_key<MODE>(vmid);
                if ((k - lo_a) <= bw) list[atomicAdd(&misc[SM_LCOUNT], 1u) & 31u] = k;
              }
            }
          }
          __syncwarp();
          if (uselist && !inreg) {
            const uint32_t wa = rka - below_a, wb = rkb - below_a;
            for (uint32_t a = j; a < cnt_a; a += T) {
              const uint32_t kk = list[a];
              uint32_t r = 0;
              for (uint32_t b = 0; b < cnt_a; b++) {
                const uint32_t ko = list[b];
                r += (ko < kk || (ko == kk && b < a)) ? 1u : 0u;
              }
              if (r == wa) misc[SM_RESA] = kk;
              if (r == wb) misc[SM_RESB] = kk;
            }
          }
          __syncwarp();
          if (uselist && !inreg) {
            ka = misc[SM_RESA];
            kb = misc[SM_RESB];
            if (j == 0) misc[SM_LCOUNT] = 0u;
          }
          if (miss) {
            klo = kmn; width = kmx - kmn;
          } else if (refine) {
            klo = lo_a; width = bw; cand_lo = lo_a; cand_w = bw;
            rka -= below_a; rkb -= below_a;
          } else {
            done = true;
          }
          __syncwarp();
        }
      }
      __syncwarp();   // arg-bin / flag atomics of every lane are visible
      if (cur_active && j == 0) {
        const bool mm = p.stats_mask & SA_STATS_MINMAX;
        uint4 r0, r1;
        r0.x = mm ? __float_as_uint(v2_unkey<MODE>(kmn)) : 0u;
        r0.y = mm ? misc[SM_ARGMIN] : 0u;
        r0.z = mm ? __float_as_uint(v2_unkey<MODE>(kmx)) : 0u;
        r0.w = mm ? misc[SM_ARGMAX] : 0u;
        r1.x = (p.stats_mask & SA_STATS_AVERAGE) ? __float_as_uint(__fdiv_rn(sum, (float)count)) : 0u;
        float med = 0.f;
        if (want_median)
          med = (count & 1u) ? v2_unkey<MODE>(kb) : __fdiv_rn(__fadd_rn(v2_unkey<MODE>(ka), v2_unkey<MODE>(kb)), 2.0f);
        r1.y = __float_as_uint(med);
        const uint32_t fg = misc[SM_FLAGS];
        r1.z = (fg & 1u) ? SA_ERR_NAN_VALUES : (fg & 2u) ? SA_ERR_INFINITY_VALUES : (fg & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
        r1.w = 0u;
        uint4 *rec = reinterpret_cast<uint4 *>(&p.stats[cur]);
        rec[0] = r0;
        rec[1] = r1;
      }
      if (j == 0) { misc[SM_ARGMIN] = 0xffffffffu; misc[SM_ARGMAX] = 0u; misc[SM_FLAGS] = 0u; }
      __syncwarp();
    }
    // stores after the statistics (whose shuffle / shared-memory chains would queue behind them, cf. sa_spectrum_v2.cuh)
    if (cur_active) {
      float *dst = p.out_vals + cur * p.out_stride - lo;
#pragma unroll
      for (int i = 0; i < E; i++)
        if (SA_VALID(i)) dst[binof(i)] = val[i];
      if (valid_mid) dst[M / 2] = vmid;
    }
#undef SA_VALID
    if (!more) break;
  }
}

#ifndef SA_SMALL_CTAS
#define SA_SMALL_CTAS 2   // resident CTAs per SM the register budget is sized for (2: <= 128 registers)
#endif
template <class C, int MODE, bool FULL>
__global__ void __launch_bounds__(C::NT, SA_SMALL_CTAS) spectrum_kernel_small(const V2Params P) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t *slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMSLOT);
  if (threadIdx.x < 32) tmem_alloc<C::TM_COLS>(slot);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tm_base = *slot;
  spectrum_body_small<C, MODE, FULL>(P, smem, tm_base);
  tmem_fence_before_sync();
  __syncthreads();   // warps finish at different times: release the allocation when all are done
  if (threadIdx.x < 32) tmem_dealloc<C::TM_COLS>(tm_base);
}

}  // namespace sa
