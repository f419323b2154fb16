// sa_spectrum_v3.cuh -- warp-specialised variant of the tuned family (V2Cfg<.., .., SPLIT = true>).
//
// What the ncu captures of the one-role kernel showed (profiles/README.md, round 2): it is bound by instruction
// issue (4 392 warp-instructions per 4096-point frame, 65 % issue utilisation with 20 warps/SM), and 43 % of those
// instructions are statistics work that every one of a group's four warps repeats or waits for: the cross-warp
// reduction barrier, the redundant histogram scan, the per-value candidate test, the arg-bin search and 17
// scattered 4-byte stores per thread.  Here the T threads of a frame group only
//   transform -> scale -> per-warp min / max / sum -> histogram atomics -> copy the scaled values to shared memory
// and ONE extra warp per group (the "statistics warp") finishes the frame while the group is already transforming
// its next one:
//   arg bins (src/spectrum.rs:496-499, 529-530 tie rules), exact median (src/spectrum.rs:515-524), average,
//   the 32-byte record, and the output row as 16-byte vector stores.
// The two roles meet at two mbarriers per group (full: values + partial results of a frame are in place, NW
// arrivals; empty: the statistics warp has consumed them, one arrival; the median window a frame is histogrammed
// against was published two frames earlier, so the record of frame k is finished while frame k + 1 is handed over) -- no CTA- or group-wide barrier is added, one (the statistics barrier) disappears from the
// transform warps, and nothing is computed twice.  A third mbarrier replaces that barrier's other job: it tells
// thread 0 of the group that every warp has read its last-pass inputs, so that the bulk copy of the next frame may
// overwrite the second exchange buffer.
//
// Everything that is arithmetic on spectrum values is shared with sa_spectrum_v2.cuh (same keys, same histogram
// slots, same tie rules); results are bit-identical to the one-role kernel except for nothing: the partial sums
// are formed in the same order.
#pragma once
#include "sa_spectrum_v2.cuh"

namespace sa {

// ---- bulk asynchronous copy shared -> global (TMA, 1-D), bulk-group completion ------------------------------
__device__ __forceinline__ void bulk_copy_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// the copies committed so far have finished READING shared memory (their source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// One warp's view of a histogram of 32 * HBPL buckets (hist[0] = keys below the window, hist[1 + b] = bucket b):
// HBPL consecutive buckets per lane, an inclusive warp scan of the lane sums.
template <int HBPL>
struct HistScan {
  uint32_t hb[HBPL], excl, incl, nbelow;
  __device__ __forceinline__ void load(uint32_t *hist, int lane, bool clear) {
    static_assert(HBPL == 2 || HBPL % 4 == 0, "vector loads of the lane's buckets");
    if constexpr (HBPL == 2) { const uint2 h = *reinterpret_cast<const uint2 *>(hist + 1 + 2 * lane); hb[0] = h.x; hb[1] = h.y; }
    else {
#pragma unroll
      for (int v = 0; v < HBPL / 4; v++) {
        const uint4 h = *reinterpret_cast<const uint4 *>(hist + 1 + HBPL * lane + 4 * v);
        hb[4 * v] = h.x; hb[4 * v + 1] = h.y; hb[4 * v + 2] = h.z; hb[4 * v + 3] = h.w;
      }
    }
    nbelow = hist[0];
    if (clear) {   // the lane's own buckets (same addresses as the loads); the caller clears hist[0] and hist[NB + 1]
      if constexpr (HBPL == 2) *reinterpret_cast<uint2 *>(hist + 1 + 2 * lane) = make_uint2(0u, 0u);
      else {
#pragma unroll
        for (int v = 0; v < HBPL / 4; v++) *reinterpret_cast<uint4 *>(hist + 1 + HBPL * lane + 4 * v) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    uint32_t local = 0;
#pragma unroll
    for (int b = 0; b < HBPL; b++) local += hb[b];
    incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    excl = nbelow + incl - local;        // keys in front of this lane's first bucket
  }
  // bucket of rank r (the first bucket whose inclusive count exceeds r), the keys in front of it, its count;
  // false when r lies outside the buckets
  __device__ __forceinline__ bool locate(uint32_t r, uint32_t &bucket, uint32_t &below, uint32_t &cnt, int lane) const {
    const uint32_t om = __ballot_sync(0xffffffffu, nbelow + incl > r);
    if (r < nbelow || om == 0u) return false;
    const int owner = __ffs(om) - 1;
    // every lane searches its own buckets; the owner's answer is broadcast
    uint32_t c = excl, mb = 0u, mc = 0u, ml = 0u;
    bool found = false;
#pragma unroll
    for (int b = 0; b < HBPL; b++) {
      const bool here = !found && (c + hb[b] > r);
      if (here) { mb = (uint32_t)(lane * HBPL + b); mc = hb[b]; ml = c; }
      found = found || here;
      c += hb[b];
    }
    bucket = __shfl_sync(0xffffffffu, mb, owner);
    cnt = __shfl_sync(0xffffffffu, mc, owner);
    below = __shfl_sync(0xffffffffu, ml, owner);
    return true;
  }
};

template <class C, int MODE, bool FULL>
__device__ __forceinline__ void stats_warp_v3(const V2Params &P, unsigned char *smem) {
  constexpr int M = C::M, G = C::G, NW = C::NW, NB = C::NBUCKET, HBITS = C::HBITS, HBPL = C::HBPL;
  constexpr int NCH = C::VALS_FLOATS / 4;        // 16-byte chunks of the value copy: M / 4 + 1
  static_assert((NCH - 1) % 64 == 0, "the pass over the copy takes two chunks per lane and step");
  constexpr int ITER = (NCH + 31) / 32;          // (slow path) one chunk per lane and step
  static_assert(NW <= 4, "per-warp slots of the hand-over");
  constexpr int NSW = C::NSW;
  const SpectrumParams &p = P.b;
  const int lane = threadIdx.x & 31;
  const int sw = (threadIdx.x - C::NT) >> 5;
  const int g = sw / NSW, slot = sw % NSW;       // group, hand-over slot (= parity of the frames this warp finishes when NSW = 2)
  unsigned char *gbase = smem + C::OFF_GROUPS + (size_t)g * C::G_BYTES;
  constexpr size_t OFF_HIST = C::G_BUF + (C::NBUF - 1) * C::G_BUFB;
  uint32_t *hist0 = reinterpret_cast<uint32_t *>(gbase + OFF_HIST) + slot * C::HSTRIDE;   // this warp's histogram
  uint32_t *misc = reinterpret_cast<uint32_t *>(gbase + OFF_HIST + C::G_HIST);
  uint32_t *list = reinterpret_cast<uint32_t *>(gbase + OFF_HIST + C::G_HIST + C::G_MISC) + 32 * slot;
  unsigned char *sync_base = gbase + OFF_HIST + C::G_HIST + C::G_MISC + C::G_LIST + C::G_STAGE;
  uint64_t *mb_full = reinterpret_cast<uint64_t *>(sync_base) + slot;
  uint64_t *mb_empty = reinterpret_cast<uint64_t *>(sync_base + 16) + slot;
  uint32_t *vbuf = reinterpret_cast<uint32_t *>(sync_base + C::G_SYNC) + slot * C::VALS_FLOATS;   // value bit patterns
  uint32_t *m_lcount = &misc[MS_LCOUNT + slot], *m_flags = &misc[MS_FLAGS + slot];
  if (NSW == 1 ? lane < 4 : lane < 2) misc[MS_WIN + (NSW == 1 ? lane : 2 * slot + lane)] = 0u;   // no window for the first frames
  if (lane == 4) { *m_lcount = 0u; *m_flags = 0u; }
  __syncthreads();   // pairs with the barrier that ends the transform threads' set-up

  const uint32_t lo = p.bin_lo, hi = p.bin_hi, count = hi - lo + 1;
  const uint32_t ra = (count & 1u) ? count / 2 : count / 2 - 1, rb = count / 2;   // middle ranks (0-based)
  const bool want_median = (p.stats_mask & SA_STATS_MEDIAN) != 0;
  constexpr uint32_t GW_INIT = 1u << 20, GW_MIN = 1u << SA_GW_MIN_LOG2;
  uint32_t gw = 0u;            // half-width of the median window (key units); 0: no guess yet
  // windows in flight: [par] is the one the transform warps used for the frame of parity par (published two
  // frames earlier, see the end of the loop)
  uint32_t win_lo[2] = {0u, 0u};
  int win_sh[2] = {0, 0};
  bool win_on[2] = {false, false};
  uint32_t fphase = 0;
  int par = NSW == 1 ? 0 : slot;
  const u64 fstep = (u64)gridDim.x * G;
  constexpr uint32_t INF_BITS = 0x7f800000u;
#if SA_EXP & 0x10000
  uint32_t tq_post = 0u;
#endif
  auto clear_hist = [&](uint32_t *hist_base) {   // HSTRIDE words, 16-byte aligned
    for (int b = lane; b < C::HSTRIDE / 4; b += 32) reinterpret_cast<uint4 *>(hist_base)[b] = make_uint4(0u, 0u, 0u, 0u);
  };
  static_assert(C::HSTRIDE % 4 == 0, "vector clears of a histogram");

#pragma unroll 1
  for (u64 cur = (u64)blockIdx.x * G + g + (NSW == 1 ? 0 : (u64)slot * fstep); cur < p.n_frames;
       cur += NSW * fstep, par ^= (NSW == 1 ? 1 : 0)) {
    uint32_t *hist_base = hist0;   // this warp's histogram: clean again before empty(k), i.e. before the atomics of the next frame that uses it
    uint32_t *hist = hist_base + 3;                      // hist[0] below, hist[1 + b], hist[NB + 1] above
    const uint32_t w_lo = par ? win_lo[1] : win_lo[0];
    const int w_sh = par ? win_sh[1] : win_sh[0];
    const bool w_on = par ? win_on[1] : win_on[0];
    // position of the frame's bins inside the copy: bin b at index b + o (o = output row address / 4 mod 4)
    const u64 row = (reinterpret_cast<uintptr_t>(p.out_vals) >> 2) + cur * p.out_stride - lo;
    const uint32_t o = (uint32_t)(row & 3u);
    float *grow = p.out_vals + cur * p.out_stride;      // row of this frame ...
    grow -= (lo + o);                                    // ... indexed by position in the copy: 16-byte aligned
    const uint32_t s_lo = lo + o, s_hi = hi + o;         // positions of the kept bins
    const uint32_t a_lo = (s_lo + 3u) & ~3u, a_hi = (s_hi + 1u) & ~3u;   // the 16-byte aligned part of the row

#if SA_EXP & 0x10000   // timing experiment: cycle counts into the record instead of the arg bins
    const long long tq0 = clock64();
#endif
    mbar_wait_sleep(mb_full, fphase);
    fphase ^= 1u;
#if SA_EXP & 0x10000
    const long long tq1 = clock64();
#endif
#if SA_EXP & 0x1000   // timing experiment: no statistics work at all
    __syncwarp();
    if (lane == 0) mbar_arrive(mb_empty);
    continue;
#endif

    // ---- per-warp partial results (one vector load per array) -> frame minimum / maximum / sum ------------------
    uint32_t pmn[NW], pmx[NW], psum[NW], bmn[NW], bmx[NW];
    if constexpr (NW == 4) {
      const uint4 a0 = *reinterpret_cast<const uint4 *>(&misc[MS_KMIN + 8 * slot]), a1 = *reinterpret_cast<const uint4 *>(&misc[MS_KMAX + 8 * slot]);
      const uint4 a2 = *reinterpret_cast<const uint4 *>(&misc[MS_SUM + 8 * slot]);
      const uint4 a3 = *reinterpret_cast<const uint4 *>(&misc[MS_BMIN + 4 * slot]), a4 = *reinterpret_cast<const uint4 *>(&misc[MS_BMAX + 4 * slot]);
      pmn[0] = a0.x; pmn[1] = a0.y; pmn[2] = a0.z; pmn[3] = a0.w;
      pmx[0] = a1.x; pmx[1] = a1.y; pmx[2] = a1.z; pmx[3] = a1.w;
      psum[0] = a2.x; psum[1] = a2.y; psum[2] = a2.z; psum[3] = a2.w;
      bmn[0] = a3.x; bmn[1] = a3.y; bmn[2] = a3.z; bmn[3] = a3.w;
      bmx[0] = a4.x; bmx[1] = a4.y; bmx[2] = a4.z; bmx[3] = a4.w;
    } else {
#pragma unroll
      for (int w = 0; w < NW; w++) {
        pmn[w] = misc[MS_KMIN + 8 * slot + w]; pmx[w] = misc[MS_KMAX + 8 * slot + w]; psum[w] = misc[MS_SUM + 8 * slot + w];
        bmn[w] = misc[MS_BMIN + 4 * slot + w]; bmx[w] = misc[MS_BMAX + 4 * slot + w];
      }
    }
    const uint32_t flags = *m_flags;
    // the up to three values on either side of the 16-byte aligned part of the row: loaded now, stored further down
    const uint32_t edge_s = lane < 3 ? s_lo + lane : max(a_hi, a_lo) + (lane - 3);
    const bool edge_ok = lane < 6 && (lane < 3 ? (edge_s < a_lo && edge_s <= s_hi) : (edge_s <= s_hi));
    const uint32_t edge_v = vbuf[edge_ok ? edge_s : 0u];
    uint32_t kmn = pmn[0], kmx = pmx[0];
    float sum = __uint_as_float(psum[0]);
#pragma unroll
    for (int w = 1; w < NW; w++) {
      kmn = min(kmn, pmn[w]);
      kmx = max(kmx, pmx[w]);
      sum += __uint_as_float(psum[w]);
    }
    const bool nonfinite = flags != 0u;     // set by the transform warps together with the NaN / Inf classification
    const bool all_equal = (kmx == kmn);

    // ---- median, fast path: bucket(s) of the middle rank(s) inside the guessed window ----------------------
    bool collect = false;        // the pass below gathers the keys of [c_lo, c_lo + c_w] into `list`
    bool need_slow = false, win_failed = false;
    uint32_t c_lo = 0u, c_w = 0u, c_below = 0u, c_cnt = 0u;
    uint32_t ka = kmn, kb = kmn;
    const bool want_sel = want_median && !(all_equal && !nonfinite);
    if (want_sel) {
      if (w_on && !nonfinite) {
        HistScan<HBPL> hs;
        hs.load(hist, lane, true);       // ... and clears it: this histogram is used again by the next frame
        __syncwarp();
        if (lane == 0) { hist[0] = 0u; hist[NB + 1] = 0u; }
        uint32_t ba, bb, cnt_a, cnt_b, below_b;
        bool ok = hs.locate(ra, ba, c_below, cnt_a, lane);
        if (count & 1u) { bb = ba; cnt_b = cnt_a; }
        else ok = hs.locate(rb, bb, below_b, cnt_b, lane) && ok;
        if (!ok) {
          need_slow = true; win_failed = true;
        } else {
          c_cnt = ba == bb ? cnt_a : cnt_a + cnt_b;          // buckets between two adjacent ranks are empty
          const uint32_t lo_c = w_lo + (ba << w_sh), wd_c = ((bb - ba + 1u) << w_sh) - 1u;
          if (c_cnt > 32u) { need_slow = true; c_lo = lo_c; c_w = wd_c; }   // crowded: exact loop below, narrowed
          else if (w_sh == 0) { ka = lo_c; kb = w_lo + bb; }                // buckets are single keys
          else { collect = true; c_lo = lo_c; c_w = wd_c; }
        }
      } else {
        need_slow = true;
      }
    }
    if (w_on && !(want_sel && !nonfinite)) clear_hist(hist_base);   // filled by the transform warps but not scanned
    if (edge_ok) grow[edge_s] = __uint_as_float(edge_v);
#if SA_EXP & 0x10000
    const long long tqa = clock64();   // partials + scan done
#endif

    // ---- arg bins (src/spectrum.rs:496-499, 529-530: stable sort -> min at the lowest, max at the highest bin among
    //      equals): the transform warps said which of their lanes hold the warp's extreme; look at those threads'
    //      sixteen (thread 0: seventeen) bins in the copy, one bin per lane ------------------------------------------
    uint32_t argmin = 0xffffffffu, argmax = 0u;
    if (p.stats_mask & SA_STATS_MINMAX) {
      // the threads that hold the frame's extremes: (warp, lane mask) pairs, almost always one thread each
      uint32_t mn_rest = 0u, mx_rest = 0u;       // bit w: warp w holds the minimum / maximum
#pragma unroll
      for (int w = 0; w < NW; w++) { mn_rest |= (pmn[w] == kmn) ? 1u << w : 0u; mx_rest |= (pmx[w] == kmx) ? 1u << w : 0u; }
      auto pick = [&](const uint32_t (&arr)[NW], int w) -> uint32_t {
        uint32_t r = arr[0];
#pragma unroll
        for (int q = 1; q < NW; q++) r = (w == q) ? arr[q] : r;
        return r;
      };
      uint32_t mn_lanes = 0u, mx_lanes = 0u;
      int mn_w = 0, mx_w = 0;
      // one step looks at one thread for the minimum and one for the maximum (lanes 0..15: the sixteen bins of a
      // thread; thread 0 also owns bin M/2, lane 16)
      while (mn_rest | mn_lanes | mx_rest | mx_lanes) {
        if (mn_lanes == 0u && mn_rest) { mn_w = __ffs(mn_rest) - 1; mn_rest &= mn_rest - 1u; mn_lanes = pick(bmn, mn_w); }
        if (mx_lanes == 0u && mx_rest) { mx_w = __ffs(mx_rest) - 1; mx_rest &= mx_rest - 1u; mx_lanes = pick(bmx, mx_w); }
        const bool do_mn = mn_lanes != 0u, do_mx = mx_lanes != 0u;
        const int j1 = 32 * mn_w + (__ffs(mn_lanes) - 1), j2 = 32 * mx_w + (__ffs(mx_lanes) - 1);
        mn_lanes &= mn_lanes - 1u;
        mx_lanes &= mx_lanes - 1u;
        const uint32_t bin1 = lane < 16 ? v2_binof<C>(do_mn ? j1 : 0, lane) : (uint32_t)(M / 2);
        const uint32_t bin2 = lane < 16 ? v2_binof<C>(do_mx ? j2 : 0, lane) : (uint32_t)(M / 2);
        const uint32_t k1 = v2_key<MODE>(__uint_as_float(vbuf[bin1 + o])), k2 = v2_key<MODE>(__uint_as_float(vbuf[bin2 + o]));
        const bool act1 = do_mn && (lane < 16 || (lane == 16 && j1 == 0)), act2 = do_mx && (lane < 16 || (lane == 16 && j2 == 0));
        const bool hit1 = act1 && (bin1 - lo) <= (hi - lo) && k1 == kmn, hit2 = act2 && (bin2 - lo) <= (hi - lo) && k2 == kmx;
        argmin = min(argmin, __reduce_min_sync(0xffffffffu, hit1 ? bin1 : 0xffffffffu));
        argmax = max(argmax, __reduce_max_sync(0xffffffffu, hit2 ? bin2 : 0u));
      }
    }
#if SA_EXP & 0x10000
    const long long tqb = clock64();   // arg bins done
    long long tqc = tqb;
#endif

    // ---- one pass over the copy: the output row as 16-byte vector stores (rows of M + 1 floats are only 4-byte
    //      aligned; the copy is laid out so that its 16-byte chunks coincide with those of the row) and, on the way,
    //      the candidates of the median ---------------------------------------------------------------------------
    {
      const uint4 *v4 = reinterpret_cast<const uint4 *>(vbuf);
      uint4 *g4 = reinterpret_cast<uint4 *>(grow);
      auto keys_of = [&](uint4 w) -> uint4 {
        if constexpr (MODE == SCALE_LOG10) {
          w.x = ordered_key(__uint_as_float(w.x)); w.y = ordered_key(__uint_as_float(w.y));
          w.z = ordered_key(__uint_as_float(w.z)); w.w = ordered_key(__uint_as_float(w.w));
        }
        return w;
      };
      const uint32_t ch_lo = a_lo >> 2, ch_n = a_hi > a_lo ? (a_hi - a_lo) >> 2 : 0u;   // chunks [ch_lo, ch_lo + ch_n) lie inside the row
      const uint32_t t_lo = collect ? c_lo : 0xffffffffu, t_w = collect ? c_w : 0u;
      // step 1: store; which of this lane's chunks hold a candidate (bit i: chunk lane + 32 i)
      uint32_t hits = 0u;
#pragma unroll 2
      for (int i = 0; i < (NCH - 1) / 64; i++) {
        const int c0 = lane + 64 * i, c1 = c0 + 32;
        const uint4 w0 = v4[c0], w1 = v4[c1];
        if ((uint32_t)c0 - ch_lo < ch_n) g4[c0] = w0;
        if ((uint32_t)c1 - ch_lo < ch_n) g4[c1] = w1;
        const uint32_t h0 = chunk_hit(keys_of(w0), t_lo, t_w), h1 = chunk_hit(keys_of(w1), t_lo, t_w);
        hits |= (h0 << (2 * i)) | (h1 << (2 * i + 1));
      }
      if (lane == 0) {   // the last chunk
        if ((uint32_t)(NCH - 1) - ch_lo < ch_n) g4[NCH - 1] = v4[NCH - 1];
        hits |= 1u << ((NCH - 1) / 32);     // looked at without the test
      }
#if SA_EXP & 0x10000
      tqc = clock64();   // step 1 done
#endif
      // step 2: the lanes that have such chunks look at them again, all at once (a few values per frame, spread
      // over the lanes: one round, rarely two).  Dropped bins and the slots around the frame are excluded by index.
      if (collect) {
        while (hits) {
          const int c = lane + 32 * (__ffs(hits) - 1);
          hits &= hits - 1u;
          const uint4 k = keys_of(v4[c]);
          const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
          bool ce[4];
          uint32_t n = 0;
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const uint32_t bin = 4u * (uint32_t)c + (uint32_t)e - o;
            ce[e] = (bin - lo) <= (hi - lo) && (kk[e] - c_lo) <= c_w;
            n += ce[e] ? 1u : 0u;
          }
          if (n) {
            uint32_t at = atoms_add(m_lcount, n);     // one reservation per lane and round
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (ce[e]) list[at++ & 31u] = kk[e];
          }
        }
        __syncwarp();
        if (*m_lcount != c_cnt) { collect = false; need_slow = true; c_cnt = 0u; }   // cannot happen; the loop below is always right
      }
    }
#if SA_EXP & 0x10000
    const long long tqd = clock64();   // candidates gathered
#endif
    // ---- exact selection without a usable window (first frames, window miss, crowded bucket, non-finite frame):
    //      narrowing histogram passes over the copy, this warp alone (uses the histogram just cleared) -------------
    if (need_slow) {
      uint32_t klo = kmn, width = kmx - kmn, rka = ra, rkb = rb;
      uint32_t cand_lo = 0u, cand_w = 0xffffffffu;        // keys still in play: (key - cand_lo) <= cand_w
      if (c_cnt > 32u) { klo = c_lo; width = c_w; cand_lo = c_lo; cand_w = c_w; rka = ra - c_below; rkb = rb - c_below; }
      const uint4 *v4 = reinterpret_cast<const uint4 *>(vbuf);
      // f(key) for every kept value of the copy
      auto for_each_key = [&](auto &&f) {
#pragma unroll 1
        for (int i = 0; i < ITER; i++) {
          const int c = lane + 32 * i;
          if (c < NCH) {
            const uint4 w = v4[c];
            const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const uint32_t bin = 4u * (uint32_t)c + (uint32_t)e - o;
              if ((bin - lo) <= (hi - lo)) f(v2_key<MODE>(__uint_as_float(wv[e])));
            }
          }
        }
      };
      __syncwarp();
#pragma unroll 1
      for (;;) {
        const int sh = max(0, (32 - __clz(width | 1u)) - HBITS);
        for_each_key([&](uint32_t k) {
          if ((k - cand_lo) <= cand_w) atomicAdd(&hist[hist_slot<MODE, NB>(k, klo, sh)], 1u);
        });
        __syncwarp();
        HistScan<HBPL> hs;
        hs.load(hist, lane, true);   // clean again for the next pass / frame
        __syncwarp();
        if (lane == 0) { hist[0] = 0u; hist[NB + 1] = 0u; }
        __syncwarp();
        uint32_t ba, bb, cnt_a, cnt_b, below_a, below_b;
        bool ok = hs.locate(rka, ba, below_a, cnt_a, lane);
        ok = hs.locate(rkb, bb, below_b, cnt_b, lane) && ok;
        if (!ok) {
          // only possible when a narrowed pass started from an inconsistent window: restart on the full range
          klo = kmn; width = kmx - kmn; rka = ra; rkb = rb; cand_lo = 0u; cand_w = 0xffffffffu;
          continue;
        }
        if (sh == 0) { ka = klo + ba; kb = klo + bb; break; }      // buckets are single keys
        const uint32_t lo_a = klo + (ba << sh), lo_b = klo + (bb << sh), bw = (1u << sh) - 1u;
        if (ba != bb) {
          // the two middle ranks fall into different buckets: s[ra] is the largest key of bucket a, s[rb] the
          // smallest key of bucket b
          uint32_t xa = 0u, xb = 0xffffffffu;
          for_each_key([&](uint32_t k) {
            if ((k - lo_a) <= bw) xa = max(xa, k);
            if ((k - lo_b) <= bw) xb = min(xb, k);
          });
          ka = __reduce_max_sync(0xffffffffu, xa);
          kb = __reduce_min_sync(0xffffffffu, xb);
          break;
        }
        if (cnt_a <= 32u) {   // gather the bucket, rank it below
          if (lane == 0) *m_lcount = 0u;
          __syncwarp();
          for_each_key([&](uint32_t k) {
            if ((k - lo_a) <= bw) list[atoms_inc(m_lcount) & 31u] = k;
          });
          __syncwarp();
          collect = true; c_cnt = cnt_a; c_below = below_a + (ra - rka);   // ranks relative to the whole frame
          break;
        }
        // crowded bucket (many near-equal values): narrow the candidate set and repeat
        klo = lo_a; width = bw; cand_lo = lo_a; cand_w = bw; rka -= below_a; rkb -= below_a;
      }
    }
    uint32_t mine = 0xffffffffu;
    if (collect && lane < (int)c_cnt) mine = list[lane];
    if (lane == 0) { *m_flags = 0u; *m_lcount = 0u; }
    __syncwarp();
    // ---- values, partial results and histogram of this frame are consumed: the transform warps may hand over
    //      the next frame while this warp finishes the record --------------------------------------------------
    // (one statistics warp per group: the window a frame is histogrammed against was published two frames earlier, so
    //  the hand-over may be released before the median is known.  Two warps: this warp's next frame is the next user
    //  of this slot AND of the window published below, so the release comes after the publication, at the end.)
    if (NSW == 1 && lane == 0) mbar_arrive(mb_empty);
#if SA_EXP & 0x10000
    const long long tq2 = clock64();
#endif

    // ---- rank the gathered candidates (one per lane; equal keys are equal values) ----------------------------
    if (collect) {
      uint32_t rank = 0;
#pragma unroll 4
      for (uint32_t c = 0; c < c_cnt; c++) {
        const uint32_t t = __shfl_sync(0xffffffffu, mine, c);
        rank += (t < mine || (t == mine && c < (uint32_t)lane)) ? 1u : 0u;
      }
      const uint32_t ma = __ballot_sync(0xffffffffu, lane < (int)c_cnt && rank == ra - c_below);
      const uint32_t mb = __ballot_sync(0xffffffffu, lane < (int)c_cnt && rank == rb - c_below);
      ka = __shfl_sync(0xffffffffu, mine, (__ffs(ma) - 1) & 31);
      kb = __shfl_sync(0xffffffffu, mine, (__ffs(mb) - 1) & 31);
    }

    // ---- the record ---------------------------------------------------------------------------------------
    if (lane == 0) {
      const bool mm = p.stats_mask & SA_STATS_MINMAX;
      uint4 r0, r1;
      r0.x = mm ? __float_as_uint(v2_unkey<MODE>(kmn)) : 0u;
      r0.y = mm ? argmin : 0u;
      r0.z = mm ? __float_as_uint(v2_unkey<MODE>(kmx)) : 0u;
      r0.w = mm ? argmax : 0u;
      r1.x = (p.stats_mask & SA_STATS_AVERAGE) ? __float_as_uint(__fdiv_rn(sum, (float)count)) : 0u;
      const float med = (count & 1u) ? v2_unkey<MODE>(kb)
                                     : __fdiv_rn(__fadd_rn(v2_unkey<MODE>(ka), v2_unkey<MODE>(kb)), 2.0f);
      r1.y = want_median ? __float_as_uint(med) : 0u;
      r1.z = (flags & 1u) ? SA_ERR_NAN_VALUES : (flags & 2u) ? SA_ERR_INFINITY_VALUES : (flags & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
      r1.w = 0u;
#if SA_EXP & 0x10000
      r0.y = (uint32_t)(tq1 - tq0); r0.w = (uint32_t)(tq2 - tq1); r1.w = tq_post;
      r0.x = (uint32_t)(tqa - tq1); r0.z = (uint32_t)(tqb - tqa); r1.x = (uint32_t)(tqc - tqb); r1.y = (uint32_t)(tqd - tqc);
#endif
      uint4 *rec = reinterpret_cast<uint4 *>(&p.stats[cur]);
      rec[0] = r0;
      rec[1] = r1;
    }

    // ---- window for the frame after the next one (the next one is already being histogrammed against the
    //      window published a frame ago): same adaptation as the one-role kernel ---------------------------------
    if (want_median) {
      if (gw == 0u) gw = GW_INIT;
      else if (win_failed) gw = (gw < (1u << 28)) ? gw << SA_GW_GROW : gw;
      else if (!need_slow) gw = max(gw - (gw >> SA_GW_SHRINK), GW_MIN);
      const uint32_t nlo = kb > gw ? kb - gw : 0u;
      const int nsh = max(0, (32 - __clz((2u * gw) | 1u)) - HBITS);
      if (par) { win_lo[1] = nlo; win_sh[1] = nsh; win_on[1] = true; } else { win_lo[0] = nlo; win_sh[0] = nsh; win_on[0] = true; }
      // read by the transform warps after they have waited for empty(next frame), i.e. after this store
      if (lane == 0) *reinterpret_cast<uint2 *>(&misc[MS_WIN + 2 * par]) = make_uint2(nlo, (uint32_t)nsh | 0x100u);
    }
    if (NSW > 1 && lane == 0) mbar_arrive(mb_empty);
#if SA_EXP & 0x10000
    tq_post = (uint32_t)(clock64() - tq2);
#endif
  }
}

template <class C, int MODE, bool FULL>
__global__ void __launch_bounds__(C::NT_ALL, 1) spectrum_kernel_v3(const V2Params P) {
  static_assert(C::SPLIT, "spectrum_kernel_v3 runs the split configuration");
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t *slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMSLOT);
  if (threadIdx.x < 32) tmem_alloc<C::TM_COLS>(slot);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tm_base = *slot;
  if (threadIdx.x < C::NT) spectrum_body_v2<C, MODE, FULL>(P, smem, tm_base);
  else stats_warp_v3<C, MODE, FULL>(P, smem);
  tmem_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<C::TM_COLS>(tm_base);
}

}  // namespace sa
