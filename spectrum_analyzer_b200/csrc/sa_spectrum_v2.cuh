// sa_spectrum_v2.cuh -- tuned fused kernel family for N in {1024, 2048, 4096, 8192, 16384} (T = N/32
// threads per frame, a multiple of the warp size).
//
// Same math and same outputs as sa_spectrum_kernel.cuh (the generic family), restructured around what
// the ncu profiles showed (profiles/README.md): DRAM traffic already equals the algorithmic bytes; the
// kernel is bound by per-warp latency chains, the shared-memory/LSU pipe and instruction issue.
//
//  * CTA = G independent frame groups of T threads (G*T = 512): each group walks its own frames and
//    synchronises with its own named barrier (bar.sync id, T), so groups overlap each other's
//    barrier/latency phases; 16 warps/SM at <= 128 registers.  Warp roles rotate from group to group
//    so that the extra work of a group's first and last warp is spread over the four SM sub-partitions.
//  * per-thread constants (window multipliers src/windows.rs, pass twiddles, recombination twiddles:
//    112 words per thread) live in TENSOR MEMORY and are read with tcgen05.ld (sa_tmem.cuh): TMEM has
//    its own datapath, so a quarter of the former shared-memory traffic left the LSU pipe.
//  * the next frame's samples are STAGED IN SHARED MEMORY BY A BULK ASYNCHRONOUS COPY (cp.async.bulk,
//    one per frame, completion on an mbarrier) issued right after the first barrier of the current
//    frame.  A register prefetch (ld.global ahead of the statistics) was slower than no prefetch at
//    all: ptxas shares its six scoreboards between those loads and the LDS/SHFL/ATOMS of the statistics
//    phase, which then waits for DRAM.  Frames that do not start on a 16-byte boundary are copied as
//    their aligned superset and read at an offset, so any hop / base alignment takes this family.
//  * two exchange buffers per group, alternated on every exchange: one barrier per exchange (one-warp
//    groups, N = 1024, reuse a single buffer between __syncwarp()s); mirrored last-pass butterflies
//    meet in one thread (N = 8192: in two lanes of one warp), so the real-FFT recombination needs no
//    shared-memory exchange.
//  * magnitude = sqrt.approx(|.|^2) * c, where c folds the real-FFT 1/2 and a power-of-two scaling
//    divisor (divide_by_N always, divide_by_N_sqrt when log2 N is even) into one exact multiply;
//    other divisors use an exactly-rounded FMA division.
//  * statistics: u32 min/max on the value bit patterns (3-input VIMNMX), redux.sync across the warp,
//    arg-bins resolved only by the threads that hold the extreme value (stable-sort tie rules of
//    src/spectrum.rs:496-499,529-530 via atomicMin / atomicMax on the bin index).  The record of a
//    frame is written one barrier later by the group's last warp; the spectrum values are stored
//    after the statistics so that their dependent SHFL/LDS chains do not queue behind 17 global
//    stores per thread in the memory pipe.
//  * NaN/Inf validation (src/lib.rs:168-173) costs nothing on clean frames: any non-finite input
//    makes every output bin non-finite, so the maximum flags it and only then the group re-reads
//    its samples to classify NaN vs Inf vs overflow.
//  * median (src/spectrum.rs:515-524): exact selection.  The keys are histogrammed (128 buckets +
//    below/above, unconditional shared-memory atomics) against a window guessed from the previous
//    frame's median, fused with the min/max/sum pass; after the one statistics barrier every warp
//    scans the histogram (four buckets per lane, one warp scan, the owner lane's answer in one packed
//    shuffle).  Candidates of the selected bucket: c_i = min(key_i - lo, width) (one VIADDMNMX per value),
//    a 3-input-minimum tree gives the thread's candidate and the sum of the c_i tells whether it is the
//    only one; it goes into the thread's own word of a table by an unconditional store (more of the same
//    thread: overflow list).  ONE warp picks the median from the table right after the first exchange
//    barrier of the group's NEXT frame (up to three keys: minimum / maximum / sum by three warp
//    reductions; more: stripping minima).  A missed window, an even count, a crowded bucket or a
//    non-finite frame take the barrier-synchronised exact loop (full-range histogram, narrowing passes).
//    (Round 1 appended candidates to a list and let the warp of the last pusher rank it: that warp was
//    the group's latest by construction and its chain held the group at its next barrier.)
//  * the window multiply is packed (FMUL2, contracted by ptxas into the first butterfly adds) and a
//    power-of-two scaling factor rides on the window multipliers (SA_OPT); without statistics the next
//    frame's bulk copy is issued by the last warp that has used its last-pass inputs (EARLY);
//    N = 8192 / 16384 write the first exchange with STS.128 into rows padded to 18 and reduce the
//    per-warp partial results inside a warp (their groups are limited by the memory-instruction queue).
//    DESIGN.md section 4.6 lists what each of these bought and what was measured and rejected.
#pragma once
#include "sa_common.cuh"
#include "sa_dft.cuh"
#include "sa_tmem.cuh"

// SA_EXP: timing experiments only (profiles/microbench/run_variants.sh); results are wrong when nonzero
#ifndef SA_EXP
#define SA_EXP 0
#endif

namespace sa {

enum ScaleMode { SCALE_MUL = 0, SCALE_DIV = 1, SCALE_ZERO_ONE = 2, SCALE_LOG10 = 3 };

struct V2Params {
  SpectrumParams b;
  const float2 *tw2;   // [15][16]           w_256^(k*p)
  const float2 *tw3;   // [R3-1][256]        w_M^(k*p)
  const float2 *twr;   // [M/2]              W_N^k  (recombination)
  const float2 *tw4;   // [R4-1][4096]       w_M^(k*p)  (N = 16384 only: fourth pass)
  float cmul;          // SCALE_MUL: 0.5/div (exact power of two); else 0.5
  float div;           // SCALE_DIV: divisor d
  float rdiv;          // SCALE_DIV: RN(1/d)
  unsigned stagger_ns; // start offset between the frame groups of a CTA (see the kernel)
};

template <int LOG2N_, int G_, int NSW_ = 0>
struct V2Cfg {
  static constexpr bool SPLIT_ = NSW_ > 0;
  static constexpr int LOG2N = LOG2N_, N = 1 << LOG2N_, M = N / 2, E = 16, T = M / E, G = G_, NT = T * G_;
  // SPLIT: warp-specialised variant (sa_spectrum_v3.cuh).  The T threads of a group only transform, scale and
  // histogram a frame; one extra "statistics warp" per group (threads NT .. NT + 32 G) finishes the frame from a
  // copy of the scaled values in shared memory while the group already transforms its next frame.
  // NSW statistics warps per group: with two, consecutive frames of a group go to alternate warps, each with its own
  // copy of the values, histogram and partial-result slots, so that a statistics warp has two frame times per frame.
  static constexpr bool SPLIT = SPLIT_;
  static constexpr int NSW = NSW_;
  static_assert(NSW_ == 0 || NSW_ == 1 || NSW_ == 2, "statistics warps per group");
  static constexpr int NT_ALL = NT + 32 * G_ * NSW_;
  static constexpr int NW = T / 32;                  // warps per group
  static constexpr bool FOUR = (M > 4096);           // N = 16384: passes 16,16,16,R4 instead of 16,16,R3
  static constexpr int R3 = FOUR ? 16 : M / 256;     // radix of the third pass (last pass unless FOUR)
  static constexpr int NB3 = 16 / R3;                // butterflies per thread in the third pass
  static constexpr int R4 = FOUR ? M / 4096 : 2;     // radix of the fourth (last) pass when FOUR
  static constexpr int NB4 = 16 / R4;
#ifndef SA_PAD18
#define SA_PAD18 1
#endif
  // first exchange: every thread writes 16 consecutive values and reads stride-T ones.  One pad element per 16 (17 j + k)
  // keeps both sides conflict-free with 8-byte accesses.  N = 16384 (one 16-warp group per SM, its stalls led by the
  // memory-instruction queue) pads with TWO per 16 (18 j + k): the rows stay 16-byte aligned and the writes are eight
  // STS.128 instead of sixteen STS.64 (quarter-warps hit eight distinct 16-byte columns): cfg4 0.391 -> 0.400, without
  // statistics 0.592 -> 0.636; N = 8192 +1 %.  At N = 4096 that measured 0.662 vs 0.664, so the smaller sizes keep 17.
  static constexpr bool PAD2 = SA_PAD18 && (M >= 4096);
  static constexpr int PADE = PAD2 ? 8 : 16;
  static constexpr int BUF = M + M / PADE;           // padded float2 elements of one exchange buffer
  // median histogram: 64 buckets per pass (measured with the warp-cooperative ranking: 32 buckets 1.49e8,
  // 64: 1.57e8, 128: 1.56e8, 256: 1.53e8 spectra/s at cfg2; -DSA_HBITS=n rebuilds with another count)
#ifndef SA_HBITS
#define SA_HBITS 7
#endif
  // the warp-specialised variant scans the histogram once per frame (not once per warp) and ranks the selected
  // bucket in its one statistics warp: more, finer buckets pay there (fewer candidates to gather and rank)
#ifndef SA_HBITS_SPLIT
#define SA_HBITS_SPLIT 8
#endif
  static constexpr int HBITS = SPLIT_ ? SA_HBITS_SPLIT : SA_HBITS, NBUCKET = 1 << HBITS;
  static constexpr int HBPL = NBUCKET / 32;               // buckets per lane of the scanning warp
  // the per-thread constants (window multipliers, pass twiddles, recombination
  // twiddles: 112 words per thread) in tensor memory, read with tcgen05.ld (sa_tmem.cuh): column
  // layout [0,32) window | [32,62) pass-2 twiddles | [64,94) pass-3 twiddles | [96,112) recombination |
  // [112,128) pass-4 twiddles (N = 16384).
  // A thread's lane is 32*(warp%4)+lane_id; groups wider than 128 threads use T/128 column sets.
  // Warp roles are rotated from group to group (thread 0 of a group does extra work -- the record, bin
  // M/2 -- and would otherwise always sit on the same SM sub-partition): group g's logical warp w is
  // physical warp (w - rot(g)) mod NW, rot(g) = (g / GPQ) % NROT.  One table copy per rotation.
  static constexpr int NROT = T == 128 ? 4 : (T == 64 || T == 256) ? 2 : 1;
  static constexpr int GPQ = T >= 128 ? 1 : 128 / T;          // groups per 4 physical warps
  static constexpr int TM_JSETS = T > 128 ? T / 128 : 1;      // column sets needed to cover j >= 128
  static constexpr int TM_SETS = TM_JSETS * NROT, TM_SET_COLS = 128, TM_COLS = TM_SETS * TM_SET_COLS;
  static_assert(TM_COLS <= 512 && G_ / GPQ >= NROT, "TMEM table layout");
  static constexpr int TM_WIN = 0, TM_TW2 = 32, TM_TW3 = 64, TM_TWR = 96, TM_TW4 = 112;   // TM_TW4: N = 16384 only
  static constexpr int NTW2 = 15 * 16, NTW3 = (R3 - 1) * 256, NTWR = M / 2, NTW4 = FOUR ? (R4 - 1) * 4096 : 0;
  static_assert(T % 32 == 0 && T >= 32 && T <= 512, "v2 covers N = 1024..16384");
  static_assert(NT <= 1024 && (T == 32 || G_ <= 15), "named barriers 1..15");
  // shared memory map (bytes)
  static constexpr size_t OFF_TMSLOT = 0;                                  // TMEM base address (16 B slot)
  static constexpr size_t OFF_GROUPS = 16;
  static constexpr size_t G_BUF = sizeof(float2) * BUF;
  // two exchange buffers alternated per exchange (one barrier each); one-warp groups (N = 1024) synchronise
  // with __syncwarp, which is cheap enough to reuse a single buffer and leave room for the staging buffer
  static constexpr int NBUF = T == 32 ? 1 : 2;
  static constexpr int HSTRIDE = NBUCKET + 8;                          // [below | NBUCKET buckets | above | pad]
  static constexpr size_t G_HIST = sizeof(uint32_t) * HSTRIDE * (SPLIT_ ? NSW_ : 2);   // two histograms, alternated per frame (SPLIT: one per statistics warp)
  static constexpr size_t G_MISC = sizeof(uint32_t) * 96;
  // candidate list (32 keys; the split variant: one per statistics warp) + one word per thread of the group (DEFRANK)
  static constexpr size_t G_LIST = sizeof(uint32_t) * (32 * (NSW_ > 1 ? NSW_ : 1) + (SPLIT_ ? 0 : T));
  // the next frame's samples are staged in shared memory by a bulk asynchronous copy
  // (one per frame, issued by thread 0 of the group) instead of being loaded into registers
  // N = 2048 / 4096 / 8192: the staged frame lives in the second exchange buffer (free between the last-pass reads of one frame
  // and the second exchange of the next), so a group needs 37 KB instead of 54 KB and five or six groups fit on
  // an SM (94 / 80 registers per thread; five measured best); the copy is issued after the statistics barrier
  // instead of after the first one
  static constexpr bool ALIAS = !FOUR && T >= 64;   // N = 2048, 4096, 8192 (two exchanges per frame, >= 2 warps per group)
  static constexpr size_t G_BUFB = ALIAS ? sizeof(float2) * M + 32 : G_BUF;   // ALIAS: unpadded exchange / frame + slack
  static constexpr size_t G_STAGE = ALIAS ? 16 : sizeof(float) * N + 32;      // samples (+16 B of alignment slack) + mbarrier
  // SPLIT: three more mbarriers (full / empty / second buffer free) and the scaled values of one frame, bin b at
  // float index b + o with o = (output row address / 4) mod 4, so that 16-byte chunks of the copy and of the
  // output row coincide (vector stores although rows of M + 1 floats are only 4-byte aligned)
  static constexpr size_t G_SYNC = SPLIT_ ? 48 : 0;      // full[2], empty[2], second buffer free
  static constexpr int VALS_FLOATS = M + 4;
  static constexpr size_t G_VALS = sizeof(float) * VALS_FLOATS * NSW_;
  static constexpr size_t G_BYTES = G_BUF + (NBUF - 1) * G_BUFB + G_HIST + G_MISC + G_LIST + G_STAGE + G_SYNC + G_VALS;
  static_assert(G_BYTES % 16 == 0, "group blocks stay 16-byte aligned");
  static constexpr size_t SMEM_TOTAL = OFF_GROUPS + G_BYTES * G_;
};

// misc words of a group
enum { MI_KMIN = 0, MI_KMAX = 16, MI_SUM = 32, MI_ARGMIN = 48, MI_ARGMAX = 49, MI_LCOUNT = 50,
       MI_GUESS = 51, MI_SELA = 52 /*b, below, cnt*/, MI_LDONE = 55, MI_SELB = 56, MI_RESA = 57, MI_RESB = 58,
       MI_FLAGS = 59, MI_WINL = 60 /* SPLIT: two windows published by the statistics warp: {klo, shift | use << 8} per frame parity */,
       MI_BFREE = 91 /* EARLY: warps that have read their last-pass inputs */,
       MI_DEF = 92 /* deferred ranking: {candidates, wanted rank, frame lo, frame hi} (16-byte aligned) */,
       MI_UMAX = 64, MI_BMIN = 80, MI_BMAX = 88 /* SPLIT: per warp, the lanes that hold the warp's minimum / maximum */ };
// misc words of a group in the split variant (the words above except MI_UMAX are unused there).  [s]: hand-over slot
// (statistics warp) s = frame parity when NSW = 2, 0 when NSW = 1; [p]: frame parity.
enum { MS_KMIN = 0 /* [s][8] */, MS_KMAX = 16 /* [s][8] */, MS_SUM = 32 /* [s][8] */, MS_BMIN = 48 /* [s][4] */, MS_BMAX = 56 /* [s][4] */,
       MS_WIN = 80 /* [p] {klo, shift | use << 8} */, MS_FLAGS = 84 /* [s] */, MS_LCOUNT = 86 /* [s] */ };

// Histogram slot of key k for the candidate window [klo, klo + width] with bucket shift sh:
// slot 0 = below the window, slots 1..NB = buckets, slot NB+1 = above.  The atomic increment is
// unconditional (ptxas turns a predicated ATOMS into a branch), the slot index is clamped instead.
template <int MODE, int NB>
__device__ __forceinline__ uint32_t hist_slot(uint32_t k, uint32_t klo, int sh) {
  if constexpr (MODE == SCALE_LOG10) {
    // ordered keys use all 32 bits: explicit compare for "below"
    const uint32_t d = min((k - klo) >> sh, (uint32_t)NB) + 1u;
    return k < klo ? 0u : d;
  } else {
    // keys are bit patterns of non-negative floats (< 2^31): the signed difference cannot overflow
    const int d1 = ((int)(k - klo) >> sh) + 1;      // arithmetic shift: negative when k < klo
    return (uint32_t)__vimin_s32_relu(d1, NB + 1);   // one instruction: max(min(d1, NB + 1), 0)
  }
}

// One step of an inclusive warp scan: v += (value of lane - off), for the lanes that have such a lane.  The shuffle's own
// "source lane in range" predicate guards the add: two instructions instead of shuffle + compare + select + add.
__device__ __forceinline__ uint32_t scan_step_up(uint32_t v, int off) {
  asm("{\n.reg .u32 t;\n.reg .pred p;\nshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n@p add.u32 %0, %0, t;\n}\n" : "+r"(v) : "r"(off));
  return v;
}

template <int T>
__device__ __forceinline__ void group_sync(int g) {
  if constexpr (T == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// keys: order-preserving u32 image of a value.  All modes except 20*log10 produce non-negative
// values whose raw bit pattern is already ordered.
template <int MODE>
__device__ __forceinline__ uint32_t v2_key(float v) {
  if constexpr (MODE == SCALE_LOG10) return ordered_key(v);
  else return __float_as_uint(v);
}
template <int MODE>
__device__ __forceinline__ float v2_unkey(uint32_t k) {
  if constexpr (MODE == SCALE_LOG10) return key_to_float(k);
  else return __uint_as_float(k);
}

// Exactly-rounded v/d for a launch-constant d (rd = RN(1/d)): q = RN(v*rd), r = v - q*d (exact, FMA),
// q' = RN(q + r*rd) (Markstein).  Used for divide_by_N_sqrt with odd log2 N (d = 2^k * sqrt(2)).
__device__ __forceinline__ float div_const(float v, float d, float rd) {
  const float q = v * rd;
  const float r = fmaf(-q, d, v);
  return fmaf(r, rd, q);
}

// v/d for a divisor shared by a whole frame (scale_to_zero_to_one, d = max of the frame): rd = RN(1/d), then two
// FMA correction steps.  Equals IEEE division bit for bit for 0 <= v <= d, d normal, significand of d not all
// ones (Markstein; checked against __fdiv_rn on 2e10 pairs by profiles/microbench/div_check.cu) -- the caller
// takes the IEEE path otherwise.  5 instructions instead of ~15.
__device__ __forceinline__ float div_by_frame_max(float v, float d, float rd) {
  float q = __fmul_rn(v, rd);
  float r = __fmaf_rn(-q, d, v);
  q = __fmaf_rn(r, rd, q);
  r = __fmaf_rn(-q, d, v);
  return __fmaf_rn(r, rd, q);
}
__device__ __forceinline__ bool frame_max_is_fast_divisor(float d) {
  const uint32_t b = __float_as_uint(d);
  return b >= 0x00800000u && b < 0x7f800000u && (b & 0x007fffffu) != 0x007fffffu;   // normal, finite, not all ones
}

// Bin index of val[i] of (logical) thread j -- the same dealing of the last-pass butterflies as in the body below
// (MIRROR for N = 4096, the generic mirrored list for N = 1024 / 2048 / 16384, shuffle-mirrored lanes for N = 8192),
// as a function of run-time (j, i): used by the statistics warp of the split variant to find a thread's bins.
template <class C>
__device__ __forceinline__ uint32_t v2_binof(int j, int i) {
  constexpr int M = C::M, T = C::T;
  constexpr int LR = C::FOUR ? C::R4 : C::R3, LNB = C::FOUR ? C::NB4 : C::NB3, LS = LNB * T;
  constexpr bool MIRROR = (C::NB3 == 2), MIRG = !MIRROR && LNB >= 2;
  uint32_t k;
  if constexpr (MIRROR) {
    const int jbB = j ? 256 - j : 128;
    const int q = i >> 2, r = i & 3;
    k = (uint32_t)(((r & 2) ? jbB : j) + 256 * q);
    return (r & 1) ? (uint32_t)M - k : k;
  } else if constexpr (MIRG) {
    const int pidx = i >> 1, sl = pidx / LR, rem = pidx % LR;
    const int a = j + sl * T, b = j ? LS - (j + sl * T) : (sl ? LS - sl * T : LS / 2);
    k = (uint32_t)(((rem & 1) ? b : a) + (rem >> 1) * LS);
    return (i & 1) ? (uint32_t)M - k : k;
  } else {
    const int beta = (j & 31) < 16 ? 16 * (j >> 5) + (j & 31) : (j == 16 ? 128 : 256 - (16 * (j >> 5) + (j & 31) - 16));
    k = (uint32_t)(beta + (i >> 1) * 256);
    return (i & 1) ? (uint32_t)M - k : k;
  }
}

// FASTIN: the launch's input is f32 with every frame on a 16-byte boundary (the common case) -- the two run-time
// format tests in front of every frame's loads become compile-time constants (cfg2 +1.5 %).
// SMASK >= 0: the statistics mask is a compile-time constant too (7: minimum / maximum + average + median, the mask the
// reference's `FrequencySpectrum::new` implies) -- the per-frame tests of `stats` / `stats_mask` disappear.
template <class C, int MODE, bool FULL, bool FASTIN = false, int SMASK = -1>
__device__ __forceinline__ void spectrum_body_v2(const V2Params &P, unsigned char *smem, const uint32_t tm_base) {
  constexpr int N = C::N, M = C::M, E = C::E, T = C::T, G = C::G, NW = C::NW, R3 = C::R3, NB3 = C::NB3;
  constexpr int PADT = T + T / C::PADE;  // padded stride of the strided reads (first exchange only)
  // N = 4096: thread j runs the two last-pass butterflies j and 256-j (thread 0: 0 and 128), whose
  // outputs are exactly each other's mirror bins Z[k] <-> Z[M-k]: the real-FFT recombination then
  // needs no exchange and no barrier.
  constexpr bool MIRROR = (NB3 == 2);
  // The same idea for the other sizes whose last pass runs >= 2 butterflies per thread (N = 1024, 2048 and the
  // radix-2 fourth pass of N = 16384): with S = (butterflies of the last pass) = LNB*T, thread j owns the LNB/2
  // butterflies A_i = j + i*T and their mirrors B_i = S - A_i (thread 0: B_0 = S/2, the self-mirrored one, and
  // A_0 = 0 pairs with itself), so every pair Z[k], Z[M-k] meets in one thread and the recombination needs no
  // exchange.  N = 8192 runs ONE radix-16 butterfly per thread: there the butterflies are dealt so that lane l < 16
  // of (logical) warp w owns butterfly 16w + l and lane l + 16 its mirror 256 - (16w + l) (warp 0: lane 16 owns the
  // self-mirrored 128), and the partner's outputs arrive by 16 shuffles (xor 16) instead of a shared-memory
  // exchange and a barrier.
  constexpr int LR = C::FOUR ? C::R4 : R3, LNB = C::FOUR ? C::NB4 : NB3, LS = LNB * T, LH = LNB / 2;
  constexpr bool MIRG = !MIRROR && LNB >= 2;

  const SpectrumParams &p = P.b;
  const int tid = threadIdx.x;
  const int g = tid / T;
  const int rot = (g / C::GPQ) % C::NROT;
  const int j = (tid - g * T + 32 * rot) & (T - 1);   // logical thread index inside the group
  const int lane = tid & 31, wig = j >> 5;            // logical warp index inside the group
  unsigned char *gbase = smem + C::OFF_GROUPS + (size_t)g * C::G_BYTES;
  float2 *bufA = reinterpret_cast<float2 *>(gbase);
  float2 *bufB = reinterpret_cast<float2 *>(gbase + (C::NBUF - 1) * C::G_BUF);   // == bufA when NBUF == 1
  constexpr size_t OFF_HIST = C::G_BUF + (C::NBUF - 1) * C::G_BUFB;
  uint32_t *hist0 = reinterpret_cast<uint32_t *>(gbase + OFF_HIST);
  uint32_t *misc = reinterpret_cast<uint32_t *>(gbase + OFF_HIST + C::G_HIST);
  uint32_t *list = reinterpret_cast<uint32_t *>(gbase + OFF_HIST + C::G_HIST + C::G_MISC);
  uint32_t *cand = list + 32;      // DEFRANK: candidate of thread j of the group (16-byte aligned)
  unsigned char *tail = gbase + OFF_HIST + C::G_HIST + C::G_MISC + C::G_LIST;
  constexpr bool ALIAS = C::ALIAS;
  unsigned char *stage = ALIAS ? reinterpret_cast<unsigned char *>(bufB) : tail;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(ALIAS ? tail : tail + sizeof(float) * N + 16);
  uint32_t phase = 0;              // parity of the mbarrier phase the next wait completes
  // SPLIT: hand-over to the group's statistics warp
  unsigned char *sync_base = gbase + OFF_HIST + C::G_HIST + C::G_MISC + C::G_LIST + C::G_STAGE;
  uint64_t *mb_full = reinterpret_cast<uint64_t *>(sync_base);        // [slot] NW arrivals: values + partials of a frame are in place
  uint64_t *mb_empty = reinterpret_cast<uint64_t *>(sync_base + 16);  // [slot] 1 arrival: the statistics warp is done with them
  uint32_t *bfree_cnt = reinterpret_cast<uint32_t *>(sync_base + 32);  // warps that have read their last-pass inputs: the last one stages the next frame
  uint32_t *bfree = C::SPLIT ? bfree_cnt : &misc[MI_BFREE];
  float *vbuf0 = reinterpret_cast<float *>(sync_base + C::G_SYNC);    // [slot] copies of the scaled values
  uint32_t ephase = 0;               // bit s = parity of the next wait on empty[s]
  uint32_t fcount = 0;               // frames handed over so far
  constexpr int HBITS = C::HBITS, HBPL = C::HBPL;
#ifndef SA_DEFRANK
#define SA_DEFRANK 1
#endif
  // DEFRANK: the candidates of a frame's selected bucket are ranked one group barrier later (right after the first
  // exchange barrier of the group's next frame, when the list is known to be complete) by ONE fixed warp, instead of
  // by the warp of the thread that completed the list: that warp was by construction the group's latest, and its
  // chain fence -> atomic -> vote -> fence -> list -> shuffles -> store held the whole group at its next barrier
  // (profiles/README.md, round 2: 0.63 -> 0.70 of the roofline with that chain removed).
  constexpr bool DEFRANK = SA_DEFRANK && !C::SPLIT;
  constexpr int RANKW = NW > 1 ? 1 : 0;      // logical warp of the group that ranks
#ifndef SA_EARLYSTAGE
#define SA_EARLYSTAGE 1
#endif
  // EARLY (launches without statistics only): the bulk copy of the group's next frame into the second exchange buffer is
  // issued by the LAST warp that has used its last-pass inputs from it (a counter in shared memory) instead of after an
  // extra group barrier: 0.83 -> 0.89 of the roofline at cfg2 without statistics.  With statistics the copy hangs on
  // the statistics barrier; issuing it earlier measured SLOWER there (0.660 -> 0.649).
  constexpr bool EARLY = SA_EARLYSTAGE && C::ALIAS && !C::SPLIT;
#ifndef SA_GW_MIN_LOG2
#define SA_GW_MIN_LOG2 17
#endif
#ifndef SA_GW_SHRINK
#define SA_GW_SHRINK 6
#endif
#ifndef SA_GW_GROW
#define SA_GW_GROW 2
#endif
  constexpr uint32_t GW_INIT = 1u << 20, GW_MIN = 1u << SA_GW_MIN_LOG2;   // median window half-width (key units)
  uint32_t gw = 0u;                // 0: no guess yet (first frame of the group)
  int par = 0;                     // which of the two histograms this frame uses
  // record fields written one barrier later (after every thread finished the frame) by the first lane
  // of the group's last warp; thread 0 already has bin M/2, the Nyquist bin and the bulk copies to do
  const bool rec_thread = (j == T - 32);
  bool pend = false, pend_med = false;
  u64 pend_cur = 0;
  uint32_t pend_kmn = 0, pend_kmx = 0, pend_medkey = 0;
  float pend_avg = 0.f;   // the division happens in the statistics phase, off the exchange critical path

  // ---- one-time: tables -> tensor memory (or shared / global memory for N = 16384), histogram cleared
  const bool has_win = p.window != nullptr;
#ifndef SA_OPT
#define SA_OPT 3   // bit 0: packed window multiply (FMUL2: -16 instructions per thread).  ptxas contracts a packed multiply with
                   // the packed add that consumes it into FFMA2, so half of the first-pass inputs skip the rounding of
                   // the windowed sample (src/windows.rs:47 multiplies, then microfft adds): the fused window is then more
                   // accurate than, and no longer bit-identical to, a caller-side window followed by window = none; both stay
                   // far inside the 1e-4 x frame-maximum tolerance.  Measured +1.0 % at cfg2.
                   // bit 1: a power-of-two scaling factor rides on the window multipliers instead of costing a multiply
                   // per bin (exact: every later operation commutes with a power of two).  Measured +0.9 % at cfg2.
#endif
  constexpr bool FOLD_SCALE = (MODE == SCALE_MUL) && (SA_OPT & 2);
  const float wscale = FOLD_SCALE ? P.cmul : 1.0f;
  const int jbB = j ? 256 - j : 128;   // MIRROR: second last-pass butterfly of this thread
  // N = 8192 (neither MIRROR nor MIRG): butterfly of this thread in the last pass
  const int beta = (j & 31) < 16 ? 16 * (j >> 5) + (j & 31)
                                 : (j == 16 ? 128 : 256 - (16 * (j >> 5) + (j & 31) - 16));
  auto bfA = [&](int i) -> int { return j + i * T; };                                         // MIRG
  auto bfB = [&](int i) -> int { return j ? LS - (j + i * T) : (i ? LS - i * T : LS / 2); };  // MIRG
  const uint32_t ta = tmem_addr(tm_base, tid >> 5, (rot * C::TM_JSETS + (j >> 7)) * C::TM_SET_COLS);   // this thread's table row
  {
    // the first NROT * max(T,128) threads cover every (lane quadrant, column set) once
    if (tid < C::NROT * (T > 128 ? T : 128)) {
      float v[16];
#pragma unroll
      for (int h = 0; h < 2; h++) {   // window multipliers of samples 2n, 2n+1 for n = j + s*T
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const float2 w = has_win ? __ldg(reinterpret_cast<const float2 *>(p.window) + j + (8 * h + s) * T) : make_float2(1.f, 1.f);
          // SCALE_MUL: the scaling factor (a power of two: 1/2 of the real-FFT recombination times 1/N or 1/sqrt(N))
          // rides on the window multipliers -- everything downstream is linear and a power of two commutes with
          // every rounding, so the values are the same bits as scaling at the end
          v[2 * s] = w.x * wscale; v[2 * s + 1] = w.y * wscale;
        }
        tmem_st16(ta + C::TM_WIN + 16 * h, v);
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {   // pass 2: w_256^(k*pp), k = 1..15
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int k = 8 * h + s + 1;
          const float2 w = k < 16 ? __ldg(P.tw2 + (k - 1) * 16 + (j & 15)) : make_float2(0.f, 0.f);
          v[2 * s] = w.x; v[2 * s + 1] = w.y;
        }
        tmem_st16(ta + C::TM_TW2 + 16 * h, v);
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {   // last pass: butterfly b, k = 1..R3-1 -> entry b*(R3-1) + k-1
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int e = 8 * h + s;
          float2 w = make_float2(0.f, 0.f);
          if (e < NB3 * (R3 - 1)) {
            const int b = e / (R3 - 1), k = e % (R3 - 1) + 1;
            const int jb = MIRROR ? (b ? jbB : j) : C::FOUR ? (j & 255) : MIRG ? ((b & 1) ? bfB(b >> 1) : bfA(b >> 1)) : beta;
            w = __ldg(P.tw3 + (k - 1) * 256 + jb);
          }
          v[2 * s] = w.x; v[2 * s + 1] = w.y;
        }
        tmem_st16(ta + C::TM_TW3 + 16 * h, v);
      }
#pragma unroll
      for (int s = 0; s < 8; s++) {   // recombination twiddles W_N^k of the 8 bin pairs this thread owns
        int k = MIRROR ? ((s & 1) ? jbB : j) + 256 * (s >> 1) : beta + s * 256;
        if constexpr (MIRG) { const int i = s / LR, rem = s % LR; k = ((rem & 1) ? bfB(i) : bfA(i)) + (rem >> 1) * LS; }
        const float2 w = __ldg(P.twr + k);
        v[2 * s] = w.x; v[2 * s + 1] = w.y;
      }
      tmem_st16(ta + C::TM_TWR, v);
      if constexpr (C::FOUR) {   // fourth pass (radix 2): w_M^(jb) for the 8 butterflies in the order A_0, B_0, A_1, ...
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int jb = (s & 1) ? bfB(s >> 1) : bfA(s >> 1);
          const float2 w = __ldg(P.tw4 + jb);
          v[2 * s] = w.x; v[2 * s + 1] = w.y;
        }
        tmem_st16(ta + C::TM_TW4, v);
      }
    }
    tmem_fence_before_sync();
  }
  // window multiplier from global memory (rare paths only)
  auto win1 = [&](int n) -> float { return has_win ? __ldg(p.window + n) : 1.0f; };
  for (int i = j; i < (int)(C::G_HIST / sizeof(uint32_t)); i += T) hist0[i] = 0;
  if (j == 0) { misc[MI_BFREE] = 0u; misc[MI_DEF] = 0u; misc[MI_LCOUNT] = 0u; misc[MI_LDONE] = 0u; misc[MI_ARGMIN] = 0xffffffffu; misc[MI_ARGMAX] = 0u; misc[MI_FLAGS] = 0u; }
  if (j == 0) {
    mbar_init(mbar, 1);
    if constexpr (C::SPLIT) {
      for (int q = 0; q < C::NSW; q++) { mbar_init(mb_full + q, NW); mbar_init(mb_empty + q, 1); }
      *bfree_cnt = 0u;
    }
    mbar_fence_init();
  }
  __syncthreads();
  tmem_fence_after_sync();

  const uint32_t lo = p.bin_lo, hi = p.bin_hi;
  const uint32_t count = hi - lo + 1;
  const uint32_t smask = SMASK >= 0 ? (uint32_t)SMASK : p.stats_mask;
  const bool want_stats = SMASK >= 0 ? (SMASK != 0) : (p.stats != nullptr);
  const bool want_median = want_stats && (smask & SA_STATS_MEDIAN);
  const float cscale = P.cmul;

  u64 frame = (u64)blockIdx.x * G + g;
  const u64 fstep = (u64)gridDim.x * G;
  if (frame >= p.n_frames) return;  // group-uniform; no CTA-wide barrier after this point
  // SPLIT: position of bin 0 of a frame inside its copy, (output row address / 4) mod 4 (see V2Cfg::G_VALS)
  uint32_t o_pos = (uint32_t)(((reinterpret_cast<uintptr_t>(p.out_vals) >> 2) + frame * p.out_stride - p.bin_lo) & 3u);
  const uint32_t o_step = (uint32_t)((fstep * p.out_stride) & 3u);
  // Tuning hook (SA_STAGGER_NS, default 0): start the groups a fraction of a frame time apart.  Measured: no
  // effect -- the groups drift apart by themselves.
  if (G > 1 && g > 0 && P.stagger_ns) __nanosleep((unsigned)g * P.stagger_ns);

  // ---- load of the first frame (raw samples; the window is applied at the top of the loop) ------
  float2 x[E];
  // Byte address of frame f.  Bulk copies need 16-byte aligned addresses and sizes: a frame that does not start on
  // a 16-byte boundary is fetched as its aligned superset [addr & ~15, round_up_16(addr + frame bytes)) and sits
  // `addr & 15` bytes into the staging buffer.  Where that superset would leave the caller's buffer
  // [samples, samples + extent) -- the first bytes of the first frame, the last bytes of the last one -- the copy is
  // clamped to the aligned interior and the (at most 15) bytes on that side are copied by ordinary loads: nothing
  // outside the buffer is ever read.
  const bool in_i16 = FASTIN ? false : (p.in_i16 != 0);
  const uint32_t esz = in_i16 ? 2u : 4u;
  const uintptr_t base_addr = reinterpret_cast<uintptr_t>(p.samples);
  const uintptr_t end_addr = base_addr + ((p.n_frames - 1) * p.frame_stride + (u64)N) * esz;
  auto frame_addr = [&](u64 f) -> uintptr_t { return base_addr + f * p.frame_stride * esz; };
  // thread 0 of the group: bulk copy of frame f into the staging buffer
  const bool al16 = FASTIN ? true : (p.aligned16 != 0);   // launch-uniform: every frame starts on a 16-byte boundary
  auto stage_frame = [&](u64 f) {
    // the target was last read / written through the generic proxy (exchange data, sample loads): order those
    // accesses (made visible to this thread by the preceding barrier) before the async-proxy write of the copy
    fence_proxy_async();
    const uintptr_t a = frame_addr(f);
    if (al16) {
      mbar_arrive_expect_tx(mbar, (uint32_t)N * esz);
      bulk_copy_g2s(stage, reinterpret_cast<const void *>(a), (uint32_t)N * esz, mbar);
    } else {
      const uintptr_t fe = a + (uintptr_t)N * esz;                 // end of the frame
      uintptr_t a0 = a & ~(uintptr_t)15, a1 = (fe + 15u) & ~(uintptr_t)15;
      const uintptr_t s0 = a0;                                     // address that maps to stage[0]
      if (a0 < base_addr) {                                        // head of the first frame: ordinary loads
        a0 += 16;
        for (uintptr_t q = a; q < a0; q += esz) {
          if (esz == 4u) *reinterpret_cast<float *>(stage + (q - s0)) = *reinterpret_cast<const float *>(q);
          else *reinterpret_cast<short *>(stage + (q - s0)) = *reinterpret_cast<const short *>(q);
        }
      }
      if (a1 > end_addr) {                                         // tail of the last frame
        a1 -= 16;
        for (uintptr_t q = a1 > a ? a1 : a; q < fe; q += esz) {
          if (esz == 4u) *reinterpret_cast<float *>(stage + (q - s0)) = *reinterpret_cast<const float *>(q);
          else *reinterpret_cast<short *>(stage + (q - s0)) = *reinterpret_cast<const short *>(q);
        }
      }
      const uint32_t bytes = (uint32_t)(a1 - a0);
      mbar_arrive_expect_tx(mbar, bytes);    // (release: the ordinary stores above are visible to the waiters)
      bulk_copy_g2s(stage + (a0 - s0), reinterpret_cast<const void *>(a0), bytes, mbar);
    }
  };
  if (j == 0) stage_frame(frame);
  float2 *wbuf = bufA, *obuf = bufB;  // exchange buffers: wbuf is written next

  // thread 0: write the record fields of the previous frame (everything except a median written
  // by the "last pusher"), then reset the per-frame words.  Called right after a group barrier.
  auto flush_record = [&]() {
    if (pend && !(SA_EXP & 1)) {
      const bool mm = smask & SA_STATS_MINMAX;
      uint4 r0;
      r0.x = mm ? __float_as_uint(v2_unkey<MODE>(pend_kmn)) : 0u;
      r0.y = mm ? misc[MI_ARGMIN] : 0u;
      r0.z = mm ? __float_as_uint(v2_unkey<MODE>(pend_kmx)) : 0u;
      r0.w = mm ? misc[MI_ARGMAX] : 0u;
      uint32_t *rec = reinterpret_cast<uint32_t *>(&p.stats[pend_cur]);
      *reinterpret_cast<uint4 *>(rec) = r0;
      rec[4] = (smask & SA_STATS_AVERAGE) ? __float_as_uint(pend_avg) : 0u;
      if (pend_med) rec[5] = want_median ? __float_as_uint(v2_unkey<MODE>(pend_medkey)) : 0u;
      const uint32_t fl = misc[MI_FLAGS];
      const uint32_t st = (fl & 1u) ? SA_ERR_NAN_VALUES : (fl & 2u) ? SA_ERR_INFINITY_VALUES
                        : (fl & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
      *reinterpret_cast<uint2 *>(rec + 6) = make_uint2(st, 0u);
      pend = false;
    }
    misc[MI_ARGMIN] = 0xffffffffu; misc[MI_ARGMAX] = 0u; misc[MI_FLAGS] = 0u;
  };

  // bin index of val[i]
  auto binof = [&](int i) -> uint32_t {
    if constexpr (MIRROR) {
      const int q = i >> 2, r = i & 3;
      const uint32_t k = (uint32_t)(((r & 2) ? jbB : j) + 256 * q);
      return (r & 1) ? (uint32_t)M - k : k;
    } else if constexpr (MIRG) {
      const int pidx = i >> 1, sl = pidx / LR, rem = pidx % LR;
      const uint32_t k = (uint32_t)(((rem & 1) ? bfB(sl) : bfA(sl)) + (rem >> 1) * LS);
      return (i & 1) ? (uint32_t)M - k : k;
    } else {
      const uint32_t k = (uint32_t)(beta + (i >> 1) * 256);
      return (i & 1) ? (uint32_t)M - k : k;
    }
  };

  // ---- FrequencyLimit slice (src/lib.rs:286-304): validity of the 16 (+1) owned bins -- the same for every frame, so
  // it is computed once (limited ranges +1 %).  Skipping values that NO lane of a warp keeps by warp-uniform branches
  // was measured too: the branches cost more than the predicated work they skip (cfg4 0.415 -> 0.358, Max(4000) -7 %).
  uint32_t valid = 0xffffffffu;
  const bool valid_mid = (j == 0) && (FULL || ((uint32_t)(M / 2) >= lo && (uint32_t)(M / 2) <= hi));
  if constexpr (!FULL) {
    valid = 0;
#pragma unroll
    for (int i = 0; i < E; i++) {
      const uint32_t b = binof(i);
      valid |= (b >= lo && b <= hi) ? (1u << i) : 0u;
    }
  }

  // DEFRANK: median of the group's previous frame from the candidate table (complete: a group barrier has passed since
  // the last store to it) -- the `want`-th smallest of the `cnt` keys in it (src/spectrum.rs:515-524 sorts; equal keys
  // are equal values, so ties need no order).  One warp; each lane holds T/32 table words plus one overflow key.
  // Up to three keys are settled by their minimum, maximum and sum (three independent warp reductions); more by
  // stripping minima.
  auto deferred_select = [&]() {
    if constexpr (DEFRANK) {
      if (wig == RANKW) {
        const uint4 d = *reinterpret_cast<const uint4 *>(&misc[MI_DEF]);
        if (d.x) {
          constexpr int KPL = T / 32;                 // table words per lane
          constexpr uint32_t NONE = 0xffffffffu;
          uint32_t k[KPL + 1];
          if constexpr (KPL % 4 == 0) {
#pragma unroll
            for (int q = 0; q < KPL / 4; q++) {
              const uint4 v = reinterpret_cast<const uint4 *>(cand)[lane + 32 * q];
              k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < KPL; q++) k[q] = cand[lane + 32 * q];
          }
          const uint32_t novf = misc[MI_LCOUNT], kovf = list[lane];   // (independent loads)
          k[KPL] = ((uint32_t)lane < novf) ? kovf : NONE;
          uint32_t cnt = d.x, want = d.y, res;
          for (;;) {
            uint32_t mn = NONE, mx = 0u, sm = 0u;
#pragma unroll
            for (int q = 0; q <= KPL; q++) {
              const bool v = k[q] != NONE;
              mn = min(mn, k[q]);
              mx = max(mx, v ? k[q] : 0u);
              sm += v ? k[q] : 0u;
            }
            mn = __reduce_min_sync(0xffffffffu, mn);
            if (cnt <= 3u) {
              mx = __reduce_max_sync(0xffffffffu, mx);
              sm = __reduce_add_sync(0xffffffffu, sm);
              res = want == 0u ? mn : (want == cnt - 1u ? mx : sm - mn - mx);
              break;
            }
            uint32_t c = 0u;
#pragma unroll
            for (int q = 0; q <= KPL; q++) { const bool e = (k[q] == mn); c += e ? 1u : 0u; k[q] = e ? NONE : k[q]; }
            c = __reduce_add_sync(0xffffffffu, c);
            res = mn;
            if (want < c || c == 0u) break;      // (c == 0 cannot happen: the table holds cnt keys)
            want -= c; cnt -= c;
          }
          if (lane == 0) {
            const u64 cf = (u64)d.z | ((u64)d.w << 32);
            reinterpret_cast<uint32_t *>(&p.stats[cf])[5] = __float_as_uint(v2_unkey<MODE>(res));
            misc[MI_GUESS] = res;
            misc[MI_LCOUNT] = 0u;
            misc[MI_DEF] = 0u;
          }
        }
      }
    }
  };

  // EARLY: called by every warp once the values it loaded from the second exchange buffer have been USED (so the loads
  // have completed); the last warp of the group to get here issues the bulk copy of the next frame into that buffer.
  auto early_stage = [&](bool more_, u64 next) {
    if constexpr (EARLY) {
      if (want_stats) return;
      __syncwarp();
      if (lane == 0) {
        if (atomicAdd(bfree, 1u) == (uint32_t)(NW - 1)) {
          *bfree = 0u;           // next used after two group barriers
          if (more_) stage_frame(next);
        }
      }
    }
  };

#pragma unroll 1
  for (;;) {
    const bool more = frame + fstep < p.n_frames;   // group-uniform
    {
#if defined(SA_FRAME_WAIT_HINT)
      mbar_wait_sleep(mbar, phase);
#else
      mbar_wait(mbar, phase);
#endif
      phase ^= 1u;
      if (al16) {
        if (in_i16) {
          const short2 *src = reinterpret_cast<const short2 *>(stage);
#pragma unroll
          for (int s = 0; s < E; s++) { const short2 v = src[j + s * T]; x[s] = make_float2((float)v.x, (float)v.y); }
        } else {
          const float2 *src = reinterpret_cast<const float2 *>(stage);
#pragma unroll
          for (int s = 0; s < E; s++) x[s] = src[j + s * T];
        }
      } else {
        // the frame sits `off` bytes into the buffer; pairs that straddle their natural alignment are read
        // as two scalars (off is group-uniform, so no divergence)
        const uint32_t off = (uint32_t)(frame_addr(frame) & 15u);
        if (in_i16) {
          const short *src = reinterpret_cast<const short *>(stage + off);
#pragma unroll
          for (int s = 0; s < E; s++) x[s] = make_float2((float)src[2 * (j + s * T)], (float)src[2 * (j + s * T) + 1]);
        } else {
          const float *src = reinterpret_cast<const float *>(stage + off);
#pragma unroll
          for (int s = 0; s < E; s++) x[s] = make_float2(src[2 * (j + s * T)], src[2 * (j + s * T) + 1]);
        }
      }
    }

    // ---- window (src/windows.rs: w_i * x_i) -------------------------------------------------------
    // packed by default (SA_OPT bit 0): ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, so half of the
    // windowed samples are not rounded on their own (src/windows.rs:47 rounds each) -- see the SA_OPT note above;
    // -DSA_OPT=2 builds the scalar mul.rn that rounds like the reference
    {
      float w[32];
      tmem_ld32(ta + C::TM_WIN, w);
#pragma unroll
      for (int s = 0; s < E; s++) {
        if constexpr (SA_OPT & 1) {
          x[s] = mul2(make_float2(w[2 * s], w[2 * s + 1]), x[s]);
        } else {
          x[s].x = __fmul_rn(w[2 * s], x[s].x);
          x[s].y = __fmul_rn(w[2 * s + 1], x[s].y);
        }
      }
    }

    // ---- pass 1: radix 16, NS = 1 (no twiddles) -------------------------------------------------
    dft16(x);
    {
      if constexpr (C::PAD2) {
        float4 *w = reinterpret_cast<float4 *>(wbuf + 18 * j);  // pad(16 j + k) = 18 j + k: 16-byte aligned rows
#pragma unroll
        for (int k = 0; k < 16; k += 2) w[k / 2] = make_float4(x[k].x, x[k].y, x[k + 1].x, x[k + 1].y);
      } else {
        float2 *w = wbuf + 17 * j;  // pad(16 j + k) = 17 j + k
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = x[k];
      }
    }
    group_sync<T>(g);
    // every thread of the group has consumed the staged samples and finished the previous frame
    if constexpr (!ALIAS) { if (j == 0 && more) stage_frame(frame + fstep); }
    if constexpr (!C::SPLIT) { if (want_stats && rec_thread) flush_record(); }
    if (want_median) deferred_select();
    {
      const float2 *r = wbuf + (j + (C::PAD2 ? 2 : 1) * (j >> 4));
#pragma unroll
      for (int s = 0; s < 16; s++) x[s] = r[s * PADT];
    }
    { float2 *t = wbuf; wbuf = obuf; obuf = t; if constexpr (C::NBUF == 1) group_sync<T>(g); }

    // ---- pass 2: radix 16, NS = 16 ---------------------------------------------------------------
    {
      float t[32];
      tmem_ld32(ta + C::TM_TW2, t);
#pragma unroll
      for (int k = 1; k < 16; k++) x[k] = cmul(x[k], make_float2(t[2 * (k - 1)], t[2 * (k - 1) + 1]));
    }
    if (!(SA_EXP & 0x80000)) dft16(x);   // (timing experiment 0x80000: no second-pass butterflies)
    if (!(SA_EXP & 0x40000)) {          // (timing experiment 0x40000: no second exchange, registers passed through)
      // second exchange: unit-stride runs of 16 on both sides -> no padding needed
      float2 *w = wbuf + (j >> 4) * 256 + (j & 15);
#pragma unroll
      for (int k = 0; k < 16; k++) w[16 * k] = x[k];
    }
    group_sync<T>(g);
    float val[E];        // magnitudes of the 16 bins this thread owns (bin index: binof(i))
    float vmid = 0.f;    // bin M/2, thread 0 only
    if constexpr (MIRROR) {
      // ---- pass 3 (last): two radix-8 butterflies, jbA = j -> x[0..7], jbB -> x[8..15] ------------
      if (!(SA_EXP & 0x40000)) {
#pragma unroll
      for (int k = 0; k < 8; k++) { x[k] = wbuf[j + 256 * k]; x[8 + k] = wbuf[jbB + 256 * k]; }
      }
      { float2 *t = wbuf; wbuf = obuf; obuf = t; if constexpr (C::NBUF == 1) group_sync<T>(g); }
      if constexpr (C::SPLIT && ALIAS) {
        // every warp has read its last-pass inputs -> the next frame may land in buffer B: the LAST warp to get
        // here issues the bulk copy (as early as possible: the data is needed at the top of the next iteration)
        __syncwarp();
        if (lane == 0) {
          if (atoms_add_acq_rel(bfree, 1u) == (uint32_t)(NW - 1)) {
            *bfree = 0u;           // next used after two group barriers
            if (more) stage_frame(frame + fstep);
          }
        }
      }
      {
        float2 va[8], vb[8];
        va[0] = x[0]; vb[0] = x[8];
        {
          float t[32];
          tmem_ld32(ta + C::TM_TW3, t);   // entries 0..6: butterfly j, 7..13: butterfly jbB
#pragma unroll
          for (int k = 1; k < 8; k++) {
            va[k] = cmul(x[k], make_float2(t[2 * (k - 1)], t[2 * (k - 1) + 1]));
            vb[k] = cmul(x[8 + k], make_float2(t[2 * (k + 6)], t[2 * (k + 6) + 1]));
          }
        }
        float tr[16];
        tmem_ld16(ta + C::TM_TWR, tr);    // W_N^k of the pairs (j + 256 q), (jbB + 256 q)
        dft8(va);     // va[q] = Z[j   + 256 q]
        dft8(vb);     // vb[q] = Z[jbB + 256 q]
        early_stage(more, frame + fstep);
        // ---- recombination (src/fft.rs:126-132 bin convention) + magnitude (src/lib.rs:366-372) ----
        // pair (k, M-k): 2X[k] = P + T, 2conj(X[M-k]) = P - T with P = A + conj(B), Q = A - conj(B), T = -i w_k Q
        auto pair_mag = [&](float2 a, float2 bq, float2 w, float &m0, float &m1) {
          const float2 cb = make_float2(bq.x, -bq.y);
          const float2 pp = add2(a, cb), qq = sub2(a, cb);
          const float2 tt = fma2(bcast2(w.x), make_float2(qq.y, -qq.x), mul2(bcast2(w.y), qq));
          const float2 uu = add2(pp, tt), vv = sub2(pp, tt);
          m0 = sqrt_approx(fmaf(uu.x, uu.x, uu.y * uu.y));
          m1 = sqrt_approx(fmaf(vv.x, vv.x, vv.y * vv.y));
        };
        const bool t0 = (j == 0);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // kA = j + 256 q pairs with M - kA: held in vb[7-q] (thread 0: its own va[8-q])
          const float2 p1 = t0 ? va[(8 - q) & 7] : vb[7 - q];
          pair_mag(va[q], p1, make_float2(tr[4 * q], tr[4 * q + 1]), val[4 * q], val[4 * q + 1]);
          // kB = jbB + 256 q pairs with M - kB: held in va[7-q] (thread 0: its own vb[7-q])
          const float2 p2 = t0 ? vb[7 - q] : va[7 - q];
          pair_mag(vb[q], p2, make_float2(tr[4 * q + 2], tr[4 * q + 3]), val[4 * q + 2], val[4 * q + 3]);
        }
        if (t0) {
          val[0] = 2.0f * fabsf(va[0].x + va[0].y);                         // X[0]   = Zr + Zi
          val[1] = 2.0f * fabsf(va[0].x - va[0].y);                         // X[N/2] = Zr - Zi (packed Nyquist)
          vmid = 2.0f * sqrt_approx(fmaf(va[4].x, va[4].x, va[4].y * va[4].y));  // X[N/4] = conj(Z[M/2])
        }
      }
    } else {
      // inputs of the last pass in the order A_0, B_0, A_1, B_1, ...: butterfly b reads positions b + k*LS
      auto read_mirror_list = [&]() {
#pragma unroll
        for (int i = 0; i < LH; i++) {
          const float2 *ra = wbuf + bfA(i), *rb = wbuf + bfB(i);
#pragma unroll
          for (int k = 0; k < LR; k++) { x[(2 * i) * LR + k] = ra[k * LS]; x[(2 * i + 1) * LR + k] = rb[k * LS]; }
        }
      };
      if constexpr (MIRG && !C::FOUR) {
        read_mirror_list();
      } else {
        const float2 *r = wbuf + (C::FOUR ? j : beta);   // N = 8192: inputs of butterfly beta (T = 256 = stride)
#pragma unroll
        for (int s = 0; s < 16; s++) x[s] = r[s * T];
      }
      { float2 *t = wbuf; wbuf = obuf; obuf = t; if constexpr (C::NBUF == 1) group_sync<T>(g); }
      if constexpr (C::SPLIT && ALIAS) {
        // every warp has read its last-pass inputs -> the next frame may land in buffer B: the LAST warp to get
        // here issues the bulk copy (as early as possible: the data is needed at the top of the next iteration)
        __syncwarp();
        if (lane == 0) {
          if (atoms_add_acq_rel(bfree, 1u) == (uint32_t)(NW - 1)) {
            *bfree = 0u;           // next used after two group barriers
            if (more) stage_frame(frame + fstep);
          }
        }
      }

      if constexpr (C::FOUR) {
        // ---- pass 3: radix 16, NS = 256 (one butterfly per thread) ---------------------------------
        {
          float t[32];
          tmem_ld32(ta + C::TM_TW3, t);     // w_4096^(k (j mod 256)), k = 1..15
#pragma unroll
          for (int k = 1; k < 16; k++) x[k] = cmul(x[k], make_float2(t[2 * (k - 1)], t[2 * (k - 1) + 1]));
        }
        dft16(x);
        {
          float2 *w = wbuf + (j >> 8) * 4096 + (j & 255);
#pragma unroll
          for (int k = 0; k < 16; k++) w[256 * k] = x[k];
        }
        group_sync<T>(g);
        read_mirror_list();
        { float2 *t = wbuf; wbuf = obuf; obuf = t; if constexpr (C::NBUF == 1) group_sync<T>(g); }
      }

      if constexpr (MIRG) {
        // ---- last pass: LNB butterflies of radix LR, x[b*LR + q] = Z[butterfly_b + q*LS] afterwards --------
        {
          float t[32];
          if constexpr (C::FOUR) {
            float t4[16];
            tmem_ld16(ta + C::TM_TW4, t4);
#pragma unroll
            for (int e = 0; e < 16; e++) t[e] = t4[e];
          } else {
            tmem_ld32(ta + C::TM_TW3, t);   // entry b*(LR-1) + k-1
          }
#pragma unroll
          for (int b = 0; b < LNB; b++) {
            float2 v[LR];
#pragma unroll
            for (int k = 0; k < LR; k++) v[k] = x[b * LR + k];
#pragma unroll
            for (int k = 1; k < LR; k++) {
              const int e = b * (LR - 1) + k - 1;
              v[k] = cmul(v[k], make_float2(t[2 * e], t[2 * e + 1]));
            }
            dft<LR>(v);
#pragma unroll
            for (int k = 0; k < LR; k++) x[b * LR + k] = v[k];
          }
        }
        // ---- recombination in registers (src/fft.rs:126-132 bin convention) + magnitude (src/lib.rs:366-372) ----
        float tr[16];
        tmem_ld16(ta + C::TM_TWR, tr);     // W_N^k in the order of the pairs below
        auto pair_mag = [&](float2 a, float2 bq, float2 w, float &m0, float &m1) {
          // P = A + conj(B), Q = A - conj(B), T = -i w Q (w = (cos,-sin)), 2X[k] = P + T, 2conj(X[M-k]) = P - T
          const float2 cb = make_float2(bq.x, -bq.y);
          const float2 pp = add2(a, cb), qq = sub2(a, cb);
          const float2 tt = fma2(bcast2(w.x), make_float2(qq.y, -qq.x), mul2(bcast2(w.y), qq));
          const float2 uu = add2(pp, tt), vv = sub2(pp, tt);
          m0 = sqrt_approx(fmaf(uu.x, uu.x, uu.y * uu.y));
          m1 = sqrt_approx(fmaf(vv.x, vv.x, vv.y * vv.y));
        };
#pragma unroll
        for (int i = 0; i < LH; i++) {
          const float2 *oa = x + (2 * i) * LR, *ob = x + (2 * i + 1) * LR;
          const bool t0s = (i == 0) && (j == 0);   // thread 0: A_0 = 0 pairs with itself, B_0 = S/2 with itself
#pragma unroll
          for (int q = 0; q < LR / 2; q++) {
            const int pidx = i * LR + 2 * q;
            // k = A_i + q*LS pairs with M - k = B_i + (LR-1-q)*LS
            const float2 p1 = t0s ? oa[(LR - q) & (LR - 1)] : ob[LR - 1 - q];
            pair_mag(oa[q], p1, make_float2(tr[2 * pidx], tr[2 * pidx + 1]), val[2 * pidx], val[2 * pidx + 1]);
            // k = B_i + q*LS pairs with M - k = A_i + (LR-1-q)*LS
            const float2 p2 = t0s ? ob[LR - 1 - q] : oa[LR - 1 - q];
            pair_mag(ob[q], p2, make_float2(tr[2 * pidx + 2], tr[2 * pidx + 3]), val[2 * pidx + 2], val[2 * pidx + 3]);
          }
        }
        if (j == 0) {
          const float2 z0 = x[0], zm = x[LR / 2];                    // Z[0] and Z[(LR/2)*LS] = Z[M/2]
          val[0] = 2.0f * fabsf(z0.x + z0.y);                        // X[0]   = Zr + Zi
          val[1] = 2.0f * fabsf(z0.x - z0.y);                        // X[N/2] = Zr - Zi (packed Nyquist)
          vmid = 2.0f * sqrt_approx(fmaf(zm.x, zm.x, zm.y * zm.y));  // X[N/4] = conj(Z[M/2])
        }
      } else {
        // ---- N = 8192: pass 3 (last), one radix-16 butterfly per thread: x[s] = Z[j + s*T] -------------
        {
          float t[32];
          tmem_ld32(ta + C::TM_TW3, t);     // entry k-1
#pragma unroll
          for (int k = 1; k < R3; k++) x[k] = cmul(x[k], make_float2(t[2 * (k - 1)], t[2 * (k - 1) + 1]));
          dft<R3>(x);
        }
        // ---- recombination (src/fft.rs:126-132 bin convention) + magnitude (src/lib.rs:366-372) ------
        // k = beta + 256 s pairs with M - k = (256 - beta) + 256 (15 - s): output 15 - s of the partner lane
        static_assert(T == 256 && R3 == 16, "shuffle-mirrored last pass is written for N = 8192");
        {
          float tr[16];
          tmem_ld16(ta + C::TM_TWR, tr);
          const bool self0 = (beta == 0), self128 = (beta == 128);   // butterflies that pair with themselves
#pragma unroll
          for (int s = 0; s < E / 2; s++) {
            const float2 a = x[s];
            float2 bq;
            bq.x = __shfl_xor_sync(0xffffffffu, x[15 - s].x, 16);
            bq.y = __shfl_xor_sync(0xffffffffu, x[15 - s].y, 16);
            if (self128) bq = x[15 - s];
            if (self0) bq = x[(16 - s) & 15];
            const float2 w = make_float2(tr[2 * s], tr[2 * s + 1]);
            // P = A + conj(B), Q = A - conj(B), T = -i w Q (w = (cos,-sin)), 2X[k] = P + T, 2conj(X[M-k]) = P - T
            const float2 cb = make_float2(bq.x, -bq.y);
            const float2 pp = add2(a, cb), qq = sub2(a, cb);
            const float2 tt = fma2(bcast2(w.x), make_float2(qq.y, -qq.x), mul2(bcast2(w.y), qq));
            const float2 uu = add2(pp, tt), vv = sub2(pp, tt);
            val[2 * s] = sqrt_approx(fmaf(uu.x, uu.x, uu.y * uu.y));
            val[2 * s + 1] = sqrt_approx(fmaf(vv.x, vv.x, vv.y * vv.y));
          }
          if (j == 0) {
            const float2 z0 = x[0], zm = x[8];                        // Z[0] and Z[8*256] = Z[M/2]
            val[0] = 2.0f * fabsf(z0.x + z0.y);                      // X[0]   = Zr + Zi
            val[1] = 2.0f * fabsf(z0.x - z0.y);                      // X[N/2] = Zr - Zi (packed Nyquist)
            vmid = 2.0f * sqrt_approx(fmaf(zm.x, zm.x, zm.y * zm.y)); // X[N/4] = conj(Z[M/2])
          }
        }
      }
    }
    if constexpr (!MIRROR) early_stage(more, frame + fstep);   // (N = 4096: right after its last-pass butterflies)
    const u64 cur = frame;
    frame += fstep;

    // FrequencyLimit slice: `valid` and `valid_mid` are launch constants, computed once in front of the frame loop
#define SA_VALID(i) (FULL || ((valid >> (i)) & 1u))

    // ---- unscaled values u = 0.5*sqrt(..) and, for the modes that need it, max|u| ----------------
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
      ub = __reduce_max_sync(0xffffffffu, ub);
      if constexpr (NW > 1) {
        if (lane == 0) misc[MI_UMAX + wig] = ub;
        group_sync<T>(g);
        if constexpr (NW > 4) {
          ub = __reduce_max_sync(0xffffffffu, lane < NW ? misc[MI_UMAX + lane] : 0u);
        } else {
#pragma unroll
          for (int w = 0; w < NW; w++) ub = max(ub, misc[MI_UMAX + w]);
        }
      }
      umax = __uint_as_float(ub);
    }

    // ---- scaling (src/scaling.rs:104-162) ---------------------------------------------------------
    auto scale1 = [&](float v) -> float {
      if constexpr (MODE == SCALE_MUL) return v * cscale;                                    // exact
      else if constexpr (MODE == SCALE_DIV) return div_const(v * 0.5f, P.div, P.rdiv);
      else if constexpr (MODE == SCALE_ZERO_ONE) return umax != 0.0f ? __fdiv_rn(v, umax) : 0.0f;
      else return (v == 0.0f) ? 0.0f : __fmul_rn(20.0f, log10f_musl(v));
    };
    if constexpr (FOLD_SCALE) {
      // already scaled (the factor rides on the window multipliers)
    } else if constexpr (MODE == SCALE_MUL) {
#pragma unroll
      for (int i = 0; i < E; i += 2) {
        const float2 r = mul2(make_float2(val[i], val[i + 1]), bcast2(cscale));
        val[i] = r.x;
        val[i + 1] = r.y;
      }
    } else if (MODE == SCALE_ZERO_ONE && frame_max_is_fast_divisor(umax)) {   // group-uniform
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
    if constexpr (MODE == SCALE_MUL && !FOLD_SCALE) vmid = scale1(vmid);

    if constexpr (C::SPLIT) {
      // ---- hand-over to the group's statistics warp (sa_spectrum_v3.cuh) -----------------------------
      // per-thread min / max / sum and the non-finite check stay here (the values are in registers); everything
      // that needs the whole frame -- arg bins, median, average, record, the stores -- happens in the other warp
      uint32_t kmn = 0xffffffffu, kmx = 0u;
      float sum = 0.f;
      if constexpr (FULL) {
        float2 acc = make_float2(val[0], val[1]);
#pragma unroll
        for (int i = 2; i < E; i += 2) acc = add2(acc, make_float2(val[i], val[i + 1]));
        sum = acc.x + acc.y;
#pragma unroll
        for (int i = 0; i < E; i += 2) {
          const uint32_t k0 = v2_key<MODE>(val[i]), k1 = v2_key<MODE>(val[i + 1]);
          kmn = min(kmn, min(k0, k1));
          kmx = max(kmx, max(k0, k1));
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; i++) {   // dropped bins become +Inf: above every finite key, never a candidate
          const bool keep = SA_VALID(i);
          if (keep) {
            const uint32_t k = v2_key<MODE>(val[i]);
            kmn = min(kmn, k);
            kmx = max(kmx, k);
            sum += val[i];
          }
          val[i] = keep ? val[i] : __uint_as_float(0x7f800000u);
        }
      }
      if (valid_mid) {
        const uint32_t k = v2_key<MODE>(vmid);
        kmn = min(kmn, k);
        kmx = max(kmx, k);
        sum += vmid;
      }
      const uint32_t my_kmn = kmn, my_kmx = kmx;
      kmn = __reduce_min_sync(0xffffffffu, kmn);
      kmx = __reduce_max_sync(0xffffffffu, kmx);
      // which lanes hold the warp's extremes: the statistics warp only looks at their bins for the arg bins
      const uint32_t bmn = __ballot_sync(0xffffffffu, my_kmn == kmn), bmx = __ballot_sync(0xffffffffu, my_kmx == kmx);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      // non-finite spectrum value (warp-uniform; a NaN / Inf sample makes every bin non-finite, so every warp of
      // the group sees it): classify NaN / Inf in the windowed input, each warp over all N samples
      bool nonfinite;
      if constexpr (MODE == SCALE_LOG10 || MODE == SCALE_ZERO_ONE) nonfinite = (__float_as_uint(umax) >= 0x7f800000u);
      else nonfinite = (kmx >= 0x7f800000u);
      uint32_t fl = 0u;
      if (nonfinite) {
        fl = 4u;
        for (int n = lane; n < N; n += 32) {
          const float v = __fmul_rn(win1(n), load_sample(p, cur * p.frame_stride + n));
          if (v != v) fl |= 1u;
          if (fabsf(v) == INFINITY) fl |= 2u;
        }
        fl = __reduce_or_sync(0xffffffffu, fl);
      }
      // the statistics warp is done with the previous frame: its window for this frame is published
      const int slot = C::NSW > 1 ? par : 0;          // hand-over slot = statistics warp of this frame
      if (fcount >= (uint32_t)C::NSW) { mbar_wait_sleep(mb_empty + slot, (ephase >> slot) & 1u); ephase ^= 1u << slot; }
      fcount++;
      float *vbuf = vbuf0 + slot * C::VALS_FLOATS;
      {
        // window of this frame's parity: published by the statistics warp two frames ago
        const uint2 wp = *reinterpret_cast<const uint2 *>(&misc[MS_WIN + 2 * par]);   // {klo, shift | use << 8}
        if ((wp.y & 0x100u) && !(SA_EXP & 0x8000)) {
          uint32_t *hist = hist0 + slot * C::HSTRIDE + 3;   // hist[0] = below, hist[1 + b] = bucket b, hist[NBUCKET + 1] = above (clean: see the statistics warp)
          const int shw = (int)(wp.y & 0xffu);
#pragma unroll
          for (int i = 0; i < E; i++)   // unconditional: dropped bins are +Inf and count as "above"
            atomicAdd(&hist[hist_slot<MODE, C::NBUCKET>(v2_key<MODE>(val[i]), wp.x, shw)], 1u);
          if (valid_mid) atomicAdd(&hist[hist_slot<MODE, C::NBUCKET>(v2_key<MODE>(vmid), wp.x, shw)], 1u);
        }
      }
      {
        // bin b of the frame at float index b + o of the copy (o: see V2Cfg::G_VALS)
        const uint32_t o = o_pos;
        o_pos = (o_pos + o_step) & 3u;
        float *vb = vbuf + o;
        if constexpr (MIRROR) {
          float *d0 = vb + j, *d1 = vb + (M - j), *d2 = vb + jbB, *d3 = vb + (M - jbB);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            d0[256 * q] = val[4 * q];
            d1[-256 * q] = val[4 * q + 1];
            d2[256 * q] = val[4 * q + 2];
            d3[-256 * q] = val[4 * q + 3];
          }
        } else {
#pragma unroll
          for (int i = 0; i < E; i++) vb[binof(i)] = val[i];
        }
        if (j == 0) vb[M / 2] = valid_mid ? vmid : __uint_as_float(0x7f800000u);
      }
      if (lane == 0) {
        misc[MS_KMIN + 8 * slot + wig] = kmn; misc[MS_KMAX + 8 * slot + wig] = kmx; misc[MS_SUM + 8 * slot + wig] = __float_as_uint(sum);
        misc[MS_BMIN + 4 * slot + wig] = bmn; misc[MS_BMAX + 4 * slot + wig] = bmx;
        if (fl) atomicOr(&misc[MS_FLAGS + slot], fl);
      }
      par ^= 1;
      __syncwarp();
      if (lane == 0) mbar_arrive(mb_full + slot);
    } else {
    if constexpr (ALIAS && !EARLY) {
      if (!want_stats) {   // no statistics barrier to hang the copy on: one extra group barrier
        group_sync<T>(g);
        if (j == 0 && more) stage_frame(frame);
      }
    }
    // ---- statistics of the scaled values (src/spectrum.rs:481-539) --------------------------------
    if (want_stats) {
      // layout of one histogram (HSTRIDE words): [3 pad | below | NBUCKET buckets (16-byte aligned) | above | pad]
      uint32_t *hist_base = hist0 + par * C::HSTRIDE;     // clean on entry
      uint32_t *hist = hist_base + 3;                     // hist[0] = below, hist[1 + b] = bucket b, hist[NBUCKET + 1] = above
      uint32_t *hist_other = hist0 + (par ^ 1) * C::HSTRIDE;  // possibly dirty from the previous frame
      par ^= 1;
      const uint32_t ra = (count & 1u) ? count / 2 : count / 2 - 1, rb = count / 2;   // middle ranks (0-based)
      // Median window guess: keys within gw of the previous frame's median.  The fused pass below
      // histograms all keys against that window (below / NBUCKET buckets / above); the selection stays
      // exact because the ranks are verified against the counts, with the full-range pass as fallback.
      const bool use_win = want_median && gw != 0u;
      const uint32_t gk = use_win ? misc[MI_GUESS] : 0u;
      const uint32_t winL = gk > gw ? gk - gw : 0u;
      const uint32_t winW = 2u * gw;
      const int shw = max(0, (32 - __clz(winW | 1u)) - HBITS);

      uint32_t kmn = 0xffffffffu, kmx = 0u;
      float sum = 0.f;
      if constexpr (FULL) {
        // all 16 owned bins are kept: pairwise (packed) sums and 3-input min/max
        float2 acc = make_float2(val[0], val[1]);
#pragma unroll
        for (int i = 2; i < E; i += 2) acc = add2(acc, make_float2(val[i], val[i + 1]));
        sum = acc.x + acc.y;
#pragma unroll
        for (int i = 0; i < E; i += 2) {
          const uint32_t k0 = v2_key<MODE>(val[i]), k1 = v2_key<MODE>(val[i + 1]);
          kmn = min(kmn, min(k0, k1));
          kmx = max(kmx, max(k0, k1));
        }
      } else {
        // FrequencyLimit slices: min / max / sum under the per-bin predicate; then the dropped bins are replaced
        // by +Inf (its key lies above every finite key in all modes) so that the histogram atomics below stay
        // unconditional like in the All case (a predicated shared-memory atomic becomes a branch): +Inf lands in
        // the "above" slot, which no rank looks at; every later use of a dropped bin is still predicated.
#pragma unroll
        for (int i = 0; i < E; i++) {
          const bool keep = SA_VALID(i);
          if (keep) {
            const uint32_t k = v2_key<MODE>(val[i]);
            kmn = min(kmn, k);
            kmx = max(kmx, k);
            sum += val[i];
          }
          val[i] = keep ? val[i] : __uint_as_float(0x7f800000u);
        }
      }
      if (valid_mid) {
        const uint32_t k = v2_key<MODE>(vmid);
        kmn = min(kmn, k);
        kmx = max(kmx, k);
        sum += vmid;
      }
      if (use_win) {
#pragma unroll
        for (int i = 0; i < E; i++) {   // unconditional: dropped bins are +Inf and count as "above"
          const uint32_t sl = hist_slot<MODE, C::NBUCKET>(v2_key<MODE>(val[i]), winL, shw);
          if (SA_EXP & 0x2000) atomicAdd(&hist[0], 1u);
          else if (!(SA_EXP & 0x800)) atomicAdd(&hist[sl], 1u);
        }
        if (valid_mid) atomicAdd(&hist[hist_slot<MODE, C::NBUCKET>(v2_key<MODE>(vmid), winL, shw)], 1u);
      }
      const uint32_t my_kmn = kmn, my_kmx = kmx;
      if (!(SA_EXP & 32)) {
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        kmx = __reduce_max_sync(0xffffffffu, kmx);
      }
      if (!(SA_EXP & 8)) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      }
      if constexpr (NW > 1 && !(SA_EXP & 4)) {
        if (lane == 0) { misc[MI_KMIN + wig] = kmn; misc[MI_KMAX + wig] = kmx; misc[MI_SUM + wig] = __float_as_uint(sum); }
        group_sync<T>(g);
        // every thread has read its last-pass inputs from the second exchange buffer: the next frame may land there
        if constexpr (ALIAS) { if (j == 0 && more) stage_frame(frame); }
        // the NW partial results of each kind as 16-byte loads (up to four warps per group; with more, holding all of
        // them at once cost more in registers than the loads saved: N = 16384 zero-to-one -3 %)
        if constexpr (NW > 4) {
          // eight or sixteen partial results of each kind: lane w of every warp loads the w-th and the warp reduces them
          // (3 loads + 2 redux + log2(NW) shuffles instead of 3 NW loads and as many min / max / add; the groups of
          // these sizes are limited by the memory-instruction queue, profiles/r02b_cfg4_ncu_lines.txt)
          const bool has = lane < NW;
          const uint32_t pmn = has ? misc[MI_KMIN + lane] : 0xffffffffu, pmx = has ? misc[MI_KMAX + lane] : 0u;
          float ps = has ? __uint_as_float(misc[MI_SUM + lane]) : 0.f;
          kmn = __reduce_min_sync(0xffffffffu, pmn);
          kmx = __reduce_max_sync(0xffffffffu, pmx);
#pragma unroll
          for (int off = NW / 2; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
          sum = __shfl_sync(0xffffffffu, ps, 0);
        } else {
          uint32_t pmn[NW], pmx[NW], psm[NW];
          if constexpr (NW % 4 == 0) {
#pragma unroll
            for (int q = 0; q < NW / 4; q++) {
              const uint4 a = reinterpret_cast<const uint4 *>(misc + MI_KMIN)[q];
              const uint4 b = reinterpret_cast<const uint4 *>(misc + MI_KMAX)[q];
              const uint4 c = reinterpret_cast<const uint4 *>(misc + MI_SUM)[q];
              pmn[4 * q] = a.x; pmn[4 * q + 1] = a.y; pmn[4 * q + 2] = a.z; pmn[4 * q + 3] = a.w;
              pmx[4 * q] = b.x; pmx[4 * q + 1] = b.y; pmx[4 * q + 2] = b.z; pmx[4 * q + 3] = b.w;
              psm[4 * q] = c.x; psm[4 * q + 1] = c.y; psm[4 * q + 2] = c.z; psm[4 * q + 3] = c.w;
            }
          } else {
            static_assert(NW == 2 || NW % 4 == 0, "partial-result loads");
            const uint2 a = *reinterpret_cast<const uint2 *>(misc + MI_KMIN);
            const uint2 b = *reinterpret_cast<const uint2 *>(misc + MI_KMAX);
            const uint2 c = *reinterpret_cast<const uint2 *>(misc + MI_SUM);
            pmn[0] = a.x; pmn[1] = a.y; pmx[0] = b.x; pmx[1] = b.y; psm[0] = c.x; psm[1] = c.y;
          }
          kmn = pmn[0]; kmx = pmx[0]; sum = __uint_as_float(psm[0]);
#pragma unroll
          for (int w = 1; w < NW; w++) {
            kmn = min(kmn, pmn[w]);
            kmx = max(kmx, pmx[w]);
            sum += __uint_as_float(psm[w]);
          }
        }
      } else {
        __syncwarp();
      }
      // arg bins: only threads that hold the extreme value search (min -> lowest bin, max -> highest)
      if (my_kmn == kmn && !(SA_EXP & 2)) {
        uint32_t best = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < E; i++)
          if (SA_VALID(i) && v2_key<MODE>(val[i]) == kmn) best = min(best, binof(i));
        if (valid_mid && v2_key<MODE>(vmid) == kmn) best = min(best, (uint32_t)(M / 2));
        atomicMin(&misc[MI_ARGMIN], best);
      }
      if (my_kmx == kmx && !(SA_EXP & 2)) {
        uint32_t best = 0u;
#pragma unroll
        for (int i = 0; i < E; i++)
          if (SA_VALID(i) && v2_key<MODE>(val[i]) == kmx) best = max(best, binof(i));
        if (valid_mid && v2_key<MODE>(vmid) == kmx) best = max(best, (uint32_t)(M / 2));
        atomicMax(&misc[MI_ARGMAX], best);
      }

      // non-finite spectrum value?  (keys of NaN/Inf sit above every finite key in all modes
      // except 20*log10 / zero-to-one, which check the unscaled maximum)
      bool nonfinite;
      if constexpr (MODE == SCALE_LOG10 || MODE == SCALE_ZERO_ONE) nonfinite = (__float_as_uint(umax) >= 0x7f800000u);
      else nonfinite = (kmx >= 0x7f800000u);
      if (nonfinite) {  // group-uniform rare path: classify NaN / Inf in the windowed input
        uint32_t f = 4u;
        for (int n = j; n < N; n += T) {
          const float v = __fmul_rn(win1(n), load_sample(p, cur * p.frame_stride + n));
          if (v != v) f |= 1u;
          if (fabsf(v) == INFINITY) f |= 2u;
        }
        atomicOr(&misc[MI_FLAGS], f);
      }

      const bool all_equal = (kmx == kmn);
      bool need_slow = false;              // exact selection through the barrier-synchronised loop below
      bool win_failed = false;
      uint32_t klo = kmn, width = kmx - kmn, rka = ra, rkb = rb;
      uint32_t cand_lo = 0u, cand_w = 0xffffffffu;   // keys still in play: (key - cand_lo) <= cand_w
      bool hist_ready = false;

      if (!want_median || (all_equal && !nonfinite)) {
        // nothing to select: thread 0 writes the whole record one barrier later
        if (rec_thread) {
          pend = true; pend_med = true; pend_cur = cur; pend_kmn = kmn; pend_kmx = kmx; pend_avg = __fdiv_rn(sum, (float)count); pend_medkey = kmn;
          if (want_median) misc[MI_GUESS] = kmn;
        }
        if (use_win) {   // the fused pass dirtied this histogram: every warp clears a quarter, used again in 2 frames
          for (int i = j; i < C::HSTRIDE; i += T) hist_other[i] = 0;
        }
        // note: `hist` (dirty when use_win) is cleaned by the next frame's pass over hist_other
        if (gw == 0u && want_median) gw = GW_INIT;
      } else if (use_win && (count & 1u) && !nonfinite) {
        // ---- fast path: no further group barrier -------------------------------------------------
        // every warp scans the buckets redundantly and finds the bucket of the middle rank; the histogram
        // of the previous frame is cleared on the side
        static_assert(HBPL == 8 || HBPL == 4 || HBPL == 2 || HBPL == 1, "scan written for 1, 2, 4 or 8 buckets per lane");
        uint32_t ba, cnt_a, below_a, om;
        if constexpr ((SA_EXP & 0x20000) != 0) {   // timing experiment: no scan
          for (int i = j; i < C::HSTRIDE; i += T) hist_other[i] = 0;
          om = 1u; cnt_a = 1u; ba = 0u; below_a = ra;
        } else if constexpr (HBPL <= 4) {
          // one level: HBPL buckets per lane (one vector load), a warp scan of the lane sums, then the
          // owner lane's buckets are broadcast and searched
          uint32_t hb[HBPL];
          if constexpr (HBPL == 4) { const uint4 h = *reinterpret_cast<const uint4 *>(hist + 1 + lane * 4); hb[0] = h.x; hb[1] = h.y; hb[2] = h.z; hb[3] = h.w; }
          else if constexpr (HBPL == 2) { const uint2 h = *reinterpret_cast<const uint2 *>(hist + 1 + lane * 2); hb[0] = h.x; hb[1] = h.y; }
          else hb[0] = hist[1 + lane];
          const uint32_t nbelow = hist[0];
          for (int i = j; i < C::HSTRIDE; i += T) hist_other[i] = 0;
          uint32_t local = 0;
#pragma unroll
          for (int b = 0; b < HBPL; b++) local += hb[b];
          uint32_t incl = local;
          // inclusive prefix sums of the lane totals (six independent vote + popcount pairs instead of the five
          // dependent shuffles measured SLOWER: 0.638 -> 0.623; the instructions count, not the chain)
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) incl = scan_step_up(incl, off);
          om = __ballot_sync(0xffffffffu, nbelow + incl > ra);
          if (ra < nbelow) om = 0u;
          const int owner = (__ffs(om) - 1) & 31;
          // every lane searches its own buckets as if it were the owner; the owner's answer travels in ONE shuffle:
          // bucket (8 bits) | min(count, 63) << 8 | keys below the bucket << 14  (below <= M + 1 <= 8193 < 2^14)
          uint32_t c = nbelow + incl - local;
          uint32_t ba_l = (uint32_t)(lane * HBPL), cnt_l = 0u, below_l = c;
          bool found = false;
#pragma unroll
          for (int b = 0; b < HBPL; b++) {
            const uint32_t h = hb[b];
            const bool here = !found && (c + h > ra);
            if (here) { ba_l = (uint32_t)(lane * HBPL + b); cnt_l = h; below_l = c; }
            found = found || here;
            c += h;
          }
          static_assert(M + 1 < (1 << 14) && C::NBUCKET <= 256, "packed scan result");
          const uint32_t pk = __shfl_sync(0xffffffffu, ba_l | (min(cnt_l, 63u) << 8) | (below_l << 14), owner);
          ba = pk & 0xffu; cnt_a = (pk >> 8) & 63u; below_a = pk >> 14;
        } else {
        const uint4 h0 = *reinterpret_cast<const uint4 *>(hist + 1 + lane * 8);
        const uint4 h1 = *reinterpret_cast<const uint4 *>(hist + 5 + lane * 8);
        const uint32_t nbelow = hist[0];
        for (int i = j; i < C::HSTRIDE; i += T) hist_other[i] = 0;
        const uint32_t local = (h0.x + h0.y) + (h0.z + h0.w) + ((h1.x + h1.y) + (h1.z + h1.w));
        uint32_t incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += o;
        }
        // first run whose inclusive count exceeds the rank (none: the window missed the middle rank)
        om = __ballot_sync(0xffffffffu, nbelow + incl > ra);
        if (ra < nbelow) om = 0u;
        const int owner = (__ffs(om) - 1) & 31;
        const uint32_t c0o = nbelow + __shfl_sync(0xffffffffu, incl - local, owner);
        const uint32_t h8 = hist[1 + owner * 8 + (lane & 7)];
        uint32_t incl8 = h8;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl8, off, 8);
          if ((lane & 7) >= off) incl8 += o;
        }
        const uint32_t om8 = __ballot_sync(0xffffffffu, c0o + incl8 > ra) & 0xffu;
        const int b8 = (__ffs(om8) - 1) & 7;
        ba = (uint32_t)(owner * 8 + b8);
        cnt_a = __shfl_sync(0xffffffffu, h8, b8);
        below_a = c0o + __shfl_sync(0xffffffffu, incl8 - h8, b8);
        }
        if (SA_EXP & 0x4000) { om = 1u; cnt_a = 1u; ba = 0u; below_a = ra; }   // timing experiment: the fast path always "succeeds"
        const uint32_t lo_a = winL + (ba << shw), bw = (1u << shw) - 1u;
        if (om != 0u && cnt_a <= 32u) {
          if constexpr (DEFRANK) {
            // Every thread keeps its (almost always at most one) candidate of the selected bucket in its own word of
            // a table -- an unconditional store: no branch, no atomic, no return value to wait for.  Further candidates
            // of the same thread (rare) go to an overflow list.  The median is picked from the table one group barrier
            // later (deferred_select, after the first exchange barrier of the group's next frame).
            if (shw == 0) {
              // buckets are single keys: the median is lo_a itself; the record thread writes it with the record
              if (rec_thread) { pend_med = true; pend_medkey = lo_a; misc[MI_GUESS] = lo_a; }
            } else {
              // c_i = min(key_i - lo_a, bw + 1) in one instruction each (the difference wraps for keys below the bucket):
              // <= bw exactly for the keys of the bucket.  Their minimum is this thread's candidate; the sum of the
              // c_i tells whether it is the only one (sum == 15 (bw + 1) + minimum), without a test per value.
              constexpr uint32_t NONE = 0xffffffffu;
              const uint32_t nlo = 0u - lo_a, cap = bw + 1u;
              uint32_t c[E];
#pragma unroll
              for (int i = 0; i < E; i++) c[i] = __viaddmin_u32(v2_key<MODE>(val[i]), nlo, cap);
              uint32_t m = cap, sm = 0u;
              if (!(SA_EXP & 0x400)) {
#pragma unroll
                for (int i = 0; i < E; i += 2) { m = __vimin3_u32(m, c[i], c[i + 1]); sm += c[i] + c[i + 1]; }
              }
              const bool hit = m <= bw;
              if (!(SA_EXP & 0x100000)) cand[j] = hit ? m + lo_a : NONE;          // NONE is never the key of a finite value
              if (!(SA_EXP & 0x100000) && hit && sm != (uint32_t)(E - 1) * cap + m) {
                // rare: several candidates in this thread -- all but one instance of the minimum go to the overflow list
                bool skipped = false;
#pragma unroll
                for (int i = 0; i < E; i++) {
                  if (c[i] <= bw) {
                    if (c[i] == m && !skipped) skipped = true;
                    else list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = c[i] + lo_a;
                  }
                }
              }
              if (valid_mid) {   // thread 0: the bin it owns on top of its sixteen
                const uint32_t k = v2_key<MODE>(vmid);
                if ((k - lo_a) <= bw) list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = k;
              }
              if (rec_thread) {
                pend_med = false;
                *reinterpret_cast<uint4 *>(&misc[MI_DEF]) = make_uint4((SA_EXP & 0x200000) ? 0u : cnt_a, ra - below_a, (uint32_t)cur, (uint32_t)(cur >> 32));
              }
            }
            if (rec_thread) { pend = true; pend_cur = cur; pend_kmn = kmn; pend_kmx = kmx; pend_avg = __fdiv_rn(sum, (float)count); }
            gw = max(gw - (gw >> SA_GW_SHRINK), GW_MIN);
          } else {
          if (shw == 0) {
            // buckets are single keys: the median is lo_a itself; thread 0 writes it with the record
            if (rec_thread) { pend_med = true; pend_medkey = lo_a; misc[MI_GUESS] = lo_a; }
          } else {
            // candidates of the selected bucket go to the list; the thread that completes the list
            // ranks it and writes the median (and the next guess) -- nobody waits for it
            // branch-free count of this thread's candidates (plus the last one); almost always 0 or 1
            uint32_t npush = 0, ck = 0;
#pragma unroll
            for (int i = 0; i < E; i++) {
              const uint32_t k = v2_key<MODE>(val[i]);
              const bool c = SA_VALID(i) && (k - lo_a) <= bw;
              ck = c ? k : ck;
              npush += c ? 1u : 0u;
            }
            if (valid_mid) {
              const uint32_t k = v2_key<MODE>(vmid);
              const bool c = (k - lo_a) <= bw;
              ck = c ? k : ck;
              npush += c ? 1u : 0u;
            }
            if (npush == 1u) {
              list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = ck;
            } else if (npush > 1u) {
#pragma unroll
              for (int i = 0; i < E; i++) {   // rare: several candidates in one thread
                const uint32_t k = v2_key<MODE>(val[i]);
                if (SA_VALID(i) && (k - lo_a) <= bw) list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = k;
              }
              if (valid_mid) {
                const uint32_t k = v2_key<MODE>(vmid);
                if ((k - lo_a) <= bw) list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = k;
              }
            }
            bool fin = false;
            if (npush) {
              __threadfence_block();
              fin = (atomicAdd(&misc[MI_LDONE], npush) + npush == cnt_a);
            }
            // the warp of the thread that completed the list ranks it together: one candidate per lane,
            // stable order (src/spectrum.rs:515-524 sorts, ties keep their order -- equal keys are equal values)
            if ((SA_EXP & 128) && fin) {   // experiment: the completing thread ranks the list alone
              __threadfence_block();
              const volatile uint32_t *vl = list;
              const uint32_t want = ra - below_a;
              uint32_t res = 0;
              for (uint32_t a = 0; a < cnt_a; a++) {
                const uint32_t kk = vl[a];
                uint32_t r = 0;
                for (uint32_t b = 0; b < cnt_a; b++) {
                  const uint32_t ko = vl[b];
                  r += (ko < kk || (ko == kk && b < a)) ? 1u : 0u;
                }
                if (r == want) res = kk;
              }
              reinterpret_cast<uint32_t *>(&p.stats[cur])[5] = __float_as_uint(v2_unkey<MODE>(res));
              misc[MI_GUESS] = res;
              misc[MI_LCOUNT] = 0u;
              misc[MI_LDONE] = 0u;
            }
            if (!(SA_EXP & 128) && __any_sync(0xffffffffu, fin)) {
              __threadfence_block();
              const uint32_t mine = (lane < (int)cnt_a) ? reinterpret_cast<const volatile uint32_t *>(list)[lane] : 0xffffffffu;
              uint32_t rank = 0;
#pragma unroll 4
              for (uint32_t c = 0; c < cnt_a; c++) {
                const uint32_t o = __shfl_sync(0xffffffffu, mine, c);
                rank += (o < mine || (o == mine && c < (uint32_t)lane)) ? 1u : 0u;
              }
              const uint32_t hit = __ballot_sync(0xffffffffu, lane < (int)cnt_a && rank == ra - below_a);
              const uint32_t res = __shfl_sync(0xffffffffu, mine, (__ffs(hit) - 1) & 31);
              if (lane == 0) {
                reinterpret_cast<uint32_t *>(&p.stats[cur])[5] = __float_as_uint(v2_unkey<MODE>(res));
                misc[MI_GUESS] = res;
                misc[MI_LCOUNT] = 0u;
                misc[MI_LDONE] = 0u;
              }
            }
            if (rec_thread) pend_med = false;
          }
          if (rec_thread) { pend = true; pend_cur = cur; pend_kmn = kmn; pend_kmx = kmx; pend_avg = __fdiv_rn(sum, (float)count); }
          gw = max(gw - (gw >> SA_GW_SHRINK), GW_MIN);
          }
        } else {
          // the guessed window missed the middle rank, or its bucket is crowded: exact loop below
          group_sync<T>(g);                                   // every warp has finished scanning
          for (int i = j; i < C::HSTRIDE; i += T) hist_base[i] = 0;
          group_sync<T>(g);
          need_slow = true;
          if (om == 0u) { win_failed = true; }
          else { klo = lo_a; width = bw; cand_lo = lo_a; cand_w = bw; rka = ra - below_a; rkb = rka; }
        }
      } else {
        need_slow = true;
        for (int i = j; i < C::HSTRIDE; i += T) hist_other[i] = 0;   // may be dirty from a fast-path frame
        if (use_win) { klo = winL; width = winW; hist_ready = true; }
      }

      // ---- exact selection, barrier-synchronised (first frame, even counts, fallbacks) ---------------
      if (need_slow) {
        uint32_t ka = kmn, kb = kmn;
        bool synced = false;                 // a group barrier has passed since the arg-bin / flag atomics
#pragma unroll 1
        while (!all_equal || hist_ready) {   // all-equal frames only need the fused histogram cleared
          const int sh = max(0, (32 - __clz(width | 1u)) - HBITS);
          if (!hist_ready) {
            // narrowed passes only count keys that were inside the previous bucket: cand_lo/cand_w
#pragma unroll
            for (int i = 0; i < E; i++) {
              const uint32_t k = v2_key<MODE>(val[i]);
              if (SA_VALID(i) && (k - cand_lo) <= cand_w) atomicAdd(&hist[hist_slot<MODE, C::NBUCKET>(k, klo, sh)], 1u);
            }
            if (valid_mid) {
              const uint32_t k = v2_key<MODE>(vmid);
              if ((k - cand_lo) <= cand_w) atomicAdd(&hist[hist_slot<MODE, C::NBUCKET>(k, klo, sh)], 1u);
            }
            group_sync<T>(g);
          }
          hist_ready = false;
          // warp 0 of the group scans the buckets (8 per lane), finds the buckets of the two middle
          // ranks and clears the histogram for its next use
          if (wig == 0) {
            uint32_t hb[HBPL];
            uint32_t local = 0;
#pragma unroll
            for (int b = 0; b < HBPL; b++) { hb[b] = hist[1 + lane * HBPL + b]; local += hb[b]; }
            const uint32_t nbelow = hist[0];
#pragma unroll
            for (int b = 0; b < HBPL; b++) hist[1 + lane * HBPL + b] = 0;
            __syncwarp();
            if (lane == 0) { hist[0] = 0; hist[C::NBUCKET + 1] = 0; misc[MI_SELA] = 0xffffffffu; misc[MI_SELB] = 0xffffffffu; }
            uint32_t incl = local;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
              if (lane >= off) incl += o;
            }
            __syncwarp();
            uint32_t c = nbelow + incl - local;   // keys counted before this lane's first bucket
#pragma unroll
            for (int b = 0; b < HBPL; b++) {
              if (rka >= c && rka < c + hb[b]) { misc[MI_SELA] = lane * HBPL + b; misc[MI_SELA + 1] = c; misc[MI_SELA + 2] = hb[b]; }
              if (rkb >= c && rkb < c + hb[b]) { misc[MI_SELB] = lane * HBPL + b; }
              c += hb[b];
            }
          }
          group_sync<T>(g);
          synced = true;
          if (all_equal) break;
          const uint32_t ba = misc[MI_SELA], below_a = misc[MI_SELA + 1], cnt_a = misc[MI_SELA + 2];
          const uint32_t bb = misc[MI_SELB];
          if (ba == 0xffffffffu || bb == 0xffffffffu) {
            // the guessed window does not contain both middle ranks: exact full-range pass instead
            win_failed = true;
            klo = kmn; width = kmx - kmn; rka = ra; rkb = rb;
            cand_lo = 0u; cand_w = 0xffffffffu;
            continue;
          }
          if (sh == 0) {  // buckets are single keys
            ka = klo + ba;
            kb = klo + bb;
            break;
          }
          const uint32_t lo_a = klo + (ba << sh), lo_b = klo + (bb << sh), bw = (1u << sh) - 1u;
          if (ba != bb) {
            // the two middle ranks fall into different buckets: s[ra] is the largest key of bucket a,
            // s[rb] the smallest key of bucket b
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
            ma = __reduce_max_sync(0xffffffffu, ma);
            mb = __reduce_min_sync(0xffffffffu, mb);
            if (j == 0) { misc[MI_RESA] = 0u; misc[MI_RESB] = 0xffffffffu; }
            group_sync<T>(g);
            if (lane == 0) { atomicMax(&misc[MI_RESA], ma); atomicMin(&misc[MI_RESB], mb); }
            group_sync<T>(g);
            ka = misc[MI_RESA];
            kb = misc[MI_RESB];
            break;
          }
          if (cnt_a <= 32) {
            // compaction of the selected bucket: a warp vote skips the warps that hold no candidate
            uint32_t tmin = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < E; i++)
              if (SA_VALID(i)) tmin = min(tmin, v2_key<MODE>(val[i]) - lo_a);
            if (valid_mid) tmin = min(tmin, v2_key<MODE>(vmid) - lo_a);
            if (__any_sync(0xffffffffu, tmin <= bw)) {
              if (tmin <= bw) {
#pragma unroll
                for (int i = 0; i < E; i++) {
                  const uint32_t k = v2_key<MODE>(val[i]);
                  if (SA_VALID(i) && (k - lo_a) <= bw) list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = k;
                }
                if (valid_mid) {
                  const uint32_t k = v2_key<MODE>(vmid);
                  if ((k - lo_a) <= bw) list[atomicAdd(&misc[MI_LCOUNT], 1u) & 31u] = k;
                }
              }
            }
            group_sync<T>(g);
            if (wig == 0) {
              const uint32_t mine = (lane < cnt_a) ? list[lane] : 0xffffffffu;
              uint32_t rank = 0;
              for (uint32_t c = 0; c < cnt_a; c++) {
                const uint32_t o = __shfl_sync(0xffffffffu, mine, c);
                rank += (o < mine || (o == mine && c < (uint32_t)lane)) ? 1u : 0u;
              }
              const uint32_t wa = rka - below_a, wb = rkb - below_a;
              const uint32_t ma = __ballot_sync(0xffffffffu, lane < cnt_a && rank == wa);
              const uint32_t mb = __ballot_sync(0xffffffffu, lane < cnt_a && rank == wb);
              ka = __shfl_sync(0xffffffffu, mine, __ffs(ma) - 1);
              kb = __shfl_sync(0xffffffffu, mine, __ffs(mb) - 1);
              if (lane == 0) misc[MI_LCOUNT] = 0u;
            }
            break;  // only warp 0 (thread 0 writes the record) holds ka/kb
          }
          // crowded bucket (many near-equal values): narrow the candidate set and repeat
          klo = lo_a;
          width = bw;
          cand_lo = lo_a;
          cand_w = bw;
          rka -= below_a;
          rkb -= below_a;
        }
        // adapt the window for the next frame (all values here are group-uniform)
        if (gw == 0u) gw = GW_INIT;
        else if (win_failed) gw = (gw < (1u << 28)) ? gw << SA_GW_GROW : gw;
        if (!synced) group_sync<T>(g);
        if (j == 0) {
          // a group barrier has passed since the arg-bin / flag atomics: write the record now
          misc[MI_GUESS] = kb;
          const bool mm = smask & SA_STATS_MINMAX;
          uint4 r0, r1;
          r0.x = mm ? __float_as_uint(v2_unkey<MODE>(kmn)) : 0u;
          r0.y = mm ? misc[MI_ARGMIN] : 0u;
          r0.z = mm ? __float_as_uint(v2_unkey<MODE>(kmx)) : 0u;
          r0.w = mm ? misc[MI_ARGMAX] : 0u;
          r1.x = (smask & SA_STATS_AVERAGE) ? __float_as_uint(__fdiv_rn(sum, (float)count)) : 0u;
          const float med = (count & 1u) ? v2_unkey<MODE>(kb)
                                         : __fdiv_rn(__fadd_rn(v2_unkey<MODE>(ka), v2_unkey<MODE>(kb)), 2.0f);
          r1.y = __float_as_uint(med);
          const uint32_t fl = misc[MI_FLAGS];
          r1.z = (fl & 1u) ? SA_ERR_NAN_VALUES : (fl & 2u) ? SA_ERR_INFINITY_VALUES : (fl & 4u) ? SA_ERR_NON_FINITE_RESULT : SA_OK;
          r1.w = 0u;
          uint4 *rec = reinterpret_cast<uint4 *>(&p.stats[cur]);
          rec[0] = r0;
          rec[1] = r1;
        }
      }
    }
    // ---- store the surviving bins: after the statistics, whose dependent shuffle / shared-memory chains
    // would otherwise queue behind 17 global stores per thread in the memory pipe -----------------------
    if (!(SA_EXP & 256)) {
      float *dst = p.out_vals + cur * p.out_stride - lo;
      if constexpr (MIRROR) {
        // four address registers, compile-time offsets of +-256 q
        float *d0 = dst + j, *d1 = dst + (M - j), *d2 = dst + jbB, *d3 = dst + (M - jbB);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (SA_VALID(4 * q)) d0[256 * q] = val[4 * q];
          if (SA_VALID(4 * q + 1)) d1[-256 * q] = val[4 * q + 1];
          if (SA_VALID(4 * q + 2)) d2[256 * q] = val[4 * q + 2];
          if (SA_VALID(4 * q + 3)) d3[-256 * q] = val[4 * q + 3];
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; i++)
          if (SA_VALID(i)) dst[binof(i)] = val[i];
      }
      if (valid_mid) dst[M / 2] = vmid;
    }
    }   // !SPLIT

#undef SA_VALID
    if (!more) break;
  }
  if (want_stats && !C::SPLIT) {
    group_sync<T>(g);                 // every thread of the group has finished its last frame
    if (want_median) deferred_select();
    if (rec_thread) flush_record();
  }
}

template <class C, int MODE, bool FULL, bool FASTIN = false, int SMASK = -1>
__global__ void __launch_bounds__(C::NT, 1) spectrum_kernel_v2(const V2Params P) {
  extern __shared__ __align__(16) unsigned char smem[];
  // tensor-memory allocation for the per-thread tables (one warp allocates, everybody reads the address)
  uint32_t *slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMSLOT);
  if (threadIdx.x < 32) tmem_alloc<C::TM_COLS>(slot);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tm_base = *slot;
  spectrum_body_v2<C, MODE, FULL, FASTIN, SMASK>(P, smem, tm_base);
  // groups finish at different times: the allocation is released once every thread is done with it
  tmem_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<C::TM_COLS>(tm_base);
}

}  // namespace sa
