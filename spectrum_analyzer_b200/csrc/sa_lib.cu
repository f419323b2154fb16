// sa_lib.cu -- C ABI of libspectrum_analyzer_cuda.so (see include/spectrum_analyzer_cuda.h).
//
// Host side: context (device, stream, cached twiddle/window tables, staging buffers), argument
// validation with the reference's error taxonomy and precedence (src/lib.rs:164-181, 317), kernel
// dispatch per N, and the host-buffer pipeline (pinned staging, H2D / compute / D2H on three
// streams).  There is no CPU compute path: every spectrum is produced by the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

// frame groups per CTA of the 4096-point kernel (measured: 4 groups 1.58e8, 5 groups at 94 registers 1.64e8, 6 groups at 80 registers 1.61e8 spectra/s; -DSA_G12=n)
#ifndef SA_G10
#define SA_G10 20   // 1024-point kernel: 20 one-warp groups
#endif
#ifndef SA_G11
#define SA_G11 10   // 2048-point kernel: 10 groups of 64 threads
#endif
#ifndef SA_G13
#define SA_G13 3    // 8192-point kernel: 3 groups of 256 threads
#endif
#ifndef SA_G12
#define SA_G12 5
#endif
// frame groups per CTA of the warp-specialised 4096-point kernel (sa_spectrum_v3.cuh): G groups of 4 transform
// warps + G statistics warps
#ifndef SA_G12S
#define SA_G12S 4
#endif
#ifndef SA_NSW12     // statistics warps per group
#define SA_NSW12 2
#endif

#include "sa_aux_kernels.cuh"
#include "sa_host_math.h"
#include "sa_spectrum_kernel.cuh"
#include "sa_spectrum_v2.cuh"
#include "sa_spectrum_v3.cuh"
#include "sa_spectrum_small.cuh"

namespace {

thread_local std::string g_last_error;

int cuda_fail(cudaError_t e, const char *what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return SA_ERR_CUDA;
}
#define SA_CUDA(call)                                          \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
  } while (0)

int log2_exact(uint64_t n) {
  int l = 0;
  while ((1ull << l) < n) l++;
  return l;
}

struct KernelPlan {
  int max_blocks = 0;   // resident CTAs on the whole device
  bool ready = false;
};

}  // namespace

struct sa_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  uint64_t launches = 0;
  char last_kernel[96] = "";  // name of the last spectrum kernel launched (sa_ctx_last_kernel)
  bool use_generic = false;   // SA_FORCE_GENERIC=1: run the generic kernel family for every N (A/B testing)
  unsigned stagger_ns = 0;    // SA_STAGGER_NS: start offset between the frame groups of a v2 CTA (tuning; measured: no effect)
  bool fast_log10 = false;    // sa_ctx_set_fast_log10
  bool split = false;         // SA_SPLIT=1: the warp-specialised variant (sa_spectrum_v3.cuh) where one exists; measured slower
                              // than the one-role kernel at cfg2 (1.65e8 vs 1.70e8 spectra/s, profiles/README.md), so opt-in
  std::map<uint32_t, float2 *> twiddles;                   // n -> W_n table
  std::map<std::pair<int, uint32_t>, float *> windows;     // (kind, n) -> multipliers
  KernelPlan plans[16];                                     // per log2n (generic family)
  KernelPlan plans_v2[16][4][2][3];                         // tuned family: [log2n][scale mode][full range][general | f32 input on 16-byte boundaries | + all statistics]
  KernelPlan plans_v3[16][4][2];                            // warp-specialised variant of the tuned family
  std::map<uint32_t, float2 *> v2_tables;                   // n -> tw2 | tw3 | twr (one allocation)
  KernelPlan plans_small[16][4][2];                         // small-N family
  std::map<uint32_t, float2 *> small_tables;                // n -> tw2 | twr
  // host-buffer pipeline
  cudaStream_t h2d = nullptr, d2h = nullptr;
  static constexpr int kSlots = 3;
  float *slot_in[kSlots] = {nullptr, nullptr, nullptr};
  float *slot_out[kSlots] = {nullptr, nullptr, nullptr};
  sa_frame_stats *slot_stats[kSlots] = {nullptr, nullptr, nullptr};
  size_t slot_in_bytes = 0, slot_out_bytes = 0, slot_stats_bytes = 0;
  cudaEvent_t ev_h2d[kSlots] = {}, ev_comp[kSlots] = {}, ev_d2h[kSlots] = {};
  bool pipeline_ready = false;
  // pageable callers: pinned staging ring (one block per pipeline slot and direction) instead of page-locking the
  // caller's whole buffers for the call (measured 5x slower than pinned buffers: cudaHostRegister of gigabytes)
  unsigned char *pin_in[kSlots] = {nullptr, nullptr, nullptr};
  unsigned char *pin_out[kSlots] = {nullptr, nullptr, nullptr};
  unsigned char *pin_st[kSlots] = {nullptr, nullptr, nullptr};
  size_t pin_in_bytes = 0, pin_out_bytes = 0, pin_st_bytes = 0;
  bool host_register = false;   // SA_HOST_PIN=register: the old behaviour (A/B testing)
  // single-frame call (sa_spectrum_single): one pinned host block and one device block, reused
  unsigned char *single_pin = nullptr, *single_dev = nullptr;
  size_t single_bytes = 0;
};

namespace {

using namespace sa;

int get_twiddle(sa_ctx *ctx, uint32_t n, const float2 **out) {
  auto it = ctx->twiddles.find(n);
  if (it == ctx->twiddles.end()) {
    std::vector<float> host = sa_host::twiddle_table(n);
    float2 *dev = nullptr;
    SA_CUDA(cudaMalloc(&dev, host.size() * sizeof(float)));
    SA_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` goes out of scope
    it = ctx->twiddles.emplace(n, dev).first;
  }
  *out = it->second;
  return SA_OK;
}

int get_window(sa_ctx *ctx, int kind, uint32_t n, const float **out) {
  if (kind == SA_WINDOW_NONE) { *out = nullptr; return SA_OK; }
  auto key = std::make_pair(kind, n);
  auto it = ctx->windows.find(key);
  if (it == ctx->windows.end()) {
    std::vector<float> host(n);
    for (uint32_t i = 0; i < n; i++) host[i] = sa_host::window_multiplier(kind, i, n);
    float *dev = nullptr;
    SA_CUDA(cudaMalloc(&dev, n * sizeof(float)));
    SA_CUDA(cudaMemcpyAsync(dev, host.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->windows.emplace(key, dev).first;
  }
  *out = it->second;
  return SA_OK;
}

template <class C>
int launch_cfg(sa_ctx *ctx, const SpectrumParams &p, int log2n) {
  KernelPlan &plan = ctx->plans[log2n];
  auto kernel = spectrum_kernel<C>;
  if (!plan.ready) {
    SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_TOTAL));
    int per_sm = 0;
    SA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, C::NT, C::SMEM_TOTAL));
    if (per_sm < 1) { g_last_error = "kernel does not fit on an SM"; return SA_ERR_CUDA; }
    plan.max_blocks = per_sm * ctx->num_sms;
    plan.ready = true;
  }
  const u64 needed = (p.n_frames + C::FPB - 1) / C::FPB;
  const unsigned grid = (unsigned)std::min<u64>(needed, (u64)plan.max_blocks);
  kernel<<<grid, C::NT, C::SMEM_TOTAL, ctx->stream>>>(p);
  ctx->launches++;
  snprintf(ctx->last_kernel, sizeof ctx->last_kernel, "sa::spectrum_kernel<SpectrumCfg<%d,%d,%d>>", log2n, C::E, C::FPB);
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

// ---- tuned family (N = 1024..8192) ----------------------------------------------------------------
template <class C>
int get_v2_tables(sa_ctx *ctx, V2Params &P) {
  auto it = ctx->v2_tables.find(C::N);
  if (it == ctx->v2_tables.end()) {
    const std::vector<float> w = sa_host::twiddle_table(C::N);  // W_N[i], i < N
    std::vector<float> host(2 * (size_t)(C::NTW2 + C::NTW3 + C::NTWR + C::NTW4));
    size_t o = 0;
    auto put = [&](size_t idx) { host[o++] = w[2 * idx]; host[o++] = w[2 * idx + 1]; };
    for (int k = 1; k < 16; k++)                       // pass 2: w_256^(k p) = W_N[k p N/256]
      for (int pp = 0; pp < 16; pp++) put((size_t)k * pp * (C::N / 256));
    for (int k = 1; k < C::R3; k++)                    // pass 3: w_(256 R3)^(k p) = W_N[k p N/(256 R3)]
      for (int pp = 0; pp < 256; pp++) put((size_t)k * pp * (C::N / (256 * C::R3)));
    for (int k = 0; k < C::NTWR; k++) put((size_t)k);  // recombination: W_N^k
    if (C::FOUR)
      for (int k = 1; k < C::R4; k++)                  // pass 4: w_M^(k p) = W_N[2 k p]
        for (int pp = 0; pp < 4096; pp++) put((size_t)2 * k * pp);
    float2 *dev = nullptr;
    SA_CUDA(cudaMalloc(&dev, host.size() * sizeof(float)));
    SA_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->v2_tables.emplace(C::N, dev).first;
  }
  P.tw2 = it->second;
  P.tw3 = P.tw2 + C::NTW2;
  P.twr = P.tw3 + C::NTW3;
  P.tw4 = P.twr + C::NTWR;
  return SA_OK;
}

template <class C, int MODE, bool FULL, bool FASTIN = false, int SMASK = -1>
int launch_v2_inst(sa_ctx *ctx, const V2Params &P) {
  if constexpr (!FASTIN) {
    // f32 frames that all start on 16-byte boundaries: the specialisation without the per-frame input-format tests
    if (P.b.aligned16 && !P.b.in_i16) return launch_v2_inst<C, MODE, FULL, true>(ctx, P);
  } else if constexpr (SMASK < 0) {
    // ... and with all statistics wanted (what the reference always computes): the mask is a constant as well
    if (P.b.stats != nullptr && P.b.stats_mask == SA_STATS_ALL) return launch_v2_inst<C, MODE, FULL, true, (int)SA_STATS_ALL>(ctx, P);
  }
  KernelPlan &plan = ctx->plans_v2[C::LOG2N][MODE][FULL ? 1 : 0][FASTIN ? (SMASK >= 0 ? 2 : 1) : 0];
  auto kernel = spectrum_kernel_v2<C, MODE, FULL, FASTIN, SMASK>;
  if (!plan.ready) {
    SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_TOTAL));
    int per_sm = 0;
    SA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, C::NT, C::SMEM_TOTAL));
    if (per_sm < 1) { g_last_error = "v2 kernel does not fit on an SM"; return SA_ERR_CUDA; }
    plan.max_blocks = per_sm * ctx->num_sms;
    plan.ready = true;
  }
  const u64 needed = (P.b.n_frames + C::G - 1) / C::G;
  const unsigned grid = (unsigned)std::min<u64>(needed, (u64)plan.max_blocks);
  kernel<<<grid, C::NT, C::SMEM_TOTAL, ctx->stream>>>(P);
  ctx->launches++;
  snprintf(ctx->last_kernel, sizeof ctx->last_kernel, "sa::spectrum_kernel_v2<V2Cfg<%d,%d,0>,%d,%d,%d,%d>", C::LOG2N, C::G, MODE, FULL ? 1 : 0, FASTIN ? 1 : 0, SMASK);
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

// warp-specialised variant: same tables, NT + 32 G threads
template <class C, int MODE, bool FULL>
int launch_v3_inst(sa_ctx *ctx, const V2Params &P) {
  KernelPlan &plan = ctx->plans_v3[C::LOG2N][MODE][FULL ? 1 : 0];
  auto kernel = spectrum_kernel_v3<C, MODE, FULL>;
  if (!plan.ready) {
    SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_TOTAL));
    int per_sm = 0;
    SA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, C::NT_ALL, C::SMEM_TOTAL));
    if (per_sm < 1) { g_last_error = "v3 kernel does not fit on an SM"; return SA_ERR_CUDA; }
    plan.max_blocks = per_sm * ctx->num_sms;
    plan.ready = true;
  }
  const u64 needed = (P.b.n_frames + C::G - 1) / C::G;
  const unsigned grid = (unsigned)std::min<u64>(needed, (u64)plan.max_blocks);
  kernel<<<grid, C::NT_ALL, C::SMEM_TOTAL, ctx->stream>>>(P);
  ctx->launches++;
  snprintf(ctx->last_kernel, sizeof ctx->last_kernel, "sa::spectrum_kernel_v3<V2Cfg<%d,%d,%d>,%d,%d>", C::LOG2N, C::G, C::NSW, MODE, FULL ? 1 : 0);
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

// the warp-specialised configuration of a size (void: none)
template <int LOG2N> struct SplitCfg { using type = void; };
template <> struct SplitCfg<12> { using type = V2Cfg<12, SA_G12S, SA_NSW12>; };

template <class C>
int launch_v2(sa_ctx *ctx, const SpectrumParams &p) {
  V2Params P;
  P.b = p;
  int e = get_v2_tables<C>(ctx, P);
  if (e != SA_OK) return e;
  const bool full = p.bin_lo == 0 && p.bin_hi == (uint32_t)C::M;
  int mode = SCALE_MUL;
  P.cmul = 0.5f;
  P.div = 1.0f;
  P.rdiv = 1.0f;
  P.stagger_ns = ctx->stagger_ns;
  switch (p.scaling_kind) {
    case SA_SCALING_NONE: break;
    case SA_SCALING_DIVIDE_BY_N: P.cmul = 0.5f / (float)C::N; break;                  // exact: N = 2^k
    case SA_SCALING_DIVIDE_BY_N_SQRT:
      if (C::LOG2N % 2 == 0) P.cmul = 0.5f / p.scale_div;                              // sqrt(N) = 2^(k/2), exact
      else { mode = SCALE_DIV; P.div = p.scale_div; P.rdiv = 1.0f / p.scale_div; }
      break;
    case SA_SCALING_ZERO_TO_ONE: mode = SCALE_ZERO_ONE; break;
    default: mode = SCALE_LOG10; break;
  }
  using CS = typename SplitCfg<C::LOG2N>::type;
  if constexpr (!std::is_void<CS>::value) {
    // statistics wanted: the warp-specialised kernel (the statistics warp is what writes records and rows)
    if (p.stats != nullptr && ctx->split) {
      switch (mode) {
        case SCALE_MUL: return full ? launch_v3_inst<CS, SCALE_MUL, true>(ctx, P) : launch_v3_inst<CS, SCALE_MUL, false>(ctx, P);
        case SCALE_DIV: return full ? launch_v3_inst<CS, SCALE_DIV, true>(ctx, P) : launch_v3_inst<CS, SCALE_DIV, false>(ctx, P);
        case SCALE_ZERO_ONE: return full ? launch_v3_inst<CS, SCALE_ZERO_ONE, true>(ctx, P) : launch_v3_inst<CS, SCALE_ZERO_ONE, false>(ctx, P);
        default: return full ? launch_v3_inst<CS, SCALE_LOG10, true>(ctx, P) : launch_v3_inst<CS, SCALE_LOG10, false>(ctx, P);
      }
    }
  }
#define SA_V2_CASE(M_) \
  case M_: return full ? launch_v2_inst<C, M_, true>(ctx, P) : launch_v2_inst<C, M_, false>(ctx, P);
  switch (mode) {
    SA_V2_CASE(SCALE_MUL)
    SA_V2_CASE(SCALE_DIV)
    SA_V2_CASE(SCALE_ZERO_ONE)
    default: return full ? launch_v2_inst<C, SCALE_LOG10, true>(ctx, P) : launch_v2_inst<C, SCALE_LOG10, false>(ctx, P);
  }
#undef SA_V2_CASE
}

// ---- small-N family (N = 64..512) -------------------------------------------------------------------
template <class C>
int get_small_tables(sa_ctx *ctx, V2Params &P) {
  auto it = ctx->small_tables.find(C::N);
  if (it == ctx->small_tables.end()) {
    const std::vector<float> w = sa_host::twiddle_table(C::N);
    std::vector<float> host(2 * (size_t)(C::NTW2 + C::NTWR));
    size_t o = 0;
    auto put = [&](size_t idx) { host[o++] = w[2 * idx]; host[o++] = w[2 * idx + 1]; };
    for (int k = 1; k < C::R2; k++)                    // pass 2: w_M^(k p) = W_N[2 k p]
      for (int pp = 0; pp < 16; pp++) put((size_t)2 * k * pp);
    for (int k = 0; k < C::NTWR; k++) put((size_t)k);  // recombination: W_N^k
    float2 *dev = nullptr;
    SA_CUDA(cudaMalloc(&dev, host.size() * sizeof(float)));
    SA_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->small_tables.emplace(C::N, dev).first;
  }
  P.tw2 = it->second;
  P.tw3 = nullptr;
  P.twr = P.tw2 + C::NTW2;
  return SA_OK;
}

template <class C, int MODE, bool FULL>
int launch_small_inst(sa_ctx *ctx, const V2Params &P) {
  KernelPlan &plan = ctx->plans_small[C::LOG2N][MODE][FULL ? 1 : 0];
  auto kernel = spectrum_kernel_small<C, MODE, FULL>;
  if (!plan.ready) {
    SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_TOTAL));
    int per_sm = 0;
    SA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, C::NT, C::SMEM_TOTAL));
    if (per_sm < 1) { g_last_error = "small kernel does not fit on an SM"; return SA_ERR_CUDA; }
    // The occupancy calculator answers 1 for kernels that allocate tensor memory; two CTAs of 256 threads
    // (<= 128 registers by __launch_bounds__, 48 KB of shared memory, 128 of 512 TMEM columns each) do fit.
    // The grid is persistent (frames strided by gridDim), so a CTA that had to wait would still be correct.
    per_sm = std::max(per_sm, SA_SMALL_CTAS);
    plan.max_blocks = per_sm * ctx->num_sms;
    plan.ready = true;
  }
  const u64 per_cta = (u64)C::WARPS * C::FPW;
  const u64 needed = (P.b.n_frames + per_cta - 1) / per_cta;
  const unsigned grid = (unsigned)std::min<u64>(needed, (u64)plan.max_blocks);
  kernel<<<grid, C::NT, C::SMEM_TOTAL, ctx->stream>>>(P);
  ctx->launches++;
  snprintf(ctx->last_kernel, sizeof ctx->last_kernel, "sa::spectrum_kernel_small<SmallCfg<%d>,%d,%d>", C::LOG2N, MODE, FULL ? 1 : 0);
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

template <class C>
int launch_small(sa_ctx *ctx, const SpectrumParams &p) {
  V2Params P;
  P.b = p;
  int e = get_small_tables<C>(ctx, P);
  if (e != SA_OK) return e;
  const bool full = p.bin_lo == 0 && p.bin_hi == (uint32_t)C::M;
  int mode = SCALE_MUL;
  P.cmul = 0.5f;
  P.div = 1.0f;
  P.rdiv = 1.0f;
  switch (p.scaling_kind) {
    case SA_SCALING_NONE: break;
    case SA_SCALING_DIVIDE_BY_N: P.cmul = 0.5f / (float)C::N; break;
    case SA_SCALING_DIVIDE_BY_N_SQRT:
      if (C::LOG2N % 2 == 0) P.cmul = 0.5f / p.scale_div;
      else { mode = SCALE_DIV; P.div = p.scale_div; P.rdiv = 1.0f / p.scale_div; }
      break;
    case SA_SCALING_ZERO_TO_ONE: mode = SCALE_ZERO_ONE; break;
    default: mode = SCALE_LOG10; break;
  }
  switch (mode) {
    case SCALE_MUL: return full ? launch_small_inst<C, SCALE_MUL, true>(ctx, P) : launch_small_inst<C, SCALE_MUL, false>(ctx, P);
    case SCALE_DIV: return full ? launch_small_inst<C, SCALE_DIV, true>(ctx, P) : launch_small_inst<C, SCALE_DIV, false>(ctx, P);
    case SCALE_ZERO_ONE: return full ? launch_small_inst<C, SCALE_ZERO_ONE, true>(ctx, P) : launch_small_inst<C, SCALE_ZERO_ONE, false>(ctx, P);
    default: return full ? launch_small_inst<C, SCALE_LOG10, true>(ctx, P) : launch_small_inst<C, SCALE_LOG10, false>(ctx, P);
  }
}

int launch_spectrum(sa_ctx *ctx, const SpectrumParams &p, int log2n) {
  if (p.n_frames == 0) return SA_OK;
#ifdef SA_ONLY_4096   // experiment builds (profiles/microbench): compile the headline size only
  if (log2n != 12) return SA_ERR_UNSUPPORTED_LENGTH;
  return launch_v2<V2Cfg<12, SA_G12>>(ctx, p);
#else
  switch (log2n) {
    case 1: case 2: case 3: case 4: case 5: {
      const unsigned grid = (unsigned)((p.n_frames + 127) / 128);
      spectrum_tiny_kernel<<<grid, 128, 0, ctx->stream>>>(p, log2n);
      ctx->launches++;
      snprintf(ctx->last_kernel, sizeof ctx->last_kernel, "sa::spectrum_tiny_kernel");
      SA_CUDA(cudaGetLastError());
      return SA_OK;
    }
    case 6: return (ctx->use_generic || !p.aligned8) ? launch_cfg<SpectrumCfg<6, 8, 32>>(ctx, p, log2n) : launch_small<SmallCfg<6>>(ctx, p);
    case 7: return (ctx->use_generic || !p.aligned8) ? launch_cfg<SpectrumCfg<7, 8, 16>>(ctx, p, log2n) : launch_small<SmallCfg<7>>(ctx, p);
    case 8: return (ctx->use_generic || !p.aligned8) ? launch_cfg<SpectrumCfg<8, 16, 16>>(ctx, p, log2n) : launch_small<SmallCfg<8>>(ctx, p);
    case 9: return (ctx->use_generic || !p.aligned8) ? launch_cfg<SpectrumCfg<9, 16, 8>>(ctx, p, log2n) : launch_small<SmallCfg<9>>(ctx, p);
    // 1024..16384 stage frames with bulk copies of the 16-byte aligned superset: any sample alignment
    case 10: return ctx->use_generic ? launch_cfg<SpectrumCfg<10, 16, 4>>(ctx, p, log2n) : launch_v2<V2Cfg<10, SA_G10>>(ctx, p);
    case 11: return ctx->use_generic ? launch_cfg<SpectrumCfg<11, 16, 2>>(ctx, p, log2n) : launch_v2<V2Cfg<11, SA_G11>>(ctx, p);
    case 12: return ctx->use_generic ? launch_cfg<SpectrumCfg<12, 16, 1>>(ctx, p, log2n) : launch_v2<V2Cfg<12, SA_G12>>(ctx, p);
    case 13: return ctx->use_generic ? launch_cfg<SpectrumCfg<13, 16, 1>>(ctx, p, log2n) : launch_v2<V2Cfg<13, SA_G13>>(ctx, p);
    case 14: return ctx->use_generic ? launch_cfg<SpectrumCfg<14, 16, 1>>(ctx, p, log2n) : launch_v2<V2Cfg<14, 1>>(ctx, p);
    case 15: return launch_cfg<SpectrumCfg<15, 16, 1>>(ctx, p, log2n);
    default: return SA_ERR_UNSUPPORTED_LENGTH;
  }
#endif
}

// Call-level validation in the reference's order (src/lib.rs:164-181, src/fft.rs:52, src/lib.rs:317).
int validate_call(uint64_t n, uint32_t sampling_rate, int limit_kind, float lo, float hi, int window_kind,
                  int scaling_kind, uint32_t *bin_lo, uint32_t *n_bins) {
  if (window_kind < SA_WINDOW_NONE || window_kind > SA_WINDOW_BLACKMAN_HARRIS_7TERM) return SA_ERR_INVALID_ARGUMENT;
  if (scaling_kind < SA_SCALING_NONE || scaling_kind > SA_SCALING_20_TIMES_LOG10) return SA_ERR_INVALID_ARGUMENT;
  if (limit_kind < SA_LIMIT_ALL || limit_kind > SA_LIMIT_RANGE) return SA_ERR_INVALID_ARGUMENT;
  if (n < 2) return SA_ERR_TOO_FEW_SAMPLES;
  if (n & (n - 1)) return SA_ERR_NOT_POWER_OF_TWO;
  const int e = sa_host::limit_verify(limit_kind, lo, hi, (float)sampling_rate / 2.0f);
  if (e != SA_OK) return e;
  if (n > 32768) return SA_ERR_UNSUPPORTED_LENGTH;
  sa_host::limit_bins((uint32_t)n, sampling_rate, limit_kind, lo, hi, bin_lo, n_bins);
  if (*n_bins < 2) return SA_ERR_LIMIT_TOO_NARROW;
  return SA_OK;
}

int fill_params(sa_ctx *ctx, SpectrumParams &p, const void *d_samples, int in_i16, uint64_t n_frames, uint32_t n,
                uint64_t frame_stride, int window_kind, int scaling_kind, uint32_t stats_mask, float *d_out,
                uint64_t out_stride, sa_frame_stats *d_stats, uint32_t bin_lo, uint32_t n_bins) {
  int e;
  if ((e = get_twiddle(ctx, n, &p.twiddle)) != SA_OK) return e;
  if ((e = get_window(ctx, window_kind, n, &p.window)) != SA_OK) return e;
  p.samples = static_cast<const float *>(d_samples);
  p.in_i16 = in_i16;
  p.fast_log10 = ctx->fast_log10 ? 1 : 0;
  p.out_vals = d_out;
  p.stats = (stats_mask & SA_STATS_ALL) ? d_stats : nullptr;
  p.n_frames = n_frames;
  p.frame_stride = frame_stride;
  p.out_stride = out_stride;
  p.bin_lo = bin_lo;
  p.bin_hi = bin_lo + n_bins - 1;
  p.scaling_kind = scaling_kind;
  p.stats_mask = stats_mask & SA_STATS_ALL;
  const float nf = (float)n;  // stats.n = samples_len as f32 (src/spectrum.rs:144)
  p.scale_div = scaling_kind == SA_SCALING_DIVIDE_BY_N_SQRT ? sqrtf(nf) : nf;
  const uintptr_t pair_mask = in_i16 ? 3u : 7u;   // a (2n, 2n+1) sample pair must be naturally aligned
  p.aligned8 = ((reinterpret_cast<uintptr_t>(d_samples) & pair_mask) == 0 && (frame_stride & 1u) == 0) ? 1 : 0;
  p.aligned16 = ((reinterpret_cast<uintptr_t>(d_samples) & 15u) == 0 && (frame_stride * (in_i16 ? 2u : 4u)) % 16u == 0) ? 1 : 0;
  return SA_OK;
}

int ensure_pipeline(sa_ctx *ctx, size_t in_bytes, size_t out_bytes, size_t stats_bytes) {
  if (!ctx->pipeline_ready) {
    SA_CUDA(cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
    SA_CUDA(cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
    for (int s = 0; s < sa_ctx::kSlots; s++) {
      SA_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[s], cudaEventDisableTiming));
      SA_CUDA(cudaEventCreateWithFlags(&ctx->ev_comp[s], cudaEventDisableTiming));
      SA_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[s], cudaEventDisableTiming));
    }
    ctx->pipeline_ready = true;
  }
  for (int s = 0; s < sa_ctx::kSlots; s++) {
    if (ctx->slot_in_bytes < in_bytes) {
      if (ctx->slot_in[s]) SA_CUDA(cudaFree(ctx->slot_in[s]));
      ctx->slot_in[s] = nullptr;
      SA_CUDA(cudaMalloc(&ctx->slot_in[s], in_bytes));
    }
    if (ctx->slot_out_bytes < out_bytes) {
      if (ctx->slot_out[s]) SA_CUDA(cudaFree(ctx->slot_out[s]));
      ctx->slot_out[s] = nullptr;
      SA_CUDA(cudaMalloc(&ctx->slot_out[s], out_bytes));
    }
    if (ctx->slot_stats_bytes < stats_bytes) {
      if (ctx->slot_stats[s]) SA_CUDA(cudaFree(ctx->slot_stats[s]));
      ctx->slot_stats[s] = nullptr;
      SA_CUDA(cudaMalloc(&ctx->slot_stats[s], stats_bytes));
    }
  }
  ctx->slot_in_bytes = std::max(ctx->slot_in_bytes, in_bytes);
  ctx->slot_out_bytes = std::max(ctx->slot_out_bytes, out_bytes);
  ctx->slot_stats_bytes = std::max(ctx->slot_stats_bytes, stats_bytes);
  return SA_OK;
}

// memcpy split over a few host threads (a single thread moves ~10 GB/s, PCIe Gen5 x16 ~55 GB/s)
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  static const unsigned nthreads = [] {
    const char *e = getenv("SA_HOST_COPY_THREADS");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return e ? (unsigned)std::max(1, atoi(e)) : std::min(8u, std::max(1u, hw / 2));
  }();
  if (bytes < (8u << 20) || nthreads == 1) { std::memcpy(dst, src, bytes); return; }
  const size_t part = ((bytes / nthreads) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nthreads; t++) {
    const size_t off = (size_t)t * part;
    if (off >= bytes) break;
    th.emplace_back([=] { std::memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, std::min(part, bytes - off)); });
  }
  std::memcpy(dst, src, std::min(part, bytes));
  for (auto &t : th) t.join();
}

bool is_cuda_host_memory(const void *p) {
  cudaPointerAttributes a;
  const bool yes = cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type != cudaMemoryTypeUnregistered;
  (void)cudaGetLastError();
  return yes;
}

int ensure_staging(sa_ctx *ctx, bool in, size_t in_bytes, bool out, size_t out_bytes, bool st, size_t st_bytes) {
  auto grow = [&](unsigned char *(&blk)[sa_ctx::kSlots], size_t &have, size_t need) -> int {
    if (have >= need) return SA_OK;
    for (int s = 0; s < sa_ctx::kSlots; s++) {
      if (blk[s]) SA_CUDA(cudaFreeHost(blk[s]));
      blk[s] = nullptr;
      SA_CUDA(cudaMallocHost(&blk[s], need));
    }
    have = need;
    return SA_OK;
  };
  int e;
  if (in && (e = grow(ctx->pin_in, ctx->pin_in_bytes, in_bytes)) != SA_OK) return e;
  if (out && (e = grow(ctx->pin_out, ctx->pin_out_bytes, out_bytes)) != SA_OK) return e;
  if (st && (e = grow(ctx->pin_st, ctx->pin_st_bytes, st_bytes)) != SA_OK) return e;
  return SA_OK;
}

// Page-locks a caller buffer for the duration of a call unless it already is CUDA host memory.
struct ScopedPin {
  void *ptr = nullptr;
  bool registered = false;
  void pin(const void *p, size_t bytes) {
    if (!p || bytes < (1u << 20)) return;  // small copies: pageable is fine
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) return;
    (void)cudaGetLastError();
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) {
      ptr = const_cast<void *>(p);
      registered = true;
    } else {
      (void)cudaGetLastError();  // fall back to pageable copies (slower, still correct)
    }
  }
  ~ScopedPin() { if (registered) cudaHostUnregister(ptr); }
};

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

uint32_t sa_abi_version(void) { return SA_ABI_VERSION; }

const char *sa_status_string(int status) {
  switch (status) {
    case SA_OK: return "ok";
    case SA_ERR_TOO_FEW_SAMPLES: return "Too few samples!";
    case SA_ERR_NAN_VALUES: return "NaN values are not supported!";
    case SA_ERR_INFINITY_VALUES: return "Infinity values are not supported!";
    case SA_ERR_NOT_POWER_OF_TWO: return "Samples length must be a power of two!";
    case SA_ERR_LIMIT_BELOW_MINIMUM: return "Invalid frequency limit: Value below minimum";
    case SA_ERR_LIMIT_ABOVE_NYQUIST: return "Invalid frequency limit: Value above Nyquist";
    case SA_ERR_LIMIT_INVALID_RANGE: return "Invalid frequency limit: Invalid range";
    case SA_ERR_LIMIT_TOO_NARROW: return "Frequency limit leaves too few frequency bins!";
    case SA_ERR_SCALING: return "Scaling error";
    case SA_ERR_UNSUPPORTED_LENGTH: return "should be one of the supported buffer lengths (2..32768)";
    case SA_ERR_NON_FINITE_RESULT: return "NaN/Infinite-values are not supported! (non-finite spectrum value)";
    case SA_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SA_ERR_CUDA: return "CUDA error (see sa_last_error_string)";
    case SA_ERR_OUT_OF_MEMORY: return "out of memory";
    default: return "unknown status";
  }
}

const char *sa_last_error_string(void) { return g_last_error.c_str(); }

int sa_ctx_create_on_stream(int device, void *cuda_stream, sa_ctx **out_ctx) {
  if (!out_ctx) return SA_ERR_INVALID_ARGUMENT;
  *out_ctx = nullptr;
  int count = 0;
  SA_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) { g_last_error = "no such CUDA device"; return SA_ERR_CUDA; }
  SA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    g_last_error = std::string("device is not sm_100 class: ") + prop.name;
    return SA_ERR_CUDA;
  }
  sa_ctx *ctx = new sa_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  const char *fg = getenv("SA_FORCE_GENERIC");
  ctx->use_generic = fg && fg[0] == '1';
  const char *hp = getenv("SA_HOST_PIN");
  ctx->host_register = hp && std::string(hp) == "register";
  const char *ns = getenv("SA_SPLIT");
  ctx->split = ns && ns[0] == '1';
  if (const char *st = getenv("SA_STAGGER_NS")) ctx->stagger_ns = (unsigned)strtoul(st, nullptr, 10);
  *out_ctx = ctx;
  return SA_OK;
}

int sa_ctx_create(int device, sa_ctx **out_ctx) {
  int e = sa_ctx_create_on_stream(device, nullptr, out_ctx);
  if (e != SA_OK) return e;
  cudaStream_t s;
  cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete *out_ctx; *out_ctx = nullptr; return cuda_fail(ce, "cudaStreamCreateWithFlags"); }
  (*out_ctx)->stream = s;
  (*out_ctx)->owns_stream = true;
  return SA_OK;
}

int sa_ctx_destroy(sa_ctx *ctx) {
  if (!ctx) return SA_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->twiddles) cudaFree(kv.second);
  for (auto &kv : ctx->windows) cudaFree(kv.second);
  for (auto &kv : ctx->v2_tables) cudaFree(kv.second);
  for (auto &kv : ctx->small_tables) cudaFree(kv.second);
  for (int s = 0; s < sa_ctx::kSlots; s++) {
    cudaFree(ctx->slot_in[s]); cudaFree(ctx->slot_out[s]); cudaFree(ctx->slot_stats[s]);
    if (ctx->pipeline_ready) { cudaEventDestroy(ctx->ev_h2d[s]); cudaEventDestroy(ctx->ev_comp[s]); cudaEventDestroy(ctx->ev_d2h[s]); }
  }
  if (ctx->pipeline_ready) { cudaStreamDestroy(ctx->h2d); cudaStreamDestroy(ctx->d2h); }
  for (int s = 0; s < sa_ctx::kSlots; s++) {
    if (ctx->pin_in[s]) cudaFreeHost(ctx->pin_in[s]);
    if (ctx->pin_out[s]) cudaFreeHost(ctx->pin_out[s]);
    if (ctx->pin_st[s]) cudaFreeHost(ctx->pin_st[s]);
  }
  if (ctx->single_pin) cudaFreeHost(ctx->single_pin);
  if (ctx->single_dev) cudaFree(ctx->single_dev);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SA_OK;
}

int sa_ctx_set_fast_log10(sa_ctx *ctx, int enable) {
  if (!ctx) return SA_ERR_INVALID_ARGUMENT;
  ctx->fast_log10 = enable != 0;
  return SA_OK;
}

int sa_ctx_sync(sa_ctx *ctx) {
  if (!ctx) return SA_ERR_INVALID_ARGUMENT;
  SA_CUDA(cudaStreamSynchronize(ctx->stream));
  return SA_OK;
}

uint64_t sa_ctx_launch_count(const sa_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *sa_ctx_last_kernel(const sa_ctx *ctx) { return ctx ? ctx->last_kernel : ""; }

int sa_window_coeffs(int window_kind, uint32_t n, float *host_out) {
  if (!host_out || window_kind < SA_WINDOW_NONE || window_kind > SA_WINDOW_BLACKMAN_HARRIS_7TERM || n > 32768)
    return SA_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < n; i++) host_out[i] = sa_host::window_multiplier(window_kind, i, n);
  return SA_OK;
}

int sa_limit_to_bins(uint32_t n, uint32_t sampling_rate, int limit_kind, float limit_lo, float limit_hi,
                     uint32_t *bin_lo, uint32_t *n_bins) {
  uint32_t lo = 0, nb = 0;
  const int e = validate_call(n, sampling_rate, limit_kind, limit_lo, limit_hi, SA_WINDOW_NONE, SA_SCALING_NONE, &lo, &nb);
  if (bin_lo) *bin_lo = lo;
  if (n_bins) *n_bins = nb;
  return e;
}

int sa_bin_frequencies(uint32_t n, uint32_t sampling_rate, uint32_t bin_lo, uint32_t n_bins, float *host_out) {
  if (!host_out || n == 0) return SA_ERR_INVALID_ARGUMENT;
  const volatile float res = (float)sampling_rate / (float)n;  // src/lib.rs:356-358
  for (uint32_t i = 0; i < n_bins; i++) {
    volatile float f = (float)(bin_lo + i) * res;              // src/lib.rs:278
    host_out[i] = f;
  }
  return SA_OK;
}

static int spectrum_batched_impl(sa_ctx *ctx, const void *d_samples, int in_i16, uint64_t n_frames, uint32_t n,
                                 uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                                 float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask,
                                 float *d_out_vals, uint64_t out_stride, sa_frame_stats *d_stats, uint32_t *bin_lo,
                                 uint32_t *n_bins) {
  if (!ctx) return SA_ERR_INVALID_ARGUMENT;
  uint32_t lo = 0, nb = 0;
  const int e = validate_call(n, sampling_rate, limit_kind, limit_lo, limit_hi, window_kind, scaling_kind, &lo, &nb);
  if (bin_lo) *bin_lo = lo;
  if (n_bins) *n_bins = nb;
  if (e != SA_OK) return e;
  if (n_frames == 0) return SA_OK;
  if (!d_samples || !d_out_vals || out_stride < nb || frame_stride == 0) return SA_ERR_INVALID_ARGUMENT;
  if ((stats_mask & SA_STATS_ALL) && !d_stats) return SA_ERR_INVALID_ARGUMENT;
  // natural alignment of the element types (a misaligned pointer would fault inside the kernel)
  if (reinterpret_cast<uintptr_t>(d_samples) % (in_i16 ? alignof(int16_t) : alignof(float)) != 0 ||
      reinterpret_cast<uintptr_t>(d_out_vals) % alignof(float) != 0 ||
      ((stats_mask & SA_STATS_ALL) && reinterpret_cast<uintptr_t>(d_stats) % 16 != 0)) {
    g_last_error = "misaligned pointer: samples need their element alignment, values 4 bytes, records 16 bytes";
    return SA_ERR_INVALID_ARGUMENT;
  }
  SA_CUDA(cudaSetDevice(ctx->device));
  SpectrumParams p;
  int pe = fill_params(ctx, p, d_samples, in_i16, n_frames, n, frame_stride, window_kind, scaling_kind, stats_mask,
                       d_out_vals, out_stride, d_stats, lo, nb);
  if (pe != SA_OK) return pe;
  return launch_spectrum(ctx, p, log2_exact(n));
}

int sa_spectrum_batched(sa_ctx *ctx, const float *d_samples, uint64_t n_frames, uint32_t n, uint64_t frame_stride,
                        uint32_t sampling_rate, int limit_kind, float limit_lo, float limit_hi, int window_kind,
                        int scaling_kind, uint32_t stats_mask, float *d_out_vals, uint64_t out_stride,
                        sa_frame_stats *d_stats, uint32_t *bin_lo, uint32_t *n_bins) {
  return spectrum_batched_impl(ctx, d_samples, 0, n_frames, n, frame_stride, sampling_rate, limit_kind, limit_lo,
                               limit_hi, window_kind, scaling_kind, stats_mask, d_out_vals, out_stride, d_stats, bin_lo,
                               n_bins);
}

int sa_spectrum_batched_i16(sa_ctx *ctx, const int16_t *d_samples, uint64_t n_frames, uint32_t n,
                            uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                            float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask, float *d_out_vals,
                            uint64_t out_stride, sa_frame_stats *d_stats, uint32_t *bin_lo, uint32_t *n_bins) {
  return spectrum_batched_impl(ctx, d_samples, 1, n_frames, n, frame_stride, sampling_rate, limit_kind, limit_lo,
                               limit_hi, window_kind, scaling_kind, stats_mask, d_out_vals, out_stride, d_stats, bin_lo,
                               n_bins);
}

static int spectrum_batched_host_impl(sa_ctx *ctx, const void *h_samples_v, int in_i16, uint64_t n_frames, uint32_t n,
                                      uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                                      float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask,
                                      float *h_out_vals, uint64_t out_stride, sa_frame_stats *h_stats,
                                      uint32_t *bin_lo, uint32_t *n_bins) {
  const size_t esz = in_i16 ? sizeof(int16_t) : sizeof(float);
  const unsigned char *h_samples = static_cast<const unsigned char *>(h_samples_v);
  if (!ctx) return SA_ERR_INVALID_ARGUMENT;
  uint32_t lo = 0, nb = 0;
  const int e = validate_call(n, sampling_rate, limit_kind, limit_lo, limit_hi, window_kind, scaling_kind, &lo, &nb);
  if (bin_lo) *bin_lo = lo;
  if (n_bins) *n_bins = nb;
  if (e != SA_OK) return e;
  if (n_frames == 0) return SA_OK;
  if (!h_samples || !h_out_vals || out_stride < nb || frame_stride == 0) return SA_ERR_INVALID_ARGUMENT;
  const bool want_stats = (stats_mask & SA_STATS_ALL) != 0;
  if (want_stats && !h_stats) return SA_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(h_samples) % esz != 0 || reinterpret_cast<uintptr_t>(h_out_vals) % alignof(float) != 0 ||
      (want_stats && reinterpret_cast<uintptr_t>(h_stats) % alignof(uint32_t) != 0)) {
    g_last_error = "misaligned host pointer";
    return SA_ERR_INVALID_ARGUMENT;
  }
  SA_CUDA(cudaSetDevice(ctx->device));

  // chunking: ~64 MiB of unique input per chunk
  // input bytes per pipeline chunk (SA_PIPE_CHUNK_MB overrides for tuning)
  static const uint64_t target = [] { const char *e = getenv("SA_PIPE_CHUNK_MB"); return (uint64_t)(e ? std::max(1, atoi(e)) : 64) << 20; }();
  uint64_t chunk = std::max<uint64_t>(1, target / (frame_stride * esz));
  chunk = std::min(chunk, n_frames);
  const size_t in_bytes = ((chunk - 1) * frame_stride + n) * esz;
  const size_t out_bytes = chunk * out_stride * sizeof(float);
  const size_t st_bytes = chunk * sizeof(sa_frame_stats);
  int pe = ensure_pipeline(ctx, in_bytes, out_bytes, st_bytes);
  if (pe != SA_OK) return pe;

  // Declared BEFORE the pins: on every exit path (errors included) the three streams are drained first, so no copy
  // into or out of the caller's buffers is in flight when the buffers are unpinned and handed back.
  struct Drain {
    sa_ctx *c;
    ~Drain() { cudaStreamSynchronize(c->d2h); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->h2d); }
  };
  ScopedPin pin_in, pin_out, pin_st;
  Drain drain{ctx};   // destroyed first (reverse order of declaration)
  // Pageable caller memory goes through a pinned staging ring (host threads copy a chunk while the previous ones
  // are on the bus); memory that already is CUDA host memory is used in place.
  bool stage_in = false, stage_out = false, stage_st = false;
  if (ctx->host_register) {
    pin_in.pin(h_samples, ((n_frames - 1) * frame_stride + n) * esz);
    pin_out.pin(h_out_vals, ((n_frames - 1) * out_stride + nb) * sizeof(float));
    if (want_stats) pin_st.pin(h_stats, n_frames * sizeof(sa_frame_stats));
  } else {
    stage_in = in_bytes >= (1u << 20) && !is_cuda_host_memory(h_samples);
    stage_out = out_bytes >= (1u << 20) && !is_cuda_host_memory(h_out_vals);
    stage_st = want_stats && stage_out && !is_cuda_host_memory(h_stats);
    pe = ensure_staging(ctx, stage_in, in_bytes, stage_out, out_bytes, stage_st, st_bytes);
    if (pe != SA_OK) return pe;
  }
  // staged outputs of chunk number `cc` (slot cc % kSlots): wait for its D2H copies, then hand the rows to the caller
  auto deliver = [&](uint64_t cc) -> int {
    const int sl = (int)(cc % sa_ctx::kSlots);
    const uint64_t g0 = cc * chunk, gn = std::min(chunk, n_frames - g0);
    SA_CUDA(cudaEventSynchronize(ctx->ev_d2h[sl]));
    if (stage_out) parallel_memcpy(h_out_vals + g0 * out_stride, ctx->pin_out[sl], ((gn - 1) * out_stride + nb) * sizeof(float));
    if (stage_st) std::memcpy(h_stats + g0, ctx->pin_st[sl], gn * sizeof(sa_frame_stats));
    return SA_OK;
  };

  const int log2n = log2_exact(n);
  uint64_t c = 0;
  for (uint64_t f0 = 0; f0 < n_frames; f0 += chunk, c++) {
    const int s = (int)(c % sa_ctx::kSlots);
    const uint64_t nf = std::min(chunk, n_frames - f0);
    if (c >= (uint64_t)sa_ctx::kSlots) {
      if (stage_out && (pe = deliver(c - sa_ctx::kSlots)) != SA_OK) return pe;   // frees this slot's staged outputs
      SA_CUDA(cudaStreamWaitEvent(ctx->h2d, ctx->ev_comp[s], 0));     // input slot consumed
      SA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[s], 0));   // output slot drained
    }
    const unsigned char *src = h_samples + f0 * frame_stride * esz;
    const size_t src_bytes = ((nf - 1) * frame_stride + n) * esz;
    if (stage_in) {
      if (c >= (uint64_t)sa_ctx::kSlots) SA_CUDA(cudaEventSynchronize(ctx->ev_h2d[s]));   // the block's previous copy has left
      parallel_memcpy(ctx->pin_in[s], src, src_bytes);
      src = ctx->pin_in[s];
    }
    SA_CUDA(cudaMemcpyAsync(ctx->slot_in[s], src, src_bytes, cudaMemcpyHostToDevice, ctx->h2d));
    SA_CUDA(cudaEventRecord(ctx->ev_h2d[s], ctx->h2d));
    SA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[s], 0));
    SpectrumParams p;
    pe = fill_params(ctx, p, ctx->slot_in[s], in_i16, nf, n, frame_stride, window_kind, scaling_kind, stats_mask,
                     ctx->slot_out[s], out_stride, ctx->slot_stats[s], lo, nb);
    if (pe != SA_OK) return pe;
    pe = launch_spectrum(ctx, p, log2n);
    if (pe != SA_OK) return pe;
    SA_CUDA(cudaEventRecord(ctx->ev_comp[s], ctx->stream));
    SA_CUDA(cudaStreamWaitEvent(ctx->d2h, ctx->ev_comp[s], 0));
    SA_CUDA(cudaMemcpyAsync(stage_out ? reinterpret_cast<float *>(ctx->pin_out[s]) : h_out_vals + f0 * out_stride, ctx->slot_out[s],
                            ((nf - 1) * out_stride + nb) * sizeof(float), cudaMemcpyDeviceToHost, ctx->d2h));
    if (want_stats)
      SA_CUDA(cudaMemcpyAsync(stage_st ? reinterpret_cast<sa_frame_stats *>(ctx->pin_st[s]) : h_stats + f0, ctx->slot_stats[s],
                              nf * sizeof(sa_frame_stats), cudaMemcpyDeviceToHost, ctx->d2h));
    SA_CUDA(cudaEventRecord(ctx->ev_d2h[s], ctx->d2h));
  }
  if (stage_out)   // the last (up to kSlots) chunks are still in the staging blocks
    for (uint64_t cc = c > (uint64_t)sa_ctx::kSlots ? c - sa_ctx::kSlots : 0; cc < c; cc++)
      if ((pe = deliver(cc)) != SA_OK) return pe;
  SA_CUDA(cudaStreamSynchronize(ctx->d2h));
  SA_CUDA(cudaStreamSynchronize(ctx->stream));
  SA_CUDA(cudaStreamSynchronize(ctx->h2d));
  return SA_OK;
}

int sa_spectrum_batched_host(sa_ctx *ctx, const float *h_samples, uint64_t n_frames, uint32_t n,
                             uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                             float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask,
                             float *h_out_vals, uint64_t out_stride, sa_frame_stats *h_stats, uint32_t *bin_lo,
                             uint32_t *n_bins) {
  return spectrum_batched_host_impl(ctx, h_samples, 0, n_frames, n, frame_stride, sampling_rate, limit_kind, limit_lo,
                                    limit_hi, window_kind, scaling_kind, stats_mask, h_out_vals, out_stride, h_stats,
                                    bin_lo, n_bins);
}

int sa_spectrum_batched_host_i16(sa_ctx *ctx, const int16_t *h_samples, uint64_t n_frames, uint32_t n,
                                 uint64_t frame_stride, uint32_t sampling_rate, int limit_kind, float limit_lo,
                                 float limit_hi, int window_kind, int scaling_kind, uint32_t stats_mask,
                                 float *h_out_vals, uint64_t out_stride, sa_frame_stats *h_stats, uint32_t *bin_lo,
                                 uint32_t *n_bins) {
  return spectrum_batched_host_impl(ctx, h_samples, 1, n_frames, n, frame_stride, sampling_rate, limit_kind, limit_lo,
                                    limit_hi, window_kind, scaling_kind, stats_mask, h_out_vals, out_stride, h_stats,
                                    bin_lo, n_bins);
}

int sa_spectrum_single(sa_ctx *ctx, const float *h_samples, uint64_t n_samples, uint32_t sampling_rate,
                       int limit_kind, float limit_lo, float limit_hi, int window_kind, int scaling_kind,
                       float *h_freqs, float *h_vals, uint32_t *n_bins, sa_frame_stats *h_stats) {
  if (!ctx) return SA_ERR_INVALID_ARGUMENT;
  if (n_bins) *n_bins = 0;
  if (window_kind < SA_WINDOW_NONE || window_kind > SA_WINDOW_BLACKMAN_HARRIS_7TERM) return SA_ERR_INVALID_ARGUMENT;
  // src/lib.rs:164-176: length, then NaN, then Inf (of the samples handed to the FFT, i.e. after
  // the window), then power of two.  This scan is argument validation only.
  if (n_samples < 2) return SA_ERR_TOO_FEW_SAMPLES;
  if (!h_samples) return SA_ERR_INVALID_ARGUMENT;
  {
    // the multiplier only matters where a sample is infinite (w == 0 turns Inf into NaN): the window's cosf is
    // not evaluated for the finite samples of an ordinary call
    bool any_nan = false, any_inf = false;
    for (uint64_t i = 0; i < n_samples; i++) {
      const float v = h_samples[i];
      if (std::fabs(v) <= std::numeric_limits<float>::max()) continue;   // finite
      if (std::isnan(v)) { any_nan = true; continue; }
      const float w = window_kind == SA_WINDOW_NONE ? 1.0f : sa_host::window_multiplier(window_kind, (uint32_t)i, (uint32_t)n_samples);
      const float wv = w * v;
      any_nan |= std::isnan(wv);
      any_inf |= std::isinf(wv);
    }
    if (any_nan) return SA_ERR_NAN_VALUES;
    if (any_inf) return SA_ERR_INFINITY_VALUES;
  }
  uint32_t lo = 0, nb = 0;
  const int e = validate_call(n_samples, sampling_rate, limit_kind, limit_lo, limit_hi, window_kind, scaling_kind, &lo, &nb);
  if (e != SA_OK) return e;
  if (!h_vals) return SA_ERR_INVALID_ARGUMENT;
  const uint32_t n = (uint32_t)n_samples;
  // latency path: samples -> pinned block -> device, one launch, [values | record] back in one copy
  SA_CUDA(cudaSetDevice(ctx->device));
  const size_t in_bytes = ((size_t)n * sizeof(float) + 255) & ~(size_t)255;
  const size_t val_bytes = ((size_t)nb * sizeof(float) + 31) & ~(size_t)31;
  const size_t total = in_bytes + val_bytes + sizeof(sa_frame_stats);
  if (ctx->single_bytes < total) {
    if (ctx->single_pin) SA_CUDA(cudaFreeHost(ctx->single_pin));
    if (ctx->single_dev) SA_CUDA(cudaFree(ctx->single_dev));
    ctx->single_pin = ctx->single_dev = nullptr;
    ctx->single_bytes = 0;
    SA_CUDA(cudaMallocHost(&ctx->single_pin, total));
    SA_CUDA(cudaMalloc(&ctx->single_dev, total));
    ctx->single_bytes = total;
  }
  std::memcpy(ctx->single_pin, h_samples, (size_t)n * sizeof(float));
  SA_CUDA(cudaMemcpyAsync(ctx->single_dev, ctx->single_pin, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  float *d_vals = reinterpret_cast<float *>(ctx->single_dev + in_bytes);
  sa_frame_stats *d_st = reinterpret_cast<sa_frame_stats *>(ctx->single_dev + in_bytes + val_bytes);
  SpectrumParams p;
  int pe = fill_params(ctx, p, ctx->single_dev, 0, 1, n, n, window_kind, scaling_kind, SA_STATS_ALL, d_vals, nb, d_st, lo, nb);
  if (pe != SA_OK) return pe;
  pe = launch_spectrum(ctx, p, log2_exact(n));
  if (pe != SA_OK) return pe;
  SA_CUDA(cudaMemcpyAsync(ctx->single_pin + in_bytes, ctx->single_dev + in_bytes, val_bytes + sizeof(sa_frame_stats),
                          cudaMemcpyDeviceToHost, ctx->stream));
  SA_CUDA(cudaStreamSynchronize(ctx->stream));
  sa_frame_stats st;
  std::memcpy(&st, ctx->single_pin + in_bytes + val_bytes, sizeof st);
  if (st.status != SA_OK) return (int)st.status;
  std::memcpy(h_vals, ctx->single_pin + in_bytes, (size_t)nb * sizeof(float));
  if (h_freqs) sa_bin_frequencies(n, sampling_rate, lo, nb, h_freqs);
  if (n_bins) *n_bins = nb;
  if (h_stats) *h_stats = st;
  return SA_OK;
}

int sa_apply_scaling_chain_batched(sa_ctx *ctx, float *d_vals, uint64_t n_frames, uint32_t n_bins, uint64_t stride,
                                   uint32_t bin_lo, uint32_t n, const int *scaling_kinds, uint32_t n_kinds,
                                   uint32_t stats_mask, sa_frame_stats *d_stats) {
  if (!ctx || !d_vals || !d_stats || n_bins < 2 || stride < n_bins || n < 2) return SA_ERR_INVALID_ARGUMENT;
  if (n_kinds > 8 || (n_kinds && !scaling_kinds)) return SA_ERR_INVALID_ARGUMENT;
  sa::ScalingChain chain;
  chain.count = (int)n_kinds;
  const float nf = (float)n;
  for (uint32_t c = 0; c < 8; c++) {
    const int k = c < n_kinds ? scaling_kinds[c] : SA_SCALING_NONE;
    if (k < SA_SCALING_NONE || k > SA_SCALING_20_TIMES_LOG10) return SA_ERR_INVALID_ARGUMENT;
    chain.kind[c] = k;
    chain.div[c] = k == SA_SCALING_DIVIDE_BY_N_SQRT ? sqrtf(nf) : nf;
  }
  if (n_frames == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)std::min<uint64_t>(n_frames, (uint64_t)ctx->num_sms * 8);
  sa::rescale_kernel<<<grid, sa::RESCALE_THREADS, 0, ctx->stream>>>(d_vals, n_frames, n_bins, stride, bin_lo, chain,
                                                                    stats_mask & SA_STATS_ALL, d_stats);
  ctx->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

int sa_apply_scaling_batched(sa_ctx *ctx, float *d_vals, uint64_t n_frames, uint32_t n_bins, uint64_t stride,
                             uint32_t bin_lo, uint32_t n, int scaling_kind, uint32_t stats_mask,
                             sa_frame_stats *d_stats) {
  return sa_apply_scaling_chain_batched(ctx, d_vals, n_frames, n_bins, stride, bin_lo, n, &scaling_kind, 1, stats_mask,
                                        d_stats);
}

int sa_apply_scaling_chain_host(sa_ctx *ctx, float *h_vals, uint64_t n_frames, uint32_t n_bins, uint64_t stride,
                                uint32_t bin_lo, uint32_t n, const int *scaling_kinds, uint32_t n_kinds,
                                uint32_t stats_mask, sa_frame_stats *h_stats) {
  if (!ctx || !h_vals || !h_stats || n_bins < 2 || stride < n_bins || n < 2) return SA_ERR_INVALID_ARGUMENT;
  if (n_frames == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(ctx->device));
  const uint64_t cf = std::min<uint64_t>(n_frames, std::max<uint64_t>(1, (64ull << 20) / (stride * sizeof(float))));
  int pe = ensure_pipeline(ctx, cf * stride * sizeof(float), 16, cf * sizeof(sa_frame_stats));
  if (pe != SA_OK) return pe;
  for (uint64_t f0 = 0; f0 < n_frames; f0 += cf) {   // simple serial chunks: this is not the hot entry point
    const uint64_t nf = std::min(cf, n_frames - f0);
    float *d_vals = static_cast<float *>(ctx->slot_in[0]);
    SA_CUDA(cudaMemcpyAsync(d_vals, h_vals + f0 * stride, ((nf - 1) * stride + n_bins) * sizeof(float),
                            cudaMemcpyHostToDevice, ctx->stream));
    SA_CUDA(cudaMemcpyAsync(ctx->slot_stats[0], h_stats + f0, nf * sizeof(sa_frame_stats), cudaMemcpyHostToDevice,
                            ctx->stream));
    pe = sa_apply_scaling_chain_batched(ctx, d_vals, nf, n_bins, stride, bin_lo, n, scaling_kinds, n_kinds, stats_mask,
                                        ctx->slot_stats[0]);
    if (pe != SA_OK) return pe;
    SA_CUDA(cudaMemcpyAsync(h_vals + f0 * stride, d_vals, ((nf - 1) * stride + n_bins) * sizeof(float),
                            cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(cudaMemcpyAsync(h_stats + f0, ctx->slot_stats[0], nf * sizeof(sa_frame_stats), cudaMemcpyDeviceToHost,
                            ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SA_OK;
}

int sa_window_batched(sa_ctx *ctx, int window_kind, const float *d_in, uint64_t n_frames, uint32_t n,
                      uint64_t frame_stride, float *d_out) {
  if (!ctx || !d_in || !d_out || n == 0 || n > 32768) return SA_ERR_INVALID_ARGUMENT;
  if (window_kind < SA_WINDOW_NONE || window_kind > SA_WINDOW_BLACKMAN_HARRIS_7TERM) return SA_ERR_INVALID_ARGUMENT;
  if (n_frames == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(ctx->device));
  const float *w = nullptr;
  int e = get_window(ctx, window_kind, n, &w);
  if (e != SA_OK) return e;
  const uint64_t total = n_frames * n;
  const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->num_sms * 16);
  sa::window_kernel<<<grid, 256, 0, ctx->stream>>>(d_in, w, d_out, n_frames, n, frame_stride);
  ctx->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

int sa_window_host(sa_ctx *ctx, int window_kind, const float *h_in, uint64_t n_frames, uint32_t n,
                   uint64_t frame_stride, float *h_out) {
  if (!ctx || !h_in || !h_out || n == 0 || n > 32768 || frame_stride == 0) return SA_ERR_INVALID_ARGUMENT;
  if (window_kind < SA_WINDOW_NONE || window_kind > SA_WINDOW_BLACKMAN_HARRIS_7TERM) return SA_ERR_INVALID_ARGUMENT;
  if (n_frames == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(ctx->device));
  const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / (std::max<uint64_t>(frame_stride, n) * sizeof(float)));
  const uint64_t cf = std::min(chunk, n_frames);
  int pe = ensure_pipeline(ctx, ((cf - 1) * frame_stride + n) * sizeof(float), cf * n * sizeof(float), 32);
  if (pe != SA_OK) return pe;
  for (uint64_t f0 = 0; f0 < n_frames; f0 += cf) {   // simple serial chunks: this is not the hot entry point
    const uint64_t nf = std::min(cf, n_frames - f0);
    SA_CUDA(cudaMemcpyAsync(ctx->slot_in[0], h_in + f0 * frame_stride, ((nf - 1) * frame_stride + n) * sizeof(float),
                            cudaMemcpyHostToDevice, ctx->stream));
    pe = sa_window_batched(ctx, window_kind, ctx->slot_in[0], nf, n, frame_stride, ctx->slot_out[0]);
    if (pe != SA_OK) return pe;
    SA_CUDA(cudaMemcpyAsync(h_out + f0 * n, ctx->slot_out[0], nf * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SA_OK;
}

int sa_generate_noise(sa_ctx *ctx, float *d_out, uint64_t first_index, uint64_t count, uint64_t seed) {
  if (!ctx || (!d_out && count)) return SA_ERR_INVALID_ARGUMENT;
  if (count == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)std::min<uint64_t>((count + 255) / 256, (uint64_t)ctx->num_sms * 16);
  sa::noise_kernel<<<grid, 256, 0, ctx->stream>>>(d_out, first_index, count, seed);
  ctx->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

}  // extern "C"
