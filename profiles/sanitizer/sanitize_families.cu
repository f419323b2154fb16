// Small runs of every kernel family through the C ABI, for compute-sanitizer (memcheck / racecheck / synccheck /
// initcheck: one tool per run).  Buffers are cudaMalloc'ed with their EXACT extents, so a read past the last sample
// or a write past the last row is an error the tools see; hops and offsets that are not multiples of 16 bytes
// exercise the clamped staging copies.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include profiles/sanitizer/sanitize_families.cu \
//        -L spectrum_analyzer_b200/lib -lspectrum_analyzer_cuda -o profiles/sanitizer/sanitize_families
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "spectrum_analyzer_cuda.h"

#define CK(x) do { int e_ = (x); if (e_ != 0) { printf("FAIL %s -> %d (%s)\n", #x, e_, sa_last_error_string()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static int run_case(sa_ctx *ctx, uint32_t n, uint64_t frames, uint64_t hop, uint64_t offset, int limit_kind, float lo, float hi,
                    int window, int scaling, uint32_t mask, bool i16) {
  const uint64_t samples = (frames - 1) * hop + n;
  float *d_all = nullptr;
  // the frames start `offset` elements into an allocation that ends exactly at the last sample
  CU(cudaMalloc(&d_all, (offset + samples) * (i16 ? 2 : 4)));
  if (i16) {
    std::vector<int16_t> h(offset + samples);
    for (size_t i = 0; i < h.size(); i++) h[i] = (int16_t)((i * 2654435761u) >> 16);
    CU(cudaMemcpy(d_all, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  } else {
    CK(sa_generate_noise(ctx, d_all, 0, offset + samples, 1234 + n));
  }
  uint32_t bin_lo = 0, n_bins = 0;
  CK(sa_limit_to_bins(n, 44100, limit_kind, lo, hi, &bin_lo, &n_bins));
  float *d_vals = nullptr;
  sa_frame_stats *d_stats = nullptr;
  CU(cudaMalloc(&d_vals, frames * n_bins * sizeof(float)));
  CU(cudaMalloc(&d_stats, frames * sizeof(sa_frame_stats)));
  if (i16)
    CK(sa_spectrum_batched_i16(ctx, reinterpret_cast<const int16_t *>(d_all) + offset, frames, n, hop, 44100, limit_kind, lo, hi, window,
                               scaling, mask, d_vals, n_bins, mask ? d_stats : nullptr, nullptr, nullptr));
  else
    CK(sa_spectrum_batched(ctx, d_all + offset, frames, n, hop, 44100, limit_kind, lo, hi, window, scaling, mask, d_vals, n_bins,
                           mask ? d_stats : nullptr, nullptr, nullptr));
  CK(sa_ctx_sync(ctx));
  if (mask == SA_STATS_ALL) {   // rescale kernel on the resident batch (a chain, then dB twice: ScalingError frames)
    const int chain[3] = {SA_SCALING_DIVIDE_BY_N_SQRT, SA_SCALING_DIVIDE_BY_N_SQRT, SA_SCALING_DIVIDE_BY_N_SQRT};
    CK(sa_apply_scaling_chain_batched(ctx, d_vals, frames, n_bins, n_bins, bin_lo, n, chain, 3, SA_STATS_ALL, d_stats));
    CK(sa_apply_scaling_batched(ctx, d_vals, frames, n_bins, n_bins, bin_lo, n, SA_SCALING_20_TIMES_LOG10, SA_STATS_ALL, d_stats));
    CK(sa_apply_scaling_batched(ctx, d_vals, frames, n_bins, n_bins, bin_lo, n, SA_SCALING_20_TIMES_LOG10, SA_STATS_ALL, d_stats));
    CK(sa_ctx_sync(ctx));
  }
  std::vector<sa_frame_stats> st(frames);
  CU(cudaMemcpy(st.data(), d_stats, frames * sizeof(sa_frame_stats), cudaMemcpyDeviceToHost));
  CU(cudaFree(d_all)); CU(cudaFree(d_vals)); CU(cudaFree(d_stats));
  printf("ok n=%u frames=%llu hop=%llu off=%llu limit=%d win=%d scale=%d mask=%u %s\n", n, (unsigned long long)frames,
         (unsigned long long)hop, (unsigned long long)offset, limit_kind, window, scaling, mask, i16 ? "i16" : "f32");
  return 0;
}

int main(int argc, char **argv) {
  const uint64_t frames = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100;
  for (int split = 0; split < 2; split++) {
    if (split) setenv("SA_SPLIT", "1", 1);
    sa_ctx *ctx = nullptr;
    CK(sa_ctx_create(0, &ctx));
    const uint32_t sizes[] = {16, 64, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    for (uint32_t n : sizes) {
      if (split && n != 4096) continue;                        // the split variant exists for 4096 only
      const uint64_t f = n >= 16384 ? (frames + 3) / 4 : frames;
      if (run_case(ctx, n, f, n, 0, SA_LIMIT_ALL, 0, 0, SA_WINDOW_HANN, SA_SCALING_DIVIDE_BY_N_SQRT, SA_STATS_ALL, false)) return 1;
      if (run_case(ctx, n, f, n / 4 + 1, 1, SA_LIMIT_ALL, 0, 0, SA_WINDOW_HAMMING, SA_SCALING_20_TIMES_LOG10, SA_STATS_ALL, false)) return 1;
      if (run_case(ctx, n, f, n / 2 + 3, 3, SA_LIMIT_RANGE, 1500.f, 15000.f, SA_WINDOW_BLACKMAN_HARRIS_4TERM, SA_SCALING_ZERO_TO_ONE,
                   SA_STATS_MINMAX, true)) return 1;
      if (run_case(ctx, n, 1, n, 0, SA_LIMIT_MAX, 0.f, 9000.f, SA_WINDOW_NONE, SA_SCALING_NONE, 0, false)) return 1;
    }
    // stand-alone window kernel
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, (99 * 700 + 1024) * sizeof(float)));
    CU(cudaMalloc(&d_out, 100 * 1024 * sizeof(float)));
    CK(sa_generate_noise(ctx, d_in, 0, 99 * 700 + 1024, 7));
    CK(sa_window_batched(ctx, SA_WINDOW_BLACKMAN_HARRIS_7TERM, d_in, 100, 1024, 700, d_out));
    CK(sa_ctx_sync(ctx));
    CU(cudaFree(d_in)); CU(cudaFree(d_out));
    CK(sa_ctx_destroy(ctx));
  }
  printf("all cases done\n");
  return 0;
}
