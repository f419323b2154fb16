// spectrum_analyzer.hpp -- C++17 host-side mirror of the reference crate's public API for the hot
// path, written above the C ABI (include/spectrum_analyzer_cuda.h).  Header-only.
//
// The reference is a Rust crate; this image has no Rust toolchain, so the host layer a Rust caller
// would get (rust/spectrum-analyzer-cuda, source only) is mirrored here in C++ with the same names,
// argument meaning and error behaviour:
//
//   spectrum_analyzer::samples_fft_to_spectrum(samples, sampling_rate, FrequencyLimit, scaling)
//       reference src/lib.rs:157-162            -> Result<FrequencySpectrum>
//   spectrum_analyzer::samples_fft_to_spectrum_batched(...)      the new [frames x N] entry point
//   spectrum_analyzer::windows::{hann_window, hamming_window, blackman_harris_4term, blackman_harris_7term}
//       reference src/windows.rs:40,58,79,97
//   spectrum_analyzer::scaling::{divide_by_N, divide_by_N_sqrt, scale_to_zero_to_one, scale_20_times_log10}
//       reference src/scaling.rs:104-162 (tokens: arbitrary closures are not representable, by design)
//   spectrum_analyzer::FrequencyLimit::{All, Min, Max, Range}     reference src/limit.rs:38-53
//   spectrum_analyzer::FrequencySpectrum getters, lookups (freq_val_exact / freq_val_closest / mel_val /
//       to_map / to_mel_map) and apply_scaling_fn                reference src/spectrum.rs:128-474
//   spectrum_analyzer::scaling::combined                          reference src/scaling.rs:174-182
//   spectrum_analyzer::SpectrumAnalyzerError                      reference src/error.rs:37-54
//
// Everything numeric happens in libspectrum_analyzer_cuda.so; there is no CPU fallback.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

#include "../../include/spectrum_analyzer_cuda.h"

namespace spectrum_analyzer {

// ---- errors (src/error.rs:37-54) ---------------------------------------------------------------
enum class SpectrumAnalyzerError : int {
  TooFewSamples = SA_ERR_TOO_FEW_SAMPLES,
  NaNValuesNotSupported = SA_ERR_NAN_VALUES,
  InfinityValuesNotSupported = SA_ERR_INFINITY_VALUES,
  SamplesLengthNotAPowerOfTwo = SA_ERR_NOT_POWER_OF_TWO,
  InvalidFrequencyLimitValueBelowMinimum = SA_ERR_LIMIT_BELOW_MINIMUM,
  InvalidFrequencyLimitValueAboveNyquist = SA_ERR_LIMIT_ABOVE_NYQUIST,
  InvalidFrequencyLimitInvalidRange = SA_ERR_LIMIT_INVALID_RANGE,
  FrequencyLimitTooNarrow = SA_ERR_LIMIT_TOO_NARROW,
  ScalingError = SA_ERR_SCALING,
  UnsupportedLength = SA_ERR_UNSUPPORTED_LENGTH,   // the reference panics (src/fft.rs:52)
  NonFiniteResult = SA_ERR_NON_FINITE_RESULT,      // the reference panics (src/frequency.rs:52-53)
  InvalidArgument = SA_ERR_INVALID_ARGUMENT,
  Cuda = SA_ERR_CUDA,
};
inline std::string to_string(SpectrumAnalyzerError e) { return sa_status_string(static_cast<int>(e)); }

// Result<T, SpectrumAnalyzerError>
template <class T>
class Result {
 public:
  Result(T v) : v_(std::move(v)) {}
  Result(SpectrumAnalyzerError e) : v_(e) {}
  bool is_ok() const { return v_.index() == 0; }
  bool is_err() const { return !is_ok(); }
  T &unwrap() {
    if (is_err()) throw std::runtime_error("called unwrap() on Err: " + to_string(std::get<1>(v_)));
    return std::get<0>(v_);
  }
  SpectrumAnalyzerError unwrap_err() const {
    if (is_ok()) throw std::runtime_error("called unwrap_err() on Ok");
    return std::get<1>(v_);
  }
 private:
  std::variant<T, SpectrumAnalyzerError> v_;
};

// ---- FrequencyLimit (src/limit.rs:38-118) ------------------------------------------------------
struct FrequencyLimit {
  int kind = SA_LIMIT_ALL;
  float lo = 0.f, hi = 0.f;
  static FrequencyLimit All() { return {}; }
  static FrequencyLimit Min(float x) { return {SA_LIMIT_MIN, x, 0.f}; }
  static FrequencyLimit Max(float x) { return {SA_LIMIT_MAX, 0.f, x}; }
  static FrequencyLimit Range(float a, float b) { return {SA_LIMIT_RANGE, a, b}; }
};

// ---- scaling (src/scaling.rs:104-162) ----------------------------------------------------------
namespace scaling {
struct Fn { int kind; };
inline constexpr Fn divide_by_N{SA_SCALING_DIVIDE_BY_N};
inline constexpr Fn divide_by_N_sqrt{SA_SCALING_DIVIDE_BY_N_SQRT};
inline constexpr Fn scale_to_zero_to_one{SA_SCALING_ZERO_TO_ONE};
inline constexpr Fn scale_20_times_log10{SA_SCALING_20_TIMES_LOG10};
// scaling::combined(&[...]) (src/scaling.rs:174-182): every function of the chain sees the statistics of the
// spectrum before the call.  At most 8 built-in functions.
struct Chain { std::vector<int> kinds; };
inline Chain combined(std::initializer_list<Fn> fns) {
  Chain c;
  for (const Fn &f : fns) c.kinds.push_back(f.kind);
  return c;
}
}  // namespace scaling

// ---- context -----------------------------------------------------------------------------------
class Context {
 public:
  explicit Context(int device = 0) {
    if (int e = sa_ctx_create(device, &h_); e != SA_OK) throw std::runtime_error(std::string("sa_ctx_create: ") + sa_last_error_string());
  }
  Context(int device, void *cuda_stream) {
    if (int e = sa_ctx_create_on_stream(device, cuda_stream, &h_); e != SA_OK) throw std::runtime_error(std::string("sa_ctx_create_on_stream: ") + sa_last_error_string());
  }
  ~Context() { sa_ctx_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  sa_ctx *handle() const { return h_; }
  void sync() const { sa_ctx_sync(h_); }
 private:
  sa_ctx *h_ = nullptr;
};

// ---- FrequencySpectrum (src/spectrum.rs:48-253) --------------------------------------------------
class FrequencySpectrum {
 public:
  using Pair = std::pair<float, float>;  // (Frequency, FrequencyValue)
  FrequencySpectrum(std::vector<Pair> data, float resolution, uint32_t samples_len, const sa_frame_stats &st,
                    uint32_t bin_lo)
      : data_(std::move(data)), res_(resolution), n_(samples_len), st_(st), bin_lo_(bin_lo) {}
  float average() const { return st_.average; }
  float median() const { return st_.median; }
  Pair max() const { return {data_[st_.max_bin - bin_lo_].first, st_.max_val}; }
  Pair min() const { return {data_[st_.min_bin - bin_lo_].first, st_.min_val}; }
  float range() const { return st_.max_val - st_.min_val; }
  const std::vector<Pair> &data() const { return data_; }
  float frequency_resolution() const { return res_; }
  uint32_t samples_len() const { return n_; }
  float max_fr() const { return data_.back().first; }
  float min_fr() const { return data_.front().first; }
  bool dc_component(float *out) const {
    if (data_.front().first != 0.0f) return false;
    *out = data_.front().second;
    return true;
  }

  // ---- lookups (src/spectrum.rs:300-474).  Out-of-range frequencies throw std::out_of_range where the
  // reference panics.  The reference walks the bin pairs linearly; the bins are sorted, so the first pair
  // whose upper frequency is >= search_fr is found by bisection here.
  float freq_val_exact(float search_fr) const {
    if (approx_eq_ulps3(data_.front().first, search_fr)) return data_.front().second;
    if (approx_eq_ulps3(data_.back().first, search_fr)) return data_.back().second;
    const size_t b = upper_pair(search_fr);
    const Pair &pa = data_[b - 1], &pb = data_[b];
    if (approx_eq_ulps3(pa.first, search_fr)) return pa.second;
    // line through A and B evaluated at search_fr, f32 operation by operation (src/spectrum.rs:584-588)
    const float slope = (pb.second - pa.second) / (pb.first - pa.first);
    const float c = pa.second - slope * pa.first;
    return slope * search_fr + c;
  }
  Pair freq_val_closest(float search_fr) const {
    if (approx_eq_ulps3(data_.front().first, search_fr)) return data_.front();
    if (approx_eq_ulps3(data_.back().first, search_fr)) return data_.back();
    const size_t b = upper_pair(search_fr);
    const Pair &pa = data_[b - 1], &pb = data_[b];
    if (approx_eq_ulps3(pa.first, search_fr)) return pa;
    return ((search_fr - pa.first) / res_ < 0.5f) ? pa : pb;
  }
  // mel = 2595 log10(1 + hz/700) (src/spectrum.rs:557-601)
  static float mel_to_hertz(float mel) { return 700.0f * (std::pow(10.0f, mel / 2595.0f) - 1.0f); }
  static float hertz_to_mel(float hz) { return 2595.0f * std::log10(1.0f + hz / 700.0f); }
  float mel_val(float mel) const { return freq_val_exact(mel_to_hertz(mel)); }
  std::map<uint32_t, float> to_map() const {
    std::map<uint32_t, float> m;
    for (const Pair &p : data_) m[static_cast<uint32_t>(p.first)] = p.second;   // later bins overwrite, as BTreeMap::collect does
    return m;
  }
  std::map<uint32_t, float> to_mel_map() const {
    std::map<uint32_t, float> m;
    for (const Pair &p : data_) m[static_cast<uint32_t>(hertz_to_mel(p.first))] = p.second;
    return m;
  }

  // ---- FrequencySpectrum::apply_scaling_fn (src/spectrum.rs:128-168), computed on the device: values are
  // rescaled with the current statistics, then the statistics are recalculated.
  Result<bool> apply_scaling_fn(Context &ctx, const scaling::Chain &chain) {
    std::vector<float> vals(data_.size());
    for (size_t i = 0; i < data_.size(); i++) vals[i] = data_[i].second;
    const int e = sa_apply_scaling_chain_host(ctx.handle(), vals.data(), 1, static_cast<uint32_t>(vals.size()), vals.size(),
                                              bin_lo_, n_, chain.kinds.data(), static_cast<uint32_t>(chain.kinds.size()),
                                              SA_STATS_ALL, &st_);
    if (e != SA_OK) return static_cast<SpectrumAnalyzerError>(e);
    if (st_.status != SA_OK) return static_cast<SpectrumAnalyzerError>(st_.status);   // ScalingError in the reference
    for (size_t i = 0; i < data_.size(); i++) data_[i].second = vals[i];
    return true;
  }
  Result<bool> apply_scaling_fn(Context &ctx, const scaling::Fn &fn) { return apply_scaling_fn(ctx, scaling::Chain{{fn.kind}}); }

 private:
  // float_cmp::approx_eq!(f32, a, b, ulps = 3): equal, or within f32::EPSILON, or within 3 units in the last place
  static bool approx_eq_ulps3(float a, float b) {
    if (a == b) return true;
    if (std::isnan(a) || std::isnan(b)) return false;
    if (std::fabs(a - b) <= std::numeric_limits<float>::epsilon()) return true;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    if ((ia < 0) != (ib < 0)) return false;
    const int64_t d = static_cast<int64_t>(ia) - static_cast<int64_t>(ib);
    return (d < 0 ? -d : d) <= 3;
  }
  // index b >= 1 of the first bin with frequency >= search_fr (bounds-checked like src/spectrum.rs:326-334)
  size_t upper_pair(float search_fr) const {
    if (search_fr < data_.front().first || search_fr > data_.back().first)
      throw std::out_of_range("Frequency " + std::to_string(search_fr) + "Hz is out of bounds [" +
                              std::to_string(data_.front().first) + "; " + std::to_string(data_.back().first) + "]!");
    size_t lo = 1, hi = data_.size() - 1;   // invariant: data_[hi].first >= search_fr
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (data_[mid].first >= search_fr) hi = mid; else lo = mid + 1;
    }
    return lo;
  }
  std::vector<Pair> data_;
  float res_;
  uint32_t n_;
  sa_frame_stats st_;
  uint32_t bin_lo_;
};

// ---- samples_fft_to_spectrum (src/lib.rs:157-162) -------------------------------------------------
// `window_kind` optionally fuses the caller-side windows::* step (SA_WINDOW_NONE = already windowed).
inline Result<FrequencySpectrum> samples_fft_to_spectrum(Context &ctx, const std::vector<float> &samples,
                                                         uint32_t sampling_rate, FrequencyLimit limit,
                                                         const scaling::Fn *scaling_fn = nullptr,
                                                         int window_kind = SA_WINDOW_NONE) {
  const size_t cap = samples.size() / 2 + 2;
  std::vector<float> freqs(cap), vals(cap);
  uint32_t n_bins = 0;
  sa_frame_stats st{};
  const int e = sa_spectrum_single(ctx.handle(), samples.data(), samples.size(), sampling_rate, limit.kind, limit.lo,
                                   limit.hi, window_kind, scaling_fn ? scaling_fn->kind : SA_SCALING_NONE, freqs.data(),
                                   vals.data(), &n_bins, &st);
  if (e != SA_OK) return static_cast<SpectrumAnalyzerError>(e);
  uint32_t bin_lo = 0, nb = 0;
  sa_limit_to_bins(static_cast<uint32_t>(samples.size()), sampling_rate, limit.kind, limit.lo, limit.hi, &bin_lo, &nb);
  std::vector<FrequencySpectrum::Pair> data(n_bins);
  for (uint32_t i = 0; i < n_bins; i++) data[i] = {freqs[i], vals[i]};
  const float res = static_cast<float>(sampling_rate) / static_cast<float>(samples.size());
  return FrequencySpectrum(std::move(data), res, static_cast<uint32_t>(samples.size()), st, bin_lo);
}

// ---- the batched entry point (host buffers) -------------------------------------------------------
struct BatchedSpectrum {
  std::vector<float> vals;             // [n_frames x n_bins]
  std::vector<sa_frame_stats> stats;   // [n_frames]
  std::vector<float> frequencies;      // [n_bins], shared by every frame
  uint32_t bin_lo = 0, n_bins = 0;
};

// `Sample` is float, or int16_t for PCM (converted `x as f32` on the device while the samples are loaded)
template <class Sample>
inline Result<BatchedSpectrum> samples_fft_to_spectrum_batched(Context &ctx, const Sample *frames, uint64_t n_frames,
                                                               uint32_t n, uint64_t frame_stride, uint32_t sampling_rate,
                                                               FrequencyLimit limit, const scaling::Fn *scaling_fn = nullptr,
                                                               int window_kind = SA_WINDOW_NONE,
                                                               uint32_t stats_mask = SA_STATS_ALL) {
  static_assert(std::is_same_v<Sample, float> || std::is_same_v<Sample, int16_t>, "f32 or int16 PCM samples");
  BatchedSpectrum out;
  int e = sa_limit_to_bins(n, sampling_rate, limit.kind, limit.lo, limit.hi, &out.bin_lo, &out.n_bins);
  if (e != SA_OK) return static_cast<SpectrumAnalyzerError>(e);
  out.vals.resize(n_frames * out.n_bins);
  out.stats.resize(n_frames);
  out.frequencies.resize(out.n_bins);
  sa_bin_frequencies(n, sampling_rate, out.bin_lo, out.n_bins, out.frequencies.data());
  const int sk = scaling_fn ? scaling_fn->kind : SA_SCALING_NONE;
  if constexpr (std::is_same_v<Sample, float>)
    e = sa_spectrum_batched_host(ctx.handle(), frames, n_frames, n, frame_stride, sampling_rate, limit.kind, limit.lo,
                                 limit.hi, window_kind, sk, stats_mask, out.vals.data(), out.n_bins, out.stats.data(),
                                 nullptr, nullptr);
  else
    e = sa_spectrum_batched_host_i16(ctx.handle(), frames, n_frames, n, frame_stride, sampling_rate, limit.kind, limit.lo,
                                     limit.hi, window_kind, sk, stats_mask, out.vals.data(), out.n_bins, out.stats.data(),
                                     nullptr, nullptr);
  if (e != SA_OK) return static_cast<SpectrumAnalyzerError>(e);
  return out;
}

// ---- windows (src/windows.rs:40-150): applied on the device (sa_window_host), bit-exact multipliers ----
namespace windows {
inline std::vector<float> apply(Context &ctx, int kind, const std::vector<float> &samples) {
  std::vector<float> out(samples.size());
  if (samples.empty()) return out;
  const uint32_t n = static_cast<uint32_t>(samples.size());
  if (int e = sa_window_host(ctx.handle(), kind, samples.data(), 1, n, n, out.data()); e != SA_OK)
    throw std::runtime_error(std::string("sa_window_host: ") + sa_status_string(e));
  return out;
}
inline std::vector<float> hann_window(Context &c, const std::vector<float> &s) { return apply(c, SA_WINDOW_HANN, s); }
inline std::vector<float> hamming_window(Context &c, const std::vector<float> &s) { return apply(c, SA_WINDOW_HAMMING, s); }
inline std::vector<float> blackman_harris_4term(Context &c, const std::vector<float> &s) { return apply(c, SA_WINDOW_BLACKMAN_HARRIS_4TERM, s); }
inline std::vector<float> blackman_harris_7term(Context &c, const std::vector<float> &s) { return apply(c, SA_WINDOW_BLACKMAN_HARRIS_7TERM, s); }
}  // namespace windows

}  // namespace spectrum_analyzer
