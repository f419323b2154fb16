// The reference's integration tests (src/tests/mod.rs) written against the C++ host mirror
// (spectrum_analyzer_b200/host/spectrum_analyzer.hpp) -> C ABI -> CUDA kernels.
// Built and run by tests/test_cpp_mirror_gpu.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "../../spectrum_analyzer_b200/host/spectrum_analyzer.hpp"

using namespace spectrum_analyzer;

static int g_checks = 0;
#define CHECK(cond)                                                                  \
  do {                                                                               \
    g_checks++;                                                                      \
    if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
  } while (0)

// src/tests/sine.rs:54-104 restated (test fixture)
static std::vector<float> sine_wave_audio_data_multiple(const std::vector<float> &freqs, uint32_t sr, uint32_t ms) {
  const float PI = 3.14159274101257324219f;
  const size_t count = (size_t)((float)sr * ((float)ms / 1000.0f));
  std::vector<float> out(count);
  for (size_t i = 0; i < count; i++) {
    const float t = (1.0f / (float)sr) * (float)i;
    float acc = 0.0f;
    for (float f : freqs) acc += sinf(t * f * 2.0f * PI);
    acc = acc * 32767.0f * 0.1f;
    out[i] = (float)(int16_t)acc;
  }
  return out;
}

int main() {
  Context ctx(0);

  {  // src/windows.rs:156-174
    auto w = windows::hamming_window(ctx, std::vector<float>(4, 1.0f));
    const float expect[4] = {0.08f, 0.77f, 0.77f, 0.08f};
    for (int i = 0; i < 4; i++) CHECK(std::fabs(w[i] - expect[i]) < 1e-5f);
    auto b = windows::blackman_harris_4term(ctx, std::vector<float>(4, 2.0f));
    const float eb[4] = {0.00012f, 1.04115f, 1.04115f, 0.00012f};
    for (int i = 0; i < 4; i++) CHECK(std::fabs(b[i] - eb[i]) < 1e-5f);
  }
  {  // src/tests/mod.rs:224-290 test_spectrum_frequency_limit_inclusive
    const uint32_t sr = 1024;
    auto audio = sine_wave_audio_data_multiple({512.0f}, sr, 1000);
    auto window = windows::hann_window(ctx, std::vector<float>(audio.begin(), audio.begin() + 512));
    auto s = samples_fft_to_spectrum(ctx, window, sr, FrequencyLimit::Max(400.0f)).unwrap();
    CHECK(s.min_fr() == 0.0f && s.max_fr() == 400.0f);
    s = samples_fft_to_spectrum(ctx, window, sr, FrequencyLimit::Min(100.0f)).unwrap();
    CHECK(s.min_fr() == 100.0f && s.max_fr() == (float)sr / 2.0f);
    s = samples_fft_to_spectrum(ctx, window, sr, FrequencyLimit::Range(412.0f, 510.0f)).unwrap();
    CHECK(s.min_fr() == 412.0f && s.max_fr() == 510.0f);
  }
  {  // mod.rs:293-321 test_spectrum_nyquist_theorem
    auto s = samples_fft_to_spectrum(ctx, std::vector<float>(4096, 0.0f), 44100, FrequencyLimit::All()).unwrap();
    size_t zeros = 0;
    for (auto &p : s.data()) zeros += p.second == 0.0f;
    CHECK(zeros == 4096 / 2 + 1);
    CHECK(s.min_fr() == 0.0f && s.max_fr() == 22050.0f);
  }
  {  // mod.rs:57-135 sine mix, Hann, Max(4000), zero-to-one: peaks at bins 5 / 93 / 351
    auto audio = sine_wave_audio_data_multiple({50.0f, 1000.0f, 3777.0f}, 44100, 1000);
    std::vector<float> window(audio.begin(), audio.begin() + 4096);
    auto s = samples_fft_to_spectrum(ctx, window, 44100, FrequencyLimit::Max(4000.0f), &scaling::scale_to_zero_to_one,
                                     SA_WINDOW_HANN).unwrap();
    CHECK(s.data().size() == 372);
    CHECK(s.max().second == 1.0f && std::fabs(s.max().first - 1001.29395f) < 1e-3f);
    CHECK(s.data()[351].second > 0.85f && s.data()[5].second > 0.85f && s.data()[46].second < 1e-4f);
  }
  {  // mod.rs:381-432 test_invalid_input
    const float nan = std::numeric_limits<float>::quiet_NaN(), inf = std::numeric_limits<float>::infinity();
    CHECK(samples_fft_to_spectrum(ctx, {0, 1, 2, 3, nan, 4, 5, 6}, 44100, FrequencyLimit::All()).unwrap_err() == SpectrumAnalyzerError::NaNValuesNotSupported);
    CHECK(samples_fft_to_spectrum(ctx, {0, 1, 2, 3, inf, 4, 5, 6}, 44100, FrequencyLimit::All()).unwrap_err() == SpectrumAnalyzerError::InfinityValuesNotSupported);
    CHECK(samples_fft_to_spectrum(ctx, {0.0f}, 44100, FrequencyLimit::All()).unwrap_err() == SpectrumAnalyzerError::TooFewSamples);
    CHECK(samples_fft_to_spectrum(ctx, std::vector<float>(4, 0.f), 44100, FrequencyLimit::Min(-1.0f)).unwrap_err() == SpectrumAnalyzerError::InvalidFrequencyLimitValueBelowMinimum);
    CHECK(samples_fft_to_spectrum(ctx, std::vector<float>(8, 0.f), 8, FrequencyLimit::Range(1.1f, 1.9f)).unwrap_err() == SpectrumAnalyzerError::FrequencyLimitTooNarrow);
    CHECK(samples_fft_to_spectrum(ctx, std::vector<float>(8, 0.f), 8, FrequencyLimit::Range(1.0f, 1.0f)).unwrap_err() == SpectrumAnalyzerError::FrequencyLimitTooNarrow);
    CHECK(samples_fft_to_spectrum(ctx, std::vector<float>(3, 0.f), 44100, FrequencyLimit::All()).unwrap_err() == SpectrumAnalyzerError::SamplesLengthNotAPowerOfTwo);
    CHECK(samples_fft_to_spectrum(ctx, {0.0f, 0.0f}, 44100, FrequencyLimit::All()).is_ok());     // mod.rs:434-438
  }
  {  // mod.rs:453-490 test_divide_by_n_has_effect: stats.n = samples_len -> bit-identical shared bins
    auto audio = sine_wave_audio_data_multiple({100.0f, 200.0f, 400.0f}, 1000, 2000);
    auto w = windows::hann_window(ctx, std::vector<float>(audio.begin(), audio.begin() + 1024));
    auto scaled = samples_fft_to_spectrum(ctx, w, 1000, FrequencyLimit::All(), &scaling::divide_by_N).unwrap();
    auto limited = samples_fft_to_spectrum(ctx, w, 1000, FrequencyLimit::Max(250.0f), &scaling::divide_by_N).unwrap();
    for (int i = 0; i < 256; i++) CHECK(scaled.data()[i].second == limited.data()[i].second);
  }
  {  // batched entry point: 64 frames x 2048, divide_by_N_sqrt, stats ordered, frequencies shared
    std::vector<float> x(64 * 2048);
    for (size_t i = 0; i < x.size(); i++) x[i] = std::sin(0.001f * (float)i) + 0.25f * std::cos(0.37f * (float)i);
    auto b = samples_fft_to_spectrum_batched(ctx, x.data(), 64, 2048, 2048, 44100, FrequencyLimit::All(),
                                             &scaling::divide_by_N_sqrt, SA_WINDOW_HAMMING).unwrap();
    CHECK(b.n_bins == 1025 && b.vals.size() == 64 * 1025 && b.frequencies[1024] == 22050.0f);
    for (auto &st : b.stats) CHECK(st.status == SA_OK && st.min_val <= st.median && st.median <= st.max_val);
  }
  std::printf("reference_suite OK: %d checks\n", g_checks);
  return 0;
}
