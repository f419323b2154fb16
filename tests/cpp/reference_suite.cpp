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
    // the same frames as int16 PCM (src/tests/mod.rs:70-73 converts on the caller's side): identical bits
    std::vector<int16_t> pcm(x.size());
    std::vector<float> xf(x.size());
    for (size_t i = 0; i < x.size(); i++) { pcm[i] = (int16_t)(x[i] * 20000.0f); xf[i] = (float)pcm[i]; }
    auto bf = samples_fft_to_spectrum_batched(ctx, xf.data(), 64, 2048, 2048, 44100, FrequencyLimit::All(),
                                              &scaling::divide_by_N_sqrt, SA_WINDOW_HAMMING).unwrap();
    auto bi = samples_fft_to_spectrum_batched(ctx, pcm.data(), 64, 2048, 2048, 44100, FrequencyLimit::All(),
                                              &scaling::divide_by_N_sqrt, SA_WINDOW_HAMMING).unwrap();
    CHECK(bf.vals == bi.vals);
    for (size_t f = 0; f < 64; f++) CHECK(bf.stats[f].median == bi.stats[f].median && bf.stats[f].max_bin == bi.stats[f].max_bin);
  }
  {  // src/spectrum.rs:656-792 test_spectrum_basic: lookups on a hand-made spectrum (host-side, known answers)
    std::vector<FrequencySpectrum::Pair> v = {{0.f, 5.f},    {50.f, 50.f},  {100.f, 100.f}, {150.f, 150.f}, {200.f, 100.f},
                                              {250.f, 20.f}, {300.f, 0.f},  {450.f, 200.f}, {500.f, 100.f}};
    sa_frame_stats st{};
    st.min_val = 0.f; st.min_bin = 6; st.max_val = 200.f; st.max_bin = 7; st.average = 80.55556f; st.median = 100.f;
    FrequencySpectrum s(v, 50.0f, (uint32_t)v.size(), st, 0);
    float dc = -1.f;
    CHECK(s.dc_component(&dc) && dc == 5.0f);
    CHECK(s.min() == FrequencySpectrum::Pair(300.f, 0.f) && s.max() == FrequencySpectrum::Pair(450.f, 200.f));
    CHECK(s.range() == 200.f && s.min_fr() == 0.f && s.max_fr() == 500.f && s.frequency_resolution() == 50.f);
    const float fx[8] = {0.f, 50.f, 150.f, 200.f, 250.f, 300.f, 375.f, 450.f};
    const float fy[8] = {5.f, 50.f, 150.f, 100.f, 20.f, 0.f, 100.f, 200.f};
    for (int i = 0; i < 8; i++) CHECK(s.freq_val_exact(fx[i]) == fy[i]);
    CHECK(s.freq_val_closest(0.f) == FrequencySpectrum::Pair(0.f, 5.f));
    CHECK(s.freq_val_closest(50.f) == FrequencySpectrum::Pair(50.f, 50.f));
    CHECK(s.freq_val_closest(450.f) == FrequencySpectrum::Pair(450.f, 200.f));
    CHECK(s.freq_val_closest(448.f) == FrequencySpectrum::Pair(450.f, 200.f));
    CHECK(s.freq_val_closest(400.f) == FrequencySpectrum::Pair(450.f, 200.f));
    CHECK(s.freq_val_closest(47.3f) == FrequencySpectrum::Pair(50.f, 50.f));
    CHECK(s.freq_val_closest(51.3f) == FrequencySpectrum::Pair(50.f, 50.f));
    // src/spectrum.rs:795-867: out-of-range lookups panic in the reference -> std::out_of_range here
    int thrown = 0;
    for (float bad : {-1.0f, 501.0f}) {
      try { (void)s.freq_val_exact(bad); } catch (const std::out_of_range &) { thrown++; }
      try { (void)s.freq_val_closest(bad); } catch (const std::out_of_range &) { thrown++; }
    }
    CHECK(thrown == 4);
    // src/spectrum.rs:984-997 test_mel_getter: 450 mel = 343.6 Hz is inside, 4500 mel is far outside
    std::vector<FrequencySpectrum::Pair> two = {{0.f, 5.f}, {450.f, 200.f}};
    FrequencySpectrum t(two, 50.0f, 2, st, 0);
    CHECK(std::fabs(FrequencySpectrum::mel_to_hertz(450.f) - 343.6f) < 0.5f);
    thrown = 0;
    try { (void)t.mel_val(4500.f); } catch (const std::out_of_range &) { thrown++; }
    CHECK(thrown == 1);
    CHECK(std::fabs(FrequencySpectrum::hertz_to_mel(FrequencySpectrum::mel_to_hertz(1000.f)) - 1000.f) < 1e-2f);
    auto m = s.to_map();
    CHECK(m.size() == 9 && m.at(450) == 200.f && m.at(0) == 5.f);
    CHECK(s.to_mel_map().size() == 9);
  }
  {  // FrequencySpectrum::apply_scaling_fn (src/spectrum.rs:128-168) and scaling::combined (src/scaling.rs:174-182):
     // scaling after the fact == scaling in the call, bit for bit, and the statistics are recalculated
    auto audio = sine_wave_audio_data_multiple({50.0f, 1000.0f, 3777.0f}, 44100, 1000);
    std::vector<float> frame(audio.begin(), audio.begin() + 2048);
    auto plain = samples_fft_to_spectrum(ctx, frame, 44100, FrequencyLimit::Max(4000.0f), nullptr, SA_WINDOW_HANN).unwrap();
    auto direct = samples_fft_to_spectrum(ctx, frame, 44100, FrequencyLimit::Max(4000.0f), &scaling::divide_by_N_sqrt,
                                          SA_WINDOW_HANN).unwrap();
    CHECK(plain.apply_scaling_fn(ctx, scaling::divide_by_N_sqrt).is_ok());
    CHECK(plain.data().size() == direct.data().size());
    for (size_t i = 0; i < plain.data().size(); i++) CHECK(plain.data()[i].second == direct.data()[i].second);
    CHECK(plain.max() == direct.max() && plain.min() == direct.min() && plain.median() == direct.median());
    // combined(divide_by_N, zero_to_one): both steps see the statistics from before the call
    auto a = samples_fft_to_spectrum(ctx, frame, 44100, FrequencyLimit::Max(4000.0f), nullptr, SA_WINDOW_HANN).unwrap();
    const float old_max = a.max().second;
    std::vector<float> before;
    for (auto &p : a.data()) before.push_back(p.second);
    CHECK(a.apply_scaling_fn(ctx, scaling::combined({scaling::divide_by_N, scaling::scale_to_zero_to_one})).is_ok());
    for (size_t i = 0; i < before.size(); i++) CHECK(a.data()[i].second == (before[i] / 2048.0f) / old_max);
    CHECK(a.max().second == (old_max / 2048.0f) / old_max);
  }
  std::printf("reference_suite OK: %d checks\n", g_checks);
  return 0;
}
