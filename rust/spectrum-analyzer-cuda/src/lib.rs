//! `spectrum-analyzer-cuda`: safe wrappers with the reference crate's signatures over the C ABI.
//!
//! SOURCE ONLY -- never compiled in this repository's environment (no Rust toolchain).  The names
//! follow phip1611/spectrum-analyzer v1.8.0 so that a caller switches by changing the `use` line:
//!
//! ```ignore
//! use spectrum_analyzer_cuda::{samples_fft_to_spectrum, FrequencyLimit, scaling, windows, Context};
//! let ctx = Context::new(0)?;
//! let hann = windows::hann_window(&ctx, &samples)?;
//! let spectrum = samples_fft_to_spectrum(&ctx, &hann, 44100, FrequencyLimit::All, Some(scaling::divide_by_N_sqrt))?;
//! ```
pub mod device;
pub mod ffi;

use core::ffi::c_int;
use std::ptr;

/// `SpectrumAnalyzerError` (reference src/error.rs:37-54) plus the two conditions the reference
/// panics on (unsupported length, non-finite spectrum value) and backend failures.
#[derive(Debug, Clone, PartialEq)]
pub enum SpectrumAnalyzerError {
    TooFewSamples,
    NaNValuesNotSupported,
    InfinityValuesNotSupported,
    InvalidFrequencyLimit(FrequencyLimitError),
    FrequencyLimitTooNarrow,
    SamplesLengthNotAPowerOfTwo,
    ScalingError(f32, f32),
    UnsupportedLength,
    NonFiniteResult,
    Backend(String),
}

#[derive(Debug, Clone, PartialEq)]
pub enum FrequencyLimitError {
    ValueBelowMinimum(f32),
    ValueAboveNyquist(f32),
    InvalidRange(f32, f32),
}

/// `FrequencyLimit` (reference src/limit.rs:38-53).
#[derive(Debug, Copy, Clone)]
pub enum FrequencyLimit {
    All,
    Min(f32),
    Max(f32),
    Range(f32, f32),
}

impl FrequencyLimit {
    fn raw(&self) -> (c_int, f32, f32) {
        match *self {
            FrequencyLimit::All => (0, 0.0, 0.0),
            FrequencyLimit::Min(x) => (1, x, 0.0),
            FrequencyLimit::Max(x) => (2, 0.0, x),
            FrequencyLimit::Range(a, b) => (3, a, b),
        }
    }
}

/// Built-in scaling functions (reference src/scaling.rs:104-162).  Arbitrary closures
/// (`&dyn Fn(f32, &SpectrumDataStats) -> f32`) cannot run on the device: the type system rejects them.
#[allow(non_camel_case_types)]
#[derive(Debug, Copy, Clone)]
pub enum ScalingFn {
    divide_by_N = 1,
    divide_by_N_sqrt = 2,
    scale_to_zero_to_one = 3,
    scale_20_times_log10 = 4,
}
pub mod scaling {
    pub use super::ScalingFn::*;
}

fn map_err(code: c_int, limit: FrequencyLimit) -> SpectrumAnalyzerError {
    use SpectrumAnalyzerError::*;
    let (lo, hi) = match limit {
        FrequencyLimit::Min(x) => (x, x),
        FrequencyLimit::Max(x) => (x, x),
        FrequencyLimit::Range(a, b) => (a, b),
        FrequencyLimit::All => (0.0, 0.0),
    };
    match code {
        1 => TooFewSamples,
        2 => NaNValuesNotSupported,
        3 => InfinityValuesNotSupported,
        4 => SamplesLengthNotAPowerOfTwo,
        5 => InvalidFrequencyLimit(FrequencyLimitError::ValueBelowMinimum(if lo < 0.0 { lo } else { hi })),
        6 => InvalidFrequencyLimit(FrequencyLimitError::ValueAboveNyquist(hi.max(lo))),
        7 => InvalidFrequencyLimit(FrequencyLimitError::InvalidRange(lo, hi)),
        8 => FrequencyLimitTooNarrow,
        9 => ScalingError(f32::NAN, f32::NAN),
        10 => UnsupportedLength,
        11 => NonFiniteResult,
        _ => Backend(unsafe { std::ffi::CStr::from_ptr(ffi::sa_last_error_string()) }.to_string_lossy().into_owned()),
    }
}

/// One GPU + one stream.  `!Sync`: use one context per host thread (one per GPU when sharding frames).
pub struct Context {
    raw: *mut ffi::sa_ctx,
}
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, SpectrumAnalyzerError> {
        let mut raw = ptr::null_mut();
        // SAFETY: out-pointer is valid; the library returns an owned handle or an error code.
        let e = unsafe { ffi::sa_ctx_create(device, &mut raw) };
        if e != ffi::SA_OK { return Err(map_err(e, FrequencyLimit::All)); }
        Ok(Self { raw })
    }
    /// A context that launches on a caller-owned stream (see `device::Stream`).
    pub fn on_stream(device: i32, stream: &device::Stream) -> Result<Self, SpectrumAnalyzerError> {
        let mut raw = ptr::null_mut();
        let e = unsafe { ffi::sa_ctx_create_on_stream(device, stream.as_raw(), &mut raw) };
        if e != ffi::SA_OK { return Err(map_err(e, FrequencyLimit::All)); }
        Ok(Self { raw })
    }
    pub fn sync(&self) { unsafe { ffi::sa_ctx_sync(self.raw); } }
    pub(crate) fn raw(&self) -> *mut ffi::sa_ctx { self.raw }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { ffi::sa_ctx_destroy(self.raw); } }
}

/// `FrequencySpectrum` (reference src/spectrum.rs:48-253): same getters, data as (frequency, value) pairs.
#[derive(Debug, Clone)]
pub struct FrequencySpectrum {
    data: Vec<(f32, f32)>,
    frequency_resolution: f32,
    samples_len: u32,
    stats: ffi::sa_frame_stats,
    bin_lo: u32,
}
impl FrequencySpectrum {
    pub fn average(&self) -> f32 { self.stats.average }
    pub fn median(&self) -> f32 { self.stats.median }
    pub fn max(&self) -> (f32, f32) { (self.data[(self.stats.max_bin - self.bin_lo) as usize].0, self.stats.max_val) }
    pub fn min(&self) -> (f32, f32) { (self.data[(self.stats.min_bin - self.bin_lo) as usize].0, self.stats.min_val) }
    pub fn range(&self) -> f32 { self.stats.max_val - self.stats.min_val }
    pub fn data(&self) -> &[(f32, f32)] { &self.data }
    pub fn frequency_resolution(&self) -> f32 { self.frequency_resolution }
    pub fn samples_len(&self) -> u32 { self.samples_len }
    pub fn max_fr(&self) -> f32 { self.data[self.data.len() - 1].0 }
    pub fn min_fr(&self) -> f32 { self.data[0].0 }
    pub fn dc_component(&self) -> Option<f32> { if self.data[0].0 == 0.0 { Some(self.data[0].1) } else { None } }

    /// `apply_scaling_fn` (reference src/spectrum.rs:128-168) for a chain of built-in functions
    /// (`scaling::combined`, src/scaling.rs:174-182: every function sees the statistics from before the call);
    /// computed on the device, statistics recalculated.
    pub fn apply_scaling_fn(&mut self, ctx: &Context, chain: &[ScalingFn]) -> Result<(), SpectrumAnalyzerError> {
        let kinds: Vec<c_int> = chain.iter().map(|s| *s as c_int).collect();
        let mut vals: Vec<f32> = self.data.iter().map(|p| p.1).collect();
        let e = unsafe {
            ffi::sa_apply_scaling_chain_host(ctx.raw, vals.as_mut_ptr(), 1, vals.len() as u32, vals.len() as u64,
                                             self.bin_lo, self.samples_len, kinds.as_ptr(), kinds.len() as u32,
                                             ffi::SA_STATS_ALL, &mut self.stats)
        };
        if e != ffi::SA_OK { return Err(map_err(e, FrequencyLimit::All)); }
        // the values in front of an offender are replaced, like the reference's in-place loop (src/spectrum.rs:150-163)
        for (p, v) in self.data.iter_mut().zip(vals.iter()) { p.1 = *v; }
        if self.stats.status == ffi::SA_ERR_SCALING {
            let r = self.stats.reserved;
            let new = if r & ffi::SA_SCALING_ERR_NAN != 0 { f32::NAN }
                      else if r & ffi::SA_SCALING_ERR_POS_INF != 0 { f32::INFINITY } else { f32::NEG_INFINITY };
            return Err(SpectrumAnalyzerError::ScalingError(vals[(r & ffi::SA_SCALING_ERR_INDEX_MASK) as usize], new));
        }
        Ok(())
    }
}

/// Drop-in for `samples_fft_to_spectrum` (reference src/lib.rs:157-162); `ctx` is the only extra argument.
pub fn samples_fft_to_spectrum(ctx: &Context, samples: &[f32], sampling_rate: u32, frequency_limit: FrequencyLimit,
                               scaling_fn: Option<ScalingFn>) -> Result<FrequencySpectrum, SpectrumAnalyzerError> {
    let cap = samples.len() / 2 + 2;
    let (mut freqs, mut vals) = (vec![0f32; cap], vec![0f32; cap]);
    let (mut n_bins, mut bin_lo, mut nb2) = (0u32, 0u32, 0u32);
    let mut st = ffi::sa_frame_stats::default();
    let (lk, lo, hi) = frequency_limit.raw();
    // SAFETY: all pointers reference live buffers of the documented sizes.
    let e = unsafe {
        ffi::sa_spectrum_single(ctx.raw, samples.as_ptr(), samples.len() as u64, sampling_rate, lk, lo, hi, 0,
                                scaling_fn.map_or(0, |s| s as c_int), freqs.as_mut_ptr(), vals.as_mut_ptr(),
                                &mut n_bins, &mut st)
    };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    unsafe { ffi::sa_limit_to_bins(samples.len() as u32, sampling_rate, lk, lo, hi, &mut bin_lo, &mut nb2); }
    let data = freqs[..n_bins as usize].iter().copied().zip(vals[..n_bins as usize].iter().copied()).collect();
    Ok(FrequencySpectrum { data, frequency_resolution: sampling_rate as f32 / samples.len() as f32,
                           samples_len: samples.len() as u32, stats: st, bin_lo })
}

/// Result of the batched entry point (SoA): `vals[frames x n_bins]`, one shared frequency vector, per-frame stats.
pub struct BatchedSpectrum {
    pub vals: Vec<f32>,
    pub stats: Vec<ffi::sa_frame_stats>,
    pub frequencies: Vec<f32>,
    pub bin_lo: u32,
    pub n_bins: u32,
}

/// The new `[frames x N]` entry point, same arguments as the single-frame call plus the window to fuse.
#[allow(clippy::too_many_arguments)]
pub fn samples_fft_to_spectrum_batched(ctx: &Context, frames: &[f32], n: u32, frame_stride: u64, n_frames: u64,
                                       sampling_rate: u32, frequency_limit: FrequencyLimit,
                                       scaling_fn: Option<ScalingFn>, window: windows::Window)
                                       -> Result<BatchedSpectrum, SpectrumAnalyzerError> {
    let (lk, lo, hi) = frequency_limit.raw();
    let (mut bin_lo, mut n_bins) = (0u32, 0u32);
    let e = unsafe { ffi::sa_limit_to_bins(n, sampling_rate, lk, lo, hi, &mut bin_lo, &mut n_bins) };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    assert!(n_frames == 0 || frames.len() as u64 >= (n_frames - 1) * frame_stride + n as u64);
    let mut out = BatchedSpectrum { vals: vec![0.0; (n_frames * n_bins as u64) as usize],
                                    stats: vec![ffi::sa_frame_stats::default(); n_frames as usize],
                                    frequencies: vec![0.0; n_bins as usize], bin_lo, n_bins };
    unsafe { ffi::sa_bin_frequencies(n, sampling_rate, bin_lo, n_bins, out.frequencies.as_mut_ptr()); }
    let e = unsafe {
        ffi::sa_spectrum_batched_host(ctx.raw, frames.as_ptr(), n_frames, n, frame_stride, sampling_rate, lk, lo, hi,
                                      window as c_int, scaling_fn.map_or(0, |s| s as c_int), ffi::SA_STATS_ALL,
                                      out.vals.as_mut_ptr(), n_bins as u64, out.stats.as_mut_ptr(),
                                      ptr::null_mut(), ptr::null_mut())
    };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    Ok(out)
}

/// The batched entry point on int16 PCM: replaces the caller-side `samples.iter().map(|x| *x as f32)` of the
/// reference's tests and examples (src/tests/mod.rs:70-73); the conversion happens while the kernel loads.
#[allow(clippy::too_many_arguments)]
pub fn samples_fft_to_spectrum_batched_i16(ctx: &Context, frames: &[i16], n: u32, frame_stride: u64, n_frames: u64,
                                           sampling_rate: u32, frequency_limit: FrequencyLimit,
                                           scaling_fn: Option<ScalingFn>, window: windows::Window)
                                           -> Result<BatchedSpectrum, SpectrumAnalyzerError> {
    let (lk, lo, hi) = frequency_limit.raw();
    let (mut bin_lo, mut n_bins) = (0u32, 0u32);
    let e = unsafe { ffi::sa_limit_to_bins(n, sampling_rate, lk, lo, hi, &mut bin_lo, &mut n_bins) };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    assert!(n_frames == 0 || frames.len() as u64 >= (n_frames - 1) * frame_stride + n as u64);
    let mut out = BatchedSpectrum { vals: vec![0.0; (n_frames * n_bins as u64) as usize],
                                    stats: vec![ffi::sa_frame_stats::default(); n_frames as usize],
                                    frequencies: vec![0.0; n_bins as usize], bin_lo, n_bins };
    unsafe { ffi::sa_bin_frequencies(n, sampling_rate, bin_lo, n_bins, out.frequencies.as_mut_ptr()); }
    let e = unsafe {
        ffi::sa_spectrum_batched_host_i16(ctx.raw, frames.as_ptr(), n_frames, n, frame_stride, sampling_rate, lk, lo, hi,
                                          window as c_int, scaling_fn.map_or(0, |s| s as c_int), ffi::SA_STATS_ALL,
                                          out.vals.as_mut_ptr(), n_bins as u64, out.stats.as_mut_ptr(),
                                          ptr::null_mut(), ptr::null_mut())
    };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    Ok(out)
}

/// `windows::*` (reference src/windows.rs:40,58,79,97), applied on the device with bit-exact multipliers.
pub mod windows {
    use super::*;
    #[derive(Debug, Copy, Clone)]
    pub enum Window { None = 0, Hann = 1, Hamming = 2, BlackmanHarris4Term = 3, BlackmanHarris7Term = 4 }
    fn apply(ctx: &Context, w: Window, samples: &[f32]) -> Result<Vec<f32>, SpectrumAnalyzerError> {
        let mut out = vec![0f32; samples.len()];
        if samples.is_empty() { return Ok(out); }
        let n = samples.len() as u32;
        let e = unsafe { ffi::sa_window_host(ctx.raw, w as c_int, samples.as_ptr(), 1, n, n as u64, out.as_mut_ptr()) };
        if e != ffi::SA_OK { return Err(map_err(e, FrequencyLimit::All)); }
        Ok(out)
    }
    pub fn hann_window(ctx: &Context, s: &[f32]) -> Result<Vec<f32>, SpectrumAnalyzerError> { apply(ctx, Window::Hann, s) }
    pub fn hamming_window(ctx: &Context, s: &[f32]) -> Result<Vec<f32>, SpectrumAnalyzerError> { apply(ctx, Window::Hamming, s) }
    pub fn blackman_harris_4term(ctx: &Context, s: &[f32]) -> Result<Vec<f32>, SpectrumAnalyzerError> { apply(ctx, Window::BlackmanHarris4Term, s) }
    pub fn blackman_harris_7term(ctx: &Context, s: &[f32]) -> Result<Vec<f32>, SpectrumAnalyzerError> { apply(ctx, Window::BlackmanHarris7Term, s) }
}
