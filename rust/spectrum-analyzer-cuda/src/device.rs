//! Device-resident side of the crate: Rust owns the device buffers and the stream, the kernels never leave the GPU
//! (BASELINE.json north_star: "Rust host code owning device buffers and streams").
//!
//! SOURCE ONLY -- never compiled in this repository's environment (no Rust toolchain); links libcudart for the four
//! runtime calls it needs.  Shape of use, one `Stream` + `Context` per GPU and host thread:
//!
//! ```ignore
//! let stream = Stream::new(0)?;
//! let ctx = Context::on_stream(0, &stream)?;
//! let samples = DeviceBuffer::<f32>::from_host(&pcm_as_f32)?;          // or filled by another kernel
//! let mut out = DeviceSpectra::alloc(n_frames, 4096, 44100, FrequencyLimit::All)?;
//! samples_fft_to_spectrum_batched_device(&ctx, &samples, 4096, 4096, n_frames, 44100, FrequencyLimit::All,
//!                                        Some(scaling::divide_by_N_sqrt), Window::Hann, &mut out)?;   // asynchronous
//! out.apply_scaling_fn(&ctx, &[scaling::scale_20_times_log10])?;      // still on the device
//! stream.synchronize()?;
//! let stats = out.stats.to_host()?;                                    // per-frame min/max/average/median + status
//! ```
use core::ffi::{c_int, c_void};
use core::marker::PhantomData;
use std::ptr;

use crate::{ffi, map_err, windows::Window, Context, FrequencyLimit, ScalingFn, SpectrumAnalyzerError};

extern "C" {
    fn cudaSetDevice(device: c_int) -> c_int;
    fn cudaMalloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn cudaFree(ptr: *mut c_void) -> c_int;
    fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
    fn cudaStreamCreateWithFlags(stream: *mut *mut c_void, flags: u32) -> c_int;
    fn cudaStreamSynchronize(stream: *mut c_void) -> c_int;
    fn cudaStreamDestroy(stream: *mut c_void) -> c_int;
}
const H2D: c_int = 1;
const D2H: c_int = 2;

fn cuda(e: c_int, what: &str) -> Result<(), SpectrumAnalyzerError> {
    if e == 0 { Ok(()) } else { Err(SpectrumAnalyzerError::Backend(format!("{what}: cudaError {e}"))) }
}

/// An owned, non-blocking CUDA stream.
pub struct Stream { raw: *mut c_void }
unsafe impl Send for Stream {}
impl Stream {
    pub fn new(device: i32) -> Result<Self, SpectrumAnalyzerError> {
        let mut raw = ptr::null_mut();
        unsafe {
            cuda(cudaSetDevice(device), "cudaSetDevice")?;
            cuda(cudaStreamCreateWithFlags(&mut raw, 1 /* cudaStreamNonBlocking */), "cudaStreamCreateWithFlags")?;
        }
        Ok(Self { raw })
    }
    pub fn synchronize(&self) -> Result<(), SpectrumAnalyzerError> { cuda(unsafe { cudaStreamSynchronize(self.raw) }, "cudaStreamSynchronize") }
    pub(crate) fn as_raw(&self) -> *mut c_void { self.raw }
}
impl Drop for Stream {
    fn drop(&mut self) { unsafe { cudaStreamDestroy(self.raw); } }
}

/// An owned device allocation of `len` elements of `T` (plain data: f32, i16, sa_frame_stats).
pub struct DeviceBuffer<T: Copy> { ptr: *mut T, len: usize, _t: PhantomData<T> }
unsafe impl<T: Copy> Send for DeviceBuffer<T> {}
impl<T: Copy> DeviceBuffer<T> {
    pub fn alloc(len: usize) -> Result<Self, SpectrumAnalyzerError> {
        let mut p: *mut c_void = ptr::null_mut();
        cuda(unsafe { cudaMalloc(&mut p, len.max(1) * core::mem::size_of::<T>()) }, "cudaMalloc")?;
        Ok(Self { ptr: p as *mut T, len, _t: PhantomData })
    }
    pub fn from_host(src: &[T]) -> Result<Self, SpectrumAnalyzerError> {
        let b = Self::alloc(src.len())?;
        cuda(unsafe { cudaMemcpy(b.ptr as *mut c_void, src.as_ptr() as *const c_void, core::mem::size_of_val(src), H2D) }, "cudaMemcpy")?;
        Ok(b)
    }
    pub fn to_host(&self) -> Result<Vec<T>, SpectrumAnalyzerError> where T: Default {
        let mut v = vec![T::default(); self.len];
        cuda(unsafe { cudaMemcpy(v.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, self.len * core::mem::size_of::<T>(), D2H) }, "cudaMemcpy")?;
        Ok(v)
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn as_ptr(&self) -> *const T { self.ptr }
    pub fn as_mut_ptr(&mut self) -> *mut T { self.ptr }
}
impl<T: Copy> Drop for DeviceBuffer<T> {
    fn drop(&mut self) { unsafe { cudaFree(self.ptr as *mut c_void); } }
}

/// Device-resident result of the batched call: `vals[f * n_bins + i]` is bin `bin_lo + i` of frame `f`.
pub struct DeviceSpectra {
    pub vals: DeviceBuffer<f32>,
    pub stats: DeviceBuffer<ffi::sa_frame_stats>,
    pub frequencies: Vec<f32>,
    pub n_frames: u64,
    pub bin_lo: u32,
    pub n_bins: u32,
    pub samples_len: u32,
}
impl DeviceSpectra {
    pub fn alloc(n_frames: u64, n: u32, sampling_rate: u32, limit: FrequencyLimit) -> Result<Self, SpectrumAnalyzerError> {
        let (lk, lo, hi) = limit.raw();
        let (mut bin_lo, mut n_bins) = (0u32, 0u32);
        let e = unsafe { ffi::sa_limit_to_bins(n, sampling_rate, lk, lo, hi, &mut bin_lo, &mut n_bins) };
        if e != ffi::SA_OK { return Err(map_err(e, limit)); }
        let mut frequencies = vec![0f32; n_bins as usize];
        unsafe { ffi::sa_bin_frequencies(n, sampling_rate, bin_lo, n_bins, frequencies.as_mut_ptr()); }
        Ok(Self { vals: DeviceBuffer::alloc((n_frames * n_bins as u64) as usize)?, stats: DeviceBuffer::alloc(n_frames as usize)?,
                  frequencies, n_frames, bin_lo, n_bins, samples_len: n })
    }
    /// `FrequencySpectrum::apply_scaling_fn` (reference src/spectrum.rs:128-168) with `scaling::combined(chain)` on every
    /// spectrum of the batch, on the device.  Per-frame `ScalingError`s are reported in `stats[f].status` / `.reserved`.
    pub fn apply_scaling_fn(&mut self, ctx: &Context, chain: &[ScalingFn]) -> Result<(), SpectrumAnalyzerError> {
        let kinds: Vec<c_int> = chain.iter().map(|s| *s as c_int).collect();
        let e = unsafe {
            ffi::sa_apply_scaling_chain_batched(ctx.raw(), self.vals.as_mut_ptr(), self.n_frames, self.n_bins, self.n_bins as u64,
                                                self.bin_lo, self.samples_len, kinds.as_ptr(), kinds.len() as u32,
                                                ffi::SA_STATS_ALL, self.stats.as_mut_ptr())
        };
        if e != ffi::SA_OK { return Err(map_err(e, FrequencyLimit::All)); }
        Ok(())
    }
}

/// The batched `[frames x N]` entry point on device-resident samples; asynchronous on the context's stream.
#[allow(clippy::too_many_arguments)]
pub fn samples_fft_to_spectrum_batched_device(ctx: &Context, samples: &DeviceBuffer<f32>, n: u32, frame_stride: u64,
                                              n_frames: u64, sampling_rate: u32, frequency_limit: FrequencyLimit,
                                              scaling_fn: Option<ScalingFn>, window: Window, out: &mut DeviceSpectra)
                                              -> Result<(), SpectrumAnalyzerError> {
    assert!(n_frames == 0 || samples.len() as u64 >= (n_frames - 1) * frame_stride + n as u64);
    assert!(out.n_frames >= n_frames && out.samples_len == n);
    let (lk, lo, hi) = frequency_limit.raw();
    let e = unsafe {
        ffi::sa_spectrum_batched(ctx.raw(), samples.as_ptr(), n_frames, n, frame_stride, sampling_rate, lk, lo, hi,
                                 window as c_int, scaling_fn.map_or(0, |s| s as c_int), ffi::SA_STATS_ALL,
                                 out.vals.as_mut_ptr(), out.n_bins as u64, out.stats.as_mut_ptr(), ptr::null_mut(), ptr::null_mut())
    };
    if e != ffi::SA_OK { return Err(map_err(e, frequency_limit)); }
    Ok(())
}
