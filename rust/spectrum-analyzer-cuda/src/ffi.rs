//! Raw `extern "C"` declarations of include/spectrum_analyzer_cuda.h (ABI version 1).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct sa_ctx {
    _private: [u8; 0],
}

/// One 32-byte record per frame (statistics of the scaled values + per-frame status).
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct sa_frame_stats {
    pub min_val: f32,
    pub min_bin: u32,
    pub max_val: f32,
    pub max_bin: u32,
    pub average: f32,
    pub median: f32,
    pub status: u32,
    pub reserved: u32,
}

pub const SA_OK: c_int = 0;
pub const SA_ERR_SCALING: u32 = 9;
pub const SA_STATS_ALL: u32 = 7;
pub const SA_SCALING_ERR_INDEX_MASK: u32 = 0x0000_ffff;
pub const SA_SCALING_ERR_NAN: u32 = 0x8000_0000;
pub const SA_SCALING_ERR_POS_INF: u32 = 0x4000_0000;
pub const SA_SCALING_ERR_NEG_INF: u32 = 0x2000_0000;

extern "C" {
    pub fn sa_abi_version() -> u32;
    pub fn sa_status_string(status: c_int) -> *const c_char;
    pub fn sa_last_error_string() -> *const c_char;
    pub fn sa_ctx_create(device: c_int, out_ctx: *mut *mut sa_ctx) -> c_int;
    pub fn sa_ctx_create_on_stream(device: c_int, cuda_stream: *mut c_void, out_ctx: *mut *mut sa_ctx) -> c_int;
    pub fn sa_ctx_destroy(ctx: *mut sa_ctx) -> c_int;
    pub fn sa_ctx_sync(ctx: *mut sa_ctx) -> c_int;
    pub fn sa_ctx_set_fast_log10(ctx: *mut sa_ctx, enable: c_int) -> c_int;
    pub fn sa_ctx_launch_count(ctx: *const sa_ctx) -> u64;
    pub fn sa_ctx_last_kernel(ctx: *const sa_ctx) -> *const c_char;
    pub fn sa_window_coeffs(window_kind: c_int, n: u32, host_out: *mut f32) -> c_int;
    pub fn sa_limit_to_bins(n: u32, sampling_rate: u32, limit_kind: c_int, limit_lo: f32, limit_hi: f32,
                            bin_lo: *mut u32, n_bins: *mut u32) -> c_int;
    pub fn sa_bin_frequencies(n: u32, sampling_rate: u32, bin_lo: u32, n_bins: u32, host_out: *mut f32) -> c_int;
    pub fn sa_spectrum_batched(ctx: *mut sa_ctx, d_samples: *const f32, n_frames: u64, n: u32, frame_stride: u64,
                               sampling_rate: u32, limit_kind: c_int, limit_lo: f32, limit_hi: f32,
                               window_kind: c_int, scaling_kind: c_int, stats_mask: u32, d_out_vals: *mut f32,
                               out_stride: u64, d_stats: *mut sa_frame_stats, bin_lo: *mut u32,
                               n_bins: *mut u32) -> c_int;
    pub fn sa_spectrum_batched_i16(ctx: *mut sa_ctx, d_samples: *const i16, n_frames: u64, n: u32, frame_stride: u64,
                               sampling_rate: u32, limit_kind: c_int, limit_lo: f32, limit_hi: f32,
                               window_kind: c_int, scaling_kind: c_int, stats_mask: u32, d_out_vals: *mut f32,
                               out_stride: u64, d_stats: *mut sa_frame_stats, bin_lo: *mut u32,
                               n_bins: *mut u32) -> c_int;
    pub fn sa_spectrum_batched_host(ctx: *mut sa_ctx, h_samples: *const f32, n_frames: u64, n: u32,
                                    frame_stride: u64, sampling_rate: u32, limit_kind: c_int, limit_lo: f32,
                                    limit_hi: f32, window_kind: c_int, scaling_kind: c_int, stats_mask: u32,
                                    h_out_vals: *mut f32, out_stride: u64, h_stats: *mut sa_frame_stats,
                                    bin_lo: *mut u32, n_bins: *mut u32) -> c_int;
    pub fn sa_spectrum_batched_host_i16(ctx: *mut sa_ctx, h_samples: *const i16, n_frames: u64, n: u32,
                                    frame_stride: u64, sampling_rate: u32, limit_kind: c_int, limit_lo: f32,
                                    limit_hi: f32, window_kind: c_int, scaling_kind: c_int, stats_mask: u32,
                                    h_out_vals: *mut f32, out_stride: u64, h_stats: *mut sa_frame_stats,
                                    bin_lo: *mut u32, n_bins: *mut u32) -> c_int;
    pub fn sa_spectrum_single(ctx: *mut sa_ctx, h_samples: *const f32, n_samples: u64, sampling_rate: u32,
                              limit_kind: c_int, limit_lo: f32, limit_hi: f32, window_kind: c_int,
                              scaling_kind: c_int, h_freqs: *mut f32, h_vals: *mut f32, n_bins: *mut u32,
                              h_stats: *mut sa_frame_stats) -> c_int;
    pub fn sa_apply_scaling_batched(ctx: *mut sa_ctx, d_vals: *mut f32, n_frames: u64, n_bins: u32, stride: u64,
                                    bin_lo: u32, n: u32, scaling_kind: c_int, stats_mask: u32,
                                    d_stats: *mut sa_frame_stats) -> c_int;
    pub fn sa_apply_scaling_chain_batched(ctx: *mut sa_ctx, d_vals: *mut f32, n_frames: u64, n_bins: u32, stride: u64,
                                          bin_lo: u32, n: u32, scaling_kinds: *const c_int, n_kinds: u32,
                                          stats_mask: u32, d_stats: *mut sa_frame_stats) -> c_int;
    pub fn sa_apply_scaling_chain_host(ctx: *mut sa_ctx, h_vals: *mut f32, n_frames: u64, n_bins: u32, stride: u64,
                                          bin_lo: u32, n: u32, scaling_kinds: *const c_int, n_kinds: u32,
                                          stats_mask: u32, h_stats: *mut sa_frame_stats) -> c_int;
    pub fn sa_window_batched(ctx: *mut sa_ctx, window_kind: c_int, d_in: *const f32, n_frames: u64, n: u32,
                             frame_stride: u64, d_out: *mut f32) -> c_int;
    pub fn sa_window_host(ctx: *mut sa_ctx, window_kind: c_int, h_in: *const f32, n_frames: u64, n: u32,
                          frame_stride: u64, h_out: *mut f32) -> c_int;
    pub fn sa_generate_noise(ctx: *mut sa_ctx, d_out: *mut f32, first_index: u64, count: u64, seed: u64) -> c_int;
}
