//! Raw bindings of `include/soundkit_cuda.h` (ABI version 1). One-to-one with the C header; the safe
//! drop-in functions live in the `soundkit-cuda` crate. SOURCE ONLY in this repository: the build
//! image has no `cargo`/`rustc`, the same surface is exercised through ctypes (`soundkit_b200/capi.py`)
//! and C++ (`soundkit_b200/csrc/host/soundkit_host.hpp`).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)] pub struct sk_ctx { _private: [u8; 0] }
#[repr(C)] pub struct sk_resampler { _private: [u8; 0] }
#[repr(C)] pub struct sk_resampler_batch { _private: [u8; 0] }

pub const SK_OK: c_int = 0;
pub const SK_E_ARG: c_int = -1;
pub const SK_E_ALLOC: c_int = -2;
pub const SK_E_CUDA: c_int = -3;
pub const SK_E_STATE: c_int = -4;
pub const SK_E_NODEVICE: c_int = -5;

pub const SK_LAW_MULAW: c_int = 0;
pub const SK_LAW_ALAW: c_int = 1;
pub const SK_KIND_I16: c_int = 0;
pub const SK_KIND_I32: c_int = 1;
pub const SK_KIND_F32: c_int = 2;
pub const SK_FMT_SIGNED: c_int = 0;
pub const SK_FMT_FLOAT: c_int = 1;
pub const SK_MODE_A: c_int = 0;
pub const SK_MODE_B: c_int = 1;
pub const SK_MIX_NONE: c_int = 0;
pub const SK_MIX_SUM_THEN_SCALE: c_int = 1;
pub const SK_MIX_SCALE_THEN_SUM: c_int = 2;
pub const SK_SEM_ONESHOT: c_int = 0;
pub const SK_SEM_STREAMING: c_int = 1;
pub const SK_RESAMPLE_BIT_EXACT: c_int = 0;
pub const SK_RESAMPLE_FAST: c_int = 1;

/// `enum sk_flat_op`: the flat functions of `soundkit/src/audio_bytes.rs`
#[repr(C)] #[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum sk_flat_op {
    I16LE_TO_F32 = 0, S16LE_TO_I16 = 1, S16BE_TO_I16 = 2, S16LE_TO_I32 = 3, S24LE_TO_I32 = 4, S24LE_TO_I16 = 5,
    S24BE_TO_I16 = 6, S32LE_TO_I32 = 7, S32BE_TO_I32 = 8, S32LE_TO_S24 = 9, S32BE_TO_S24 = 10, S32LE_TO_F32 = 11,
    S32BE_TO_F32 = 12, S32LE_TO_I16 = 13, S32BE_TO_I16 = 14, F32LE_TO_I16 = 15, F32BE_TO_I16 = 16, F32LE_TO_I32 = 17,
    F32LE_TO_S24 = 18,
    // packet scalers of soundkit/src/audio_packet.rs and container-PCM unpackers
    PKT_S16LE_TO_F32 = 19, PKT_S24LE_TO_F32 = 20, PKT_S32LE_TO_I16 = 21, PKT_F32LE_TO_I16 = 22, U8_TO_S16 = 23,
}

#[repr(C)] pub struct sk_norm48k { _private: [u8; 0] }
pub const SK_G726_STATE_INTS: usize = 24;
pub const SK_G722_STATE_INTS: usize = 68;

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct sk_resample_params {
    pub sinc_len: u32, pub oversampling: u32, pub f_cutoff: c_float, pub interp: c_int, pub window: c_int,
    pub max_ratio_rel: c_double,
}

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct sk_batch_desc {
    pub input: *const c_void, pub in_stride: usize, pub n_streams: usize, pub frames: usize,
    pub bits: c_int, pub fmt: c_int, pub big_endian: c_int, pub channels: c_int, pub scale_mode: c_int,
    pub in_rate: u32, pub out_rate: u32, pub semantics: c_int, pub mix_mode: c_int,
    pub out: *mut c_float, pub out_stride: usize, pub out_stride_ch: usize,
}

extern "C" {
    pub fn sk_abi_version() -> c_int;
    pub fn sk_ctx_create(device: c_int, out: *mut *mut sk_ctx) -> c_int;
    pub fn sk_ctx_destroy(ctx: *mut sk_ctx);
    pub fn sk_last_error(ctx: *const sk_ctx) -> *const c_char;
    pub fn sk_ctx_set_stream(ctx: *mut sk_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn sk_ctx_reset_stream(ctx: *mut sk_ctx) -> c_int;
    pub fn sk_ctx_synchronize(ctx: *mut sk_ctx) -> c_int;
    pub fn sk_ctx_launch_count(ctx: *const sk_ctx) -> u64;
    pub fn sk_ctx_set_batch_slab_bytes(ctx: *mut sk_ctx, bytes: usize) -> c_int;
    pub fn sk_ctx_set_resample_mode(ctx: *mut sk_ctx, mode: c_int) -> c_int;
    pub fn sk_ctx_resample_fallback_outputs(ctx: *mut sk_ctx, count: *mut u64) -> c_int;
    pub fn sk_ctx_measure_ffma_peak(ctx: *mut sk_ctx, tflops: *mut c_double) -> c_int;
    pub fn sk_ctx_measure_ffma_peaks(ctx: *mut sk_ctx, scalar: *mut c_double, packed: *mut c_double) -> c_int;
    pub fn sk_ctx_poison_shared_memory(ctx: *mut sk_ctx, pattern: u32) -> c_int;

    pub fn sk_g711_decode(ctx: *mut sk_ctx, law: c_int, input: *const u8, n: usize, out_kind: c_int, out: *mut c_void) -> c_int;
    pub fn sk_g711_decode_dev(ctx: *mut sk_ctx, law: c_int, input: *const u8, n: usize, out_kind: c_int, out: *mut c_void) -> c_int;
    pub fn sk_g711_encode(ctx: *mut sk_ctx, law: c_int, in_kind: c_int, input: *const c_void, n: usize, out: *mut u8) -> c_int;
    pub fn sk_g711_encode_dev(ctx: *mut sk_ctx, law: c_int, in_kind: c_int, input: *const c_void, n: usize, out: *mut u8) -> c_int;

    pub fn sk_pcm_flat_convert(ctx: *mut sk_ctx, op: c_int, input: *const u8, bytes: usize, out: *mut c_void, n_out: *mut usize) -> c_int;
    pub fn sk_pcm_flat_convert_dev(ctx: *mut sk_ctx, op: c_int, input: *const u8, bytes: usize, out: *mut c_void) -> c_int;
    pub fn sk_flat_op_widths(op: c_int, in_width: *mut c_int, out_width: *mut c_int) -> c_int;
    pub fn sk_deinterleave(ctx: *mut sk_ctx, kind: c_int, input: *const u8, bytes: usize, channels: c_int, out: *mut c_void, frames: *mut usize) -> c_int;
    pub fn sk_interleave_i16(ctx: *mut sk_ctx, planar: *const i16, channels: c_int, frames: usize, out: *mut u8) -> c_int;
    pub fn sk_interleave_planar(ctx: *mut sk_ctx, elem_bytes: c_int, planar: *const c_void, channels: c_int, frames: usize, out: *mut u8) -> c_int;
    pub fn sk_interleave_planar_dev(ctx: *mut sk_ctx, elem_bytes: c_int, planar: *const c_void, channels: c_int, frames: usize, out: *mut u8) -> c_int;
    pub fn sk_stereo_to_mono_i16(ctx: *mut sk_ctx, avg: c_int, input: *const i16, n: usize, out: *mut i16) -> c_int;
    pub fn sk_pcm_to_planar_f32(ctx: *mut sk_ctx, input: *const u8, bytes: usize, bits: c_int, fmt: c_int, big_endian: c_int, channels: c_int,
                                scale_mode: c_int, out: *mut c_float, frames_cap: usize, frames: *mut usize) -> c_int;
    pub fn sk_pcm_to_planar_f32_dev(ctx: *mut sk_ctx, input: *const u8, in_stride: usize, n_streams: usize, frames: usize, bits: c_int, fmt: c_int,
                                    big_endian: c_int, channels: c_int, scale_mode: c_int, out: *mut c_float, out_stride: usize) -> c_int;
    pub fn sk_planar_f32_to_pcm(ctx: *mut sk_ctx, planar: *const c_float, frames: usize, channels: c_int, bits: c_int, fmt: c_int, out: *mut u8) -> c_int;
    pub fn sk_planar_f32_to_pcm_dev(ctx: *mut sk_ctx, planar: *const c_float, frames: usize, channels: c_int, bits: c_int, fmt: c_int, out: *mut u8) -> c_int;
    pub fn sk_pcm_narrow_i16(ctx: *mut sk_ctx, input: *const u8, bytes: usize, bits: c_int, big_endian: c_int, out: *mut u8) -> c_int;
    pub fn sk_mixdown_mono(ctx: *mut sk_ctx, planar: *const c_float, frames: usize, channels: c_int, mix_mode: c_int, out: *mut c_float) -> c_int;
    pub fn sk_mixdown_mono_dev(ctx: *mut sk_ctx, planar: *const c_float, frames: usize, channels: c_int, mix_mode: c_int, out: *mut c_float) -> c_int;
    pub fn sk_downmix_stereo(ctx: *mut sk_ctx, planar: *const c_float, frames: usize, channels: c_int, out: *mut c_float) -> c_int;

    pub fn sk_resampler_create(ctx: *mut sk_ctx, in_rate: u32, out_rate: u32, channels: c_int, chunk: usize, params: *const sk_resample_params,
                               out: *mut *mut sk_resampler) -> c_int;
    pub fn sk_resampler_destroy(r: *mut sk_resampler);
    pub fn sk_resampler_out_max(r: *const sk_resampler) -> usize;
    pub fn sk_resampler_last_index(r: *const sk_resampler) -> c_double;
    pub fn sk_resampler_process(r: *mut sk_resampler, input: *const *const c_float, out: *const *mut c_float, n_out: *mut usize) -> c_int;
    pub fn sk_resampler_process_partial(r: *mut sk_resampler, input_or_null: *const *const c_float, frames: usize, out: *const *mut c_float,
                                        n_out: *mut usize) -> c_int;
    pub fn sk_resampler_copy_table(r: *const sk_resampler, out65536: *mut c_float) -> c_int;

    // N StreamingResamplers in lock step, device-resident state (soundkit-decoder/src/lib.rs:1917-2060)
    pub fn sk_resampler_batch_create(ctx: *mut sk_ctx, in_rate: u32, out_rate: u32, channels: c_int, n_streams: usize, max_push_frames: usize,
                                     out: *mut *mut sk_resampler_batch) -> c_int;
    pub fn sk_resampler_batch_destroy(b: *mut sk_resampler_batch);
    pub fn sk_resampler_batch_out_cap(b: *const sk_resampler_batch) -> usize;
    pub fn sk_resampler_batch_pending(b: *const sk_resampler_batch) -> usize;
    pub fn sk_resampler_batch_last_index(b: *const sk_resampler_batch) -> c_double;
    pub fn sk_resampler_batch_process_dev(b: *mut sk_resampler_batch, input: *const c_float, in_stream_stride: usize, in_ch_stride: usize, frames: usize,
                                          out: *mut c_float, out_stream_stride: usize, out_ch_stride: usize, n_out: *mut usize) -> c_int;
    pub fn sk_resampler_batch_input_dev(b: *mut sk_resampler_batch, input: *mut *mut c_float, lane_pitch: *mut usize) -> c_int;
    pub fn sk_resampler_batch_flush_dev(b: *mut sk_resampler_batch, out: *mut c_float, out_stream_stride: usize, out_ch_stride: usize,
                                        n_out: *mut usize) -> c_int;

    pub fn sk_normalize_batch_plan(ctx: *mut sk_ctx, in_rate: u32, out_rate: u32, frames: usize, semantics: c_int, n_out: *mut usize) -> c_int;
    pub fn sk_normalize_batch_dev(ctx: *mut sk_ctx, desc: *const sk_batch_desc) -> c_int;
    pub fn sk_normalize_batch(ctx: *mut sk_ctx, desc: *const sk_batch_desc) -> c_int;
    pub fn sk_normalize_batch_pcm_dev(ctx: *mut sk_ctx, desc: *const sk_batch_desc, out_bits: c_int, out_fmt: c_int, out: *mut c_void,
                                      out_stride_bytes: usize) -> c_int;
    pub fn sk_normalize_stream(ctx: *mut sk_ctx, input: *const u8, bytes: usize, bits: c_int, fmt: c_int, big_endian: c_int, channels: c_int,
                               scale_mode: c_int, in_rate: u32, out_rate: u32, semantics: c_int, mix_mode: c_int, out: *mut c_float,
                               out_cap_per_plane: usize, n_out: *mut usize) -> c_int;

    // unpack_dvd_lpcm (soundkit-decoder/src/lib.rs:750-804)
    pub fn sk_dvd_lpcm_unpack(ctx: *mut sk_ctx, packed: *const u8, bytes: usize, bits: c_int, channels: c_int, out: *mut u8) -> c_int;
    pub fn sk_dvd_lpcm_unpack_dev(ctx: *mut sk_ctx, packed: *const u8, bytes: usize, bits: c_int, channels: c_int, out: *mut u8) -> c_int;

    // G.726 ADPCM (soundkit-g726/src/lib.rs): whole code groups, explicit G726Core state
    pub fn sk_g726_state_init(state: *mut i32);
    pub fn sk_g726_decode(ctx: *mut sk_ctx, bits: c_int, packing: c_int, state: *mut i32, input: *const u8, bytes: usize, out_kind: c_int,
                          out: *mut c_void) -> c_int;
    pub fn sk_g726_encode(ctx: *mut sk_ctx, bits: c_int, packing: c_int, state: *mut i32, in_kind: c_int, input: *const c_void, n: usize,
                          out: *mut u8) -> c_int;
    pub fn sk_g726_decode_batch_dev(ctx: *mut sk_ctx, bits: c_int, packing: c_int, input: *const u8, in_stride: usize, n_streams: usize,
                                    bytes: usize, out_kind: c_int, out: *mut c_void, out_stride: usize, state: *mut i32) -> c_int;
    pub fn sk_g726_encode_batch_dev(ctx: *mut sk_ctx, bits: c_int, packing: c_int, in_kind: c_int, input: *const c_void, in_stride: usize,
                                    n_streams: usize, n: usize, out: *mut u8, out_stride: usize, state: *mut i32) -> c_int;

    // G.722 64 kbit/s (soundkit-g722/src/lib.rs over ezk-g722): explicit 68-word codec state, all zero at reset
    pub fn sk_g722_state_init(state: *mut i32);
    pub fn sk_g722_decode(ctx: *mut sk_ctx, state: *mut i32, input: *const u8, bytes: usize, out_kind: c_int, out: *mut c_void) -> c_int;
    pub fn sk_g722_encode(ctx: *mut sk_ctx, state: *mut i32, in_kind: c_int, input: *const c_void, n: usize, out: *mut u8) -> c_int;
    pub fn sk_g722_decode_batch_dev(ctx: *mut sk_ctx, input: *const u8, in_stride: usize, n_streams: usize, bytes: usize, out_kind: c_int,
                                    out: *mut c_void, out_stride: usize, state: *mut i32) -> c_int;
    pub fn sk_g722_encode_batch_dev(ctx: *mut sk_ctx, in_kind: c_int, input: *const c_void, in_stride: usize, n_streams: usize, n: usize,
                                    out: *mut u8, out_stride: usize, state: *mut i32) -> c_int;

    // StreamingStereo48kNormalizer (soundkit/src/audio_pipeline.rs:256-394)
    pub fn sk_norm48k_create(ctx: *mut sk_ctx, out: *mut *mut sk_norm48k) -> c_int;
    pub fn sk_norm48k_destroy(n: *mut sk_norm48k);
    pub fn sk_norm48k_push(n: *mut sk_norm48k, data: *const u8, bytes: usize, bits: c_int, fmt: c_int, big_endian: c_int, channels: c_int,
                           rate: u32, left: *mut c_float, right: *mut c_float, cap: usize, n_out: *mut usize) -> c_int;
    pub fn sk_norm48k_finish(n: *mut sk_norm48k, left: *mut c_float, right: *mut c_float, cap: usize, n_out: *mut usize) -> c_int;
    pub fn sk_norm48k_source_frames(n: *const sk_norm48k) -> u64;
    pub fn sk_norm48k_output_frames(n: *const sk_norm48k) -> u64;
    pub fn sk_norm48k_source_sample_rate(n: *const sk_norm48k) -> u32;
    pub fn sk_norm48k_source_channels(n: *const sk_norm48k) -> c_int;
    pub fn sk_norm48k_max_output(frames: usize, rate: u32) -> usize;
    pub fn sk_normalize48k_out_frames(frames: usize, rate: u32) -> usize;
    pub fn sk_normalize48k_batch_dev(ctx: *mut sk_ctx, input: *const c_void, in_stride: usize, n_streams: usize, frames: usize, bits: c_int,
                                     fmt: c_int, big_endian: c_int, channels: c_int, rate: u32, out: *mut c_float, out_stride: usize,
                                     out_stride_ch: usize, n_out: *mut usize) -> c_int;
}
