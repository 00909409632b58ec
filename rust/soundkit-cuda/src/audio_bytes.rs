//! `soundkit::audio_bytes` (soundkit/src/audio_bytes.rs) — same signatures, GPU arithmetic.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

fn flat<T: Default + Clone>(op: sys::sk_flat_op, in_width: usize, data: &[u8]) -> Vec<T> {
    let mut out = vec![T::default(); data.len() / in_width];
    let mut n = 0usize;
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_pcm_flat_convert(ctx, op as i32, data.as_ptr(), data.len(), out.as_mut_ptr() as *mut _, &mut n)
    })).expect("soundkit-cuda flat conversion");
    out.truncate(n);
    out
}

pub fn i16le_to_f32(bytes: &[u8]) -> Vec<f32> {                                   // :3
    assert!(bytes.len() % 2 == 0, "Bytes length must be a multiple of 2");
    flat(sys::sk_flat_op::I16LE_TO_F32, 2, bytes)
}
pub fn i16le_to_i16(bytes: &[u8]) -> Vec<i16> {                                   // :25
    assert!(bytes.len() % 2 == 0, "Bytes length must be a multiple of 2");
    flat(sys::sk_flat_op::S16LE_TO_I16, 2, bytes)
}
pub fn s24le_to_i32(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S24LE_TO_I32, 3, data) }  // :36
pub fn s24le_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S24LE_TO_I16, 3, data) }  // :51
pub fn s24be_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S24BE_TO_I16, 3, data) }  // :66
pub fn s32le_to_i32(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S32LE_TO_I32, 4, data) }  // :81
pub fn s32be_to_i32(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S32BE_TO_I32, 4, data) }  // :91
pub fn s32le_to_s24(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S32LE_TO_S24, 4, data) }  // :101
pub fn s32be_to_s24(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S32BE_TO_S24, 4, data) }  // :112
pub fn s32le_to_f32(data: &[u8]) -> Vec<f32> { flat(sys::sk_flat_op::S32LE_TO_F32, 4, data) }  // :123
pub fn s32be_to_f32(data: &[u8]) -> Vec<f32> { flat(sys::sk_flat_op::S32BE_TO_F32, 4, data) }  // :134
pub fn s32le_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S32LE_TO_I16, 4, data) }  // :145
pub fn s32be_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S32BE_TO_I16, 4, data) }  // :156
pub fn f32le_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::F32LE_TO_I16, 4, data) }  // :167
pub fn f32be_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::F32BE_TO_I16, 4, data) }  // :178
pub fn f32le_to_i32(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::F32LE_TO_I32, 4, data) }  // :189
pub fn f32le_to_s24(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::F32LE_TO_S24, 4, data) }  // :205
pub fn s16be_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S16BE_TO_I16, 2, data) }  // :222
pub fn s16le_to_i16(data: &[u8]) -> Vec<i16> { flat(sys::sk_flat_op::S16LE_TO_I16, 2, data) }  // :231
pub fn s16le_to_i32(data: &[u8]) -> Vec<i32> { flat(sys::sk_flat_op::S16LE_TO_I32, 2, data) }  // :240

fn deinterleave<T: Default + Clone>(kind: i32, width: usize, input: &[u8], channel_count: usize) -> Vec<Vec<T>> {
    let frames = input.len() / (channel_count * width);
    let mut flat = vec![T::default(); frames * channel_count];
    let mut got = 0usize;
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_deinterleave(ctx, kind, input.as_ptr(), input.len(), channel_count as i32, flat.as_mut_ptr() as *mut _, &mut got)
    })).expect("soundkit-cuda deinterleave");
    flat.chunks(frames.max(1)).take(channel_count).map(|c| c[..frames].to_vec()).collect()
}
pub fn deinterleave_vecs_i16(input: &[u8], channel_count: usize) -> Vec<Vec<i16>> { deinterleave(0, 2, input, channel_count) }  // :264
pub fn deinterleave_vecs_s24(input: &[u8], channel_count: usize) -> Vec<Vec<i32>> { deinterleave(1, 3, input, channel_count) }  // :280
pub fn deinterleave_vecs_f32(input: &[u8], channel_count: usize) -> Vec<Vec<f32>> { deinterleave(2, 4, input, channel_count) }  // :296

pub fn interleave_vecs_i16(channels: &[Vec<i16>]) -> Vec<u8> {                    // :250
    let frames = channels[0].len();
    let planar: Vec<i16> = channels.iter().flat_map(|c| c[..frames].iter().copied()).collect();
    let mut out = vec![0u8; planar.len() * 2];
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_interleave_i16(ctx, planar.as_ptr(), channels.len() as i32, frames, out.as_mut_ptr())
    })).expect("soundkit-cuda interleave");
    out
}

fn stereo_to_mono(input: &[i16], avg: bool) -> Vec<i16> {
    assert!(input.len() % 2 == 0, "Stereo buffer must contain an even number of samples");
    let mut out = vec![0i16; input.len() / 2];
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_stereo_to_mono_i16(ctx, avg as i32, input.as_ptr(), input.len(), out.as_mut_ptr())
    })).expect("soundkit-cuda stereo_to_mono");
    out
}
pub fn stereo_to_mono_take_left(input: &[i16]) -> Vec<i16> { stereo_to_mono(input, false) }  // :317
pub fn stereo_to_mono_avg(input: &[i16]) -> Vec<i16> { stereo_to_mono(input, true) }         // :344

// ---- the remaining items of soundkit/src/audio_bytes.rs -------------------------------------------------------------
/// `i16_to_i16le` (:17-23): on a little-endian host this is the identity on bytes — the one conversion of the module that
/// moves no arithmetic at all, so it stays a host copy (a PCIe round trip would only add latency).
pub fn i16_to_i16le(data: &[i16]) -> Vec<u8> {
    let mut bytes = Vec::with_capacity(data.len() * 2);
    for value in data { bytes.extend(&value.to_le_bytes()); }
    bytes
}
/// `s24le_to_i32_sample` (:312-315): a single sample, scalar by definition (callers that have a buffer use `s24le_to_i32`).
pub fn s24le_to_i32_sample(sample_bytes: [u8; 3]) -> i32 {
    let sample = i32::from_le_bytes([sample_bytes[0], sample_bytes[1], sample_bytes[2], 0]);
    (sample << 8) >> 8
}
/// `stereo_to_mono_inplace_take_left` (:331-342): device result copied back over the front half of the caller's buffer
pub fn stereo_to_mono_inplace_take_left(samples: &mut [i16]) -> &mut [i16] {
    let mono = stereo_to_mono(samples, false);
    let frames = mono.len();
    samples[..frames].copy_from_slice(&mono);
    &mut samples[..frames]
}
/// `stereo_to_mono_inplace_avg` (:360-373)
pub fn stereo_to_mono_inplace_avg(samples: &mut [i16]) -> &mut [i16] {
    let mono = stereo_to_mono(samples, true);
    let frames = mono.len();
    samples[..frames].copy_from_slice(&mono);
    &mut samples[..frames]
}
