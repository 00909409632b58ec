//! The sample-moving half of `WavStreamEncoder::push_planar_{i16,i32,f32}` and `push` (soundkit/src/wav.rs:444-540):
//! planar `[channel * frames_per_channel + frame]` -> interleaved little-endian bytes, a bit copy done on the device.
//! `require_format`, `reserve_chunk` and `frames_written` bookkeeping stay in `WavStreamEncoder`; it calls these where the
//! reference has its nested `for frame { for channel { extend_from_slice(to_le_bytes) } }` loops.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

fn interleave<T>(planar: &[T], frames_per_channel: usize, channel_count: usize) -> Result<Vec<u8>, String> {
    let width = std::mem::size_of::<T>();
    debug_assert!(width == 2 || width == 4);
    if planar.len() != frames_per_channel * channel_count {
        return Err(format!("planar buffer holds {} samples, expected {} x {}", planar.len(), channel_count, frames_per_channel));
    }
    let mut output = vec![0u8; planar.len() * width];
    if planar.is_empty() { return Ok(output); }
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_interleave_planar(ctx, width as i32, planar.as_ptr().cast(), channel_count as i32, frames_per_channel, output.as_mut_ptr())
    }))?;
    Ok(output)
}
pub fn push_planar_i16(planar: &[i16], frames_per_channel: usize, channel_count: usize) -> Result<Vec<u8>, String> { interleave(planar, frames_per_channel, channel_count) }  // :444-460
pub fn push_planar_i32(planar: &[i32], frames_per_channel: usize, channel_count: usize) -> Result<Vec<u8>, String> { interleave(planar, frames_per_channel, channel_count) }  // :462-477
pub fn push_planar_f32(planar: &[f32], frames_per_channel: usize, channel_count: usize) -> Result<Vec<u8>, String> { interleave(planar, frames_per_channel, channel_count) }  // :479-494

/// `push(&PcmData)` (:496-540): `Vec<Vec<T>>` channels of equal length (validate_planar_channels stays with the caller)
pub fn push_channels<T: Copy>(channels: &[Vec<T>]) -> Result<Vec<u8>, String> {
    let frames = channels.first().map_or(0, |c| c.len());
    let planar: Vec<T> = channels.iter().flat_map(|c| c[..frames].iter().copied()).collect();
    interleave(&planar, frames, channels.len())
}
