//! `soundkit::audio_pipeline::downsample_audio` (soundkit/src/audio_pipeline.rs:438-493): the fused
//! convert + one-shot sinc resample call.  `AudioData` is the reference's own type.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

const COMMON_SAMPLE_RATES: [u32; 9] = [8000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000];
const COMMON_BITS_PER_SAMPLE: [u8; 3] = [16, 24, 32];

pub fn downsample_audio(bits_per_sample: u8, channel_count: u8, input_rate: u32, is_float: bool, data: &[u8], sampling_rate: usize)
    -> Result<Vec<Vec<f32>>, String> {
    let channels = channel_count as usize;
    if channels == 0 { return Err("Channel count must be > 0".to_string()); }
    if !COMMON_BITS_PER_SAMPLE.contains(&bits_per_sample) { return Err(format!("Unsupported bits_per_sample: {}", bits_per_sample)); }
    let output_rate = u32::try_from(sampling_rate).map_err(|_| "sampling_rate out of range".to_string())?;
    if input_rate == 0 || output_rate == 0 { return Err("sampling_rate must be > 0".to_string()); }
    if !COMMON_SAMPLE_RATES.contains(&input_rate) { return Err(format!("Unsupported input sample_rate: {}", input_rate)); }
    if !COMMON_SAMPLE_RATES.contains(&output_rate) { return Err(format!("Unsupported output sample_rate: {}", output_rate)); }
    let frames = data.len() / (bits_per_sample as usize / 8 * channels);
    with_ctx(|ctx| {
        let mut n = 0usize;
        check(ctx, unsafe { sys::sk_normalize_batch_plan(ctx, input_rate, output_rate, frames, sys::SK_SEM_ONESHOT, &mut n) })?;
        let cap = n.max(1);
        let mut flat = vec![0f32; cap * channels];
        let mut got = 0usize;
        let fmt = if bits_per_sample == 32 && is_float { sys::SK_FMT_FLOAT } else { sys::SK_FMT_SIGNED };
        check(ctx, unsafe {
            sys::sk_normalize_stream(ctx, data.as_ptr(), data.len(), bits_per_sample as i32, fmt, 0, channels as i32, sys::SK_MODE_A, input_rate,
                                     output_rate, sys::SK_SEM_ONESHOT, sys::SK_MIX_NONE, flat.as_mut_ptr(), cap, &mut got)
        })?;
        Ok((0..channels).map(|c| flat[c * cap..c * cap + got].to_vec()).collect())
    })
}

/// One bounded block of canonical 48 kHz stereo PCM (soundkit/src/audio_pipeline.rs:234-248).
#[derive(Debug)]
pub struct Stereo48kBlock { pub left: Vec<f32>, pub right: Vec<f32> }
impl Stereo48kBlock {
    pub fn frame_count(&self) -> usize { self.left.len() }
    pub fn is_empty(&self) -> bool { self.left.is_empty() }
}

/// `StreamingStereo48kNormalizer` (soundkit/src/audio_pipeline.rs:256-394) with the conversion and the interpolation on
/// the device (`sk_norm48k`).  The geometry / depth / frame checks and their strings stay here, in the reference's order.
pub struct StreamingStereo48kNormalizer {
    handle: *mut sys::sk_norm48k,
    source_sample_rate: u32,
    source_channels: u8,
    finished: bool,
}
impl Drop for StreamingStereo48kNormalizer { fn drop(&mut self) { unsafe { sys::sk_norm48k_destroy(self.handle) } } }

impl StreamingStereo48kNormalizer {
    pub fn new() -> Result<Self, String> {
        with_ctx(|ctx| {
            let mut h = std::ptr::null_mut();
            check(ctx, unsafe { sys::sk_norm48k_create(ctx, &mut h) })?;
            Ok(Self { handle: h, source_sample_rate: 0, source_channels: 0, finished: false })
        })
    }
    pub fn source_sample_rate(&self) -> u32 { self.source_sample_rate }
    pub fn source_channels(&self) -> u8 { self.source_channels }
    pub fn source_frames(&self) -> u64 { unsafe { sys::sk_norm48k_source_frames(self.handle) } }
    pub fn output_frames(&self) -> u64 { unsafe { sys::sk_norm48k_output_frames(self.handle) } }

    /// `push(&AudioData)`: bits / channels / rate / float / big-endian are the fields of the reference's `AudioData`
    pub fn push(&mut self, bits: u8, channel_count: u8, sample_rate: u32, is_float: bool, big_endian: bool, data: &[u8])
        -> Result<Option<Stereo48kBlock>, String> {
        if self.finished { return Err("48 kHz stereo normalizer is already finished".to_owned()); }
        if sample_rate == 0 || channel_count == 0 { return Err("decoded audio has invalid PCM geometry".to_owned()); }
        if self.source_sample_rate == 0 {
            self.source_sample_rate = sample_rate;
            self.source_channels = channel_count;
        } else if self.source_sample_rate != sample_rate || self.source_channels != channel_count {
            return Err(format!("decoded PCM geometry changed from {} Hz/{} ch to {sample_rate} Hz/{channel_count} ch",
                               self.source_sample_rate, self.source_channels));
        }
        let sample_bytes = match bits { 16 => 2usize, 24 => 3, 32 => 4, b => return Err(format!("unsupported PCM bit depth {b}")) };
        let frame_bytes = sample_bytes * channel_count as usize;
        if data.is_empty() || data.len() % frame_bytes != 0 { return Err("decoded audio contained an incomplete PCM frame".to_owned()); }
        let cap = unsafe { sys::sk_norm48k_max_output(data.len() / frame_bytes, sample_rate) };
        let (mut left, mut right) = (vec![0f32; cap], vec![0f32; cap]);
        let mut n = 0usize;
        let fmt = if is_float { sys::SK_FMT_FLOAT } else { sys::SK_FMT_SIGNED };
        let h = self.handle;
        with_ctx(|ctx| check(ctx, unsafe {
            sys::sk_norm48k_push(h, data.as_ptr(), data.len(), bits as i32, fmt, big_endian as i32, channel_count as i32, sample_rate,
                                 left.as_mut_ptr(), right.as_mut_ptr(), cap, &mut n)
        }))?;
        left.truncate(n);
        right.truncate(n);
        Ok((n != 0).then_some(Stereo48kBlock { left, right }))
    }

    /// `finish()`: the exact rounded 48 kHz duration, repeating the final sample
    pub fn finish(&mut self) -> Result<Option<Stereo48kBlock>, String> {
        if self.finished { return Err("48 kHz stereo normalizer is already finished".to_owned()); }
        self.finished = true;
        if self.source_frames() == 0 || self.source_sample_rate == 0 { return Err("decoded source contained no PCM".to_owned()); }
        let cap = (48_000 / self.source_sample_rate) as usize + 8;
        let (mut left, mut right) = (vec![0f32; cap], vec![0f32; cap]);
        let mut n = 0usize;
        let h = self.handle;
        with_ctx(|ctx| check(ctx, unsafe { sys::sk_norm48k_finish(h, left.as_mut_ptr(), right.as_mut_ptr(), cap, &mut n) }))?;
        left.truncate(n);
        right.truncate(n);
        Ok((n != 0).then_some(Stereo48kBlock { left, right }))
    }
}
