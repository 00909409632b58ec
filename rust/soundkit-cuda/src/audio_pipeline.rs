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

// ---- the free functions of soundkit/src/audio_pipeline.rs:17-131, 415-436 ----------------------------------------------
fn flat_vec<I, O: Default + Clone>(op: sys::sk_flat_op, input: &[I]) -> Vec<O> {
    let mut out = vec![O::default(); input.len()];
    if input.is_empty() { return out; }
    let mut n = 0usize;
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_pcm_flat_convert(ctx, op as i32, input.as_ptr().cast(), std::mem::size_of_val(input), out.as_mut_ptr().cast(), &mut n)
    })).expect("soundkit-cuda flat conversion");
    out.truncate(n);
    out
}
/// `(clamp(x, -1, 1) * 32767.0) as i16` (:17-27) == f32le_to_i16 on the samples' own little-endian bytes
pub fn vec_f32_to_i16(input: Vec<f32>) -> Vec<i16> { flat_vec(sys::sk_flat_op::F32LE_TO_I16, &input) }
/// `value as f32 / 32768.0` (:29-38)
pub fn vec_i16_to_f32(input: Vec<i16>) -> Vec<f32> { flat_vec(sys::sk_flat_op::I16LE_TO_F32, &input) }
/// `value as f32 / 2147483647.0` — the literal rounds to 2^31 in f32 (:40-50)
pub fn vec_i32_to_f32(input: Vec<i32>) -> Vec<f32> { flat_vec(sys::sk_flat_op::S32LE_TO_F32, &input) }

/// `audio_to_f32_channels` (:74-98), soundkit convention ("mode A": always little-endian, 24-bit / 2^31, no NaN scrub).
/// Takes the fields of `AudioData` so the crate does not depend on soundkit; the caller passes `audio.data()` etc.
pub fn audio_to_f32_channels(bits_per_sample: u8, channel_count: u8, is_float: bool, data: &[u8]) -> Result<Vec<Vec<f32>>, String> {
    let channels = channel_count as usize;
    if channels == 0 { return Err("Channel count must be > 0".to_string()); }
    let fmt = if bits_per_sample == 32 && is_float { sys::SK_FMT_FLOAT } else { sys::SK_FMT_SIGNED };
    if !(bits_per_sample == 16 || bits_per_sample == 24 || bits_per_sample == 32) {
        return Err("deserialize_audio failed: unsuporrted type".to_string());      // sic, :71 and :91
    }
    let frames = data.len() / (bits_per_sample as usize / 8 * channels);
    let mut flat = vec![0f32; frames.max(1) * channels];
    let mut got = 0usize;
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_pcm_to_planar_f32(ctx, data.as_ptr(), data.len(), bits_per_sample as i32, fmt, 0, channels as i32, sys::SK_MODE_A,
                                  flat.as_mut_ptr(), frames.max(1), &mut got)
    }))?;
    Ok((0..channels).map(|c| flat[c * frames.max(1)..c * frames.max(1) + got].to_vec()).collect())
}

/// `mixdown_to_mono_f32` (:105-131): `(((0 + c0) + c1) + ...) * (1 / n)`
pub fn mixdown_to_mono_f32(channels: &[Vec<f32>]) -> Result<Vec<f32>, String> {
    if channels.is_empty() { return Ok(Vec::new()); }
    if channels.len() == 1 { return Ok(channels[0].clone()); }
    let sample_count = channels[0].len();
    if channels.iter().any(|channel| channel.len() != sample_count) { return Err("channel length mismatch".to_string()); }
    if sample_count == 0 { return Ok(Vec::new()); }
    let planar: Vec<f32> = channels.iter().flat_map(|c| c.iter().copied()).collect();
    let mut mono = vec![0f32; sample_count];
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_mixdown_mono(ctx, planar.as_ptr(), sample_count, channels.len() as i32, sys::SK_MIX_SUM_THEN_SCALE, mono.as_mut_ptr())
    }))?;
    Ok(mono)
}
/// `audio_to_mono_f32` (:100-103)
pub fn audio_to_mono_f32(bits_per_sample: u8, channel_count: u8, is_float: bool, data: &[u8]) -> Result<Vec<f32>, String> {
    mixdown_to_mono_f32(&audio_to_f32_channels(bits_per_sample, channel_count, is_float, data)?)
}

/// `f32s_to_le_bytes` / `f32s_from_le_bytes` (:415-436): bit copies on a little-endian host — no arithmetic to offload
pub fn f32s_to_le_bytes(samples: &[f32]) -> Vec<u8> {
    let mut bytes = Vec::with_capacity(std::mem::size_of_val(samples));
    for sample in samples { bytes.extend_from_slice(&sample.to_le_bytes()); }
    bytes
}
pub fn f32s_from_le_bytes(bytes: &[u8]) -> Result<Vec<f32>, String> {
    if bytes.len() % std::mem::size_of::<f32>() != 0 {
        return Err(format!("invalid f32le byte length {}; expected multiple of 4", bytes.len()));
    }
    Ok(bytes.chunks_exact(4).map(|c| f32::from_le_bytes([c[0], c[1], c[2], c[3]])).collect())
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
