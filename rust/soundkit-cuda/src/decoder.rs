//! The output-conversion stage of `soundkit-decoder` (soundkit-decoder/src/lib.rs:750-804, 1815-2060, 3458-3701) on the
//! device. Same names, argument meaning and error strings as the private functions they replace; the bookkeeping that is
//! not arithmetic (`accum`, the flush trim, every validation string) stays in Rust.  SOURCE ONLY — the compiled and
//! tested twin is namespace `soundkit_decoder` in `soundkit_b200/csrc/host/soundkit_host.hpp`.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

pub const RESAMPLE_CHUNK_SIZE: usize = 4096; // soundkit-decoder/src/lib.rs:79

fn flatten(channels: &[Vec<f32>], frames: usize) -> Vec<f32> {
    let mut planar = Vec::with_capacity(channels.len() * frames);
    for c in channels { planar.extend_from_slice(&c[..frames]); }
    planar
}
fn split_planes(flat: Vec<f32>, channels: usize, frames: usize) -> Vec<Vec<f32>> {
    (0..channels).map(|c| flat[c * frames..(c + 1) * frames].to_vec()).collect()
}

/// `audio_data_to_f32_channels` (:3563-3617): endian-aware, full-scale divisors 2^15 / 2^23 / 2^31, non-finite floats → 0.0
pub fn audio_data_to_f32_channels(bits: u8, channel_count: u8, is_float: bool, big_endian: bool, data: &[u8]) -> Result<Vec<Vec<f32>>, String> {
    let ch = channel_count as usize;
    if ch == 0 { return Err("Channel count must be > 0".to_string()); }
    let bps = (bits as usize + 7) / 8;
    if !matches!(bits, 16 | 24 | 32) || data.len() % bps != 0 || (data.len() / bps) % ch != 0 {
        return Err("PCM data is unsupported or contains a partial frame".to_string());
    }
    if is_float && bits != 32 { return Err("floating-point PCM must contain 32-bit samples".to_string()); }
    let frames = data.len() / bps / ch;
    let mut flat = vec![0f32; ch * frames];
    let mut got = 0usize;
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_pcm_to_planar_f32(ctx, data.as_ptr(), data.len(), bits as i32, if is_float { sys::SK_FMT_FLOAT } else { sys::SK_FMT_SIGNED },
                                  big_endian as i32, ch as i32, sys::SK_MODE_B, flat.as_mut_ptr(), frames, &mut got)
    }))?;
    Ok(split_planes(flat, ch, frames))
}

/// `exact_signed_pcm_to_i16` (:3458-3489): 24/32-bit signed LE/BE → s16 LE bytes by an arithmetic shift
pub fn exact_signed_pcm_to_i16(bits: u8, big_endian: bool, data: &[u8]) -> Result<Vec<u8>, String> {
    let bps = bits as usize / 8;
    if bps == 0 || data.len() % bps != 0 { return Err("integer PCM contains a partial sample".to_string()); }
    let mut out = vec![0u8; data.len() / bps * 2];
    with_ctx(|ctx| check(ctx, unsafe { sys::sk_pcm_narrow_i16(ctx, data.as_ptr(), data.len(), bits as i32, big_endian as i32, out.as_mut_ptr()) }))?;
    Ok(out)
}

/// `downmix_channels` (:3492-3561): → mono by scale-then-sum; → stereo from > 2 channels with the 0.707 matrix and the
/// block-wide peak normalisation; otherwise the first N channels
pub fn downmix_channels(channels: &[Vec<f32>], target_channels: u8) -> Vec<Vec<f32>> {
    if channels.is_empty() || target_channels == 0 { return Vec::new(); }
    let n = channels[0].len();
    let gpu = |stereo: bool| -> Result<Vec<Vec<f32>>, String> {
        let planar = flatten(channels, n);
        let planes = if stereo { 2 } else { 1 };
        let mut out = vec![0f32; planes * n];
        with_ctx(|ctx| check(ctx, unsafe {
            if stereo { sys::sk_downmix_stereo(ctx, planar.as_ptr(), n, channels.len() as i32, out.as_mut_ptr()) }
            else { sys::sk_mixdown_mono(ctx, planar.as_ptr(), n, channels.len() as i32, sys::SK_MIX_SCALE_THEN_SUM, out.as_mut_ptr()) }
        }))?;
        Ok(split_planes(out, planes, n))
    };
    // the reference function is infallible; a device failure here is a broken installation, not a data error
    if target_channels == 1 { return gpu(false).expect("soundkit-cuda: mixdown failed"); }
    if target_channels == 2 && channels.len() > 2 { return gpu(true).expect("soundkit-cuda: stereo downmix failed"); }
    channels.iter().take(target_channels as usize).cloned().collect()
}

/// `f32_channels_to_bytes` (:3619-3683) incl. `float_sample_to_i16` (:1815-1827) and `interleave_vecs_f32` (:3685-3701)
pub fn f32_channels_to_bytes(channels: &[Vec<f32>], bits_per_sample: u8, pcm_float: bool) -> Result<Vec<u8>, String> {
    if channels.is_empty() { return Ok(Vec::new()); }
    let n = channels[0].len();
    if channels.iter().any(|c| c.len() != n) { return Err("Channel length mismatch".to_string()); }
    if pcm_float && bits_per_sample != 32 { return Err("PCMFloat output requires 32-bit samples".to_string()); }
    if !pcm_float && !matches!(bits_per_sample, 16 | 24 | 32) { return Err(format!("Unsupported output bits per sample: {bits_per_sample}")); }
    let planar = flatten(channels, n);
    let mut out = vec![0u8; n * channels.len() * (bits_per_sample as usize / 8)];
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_planar_f32_to_pcm(ctx, planar.as_ptr(), n, channels.len() as i32, bits_per_sample as i32,
                                  if pcm_float { sys::SK_FMT_FLOAT } else { sys::SK_FMT_SIGNED }, out.as_mut_ptr())
    }))?;
    Ok(out)
}

/// `unpack_dvd_lpcm` (:750-804): 16-bit = byte swap, 20/24-bit = the DVD group layout → s24/s32 LE as the reference emits
/// The output has exactly the input's byte length (the C side writes `packed.len()` bytes): the buffer is sized here, not by
/// the caller, so a wrong size can never reach the device.
pub fn unpack_dvd_lpcm(packed: &[u8], bits: u8, channels: u8) -> Result<Vec<u8>, String> {
    let mut out = vec![0u8; packed.len()];
    if packed.is_empty() { return Ok(out); }
    with_ctx(|ctx| check(ctx, unsafe { sys::sk_dvd_lpcm_unpack(ctx, packed.as_ptr(), packed.len(), bits as i32, channels as i32, out.as_mut_ptr()) }))?;
    Ok(out)
}

/// `StreamingResampler` (:1917-2060). rubato's `SincFixedIn` state (512-sample history per channel + the f64 index) lives
/// on the device inside `sk_resampler`; `accum`, the compaction rule and the flush trim are the reference's.
pub struct StreamingResampler {
    handle: *mut sys::sk_resampler,
    out_max: usize,
    accum: Vec<Vec<f32>>,
    accum_start: usize,
    pub chunk_size: usize,
    pub channels: usize,
    pub input_sample_rate: u32,
    pub output_sample_rate: u32,
}
impl Drop for StreamingResampler { fn drop(&mut self) { unsafe { sys::sk_resampler_destroy(self.handle) } } }

impl StreamingResampler {
    pub fn new(input_sample_rate: u32, output_sample_rate: u32, channels: usize) -> Result<Self, String> {
        with_ctx(|ctx| {
            let mut h = std::ptr::null_mut();
            check(ctx, unsafe { sys::sk_resampler_create(ctx, input_sample_rate, output_sample_rate, channels as i32, RESAMPLE_CHUNK_SIZE, std::ptr::null(), &mut h) })
                .map_err(|e| format!("Failed to create resampler: {e}"))?;
            Ok(Self { handle: h, out_max: unsafe { sys::sk_resampler_out_max(h) }, accum: vec![Vec::new(); channels], accum_start: 0,
                      chunk_size: RESAMPLE_CHUNK_SIZE, channels, input_sample_rate, output_sample_rate })
        })
    }

    fn run(&mut self, start: usize, frames: usize, partial: bool) -> Result<Vec<Vec<f32>>, String> {
        let ip: Vec<*const f32> = self.accum.iter().map(|c| unsafe { c.as_ptr().add(start) }).collect();
        let mut out = vec![vec![0f32; self.out_max]; self.channels];
        let op: Vec<*mut f32> = out.iter_mut().map(|c| c.as_mut_ptr()).collect();
        let mut n = 0usize;
        let h = self.handle;
        with_ctx(|ctx| check(ctx, unsafe {
            if partial { sys::sk_resampler_process_partial(h, if frames != 0 { ip.as_ptr() } else { std::ptr::null() }, frames, op.as_ptr(), &mut n) }
            else { sys::sk_resampler_process(h, ip.as_ptr(), op.as_ptr(), &mut n) }
        })).map_err(|e| format!("{}{e}", if partial { "Resample partial failed: " } else { "Resample failed: " }))?;
        for c in &mut out { c.truncate(n); }
        Ok(out)
    }

    pub fn process(&mut self, input: &[Vec<f32>]) -> Result<Vec<Vec<Vec<f32>>>, String> {
        if input.len() != self.channels {
            return Err(format!("Channel count changed mid-stream: expected {}, got {}", self.channels, input.len()));
        }
        if input.iter().any(|c| c.len() != input[0].len()) { return Err("Channel sample counts changed mid-stream".to_string()); }
        for (acc, c) in self.accum.iter_mut().zip(input) { acc.extend_from_slice(c); }
        let mut outputs = Vec::new();
        while self.accum[0].len() - self.accum_start >= self.chunk_size {
            let res = self.run(self.accum_start, self.chunk_size, false)?;
            self.accum_start += self.chunk_size;
            if !res.is_empty() && !res[0].is_empty() { outputs.push(res); }
        }
        if self.accum_start >= self.chunk_size * 8 && self.accum_start * 2 >= self.accum[0].len() {
            let s = self.accum_start;
            for c in &mut self.accum { c.drain(..s); }
            self.accum_start = 0;
        }
        Ok(outputs)
    }

    pub fn flush(&mut self) -> Result<Vec<Vec<Vec<f32>>>, String> {
        let mut outputs = Vec::new();
        let remaining = self.accum[0].len() - self.accum_start;
        if remaining > 0 {
            let padded = self.chunk_size - remaining;
            let mut res = self.run(self.accum_start, remaining, true)?;
            if padded > 0 {
                let trim = ((padded as f64 * self.output_sample_rate as f64) / self.input_sample_rate as f64).round() as usize;
                for c in &mut res { let keep = c.len().saturating_sub(trim); c.truncate(keep); }
            }
            if !res.is_empty() && !res[0].is_empty() { outputs.push(res); }
            for c in &mut self.accum { c.clear(); }
            self.accum_start = 0;
        } else {
            let res = self.run(0, 0, true)?;
            if !res.is_empty() && !res[0].is_empty() { outputs.push(res); }
        }
        Ok(outputs)
    }
}


/// N `StreamingResampler`s in lock step (`sk_resampler_batch_*`): the per-stream rubato history and the pending input stay on the
/// device between pushes, every push of the whole batch is one kernel launch.  `input` / `output` are DEVICE pointers to planar
/// f32 (`[stream][channel][frame]`, strides in floats) owned by the caller — a serving front-end keeps its streams resident.
pub struct StreamingResamplerBatch { handle: *mut sys::sk_resampler_batch, pub out_cap: usize }
impl Drop for StreamingResamplerBatch { fn drop(&mut self) { unsafe { sys::sk_resampler_batch_destroy(self.handle) } } }
impl StreamingResamplerBatch {
    pub fn new(input_sample_rate: u32, output_sample_rate: u32, channels: usize, n_streams: usize, max_push_frames: usize) -> Result<Self, String> {
        if input_sample_rate == 0 { return Err("Input sample rate must be > 0".to_string()); }       // apply_output_options' strings (:3330-3370)
        if output_sample_rate == 0 { return Err("Output sample rate must be > 0".to_string()); }
        if channels == 0 { return Err("Channel count must be > 0".to_string()); }
        with_ctx(|ctx| {
            let mut h = std::ptr::null_mut();
            check(ctx, unsafe {
                sys::sk_resampler_batch_create(ctx, input_sample_rate, output_sample_rate, channels as i32, n_streams, max_push_frames, &mut h)
            }).map_err(|e| format!("Failed to create resampler: {e}"))?;
            Ok(Self { handle: h, out_cap: unsafe { sys::sk_resampler_batch_out_cap(h) } })
        })
    }
    /// # Safety
    /// `input` must address `frames` readable floats per (stream, channel) at the given strides and `output` `out_cap` writable
    /// floats per (stream, channel), both in device memory of the context's device.
    pub unsafe fn process_dev(&mut self, input: *const f32, in_stream_stride: usize, in_ch_stride: usize, frames: usize, output: *mut f32,
                              out_stream_stride: usize, out_ch_stride: usize) -> Result<usize, String> {
        let mut n = 0usize;
        with_ctx(|ctx| check(ctx, sys::sk_resampler_batch_process_dev(self.handle, input, in_stream_stride, in_ch_stride, frames, output,
                                                                         out_stream_stride, out_ch_stride, &mut n))
            .map_err(|e| format!("Resample failed: {e}")))?;
        Ok(n)
    }
    /// Where a device-side producer writes the next push (zero-copy append): device address of lane (stream 0, channel 0) and the
    /// pitch in floats between (stream, channel) lanes; follow it with `process_dev(std::ptr::null(), 0, 0, frames, ..)`.
    pub fn input_dev(&mut self) -> Result<(*mut f32, usize), String> {
        let (mut p, mut pitch) = (std::ptr::null_mut(), 0usize);
        with_ctx(|ctx| check(ctx, unsafe { sys::sk_resampler_batch_input_dev(self.handle, &mut p, &mut pitch) }))?;
        Ok((p, pitch))
    }
    /// # Safety
    /// as `process_dev`
    pub unsafe fn flush_dev(&mut self, output: *mut f32, out_stream_stride: usize, out_ch_stride: usize) -> Result<usize, String> {
        let mut n = 0usize;
        with_ctx(|ctx| check(ctx, sys::sk_resampler_batch_flush_dev(self.handle, output, out_stream_stride, out_ch_stride, &mut n))
            .map_err(|e| format!("Resample flush failed: {e}")))?;
        Ok(n)
    }
}
