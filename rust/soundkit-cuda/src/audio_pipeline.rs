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
