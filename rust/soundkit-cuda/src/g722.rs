//! `soundkit_g722` (soundkit-g722/src/lib.rs): G.722 64 kbit/s with the sub-band ADPCM codec on the device.  The odd pending
//! sample of the encoder, the buffer checks and their error strings stay here, exactly as in the reference.
//! SOURCE ONLY (see lib.rs); mirrors `soundkit_g722` in soundkit_host.hpp, which is compiled and tested.
use crate::audio_packet::{Decoder, Encoder};
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

pub const G722_SAMPLE_RATE: u32 = 16_000;                                         // :6
pub const G722_CHANNELS: u8 = 1;                                                  // :7

fn reset_state() -> [i32; sys::SK_G722_STATE_INTS] {
    let mut s = [0i32; sys::SK_G722_STATE_INTS];
    unsafe { sys::sk_g722_state_init(s.as_mut_ptr()) };
    s
}
fn max_encoded_len(input_samples: usize) -> usize { input_samples.div_ceil(2) }   // :234-236

pub struct G722Encoder { state: [i32; sys::SK_G722_STATE_INTS], pending_sample: Option<i16> }   // :9-12
impl G722Encoder {
    pub fn new_64k() -> Self { Self { state: reset_state(), pending_sample: None } }             // :15
    fn encode_pairs(&mut self, samples: &[i16], output: &mut [u8]) -> Result<usize, String> {
        debug_assert!(samples.len() % 2 == 0 && output.len() >= samples.len() / 2);
        if samples.is_empty() { return Ok(0); }
        with_ctx(|ctx| check(ctx, unsafe {
            sys::sk_g722_encode(ctx, self.state.as_mut_ptr(), sys::SK_KIND_I16, samples.as_ptr().cast(), samples.len(), output.as_mut_ptr())
        }))?;
        Ok(samples.len() / 2)
    }
    pub fn encode_to_vec(&mut self, input: &[i16], output: &mut Vec<u8>) -> Result<usize, String> {   // :23-31
        let start = output.len();
        let mut scratch = vec![0u8; max_encoded_len(input.len() + self.pending_sample.is_some() as usize)];
        let written = self.encode_i16(input, &mut scratch)?;
        output.extend_from_slice(&scratch[..written]);
        Ok(output.len() - start)
    }
    pub fn flush_into(&mut self, output: &mut [u8]) -> Result<usize, String> {                         // :33-46
        let Some(sample) = self.pending_sample.take() else { return Ok(0) };
        if output.is_empty() {
            self.pending_sample = Some(sample);
            return Err("Output buffer too small for G.722 flush: need 1, have 0".to_string());
        }
        self.encode_pairs(&[sample, 0], &mut output[..1])
    }
    pub fn flush_to_vec(&mut self, output: &mut Vec<u8>) -> Result<usize, String> {                    // :48-54
        let start = output.len();
        output.resize(start + 1, 0);
        let written = self.flush_into(&mut output[start..])?;
        output.truncate(start + written);
        Ok(written)
    }
}
impl Default for G722Encoder { fn default() -> Self { Self::new_64k() } }
impl Encoder for G722Encoder {                                                                         // :63-129
    fn new(_sample_rate: u32, _bits_per_sample: u32, _channels: u32, _frame_size: u32, _bitrate: u32) -> Self { Self::new_64k() }
    fn init(&mut self) -> Result<(), String> { Ok(()) }
    fn encode_i16(&mut self, input: &[i16], output: &mut [u8]) -> Result<usize, String> {              // :83-117
        let total_samples = input.len() + self.pending_sample.is_some() as usize;
        let complete_samples = total_samples - (total_samples % 2);
        let required = complete_samples / 2;
        if output.len() < required {
            return Err(format!("Output buffer too small for G.722 encode: need {}, have {}", required, output.len()));
        }
        if total_samples < 2 {
            if self.pending_sample.is_none() { self.pending_sample = input.first().copied(); }
            return Ok(0);
        }
        let mut samples = Vec::with_capacity(complete_samples);
        if let Some(sample) = self.pending_sample.take() { samples.push(sample); }
        samples.extend_from_slice(input);
        if samples.len() % 2 != 0 { self.pending_sample = samples.pop(); }
        self.encode_pairs(&samples, &mut output[..required])
    }
    fn encode_i32(&mut self, input: &[i32], output: &mut [u8]) -> Result<usize, String> {              // :119-122
        let samples: Vec<i16> = input.iter().map(|sample| (sample >> 16) as i16).collect();
        self.encode_i16(&samples, output)
    }
    fn reset(&mut self) -> Result<(), String> { self.state = reset_state(); self.pending_sample = None; Ok(()) }   // :124-128
}

pub struct G722Decoder { state: [i32; sys::SK_G722_STATE_INTS] }                                       // :131-133
impl G722Decoder {
    pub fn new_64k() -> Self { Self { state: reset_state() } }
    pub fn sample_rate(&self) -> u32 { G722_SAMPLE_RATE }
    pub fn channels(&self) -> u8 { G722_CHANNELS }
    fn run<T>(&mut self, input: &[u8], output: &mut [T], kind: i32) -> Result<usize, String> {
        let required = input.len() * 2;
        if output.len() < required {
            return Err(format!("Output buffer too small for G.722 decode: need {}, have {}", required, output.len()));
        }
        if input.is_empty() { return Ok(0); }
        with_ctx(|ctx| check(ctx, unsafe {
            sys::sk_g722_decode(ctx, self.state.as_mut_ptr(), input.as_ptr(), input.len(), kind, output.as_mut_ptr().cast())
        }))?;
        Ok(required)
    }
}
impl Default for G722Decoder { fn default() -> Self { Self::new_64k() } }
impl Decoder for G722Decoder {                                                                         // :161-232
    fn decode_i16(&mut self, input: &[u8], output: &mut [i16], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_I16) }
    fn decode_i32(&mut self, input: &[u8], output: &mut [i32], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_I32) }
    fn decode_f32(&mut self, input: &[u8], output: &mut [f32], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_F32) }
}
