//! `soundkit_g711` (soundkit-g711/src/lib.rs) — same items, same signatures, same error strings; the per-sample
//! arithmetic (:235-322) runs on the device.  SOURCE ONLY (see lib.rs); mirrors `soundkit_g711` in soundkit_host.hpp.
use crate::audio_packet::{Decoder, Encoder};
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum G711Law { MuLaw, ALaw }                                                 // :17
fn law_id(law: G711Law) -> i32 { match law { G711Law::MuLaw => sys::SK_LAW_MULAW, G711Law::ALaw => sys::SK_LAW_ALAW } }

fn too_small(what: &str, need: usize, have: usize) -> String { format!("Output buffer too small for G.711 {what}: need {need}, have {have}") }

fn decode_kind<T>(law: G711Law, input: &[u8], output: &mut [T], kind: i32) -> Result<usize, String> {
    if output.len() < input.len() { return Err(too_small("decode", input.len(), output.len())); }
    if input.is_empty() { return Ok(0); }
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g711_decode(ctx, law_id(law), input.as_ptr(), input.len(), kind, output.as_mut_ptr() as *mut _)
    }))?;
    Ok(input.len())
}
fn encode_kind<T>(law: G711Law, input: &[T], output: &mut [u8], kind: i32) -> Result<usize, String> {
    if output.len() < input.len() { return Err(too_small("encode", input.len(), output.len())); }
    if input.is_empty() { return Ok(0); }
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g711_encode(ctx, law_id(law), kind, input.as_ptr() as *const _, input.len(), output.as_mut_ptr())
    }))?;
    Ok(input.len())
}

pub fn encode_i16(law: G711Law, input: &[i16], output: &mut [u8]) -> Result<usize, String> { encode_kind(law, input, output, sys::SK_KIND_I16) }  // :36
pub fn decode_i16(law: G711Law, input: &[u8], output: &mut [i16]) -> Result<usize, String> { decode_kind(law, input, output, sys::SK_KIND_I16) }  // :52
pub fn encode_sample(law: G711Law, sample: i16) -> u8 { let mut o = [0u8]; encode_i16(law, &[sample], &mut o).unwrap(); o[0] }  // :22
pub fn decode_sample(law: G711Law, sample: u8) -> i16 { let mut o = [0i16]; decode_i16(law, &[sample], &mut o).unwrap(); o[0] } // :29

#[derive(Debug)]
pub struct G711Encoder { law: G711Law, sample_rate: u32, channels: u8 }          // :68-73
impl G711Encoder {
    pub fn new_with_law(law: G711Law, sample_rate: u32, channels: u8) -> Self { Self { law, sample_rate, channels } }   // :76
    pub fn new_mulaw(sample_rate: u32, channels: u8) -> Self { Self::new_with_law(G711Law::MuLaw, sample_rate, channels) }
    pub fn new_alaw(sample_rate: u32, channels: u8) -> Self { Self::new_with_law(G711Law::ALaw, sample_rate, channels) }
    pub fn law(&self) -> G711Law { self.law }
    pub fn sample_rate(&self) -> u32 { self.sample_rate }
    pub fn channels(&self) -> u8 { self.channels }
}
impl Encoder for G711Encoder {                                                    // :105-143
    fn new(sample_rate: u32, _bits_per_sample: u32, channels: u32, _frame_size: u32, _bitrate: u32) -> Self {
        Self::new_mulaw(sample_rate, channels as u8)                               // the trait constructor defaults to mu-law
    }
    fn init(&mut self) -> Result<(), String> { Ok(()) }
    fn encode_i16(&mut self, input: &[i16], output: &mut [u8]) -> Result<usize, String> { encode_i16(self.law, input, output) }
    /// `(sample >> 16) as i16` then the i16 encoder (:124-138): the kernel takes the i32 slice and narrows on the device
    fn encode_i32(&mut self, input: &[i32], output: &mut [u8]) -> Result<usize, String> { encode_kind(self.law, input, output, sys::SK_KIND_I32) }
    fn reset(&mut self) -> Result<(), String> { Ok(()) }
}

#[derive(Debug)]
pub struct G711Decoder { law: G711Law, sample_rate: u32, channels: u8 }          // :145-150
impl G711Decoder {
    pub fn new(law: G711Law, sample_rate: u32, channels: u8) -> Self { Self { law, sample_rate, channels } }            // :153
    pub fn new_mulaw(sample_rate: u32, channels: u8) -> Self { Self::new(G711Law::MuLaw, sample_rate, channels) }
    pub fn new_alaw(sample_rate: u32, channels: u8) -> Self { Self::new(G711Law::ALaw, sample_rate, channels) }
    pub fn law(&self) -> G711Law { self.law }
    pub fn sample_rate(&self) -> u32 { self.sample_rate }
    pub fn channels(&self) -> u8 { self.channels }
}
impl Decoder for G711Decoder {                                                    // :182-233; `fec` is ignored by the reference too
    fn decode_i16(&mut self, input: &[u8], output: &mut [i16], _fec: bool) -> Result<usize, String> { decode_i16(self.law, input, output) }
    /// `i32::from(decode_sample) << 16` (:206-208)
    fn decode_i32(&mut self, input: &[u8], output: &mut [i32], _fec: bool) -> Result<usize, String> { decode_kind(self.law, input, output, sys::SK_KIND_I32) }
    /// `f32::from(decode_sample) / 32768.0` (:227-229)
    fn decode_f32(&mut self, input: &[u8], output: &mut [f32], _fec: bool) -> Result<usize, String> { decode_kind(self.law, input, output, sys::SK_KIND_F32) }
}
