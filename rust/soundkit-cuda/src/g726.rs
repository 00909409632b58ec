//! `soundkit_g726` (soundkit-g726/src/lib.rs): the ADPCM predictor (`G726Core`, :198-596) runs on the device; the rate /
//! packing types, `pending_samples` / `pending_bytes`, every buffer-size check and its error string stay here, exactly as
//! in the reference.  SOURCE ONLY (see lib.rs); mirrors `soundkit_g726` in soundkit_host.hpp, which is compiled and tested.
use crate::audio_packet::{Decoder, Encoder};
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

pub const G726_SAMPLE_RATE: u32 = 8_000;                                          // :7
pub const G726_CHANNELS: u8 = 1;                                                  // :8
pub const G726_16K_BIT_RATE: u32 = 16_000;
pub const G726_24K_BIT_RATE: u32 = 24_000;
pub const G726_32K_BIT_RATE: u32 = 32_000;
pub const G726_40K_BIT_RATE: u32 = 40_000;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum G726Rate { Rate16000, Rate24000, Rate32000, Rate40000 }                  // :62-72
impl G726Rate {
    pub fn from_bitrate(bit_rate: u32) -> Self {                                  // :75-82
        match bit_rate { G726_16K_BIT_RATE => Self::Rate16000, G726_24K_BIT_RATE => Self::Rate24000, G726_40K_BIT_RATE => Self::Rate40000, _ => Self::Rate32000 }
    }
    pub fn bits_per_sample(self) -> usize { match self { Self::Rate16000 => 2, Self::Rate24000 => 3, Self::Rate32000 => 4, Self::Rate40000 => 5 } }
    pub fn bit_rate(self) -> u32 { match self { Self::Rate16000 => G726_16K_BIT_RATE, Self::Rate24000 => G726_24K_BIT_RATE, Self::Rate32000 => G726_32K_BIT_RATE, Self::Rate40000 => G726_40K_BIT_RATE } }
    fn samples_per_byte_group(self) -> usize { match self { Self::Rate16000 => 4, Self::Rate24000 => 8, Self::Rate32000 => 2, Self::Rate40000 => 8 } }  // :104-111
    fn bytes_per_group(self) -> usize { match self { Self::Rate16000 => 1, Self::Rate24000 => 3, Self::Rate32000 => 1, Self::Rate40000 => 5 } }         // :113-120
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum G726Packing { Left, Right }                                              // :191-197 (`ffmpeg -f g726` / `g726le`)

/// `G726Core` (:198-230) as the ABI's 24 words: yl yu dms dml ap a[2] b[6] pk[2] dq[6] sr[2] td
#[derive(Clone)]
struct Core([i32; sys::SK_G726_STATE_INTS]);
impl Default for Core {
    fn default() -> Self { let mut s = [0i32; sys::SK_G726_STATE_INTS]; unsafe { sys::sk_g726_state_init(s.as_mut_ptr()) }; Core(s) }
}

pub struct G726Encoder { core: Core, rate: G726Rate, packing: G726Packing, pending_samples: Vec<i16> }   // :563-569
impl G726Encoder {
    pub fn try_new(rate: G726Rate, packing: G726Packing) -> Result<Self, String> {
        Ok(Self { core: Core::default(), rate, packing, pending_samples: Vec::with_capacity(rate.samples_per_byte_group()) })
    }
    pub fn new_32k_left() -> Self { Self::try_new(G726Rate::Rate32000, G726Packing::Left).expect("failed to initialize G.726 encoder") }
    pub fn new_32k_right() -> Self { Self::try_new(G726Rate::Rate32000, G726Packing::Right).expect("failed to initialize G.726 encoder") }
    pub fn new_16k_left() -> Self { Self::try_new(G726Rate::Rate16000, G726Packing::Left).expect("failed to initialize G.726-16 encoder") }
    pub fn new_24k_left() -> Self { Self::try_new(G726Rate::Rate24000, G726Packing::Left).expect("failed to initialize G.726-24 encoder") }
    pub fn new_40k_left() -> Self { Self::try_new(G726Rate::Rate40000, G726Packing::Left).expect("failed to initialize G.726-40 encoder") }
    pub fn rate(&self) -> G726Rate { self.rate }
    pub fn packing(&self) -> G726Packing { self.packing }

    pub fn encode_to_vec(&mut self, input: &[i16], output: &mut Vec<u8>) -> Result<usize, String> {   // :598-605
        let start = output.len();
        let required = self.required_encode_bytes(input.len(), false);
        output.resize(start + required, 0);
        let written = self.encode_i16(input, &mut output[start..])?;
        output.truncate(start + written);
        Ok(written)
    }
    pub fn flush_into(&mut self, output: &mut [u8]) -> Result<usize, String> {                         // :607-625
        let required = self.required_encode_bytes(0, true);
        if output.len() < required {
            return Err(format!("Output buffer too small for G.726 flush: need {}, have {}", required, output.len()));
        }
        if self.pending_samples.is_empty() { return Ok(0); }
        let mut samples = self.pending_samples.clone();
        samples.resize(self.rate.samples_per_byte_group(), 0);
        self.pending_samples.clear();
        self.encode_complete_samples(&samples, output)
    }
    pub fn flush_to_vec(&mut self, output: &mut Vec<u8>) -> Result<usize, String> {                    // :627-634
        let start = output.len();
        let required = self.required_encode_bytes(0, true);
        output.resize(start + required, 0);
        let written = self.flush_into(&mut output[start..])?;
        output.truncate(start + written);
        Ok(written)
    }
    fn required_encode_bytes(&self, input_samples: usize, flush: bool) -> usize {                      // :644-653
        let group_samples = self.rate.samples_per_byte_group();
        let total_samples = self.pending_samples.len() + input_samples;
        let groups = if flush { total_samples.div_ceil(group_samples) } else { total_samples / group_samples };
        groups * self.rate.bytes_per_group()
    }
    /// whole groups only (:655-677): every caller has checked `output.len()` against `required_encode_bytes` already;
    /// the length is asserted again here because the device writes `groups * bytes_per_group` bytes unconditionally
    fn encode_complete_samples(&mut self, samples: &[i16], output: &mut [u8]) -> Result<usize, String> {
        let (gs, gb) = (self.rate.samples_per_byte_group(), self.rate.bytes_per_group());
        assert!(samples.len() % gs == 0, "partial G.726 code group");
        let bytes = samples.len() / gs * gb;
        assert!(output.len() >= bytes, "G.726 output slice shorter than the encoded groups");
        if samples.is_empty() { return Ok(0); }
        with_ctx(|ctx| check(ctx, unsafe {
            sys::sk_g726_encode(ctx, self.rate.bits_per_sample() as i32, (self.packing == G726Packing::Right) as i32, self.core.0.as_mut_ptr(),
                                sys::SK_KIND_I16, samples.as_ptr().cast(), samples.len(), output.as_mut_ptr())
        }))?;
        Ok(bytes)
    }
}
impl Default for G726Encoder { fn default() -> Self { Self::new_32k_left() } }
impl Encoder for G726Encoder {                                                                         // :685-744
    fn new(_sample_rate: u32, _bits_per_sample: u32, _channels: u32, _frame_size: u32, bitrate: u32) -> Self {
        Self::try_new(G726Rate::from_bitrate(bitrate), G726Packing::Left).expect("failed to initialize G.726 encoder")
    }
    fn init(&mut self) -> Result<(), String> { Ok(()) }
    fn encode_i16(&mut self, input: &[i16], output: &mut [u8]) -> Result<usize, String> {              // :700-733
        let required = self.required_encode_bytes(input.len(), false);
        if output.len() < required {
            return Err(format!("Output buffer too small for G.726 encode: need {}, have {}", required, output.len()));
        }
        let group_samples = self.rate.samples_per_byte_group();
        let total_samples = self.pending_samples.len() + input.len();
        let complete_samples = (total_samples / group_samples) * group_samples;
        let mut samples = Vec::with_capacity(total_samples);
        samples.extend_from_slice(&self.pending_samples);
        samples.extend_from_slice(input);
        let written = if complete_samples == 0 { 0 } else { self.encode_complete_samples(&samples[..complete_samples], output)? };
        self.pending_samples.clear();
        self.pending_samples.extend_from_slice(&samples[complete_samples..]);
        Ok(written)
    }
    fn encode_i32(&mut self, input: &[i32], output: &mut [u8]) -> Result<usize, String> {              // :735-738
        let samples: Vec<i16> = input.iter().map(|sample| (sample >> 16) as i16).collect();
        self.encode_i16(&samples, output)
    }
    fn reset(&mut self) -> Result<(), String> { self.core = Core::default(); self.pending_samples.clear(); Ok(()) }
}

pub struct G726Decoder { core: Core, rate: G726Rate, packing: G726Packing, pending_bytes: Vec<u8> }    // :746-751
impl G726Decoder {
    pub fn try_new(rate: G726Rate, packing: G726Packing) -> Result<Self, String> {
        Ok(Self { core: Core::default(), rate, packing, pending_bytes: Vec::with_capacity(rate.bytes_per_group()) })
    }
    pub fn new_32k_left() -> Self { Self::try_new(G726Rate::Rate32000, G726Packing::Left).expect("failed to initialize G.726 decoder") }
    pub fn new_32k_right() -> Self { Self::try_new(G726Rate::Rate32000, G726Packing::Right).expect("failed to initialize G.726 decoder") }
    pub fn new_16k_left() -> Self { Self::try_new(G726Rate::Rate16000, G726Packing::Left).expect("failed to initialize G.726-16 decoder") }
    pub fn new_24k_left() -> Self { Self::try_new(G726Rate::Rate24000, G726Packing::Left).expect("failed to initialize G.726-24 decoder") }
    pub fn new_40k_left() -> Self { Self::try_new(G726Rate::Rate40000, G726Packing::Left).expect("failed to initialize G.726-40 decoder") }
    pub fn sample_rate(&self) -> u32 { G726_SAMPLE_RATE }
    pub fn channels(&self) -> u8 { G726_CHANNELS }
    pub fn rate(&self) -> G726Rate { self.rate }
    pub fn packing(&self) -> G726Packing { self.packing }
    pub fn flush(&mut self) -> Result<(), String> {                                                    // :802-811
        if self.pending_bytes.is_empty() { Ok(()) }
        else { Err(format!("G.726 stream ended with {} trailing partial-packet byte(s)", self.pending_bytes.len())) }
    }
    /// the shared body of decode_{i16,i32,f32} (:821-951): i32 = `i32::from(sample) << 16`, f32 = `f32::from(sample) / 32768.0`
    /// are applied on the device (`out_kind`); `T` is fixed by the three trait methods below, never chosen by a caller
    fn run<T>(&mut self, input: &[u8], output: &mut [T], out_kind: i32) -> Result<usize, String> {
        let group_bytes = self.rate.bytes_per_group();
        let total_bytes = self.pending_bytes.len() + input.len();
        let complete_bytes = (total_bytes / group_bytes) * group_bytes;
        let required = (complete_bytes / group_bytes) * self.rate.samples_per_byte_group();
        if output.len() < required {
            return Err(format!("Output buffer too small for G.726 decode: need {}, have {}", required, output.len()));
        }
        let mut bytes = Vec::with_capacity(total_bytes);
        bytes.extend_from_slice(&self.pending_bytes);
        bytes.extend_from_slice(input);
        if complete_bytes > 0 {
            with_ctx(|ctx| check(ctx, unsafe {
                sys::sk_g726_decode(ctx, self.rate.bits_per_sample() as i32, (self.packing == G726Packing::Right) as i32, self.core.0.as_mut_ptr(),
                                    bytes.as_ptr(), complete_bytes, out_kind, output.as_mut_ptr().cast())
            }))?;
        }
        self.pending_bytes.clear();
        self.pending_bytes.extend_from_slice(&bytes[complete_bytes..]);
        Ok(required)
    }
}
impl Default for G726Decoder { fn default() -> Self { Self::new_32k_left() } }
impl Decoder for G726Decoder {
    fn decode_i16(&mut self, input: &[u8], output: &mut [i16], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_I16) }
    fn decode_i32(&mut self, input: &[u8], output: &mut [i32], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_I32) }
    fn decode_f32(&mut self, input: &[u8], output: &mut [f32], _fec: bool) -> Result<usize, String> { self.run(input, output, sys::SK_KIND_F32) }
}
