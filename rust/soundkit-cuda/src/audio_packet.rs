//! The two codec traits of `soundkit::audio_packet` (soundkit/src/audio_packet.rs:6-26), restated so that this crate
//! builds on its own.  Inside the soundkit workspace enable the `soundkit-traits` feature: the structs below then
//! implement soundkit's own `Encoder` / `Decoder` (same items, re-exported) and drop into every `impl Encoder` /
//! `impl Decoder` call site unchanged.
#[cfg(feature = "soundkit-traits")]
pub use soundkit::audio_packet::{Decoder, Encoder};

#[cfg(not(feature = "soundkit-traits"))]
pub trait Encoder {
    fn new(sample_rate: u32, bits_per_sample: u32, channels: u32, frame_size: u32, bitrate: u32) -> Self;
    fn init(&mut self) -> Result<(), String>;
    fn encode_i16(&mut self, input: &[i16], output: &mut [u8]) -> Result<usize, String>;
    fn encode_i32(&mut self, input: &[i32], output: &mut [u8]) -> Result<usize, String>;
    fn reset(&mut self) -> Result<(), String>;
}

#[cfg(not(feature = "soundkit-traits"))]
pub trait Decoder {
    fn decode_i16(&mut self, input: &[u8], output: &mut [i16], fec: bool) -> Result<usize, String>;
    fn decode_i32(&mut self, input: &[u8], output: &mut [i32], fec: bool) -> Result<usize, String>;
    fn decode_f32(&mut self, input: &[u8], output: &mut [f32], fec: bool) -> Result<usize, String>;
}
