//! Emits true rubato 0.14.1 vectors for the seeded inputs the parity tests use, so the "parity
//! unpinned" status of the resampler VALUES can be lifted by anyone with cargo:
//!   cargo run --release --example dump_rubato -- 48000 16000 2 30011 > rubato_48k_16k.f32
//! Input: the same white noise tests/test_gpu_parity.py feeds (written by the test with --dump-inputs).
use rubato::{Resampler, SincFixedIn, SincInterpolationParameters, SincInterpolationType, WindowFunction};
use std::io::{Read, Write};

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let (in_rate, out_rate, ch, frames): (f64, f64, usize, usize) = (a[1].parse().unwrap(), a[2].parse().unwrap(), a[3].parse().unwrap(), a[4].parse().unwrap());
    let mut raw = Vec::new();
    std::io::stdin().read_to_end(&mut raw).unwrap();            // planar f32 LE [ch][frames]
    let data: Vec<Vec<f32>> = (0..ch).map(|c| (0..frames).map(|i| {
        let o = (c * frames + i) * 4; f32::from_le_bytes([raw[o], raw[o + 1], raw[o + 2], raw[o + 3]]) }).collect()).collect();
    let params = SincInterpolationParameters { sinc_len: 256, f_cutoff: 0.95, interpolation: SincInterpolationType::Linear,
                                               oversampling_factor: 256, window: WindowFunction::BlackmanHarris2 };
    let mut r = SincFixedIn::<f32>::new(out_rate / in_rate, 2.0, params, frames, ch).unwrap();
    let out = r.process(&data, None).unwrap();
    let mut so = std::io::stdout().lock();
    for c in &out { for v in c { so.write_all(&v.to_le_bytes()).unwrap(); } }
}
