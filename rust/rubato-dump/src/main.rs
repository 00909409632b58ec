//! rubato-dump <dir>: for every `<case>.json` + `<case>.in.f32` in <dir> (written by tools/dump_rubato_inputs.py, planar f32 LE
//! [channels][frames]) write `<case>.out.f32` (planar f32 LE [channels][n_out]) produced by the real rubato 0.14.1 exactly the way
//! soundkit drives it:
//!   "oneshot"   -> soundkit::audio_pipeline::downsample_audio (soundkit/src/audio_pipeline.rs:474-491): SincFixedIn with
//!                  chunk_size = the whole buffer, one process() call;
//!   "streaming" -> soundkit-decoder's StreamingResampler (soundkit-decoder/src/lib.rs:1931-2059): chunk 4096, process() per
//!                  chunk, then flush: process_partial(Some(rest)) and drop round(padded * out / in) frames, or
//!                  process_partial(None) when nothing is pending.
//! Then `python -m pytest tests/test_oracle_golden.py -k rubato` (CPU checker) and `-m gpu -k rubato` (CUDA kernels) compare.
//! The JSON is read with a minimal hand parser to keep the dependency list at exactly one crate.
use rubato::{Resampler, SincFixedIn, SincInterpolationParameters, SincInterpolationType, WindowFunction};
use std::fs;

fn field(json: &str, key: &str) -> String {
    let k = format!("\"{}\"", key);
    let at = json.find(&k).unwrap_or_else(|| panic!("missing {key}")) + k.len();
    let rest = json[at..].trim_start().trim_start_matches(':').trim_start();
    let end = rest.find(|c| c == ',' || c == '}').unwrap();
    rest[..end].trim().trim_matches('"').to_string()
}

fn params() -> SincInterpolationParameters {
    SincInterpolationParameters { sinc_len: 256, f_cutoff: 0.95, interpolation: SincInterpolationType::Linear,
                                  oversampling_factor: 256, window: WindowFunction::BlackmanHarris2 }
}

const CHUNK: usize = 4096; // RESAMPLE_CHUNK_SIZE, soundkit-decoder/src/lib.rs:79

fn main() {
    let dir = std::env::args().nth(1).expect("usage: rubato-dump <dir>");
    for e in fs::read_dir(&dir).unwrap() {
        let p = e.unwrap().path();
        if p.extension().and_then(|x| x.to_str()) != Some("json") { continue; }
        let meta = fs::read_to_string(&p).unwrap();
        let (in_rate, out_rate): (f64, f64) = (field(&meta, "in_rate").parse().unwrap(), field(&meta, "out_rate").parse().unwrap());
        let ch: usize = field(&meta, "channels").parse().unwrap();
        let frames: usize = field(&meta, "frames").parse().unwrap();
        let semantics = field(&meta, "semantics");
        let raw = fs::read(p.with_extension("in.f32")).unwrap();
        assert_eq!(raw.len(), ch * frames * 4);
        let data: Vec<Vec<f32>> = (0..ch).map(|c| (0..frames).map(|i| {
            let o = (c * frames + i) * 4; f32::from_le_bytes([raw[o], raw[o + 1], raw[o + 2], raw[o + 3]]) }).collect()).collect();
        let mut out: Vec<Vec<f32>> = vec![Vec::new(); ch];
        if semantics == "oneshot" {
            let mut r = SincFixedIn::<f32>::new(out_rate / in_rate, 2.0, params(), frames, ch).unwrap();
            out = r.process(&data, None).unwrap();
        } else {
            let mut r = SincFixedIn::<f32>::new(out_rate / in_rate, 2.0, params(), CHUNK, ch).unwrap();
            let mut pos = 0;
            while frames - pos >= CHUNK {
                let chunk: Vec<&[f32]> = data.iter().map(|c| &c[pos..pos + CHUNK]).collect();
                let o = r.process(&chunk, None).unwrap();
                for (dst, src) in out.iter_mut().zip(o) { dst.extend(src); }
                pos += CHUNK;
            }
            let remaining = frames - pos;
            if remaining > 0 {
                let rest: Vec<&[f32]> = data.iter().map(|c| &c[pos..]).collect();
                let mut o = r.process_partial(Some(&rest), None).unwrap();
                let padded = CHUNK - remaining;
                let trim = ((padded as f64 * out_rate) / in_rate).round() as usize;
                for c in o.iter_mut() { let keep = c.len().saturating_sub(trim); c.truncate(keep); }
                for (dst, src) in out.iter_mut().zip(o) { dst.extend(src); }
            } else {
                let o = r.process_partial::<&[f32]>(None, None).unwrap();
                for (dst, src) in out.iter_mut().zip(o) { dst.extend(src); }
            }
        }
        let mut bytes = Vec::with_capacity(out.iter().map(|c| c.len() * 4).sum());
        for c in &out { for v in c { bytes.extend_from_slice(&v.to_le_bytes()); } }
        fs::write(p.with_extension("out.f32"), bytes).unwrap();
        eprintln!("{}: {} x {} frames", p.display(), ch, out[0].len());
    }
}
