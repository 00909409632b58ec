//! `soundkit_g711` (soundkit-g711/src/lib.rs) — same free functions and error strings.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum G711Law { MuLaw, ALaw }                                                 // :17
fn law_id(law: G711Law) -> i32 { match law { G711Law::MuLaw => sys::SK_LAW_MULAW, G711Law::ALaw => sys::SK_LAW_ALAW } }

pub fn encode_i16(law: G711Law, input: &[i16], output: &mut [u8]) -> Result<usize, String> {   // :36
    if output.len() < input.len() {
        return Err(format!("Output buffer too small for G.711 encode: need {}, have {}", input.len(), output.len()));
    }
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g711_encode(ctx, law_id(law), sys::SK_KIND_I16, input.as_ptr() as *const _, input.len(), output.as_mut_ptr())
    }))?;
    Ok(input.len())
}

pub fn decode_i16(law: G711Law, input: &[u8], output: &mut [i16]) -> Result<usize, String> {   // :52
    if output.len() < input.len() {
        return Err(format!("Output buffer too small for G.711 decode: need {}, have {}", input.len(), output.len()));
    }
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g711_decode(ctx, law_id(law), input.as_ptr(), input.len(), sys::SK_KIND_I16, output.as_mut_ptr() as *mut _)
    }))?;
    Ok(input.len())
}

pub fn encode_sample(law: G711Law, sample: i16) -> u8 { let mut o = [0u8]; encode_i16(law, &[sample], &mut o).unwrap(); o[0] }  // :22
pub fn decode_sample(law: G711Law, sample: u8) -> i16 { let mut o = [0i16]; decode_i16(law, &[sample], &mut o).unwrap(); o[0] } // :29
// G711Encoder / G711Decoder (:69-233) wrap these exactly like the reference structs do; decode_i32 /
// decode_f32 / encode_i32 pass SK_KIND_I32 / SK_KIND_F32 (see soundkit_host.hpp for the full mirror).
