//! Drop-in replacements for soundkit-g726's hot loops (soundkit-g726/src/lib.rs:598-951): the predictor (`G726Core`)
//! runs on the device; `pending_samples` / `pending_bytes`, the buffer-size checks and their error strings stay here,
//! exactly as in the reference.  SOURCE ONLY (see lib.rs); mirrors `soundkit_g726` in soundkit_host.hpp.
use crate::{check, with_ctx};
use soundkit_cuda_sys as sys;

/// `G726Core` (lib.rs:198-230) as the ABI's 24 words: yl yu dms dml ap a[2] b[6] pk[2] dq[6] sr[2] td
#[derive(Clone)]
pub struct Core(pub [i32; sys::SK_G726_STATE_INTS]);
impl Default for Core {
    fn default() -> Self { let mut s = [0i32; sys::SK_G726_STATE_INTS]; unsafe { sys::sk_g726_state_init(s.as_mut_ptr()) }; Core(s) }
}

/// whole groups only: `G726Encoder::encode_complete_samples` (lib.rs:672-698)
pub fn encode_complete_samples(core: &mut Core, bits: usize, right_packing: bool, samples: &[i16], output: &mut [u8]) -> Result<(), String> {
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g726_encode(ctx, bits as i32, right_packing as i32, core.0.as_mut_ptr(), sys::SK_KIND_I16, samples.as_ptr().cast(), samples.len(),
                            output.as_mut_ptr())
    }))
}

/// whole groups only: the decode loop of `G726Decoder::decode_i16` (lib.rs:842-853); `out_kind` = SK_KIND_{I16,I32,F32}
pub fn decode_complete_bytes<T>(core: &mut Core, bits: usize, right_packing: bool, bytes: &[u8], out_kind: i32, output: &mut [T]) -> Result<(), String> {
    with_ctx(|ctx| check(ctx, unsafe {
        sys::sk_g726_decode(ctx, bits as i32, right_packing as i32, core.0.as_mut_ptr(), bytes.as_ptr(), bytes.len(), out_kind, output.as_mut_ptr().cast())
    }))
}
