//! Safe Rust layer with the reference's signatures, on top of `soundkit-cuda-sys`.
//! Each function keeps the name, argument meaning and error strings of the function it replaces, so
//! `soundkit`, `soundkit-g711` and `soundkit-decoder` can re-export these instead of their scalar loops.
//! SOURCE ONLY: mirrors `soundkit_b200/csrc/host/soundkit_host.hpp`, which is the compiled and tested
//! equivalent in this repository.
use soundkit_cuda_sys as sys;
use std::cell::RefCell;
use std::ffi::CStr;

pub mod audio_bytes;
pub mod audio_packet;
pub mod audio_pipeline;
pub mod decoder;
pub mod g711;
pub mod g722;
pub mod g726;
pub mod wav;

/// One `sk_ctx` per thread: the reference runs one worker thread per `DecodePipeline`
/// (soundkit-decoder/src/lib.rs:2764) and contexts are thread-confined.
pub(crate) struct Gpu(*mut sys::sk_ctx);
impl Drop for Gpu { fn drop(&mut self) { unsafe { sys::sk_ctx_destroy(self.0) } } }
thread_local! { static GPU: RefCell<Option<Gpu>> = RefCell::new(None); }

pub(crate) fn with_ctx<T>(f: impl FnOnce(*mut sys::sk_ctx) -> Result<T, String>) -> Result<T, String> {
    GPU.with(|cell| {
        let mut slot = cell.borrow_mut();
        if slot.is_none() {
            let mut p = std::ptr::null_mut();
            let rc = unsafe { sys::sk_ctx_create(0, &mut p) };
            if rc != sys::SK_OK {
                return Err("soundkit-cuda: no CUDA device (there is no CPU fallback)".to_string());
            }
            *slot = Some(Gpu(p));
        }
        f(slot.as_ref().unwrap().0)
    })
}

pub(crate) fn check(ctx: *mut sys::sk_ctx, rc: i32) -> Result<(), String> {
    if rc == sys::SK_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(sys::sk_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(format!("soundkit-cuda: {msg}"))
}
