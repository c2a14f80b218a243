//! GPU (NVIDIA B200) versions of radiorust's sample-chain blocks.
//!
//! The blocks in [`blocks`] have the constructors and setters of
//! `radiorust::blocks::{FreqShifter, filters::Filter, Downsampler, Upsampler, modulation::FmDemod}` and implement
//! `radiorust::flow::{Consumer, Producer}` for `Signal<Complex<Flt>>` through radiorust's own `impl_block_trait!`
//! (src/blocks/mod.rs:111-144), so they connect with `feed_into` / `feed_from` like any other block:
//!
//! ```ignore
//! use radiorust::prelude::*;
//! use radiorust_b200::blocks::GpuChain;
//!
//! let ddc = GpuChain::<f32>::builder(0)
//!     .freq_shifter(1.0, -577_000.0)
//!     .filter(|_, freq| if freq.abs() <= 3000.0 { Complex::from(1.0) } else { Complex::from(0.0) })
//!     .downsampler(4096, 48_000.0, 6_000.0)
//!     .build();
//! ddc.feed_from(&sdr_rx);
//! ddc.feed_into(&demodulator);
//! ```
//!
//! A [`blocks::GpuChain`] keeps the samples on the device between its stages: PCIe is touched once per direction,
//! at the chain's edges, from pinned buffers of a [`pool::PinnedChunkBufPool`].  There is no CPU fallback: creating
//! a block without a B200 panics in the constructor (like `SoapySdrRx::new` fails without a device).
//!
//! All sample arithmetic lives in the CUDA library behind `radiorust-b200-sys`; this crate is the host side only:
//! handles, the pinned pool, and the Tokio task loops of the reference blocks with their DSP replaced by
//! `rr_chain_push`.
#![warn(missing_docs)]

pub mod blocks;
pub mod chain;
pub mod pool;

pub use radiorust_b200_sys as sys;

use std::ffi::CStr;
use std::fmt;
use std::sync::Arc;

/// Error of the C ABI: status code (`RR_ERR_*`) and the message of `rr_last_error`
#[derive(Clone, Debug)]
pub struct Error {
    /// `sys::RR_ERR_*`
    pub code: i32,
    /// text of `rr_last_error()` read on the failing thread right after the call
    pub message: String,
}

impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "radiorust_b200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

/// Turns a status code into a `Result`; must run on the thread that made the call, before any `.await`
/// (`rr_last_error` is thread local and Tokio tasks migrate).
pub(crate) fn check(code: i32) -> Result<(), Error> {
    if code == sys::RR_OK {
        return Ok(());
    }
    let message = unsafe {
        let p = sys::rr_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    };
    Err(Error { code, message })
}

struct ContextInner(*mut sys::rr_ctx);
// rr_ctx holds the device ordinal and SM count only; every entry point selects the device itself
unsafe impl Send for ContextInner {}
unsafe impl Sync for ContextInner {}
impl Drop for ContextInner {
    fn drop(&mut self) {
        unsafe {
            sys::rr_ctx_destroy(self.0);
        }
    }
}

/// One CUDA device (`rr_ctx`); cheap to clone, shared by chains and pools
#[derive(Clone)]
pub struct Context(Arc<ContextInner>);

impl Context {
    /// Opens CUDA device `device`; fails when there is none (there is no CPU fallback) or when it is not sm_100
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::rr_ctx_create(device, &mut p) })?;
        Ok(Context(Arc::new(ContextInner(p))))
    }
    /// CUDA device ordinal
    pub fn device(&self) -> i32 {
        unsafe { sys::rr_ctx_device(self.0 .0) }
    }
    pub(crate) fn raw(&self) -> *mut sys::rr_ctx {
        self.0 .0
    }
}

/// The two precisions radiorust's `Float` covers (src/numbers.rs:23-42), with the C ABI's dtype tag
pub trait GpuFloat: radiorust::numbers::Float + Send + Sync + 'static {
    /// `RR_C32` / `RR_C64`
    const DTYPE: i32;
}
impl GpuFloat for f32 {
    const DTYPE: i32 = sys::RR_C32;
}
impl GpuFloat for f64 {
    const DTYPE: i32 = sys::RR_C64;
}
