//! concrete-ntt's `Plan` API on NVIDIA B200 GPUs.
//!
//! Same module paths, type names, method names, argument meaning, `Option` / panic behaviour and bit-exact
//! results as `concrete-ntt` 0.2.0; the bodies call `libcntt_b200.so` (include/cntt_b200.h).  Every method of
//! the reference operates on one polynomial in host memory and keeps doing so here (the library stages it
//! through the device); the `*_batch` and `*_device` methods are the extensions that make a GPU worthwhile:
//! many polynomials per call, optionally resident in device memory on a caller-provided CUDA stream.
//!
//! Not compiled in the build image of this repository (no Rust toolchain there); see INTEGRATION.md.
pub mod ffi;
pub mod native;
pub mod prime32;
pub mod prime64;
pub mod product;

pub use native::{native128, native32, native64, native_binary128, native_binary32, native_binary64};

/// `concrete_ntt::prime` -- plan-time helpers evaluated by the library's host code (same values as src/prime.rs).
pub mod prime {
    /// src/prime.rs:76-126
    pub fn is_prime64(n: u64) -> bool {
        unsafe { crate::ffi::cntt_is_prime64(n) != 0 }
    }
    /// src/prime.rs:130-180
    pub fn largest_prime_in_arithmetic_progression64(factor: u64, offset: u64, lo: u64, hi: u64) -> Option<u64> {
        let mut out = 0u64;
        (unsafe { crate::ffi::cntt_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, &mut out) } != 0).then_some(out)
    }
}

/// Version string of the loaded library, e.g. `cntt_b200 0.1 (sm_100a; concrete-ntt 0.2.0 semantics)`.
pub fn backend_version() -> String {
    unsafe { core::ffi::CStr::from_ptr(ffi::cntt_version()) }.to_string_lossy().into_owned()
}
