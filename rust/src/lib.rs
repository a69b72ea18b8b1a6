//! concrete-ntt's `Plan` API on NVIDIA B200 GPUs.
//!
//! Same module paths, type names, method names, argument meaning, `Option` / panic behaviour and bit-exact
//! results as `concrete-ntt` 0.2.0; the bodies call `libcntt_b200.so` (include/cntt_b200.h).  Every method of
//! the reference operates on one polynomial in host memory and keeps doing so here (the library stages it
//! through the device); the `*_batch` and `*_device` methods are the extensions that make a GPU worthwhile:
//! many polynomials per call, optionally resident in device memory on a caller-provided CUDA stream.
//!
//! Not compiled in the build image of this repository (no Rust toolchain there); see INTEGRATION.md.
pub mod fastdiv;
pub mod ffi;
pub mod native;
pub mod prime32;
pub mod prime64;
pub mod product;

pub use native::{native128, native32, native64, native_binary128, native_binary32, native_binary64};

/// `concrete_ntt::prime` -- plan-time helpers (same values as src/prime.rs).
pub mod prime {
    use crate::fastdiv::{Div32, Div64};

    /// src/prime.rs:4-6
    pub const fn mul_mod32(n: Div32, x: u32, y: u32) -> u32 {
        Div32::rem_u64(x as u64 * y as u64, n)
    }
    /// src/prime.rs:8-10
    pub const fn mul_mod64(n: Div64, x: u64, y: u64) -> u64 {
        Div64::rem_u128(x as u128 * y as u128, n)
    }
    /// src/prime.rs:12-29
    pub const fn exp_mod32(n: Div32, base: u32, pow: u32) -> u32 {
        let (mut result, mut base, mut pow) = (1u32 % n.divisor, base % n.divisor, pow);
        while pow > 0 {
            if pow & 1 == 1 {
                result = mul_mod32(n, result, base);
            }
            base = mul_mod32(n, base, base);
            pow >>= 1;
        }
        result
    }
    /// src/prime.rs:31-48
    pub const fn exp_mod64(n: Div64, base: u64, pow: u64) -> u64 {
        let (mut result, mut base, mut pow) = (1u64 % n.divisor, base % n.divisor, pow);
        while pow > 0 {
            if pow & 1 == 1 {
                result = mul_mod64(n, result, base);
            }
            base = mul_mod64(n, base, base);
            pow >>= 1;
        }
        result
    }
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

/// Version string of the loaded library, e.g. `cntt_b200 0.2 (sm_100a; concrete-ntt 0.2.0 semantics)`.
pub fn backend_version() -> String {
    unsafe { core::ffi::CStr::from_ptr(ffi::cntt_version()) }.to_string_lossy().into_owned()
}
