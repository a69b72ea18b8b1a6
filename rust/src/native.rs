//! `native32 / native64 / native128 / native_binary32 / native_binary64 / native_binary128 ::Plan32` and the `Plan52`
//! twins (reference: src/native*.rs, src/native_binary*.rs).
//!
//! The reference's `fwd` / `inv` take one `&mut [u32]` per prime; here they take one slice holding the planes back to
//! back (`mod_p[k * n .. (k + 1) * n]` = the reference's `mod_pk`), which is what the device kernels read and write.
//! `u128` words are passed as they lie in memory on little-endian targets (`{lo, hi}` u64 pairs).
use crate::ffi;
use core::ffi::c_void;
use core::ptr::NonNull;

macro_rules! native_module {
    ($module:ident, $word:ty, $bits:expr, $binary:expr, $nprimes:expr, $doc:expr) => {
        #[doc = $doc]
        pub mod $module {
            use super::*;

            /// Negacyclic NTT plan for multiplying two polynomials with wrapping word arithmetic.
            pub struct Plan32 {
                raw: NonNull<ffi::NativePlan>,
                n: usize,
                device: i32,
            }
            unsafe impl Send for Plan32 {}
            unsafe impl Sync for Plan32 {}

            impl Plan32 {
                /// Number of 32-bit primes of this plan kind.
                pub const NUM_PRIMES: usize = $nprimes;

                /// `Plan32::try_new(n)`: `None` unless `n` is a power of two in `32..=32768`.
                pub fn try_new(n: usize) -> Option<Self> {
                    Self::build(n, 0, false)
                }
                /// Extension: tables on CUDA device `device`.
                pub fn try_new_on(n: usize, device: i32) -> Option<Self> {
                    Self::build(n, device, false)
                }
                /// EXTENSION (no reference counterpart): the primes `k * 2^17 + 1` just below 2^30, which also admit
                /// `n = 65536`; `negacyclic_polymul` is bit-identical to `try_new(n)` plans wherever both exist.
                /// `None` for native128 (ten primes needed, nine exist).
                pub fn try_new_extended(n: usize) -> Option<Self> {
                    Self::build(n, 0, true)
                }
                fn build(n: usize, device: i32, extended: bool) -> Option<Self> {
                    let mut raw = core::ptr::null_mut();
                    let st = unsafe {
                        if extended {
                            ffi::cntt_native_plan_new_ext(n, $bits, $binary as i32, device, &mut raw)
                        } else {
                            ffi::cntt_native_plan_new(n, $bits, $binary as i32, device, &mut raw)
                        }
                    };
                    ffi::plan_status(st)?;
                    Some(Self { raw: NonNull::new(raw)?, n, device })
                }
                /// `Plan32::ntt_size`
                #[inline]
                pub fn ntt_size(&self) -> usize {
                    self.n
                }
                /// Modulus of `Plan32::ntt_i()`.
                pub fn ntt_modulus(&self, i: usize) -> u32 {
                    unsafe { ffi::cntt_native_prime(self.raw.as_ptr(), i as i32) }
                }
                /// `Plan32::ntt_0() .. ntt_9()`: an equal `prime32::Plan` (plans are pure functions of `(n, p)`).
                pub fn ntt_i(&self, i: usize) -> crate::prime32::Plan {
                    crate::prime32::Plan::try_new_on(self.n, self.ntt_modulus(i), self.device).expect("the sub-plan of an existing plan exists")
                }
                /// `Plan32::fwd(value, mod_p0, mod_p1, ..)` with the planes back to back in `mod_p`.
                pub fn fwd(&self, value: &[$word], mod_p: &mut [u32]) {
                    assert_eq!(value.len(), self.n);
                    assert_eq!(mod_p.len(), Self::NUM_PRIMES * self.n);
                    ffi::check(unsafe { ffi::cntt_native_fwd_host(self.raw.as_ptr(), value.as_ptr() as *const c_void, mod_p.as_mut_ptr(), self.n, 1) });
                }
                /// `Plan32::fwd_binary` (binary plans only; `UNSUPPORTED` panics otherwise, the reference has no such method).
                pub fn fwd_binary(&self, value: &[$word], mod_p: &mut [u32]) {
                    assert_eq!(value.len(), self.n);
                    assert_eq!(mod_p.len(), Self::NUM_PRIMES * self.n);
                    ffi::check(unsafe { ffi::cntt_native_fwd_binary_host(self.raw.as_ptr(), value.as_ptr() as *const c_void, mod_p.as_mut_ptr(), self.n, 1) });
                }
                /// `Plan32::inv(value, mod_p0, ..)`: `mod_p` is clobbered exactly like the reference's buffers.
                pub fn inv(&self, value: &mut [$word], mod_p: &mut [u32]) {
                    assert_eq!(value.len(), self.n);
                    assert_eq!(mod_p.len(), Self::NUM_PRIMES * self.n);
                    ffi::check(unsafe { ffi::cntt_native_inv_host(self.raw.as_ptr(), value.as_mut_ptr() as *mut c_void, mod_p.as_mut_ptr(), self.n, 1) });
                }
                /// `Plan32::negacyclic_polymul(prod, lhs, rhs)`
                pub fn negacyclic_polymul(&self, prod: &mut [$word], lhs: &[$word], rhs: &[$word]) {
                    let n = prod.len();
                    assert_eq!(n, lhs.len());
                    assert_eq!(n, rhs.len());
                    ffi::check(unsafe {
                        ffi::cntt_native_polymul_host(self.raw.as_ptr(), prod.as_mut_ptr() as *mut c_void, lhs.as_ptr() as *const c_void, rhs.as_ptr() as *const c_void, n, 1)
                    });
                }
                // ---- extensions ----
                /// `prod.len() / n` independent products in one call.
                pub fn negacyclic_polymul_batch(&self, prod: &mut [$word], lhs: &[$word], rhs: &[$word]) {
                    let len = prod.len();
                    assert_eq!(len, lhs.len());
                    assert_eq!(len, rhs.len());
                    assert_eq!(len % self.n, 0);
                    ffi::check(unsafe {
                        ffi::cntt_native_polymul_host(self.raw.as_ptr(), prod.as_mut_ptr() as *mut c_void, lhs.as_ptr() as *const c_void, rhs.as_ptr() as *const c_void, len, len / self.n)
                    });
                }
                /// Device-resident batch on a CUDA stream.
                ///
                /// # Safety
                /// the pointers must be valid device memory on the plan's device for `batch * n` words each.
                pub unsafe fn negacyclic_polymul_device(&self, d_prod: *mut $word, d_lhs: *const $word, d_rhs: *const $word, batch: usize, stream: ffi::Stream) {
                    ffi::check(ffi::cntt_native_polymul(self.raw.as_ptr(), d_prod as *mut c_void, d_lhs as *const c_void, d_rhs as *const c_void, batch, stream));
                }
                /// Residue planes on the device: plane `k` of polynomial `b` at `d_mod_p[(k * batch + b) * n ..]`.
                ///
                /// # Safety
                /// as [`Plan32::negacyclic_polymul_device`]; `d_mod_p` holds `NUM_PRIMES * batch * n` words.
                pub unsafe fn fwd_device(&self, d_value: *const $word, d_mod_p: *mut u32, batch: usize, stream: ffi::Stream) {
                    ffi::check(ffi::cntt_native_fwd(self.raw.as_ptr(), d_value as *const c_void, d_mod_p, batch, stream));
                }
                /// # Safety
                /// as [`Plan32::fwd_device`]
                pub unsafe fn inv_device(&self, d_value: *mut $word, d_mod_p: *mut u32, batch: usize, stream: ffi::Stream) {
                    ffi::check(ffi::cntt_native_inv(self.raw.as_ptr(), d_value as *mut c_void, d_mod_p, batch, stream));
                }
            }
            impl Drop for Plan32 {
                fn drop(&mut self) {
                    unsafe { ffi::cntt_native_plan_free(self.raw.as_ptr()) }
                }
            }
            impl Clone for Plan32 {
                fn clone(&self) -> Self {
                    Self::try_new_on(self.n, self.device).or_else(|| Self::build(self.n, self.device, true)).expect("a plan that exists can be rebuilt")
                }
            }
        }
    };
}

native_module!(native32, u32, 32, false, 3, "`native32` (src/native32.rs): products modulo 2^32 over P0..P2.");
native_module!(native64, u64, 64, false, 5, "`native64` (src/native64.rs): products modulo 2^64 over P0..P4.");
native_module!(native128, u128, 128, false, 10, "`native128` (src/native128.rs): products modulo 2^128 over P0..P9.");
native_module!(native_binary32, u32, 32, true, 2, "`native_binary32` (src/native_binary32.rs): rhs in {0, 1}, P0..P1.");
native_module!(native_binary64, u64, 64, true, 3, "`native_binary64` (src/native_binary64.rs): rhs in {0, 1}, P0..P2.");
native_module!(native_binary128, u128, 128, true, 5, "`native_binary128` (src/native_binary128.rs): rhs in {0, 1}, P0..P4.");

macro_rules! plan52 {
    ($module:ident, $word:ty, $bits:expr, $binary:expr, $nprimes:expr) => {
        /// `Plan52` of this module (reference: feature = "nightly" and AVX-512 IFMA only; always available here).
        /// Residue planes are `u64`, one per prime of `primes52`; `negacyclic_polymul` returns what `Plan32` returns.
        pub mod $module {
            use super::*;
            pub struct Plan52 {
                raw: NonNull<ffi::Native52Plan>,
                n: usize,
            }
            unsafe impl Send for Plan52 {}
            unsafe impl Sync for Plan52 {}
            impl Plan52 {
                pub const NUM_PRIMES: usize = $nprimes;
                pub fn try_new(n: usize) -> Option<Self> {
                    let mut raw = core::ptr::null_mut();
                    ffi::plan_status(unsafe { ffi::cntt_native52_plan_new(n, $bits, $binary as i32, 0, &mut raw) })?;
                    Some(Self { raw: NonNull::new(raw)?, n })
                }
                #[inline]
                pub fn ntt_size(&self) -> usize {
                    self.n
                }
                pub fn ntt_i(&self, i: usize) -> crate::prime64::Plan {
                    let p = unsafe { ffi::cntt_native52_prime(self.raw.as_ptr(), i as i32) };
                    crate::prime64::Plan::try_new(self.n, p).expect("the sub-plan of an existing plan exists")
                }
                pub fn negacyclic_polymul(&self, prod: &mut [$word], lhs: &[$word], rhs: &[$word]) {
                    let n = prod.len();
                    assert_eq!(n, lhs.len());
                    assert_eq!(n, rhs.len());
                    ffi::check(unsafe {
                        ffi::cntt_native52_polymul_host(self.raw.as_ptr(), prod.as_mut_ptr() as *mut c_void, lhs.as_ptr() as *const c_void, rhs.as_ptr() as *const c_void, n, 1)
                    });
                }
                /// # Safety
                /// device pointers on the plan's device; `d_mod_p` holds `NUM_PRIMES * batch * n` u64 words.
                pub unsafe fn fwd_device(&self, d_value: *const $word, d_mod_p: *mut u64, batch: usize, stream: ffi::Stream) {
                    ffi::check(ffi::cntt_native52_fwd(self.raw.as_ptr(), d_value as *const c_void, d_mod_p, batch, stream));
                }
                /// # Safety
                /// as [`Plan52::fwd_device`]
                pub unsafe fn inv_device(&self, d_value: *mut $word, d_mod_p: *mut u64, batch: usize, stream: ffi::Stream) {
                    ffi::check(ffi::cntt_native52_inv(self.raw.as_ptr(), d_value as *mut c_void, d_mod_p, batch, stream));
                }
            }
            impl Drop for Plan52 {
                fn drop(&mut self) {
                    unsafe { ffi::cntt_native52_plan_free(self.raw.as_ptr()) }
                }
            }
        }
    };
}
/// The `Plan52` twins live in their own modules here (`plan52::native64::Plan52` ...) because a macro-generated
/// module cannot be re-opened; a fork of the reference would place each type next to its `Plan32`.
pub mod plan52 {
    use super::*;
    plan52!(native32, u32, 32, false, 2);
    plan52!(native64, u64, 64, false, 3);
    plan52!(native_binary32, u32, 32, true, 1);
    plan52!(native_binary64, u64, 64, true, 2);
}
