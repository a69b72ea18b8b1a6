//! `prime64::Plan` (reference: src/prime64.rs:222-1129).
use crate::ffi;
use core::ptr::NonNull;

/// Negacyclic NTT plan for 64-bit primes.
pub struct Plan {
    raw: NonNull<ffi::Prime64Plan>,
    device: i32,
}

/// `prime64::Solinas` (src/prime64/generic_solinas.rs:36-40)
pub struct Solinas;
impl Solinas {
    pub const P: u64 = 0xFFFF_FFFF_0000_0001;
}
// handles are immutable after creation; calls on distinct buffers may run concurrently (the host-slice calls
// serialise on the plan's staging arena)
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// src/prime64.rs:704 -- `None` if `polynomial_size` is not a power of two >= 16, `modulus` is not prime, or no
    /// 2n-th root of unity exists; panics for `modulus <= 1` like `Div64::new`.
    pub fn try_new(polynomial_size: usize, modulus: u64) -> Option<Self> {
        Self::try_new_on(polynomial_size, modulus, 0)
    }
    /// Extension: the plan's tables live on CUDA device `device`.
    pub fn try_new_on(polynomial_size: usize, modulus: u64, device: i32) -> Option<Self> {
        let mut raw = core::ptr::null_mut();
        ffi::plan_status(unsafe { ffi::cntt_prime64_plan_new(polynomial_size, modulus, device, &mut raw) })?;
        Some(Self { raw: NonNull::new(raw)?, device })
    }
    /// src/prime64.rs:775
    #[inline]
    pub fn ntt_size(&self) -> usize {
        unsafe { ffi::cntt_prime64_ntt_size(self.raw.as_ptr()) }
    }
    /// src/prime64.rs:781
    #[inline]
    pub fn modulus(&self) -> u64 {
        unsafe { ffi::cntt_prime64_modulus(self.raw.as_ptr()) }
    }
    /// src/prime64.rs:794 -- natural order in, bit-reversed order out, values in `[0, p)`.
    pub fn fwd(&self, buf: &mut [u64]) {
        assert_eq!(buf.len(), self.ntt_size());
        ffi::check(unsafe { ffi::cntt_prime64_fwd_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), 1) });
    }
    /// src/prime64.rs:872 -- bit-reversed in, natural out, not normalised (`inv(fwd(x)) == n * x`).
    pub fn inv(&self, buf: &mut [u64]) {
        assert_eq!(buf.len(), self.ntt_size());
        ffi::check(unsafe { ffi::cntt_prime64_inv_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), 1) });
    }
    /// src/prime64.rs:947 -- `lhs[i] = lhs[i] * rhs[i] / n mod p`; zip-truncates like the reference.
    pub fn mul_assign_normalize(&self, lhs: &mut [u64], rhs: &[u64]) {
        let n = lhs.len().min(rhs.len()) & !1;
        ffi::check(unsafe { ffi::cntt_prime64_mul_assign_normalize_host(self.raw.as_ptr(), lhs.as_mut_ptr(), rhs.as_ptr(), n) });
    }
    /// src/prime64.rs:1040
    pub fn normalize(&self, values: &mut [u64]) {
        let n = values.len() & !1;
        ffi::check(unsafe { ffi::cntt_prime64_normalize_host(self.raw.as_ptr(), values.as_mut_ptr(), n) });
    }
    /// src/prime64.rs:1092 -- `acc[i] += lhs[i] * rhs[i] mod p`.
    pub fn mul_accumulate(&self, acc: &mut [u64], lhs: &[u64], rhs: &[u64]) {
        let n = acc.len().min(lhs.len()).min(rhs.len()) & !1;
        ffi::check(unsafe { ffi::cntt_prime64_mul_accumulate_host(self.raw.as_ptr(), acc.as_mut_ptr(), lhs.as_ptr(), rhs.as_ptr(), n) });
    }

    // ---- extensions: batches and device-resident buffers ------------------------------------------------------
    /// `buf` holds `buf.len() / n` polynomials back to back; one upload, one launch, one download.
    pub fn fwd_batch(&self, buf: &mut [u64]) {
        let n = self.ntt_size();
        assert_eq!(buf.len() % n, 0);
        ffi::check(unsafe { ffi::cntt_prime64_fwd_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), buf.len() / n) });
    }
    pub fn inv_batch(&self, buf: &mut [u64]) {
        let n = self.ntt_size();
        assert_eq!(buf.len() % n, 0);
        ffi::check(unsafe { ffi::cntt_prime64_inv_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), buf.len() / n) });
    }
    /// Device pointer (`batch * n` words) and CUDA stream, e.g. from `cudarc`; asynchronous.
    ///
    /// # Safety
    /// `d_buf` must be valid device memory on the plan's device for `batch * n` words until the stream has run.
    pub unsafe fn fwd_device(&self, d_buf: *mut u64, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime64_fwd(self.raw.as_ptr(), d_buf, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn inv_device(&self, d_buf: *mut u64, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime64_inv(self.raw.as_ptr(), d_buf, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]; `nwords` words per operand
    pub unsafe fn mul_accumulate_device(&self, d_acc: *mut u64, d_lhs: *const u64, d_rhs: *const u64, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime64_mul_accumulate(self.raw.as_ptr(), d_acc, d_lhs, d_rhs, nwords, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn mul_assign_normalize_device(&self, d_lhs: *mut u64, d_rhs: *const u64, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime64_mul_assign_normalize(self.raw.as_ptr(), d_lhs, d_rhs, nwords, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn normalize_device(&self, d_values: *mut u64, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime64_normalize(self.raw.as_ptr(), d_values, nwords, stream));
    }
}

impl Drop for Plan {
    fn drop(&mut self) {
        unsafe { ffi::cntt_prime64_plan_free(self.raw.as_ptr()) }
    }
}
/// The reference derives `Clone` (a deep copy of the tables); a plan is a pure function of `(n, p)`.
impl Clone for Plan {
    fn clone(&self) -> Self {
        Self::try_new_on(self.ntt_size(), self.modulus(), self.device).expect("a plan that exists can be rebuilt")
    }
}
impl core::fmt::Debug for Plan {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        f.debug_struct("Plan").field("ntt_size", &self.ntt_size()).field("modulus", &self.modulus()).finish()
    }
}
