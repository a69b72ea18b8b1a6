//! `prime32::Plan` (reference: src/prime32.rs:602-928).
use crate::ffi;
use core::ptr::NonNull;

/// Negacyclic NTT plan for 32-bit primes.
pub struct Plan {
    raw: NonNull<ffi::Prime32Plan>,
    device: i32,
}
// handles are immutable after creation; calls on distinct buffers may run concurrently (the host-slice calls
// serialise on the plan's staging arena)
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// src/prime32.rs:630 -- `None` if `polynomial_size` is not a power of two >= 32, `modulus` is not prime, or no
    /// 2n-th root of unity exists; panics for `modulus <= 1` like `Div32::new`.
    pub fn try_new(polynomial_size: usize, modulus: u32) -> Option<Self> {
        Self::try_new_on(polynomial_size, modulus, 0)
    }
    /// Extension: the plan's tables live on CUDA device `device`.
    pub fn try_new_on(polynomial_size: usize, modulus: u32, device: i32) -> Option<Self> {
        let mut raw = core::ptr::null_mut();
        ffi::plan_status(unsafe { ffi::cntt_prime32_plan_new(polynomial_size, modulus, device, &mut raw) })?;
        Some(Self { raw: NonNull::new(raw)?, device })
    }
    /// src/prime32.rs:694
    #[inline]
    pub fn ntt_size(&self) -> usize {
        unsafe { ffi::cntt_prime32_ntt_size(self.raw.as_ptr()) }
    }
    /// src/prime32.rs:700
    #[inline]
    pub fn modulus(&self) -> u32 {
        unsafe { ffi::cntt_prime32_modulus(self.raw.as_ptr()) }
    }
    /// src/prime32.rs:709 -- natural order in, bit-reversed order out, values in `[0, p)`.
    pub fn fwd(&self, buf: &mut [u32]) {
        assert_eq!(buf.len(), self.ntt_size());
        ffi::check(unsafe { ffi::cntt_prime32_fwd_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), 1) });
    }
    /// src/prime32.rs:762 -- bit-reversed in, natural out, not normalised (`inv(fwd(x)) == n * x`).
    pub fn inv(&self, buf: &mut [u32]) {
        assert_eq!(buf.len(), self.ntt_size());
        ffi::check(unsafe { ffi::cntt_prime32_inv_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), 1) });
    }
    /// src/prime32.rs:812 -- `lhs[i] = lhs[i] * rhs[i] / n mod p`; zip-truncates like the reference.
    pub fn mul_assign_normalize(&self, lhs: &mut [u32], rhs: &[u32]) {
        let n = lhs.len().min(rhs.len()) & !3;
        ffi::check(unsafe { ffi::cntt_prime32_mul_assign_normalize_host(self.raw.as_ptr(), lhs.as_mut_ptr(), rhs.as_ptr(), n) });
    }
    /// src/prime32.rs:868
    pub fn normalize(&self, values: &mut [u32]) {
        let n = values.len() & !3;
        ffi::check(unsafe { ffi::cntt_prime32_normalize_host(self.raw.as_ptr(), values.as_mut_ptr(), n) });
    }
    /// src/prime32.rs:905 -- `acc[i] += lhs[i] * rhs[i] mod p`.
    pub fn mul_accumulate(&self, acc: &mut [u32], lhs: &[u32], rhs: &[u32]) {
        let n = acc.len().min(lhs.len()).min(rhs.len()) & !3;
        ffi::check(unsafe { ffi::cntt_prime32_mul_accumulate_host(self.raw.as_ptr(), acc.as_mut_ptr(), lhs.as_ptr(), rhs.as_ptr(), n) });
    }

    // ---- extensions: batches and device-resident buffers ------------------------------------------------------
    /// `buf` holds `buf.len() / n` polynomials back to back; one upload, one launch, one download.
    pub fn fwd_batch(&self, buf: &mut [u32]) {
        let n = self.ntt_size();
        assert_eq!(buf.len() % n, 0);
        ffi::check(unsafe { ffi::cntt_prime32_fwd_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), buf.len() / n) });
    }
    pub fn inv_batch(&self, buf: &mut [u32]) {
        let n = self.ntt_size();
        assert_eq!(buf.len() % n, 0);
        ffi::check(unsafe { ffi::cntt_prime32_inv_host(self.raw.as_ptr(), buf.as_mut_ptr(), buf.len(), buf.len() / n) });
    }
    /// Device pointer (`batch * n` words) and CUDA stream, e.g. from `cudarc`; asynchronous.
    ///
    /// # Safety
    /// `d_buf` must be valid device memory on the plan's device for `batch * n` words until the stream has run.
    pub unsafe fn fwd_device(&self, d_buf: *mut u32, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime32_fwd(self.raw.as_ptr(), d_buf, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn inv_device(&self, d_buf: *mut u32, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime32_inv(self.raw.as_ptr(), d_buf, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]; `nwords` words per operand
    pub unsafe fn mul_accumulate_device(&self, d_acc: *mut u32, d_lhs: *const u32, d_rhs: *const u32, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime32_mul_accumulate(self.raw.as_ptr(), d_acc, d_lhs, d_rhs, nwords, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn mul_assign_normalize_device(&self, d_lhs: *mut u32, d_rhs: *const u32, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime32_mul_assign_normalize(self.raw.as_ptr(), d_lhs, d_rhs, nwords, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn normalize_device(&self, d_values: *mut u32, nwords: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_prime32_normalize(self.raw.as_ptr(), d_values, nwords, stream));
    }
}

impl Drop for Plan {
    fn drop(&mut self) {
        unsafe { ffi::cntt_prime32_plan_free(self.raw.as_ptr()) }
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
