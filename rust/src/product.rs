//! `product::Plan` (reference: src/product.rs:139-967): negacyclic NTT plan for a modulus that is a product of distinct
//! primes.  NTT-domain buffers keep the reference's packed layout (src/product.rs:261-278), so they are interchangeable
//! with buffers produced by the CPU crate.
use crate::ffi;
use core::ptr::NonNull;

/// src/product.rs:10-14
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum FwdMode {
    Generic,
    /// every coefficient is a centred representative of magnitude at most the bound
    Bounded(u64),
}
/// src/product.rs:16-20
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum InvMode {
    Replace,
    Accumulate,
}

pub struct Plan {
    raw: NonNull<ffi::ProductPlan>,
}
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// src/product.rs:152 -- `factors`: the prime factors of `modulus` (ones are ignored); `None` if a factor is zero
    /// or repeated, the product is not `modulus`, or a factor admits no plan of this size.
    pub fn try_new(polynomial_size: usize, modulus: u64, factors: impl IntoIterator<Item = u64>) -> Option<Self> {
        let f: Vec<u64> = factors.into_iter().collect();
        let mut raw = core::ptr::null_mut();
        ffi::plan_status(unsafe { ffi::cntt_product_plan_new(polynomial_size, modulus, f.as_ptr(), f.len(), 0, &mut raw) })?;
        Some(Self { raw: NonNull::new(raw)? })
    }
    /// src/product.rs:254
    pub fn ntt_size(&self) -> usize {
        unsafe { ffi::cntt_product_ntt_size(self.raw.as_ptr()) }
    }
    /// src/product.rs:260
    pub fn modulus(&self) -> u64 {
        unsafe { ffi::cntt_product_modulus(self.raw.as_ptr()) }
    }
    /// src/product.rs:265
    pub fn ntt_domain_len(&self) -> usize {
        unsafe { ffi::cntt_product_ntt_domain_len(self.raw.as_ptr()) }
    }
    fn mode(mode: FwdMode) -> (i32, u64) {
        match mode {
            FwdMode::Generic => (0, 0),
            FwdMode::Bounded(b) => (1, b),
        }
    }
    /// src/product.rs:276
    pub fn fwd(&self, ntt: &mut [u64], standard: &[u64], mode: FwdMode) {
        let (m, b) = Self::mode(mode);
        ffi::check(unsafe { ffi::cntt_product_fwd_host(self.raw.as_ptr(), ntt.as_mut_ptr(), standard.as_ptr(), ntt.len(), standard.len(), m, b, 1) });
    }
    /// src/product.rs:355 -- `ntt` is clobbered like in the reference (it comes back holding the inverse transforms).
    pub fn inv(&self, standard: &mut [u64], ntt: &mut [u64], mode: InvMode) {
        let m = matches!(mode, InvMode::Accumulate) as i32;
        ffi::check(unsafe { ffi::cntt_product_inv_host(self.raw.as_ptr(), standard.as_mut_ptr(), ntt.as_mut_ptr(), standard.len(), ntt.len(), m, 1) });
    }
    /// src/product.rs:884
    pub fn mul_assign_normalize(&self, lhs: &mut [u64], rhs: &[u64]) {
        assert_eq!(lhs.len(), rhs.len());
        ffi::check(unsafe { ffi::cntt_product_mul_assign_normalize_host(self.raw.as_ptr(), lhs.as_mut_ptr(), rhs.as_ptr(), lhs.len(), 1) });
    }
    /// src/product.rs:918
    pub fn normalize(&self, values: &mut [u64]) {
        ffi::check(unsafe { ffi::cntt_product_normalize_host(self.raw.as_ptr(), values.as_mut_ptr(), values.len(), 1) });
    }
    /// src/product.rs:935
    pub fn mul_accumulate(&self, acc: &mut [u64], lhs: &[u64], rhs: &[u64]) {
        assert_eq!(acc.len(), lhs.len());
        assert_eq!(acc.len(), rhs.len());
        ffi::check(unsafe { ffi::cntt_product_mul_accumulate_host(self.raw.as_ptr(), acc.as_mut_ptr(), lhs.as_ptr(), rhs.as_ptr(), acc.len(), 1) });
    }

    // ---- extensions: device-resident batches (polynomial b at `d_standard[b * n ..]` / `d_ntt[b * ntt_domain_len ..]`) ----
    /// # Safety
    /// device pointers on device 0 valid for `batch` polynomials until the stream has run
    pub unsafe fn fwd_device(&self, d_ntt: *mut u64, d_standard: *const u64, mode: FwdMode, batch: usize, stream: ffi::Stream) {
        let (m, b) = Self::mode(mode);
        ffi::check(ffi::cntt_product_fwd(self.raw.as_ptr(), d_ntt, d_standard, m, b, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn inv_device(&self, d_standard: *mut u64, d_ntt: *mut u64, mode: InvMode, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_product_inv(self.raw.as_ptr(), d_standard, d_ntt, matches!(mode, InvMode::Accumulate) as i32, batch, stream));
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn mul_accumulate_device(&self, d_acc: *mut u64, d_lhs: *const u64, d_rhs: *const u64, batch: usize, stream: ffi::Stream) {
        ffi::check(ffi::cntt_product_mul_accumulate(self.raw.as_ptr(), d_acc, d_lhs, d_rhs, batch, stream));
    }
}
impl Drop for Plan {
    fn drop(&mut self) {
        unsafe { ffi::cntt_product_plan_free(self.raw.as_ptr()) }
    }
}
