//! `extern "C"` declarations, one to one with include/cntt_b200.h.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(Prime32Plan, Prime64Plan, NativePlan, Native52Plan, ProductPlan);

pub const OK: c_int = 0;
pub const INVALID_SIZE: c_int = 1; // try_new -> None
pub const INVALID_MODULUS: c_int = 2; // try_new -> None
pub const NO_ROOT: c_int = 3; // try_new -> None
pub const LENGTH_MISMATCH: c_int = 4; // assert_eq!(buf.len(), n) -> panic
pub const CUDA_ERROR: c_int = 5;
pub const NULL_POINTER: c_int = 6;
pub const UNSUPPORTED: c_int = 7;
pub const PANIC_MODULUS: c_int = 8; // Div32::new / Div64::new assert (divisor > 1) -> panic
pub const MISALIGNED: c_int = 9; // a device batch pointer is not 16-byte aligned

pub type Stream = *mut c_void; // cudaStream_t

extern "C" {
    pub fn cntt_status_string(status: c_int) -> *const c_char;
    pub fn cntt_last_cuda_error() -> *const c_char;
    pub fn cntt_version() -> *const c_char;
    pub fn cntt_is_prime64(n: u64) -> c_int;
    pub fn cntt_largest_prime_in_arithmetic_progression64(factor: u64, offset: u64, lo: u64, hi: u64, out: *mut u64) -> c_int;
    pub fn cntt_find_primitive_root64(p: u64, degree: u64, out: *mut u64) -> c_int;
    pub fn cntt_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int; // pinned host memory for the *_host calls
    pub fn cntt_host_free(ptr: *mut c_void) -> c_int;

    // ---- prime32::Plan ----
    pub fn cntt_prime32_plan_new(n: usize, p: u32, device: c_int, out: *mut *mut Prime32Plan) -> c_int;
    pub fn cntt_prime32_plan_free(plan: *mut Prime32Plan);
    pub fn cntt_prime32_ntt_size(plan: *const Prime32Plan) -> usize;
    pub fn cntt_prime32_modulus(plan: *const Prime32Plan) -> u32;
    pub fn cntt_prime32_fwd(plan: *const Prime32Plan, d_buf: *mut u32, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_prime32_inv(plan: *const Prime32Plan, d_buf: *mut u32, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_prime32_mul_assign_normalize(plan: *const Prime32Plan, d_lhs: *mut u32, d_rhs: *const u32, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime32_normalize(plan: *const Prime32Plan, d_values: *mut u32, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime32_mul_accumulate(plan: *const Prime32Plan, d_acc: *mut u32, d_lhs: *const u32, d_rhs: *const u32, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime32_fwd_host(plan: *const Prime32Plan, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime32_inv_host(plan: *const Prime32Plan, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime32_fwd_inv_host(plan: *const Prime32Plan, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime32_mul_assign_normalize_host(plan: *const Prime32Plan, lhs: *mut u32, rhs: *const u32, nwords: usize) -> c_int;
    pub fn cntt_prime32_normalize_host(plan: *const Prime32Plan, values: *mut u32, nwords: usize) -> c_int;
    pub fn cntt_prime32_mul_accumulate_host(plan: *const Prime32Plan, acc: *mut u32, lhs: *const u32, rhs: *const u32, nwords: usize) -> c_int;

    // ---- one host batch over several GPUs (extension) ----
    pub fn cntt_prime32_fwd_host_multi(plans: *const *const Prime32Plan, nplans: c_int, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime32_inv_host_multi(plans: *const *const Prime32Plan, nplans: c_int, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime32_fwd_inv_host_multi(plans: *const *const Prime32Plan, nplans: c_int, h_buf: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_fwd_host_multi(plans: *const *const Prime64Plan, nplans: c_int, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_inv_host_multi(plans: *const *const Prime64Plan, nplans: c_int, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_fwd_inv_host_multi(plans: *const *const Prime64Plan, nplans: c_int, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_native_polymul_host_multi(plans: *const *const NativePlan, nplans: c_int, h_prod: *mut c_void, h_lhs: *const c_void, h_rhs: *const c_void, len: usize, batch: usize) -> c_int;

    // ---- prime64::Plan ----
    pub fn cntt_prime64_plan_new(n: usize, p: u64, device: c_int, out: *mut *mut Prime64Plan) -> c_int;
    pub fn cntt_prime64_plan_free(plan: *mut Prime64Plan);
    pub fn cntt_prime64_ntt_size(plan: *const Prime64Plan) -> usize;
    pub fn cntt_prime64_modulus(plan: *const Prime64Plan) -> u64;
    pub fn cntt_prime64_fwd(plan: *const Prime64Plan, d_buf: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_prime64_inv(plan: *const Prime64Plan, d_buf: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_prime64_mul_assign_normalize(plan: *const Prime64Plan, d_lhs: *mut u64, d_rhs: *const u64, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime64_normalize(plan: *const Prime64Plan, d_values: *mut u64, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime64_mul_accumulate(plan: *const Prime64Plan, d_acc: *mut u64, d_lhs: *const u64, d_rhs: *const u64, nwords: usize, stream: Stream) -> c_int;
    pub fn cntt_prime64_fwd_host(plan: *const Prime64Plan, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_inv_host(plan: *const Prime64Plan, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_fwd_inv_host(plan: *const Prime64Plan, h_buf: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_prime64_mul_assign_normalize_host(plan: *const Prime64Plan, lhs: *mut u64, rhs: *const u64, nwords: usize) -> c_int;
    pub fn cntt_prime64_normalize_host(plan: *const Prime64Plan, values: *mut u64, nwords: usize) -> c_int;
    pub fn cntt_prime64_mul_accumulate_host(plan: *const Prime64Plan, acc: *mut u64, lhs: *const u64, rhs: *const u64, nwords: usize) -> c_int;

    // ---- native{32,64,128} / native_binary{32,64,128} ::Plan32 ----
    pub fn cntt_native_plan_new(n: usize, word_bits: c_int, binary: c_int, device: c_int, out: *mut *mut NativePlan) -> c_int;
    pub fn cntt_native_plan_new_ext(n: usize, word_bits: c_int, binary: c_int, device: c_int, out: *mut *mut NativePlan) -> c_int;
    pub fn cntt_native_plan_free(plan: *mut NativePlan);
    pub fn cntt_native_ntt_size(plan: *const NativePlan) -> usize;
    pub fn cntt_native_num_primes(plan: *const NativePlan) -> c_int;
    pub fn cntt_native_prime(plan: *const NativePlan, i: c_int) -> u32;
    pub fn cntt_native_fwd(plan: *const NativePlan, d_value: *const c_void, d_mod_p: *mut u32, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native_fwd_binary(plan: *const NativePlan, d_value: *const c_void, d_mod_p: *mut u32, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native_inv(plan: *const NativePlan, d_value: *mut c_void, d_mod_p: *mut u32, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native_fwd_host(plan: *const NativePlan, h_value: *const c_void, h_mod_p: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_native_fwd_binary_host(plan: *const NativePlan, h_value: *const c_void, h_mod_p: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_native_inv_host(plan: *const NativePlan, h_value: *mut c_void, h_mod_p: *mut u32, len: usize, batch: usize) -> c_int;
    pub fn cntt_native_polymul(plan: *const NativePlan, d_prod: *mut c_void, d_lhs: *const c_void, d_rhs: *const c_void, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native_polymul_ntt_rhs(plan: *const NativePlan, d_prod: *mut c_void, d_lhs: *const c_void, d_rhs_planes: *const u32, rhs_batch: usize, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native_polymul_host(plan: *const NativePlan, h_prod: *mut c_void, h_lhs: *const c_void, h_rhs: *const c_void, len: usize, batch: usize) -> c_int;

    // ---- Plan52 twins ----
    pub fn cntt_native52_plan_new(n: usize, word_bits: c_int, binary: c_int, device: c_int, out: *mut *mut Native52Plan) -> c_int;
    pub fn cntt_native52_plan_free(plan: *mut Native52Plan);
    pub fn cntt_native52_ntt_size(plan: *const Native52Plan) -> usize;
    pub fn cntt_native52_num_primes(plan: *const Native52Plan) -> c_int;
    pub fn cntt_native52_prime(plan: *const Native52Plan, i: c_int) -> u64;
    pub fn cntt_native52_fwd(plan: *const Native52Plan, d_value: *const c_void, d_mod_p: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native52_fwd_binary(plan: *const Native52Plan, d_value: *const c_void, d_mod_p: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native52_inv(plan: *const Native52Plan, d_value: *mut c_void, d_mod_p: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native52_fwd_host(plan: *const Native52Plan, h_value: *const c_void, h_mod_p: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_native52_fwd_binary_host(plan: *const Native52Plan, h_value: *const c_void, h_mod_p: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_native52_inv_host(plan: *const Native52Plan, h_value: *mut c_void, h_mod_p: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_native52_polymul(plan: *const Native52Plan, d_prod: *mut c_void, d_lhs: *const c_void, d_rhs: *const c_void, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_native52_polymul_host(plan: *const Native52Plan, h_prod: *mut c_void, h_lhs: *const c_void, h_rhs: *const c_void, len: usize, batch: usize) -> c_int;

    // ---- product::Plan ----
    pub fn cntt_product_plan_new(n: usize, modulus: u64, factors: *const u64, nfactors: usize, device: c_int, out: *mut *mut ProductPlan) -> c_int;
    pub fn cntt_product_plan_free(plan: *mut ProductPlan);
    pub fn cntt_product_ntt_size(plan: *const ProductPlan) -> usize;
    pub fn cntt_product_modulus(plan: *const ProductPlan) -> u64;
    pub fn cntt_product_ntt_domain_len(plan: *const ProductPlan) -> usize;
    pub fn cntt_product_num_primes(plan: *const ProductPlan, count32: *mut c_int, count64: *mut c_int) -> c_int;
    pub fn cntt_product_prime(plan: *const ProductPlan, i: c_int) -> u64;
    pub fn cntt_product_fwd(plan: *const ProductPlan, d_ntt: *mut u64, d_standard: *const u64, mode: c_int, bound: u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_product_inv(plan: *const ProductPlan, d_standard: *mut u64, d_ntt: *mut u64, mode: c_int, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_product_mul_assign_normalize(plan: *const ProductPlan, d_lhs: *mut u64, d_rhs: *const u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_product_normalize(plan: *const ProductPlan, d_values: *mut u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_product_mul_accumulate(plan: *const ProductPlan, d_acc: *mut u64, d_lhs: *const u64, d_rhs: *const u64, batch: usize, stream: Stream) -> c_int;
    pub fn cntt_product_fwd_host(plan: *const ProductPlan, h_ntt: *mut u64, h_standard: *const u64, ntt_len: usize, standard_len: usize, mode: c_int, bound: u64, batch: usize) -> c_int;
    pub fn cntt_product_inv_host(plan: *const ProductPlan, h_standard: *mut u64, h_ntt: *mut u64, standard_len: usize, ntt_len: usize, mode: c_int, batch: usize) -> c_int;
    pub fn cntt_product_mul_assign_normalize_host(plan: *const ProductPlan, h_lhs: *mut u64, h_rhs: *const u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_product_normalize_host(plan: *const ProductPlan, h_values: *mut u64, len: usize, batch: usize) -> c_int;
    pub fn cntt_product_mul_accumulate_host(plan: *const ProductPlan, h_acc: *mut u64, h_lhs: *const u64, h_rhs: *const u64, len: usize, batch: usize) -> c_int;
}

/// `!= OK` from a compute call: length mismatch is the reference's `assert_eq!` panic, anything else is a
/// backend failure (there is no CPU fallback to hide it behind).
#[track_caller]
pub fn check(status: c_int) {
    match status {
        OK => (),
        LENGTH_MISMATCH => panic!("assertion `left == right` failed (slice length != polynomial size)"),
        CUDA_ERROR => {
            let msg = unsafe { core::ffi::CStr::from_ptr(cntt_last_cuda_error()) }.to_string_lossy().into_owned();
            panic!("cntt_b200: CUDA error: {msg}")
        }
        s => panic!("cntt_b200: status {s}"),
    }
}

/// Outcome of a `*_plan_new` call in the reference's terms.
#[track_caller]
pub fn plan_status(status: c_int) -> Option<()> {
    match status {
        OK => Some(()),
        INVALID_SIZE | INVALID_MODULUS | NO_ROOT | UNSUPPORTED => None,
        PANIC_MODULUS => panic!("assertion failed: divisor > 1"), // src/fastdiv.rs:49,99
        s => {
            check(s);
            None
        }
    }
}
