/*
 * cntt_b200.h -- C ABI of libcntt_b200.so: a B200 (sm_100a) batched negacyclic NTT library that is a
 * drop-in for the hot path of zama-ai/concrete-ntt v0.2.0.
 *
 * Every entry point replaces one public method of the reference crate; the citation next to it is the
 * reference definition (paths relative to the crate root).  The reference API is "one polynomial per
 * call on a host slice"; this ABI adds an explicit batch count and comes in two flavours:
 *
 *   cntt_<plan>_<op>(plan, device pointers..., batch, stream)        device-resident batches, async
 *   cntt_<plan>_<op>_host(plan, host pointers..., batch)            host slices, staged H2D/D2H inside
 *
 * Layout: polynomial-major contiguous -- element i of polynomial b is buf[b*n + i] (the reference's
 * slices concatenated).  Residue planes of the native plans: plane k of polynomial b is
 * mod_p[(k*batch + b)*n + i] (k-th `mod_pk` slice of the reference, concatenated over the batch).
 * 128-bit words are little-endian {lo:u64, hi:u64} pairs, 16-byte aligned (Rust u128 on x86-64).
 * Device batches must be 16-byte aligned (CNTT_MISALIGNED otherwise; n >= 16 words keeps every polynomial and residue
 * plane of an aligned batch aligned); 32-byte aligned batches (cudaMalloc gives 256) also get the 256-bit load / store
 * path of the transform kernels.
 *
 * Results are bit-identical to the reference on the same inputs (same primitive root, same
 * bit-reversed order, canonical residues, same centred CRT lift).
 *
 * All functions return a cntt_status.  Plans are immutable after creation and may be used from
 * several host threads on distinct streams/buffers (the reference's Plan is Send + Sync).
 * There is no CPU fallback: without a CUDA device every compute call returns CNTT_CUDA_ERROR.
 */
#ifndef CNTT_B200_H
#define CNTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cntt_status {
    CNTT_OK = 0,
    CNTT_INVALID_SIZE = 1,    /* try_new -> None: n too small / not a power of two (prime32.rs:635, prime64.rs:709) */
    CNTT_INVALID_MODULUS = 2, /* try_new -> None: p not prime.  p in {0,1}: the reference panics (fastdiv.rs:49,99); see cntt_status_is_panic */
    CNTT_NO_ROOT = 3,         /* try_new -> None: no primitive 2n-th root of unity mod p */
    CNTT_LENGTH_MISMATCH = 4, /* the reference's assert_eq!(buf.len(), n) */
    CNTT_CUDA_ERROR = 5,
    CNTT_NULL_POINTER = 6,
    CNTT_UNSUPPORTED = 7,
    CNTT_PANIC_MODULUS = 8,   /* p in {0,1}: the reference panics before validating */
    CNTT_MISALIGNED = 9       /* a device batch pointer is not 16-byte aligned (checked before any launch) */
} cntt_status;

const char* cntt_status_string(int status);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* cntt_last_cuda_error(void);
/* library / build identification, e.g. "cntt_b200 0.2 (sm_100a; concrete-ntt 0.2.0 semantics)" */
const char* cntt_version(void);

/* ---- plan-time helpers (host only) ------------------------------------------------------------- */
/* prime::is_prime64                                   src/prime.rs:76-126  */
int cntt_is_prime64(uint64_t n);
/* prime::largest_prime_in_arithmetic_progression64    src/prime.rs:130-180 ; returns 1 and *out on Some */
int cntt_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset, uint64_t lo, uint64_t hi, uint64_t* out);
/* roots::find_primitive_root64                        src/roots.rs:68-91   ; returns 1 and *root on Some */
int cntt_find_primitive_root64(uint64_t p, uint64_t degree, uint64_t* root);

/* pinned host memory for the *_host entry points (optional; any host pointer works) */
int cntt_host_alloc(void** ptr, size_t bytes);
int cntt_host_free(void* ptr);

/* ---- prime32::Plan                                  src/prime32.rs:602-928 ------------------------ */
typedef struct cntt_prime32_plan cntt_prime32_plan;
/* Plan::try_new(polynomial_size, modulus)             src/prime32.rs:630-686 ; device = CUDA ordinal */
int cntt_prime32_plan_new(size_t n, uint32_t p, int device, cntt_prime32_plan** out);
void cntt_prime32_plan_free(cntt_prime32_plan* plan);
/* Plan::ntt_size / Plan::modulus                      src/prime32.rs:694-703 */
size_t cntt_prime32_ntt_size(const cntt_prime32_plan* plan);
uint32_t cntt_prime32_modulus(const cntt_prime32_plan* plan);
/* Plan::fwd / Plan::inv (in place)                    src/prime32.rs:709-755, 762-808 */
int cntt_prime32_fwd(const cntt_prime32_plan* plan, uint32_t* d_buf, size_t batch, void* stream);
int cntt_prime32_inv(const cntt_prime32_plan* plan, uint32_t* d_buf, size_t batch, void* stream);
/* Plan::mul_assign_normalize(lhs, rhs)                src/prime32.rs:812-864 ; nwords = words in each stream */
int cntt_prime32_mul_assign_normalize(const cntt_prime32_plan* plan, uint32_t* d_lhs, const uint32_t* d_rhs, size_t nwords, void* stream);
/* Plan::normalize(values)                             src/prime32.rs:868-902 */
int cntt_prime32_normalize(const cntt_prime32_plan* plan, uint32_t* d_values, size_t nwords, void* stream);
/* Plan::mul_accumulate(acc, lhs, rhs)                 src/prime32.rs:905-927 */
int cntt_prime32_mul_accumulate(const cntt_prime32_plan* plan, uint32_t* d_acc, const uint32_t* d_lhs, const uint32_t* d_rhs, size_t nwords, void* stream);
/* host-slice flavours: len = number of words in h_buf; must equal n*batch (CNTT_LENGTH_MISMATCH otherwise) */
int cntt_prime32_fwd_host(const cntt_prime32_plan* plan, uint32_t* h_buf, size_t len, size_t batch);
int cntt_prime32_inv_host(const cntt_prime32_plan* plan, uint32_t* h_buf, size_t len, size_t batch);
/* fwd followed by inv on the device between one upload and one download (round trip = N * x) */
int cntt_prime32_fwd_inv_host(const cntt_prime32_plan* plan, uint32_t* h_buf, size_t len, size_t batch);
int cntt_prime32_mul_assign_normalize_host(const cntt_prime32_plan* plan, uint32_t* h_lhs, const uint32_t* h_rhs, size_t nwords);
int cntt_prime32_normalize_host(const cntt_prime32_plan* plan, uint32_t* h_values, size_t nwords);
int cntt_prime32_mul_accumulate_host(const cntt_prime32_plan* plan, uint32_t* h_acc, const uint32_t* h_lhs, const uint32_t* h_rhs, size_t nwords);

/* ---- prime64::Plan                                  src/prime64.rs:222-1129 ----------------------- */
typedef struct cntt_prime64_plan cntt_prime64_plan;
/* prime64::Solinas::P = 2^64 - 2^32 + 1               src/prime64/generic_solinas.rs:36-40 */
#define CNTT_SOLINAS_P 0xFFFFFFFF00000001ull
/* Plan::try_new                                       src/prime64.rs:704-771 */
int cntt_prime64_plan_new(size_t n, uint64_t p, int device, cntt_prime64_plan** out);
void cntt_prime64_plan_free(cntt_prime64_plan* plan);
size_t cntt_prime64_ntt_size(const cntt_prime64_plan* plan);                 /* src/prime64.rs:779-781 */
uint64_t cntt_prime64_modulus(const cntt_prime64_plan* plan);                /* src/prime64.rs:785-787 */
int cntt_prime64_fwd(const cntt_prime64_plan* plan, uint64_t* d_buf, size_t batch, void* stream);   /* src/prime64.rs:794-865 */
int cntt_prime64_inv(const cntt_prime64_plan* plan, uint64_t* d_buf, size_t batch, void* stream);   /* src/prime64.rs:872-943 */
int cntt_prime64_mul_assign_normalize(const cntt_prime64_plan* plan, uint64_t* d_lhs, const uint64_t* d_rhs, size_t nwords, void* stream); /* src/prime64.rs:947-1032 */
int cntt_prime64_normalize(const cntt_prime64_plan* plan, uint64_t* d_values, size_t nwords, void* stream);                                 /* src/prime64.rs:1036-1082 */
int cntt_prime64_mul_accumulate(const cntt_prime64_plan* plan, uint64_t* d_acc, const uint64_t* d_lhs, const uint64_t* d_rhs, size_t nwords, void* stream); /* src/prime64.rs:1085-1128 */
int cntt_prime64_fwd_host(const cntt_prime64_plan* plan, uint64_t* h_buf, size_t len, size_t batch);
int cntt_prime64_inv_host(const cntt_prime64_plan* plan, uint64_t* h_buf, size_t len, size_t batch);
int cntt_prime64_fwd_inv_host(const cntt_prime64_plan* plan, uint64_t* h_buf, size_t len, size_t batch);
int cntt_prime64_mul_assign_normalize_host(const cntt_prime64_plan* plan, uint64_t* h_lhs, const uint64_t* h_rhs, size_t nwords);
int cntt_prime64_normalize_host(const cntt_prime64_plan* plan, uint64_t* h_values, size_t nwords);
int cntt_prime64_mul_accumulate_host(const cntt_prime64_plan* plan, uint64_t* h_acc, const uint64_t* h_lhs, const uint64_t* h_rhs, size_t nwords);

/* ---- one host batch over several GPUs (extension) ---------------------------------------------------
 * plans[g] = the same plan (n, p) built on device g.  The batch is cut into nplans contiguous shards, each staged and
 * transformed on its own device concurrently (one host thread and one staging pipeline per device, no collective --
 * polynomials are independent); results land in place.  This is the single-process entry point for a box of GPUs;
 * the per-device calls above remain for callers that do their own sharding (one process per GPU, shard.py). */
int cntt_prime32_fwd_host_multi(const cntt_prime32_plan* const* plans, int nplans, uint32_t* h_buf, size_t len, size_t batch);
int cntt_prime32_inv_host_multi(const cntt_prime32_plan* const* plans, int nplans, uint32_t* h_buf, size_t len, size_t batch);
int cntt_prime32_fwd_inv_host_multi(const cntt_prime32_plan* const* plans, int nplans, uint32_t* h_buf, size_t len, size_t batch);
int cntt_prime64_fwd_host_multi(const cntt_prime64_plan* const* plans, int nplans, uint64_t* h_buf, size_t len, size_t batch);
int cntt_prime64_inv_host_multi(const cntt_prime64_plan* const* plans, int nplans, uint64_t* h_buf, size_t len, size_t batch);
int cntt_prime64_fwd_inv_host_multi(const cntt_prime64_plan* const* plans, int nplans, uint64_t* h_buf, size_t len, size_t batch);

/* ---- native{32,64,128}::Plan32 and native_binary{32,64,128}::Plan32 --------------------------------
 * One handle type; `word_bits` in {32,64,128} and `binary` in {0,1} select the reference plan:
 *   native32::Plan32        (P0..P2)  src/native32.rs:8-12,335-433
 *   native64::Plan32        (P0..P4)  src/native64.rs:16-22,930-1070
 *   native128::Plan32       (P0..P9)  src/native128.rs:6-17,120-349
 *   native_binary32::Plan32 (P0..P1)  src/native_binary32.rs:11,187-263
 *   native_binary64::Plan32 (P0..P2)  src/native_binary64.rs:17-21,342-445
 *   native_binary128::Plan32(P0..P4)  src/native_binary128.rs:4-10,66-196
 * Words are passed as void*: uint32_t / uint64_t / 16-byte {lo,hi} according to word_bits.
 */
typedef struct cntt_native_plan cntt_native_plan;
/* Plan32::try_new(n)  -- CNTT_NO_ROOT when n > 32768 (P1 - 1 = 2^16 * odd, src/lib.rs:454) */
int cntt_native_plan_new(size_t n, int word_bits, int binary, int device, cntt_native_plan** out);
/* EXTENSION (no reference counterpart): the same plan kinds on the nine primes k*2^17+1 just below 2^30
 * (six of them are the reference's P0,P2,P3,P4,P8,P9), which admit n up to 65536 -- BASELINE.json configs[4],
 * native_binary64 N=65536, for which the reference's try_new returns None.  polymul results are independent of
 * the primes (exact product, wrapped), so wherever both constructors succeed their polymul outputs are
 * identical; fwd/inv residue planes are relative to cntt_native_prime(plan, i).  word_bits 128 non-binary needs
 * ten primes: CNTT_UNSUPPORTED. */
int cntt_native_plan_new_ext(size_t n, int word_bits, int binary, int device, cntt_native_plan** out);
void cntt_native_plan_free(cntt_native_plan* plan);
size_t cntt_native_ntt_size(const cntt_native_plan* plan);
int cntt_native_num_primes(const cntt_native_plan* plan);
/* Plan32::ntt_i(): modulus of the i-th prime32 sub-plan (src/native64.rs:950-969) */
uint32_t cntt_native_prime(const cntt_native_plan* plan, int i);
/* Plan32::fwd(value, mod_p0..)        value: batch*n words in, mod_p: num_primes planes out */
int cntt_native_fwd(const cntt_native_plan* plan, const void* d_value, uint32_t* d_mod_p, size_t batch, void* stream);
/* Plan32::fwd_binary(value, mod_p0..) binary plans only; CNTT_UNSUPPORTED otherwise */
int cntt_native_fwd_binary(const cntt_native_plan* plan, const void* d_value, uint32_t* d_mod_p, size_t batch, void* stream);
/* Plan32::inv(value, mod_p0..)        mod_p planes are clobbered, like in the reference */
int cntt_native_inv(const cntt_native_plan* plan, void* d_value, uint32_t* d_mod_p, size_t batch, void* stream);
/* host-slice flavours of fwd / fwd_binary / inv (the reference's call shape; `batch` polynomials concatenated, len =
 * words in value = n * batch, plane k of polynomial b at h_mod_p[(k * batch + b) * n]); inv also returns the clobbered
 * planes, which the reference leaves holding the un-normalised inverse transforms (src/native64.rs:1001-1014) */
int cntt_native_fwd_host(const cntt_native_plan* plan, const void* h_value, uint32_t* h_mod_p, size_t len, size_t batch);
int cntt_native_fwd_binary_host(const cntt_native_plan* plan, const void* h_value, uint32_t* h_mod_p, size_t len, size_t batch);
int cntt_native_inv_host(const cntt_native_plan* plan, void* h_value, uint32_t* h_mod_p, size_t len, size_t batch);
/* Plan32::negacyclic_polymul(prod, lhs, rhs) */
int cntt_native_polymul(const cntt_native_plan* plan, void* d_prod, const void* d_lhs, const void* d_rhs, size_t batch, void* stream);
/* EXTENSION: the same product with rhs already in the NTT domain -- d_rhs_planes as written by cntt_native_fwd (fwd_binary for the
 * binary plans) for rhs_batch polynomials, rhs_batch == batch or 1 (one key shared by the whole batch).  One of the three transforms
 * per prime disappears; this is the shape of a TFHE external product with the key kept transformed (src/prime32.rs:905-927 is the
 * per-prime step a reference caller composes).  256 <= n <= 4096, CNTT_UNSUPPORTED otherwise. */
int cntt_native_polymul_ntt_rhs(const cntt_native_plan* plan, void* d_prod, const void* d_lhs, const uint32_t* d_rhs_planes, size_t rhs_batch, size_t batch, void* stream);
/* host-slice flavour: len = words in each of prod/lhs/rhs; must equal n*batch */
int cntt_native_polymul_host(const cntt_native_plan* plan, void* h_prod, const void* h_lhs, const void* h_rhs, size_t len, size_t batch);
/* the same over several GPUs: plans[g] = the same plan kind and n on device g (see cntt_prime32_fwd_host_multi) */
int cntt_native_polymul_host_multi(const cntt_native_plan* const* plans, int nplans, void* h_prod, const void* h_lhs, const void* h_rhs, size_t len, size_t batch);

/* ---- Plan52 twins: native32 / native64 / native_binary32 / native_binary64 ::Plan52 ----------------------------
 * (src/native32.rs:19,435-496; native64.rs:29-34,1072-1165; native_binary32.rs:19,266-330; native_binary64.rs:29,447-521)
 * The same plans on 2 / 3 / 1 / 2 of the ~50-bit primes52 (src/lib.rs:598-652) with u64 residue planes transformed by
 * prime64 plans.  In the reference they exist only under feature = "nightly" and try_new is None without AVX-512 IFMA;
 * here they are always available.  word_bits in {32, 64}.  negacyclic_polymul returns exactly what Plan32 returns. */
typedef struct cntt_native52_plan cntt_native52_plan;
int cntt_native52_plan_new(size_t n, int word_bits, int binary, int device, cntt_native52_plan** out);
void cntt_native52_plan_free(cntt_native52_plan* plan);
size_t cntt_native52_ntt_size(const cntt_native52_plan* plan);
int cntt_native52_num_primes(const cntt_native52_plan* plan);
uint64_t cntt_native52_prime(const cntt_native52_plan* plan, int i);          /* Plan52::ntt_i().modulus() */
/* Plan52::fwd / fwd_binary / inv: value batch*n words, d_mod_p num_primes planes of batch*n u64 (plane k at k*batch*n) */
int cntt_native52_fwd(const cntt_native52_plan* plan, const void* d_value, uint64_t* d_mod_p, size_t batch, void* stream);
int cntt_native52_fwd_binary(const cntt_native52_plan* plan, const void* d_value, uint64_t* d_mod_p, size_t batch, void* stream);
int cntt_native52_inv(const cntt_native52_plan* plan, void* d_value, uint64_t* d_mod_p, size_t batch, void* stream);
/* host-slice flavours (the reference's call shape): len = words in value = n * batch, plane k of polynomial b at
 * h_mod_p[(k * batch + b) * n]; inv returns the clobbered planes too */
int cntt_native52_fwd_host(const cntt_native52_plan* plan, const void* h_value, uint64_t* h_mod_p, size_t len, size_t batch);
int cntt_native52_fwd_binary_host(const cntt_native52_plan* plan, const void* h_value, uint64_t* h_mod_p, size_t len, size_t batch);
int cntt_native52_inv_host(const cntt_native52_plan* plan, void* h_value, uint64_t* h_mod_p, size_t len, size_t batch);
int cntt_native52_polymul(const cntt_native52_plan* plan, void* d_prod, const void* d_lhs, const void* d_rhs, size_t batch, void* stream);
int cntt_native52_polymul_host(const cntt_native52_plan* plan, void* h_prod, const void* h_lhs, const void* h_rhs, size_t len, size_t batch);

/* ---- product::Plan (modulus = product of distinct primes, < 2^64)      src/product.rs:139-967 -------------
 * NTT-domain layout of one polynomial = the reference's (product.rs:261-278): the u32 planes (primes < 2^32,
 * ascending) bit-cast into the front of the u64 buffer, then the u64 planes; ntt_domain_len u64 words.  A batch is
 * the reference's slices concatenated: polynomial b at d_ntt + b * ntt_domain_len, d_standard + b * n.
 */
typedef struct cntt_product_plan cntt_product_plan;
/* Plan::try_new(polynomial_size, modulus, factors)    src/product.rs:152-251 ; != CNTT_OK <=> None */
int cntt_product_plan_new(size_t n, uint64_t modulus, const uint64_t* factors, size_t nfactors, int device, cntt_product_plan** out);
void cntt_product_plan_free(cntt_product_plan* plan);
size_t cntt_product_ntt_size(const cntt_product_plan* plan);        /* src/product.rs:254-257 */
uint64_t cntt_product_modulus(const cntt_product_plan* plan);       /* src/product.rs:260-263 */
size_t cntt_product_ntt_domain_len(const cntt_product_plan* plan);  /* src/product.rs:265-274 */
int cntt_product_num_primes(const cntt_product_plan* plan, int* count32, int* count64);
uint64_t cntt_product_prime(const cntt_product_plan* plan, int i);  /* i-th prime, ascending */
/* Plan::fwd(ntt, standard, mode)                      src/product.rs:276-353 ; mode 0 = FwdMode::Generic,
 * 1 = FwdMode::Bounded(bound) */
int cntt_product_fwd(const cntt_product_plan* plan, uint64_t* d_ntt, const uint64_t* d_standard, int mode, uint64_t bound, size_t batch, void* stream);
/* Plan::inv(standard, ntt, mode)                      src/product.rs:355-880 ; mode 0 = InvMode::Replace,
 * 1 = InvMode::Accumulate; d_ntt is clobbered like the reference's `ntt: &mut [u64]` */
int cntt_product_inv(const cntt_product_plan* plan, uint64_t* d_standard, uint64_t* d_ntt, int mode, size_t batch, void* stream);
/* Plan::mul_assign_normalize / normalize / mul_accumulate on NTT-domain buffers   src/product.rs:884-967 */
int cntt_product_mul_assign_normalize(const cntt_product_plan* plan, uint64_t* d_lhs, const uint64_t* d_rhs, size_t batch, void* stream);
int cntt_product_normalize(const cntt_product_plan* plan, uint64_t* d_values, size_t batch, void* stream);
int cntt_product_mul_accumulate(const cntt_product_plan* plan, uint64_t* d_acc, const uint64_t* d_lhs, const uint64_t* d_rhs, size_t batch, void* stream);
/* host-slice flavours: the reference's `&mut [u64]` call shape (lengths checked like its assert_eq!), `batch`
 * concatenated polynomials, synchronous, staged through the plan's device arena */
int cntt_product_fwd_host(const cntt_product_plan* plan, uint64_t* h_ntt, const uint64_t* h_standard, size_t ntt_len, size_t standard_len, int mode, uint64_t bound, size_t batch);
int cntt_product_inv_host(const cntt_product_plan* plan, uint64_t* h_standard, uint64_t* h_ntt, size_t standard_len, size_t ntt_len, int mode, size_t batch);
int cntt_product_mul_assign_normalize_host(const cntt_product_plan* plan, uint64_t* h_lhs, const uint64_t* h_rhs, size_t len, size_t batch);
int cntt_product_normalize_host(const cntt_product_plan* plan, uint64_t* h_values, size_t len, size_t batch);
int cntt_product_mul_accumulate_host(const cntt_product_plan* plan, uint64_t* h_acc, const uint64_t* h_lhs, const uint64_t* h_rhs, size_t len, size_t batch);

#ifdef __cplusplus
}
#endif
#endif /* CNTT_B200_H */
