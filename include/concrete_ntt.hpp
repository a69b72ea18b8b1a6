// concrete_ntt.hpp -- header-only C++17 mirror of the concrete-ntt Plan API over the C ABI
// (include/cntt_b200.h, libcntt_b200.so).  Same module / type / method names as the Rust crate; Rust's
// Option<Plan> becomes std::optional<Plan>, a Rust panic becomes concrete_ntt::Panic.
//
//   auto plan = concrete_ntt::prime32::Plan::try_new(1024, 1062862849u).value();
//   plan.fwd(buf.data(), buf.size());                       // one polynomial, host slice (reference shape)
//   plan.fwd_batch(host_ptr, batch);                        // `batch` polynomials, host slice
//   plan.fwd_device(dev_ptr, batch, stream);                // device-resident batch, asynchronous
//
// Reference: src/prime32.rs:627-928, src/prime64.rs:701-1129, src/native64.rs:930-1070 (and siblings).
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>

#include "cntt_b200.h"

namespace concrete_ntt {

struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

inline void check(int st)
{
    if (st == CNTT_OK) return;
    if (st == CNTT_LENGTH_MISMATCH || st == CNTT_PANIC_MODULUS) throw Panic(cntt_status_string(st));
    std::string msg = cntt_status_string(st);
    if (st == CNTT_CUDA_ERROR) msg += std::string(": ") + cntt_last_cuda_error();
    throw Error(st, msg);
}
inline bool is_none(int st) { return st == CNTT_INVALID_SIZE || st == CNTT_INVALID_MODULUS || st == CNTT_NO_ROOT; }

namespace prime {
inline bool is_prime64(uint64_t n) { return cntt_is_prime64(n) != 0; }
inline std::optional<uint64_t> largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset, uint64_t lo, uint64_t hi)
{
    uint64_t out;
    if (!cntt_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, &out)) return std::nullopt;
    return out;
}
} // namespace prime

#define CNTT_PRIME_PLAN(NS, BITS, WORD)                                                                                    \
    namespace NS {                                                                                                         \
    class Plan {                                                                                                           \
        cntt_prime##BITS##_plan* h_ = nullptr;                                                                             \
        explicit Plan(cntt_prime##BITS##_plan* h) : h_(h) {}                                                               \
                                                                                                                           \
      public:                                                                                                              \
        Plan(Plan&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}                                                      \
        Plan& operator=(Plan&& o) noexcept { std::swap(h_, o.h_); return *this; }                                          \
        Plan(const Plan&) = delete;                                                                                        \
        ~Plan() { cntt_prime##BITS##_plan_free(h_); }                                                                      \
        /* Plan::try_new(polynomial_size, modulus) -> Option<Plan> */                                                      \
        static std::optional<Plan> try_new(size_t polynomial_size, WORD modulus, int device = 0)                           \
        {                                                                                                                  \
            cntt_prime##BITS##_plan* h = nullptr;                                                                          \
            int st = cntt_prime##BITS##_plan_new(polynomial_size, modulus, device, &h);                                    \
            if (is_none(st)) return std::nullopt;                                                                          \
            check(st);                                                                                                     \
            return Plan(h);                                                                                                \
        }                                                                                                                  \
        size_t ntt_size() const { return cntt_prime##BITS##_ntt_size(h_); }                                                \
        WORD modulus() const { return cntt_prime##BITS##_modulus(h_); }                                                    \
        /* reference shape: one polynomial in a host slice; len must equal ntt_size() (assert_eq! in the crate) */         \
        void fwd(WORD* buf, size_t len) const { check(cntt_prime##BITS##_fwd_host(h_, buf, len, 1)); }                      \
        void inv(WORD* buf, size_t len) const { check(cntt_prime##BITS##_inv_host(h_, buf, len, 1)); }                      \
        void mul_assign_normalize(WORD* lhs, const WORD* rhs, size_t len) const { check(cntt_prime##BITS##_mul_assign_normalize_host(h_, lhs, rhs, len)); } \
        void normalize(WORD* values, size_t len) const { check(cntt_prime##BITS##_normalize_host(h_, values, len)); }       \
        void mul_accumulate(WORD* acc, const WORD* lhs, const WORD* rhs, size_t len) const { check(cntt_prime##BITS##_mul_accumulate_host(h_, acc, lhs, rhs, len)); } \
        /* batch extensions */                                                                                             \
        void fwd_batch(WORD* host, size_t batch) const { check(cntt_prime##BITS##_fwd_host(h_, host, ntt_size() * batch, batch)); } \
        void inv_batch(WORD* host, size_t batch) const { check(cntt_prime##BITS##_inv_host(h_, host, ntt_size() * batch, batch)); } \
        void fwd_device(WORD* dev, size_t batch, void* stream = nullptr) const { check(cntt_prime##BITS##_fwd(h_, dev, batch, stream)); } \
        void inv_device(WORD* dev, size_t batch, void* stream = nullptr) const { check(cntt_prime##BITS##_inv(h_, dev, batch, stream)); } \
        void mul_assign_normalize_device(WORD* lhs, const WORD* rhs, size_t nwords, void* stream = nullptr) const { check(cntt_prime##BITS##_mul_assign_normalize(h_, lhs, rhs, nwords, stream)); } \
        void normalize_device(WORD* v, size_t nwords, void* stream = nullptr) const { check(cntt_prime##BITS##_normalize(h_, v, nwords, stream)); } \
        void mul_accumulate_device(WORD* acc, const WORD* lhs, const WORD* rhs, size_t nwords, void* stream = nullptr) const { check(cntt_prime##BITS##_mul_accumulate(h_, acc, lhs, rhs, nwords, stream)); } \
        const cntt_prime##BITS##_plan* raw() const { return h_; }                                                          \
    };                                                                                                                     \
    }

CNTT_PRIME_PLAN(prime32, 32, uint32_t)
CNTT_PRIME_PLAN(prime64, 64, uint64_t)
#undef CNTT_PRIME_PLAN

namespace prime64 {
struct Solinas {
    static constexpr uint64_t P = CNTT_SOLINAS_P;
};
} // namespace prime64

// native{32,64,128}::Plan32 and native_binary{32,64,128}::Plan32.  Word = uint32_t / uint64_t /
// unsigned __int128 (little-endian {lo, hi}, the layout of Rust's u128 on x86-64).
template <int BITS, bool BINARY, class Word>
class NativePlan32 {
    cntt_native_plan* h_ = nullptr;
    explicit NativePlan32(cntt_native_plan* h) : h_(h) {}

  public:
    NativePlan32(NativePlan32&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    NativePlan32(const NativePlan32&) = delete;
    ~NativePlan32() { cntt_native_plan_free(h_); }
    static std::optional<NativePlan32> try_new(size_t n, int device = 0)
    {
        cntt_native_plan* h = nullptr;
        int st = cntt_native_plan_new(n, BITS, BINARY ? 1 : 0, device, &h);
        if (is_none(st)) return std::nullopt;
        check(st);
        return NativePlan32(h);
    }
    // EXTENSION (no reference counterpart): primes k*2^17+1, n up to 65536 -- see cntt_native_plan_new_ext
    static std::optional<NativePlan32> try_new_extended(size_t n, int device = 0)
    {
        cntt_native_plan* h = nullptr;
        int st = cntt_native_plan_new_ext(n, BITS, BINARY ? 1 : 0, device, &h);
        if (is_none(st) || st == CNTT_UNSUPPORTED) return std::nullopt;
        check(st);
        return NativePlan32(h);
    }
    // Plan32::ntt_0() .. ntt_9(): an equal prime32::Plan (plans are pure functions of (n, p))
    std::optional<prime32::Plan> ntt_i(int i, int device = 0) const { return prime32::Plan::try_new(ntt_size(), cntt_native_prime(h_, i), device); }
    size_t ntt_size() const { return cntt_native_ntt_size(h_); }
    int num_primes() const { return cntt_native_num_primes(h_); }
    uint32_t ntt_modulus(int i) const { return cntt_native_prime(h_, i); }
    // reference shape: host slices of n words each
    void negacyclic_polymul(Word* prod, const Word* lhs, const Word* rhs, size_t len) const { check(cntt_native_polymul_host(h_, prod, lhs, rhs, len, 1)); }
    // batch extensions
    void negacyclic_polymul_batch(Word* prod, const Word* lhs, const Word* rhs, size_t batch) const { check(cntt_native_polymul_host(h_, prod, lhs, rhs, ntt_size() * batch, batch)); }
    void negacyclic_polymul_device(Word* prod, const Word* lhs, const Word* rhs, size_t batch, void* stream = nullptr) const { check(cntt_native_polymul(h_, prod, lhs, rhs, batch, stream)); }
    // reference shape (src/native64.rs:971-1038): host slices; mod_p holds num_primes() planes of n u32 each, back to back
    // (the reference's mod_p0, mod_p1, ... concatenated); inv clobbers mod_p like the reference
    void fwd(const Word* value, uint32_t* mod_p, size_t len) const { check(cntt_native_fwd_host(h_, value, mod_p, len, 1)); }
    void fwd_binary(const Word* value, uint32_t* mod_p, size_t len) const
    {
        static_assert(BINARY, "fwd_binary exists only on native_binary* plans");
        check(cntt_native_fwd_binary_host(h_, value, mod_p, len, 1));
    }
    void inv(Word* value, uint32_t* mod_p, size_t len) const { check(cntt_native_inv_host(h_, value, mod_p, len, 1)); }
    // fwd / fwd_binary / inv on device residue planes (plane k of polynomial b at mod_p[(k*batch + b)*n])
    void fwd_device(const Word* value, uint32_t* mod_p, size_t batch, void* stream = nullptr) const { check(cntt_native_fwd(h_, value, mod_p, batch, stream)); }
    void fwd_binary_device(const Word* value, uint32_t* mod_p, size_t batch, void* stream = nullptr) const
    {
        static_assert(BINARY, "fwd_binary exists only on native_binary* plans");
        check(cntt_native_fwd_binary(h_, value, mod_p, batch, stream));
    }
    void inv_device(Word* value, uint32_t* mod_p, size_t batch, void* stream = nullptr) const { check(cntt_native_inv(h_, value, mod_p, batch, stream)); }
};

// native32 / native64 / native_binary32 / native_binary64 ::Plan52 (reference: feature = "nightly" + AVX-512 IFMA only):
// u64 residue planes over the ~50-bit primes52; negacyclic_polymul returns exactly what Plan32 returns.
template <int BITS, bool BINARY, class Word>
class NativePlan52 {
    cntt_native52_plan* h_ = nullptr;
    explicit NativePlan52(cntt_native52_plan* h) : h_(h) {}

  public:
    NativePlan52(NativePlan52&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    NativePlan52(const NativePlan52&) = delete;
    ~NativePlan52() { cntt_native52_plan_free(h_); }
    static std::optional<NativePlan52> try_new(size_t n, int device = 0)
    {
        cntt_native52_plan* h = nullptr;
        int st = cntt_native52_plan_new(n, BITS, BINARY ? 1 : 0, device, &h);
        if (is_none(st)) return std::nullopt;
        check(st);
        return NativePlan52(h);
    }
    size_t ntt_size() const { return cntt_native52_ntt_size(h_); }
    int num_primes() const { return cntt_native52_num_primes(h_); }
    uint64_t ntt_modulus(int i) const { return cntt_native52_prime(h_, i); }
    void negacyclic_polymul(Word* prod, const Word* lhs, const Word* rhs, size_t len) const { check(cntt_native52_polymul_host(h_, prod, lhs, rhs, len, 1)); }
    void negacyclic_polymul_device(Word* prod, const Word* lhs, const Word* rhs, size_t batch, void* stream = nullptr) const { check(cntt_native52_polymul(h_, prod, lhs, rhs, batch, stream)); }
    void fwd_device(const Word* value, uint64_t* mod_p, size_t batch, void* stream = nullptr) const { check(cntt_native52_fwd(h_, value, mod_p, batch, stream)); }
    void fwd_binary_device(const Word* value, uint64_t* mod_p, size_t batch, void* stream = nullptr) const
    {
        static_assert(BINARY, "fwd_binary exists only on native_binary* plans");
        check(cntt_native52_fwd_binary(h_, value, mod_p, batch, stream));
    }
    void inv_device(Word* value, uint64_t* mod_p, size_t batch, void* stream = nullptr) const { check(cntt_native52_inv(h_, value, mod_p, batch, stream)); }
};

namespace native32 { using Plan32 = NativePlan32<32, false, uint32_t>; using Plan52 = NativePlan52<32, false, uint32_t>; }
namespace native64 { using Plan32 = NativePlan32<64, false, uint64_t>; using Plan52 = NativePlan52<64, false, uint64_t>; }
namespace native128 { using Plan32 = NativePlan32<128, false, unsigned __int128>; }
namespace native_binary32 { using Plan32 = NativePlan32<32, true, uint32_t>; using Plan52 = NativePlan52<32, true, uint32_t>; }
namespace native_binary64 { using Plan32 = NativePlan32<64, true, uint64_t>; using Plan52 = NativePlan52<64, true, uint64_t>; }
namespace native_binary128 { using Plan32 = NativePlan32<128, true, unsigned __int128>; }

// product::Plan (src/product.rs:139-967): modulus = product of distinct primes.  NTT-domain buffers hold
// ntt_domain_len() u64 words per polynomial in the reference's packed layout.
namespace product {
struct FwdMode {
    int kind;       // 0 = Generic, 1 = Bounded(bound)
    uint64_t bound;
    static FwdMode Generic() { return {0, 0}; }
    static FwdMode Bounded(uint64_t b) { return {1, b}; }
};
enum class InvMode : int { Replace = 0, Accumulate = 1 };

class Plan {
    cntt_product_plan* h_ = nullptr;
    explicit Plan(cntt_product_plan* h) : h_(h) {}

  public:
    Plan(Plan&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    Plan(const Plan&) = delete;
    ~Plan() { cntt_product_plan_free(h_); }
    /* Plan::try_new(polynomial_size, modulus, factors) -> Option<Plan> */
    static std::optional<Plan> try_new(size_t polynomial_size, uint64_t modulus, const uint64_t* factors, size_t nfactors, int device = 0)
    {
        cntt_product_plan* h = nullptr;
        int st = cntt_product_plan_new(polynomial_size, modulus, factors, nfactors, device, &h);
        if (is_none(st)) return std::nullopt;
        check(st);
        return Plan(h);
    }
    size_t ntt_size() const { return cntt_product_ntt_size(h_); }
    uint64_t modulus() const { return cntt_product_modulus(h_); }
    size_t ntt_domain_len() const { return cntt_product_ntt_domain_len(h_); }
    // reference shape: host slices, one polynomial
    void fwd(uint64_t* ntt, size_t ntt_len, const uint64_t* standard, size_t standard_len, FwdMode mode) const { check(cntt_product_fwd_host(h_, ntt, standard, ntt_len, standard_len, mode.kind, mode.bound, 1)); }
    void inv(uint64_t* standard, size_t standard_len, uint64_t* ntt, size_t ntt_len, InvMode mode) const { check(cntt_product_inv_host(h_, standard, ntt, standard_len, ntt_len, (int)mode, 1)); }
    void mul_assign_normalize(uint64_t* lhs, const uint64_t* rhs, size_t len) const { check(cntt_product_mul_assign_normalize_host(h_, lhs, rhs, len, 1)); }
    void normalize(uint64_t* values, size_t len) const { check(cntt_product_normalize_host(h_, values, len, 1)); }
    void mul_accumulate(uint64_t* acc, const uint64_t* lhs, const uint64_t* rhs, size_t len) const { check(cntt_product_mul_accumulate_host(h_, acc, lhs, rhs, len, 1)); }
    // batch / device extensions
    void fwd_device(uint64_t* ntt, const uint64_t* standard, FwdMode mode, size_t batch, void* stream = nullptr) const { check(cntt_product_fwd(h_, ntt, standard, mode.kind, mode.bound, batch, stream)); }
    void inv_device(uint64_t* standard, uint64_t* ntt, InvMode mode, size_t batch, void* stream = nullptr) const { check(cntt_product_inv(h_, standard, ntt, (int)mode, batch, stream)); }
    void mul_assign_normalize_device(uint64_t* lhs, const uint64_t* rhs, size_t batch, void* stream = nullptr) const { check(cntt_product_mul_assign_normalize(h_, lhs, rhs, batch, stream)); }
    void normalize_device(uint64_t* values, size_t batch, void* stream = nullptr) const { check(cntt_product_normalize(h_, values, batch, stream)); }
    void mul_accumulate_device(uint64_t* acc, const uint64_t* lhs, const uint64_t* rhs, size_t batch, void* stream = nullptr) const { check(cntt_product_mul_accumulate(h_, acc, lhs, rhs, batch, stream)); }
};
} // namespace product

} // namespace concrete_ntt
