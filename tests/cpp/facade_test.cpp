// facade_test.cpp -- drives the C++ mirror (include/concrete_ntt.hpp) the way the reference's README and
// examples drive the crate: README.md:30-51, examples/mul_poly_prime.rs, examples/mul_poly_native.rs.
// Exit code 0 = all checks passed.  Needs a CUDA device (run by tests/test_cpp_facade.py under -m gpu).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "concrete_ntt.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

// schoolbook negacyclic product mod p (p == 0: wrapping), the reference tests' oracle (src/prime32.rs:966-978)
template <class T>
static std::vector<T> schoolbook(const std::vector<T>& a, const std::vector<T>& b, unsigned long long p)
{
    size_t n = a.size();
    std::vector<T> out(n, 0);
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            size_t k = (i + j) % n;
            bool neg = i + j >= n;
            if (p == 0) {
                T pr = (T)(a[i] * b[j]);
                out[k] = neg ? (T)(out[k] - pr) : (T)(out[k] + pr);
            } else {
                unsigned long long pr = (unsigned long long)(((unsigned __int128)a[i] * b[j]) % p);
                unsigned long long cur = out[k];
                out[k] = (T)(neg ? (cur + p - pr) % p : (cur + pr) % p);
            }
        }
    return out;
}

int main()
{
    using namespace concrete_ntt;
    // README example
    {
        auto plan = prime32::Plan::try_new(32, 1062862849u);
        REQUIRE(plan.has_value());
        std::vector<uint32_t> data(32), buf;
        for (uint32_t i = 0; i < 32; i++) data[i] = i;
        buf = data;
        plan->fwd(buf.data(), buf.size());
        plan->inv(buf.data(), buf.size());
        for (uint32_t i = 0; i < 32; i++) REQUIRE(buf[i] == data[i] * 32);
    }
    // try_new rejections and the panic
    REQUIRE(!prime32::Plan::try_new(16, 1062862849u).has_value());
    REQUIRE(!prime64::Plan::try_new(2048, 1024).has_value());
    REQUIRE(!native64::Plan32::try_new(65536).has_value());
    try { (void)prime32::Plan::try_new(32, 1u); REQUIRE(false); } catch (const Panic&) {}
    std::mt19937_64 rng(7);
    // examples/mul_poly_prime.rs: p = 1073479681, N = 1024
    {
        const uint32_t p = 1073479681u;
        const size_t n = 1024;
        auto plan = prime32::Plan::try_new(n, p).value();
        std::vector<uint32_t> a(n), b(n);
        for (auto& v : a) v = (uint32_t)(rng() % p);
        for (auto& v : b) v = (uint32_t)(rng() % p);
        auto expect = schoolbook(a, b, p);
        auto fa = a, fb = b;
        plan.fwd(fa.data(), n);
        plan.fwd(fb.data(), n);
        plan.mul_assign_normalize(fa.data(), fb.data(), n);
        plan.inv(fa.data(), n);
        REQUIRE(fa == expect);
        try { plan.fwd(fa.data(), n - 1); REQUIRE(false); } catch (const Panic&) {}
    }
    // examples/mul_poly_native.rs: native32 N = 1024; plus native64 and native128 at N = 64
    {
        const size_t n = 1024;
        auto plan = native32::Plan32::try_new(n).value();
        std::vector<uint32_t> a(n), b(n), prod(n);
        for (auto& v : a) v = (uint32_t)rng();
        for (auto& v : b) v = (uint32_t)rng();
        plan.negacyclic_polymul(prod.data(), a.data(), b.data(), n);
        REQUIRE(prod == schoolbook(a, b, 0));
    }
    {
        const size_t n = 64;
        auto plan = native64::Plan32::try_new(n).value();
        std::vector<uint64_t> a(n), b(n), prod(n);
        for (auto& v : a) v = rng();
        for (auto& v : b) v = rng();
        plan.negacyclic_polymul(prod.data(), a.data(), b.data(), n);
        REQUIRE(prod == schoolbook(a, b, 0));
        auto bplan = native_binary64::Plan32::try_new(n).value();
        for (auto& v : b) v &= 1;
        bplan.negacyclic_polymul(prod.data(), a.data(), b.data(), n);
        REQUIRE(prod == schoolbook(a, b, 0));
    }
    {
        const size_t n = 64;
        auto plan = native128::Plan32::try_new(n).value();
        std::vector<unsigned __int128> a(n), b(n), prod(n);
        for (auto& v : a) v = ((unsigned __int128)rng() << 64) | rng();
        for (auto& v : b) v = ((unsigned __int128)rng() << 64) | rng();
        plan.negacyclic_polymul(prod.data(), a.data(), b.data(), n);
        REQUIRE(prod == schoolbook(a, b, 0));
    }
    // Solinas round trip, batch of 3
    {
        const size_t n = 2048, batch = 3;
        auto plan = prime64::Plan::try_new(n, prime64::Solinas::P).value();
        std::vector<uint64_t> a(n * batch), buf;
        for (auto& v : a) v = rng() % prime64::Solinas::P;
        buf = a;
        plan.fwd_batch(buf.data(), batch);
        plan.inv_batch(buf.data(), batch);
        plan.normalize(buf.data(), buf.size());
        REQUIRE(buf == a);
    }
    // product::Plan, the tfhe-rs NTT-PBS shape (src/product.rs:444-445): two primes < 2^31; polymul mod p0*p1 through
    // fwd (Generic and Bounded) / mul_assign_normalize / inv (Replace, then Accumulate), host slices
    {
        const size_t n = 256;
        const uint64_t p0 = prime::largest_prime_in_arithmetic_progression64(2 * n, 1, 0, 1ull << 31).value();
        const uint64_t p1 = prime::largest_prime_in_arithmetic_progression64(2 * n, 1, 0, p0 - 1).value();
        const uint64_t p = p0 * p1, factors[2] = {p0, p1};
        auto plan = product::Plan::try_new(n, p, factors, 2).value();
        REQUIRE(plan.ntt_size() == n && plan.modulus() == p && plan.ntt_domain_len() == n);
        const uint64_t dup[3] = {p1, p0, p1};
        REQUIRE(!product::Plan::try_new(n, p0 * p1 * p1, dup, 3).has_value());   // src/product.rs:1162-1169
        std::vector<uint64_t> a(n), b(n);
        for (auto& v : a) v = rng() % 1000;          // small: also a valid input of FwdMode::Bounded(1000)
        for (auto& v : b) v = rng() % p;
        auto expect = schoolbook(a, b, p);
        const size_t dl = plan.ntt_domain_len();
        std::vector<uint64_t> fa(dl), fa2(dl), fb(dl), out(n), acc(n);
        plan.fwd(fa.data(), dl, a.data(), n, product::FwdMode::Generic());
        plan.fwd(fa2.data(), dl, a.data(), n, product::FwdMode::Bounded(1000));
        REQUIRE(fa == fa2);
        plan.fwd(fb.data(), dl, b.data(), n, product::FwdMode::Generic());
        plan.mul_assign_normalize(fa.data(), fb.data(), dl);
        fa2 = fa;
        plan.inv(out.data(), n, fa.data(), dl, product::InvMode::Replace);
        REQUIRE(out == expect);
        for (auto& v : acc) v = rng() % p;
        auto acc0 = acc;
        plan.inv(acc.data(), n, fa2.data(), dl, product::InvMode::Accumulate);
        for (size_t i = 0; i < n; i++) REQUIRE(acc[i] == (uint64_t)(((unsigned __int128)acc0[i] + expect[i]) % p));
        try { plan.normalize(fa.data(), dl - 1); REQUIRE(false); } catch (const Panic&) {}
    }
    std::printf("facade ok\n");
    return 0;
}
