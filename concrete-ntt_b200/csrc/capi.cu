// capi.cu -- the C ABI (include/cntt_b200.h): plan construction, argument checking, launches and the
// host-slice staging paths.  No arithmetic happens on the CPU here except plan-time table building.
#include "../../include/cntt_b200.h"
#include "dispatch.hpp"
#include "host_math.hpp"
#include "native.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

using namespace cntt;

#define CNTT_API extern "C" __attribute__((visibility("default")))

// ---- errors ----------------------------------------------------------------------------------------
static thread_local std::string t_cuda_err;

static int cuda_fail(cudaError_t e, const char* where)
{
    t_cuda_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    (void)cudaGetLastError(); // clear the sticky-less error state
    return CNTT_CUDA_ERROR;
}
#define CU(expr)                                                  \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);       \
    } while (0)

CNTT_API const char* cntt_status_string(int s)
{
    switch (s) {
    case CNTT_OK: return "ok";
    case CNTT_INVALID_SIZE: return "invalid polynomial size (try_new -> None)";
    case CNTT_INVALID_MODULUS: return "modulus is not prime (try_new -> None)";
    case CNTT_NO_ROOT: return "no primitive 2n-th root of unity (try_new -> None)";
    case CNTT_LENGTH_MISMATCH: return "buffer length does not match the plan (reference: assert_eq! panic)";
    case CNTT_CUDA_ERROR: return "CUDA error (see cntt_last_cuda_error)";
    case CNTT_NULL_POINTER: return "null pointer";
    case CNTT_UNSUPPORTED: return "operation not supported by this plan";
    case CNTT_PANIC_MODULUS: return "modulus <= 1 (reference: Div::new assert panic)";
    case CNTT_MISALIGNED: return "device buffer is not 16-byte aligned";
    default: return "unknown status";
    }
}
CNTT_API const char* cntt_last_cuda_error(void) { return t_cuda_err.c_str(); }
CNTT_API const char* cntt_version(void) { return "cntt_b200 0.2 (sm_100a; concrete-ntt 0.2.0 semantics)"; }

// device batches are accessed with 128-bit (and, when they allow it, 256-bit) loads and stores
static inline bool misaligned16(const void* a, const void* b = nullptr, const void* c = nullptr)
{
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15u) != 0;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        cur = dev;
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != cur) cudaSetDevice(prev);
    }
    int cur = -1;
};
#define GUARD(dev)                                                         \
    DeviceGuard _guard(dev);                                               \
    if (!_guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice")

// ---- plan-time helpers -------------------------------------------------------------------------------
CNTT_API int cntt_is_prime64(uint64_t n) { return host::is_prime_u64(n) ? 1 : 0; }
CNTT_API int cntt_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset, uint64_t lo, uint64_t hi, uint64_t* out)
{
    uint64_t v;
    if (!host::largest_prime_in_arithmetic_progression(factor, offset, lo, hi, &v)) return 0;
    if (out) *out = v;
    return 1;
}
CNTT_API int cntt_find_primitive_root64(uint64_t p, uint64_t degree, uint64_t* root)
{
    if (p < 2 || degree < 2 || (degree & (degree - 1))) return 0;
    uint64_t r;
    if (!host::primitive_root_pow2(p, degree, &r)) return 0;
    if (root) *root = r;
    return 1;
}
CNTT_API int cntt_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return CNTT_NULL_POINTER;
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return CNTT_OK;
}
CNTT_API int cntt_host_free(void* ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return CNTT_OK;
}

// ---- staging for the host-slice entry points ----------------------------------------------------------
// A plan owns (lazily) one internal stream and a device staging arena; host calls on one plan are
// serialised by a mutex (the reference call is synchronous too).
struct Staging {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr};   // ring of host_inplace: upload of slot s done
    cudaEvent_t ev_down[4] = {nullptr, nullptr, nullptr, nullptr}; // download of slot s done
    void* buf = nullptr;
    size_t cap = 0;
    bool ready = false; // set only after every stream and event exists
    cudaError_t create_objects()
    {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (auto* arr : {ev, ev_up, ev_down})
            for (int i = 0; i < 4; i++)
                if ((e = cudaEventCreateWithFlags(&arr[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void destroy_objects()
    {
        for (auto* arr : {ev, ev_up, ev_down})
            for (int i = 0; i < 4; i++) {
                if (arr[i]) cudaEventDestroy(arr[i]);
                arr[i] = nullptr;
            }
        if (stream) cudaStreamDestroy(stream);
        if (stream2) cudaStreamDestroy(stream2);
        stream = stream2 = nullptr;
    }
    cudaError_t ensure(size_t bytes)
    {
        cudaError_t e;
        if (!ready) {
            if ((e = create_objects()) != cudaSuccess) {
                destroy_objects(); // failure-atomic: the next call starts from scratch instead of using half a set
                return e;
            }
            ready = true;
        }
        if (bytes > cap) {
            if (buf) cudaFree(buf);
            buf = nullptr;
            cap = 0;
            if ((e = cudaMalloc(&buf, bytes)) != cudaSuccess) return e;
            cap = bytes;
        }
        return cudaSuccess;
    }
    void release()
    {
        if (buf) cudaFree(buf);
        buf = nullptr;
        cap = 0;
        destroy_objects();
        ready = false;
    }
};

// ---- prime plans ------------------------------------------------------------------------------------------
enum Cls32 { C32_L4 = 0, C32_L2 = 1, C32_G = 2 };
enum Cls64 { C64_L4 = 0, C64_L2 = 1, C64_S = 2, C64_G = 3 };

struct cntt_prime32_plan {
    size_t n;
    uint32_t p;
    int device;
    int cls;
    int logn;
    Mod32 mod;
    uint2* d_fwd;
    uint2* d_inv;
    uint2* d_fwd_last; // last-pass layouts of the CTA kernel (ntt_engine.cuh, TwSrc); nullptr when unused
    uint2* d_inv_last;
    TwHead<uint2> head_fwd, head_inv; // host copies of the first heap entries (by-value kernel parameter)
    Staging stg;
};
struct cntt_prime64_plan {
    size_t n;
    uint64_t p;
    int device;
    int cls;
    int logn;
    Mod64 mod;
    void* d_fwd; // ulonglong2[n] (Shoup classes) or uint64_t[n] (Solinas)
    void* d_inv;
    void* d_fwd_last; // as in cntt_prime32_plan
    void* d_inv_last;
    alignas(16) unsigned char head_fwd[sizeof(TwHead<ulonglong2>)]; // TwHead<Tw> of the plan's class (same bytes for all)
    alignas(16) unsigned char head_inv[sizeof(TwHead<ulonglong2>)];
    Staging stg;
};

static int validate(size_t n, uint64_t p, size_t min_n, uint64_t* psi, int* logn)
{
    if (p <= 1) return CNTT_PANIC_MODULUS; // Div32::new / Div64::new assert (fastdiv.rs:49,99)
    if (n < min_n || (n & (n - 1)) != 0) return CNTT_INVALID_SIZE;
    int lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    if (!host::is_prime_u64(p)) return CNTT_INVALID_MODULUS;
    if (lg >= 63 || !host::primitive_root_pow2(p, 2 * (uint64_t)n, psi)) return CNTT_NO_ROOT;
    // an implementation limit (table memory), not one of the reference's None cases: reported as such, after them
    if (lg > kMaxLogN) return CNTT_UNSUPPORTED;
    *logn = lg;
    return CNTT_OK;
}

static void free_tables32(cntt_prime32_plan* pl)
{
    if (pl->d_fwd) cudaFree(pl->d_fwd);
    if (pl->d_inv) cudaFree(pl->d_inv);
    if (pl->d_fwd_last) cudaFree(pl->d_fwd_last);
    if (pl->d_inv_last) cudaFree(pl->d_inv_last);
}
// device-side re-layout of the heap tables for the last pass of the CTA kernel (synchronous: plan time)
static cudaError_t build_last32(cntt_prime32_plan* pl)
{
    const bool uses = pl->cls == C32_L4 ? uses_last_A32L4(pl->logn) : pl->cls == C32_L2 ? uses_last_A32L2(pl->logn) : uses_last_A32G(pl->logn);
    if (!uses) return cudaSuccess;
    cudaError_t e;
    const size_t entries = last_table_entries<A32L4>(pl->logn); // the same for every 32-bit class: two layouts at N = 1024 (ntt_kernels.cuh)
    if ((e = cudaMalloc(&pl->d_fwd_last, entries * sizeof(uint2))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&pl->d_inv_last, entries * sizeof(uint2))) != cudaSuccess) return e;
    for (int dir = 0; dir < 2; dir++) {
        const uint2* heap = dir ? pl->d_inv : pl->d_fwd;
        uint2* o = dir ? pl->d_inv_last : pl->d_fwd_last;
        switch (pl->cls) {
        case C32_L4: e = build_last_A32L4(pl->logn, dir == 0, heap, o, nullptr); break;
        case C32_L2: e = build_last_A32L2(pl->logn, dir == 0, heap, o, nullptr); break;
        default: e = build_last_A32G(pl->logn, dir == 0, heap, o, nullptr); break;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaStreamSynchronize(nullptr);
}

static int build_prime32(size_t n, uint32_t p, int device, cntt_prime32_plan** out)
{
    uint64_t psi;
    int logn;
    int st = validate(n, p, 32, &psi, &logn);
    if (st != CNTT_OK) return st;
    std::vector<uint64_t> f, iv;
    host::negacyclic_twiddles(p, logn, psi, f, iv);
    std::vector<uint2> hf(n), hi(n);
    for (size_t i = 0; i < n; i++) {
        hf[i] = make_uint2((uint32_t)f[i], (uint32_t)((f[i] << 32) / p));
        hi[i] = make_uint2((uint32_t)iv[i], (uint32_t)((iv[i] << 32) / p));
    }
    GUARD(device);
    cntt_prime32_plan* pl = new cntt_prime32_plan();
    pl->n = n; pl->p = p; pl->device = device; pl->logn = logn;
    pl->cls = p < (1u << 30) ? C32_L4 : p < (1u << 31) ? C32_L2 : C32_G;
    host::Fp fp(p);
    Mod32& m = pl->mod;
    m.p = p; m.two_p = 2u * p; m.neg_p = 0u - p;
    m.n_inv = (uint32_t)fp.inv(n % p);
    m.n_inv_shoup = (uint32_t)(((uint64_t)m.n_inv << 32) / p);
    const int big_q = host::ilog2(p) + 1;                       // prime32.rs:669
    m.big_q_m1 = (uint32_t)(big_q - 1);
    m.p_barrett = (uint32_t)((((uint64_t)1) << (big_q + 31)) / p); // prime32.rs:670-671
    pl->d_fwd = pl->d_inv = pl->d_fwd_last = pl->d_inv_last = nullptr;
    std::memset(&pl->head_fwd, 0, sizeof(pl->head_fwd));
    std::memset(&pl->head_inv, 0, sizeof(pl->head_inv));
    for (size_t i = 0; i < n && i < sizeof(pl->head_fwd.e) / sizeof(uint2); i++) { pl->head_fwd.e[i] = hf[i]; pl->head_inv.e[i] = hi[i]; }
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_fwd, n * sizeof(uint2))) != cudaSuccess || (e = cudaMalloc(&pl->d_inv, n * sizeof(uint2))) != cudaSuccess ||
        (e = cudaMemcpy(pl->d_fwd, hf.data(), n * sizeof(uint2), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(pl->d_inv, hi.data(), n * sizeof(uint2), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = build_last32(pl)) != cudaSuccess) {
        free_tables32(pl);
        delete pl;
        return cuda_fail(e, "prime32 plan upload");
    }
    *out = pl;
    return CNTT_OK;
}

CNTT_API int cntt_prime32_plan_new(size_t n, uint32_t p, int device, cntt_prime32_plan** out)
{
    if (!out) return CNTT_NULL_POINTER;
    *out = nullptr;
    return build_prime32(n, p, device, out);
}
CNTT_API void cntt_prime32_plan_free(cntt_prime32_plan* pl)
{
    if (!pl) return;
    DeviceGuard g(pl->device);
    pl->stg.release();
    free_tables32(pl);
    delete pl;
}
CNTT_API size_t cntt_prime32_ntt_size(const cntt_prime32_plan* pl) { return pl ? pl->n : 0; }
CNTT_API uint32_t cntt_prime32_modulus(const cntt_prime32_plan* pl) { return pl ? pl->p : 0; }

template <class A> static PlanDev<A> dev32(const cntt_prime32_plan* pl)
{
    PlanDev<A> d;
    d.logn = pl->logn; d.mod = pl->mod; d.tw_fwd = pl->d_fwd; d.tw_inv = pl->d_inv;
    d.tw_fwd_last = pl->d_fwd_last; d.tw_inv_last = pl->d_inv_last;
    d.head_fwd = &pl->head_fwd; d.head_inv = &pl->head_inv;
    return d;
}
static cudaError_t run_ntt32(const cntt_prime32_plan* pl, uint32_t* d, size_t batch, bool fwd, cudaStream_t st)
{
    switch (pl->cls) {
    case C32_L4: return ntt_A32L4(dev32<A32L4>(pl), d, batch, fwd, st);
    case C32_L2: return ntt_A32L2(dev32<A32L2>(pl), d, batch, fwd, st);
    default: return ntt_A32G(dev32<A32G>(pl), d, batch, fwd, st);
    }
}
static cudaError_t run_pw32(const cntt_prime32_plan* pl, int op, uint32_t* dst, const uint32_t* a, const uint32_t* b, size_t nwords, cudaStream_t st)
{
    switch (pl->cls) {
    case C32_L4: return pointwise_A32L4(dev32<A32L4>(pl), op, dst, a, b, nwords, st);
    case C32_L2: return pointwise_A32L2(dev32<A32L2>(pl), op, dst, a, b, nwords, st);
    default: return pointwise_A32G(dev32<A32G>(pl), op, dst, a, b, nwords, st);
    }
}

CNTT_API int cntt_prime32_fwd(const cntt_prime32_plan* pl, uint32_t* d_buf, size_t batch, void* stream)
{
    if (!pl || (!d_buf && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_buf)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(run_ntt32(pl, d_buf, batch, true, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime32_inv(const cntt_prime32_plan* pl, uint32_t* d_buf, size_t batch, void* stream)
{
    if (!pl || (!d_buf && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_buf)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(run_ntt32(pl, d_buf, batch, false, (cudaStream_t)stream));
    return CNTT_OK;
}
// The reference zips the slices and its SIMD bodies only touch whole vectors (prime32.rs:348); n is a
// multiple of 32, so whole-polynomial streams are always whole vectors.  A ragged tail (nwords not a
// multiple of 4) is rejected instead of being silently dropped.
CNTT_API int cntt_prime32_mul_assign_normalize(const cntt_prime32_plan* pl, uint32_t* l, const uint32_t* r, size_t nwords, void* stream)
{
    if (!pl || ((!l || !r) && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(l, r)) return CNTT_MISALIGNED;
    if (nwords % 4) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw32(pl, OP_MUL_ASSIGN_NORMALIZE, l, r, nullptr, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime32_normalize(const cntt_prime32_plan* pl, uint32_t* v, size_t nwords, void* stream)
{
    if (!pl || (!v && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(v)) return CNTT_MISALIGNED;
    if (nwords % 4) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw32(pl, OP_NORMALIZE, v, nullptr, nullptr, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime32_mul_accumulate(const cntt_prime32_plan* pl, uint32_t* acc, const uint32_t* l, const uint32_t* r, size_t nwords, void* stream)
{
    if (!pl || ((!acc || !l || !r) && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(acc, l, r)) return CNTT_MISALIGNED;
    if (nwords % 4) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw32(pl, OP_MUL_ACCUMULATE, acc, l, r, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}

// ---- prime64 ------------------------------------------------------------------------------------------------
static void free_tables64(cntt_prime64_plan* pl)
{
    if (pl->d_fwd) cudaFree(pl->d_fwd);
    if (pl->d_inv) cudaFree(pl->d_inv);
    if (pl->d_fwd_last) cudaFree(pl->d_fwd_last);
    if (pl->d_inv_last) cudaFree(pl->d_inv_last);
}
static cudaError_t build_last64(cntt_prime64_plan* pl, size_t bytes)
{
    bool uses;
    switch (pl->cls) {
    case C64_L4: uses = uses_last_A64L4(pl->logn); break;
    case C64_L2: uses = uses_last_A64L2(pl->logn); break;
    case C64_S: uses = uses_last_A64S(pl->logn); break;
    default: uses = uses_last_A64G(pl->logn); break;
    }
    if (!uses) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_fwd_last, bytes)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&pl->d_inv_last, bytes)) != cudaSuccess) return e;
    for (int dir = 0; dir < 2; dir++) {
        const void* heap = dir ? pl->d_inv : pl->d_fwd;
        void* o = dir ? pl->d_inv_last : pl->d_fwd_last;
        switch (pl->cls) {
        case C64_L4: e = build_last_A64L4(pl->logn, dir == 0, (const ulonglong2*)heap, (ulonglong2*)o, nullptr); break;
        case C64_L2: e = build_last_A64L2(pl->logn, dir == 0, (const ulonglong2*)heap, (ulonglong2*)o, nullptr); break;
        case C64_S: e = build_last_A64S(pl->logn, dir == 0, (const uint64_t*)heap, (uint64_t*)o, nullptr); break;
        default: e = build_last_A64G(pl->logn, dir == 0, (const ulonglong2*)heap, (ulonglong2*)o, nullptr); break;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaStreamSynchronize(nullptr);
}

static int build_prime64(size_t n, uint64_t p, int device, cntt_prime64_plan** out)
{
    typedef unsigned __int128 u128;
    uint64_t psi;
    int logn;
    int st = validate(n, p, 16, &psi, &logn);
    if (st != CNTT_OK) return st;
    std::vector<uint64_t> f, iv;
    host::negacyclic_twiddles(p, logn, psi, f, iv);
    const int cls = p == CNTT_SOLINAS_P ? C64_S : p < (1ull << 62) ? C64_L4 : p < (1ull << 63) ? C64_L2 : C64_G;
    GUARD(device);
    cntt_prime64_plan* pl = new cntt_prime64_plan();
    pl->n = n; pl->p = p; pl->device = device; pl->logn = logn; pl->cls = cls;
    host::Fp fp(p);
    Mod64& m = pl->mod;
    m.p = p; m.two_p = 2ull * p; m.neg_p = 0ull - p;
    m.n_inv = fp.inv(n % p);
    m.n_inv_shoup = (uint64_t)((((u128)m.n_inv) << 64) / p);
    const int big_q = host::ilog2(p) + 1;                           // prime64.rs:754
    m.big_q_m1 = (uint32_t)(big_q - 1);
    m.p_barrett = (uint64_t)((((u128)1) << (big_q + 63)) / p);      // prime64.rs:755-756 (unused when p >= 2^63)
    m.shift_head = 0;
    if (cls == C64_S) { // the shift butterflies of the leading levels assume tw[h] = 2^shift_exp(h): check it against the real table
        bool ok = true;
        for (size_t h = 1; h < (size_t)kShiftNodes && h < n; h++) {
            uint64_t want = 1, wanti = 1;
            for (int k = 0; k < shift_exp((int)h); k++) want = fp.mul(want, 2);
            for (int k = 0; k < (192 - shift_exp((int)h)) % 192; k++) wanti = fp.mul(wanti, 2);
            ok = ok && f[h] == want && iv[h] == wanti;
        }
        m.shift_head = ok ? 1u : 0u;
    }
    pl->d_fwd = pl->d_inv = pl->d_fwd_last = pl->d_inv_last = nullptr;
    cudaError_t e;
    size_t bytes;
    std::vector<ulonglong2> sf, si;
    const void *hf, *hi;
    if (cls == C64_S) {
        bytes = n * sizeof(uint64_t);
        hf = f.data(); hi = iv.data();
    } else {
        sf.resize(n); si.resize(n);
        for (size_t i = 0; i < n; i++) {
            sf[i] = make_ulonglong2(f[i], (uint64_t)((((u128)f[i]) << 64) / p));
            si[i] = make_ulonglong2(iv[i], (uint64_t)((((u128)iv[i]) << 64) / p));
        }
        bytes = n * sizeof(ulonglong2);
        hf = sf.data(); hi = si.data();
    }
    {
        const size_t hb = bytes < sizeof(pl->head_fwd) ? bytes : sizeof(pl->head_fwd);
        std::memset(pl->head_fwd, 0, sizeof(pl->head_fwd));
        std::memset(pl->head_inv, 0, sizeof(pl->head_inv));
        std::memcpy(pl->head_fwd, hf, hb);
        std::memcpy(pl->head_inv, hi, hb);
    }
    if ((e = cudaMalloc(&pl->d_fwd, bytes)) != cudaSuccess || (e = cudaMalloc(&pl->d_inv, bytes)) != cudaSuccess ||
        (e = cudaMemcpy(pl->d_fwd, hf, bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(pl->d_inv, hi, bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = build_last64(pl, bytes)) != cudaSuccess) {
        free_tables64(pl);
        delete pl;
        return cuda_fail(e, "prime64 plan upload");
    }
    *out = pl;
    return CNTT_OK;
}
CNTT_API int cntt_prime64_plan_new(size_t n, uint64_t p, int device, cntt_prime64_plan** out)
{
    if (!out) return CNTT_NULL_POINTER;
    *out = nullptr;
    return build_prime64(n, p, device, out);
}
CNTT_API void cntt_prime64_plan_free(cntt_prime64_plan* pl)
{
    if (!pl) return;
    DeviceGuard g(pl->device);
    pl->stg.release();
    free_tables64(pl);
    delete pl;
}
CNTT_API size_t cntt_prime64_ntt_size(const cntt_prime64_plan* pl) { return pl ? pl->n : 0; }
CNTT_API uint64_t cntt_prime64_modulus(const cntt_prime64_plan* pl) { return pl ? pl->p : 0; }

template <class A> static PlanDev<A> dev64(const cntt_prime64_plan* pl)
{
    PlanDev<A> d;
    d.logn = pl->logn; d.mod = pl->mod;
    d.tw_fwd = reinterpret_cast<const typename A::Tw*>(pl->d_fwd);
    d.tw_inv = reinterpret_cast<const typename A::Tw*>(pl->d_inv);
    d.tw_fwd_last = reinterpret_cast<const typename A::Tw*>(pl->d_fwd_last);
    d.tw_inv_last = reinterpret_cast<const typename A::Tw*>(pl->d_inv_last);
    d.head_fwd = reinterpret_cast<const TwHead<typename A::Tw>*>(pl->head_fwd);
    d.head_inv = reinterpret_cast<const TwHead<typename A::Tw>*>(pl->head_inv);
    return d;
}
static cudaError_t run_ntt64(const cntt_prime64_plan* pl, uint64_t* d, size_t batch, bool fwd, cudaStream_t st)
{
    switch (pl->cls) {
    case C64_L4: return ntt_A64L4(dev64<A64L4>(pl), d, batch, fwd, st);
    case C64_L2: return ntt_A64L2(dev64<A64L2>(pl), d, batch, fwd, st);
    case C64_S: return ntt_A64S(dev64<A64S>(pl), d, batch, fwd, st);
    default: return ntt_A64G(dev64<A64G>(pl), d, batch, fwd, st);
    }
}
static cudaError_t run_pw64(const cntt_prime64_plan* pl, int op, uint64_t* dst, const uint64_t* a, const uint64_t* b, size_t nwords, cudaStream_t st)
{
    switch (pl->cls) {
    case C64_L4: return pointwise_A64L4(dev64<A64L4>(pl), op, dst, a, b, nwords, st);
    case C64_L2: return pointwise_A64L2(dev64<A64L2>(pl), op, dst, a, b, nwords, st);
    case C64_S: return pointwise_A64S(dev64<A64S>(pl), op, dst, a, b, nwords, st);
    default: return pointwise_A64G(dev64<A64G>(pl), op, dst, a, b, nwords, st);
    }
}
CNTT_API int cntt_prime64_fwd(const cntt_prime64_plan* pl, uint64_t* d_buf, size_t batch, void* stream)
{
    if (!pl || (!d_buf && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_buf)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(run_ntt64(pl, d_buf, batch, true, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime64_inv(const cntt_prime64_plan* pl, uint64_t* d_buf, size_t batch, void* stream)
{
    if (!pl || (!d_buf && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_buf)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(run_ntt64(pl, d_buf, batch, false, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime64_mul_assign_normalize(const cntt_prime64_plan* pl, uint64_t* l, const uint64_t* r, size_t nwords, void* stream)
{
    if (!pl || ((!l || !r) && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(l, r)) return CNTT_MISALIGNED;
    if (nwords % 2) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw64(pl, OP_MUL_ASSIGN_NORMALIZE, l, r, nullptr, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime64_normalize(const cntt_prime64_plan* pl, uint64_t* v, size_t nwords, void* stream)
{
    if (!pl || (!v && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(v)) return CNTT_MISALIGNED;
    if (nwords % 2) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw64(pl, OP_NORMALIZE, v, nullptr, nullptr, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_prime64_mul_accumulate(const cntt_prime64_plan* pl, uint64_t* acc, const uint64_t* l, const uint64_t* r, size_t nwords, void* stream)
{
    if (!pl || ((!acc || !l || !r) && nwords)) return CNTT_NULL_POINTER;
    if (misaligned16(acc, l, r)) return CNTT_MISALIGNED;
    if (nwords % 2) return CNTT_LENGTH_MISMATCH;
    GUARD(pl->device);
    CU(run_pw64(pl, OP_MUL_ACCUMULATE, acc, l, r, nwords, (cudaStream_t)stream));
    return CNTT_OK;
}

// ---- host-slice flavours of the prime plans --------------------------------------------------------------------
// Chunked over a ring of kHostSlots device slots: uploads run on `stream2`, compute + downloads on `stream`, so the
// two PCIe directions stay busy at the same time.  A slot is reused once the download of the chunk that last held
// it has finished; with four slots the upload engine never waits on that (two slots left a bubble of one kernel
// time per chunk on the H2D side).
constexpr int kHostSlots = 4;
template <class Fn>
static int host_inplace(Staging& stg, int device, void* h_buf, size_t total_bytes, size_t bytes_per_poly, size_t batch, Fn launch)
{
    if (total_bytes == 0) return CNTT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    std::lock_guard<std::mutex> lk(stg.mu);
    const size_t target = (size_t)32 << 20;
    size_t chunk = std::max<size_t>(1, std::min(batch, target / std::max<size_t>(1, bytes_per_poly)));
    if (chunk >= batch) chunk = batch;
    const size_t slot = chunk * bytes_per_poly;
    const size_t nchunks = (batch + chunk - 1) / chunk;
    const size_t nslots = std::min<size_t>(kHostSlots, nchunks);
    CU(stg.ensure(nslots * slot));
    char* h = static_cast<char*>(h_buf);
    for (size_t c = 0; c < nchunks; c++) {
        const size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
        const int s = (int)(c % kHostSlots);
        char* d = static_cast<char*>(stg.buf) + (size_t)s * slot;
        if (c >= (size_t)kHostSlots) CU(cudaStreamWaitEvent(stg.stream2, stg.ev_down[s], 0)); // slot free: chunk c - kHostSlots is home
        CU(cudaMemcpyAsync(d, h + b0 * bytes_per_poly, nb * bytes_per_poly, cudaMemcpyHostToDevice, stg.stream2));
        CU(cudaEventRecord(stg.ev_up[s], stg.stream2));
        CU(cudaStreamWaitEvent(stg.stream, stg.ev_up[s], 0));
        cudaError_t e = launch(d, nb, stg.stream);
        if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
        CU(cudaMemcpyAsync(h + b0 * bytes_per_poly, d, nb * bytes_per_poly, cudaMemcpyDeviceToHost, stg.stream));
        CU(cudaEventRecord(stg.ev_down[s], stg.stream));
    }
    CU(cudaStreamSynchronize(stg.stream));
    CU(cudaStreamSynchronize(stg.stream2));
    return CNTT_OK;
}

#define PRIME_HOST_NTT(BITS, T, NAME, BODY)                                                                             \
    CNTT_API int cntt_prime##BITS##_##NAME##_host(const cntt_prime##BITS##_plan* pl, T* h_buf, size_t len, size_t batch) \
    {                                                                                                                    \
        if (!pl || (!h_buf && len)) return CNTT_NULL_POINTER;                                                            \
        if (len != pl->n * batch) return CNTT_LENGTH_MISMATCH;                                                           \
        auto* mpl = const_cast<cntt_prime##BITS##_plan*>(pl);                                                            \
        return host_inplace(mpl->stg, pl->device, h_buf, len * sizeof(T), pl->n * sizeof(T), batch,                      \
                            [&](void* d, size_t nb, cudaStream_t st) -> cudaError_t { BODY });                           \
    }
PRIME_HOST_NTT(32, uint32_t, fwd, return run_ntt32(pl, (uint32_t*)d, nb, true, st);)
PRIME_HOST_NTT(32, uint32_t, inv, return run_ntt32(pl, (uint32_t*)d, nb, false, st);)
PRIME_HOST_NTT(32, uint32_t, fwd_inv, cudaError_t e = run_ntt32(pl, (uint32_t*)d, nb, true, st); if (e != cudaSuccess) return e;
               return run_ntt32(pl, (uint32_t*)d, nb, false, st);)
PRIME_HOST_NTT(64, uint64_t, fwd, return run_ntt64(pl, (uint64_t*)d, nb, true, st);)
PRIME_HOST_NTT(64, uint64_t, inv, return run_ntt64(pl, (uint64_t*)d, nb, false, st);)
PRIME_HOST_NTT(64, uint64_t, fwd_inv, cudaError_t e = run_ntt64(pl, (uint64_t*)d, nb, true, st); if (e != cudaSuccess) return e;
               return run_ntt64(pl, (uint64_t*)d, nb, false, st);)

// ---- one host batch over several devices ---------------------------------------------------------------------------
// plans[g] is the same plan (n, p) built on device g (any devices, duplicates allowed); the batch is cut into nplans contiguous
// shards and every shard runs through its plan's own staging pipeline on its own host thread, so all PCIe links and copy
// engines work at once.  No collective: polynomials are independent.  Returns the first status that is not CNTT_OK.
template <class Fn>
static int host_multi(int nplans, size_t batch, Fn shard_fn)
{
    if (nplans <= 0) return CNTT_NULL_POINTER;
    std::vector<int> rc((size_t)nplans, CNTT_OK);
    std::vector<std::thread> th;
    th.reserve((size_t)nplans);
    for (int g = 0; g < nplans; g++) {
        const size_t b0 = batch * (size_t)g / (size_t)nplans, b1 = batch * (size_t)(g + 1) / (size_t)nplans;
        if (b1 <= b0) continue;
        try {
            th.emplace_back([&, g, b0, b1]() { rc[(size_t)g] = shard_fn(g, b0, b1 - b0); });
        } catch (...) { // no thread to be had (std::system_error): run the shard on the calling thread rather than throw across the C ABI
            rc[(size_t)g] = shard_fn(g, b0, b1 - b0);
        }
    }
    for (auto& t : th) t.join();
    for (int r : rc)
        if (r != CNTT_OK) return r;
    return CNTT_OK;
}
#define PRIME_HOST_MULTI(BITS, T, NAME)                                                                                               \
    CNTT_API int cntt_prime##BITS##_##NAME##_host_multi(const cntt_prime##BITS##_plan* const* plans, int nplans, T* h_buf, size_t len, \
                                                         size_t batch)                                                                 \
    {                                                                                                                                  \
        if (!plans || nplans <= 0 || (!h_buf && len)) return CNTT_NULL_POINTER;                                                        \
        for (int g = 0; g < nplans; g++)                                                                                               \
            if (!plans[g] || plans[g]->n != plans[0]->n || plans[g]->p != plans[0]->p) return g && plans[g] ? CNTT_LENGTH_MISMATCH : CNTT_NULL_POINTER; \
        const size_t n = plans[0]->n;                                                                                                  \
        if (len != n * batch) return CNTT_LENGTH_MISMATCH;                                                                             \
        return host_multi(nplans, batch, [&](int g, size_t b0, size_t nb) {                                                            \
            return cntt_prime##BITS##_##NAME##_host(plans[g], h_buf + b0 * n, nb * n, nb);                                             \
        });                                                                                                                            \
    }
PRIME_HOST_MULTI(32, uint32_t, fwd)
PRIME_HOST_MULTI(32, uint32_t, inv)
PRIME_HOST_MULTI(32, uint32_t, fwd_inv)
PRIME_HOST_MULTI(64, uint64_t, fwd)
PRIME_HOST_MULTI(64, uint64_t, inv)
PRIME_HOST_MULTI(64, uint64_t, fwd_inv)

// pointwise host flavours: simple staged version (streams: dst [+a [+b]])
template <class T, class Fn>
static int host_pointwise(Staging& stg, int device, T* h_dst, const T* h_a, const T* h_b, size_t nwords, Fn launch)
{
    if (nwords == 0) return CNTT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    std::lock_guard<std::mutex> lk(stg.mu);
    const size_t bytes = nwords * sizeof(T);
    const int nstreams = 1 + (h_a ? 1 : 0) + (h_b ? 1 : 0);
    CU(stg.ensure(bytes * nstreams));
    T* d_dst = static_cast<T*>(stg.buf);
    T* d_a = h_a ? d_dst + nwords : nullptr;
    T* d_b = h_b ? d_dst + 2 * nwords : nullptr;
    CU(cudaMemcpyAsync(d_dst, h_dst, bytes, cudaMemcpyHostToDevice, stg.stream));
    if (h_a) CU(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, stg.stream));
    if (h_b) CU(cudaMemcpyAsync(d_b, h_b, bytes, cudaMemcpyHostToDevice, stg.stream));
    cudaError_t e = launch(d_dst, d_a, d_b, stg.stream);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    CU(cudaMemcpyAsync(h_dst, d_dst, bytes, cudaMemcpyDeviceToHost, stg.stream));
    CU(cudaStreamSynchronize(stg.stream));
    return CNTT_OK;
}
#define PRIME_HOST_PW(BITS, T, VEC)                                                                                         \
    CNTT_API int cntt_prime##BITS##_mul_assign_normalize_host(const cntt_prime##BITS##_plan* pl, T* l, const T* r, size_t nwords) \
    {                                                                                                                        \
        if (!pl || ((!l || !r) && nwords)) return CNTT_NULL_POINTER;                                                         \
        if (nwords % VEC) return CNTT_LENGTH_MISMATCH;                                                                       \
        auto* mpl = const_cast<cntt_prime##BITS##_plan*>(pl);                                                                \
        return host_pointwise<T>(mpl->stg, pl->device, l, r, nullptr, nwords, [&](T* d, T* a, T*, cudaStream_t st) {         \
            return run_pw##BITS(pl, OP_MUL_ASSIGN_NORMALIZE, d, a, nullptr, nwords, st);                                     \
        });                                                                                                                  \
    }                                                                                                                        \
    CNTT_API int cntt_prime##BITS##_normalize_host(const cntt_prime##BITS##_plan* pl, T* v, size_t nwords)                   \
    {                                                                                                                        \
        if (!pl || (!v && nwords)) return CNTT_NULL_POINTER;                                                                 \
        if (nwords % VEC) return CNTT_LENGTH_MISMATCH;                                                                       \
        auto* mpl = const_cast<cntt_prime##BITS##_plan*>(pl);                                                                \
        return host_pointwise<T>(mpl->stg, pl->device, v, nullptr, nullptr, nwords, [&](T* d, T*, T*, cudaStream_t st) {     \
            return run_pw##BITS(pl, OP_NORMALIZE, d, nullptr, nullptr, nwords, st);                                          \
        });                                                                                                                  \
    }                                                                                                                        \
    CNTT_API int cntt_prime##BITS##_mul_accumulate_host(const cntt_prime##BITS##_plan* pl, T* acc, const T* l, const T* r, size_t nwords) \
    {                                                                                                                        \
        if (!pl || ((!acc || !l || !r) && nwords)) return CNTT_NULL_POINTER;                                                 \
        if (nwords % VEC) return CNTT_LENGTH_MISMATCH;                                                                       \
        auto* mpl = const_cast<cntt_prime##BITS##_plan*>(pl);                                                                \
        return host_pointwise<T>(mpl->stg, pl->device, acc, l, r, nwords, [&](T* d, T* a, T* b, cudaStream_t st) {           \
            return run_pw##BITS(pl, OP_MUL_ACCUMULATE, d, a, b, nwords, st);                                                 \
        });                                                                                                                  \
    }
PRIME_HOST_PW(32, uint32_t, 4)
PRIME_HOST_PW(64, uint64_t, 2)

// ---- native plans ---------------------------------------------------------------------------------------------------
struct cntt_native_plan {
    size_t n;
    int kind;
    int device;
    int nprimes;
    cntt_prime32_plan* sub[10];
    uint2* d_fused_last; // fwd/inv last-pass tables of the fused kernel's engine for every prime (2 * nprimes * n entries)
    uint32_t* d_bin0;    // binary plans: pass-0 tables of the {0,1} operand (nprimes * kBin0Entries entries), see native_fused.cuh
    NativePlanDev dev;
    Staging stg;
};

static int native_plan_new_impl(size_t n, int word_bits, int binary, int device, int prime_set, cntt_native_plan** out)
{
    if (!out) return CNTT_NULL_POINTER;
    *out = nullptr;
    int kind;
    if (word_bits == 32) kind = binary ? NK_BINARY32 : NK_NATIVE32;
    else if (word_bits == 64) kind = binary ? NK_BINARY64 : NK_NATIVE64;
    else if (word_bits == 128) kind = binary ? NK_BINARY128 : NK_NATIVE128;
    else return CNTT_UNSUPPORTED;
    if (prime_set != 0 && native_num_primes(kind) > kExtPrimes) return CNTT_UNSUPPORTED; // nine extended primes: no native128
    const NativeConsts& c = native_consts(prime_set);
    cntt_native_plan* pl = new cntt_native_plan();
    pl->n = n; pl->kind = kind; pl->device = device; pl->nprimes = native_num_primes(kind);
    for (int k = 0; k < 10; k++) pl->sub[k] = nullptr;
    pl->d_fused_last = nullptr;
    pl->d_bin0 = nullptr;
    // decide Some/None on the host for every prime before touching the device
    for (int k = 0; k < pl->nprimes; k++) {
        uint64_t psi;
        int lg;
        int st = validate(n, c.P[k], 32, &psi, &lg);
        if (st != CNTT_OK) {
            delete pl;
            return st;
        }
    }
    for (int k = 0; k < pl->nprimes; k++) {
        // Plan::try_new(n, P_k)? for every prime (src/native64.rs:933-942): first failure -> None
        int st = build_prime32(n, c.P[k], device, &pl->sub[k]);
        if (st != CNTT_OK) {
            for (int j = 0; j < k; j++) cntt_prime32_plan_free(pl->sub[j]);
            delete pl;
            return st;
        }
    }
    pl->dev.kind = kind;
    pl->dev.logn = pl->sub[0]->logn;
    pl->dev.nprimes = pl->nprimes;
    pl->dev.prime_set = prime_set;
    for (int k = 0; k < pl->nprimes; k++) pl->dev.sub[k] = dev32<A32L4>(pl->sub[k]);
    // the fused polymul (N <= 4096) may carry the product on fewer primes than the plan owns (native.hpp, native_fused_np)
    native_lhs_scale(pl->dev.logn, pl->dev.lscale, prime_set,
                     native_fused_supported(pl->dev.logn) ? native_fused_np(kind, pl->nprimes) : pl->nprimes);
    for (int k = 0; k < 10; k++) { pl->dev.fused_fwd_last[k] = pl->dev.fused_inv_last[k] = nullptr; pl->dev.bin0[k] = nullptr; }
    const bool fused = native_fused_supported(pl->dev.logn), large = native_large_supported(pl->dev.logn);
    if (fused && kind >= NK_BINARY32) {
        // pass-0 tables of the binary operand (native_fused.cuh, binary_pass0): run the first r1 levels of the forward transform
        // (root sub-tree: heap nodes 2^lvl + g, the plan's own twiddles) on unit vectors, then sum the columns per nibble pattern
        const int logn = pl->dev.logn, logr = native_fused_logr(kind, logn), passes = (logn + logr - 1) / logr;
        const int r1 = logn - (passes - 1) * logr, S = 1 << r1, QB = S < 4 ? S : 4, NQ = S / QB;
        std::vector<uint32_t> tab((size_t)pl->nprimes * kBin0Entries, 0u);
        for (int k = 0; k < pl->nprimes && passes >= 2; k++) {
            const uint64_t p = pl->sub[k]->p;
            std::vector<std::vector<uint64_t>> col((size_t)S, std::vector<uint64_t>((size_t)S, 0)); // col[m][j]: output j of unit input m
            for (int mi = 0; mi < S; mi++) {
                std::vector<uint64_t>& v = col[(size_t)mi];
                v[(size_t)mi] = 1;
                for (int lvl = 0; lvl < r1; lvl++) {
                    const int half = S >> (lvl + 1);
                    for (int g = 0; g < (1 << lvl); g++) {
                        const uint64_t w = pl->sub[k]->head_fwd.e[(1 << lvl) + g].x;
                        for (int u = 0; u < half; u++) {
                            const uint64_t a = v[(size_t)(2 * half * g + u)], b = v[(size_t)(2 * half * g + u + half)] * w % p;
                            v[(size_t)(2 * half * g + u)] = (a + b) % p;
                            v[(size_t)(2 * half * g + u + half)] = (a + p - b) % p;
                        }
                    }
                }
            }
            uint32_t* t = tab.data() + (size_t)k * kBin0Entries;
            for (int q = 0; q < NQ; q++)
                for (int pat = 0; pat < 16; pat++)
                    for (int j = 0; j < S; j++) {
                        uint64_t acc = 0;
                        for (int i = 0; i < QB; i++)
                            if (pat >> i & 1) acc = (acc + col[(size_t)(QB * q + i)][(size_t)j]) % p;
                        t[(q * 16 + pat) * S + j] = (uint32_t)acc;
                    }
        }
        DeviceGuard g(device);
        cudaError_t e = g.ok ? cudaMalloc(&pl->d_bin0, tab.size() * sizeof(uint32_t)) : cudaErrorInvalidDevice;
        if (e == cudaSuccess) e = cudaMemcpy(pl->d_bin0, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cntt_native_plan_free(pl);
            return cuda_fail(e, "native plan upload");
        }
        for (int k = 0; k < pl->nprimes; k++) pl->dev.bin0[k] = pl->d_bin0 + (size_t)k * kBin0Entries;
    }
    if (fused || large) {
        DeviceGuard g(device);
        cudaError_t e = g.ok ? cudaMalloc(&pl->d_fused_last, 2 * (size_t)pl->nprimes * n * sizeof(uint2)) : cudaErrorInvalidDevice;
        for (int k = 0; k < pl->nprimes && e == cudaSuccess; k++) {
            uint2* f = pl->d_fused_last + (size_t)(2 * k) * n;
            uint2* i = f + n;
            if ((e = fused ? native_fused_build_last(kind, pl->dev.logn, pl->sub[k]->d_fwd, f, nullptr)
                           : native_large_build_last(pl->dev.logn, pl->sub[k]->d_fwd, f, nullptr)) != cudaSuccess) break;
            if ((e = fused ? native_fused_build_last(kind, pl->dev.logn, pl->sub[k]->d_inv, i, nullptr)
                           : native_large_build_last(pl->dev.logn, pl->sub[k]->d_inv, i, nullptr)) != cudaSuccess) break;
            pl->dev.fused_fwd_last[k] = f;
            pl->dev.fused_inv_last[k] = i;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) {
            cntt_native_plan_free(pl);
            return cuda_fail(e, "native plan upload");
        }
    }
    *out = pl;
    return CNTT_OK;
}
CNTT_API int cntt_native_plan_new(size_t n, int word_bits, int binary, int device, cntt_native_plan** out)
{
    return native_plan_new_impl(n, word_bits, binary, device, 0, out);
}
CNTT_API int cntt_native_plan_new_ext(size_t n, int word_bits, int binary, int device, cntt_native_plan** out)
{
    return native_plan_new_impl(n, word_bits, binary, device, 1, out);
}
CNTT_API void cntt_native_plan_free(cntt_native_plan* pl)
{
    if (!pl) return;
    {
        DeviceGuard g(pl->device);
        pl->stg.release();
        if (pl->d_fused_last) cudaFree(pl->d_fused_last);
        if (pl->d_bin0) cudaFree(pl->d_bin0);
    }
    for (int k = 0; k < pl->nprimes; k++) if (pl->sub[k]) cntt_prime32_plan_free(pl->sub[k]);
    delete pl;
}
CNTT_API size_t cntt_native_ntt_size(const cntt_native_plan* pl) { return pl ? pl->n : 0; }
CNTT_API int cntt_native_num_primes(const cntt_native_plan* pl) { return pl ? pl->nprimes : 0; }
CNTT_API uint32_t cntt_native_prime(const cntt_native_plan* pl, int i) { return (pl && i >= 0 && i < pl->nprimes) ? pl->sub[i]->p : 0; }

static cudaError_t native_fwd_impl(const cntt_native_plan* pl, const void* value, uint32_t* planes, size_t batch, bool binary_copy, cudaStream_t st)
{
    const size_t nwords = pl->n * batch;
    cudaError_t e = native_split_fused(pl->dev, const_cast<void*>(value), planes, nwords, batch, binary_copy ? 1 : 0, st);
    if (e != cudaErrorNotSupported) return e;
    (void)cudaGetLastError();
    e = native_reduce(pl->dev, value, planes, nwords, nwords, binary_copy, st);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < pl->nprimes; k++)
        if ((e = ntt_A32L4(pl->dev.sub[k], planes + (size_t)k * nwords, batch, true, st)) != cudaSuccess) return e;
    return cudaSuccess;
}
static cudaError_t native_inv_impl(const cntt_native_plan* pl, void* value, uint32_t* planes, size_t batch, cudaStream_t st)
{
    const size_t nwords = pl->n * batch;
    cudaError_t e;
    for (int k = 0; k < pl->nprimes; k++)
        if ((e = ntt_A32L4(pl->dev.sub[k], planes + (size_t)k * nwords, batch, false, st)) != cudaSuccess) return e;
    return native_crt(pl->dev, value, planes, nwords, nwords, st);
}

CNTT_API int cntt_native_fwd(const cntt_native_plan* pl, const void* d_value, uint32_t* d_mod_p, size_t batch, void* stream)
{
    if (!pl || ((!d_value || !d_mod_p) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_value, d_mod_p)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(native_fwd_impl(pl, d_value, d_mod_p, batch, false, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_native_fwd_binary(const cntt_native_plan* pl, const void* d_value, uint32_t* d_mod_p, size_t batch, void* stream)
{
    if (!pl || ((!d_value || !d_mod_p) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_value, d_mod_p)) return CNTT_MISALIGNED;
    if (pl->kind < NK_BINARY32) return CNTT_UNSUPPORTED; // only native_binary* have fwd_binary
    GUARD(pl->device);
    CU(native_fwd_impl(pl, d_value, d_mod_p, batch, true, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_native_inv(const cntt_native_plan* pl, void* d_value, uint32_t* d_mod_p, size_t batch, void* stream)
{
    if (!pl || ((!d_value || !d_mod_p) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_value, d_mod_p)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(native_inv_impl(pl, d_value, d_mod_p, batch, (cudaStream_t)stream));
    return CNTT_OK;
}

// Stream-ordered scratch comes from a library-owned memory pool per device whose release threshold is "never":
// with the default pool every synchronisation hands the freed planes back to the driver, and the next call
// pays for mapping ~1 GiB again (measured: native64 N=32768 batch 1024, 2.2 ms -> 5.4 ms per call).
static cudaError_t scratch_pool(int device, cudaMemPool_t* out)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&pools[device], &props);
        if (e != cudaSuccess) { pools[device] = nullptr; return e; }
        unsigned long long keep = ~0ull;
        if ((e = cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
    }
    *out = pools[device];
    return cudaSuccess;
}

// Pipelines through residue planes (N > 4096): the planes live in a stream-ordered scratch allocation, processed in
// batch chunks so the scratch stays bounded.  4096 < N <= 32768 (every size the native plans accept beyond the fused
// kernel) runs the three-kernel path of native_large.cuh; the plan-API composition below it is the generic fallback.
static cudaError_t native_polymul_unfused(const cntt_native_plan* pl, void* prod, const void* lhs, const void* rhs, size_t batch, cudaStream_t st)
{
    const size_t n = pl->n;
    const int np = pl->nprimes;
    const size_t wb = (size_t)native_word_bytes(pl->kind);
    const size_t per_poly = 2 * (size_t)np * n * sizeof(uint32_t);
    const size_t budget = (size_t)1 << 30;
    const size_t chunk = std::max<size_t>(1, std::min(batch, budget / per_poly));
    uint32_t* scratch = nullptr;
    cudaMemPool_t pool;
    cudaError_t e = scratch_pool(pl->device, &pool);
    if (e != cudaSuccess) return e;
    if ((e = cudaMallocFromPoolAsync((void**)&scratch, chunk * per_poly, pool, st)) != cudaSuccess) return e;
    const bool binary = pl->kind >= NK_BINARY32;
    for (size_t b0 = 0; b0 < batch && e == cudaSuccess; b0 += chunk) {
        const size_t nb = std::min(chunk, batch - b0);
        const size_t nwords = nb * n;
        uint32_t* L = scratch;
        uint32_t* R = scratch + (size_t)np * nwords;
        const char* l = static_cast<const char*>(lhs) + b0 * n * wb;
        const char* r = static_cast<const char*>(rhs) + b0 * n * wb;
        char* o = static_cast<char*>(prod) + b0 * n * wb;
        if (native_large_supported(pl->dev.logn)) { // three fused kernels around the planes (native_large.cuh)
            if ((e = native_polymul_large(pl->dev, o, l, r, nb, L, R, st)) != cudaSuccess) break;
            continue;
        }
        if ((e = native_fwd_impl(pl, l, L, nb, false, st)) != cudaSuccess) break;
        if ((e = native_fwd_impl(pl, r, R, nb, binary, st)) != cudaSuccess) break;
        for (int k = 0; k < np && e == cudaSuccess; k++)
            e = pointwise_A32L4(pl->dev.sub[k], OP_MUL_ASSIGN_NORMALIZE, L + (size_t)k * nwords, R + (size_t)k * nwords, nullptr, nwords, st);
        if (e != cudaSuccess) break;
        e = native_inv_impl(pl, o, L, nb, st);
    }
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

static cudaError_t native_polymul_impl(const cntt_native_plan* pl, void* prod, const void* lhs, const void* rhs, size_t batch, cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    cudaError_t e = native_polymul_fused(pl->dev, prod, lhs, rhs, batch, st);
    if (e != cudaErrorNotSupported) return e;
    (void)cudaGetLastError();
    return native_polymul_unfused(pl, prod, lhs, rhs, batch, st);
}

// EXTENSION: negacyclic_polymul with the rhs operand already transformed (TFHE keeps its bootstrapping key in the NTT domain).
// d_rhs_planes: what cntt_native_fwd (cntt_native_fwd_binary for binary plans) wrote for `rhs_batch` polynomials -- plane k of key b at
// d_rhs_planes[(k * rhs_batch + b) * n]; rhs_batch == batch (one key per product) or 1 (one key for the whole batch).
CNTT_API int cntt_native_polymul_ntt_rhs(const cntt_native_plan* pl, void* d_prod, const void* d_lhs, const uint32_t* d_rhs_planes, size_t rhs_batch,
                                         size_t batch, void* stream)
{
    if (!pl || ((!d_prod || !d_lhs || !d_rhs_planes) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_prod, d_lhs, d_rhs_planes)) return CNTT_MISALIGNED;
    if (rhs_batch != batch && rhs_batch != 1) return CNTT_LENGTH_MISMATCH;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    const cudaError_t e = native_polymul_fused_pre(pl->dev, d_prod, d_lhs, d_rhs_planes, batch, rhs_batch * pl->n, rhs_batch == 1 ? 0 : pl->n, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) { (void)cudaGetLastError(); return CNTT_UNSUPPORTED; } // N < 256 or N > 4096: use fwd / mul_assign_normalize / inv
    CU(e);
    return CNTT_OK;
}
CNTT_API int cntt_native_polymul(const cntt_native_plan* pl, void* d_prod, const void* d_lhs, const void* d_rhs, size_t batch, void* stream)
{
    if (!pl || ((!d_prod || !d_lhs || !d_rhs) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_prod, d_lhs, d_rhs)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(native_polymul_impl(pl, d_prod, d_lhs, d_rhs, batch, (cudaStream_t)stream));
    return CNTT_OK;
}

CNTT_API int cntt_native_polymul_host(const cntt_native_plan* pl, void* h_prod, const void* h_lhs, const void* h_rhs, size_t len, size_t batch)
{
    if (!pl || ((!h_prod || !h_lhs || !h_rhs) && len)) return CNTT_NULL_POINTER;
    if (len != pl->n * batch) return CNTT_LENGTH_MISMATCH; // assert_eq!(n, lhs.len()) (src/native64.rs:1043-1045)
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    auto* mpl = const_cast<cntt_native_plan*>(pl);
    Staging& stg = mpl->stg;
    std::lock_guard<std::mutex> lk(stg.mu);
    const size_t wb = (size_t)native_word_bytes(pl->kind);
    const size_t ppb = pl->n * wb; // bytes per polynomial
    const size_t target = (size_t)32 << 20;
    const size_t chunk = std::max<size_t>(1, std::min(batch, target / ppb));
    const size_t slot = 3 * chunk * ppb; // lhs, rhs, prod
    CU(stg.ensure(2 * slot));
    const size_t nchunks = (batch + chunk - 1) / chunk;
    for (size_t c = 0; c < nchunks; c++) {
        const size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
        char* base = static_cast<char*>(stg.buf) + (c & 1) * slot;
        char *dl = base, *dr = base + chunk * ppb, *dp = base + 2 * chunk * ppb;
        if (c >= 2) CU(cudaStreamWaitEvent(stg.stream2, stg.ev[2 + (c & 1)], 0));
        CU(cudaMemcpyAsync(dl, static_cast<const char*>(h_lhs) + b0 * ppb, nb * ppb, cudaMemcpyHostToDevice, stg.stream2));
        CU(cudaMemcpyAsync(dr, static_cast<const char*>(h_rhs) + b0 * ppb, nb * ppb, cudaMemcpyHostToDevice, stg.stream2));
        CU(cudaEventRecord(stg.ev[c & 1], stg.stream2));
        CU(cudaStreamWaitEvent(stg.stream, stg.ev[c & 1], 0));
        CU(native_polymul_impl(pl, dp, dl, dr, nb, stg.stream));
        CU(cudaMemcpyAsync(static_cast<char*>(h_prod) + b0 * ppb, dp, nb * ppb, cudaMemcpyDeviceToHost, stg.stream));
        CU(cudaEventRecord(stg.ev[2 + (c & 1)], stg.stream));
    }
    CU(cudaStreamSynchronize(stg.stream));
    CU(cudaStreamSynchronize(stg.stream2));
    return CNTT_OK;
}
// one host batch of polymuls over several devices (see host_multi above)
CNTT_API int cntt_native_polymul_host_multi(const cntt_native_plan* const* plans, int nplans, void* h_prod, const void* h_lhs, const void* h_rhs,
                                            size_t len, size_t batch)
{
    if (!plans || nplans <= 0 || ((!h_prod || !h_lhs || !h_rhs) && len)) return CNTT_NULL_POINTER;
    for (int g = 0; g < nplans; g++)
        if (!plans[g] || plans[g]->n != plans[0]->n || plans[g]->kind != plans[0]->kind) return g && plans[g] ? CNTT_LENGTH_MISMATCH : CNTT_NULL_POINTER;
    const size_t n = plans[0]->n, wb = (size_t)native_word_bytes(plans[0]->kind);
    if (len != n * batch) return CNTT_LENGTH_MISMATCH;
    return host_multi(nplans, batch, [&](int g, size_t b0, size_t nb) {
        return cntt_native_polymul_host(plans[g], (char*)h_prod + b0 * n * wb, (const char*)h_lhs + b0 * n * wb, (const char*)h_rhs + b0 * n * wb, nb * n, nb);
    });
}

// host-slice flavours of Plan32::fwd / fwd_binary / inv: the reference's call shape (value: n words, mod_p_k: n u32 each;
// here `batch` of them concatenated, plane k of polynomial b at h_mod_p[(k * batch + b) * n]).  Staged through the plan's
// arena, synchronous.  inv copies the planes back as well: the reference's inv leaves the inverse transforms in them.
static int native_split_host(const cntt_native_plan* pl, void* h_value, uint32_t* h_mod_p, size_t len, size_t batch, int what)
{
    if (!pl || ((!h_value || !h_mod_p) && len)) return CNTT_NULL_POINTER;
    if (len != pl->n * batch) return CNTT_LENGTH_MISMATCH; // assert_eq!(value.len(), n) (src/native64.rs:983-989)
    if (what == 1 && pl->kind < NK_BINARY32) return CNTT_UNSUPPORTED;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    auto* mpl = const_cast<cntt_native_plan*>(pl);
    Staging& stg = mpl->stg;
    std::lock_guard<std::mutex> lk(stg.mu);
    const size_t vbytes = len * (size_t)native_word_bytes(pl->kind);
    const size_t pbytes = (size_t)pl->nprimes * len * sizeof(uint32_t);
    CU(stg.ensure(vbytes + pbytes));
    char* dv = static_cast<char*>(stg.buf);
    uint32_t* dp = reinterpret_cast<uint32_t*>(dv + vbytes);
    cudaStream_t st = stg.stream;
    if (what <= 1) { // fwd / fwd_binary: value in, planes out
        CU(cudaMemcpyAsync(dv, h_value, vbytes, cudaMemcpyHostToDevice, st));
        CU(native_fwd_impl(pl, dv, dp, batch, what == 1, st));
        CU(cudaMemcpyAsync(h_mod_p, dp, pbytes, cudaMemcpyDeviceToHost, st));
    } else {         // inv: planes in (clobbered), value out
        CU(cudaMemcpyAsync(dp, h_mod_p, pbytes, cudaMemcpyHostToDevice, st));
        CU(native_inv_impl(pl, dv, dp, batch, st));
        CU(cudaMemcpyAsync(h_value, dv, vbytes, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_mod_p, dp, pbytes, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return CNTT_OK;
}
CNTT_API int cntt_native_fwd_host(const cntt_native_plan* pl, const void* h_value, uint32_t* h_mod_p, size_t len, size_t batch)
{
    return native_split_host(pl, const_cast<void*>(h_value), h_mod_p, len, batch, 0);
}
CNTT_API int cntt_native_fwd_binary_host(const cntt_native_plan* pl, const void* h_value, uint32_t* h_mod_p, size_t len, size_t batch)
{
    return native_split_host(pl, const_cast<void*>(h_value), h_mod_p, len, batch, 1);
}
CNTT_API int cntt_native_inv_host(const cntt_native_plan* pl, void* h_value, uint32_t* h_mod_p, size_t len, size_t batch)
{
    return native_split_host(pl, h_value, h_mod_p, len, batch, 2);
}

// ---- Plan52 twins: native32 / native64 / native_binary32 / native_binary64 ::Plan52 ---------------------------------
#include "native52_kernels.cuh"

struct cntt_native52_plan {
    size_t n;
    int device;
    int binary;
    Native52Consts c;
    cntt_prime64_plan* sub[3];
    cntt_native_plan* inner; // the Plan32 of the same kind: negacyclic_polymul results are identical (exact product, wrapped)
};

CNTT_API void cntt_native52_plan_free(cntt_native52_plan* pl)
{
    if (!pl) return;
    for (int k = 0; k < 3; k++)
        if (pl->sub[k]) cntt_prime64_plan_free(pl->sub[k]);
    if (pl->inner) cntt_native_plan_free(pl->inner);
    delete pl;
}
CNTT_API int cntt_native52_plan_new(size_t n, int word_bits, int binary, int device, cntt_native52_plan** out)
{
    // primes52::P0..P2 (src/lib.rs:600-602)
    static const uint64_t P52[3] = {0x3FFFFFE770001ull, 0x3FFFFFEB90001ull, 0x3FFFFFEC80001ull};
    if (!out) return CNTT_NULL_POINTER;
    *out = nullptr;
    if (word_bits != 32 && word_bits != 64) return CNTT_UNSUPPORTED; // no 128-bit Plan52 in the reference
    // native32: P0,P1 (native32.rs:441-445); native64: P0..P2 (native64.rs:1078-1087); binary32: P0
    // (native_binary32.rs:270-274); binary64: P0,P1 (native_binary64.rs:453-457)
    const int np = word_bits == 32 ? (binary ? 1 : 2) : (binary ? 2 : 3);
    cntt_native52_plan* pl = new cntt_native52_plan();
    pl->n = n; pl->device = device; pl->binary = binary; pl->inner = nullptr;
    for (int k = 0; k < 3; k++) pl->sub[k] = nullptr;
    for (int k = 0; k < np; k++) {
        int st = build_prime64(n, P52[k], device, &pl->sub[k]); // Plan::try_new(n, P_k)?  -> None on the first failure
        if (st != CNTT_OK) { cntt_native52_plan_free(pl); return st; }
    }
    int st = cntt_native_plan_new(n, word_bits, binary, device, &pl->inner);
    if (st != CNTT_OK) { cntt_native52_plan_free(pl); return st; }
    Native52Consts& c = pl->c;
    c.np = np; c.word_bytes = word_bits / 8; c.reduce = word_bits == 64;
    uint64_t pre = 1;
    for (int k = 0; k < 3; k++) { c.p[k] = k < np ? P52[k] : 1; c.inv[k] = 0; c.pre[k] = 0; }
    for (int k = 0; k < np; k++) {
        c.pre[k] = pre;
        if (k >= 1) {
            host::Fp f(P52[k]);
            uint64_t m = 1;
            for (int j = 0; j < k; j++) m = f.mul(m, P52[j] % P52[k]);
            c.inv[k] = f.inv(m);
        }
        pre *= P52[k]; // wrapping, like the reference's P0.wrapping_mul(P1) (native64.rs:793-794)
    }
    c.full = pre;
    *out = pl;
    return CNTT_OK;
}
CNTT_API size_t cntt_native52_ntt_size(const cntt_native52_plan* pl) { return pl ? pl->n : 0; }
CNTT_API int cntt_native52_num_primes(const cntt_native52_plan* pl) { return pl ? pl->c.np : 0; }
CNTT_API uint64_t cntt_native52_prime(const cntt_native52_plan* pl, int i) { return (pl && i >= 0 && i < pl->c.np) ? pl->c.p[i] : 0; }

static int native52_fwd(const cntt_native52_plan* pl, const void* d_value, uint64_t* d_mod_p, size_t batch, bool copy, cudaStream_t st)
{
    if (!pl || ((!d_value || !d_mod_p) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_value, d_mod_p)) return CNTT_MISALIGNED;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    const unsigned long long nwords = (unsigned long long)pl->n * batch;
    const unsigned g = (unsigned)std::min<unsigned long long>((nwords + 255) / 256, 148ull * 16ull);
    if (copy) k_native52_reduce<true><<<g, 256, 0, st>>>(pl->c, d_value, d_mod_p, nwords, nwords);
    else k_native52_reduce<false><<<g, 256, 0, st>>>(pl->c, d_value, d_mod_p, nwords, nwords);
    CU(cudaGetLastError());
    for (int k = 0; k < pl->c.np; k++) CU(run_ntt64(pl->sub[k], d_mod_p + (size_t)k * nwords, batch, true, st));
    return CNTT_OK;
}
CNTT_API int cntt_native52_fwd(const cntt_native52_plan* pl, const void* d_value, uint64_t* d_mod_p, size_t batch, void* stream)
{
    return native52_fwd(pl, d_value, d_mod_p, batch, false, (cudaStream_t)stream);
}
CNTT_API int cntt_native52_fwd_binary(const cntt_native52_plan* pl, const void* d_value, uint64_t* d_mod_p, size_t batch, void* stream)
{
    if (pl && !pl->binary) return CNTT_UNSUPPORTED; // only the native_binary* Plan52 have fwd_binary
    return native52_fwd(pl, d_value, d_mod_p, batch, true, (cudaStream_t)stream);
}
CNTT_API int cntt_native52_inv(const cntt_native52_plan* pl, void* d_value, uint64_t* d_mod_p, size_t batch, void* stream)
{
    if (!pl || ((!d_value || !d_mod_p) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_value, d_mod_p)) return CNTT_MISALIGNED;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long nwords = (unsigned long long)pl->n * batch;
    for (int k = 0; k < pl->c.np; k++) CU(run_ntt64(pl->sub[k], d_mod_p + (size_t)k * nwords, batch, false, st));
    const unsigned g = (unsigned)std::min<unsigned long long>((nwords + 255) / 256, 148ull * 16ull);
    k_native52_crt<<<g, 256, 0, st>>>(pl->c, d_value, d_mod_p, nwords, nwords);
    CU(cudaGetLastError());
    return CNTT_OK;
}
// host-slice flavours of Plan52::fwd / fwd_binary / inv (the reference's call shape, src/native64.rs:1108-1141), staged through
// the arena of the inner Plan32 handle: what = 0 fwd, 1 fwd_binary, 2 inv (planes clobbered and returned, like cntt_native_inv_host)
static int native52_split_host(const cntt_native52_plan* pl, void* h_value, uint64_t* h_mod_p, size_t len, size_t batch, int what)
{
    if (!pl || ((!h_value || !h_mod_p) && len)) return CNTT_NULL_POINTER;
    if (len != pl->n * batch) return CNTT_LENGTH_MISMATCH;
    if (what == 1 && !pl->binary) return CNTT_UNSUPPORTED;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    Staging& stg = pl->inner->stg;
    std::lock_guard<std::mutex> lk(stg.mu);
    const size_t vbytes = len * (size_t)native_word_bytes(pl->inner->kind);
    const size_t vpad = (vbytes + 15) & ~(size_t)15;
    const size_t pbytes = (size_t)pl->c.np * len * sizeof(uint64_t);
    CU(stg.ensure(vpad + pbytes));
    char* dv = static_cast<char*>(stg.buf);
    uint64_t* dp = reinterpret_cast<uint64_t*>(dv + vpad);
    cudaStream_t st = stg.stream;
    int rc;
    if (what <= 1) {
        CU(cudaMemcpyAsync(dv, h_value, vbytes, cudaMemcpyHostToDevice, st));
        if ((rc = native52_fwd(pl, dv, dp, batch, what == 1, st)) != CNTT_OK) return rc;
        CU(cudaMemcpyAsync(h_mod_p, dp, pbytes, cudaMemcpyDeviceToHost, st));
    } else {
        CU(cudaMemcpyAsync(dp, h_mod_p, pbytes, cudaMemcpyHostToDevice, st));
        if ((rc = cntt_native52_inv(pl, dv, dp, batch, st)) != CNTT_OK) return rc;
        CU(cudaMemcpyAsync(h_value, dv, vbytes, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_mod_p, dp, pbytes, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return CNTT_OK;
}
CNTT_API int cntt_native52_fwd_host(const cntt_native52_plan* pl, const void* h_value, uint64_t* h_mod_p, size_t len, size_t batch)
{
    return native52_split_host(pl, const_cast<void*>(h_value), h_mod_p, len, batch, 0);
}
CNTT_API int cntt_native52_fwd_binary_host(const cntt_native52_plan* pl, const void* h_value, uint64_t* h_mod_p, size_t len, size_t batch)
{
    return native52_split_host(pl, const_cast<void*>(h_value), h_mod_p, len, batch, 1);
}
CNTT_API int cntt_native52_inv_host(const cntt_native52_plan* pl, void* h_value, uint64_t* h_mod_p, size_t len, size_t batch)
{
    return native52_split_host(pl, h_value, h_mod_p, len, batch, 2);
}
CNTT_API int cntt_native52_polymul(const cntt_native52_plan* pl, void* d_prod, const void* d_lhs, const void* d_rhs, size_t batch, void* stream)
{
    if (!pl) return CNTT_NULL_POINTER;
    return cntt_native_polymul(pl->inner, d_prod, d_lhs, d_rhs, batch, stream);
}
CNTT_API int cntt_native52_polymul_host(const cntt_native52_plan* pl, void* h_prod, const void* h_lhs, const void* h_rhs, size_t len, size_t batch)
{
    if (!pl) return CNTT_NULL_POINTER;
    return cntt_native_polymul_host(pl->inner, h_prod, h_lhs, h_rhs, len, batch);
}

// ---- product::Plan (src/product.rs) --------------------------------------------------------------------------
#include "product_kernels.cuh"
#include "product_fused.hpp"

struct cntt_product_plan {
    size_t n;
    uint64_t modulus;
    int device;
    ProductConsts c;
    cntt_prime32_plan* p32[kProductMaxPrimes];
    cntt_prime64_plan* p64[kProductMaxPrimes];
    uint2* d_pf_last[2][2]; // [fwd, inv][prime]: last-pass tables of the fused kernels where they differ from the prime plans' (product_fused.hpp)
    Staging stg;
};

static void product_free(cntt_product_plan* pl)
{
    for (int d = 0; d < 2; d++)
        for (int j = 0; j < 2; j++)
            if (pl->d_pf_last[d][j]) cudaFree(pl->d_pf_last[d][j]);
    for (int k = 0; k < kProductMaxPrimes; k++) {
        if (pl->p32[k]) cntt_prime32_plan_free(pl->p32[k]);
        if (pl->p64[k]) cntt_prime64_plan_free(pl->p64[k]);
    }
    pl->stg.release();
    delete pl;
}

static bool product_fused_args(const cntt_product_plan* pl, ProductFusedArgs* a);

// try_new (product.rs:152-251): None (a status != OK) for odd n, zero or duplicate factors, a product of the
// factors (1s skipped, checked multiplication) different from `modulus`, or any factor rejected by its prime plan.
CNTT_API int cntt_product_plan_new(size_t n, uint64_t modulus, const uint64_t* factors, size_t nfactors, int device, cntt_product_plan** out)
{
    if (!out || (!factors && nfactors)) return CNTT_NULL_POINTER;
    *out = nullptr;
    if (n % 2 != 0) return CNTT_INVALID_SIZE;
    std::vector<uint64_t> pr(factors, factors + nfactors);
    std::sort(pr.begin(), pr.end());
    uint64_t prev = 0;
    for (uint64_t f : pr) {
        if (f == prev) return CNTT_INVALID_MODULUS; // zero or duplicate
        prev = f;
    }
    pr.erase(pr.begin(), std::find_if(pr.begin(), pr.end(), [](uint64_t f) { return f != 1; }));
    unsigned __int128 prod = 1;
    for (uint64_t f : pr) {
        prod *= f;
        if (prod >> 64) return CNTT_INVALID_MODULUS; // checked_mul overflow
    }
    if ((uint64_t)prod != modulus) return CNTT_INVALID_MODULUS;
    for (uint64_t f : pr) { // every None of the reference first (host-only checks) ...
        uint64_t psi;
        int lg;
        const int st = validate(n, f, f < (1ull << 32) ? 32 : 16, &psi, &lg);
        if (st != CNTT_OK) return st;
    }
    if (pr.size() > (size_t)kProductMaxPrimes) return CNTT_UNSUPPORTED; // ... then the implementation limit (cannot happen for accepted sizes)
    cntt_product_plan* pl = new cntt_product_plan();
    pl->n = n; pl->modulus = modulus; pl->device = device;
    for (int k = 0; k < kProductMaxPrimes; k++) { pl->p32[k] = nullptr; pl->p64[k] = nullptr; }
    for (int d = 0; d < 2; d++) pl->d_pf_last[d][0] = pl->d_pf_last[d][1] = nullptr;
    ProductConsts& c = pl->c;
    std::memset(&c, 0, sizeof(c));
    c.modulus = modulus; c.n = n;
    for (uint64_t f : pr) {
        int st;
        if (f < (1ull << 32)) st = build_prime32(n, (uint32_t)f, device, &pl->p32[c.count32]);
        else st = build_prime64(n, f, device, &pl->p64[c.count64]);
        if (st != CNTT_OK) { product_free(pl); return st; }
        const int j = c.count32 + c.count64;
        c.p[j] = f;
        c.recip[j] = ~0ull / f;
        for (int l = 0; l < 2; l++) {
            const uint64_t cst = l == 0 ? 1 % f : (((uint64_t)1 << 32) % f);
            c.red32[j][l][0] = f < ((uint64_t)1 << 32) ? (uint32_t)cst : 0;
            c.red32[j][l][1] = f < ((uint64_t)1 << 32) ? (uint32_t)((cst << 32) / f) : 0;
        }
        host::Fp fp(f);
        for (int i = 0; i < j; i++) c.inv[j][i] = fp.inv(c.p[i] % f);
        if (f < (1ull << 32)) c.count32++; else c.count64++;
    }
    c.domain_len = (n / 2) * c.count32 + n * c.count64;
    if (c.count32 == 2 && c.p[1] < (1ull << 31)) {
        c.inv10_32[0] = (uint32_t)c.inv[1][0];
        c.inv10_32[1] = (uint32_t)((c.inv[1][0] << 32) / c.p[1]);
    }
    if (ProductFusedArgs fa; product_fused_args(pl, &fa) && product_fused_own_tables(fa.logn)) {
        DeviceGuard guard(device);
        cudaError_t e = guard.ok ? cudaSuccess : cudaGetLastError();
        for (int d = 0; d < 2 && e == cudaSuccess; d++)
            for (int j = 0; j < 2 && e == cudaSuccess; j++) {
                if ((e = cudaMalloc(&pl->d_pf_last[d][j], n * sizeof(uint2))) != cudaSuccess) break;
                e = product_fused_build_last(fa.cls, fa.logn, d == 0 ? fa.tw_fwd[j] : fa.tw_inv[j], pl->d_pf_last[d][j], nullptr);
            }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { product_free(pl); return cuda_fail(e, "product plan tables"); }
    }
    *out = pl;
    return CNTT_OK;
}
CNTT_API void cntt_product_plan_free(cntt_product_plan* pl) { if (pl) product_free(pl); }
CNTT_API size_t cntt_product_ntt_size(const cntt_product_plan* pl) { return pl ? pl->n : 0; }
CNTT_API uint64_t cntt_product_modulus(const cntt_product_plan* pl) { return pl ? pl->modulus : 0; }
CNTT_API size_t cntt_product_ntt_domain_len(const cntt_product_plan* pl) { return pl ? (size_t)pl->c.domain_len : 0; }
CNTT_API int cntt_product_num_primes(const cntt_product_plan* pl, int* count32, int* count64)
{
    if (!pl) return CNTT_NULL_POINTER;
    if (count32) *count32 = pl->c.count32;
    if (count64) *count64 = pl->c.count64;
    return CNTT_OK;
}
CNTT_API uint64_t cntt_product_prime(const cntt_product_plan* pl, int i) { return (pl && i >= 0 && i < pl->c.count32 + pl->c.count64) ? pl->c.p[i] : 0; }

static cudaError_t run_ntt32_strided(const cntt_prime32_plan* pl, uint32_t* d, size_t batch, size_t stride, bool fwd, cudaStream_t st)
{
    switch (pl->cls) {
    case C32_L4: return ntt_strided_A32L4(dev32<A32L4>(pl), d, batch, stride, fwd, st);
    case C32_L2: return ntt_strided_A32L2(dev32<A32L2>(pl), d, batch, stride, fwd, st);
    default: return ntt_strided_A32G(dev32<A32G>(pl), d, batch, stride, fwd, st);
    }
}
static cudaError_t run_ntt64_strided(const cntt_prime64_plan* pl, uint64_t* d, size_t batch, size_t stride, bool fwd, cudaStream_t st)
{
    switch (pl->cls) {
    case C64_L4: return ntt_strided_A64L4(dev64<A64L4>(pl), d, batch, stride, fwd, st);
    case C64_L2: return ntt_strided_A64L2(dev64<A64L2>(pl), d, batch, stride, fwd, st);
    case C64_S: return ntt_strided_A64S(dev64<A64S>(pl), d, batch, stride, fwd, st);
    default: return ntt_strided_A64G(dev64<A64G>(pl), d, batch, stride, fwd, st);
    }
}
static cudaError_t run_pw32_strided(const cntt_prime32_plan* pl, int op, uint32_t* dst, const uint32_t* a, const uint32_t* b, size_t batch, size_t stride, cudaStream_t st)
{
    switch (pl->cls) {
    case C32_L4: return pointwise_strided_A32L4(dev32<A32L4>(pl), op, dst, a, b, batch, stride, st);
    case C32_L2: return pointwise_strided_A32L2(dev32<A32L2>(pl), op, dst, a, b, batch, stride, st);
    default: return pointwise_strided_A32G(dev32<A32G>(pl), op, dst, a, b, batch, stride, st);
    }
}
static cudaError_t run_pw64_strided(const cntt_prime64_plan* pl, int op, uint64_t* dst, const uint64_t* a, const uint64_t* b, size_t batch, size_t stride, cudaStream_t st)
{
    switch (pl->cls) {
    case C64_L4: return pointwise_strided_A64L4(dev64<A64L4>(pl), op, dst, a, b, batch, stride, st);
    case C64_L2: return pointwise_strided_A64L2(dev64<A64L2>(pl), op, dst, a, b, batch, stride, st);
    case C64_S: return pointwise_strided_A64S(dev64<A64S>(pl), op, dst, a, b, batch, stride, st);
    default: return pointwise_strided_A64G(dev64<A64G>(pl), op, dst, a, b, batch, stride, st);
    }
}
// the per-prime transforms of every plane of the batch (planes of one polynomial are domain_len u64 apart)
static cudaError_t product_planes_ntt(const cntt_product_plan* pl, uint64_t* ntt, size_t batch, bool fwd, cudaStream_t st)
{
    const ProductConsts& c = pl->c;
    cudaError_t e;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(ntt);
    uint64_t* d64 = ntt + (c.n / 2) * c.count32;
    for (int k = 0; k < c.count32; k++)
        if ((e = run_ntt32_strided(pl->p32[k], d32 + (size_t)k * c.n, batch, 2 * (size_t)c.domain_len, fwd, st)) != cudaSuccess) return e;
    for (int k = 0; k < c.count64; k++)
        if ((e = run_ntt64_strided(pl->p64[k], d64 + (size_t)k * c.n, batch, (size_t)c.domain_len, fwd, st)) != cudaSuccess) return e;
    return cudaSuccess;
}
static cudaError_t product_pointwise(const cntt_product_plan* pl, int op, uint64_t* dst, const uint64_t* a, const uint64_t* b, size_t batch, cudaStream_t st)
{
    const ProductConsts& c = pl->c;
    cudaError_t e;
    const size_t off64 = (c.n / 2) * c.count32;
    for (int k = 0; k < c.count32; k++) {
        const size_t o = (size_t)k * c.n;
        if ((e = run_pw32_strided(pl->p32[k], op, reinterpret_cast<uint32_t*>(dst) + o, a ? reinterpret_cast<const uint32_t*>(a) + o : nullptr,
                                  b ? reinterpret_cast<const uint32_t*>(b) + o : nullptr, batch, 2 * (size_t)c.domain_len, st)) != cudaSuccess) return e;
    }
    for (int k = 0; k < c.count64; k++) {
        const size_t o = off64 + (size_t)k * c.n;
        if ((e = run_pw64_strided(pl->p64[k], op, dst + o, a ? a + o : nullptr, b ? b + o : nullptr, batch, (size_t)c.domain_len, st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// the hot shape -- two u32 primes of one arithmetic class, N <= 4096 -- runs one fused kernel per call (product_fused.hpp)
static bool product_fused_args(const cntt_product_plan* pl, ProductFusedArgs* a)
{
    const ProductConsts& c = pl->c;
    if (c.count32 != 2 || c.count64 != 0) return false;
    const cntt_prime32_plan *q0 = pl->p32[0], *q1 = pl->p32[1];
    if (!q0 || !q1 || q0->cls != q1->cls || (q0->cls != C32_L4 && q0->cls != C32_L2)) return false;
    if (!product_fused_supported(q0->cls == C32_L4 ? 0 : 1, q0->logn)) return false;
    a->cls = q0->cls == C32_L4 ? 0 : 1;
    a->logn = q0->logn;
    const cntt_prime32_plan* q[2] = {q0, q1};
    for (int j = 0; j < 2; j++) {
        a->tw_fwd[j] = q[j]->d_fwd; a->tw_inv[j] = q[j]->d_inv;
        a->last_fwd[j] = pl->d_pf_last[0][j] ? pl->d_pf_last[0][j] : q[j]->d_fwd_last;
        a->last_inv[j] = pl->d_pf_last[1][j] ? pl->d_pf_last[1][j] : q[j]->d_inv_last;
        a->mod[j] = q[j]->mod;
        a->head_fwd[j] = &q[j]->head_fwd; a->head_inv[j] = &q[j]->head_inv;
    }
    return true;
}

// fwd (product.rs:272-353): d_ntt[batch * domain_len] <- d_standard[batch * n]; mode 0 = FwdMode::Generic,
// 1 = FwdMode::Bounded(bound)
CNTT_API int cntt_product_fwd(const cntt_product_plan* pl, uint64_t* d_ntt, const uint64_t* d_standard, int mode, uint64_t bound, size_t batch, void* stream)
{
    if (!pl || ((!d_ntt || !d_standard) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_ntt, d_standard)) return CNTT_MISALIGNED;
    if (mode != PF_GENERIC && mode != PF_BOUNDED) return CNTT_UNSUPPORTED;
    if (batch == 0 || pl->c.count32 + pl->c.count64 == 0) return CNTT_OK;
    GUARD(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProductFusedArgs fa;
    if (product_fused_args(pl, &fa)) {
        CU(product_fused_fwd(pl->c, fa, d_ntt, d_standard, mode, bound, batch, st));
        return CNTT_OK;
    }
    const unsigned long long ncoef = (unsigned long long)batch * pl->n;
    k_product_reduce<<<(unsigned)((ncoef + 255) / 256), 256, 0, st>>>(pl->c, d_ntt, d_standard, mode, bound, ncoef);
    CU(cudaGetLastError());
    CU(product_planes_ntt(pl, d_ntt, batch, true, st));
    return CNTT_OK;
}
// inv (product.rs:355-880): d_standard[batch * n] <- (or += mod modulus) lift of the inverse transforms of d_ntt, which
// is clobbered like the reference's `ntt: &mut [u64]`; mode 0 = InvMode::Replace, 1 = InvMode::Accumulate
CNTT_API int cntt_product_inv(const cntt_product_plan* pl, uint64_t* d_standard, uint64_t* d_ntt, int mode, size_t batch, void* stream)
{
    if (!pl || ((!d_ntt || !d_standard) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_ntt, d_standard)) return CNTT_MISALIGNED;
    if (mode != PI_REPLACE && mode != PI_ACCUMULATE) return CNTT_UNSUPPORTED;
    if (batch == 0) return CNTT_OK;
    GUARD(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    ProductFusedArgs fa;
    if (product_fused_args(pl, &fa)) {
        CU(product_fused_inv(pl->c, fa, d_standard, d_ntt, mode, batch, st));
        return CNTT_OK;
    }
    CU(product_planes_ntt(pl, d_ntt, batch, false, st));
    const unsigned long long ncoef = (unsigned long long)batch * pl->n;
    k_product_crt<<<(unsigned)((ncoef + 255) / 256), 256, 0, st>>>(pl->c, d_standard, d_ntt, mode, ncoef);
    CU(cudaGetLastError());
    return CNTT_OK;
}
CNTT_API int cntt_product_mul_assign_normalize(const cntt_product_plan* pl, uint64_t* d_lhs, const uint64_t* d_rhs, size_t batch, void* stream)
{
    if (!pl || ((!d_lhs || !d_rhs) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_lhs, d_rhs)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(product_pointwise(pl, OP_MUL_ASSIGN_NORMALIZE, d_lhs, d_rhs, nullptr, batch, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_product_normalize(const cntt_product_plan* pl, uint64_t* d_values, size_t batch, void* stream)
{
    if (!pl || (!d_values && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_values)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(product_pointwise(pl, OP_NORMALIZE, d_values, nullptr, nullptr, batch, (cudaStream_t)stream));
    return CNTT_OK;
}
CNTT_API int cntt_product_mul_accumulate(const cntt_product_plan* pl, uint64_t* d_acc, const uint64_t* d_lhs, const uint64_t* d_rhs, size_t batch, void* stream)
{
    if (!pl || ((!d_acc || !d_lhs || !d_rhs) && batch)) return CNTT_NULL_POINTER;
    if (misaligned16(d_acc, d_lhs, d_rhs)) return CNTT_MISALIGNED;
    GUARD(pl->device);
    CU(product_pointwise(pl, OP_MUL_ACCUMULATE, d_acc, d_lhs, d_rhs, batch, (cudaStream_t)stream));
    return CNTT_OK;
}

// host-slice flavours (the reference's call shape, one or `batch` concatenated polynomials): staged through the plan's
// arena, synchronous.  h_bufs[i] holds words[i] u64 words; copy_back[i] says whether the reference mutates it.
static int product_host(const cntt_product_plan* pl, int nbuf, uint64_t* const* h_bufs, const size_t* words, const bool* copy_back,
                        const std::function<int(uint64_t* const*, cudaStream_t)>& body)
{
    auto* mpl = const_cast<cntt_product_plan*>(pl);
    DeviceGuard guard(pl->device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    std::lock_guard<std::mutex> lk(mpl->stg.mu);
    size_t total = 0;
    for (int i = 0; i < nbuf; i++) total += (words[i] + 1) & ~(size_t)1; // keep every buffer 16-byte aligned
    CU(mpl->stg.ensure(std::max<size_t>(total, 2) * sizeof(uint64_t)));
    uint64_t* d[4];
    size_t off = 0;
    for (int i = 0; i < nbuf; i++) {
        d[i] = static_cast<uint64_t*>(mpl->stg.buf) + off;
        off += (words[i] + 1) & ~(size_t)1;
        if (words[i]) CU(cudaMemcpyAsync(d[i], h_bufs[i], words[i] * sizeof(uint64_t), cudaMemcpyHostToDevice, mpl->stg.stream));
    }
    const int st = body(d, mpl->stg.stream);
    if (st != CNTT_OK) return st;
    for (int i = 0; i < nbuf; i++)
        if (copy_back[i] && words[i]) CU(cudaMemcpyAsync(h_bufs[i], d[i], words[i] * sizeof(uint64_t), cudaMemcpyDeviceToHost, mpl->stg.stream));
    CU(cudaStreamSynchronize(mpl->stg.stream));
    return CNTT_OK;
}
CNTT_API int cntt_product_fwd_host(const cntt_product_plan* pl, uint64_t* h_ntt, const uint64_t* h_standard, size_t ntt_len, size_t standard_len,
                                   int mode, uint64_t bound, size_t batch)
{
    if (!pl || ((!h_ntt && ntt_len) || (!h_standard && standard_len))) return CNTT_NULL_POINTER;
    if (standard_len != pl->n * batch || ntt_len != (size_t)pl->c.domain_len * batch) return CNTT_LENGTH_MISMATCH; // assert_eq!, product.rs:278-279
    uint64_t* bufs[2] = {h_ntt, const_cast<uint64_t*>(h_standard)};
    const size_t words[2] = {ntt_len, standard_len};
    const bool back[2] = {true, false};
    return product_host(pl, 2, bufs, words, back, [&](uint64_t* const* d, cudaStream_t st) { return cntt_product_fwd(pl, d[0], d[1], mode, bound, batch, st); });
}
CNTT_API int cntt_product_inv_host(const cntt_product_plan* pl, uint64_t* h_standard, uint64_t* h_ntt, size_t standard_len, size_t ntt_len, int mode,
                                   size_t batch)
{
    if (!pl || ((!h_ntt && ntt_len) || (!h_standard && standard_len))) return CNTT_NULL_POINTER;
    if (standard_len != pl->n * batch || ntt_len != (size_t)pl->c.domain_len * batch) return CNTT_LENGTH_MISMATCH; // product.rs:357-358
    uint64_t* bufs[2] = {h_standard, h_ntt};
    const size_t words[2] = {standard_len, ntt_len};
    const bool back[2] = {true, true}; // the reference's inv leaves the inverse transforms in `ntt`
    return product_host(pl, 2, bufs, words, back, [&](uint64_t* const* d, cudaStream_t st) { return cntt_product_inv(pl, d[0], d[1], mode, batch, st); });
}
CNTT_API int cntt_product_mul_assign_normalize_host(const cntt_product_plan* pl, uint64_t* h_lhs, const uint64_t* h_rhs, size_t len, size_t batch)
{
    if (!pl || ((!h_lhs || !h_rhs) && len)) return CNTT_NULL_POINTER;
    if (len != (size_t)pl->c.domain_len * batch) return CNTT_LENGTH_MISMATCH; // product.rs:887-888
    uint64_t* bufs[2] = {h_lhs, const_cast<uint64_t*>(h_rhs)};
    const size_t words[2] = {len, len};
    const bool back[2] = {true, false};
    return product_host(pl, 2, bufs, words, back, [&](uint64_t* const* d, cudaStream_t st) { return cntt_product_mul_assign_normalize(pl, d[0], d[1], batch, st); });
}
CNTT_API int cntt_product_normalize_host(const cntt_product_plan* pl, uint64_t* h_values, size_t len, size_t batch)
{
    if (!pl || (!h_values && len)) return CNTT_NULL_POINTER;
    if (len != (size_t)pl->c.domain_len * batch) return CNTT_LENGTH_MISMATCH; // product.rs:920
    uint64_t* bufs[1] = {h_values};
    const size_t words[1] = {len};
    const bool back[1] = {true};
    return product_host(pl, 1, bufs, words, back, [&](uint64_t* const* d, cudaStream_t st) { return cntt_product_normalize(pl, d[0], batch, st); });
}
CNTT_API int cntt_product_mul_accumulate_host(const cntt_product_plan* pl, uint64_t* h_acc, const uint64_t* h_lhs, const uint64_t* h_rhs, size_t len,
                                              size_t batch)
{
    if (!pl || ((!h_acc || !h_lhs || !h_rhs) && len)) return CNTT_NULL_POINTER;
    if (len != (size_t)pl->c.domain_len * batch) return CNTT_LENGTH_MISMATCH; // product.rs:938-939
    uint64_t* bufs[3] = {h_acc, const_cast<uint64_t*>(h_lhs), const_cast<uint64_t*>(h_rhs)};
    const size_t words[3] = {len, len, len};
    const bool back[3] = {true, false, false};
    return product_host(pl, 3, bufs, words, back, [&](uint64_t* const* d, cudaStream_t st) { return cntt_product_mul_accumulate(pl, d[0], d[1], d[2], batch, st); });
}
