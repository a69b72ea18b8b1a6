// arith.cuh -- modular-arithmetic policies for the sm_100a NTT kernels.
//
// One policy per "bit class" of the reference's dispatch (prime32.rs:713-754, prime64.rs:812-864).
// Only the *final canonical values* are observable (every reference path returns residues in
// [0,p)), so each policy is free to pick its own lazy ranges; they are documented per policy.
//
// A policy provides
//   W            word type (uint32_t / uint64_t)
//   Tw           twiddle record as stored in the device heap table
//   Mod          per-plan constants passed by value in kernel params
//   fwd_bf       Cooley-Tukey butterfly  (z0,z1) -> (z0 + w z1, z0 - w z1)   [lazy]
//   inv_bf       Gentleman-Sande butterfly (z0,z1) -> (z0 + z1, (z0 - z1) w) [lazy]
//   canon_fwd / canon_inv   map the lazy range after the last fwd / inv stage to [0,p)
//   mul_norm     lhs*rhs*n^-1, norm: v*n^-1, mul_acc: acc+lhs*rhs   -- the reference's own
//                scalar formulas (prime32.rs:383-408,477-486,575-598; prime64.rs:534-584,690-699)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cntt {

typedef unsigned __int128 u128_t;

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------
// 32-bit constants shared by all u32 policies
// ------------------------------------------------------------------------------------------
struct Mod32 {
    uint32_t p, two_p, neg_p;
    uint32_t n_inv, n_inv_shoup;  // N^-1 mod p and floor(N^-1 * 2^32 / p)
    uint32_t p_barrett, big_q_m1; // prime32.rs:667-671
};

// Shoup multiply: a * w mod p in [0, 2p) for any 32-bit a (p < 2^31).
__device__ __forceinline__ uint32_t shoup32(uint32_t a, uint2 t, const Mod32& m)
{
    uint32_t q = __umulhi(a, t.y);
    return a * t.x + q * m.neg_p;
}

// ---- p < 2^30 : Harvey lazy butterflies, values in [0, 4p) between forward stages and
//      [0, 2p) between inverse stages (same ranges as prime32/less_than_30bit.rs:115-129,265-282)
struct A32L4 {
    typedef uint32_t W;
    typedef uint2 Tw;
    typedef Mod32 Mod;
    static constexpr bool kShoupTable = true;

    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W c = umin32(z0, z0 - m.two_p);
        W x = shoup32(z1, t, m);
        z0 = c + x;
        z1 = c - x + m.two_p;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W s = z0 + z1;
        W d = z0 - z1 + m.two_p;
        z0 = umin32(s, s - m.two_p);
        z1 = shoup32(d, t, m);
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod& m)
    {
        x = umin32(x, x - m.two_p);
        return umin32(x, x - m.p);
    }
    static __device__ __forceinline__ W canon_inv(W x, const Mod& m) { return umin32(x, x - m.p); }
    // any 32-bit value is a legal forward input for the z1 role; the z0 role needs [0,4p)
    static __device__ __forceinline__ W mul_norm(W a, W b, const Mod& m)
    {
        uint64_t d = (uint64_t)a * b;
        uint32_t c1 = (uint32_t)(d >> m.big_q_m1);
        uint32_t c3 = __umulhi(c1, m.p_barrett);
        uint32_t prod = (uint32_t)d - m.p * c3;
        uint32_t q = __umulhi(prod, m.n_inv_shoup);
        uint32_t t = prod * m.n_inv - q * m.p;
        return umin32(t, t - m.p);
    }
    static __device__ __forceinline__ W norm(W v, const Mod& m)
    {
        uint32_t q = __umulhi(v, m.n_inv_shoup);
        uint32_t t = v * m.n_inv - q * m.p;
        return umin32(t, t - m.p);
    }
    static __device__ __forceinline__ W mul_acc(W acc, W a, W b, const Mod& m)
    {
        uint64_t d = (uint64_t)a * b;
        uint32_t c1 = (uint32_t)(d >> m.big_q_m1);
        uint32_t c3 = __umulhi(c1, m.p_barrett);
        uint32_t prod = (uint32_t)d - m.p * c3;
        prod = umin32(prod, prod - m.p);
        uint32_t s = prod + acc;
        return umin32(s, s - m.p);
    }
};

// ---- 2^30 <= p < 2^31 : values in [0, 2p) between forward stages, [0, p) between inverse stages
//      (prime32/less_than_31bit.rs:117-133, 214-234)
struct A32L2 : A32L4 {
    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W c = umin32(z0, z0 - m.p);
        W x = shoup32(z1, t, m);
        x = umin32(x, x - m.p);
        z0 = c + x;
        z1 = c - x + m.p;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W s = z0 + z1;
        W d = z0 - z1 + m.p;
        z0 = umin32(s, s - m.p);
        W y = shoup32(d, t, m);
        z1 = umin32(y, y - m.p);
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod& m) { return umin32(x, x - m.p); }
    static __device__ __forceinline__ W canon_inv(W x, const Mod&) { return x; }
};

// ---- p >= 2^31 : fully reduced values.  The reference has no Shoup table for this class
//      (prime32.rs:645-652) and divides with Div32; here the device table still carries
//      floor(w 2^32 / p) and the product is finished in 64 bits (internal choice, same results).
struct A32G {
    typedef uint32_t W;
    typedef uint2 Tw;
    typedef Mod32 Mod;
    static constexpr bool kShoupTable = true;

    static __device__ __forceinline__ W mulw(W a, Tw t, const Mod& m) // a < p  ->  a*w mod p
    {
        uint32_t q = __umulhi(a, t.y);
        uint64_t r = (uint64_t)a * t.x - (uint64_t)q * m.p; // in [0, 2p)
        if (r >= m.p) r -= m.p;
        return (W)r;
    }
    static __device__ __forceinline__ W add(W a, W b, const Mod& m)
    {
        W nb = m.p - b;
        return a >= nb ? a - nb : a + b;
    }
    static __device__ __forceinline__ W sub(W a, W b, const Mod& m) { return a >= b ? a - b : a + (m.p - b); }
    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W x = mulw(z1, t, m);
        W a = add(z0, x, m), b = sub(z0, x, m);
        z0 = a; z1 = b;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W a = add(z0, z1, m), b = mulw(sub(z0, z1, m), t, m);
        z0 = a; z1 = b;
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod&) { return x; }
    static __device__ __forceinline__ W canon_inv(W x, const Mod&) { return x; }
    static __device__ __forceinline__ W mulmod(W a, W b, const Mod& m) { return (W)(((uint64_t)a * b) % m.p); }
    // prime32.rs:853-863, 893-901, 919-926
    static __device__ __forceinline__ W mul_norm(W a, W b, const Mod& m) { return mulmod(mulmod(a, b, m), m.n_inv, m); }
    static __device__ __forceinline__ W norm(W v, const Mod& m) { return mulmod(v, m.n_inv, m); }
    static __device__ __forceinline__ W mul_acc(W acc, W a, W b, const Mod& m) { return add(acc, mulmod(a, b, m), m); }
};

// ------------------------------------------------------------------------------------------
// 64-bit
// ------------------------------------------------------------------------------------------
struct Mod64 {
    uint64_t p, two_p, neg_p;
    uint64_t n_inv, n_inv_shoup;
    uint64_t p_barrett;
    uint32_t big_q_m1;
    uint32_t shift_head; // Solinas: the plan's first heap entries are the powers of two the shift butterflies assume (checked at plan time)
};

__device__ __forceinline__ uint64_t shoup64(uint64_t a, ulonglong2 t, const Mod64& m)
{
    uint64_t q = __umul64hi(a, t.y);
    return a * t.x + q * m.neg_p;
}

// ---- p < 2^62 (prime64/less_than_62bit.rs:117-131, 271-288)
struct A64L4 {
    typedef uint64_t W;
    typedef ulonglong2 Tw;
    typedef Mod64 Mod;
    static constexpr bool kShoupTable = true;

    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W c = umin64(z0, z0 - m.two_p);
        W x = shoup64(z1, t, m);
        z0 = c + x;
        z1 = c - x + m.two_p;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W s = z0 + z1;
        W d = z0 - z1 + m.two_p;
        z0 = umin64(s, s - m.two_p);
        z1 = shoup64(d, t, m);
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod& m)
    {
        x = umin64(x, x - m.two_p);
        return umin64(x, x - m.p);
    }
    static __device__ __forceinline__ W canon_inv(W x, const Mod& m) { return umin64(x, x - m.p); }
    // prime64.rs:534-559
    static __device__ __forceinline__ W barrett(W a, W b, const Mod& m)
    {
        uint64_t lo = a * b, hi = __umul64hi(a, b);
        uint64_t c1 = m.big_q_m1 == 0 ? lo : (m.big_q_m1 >= 64 ? hi >> (m.big_q_m1 - 64) : (lo >> m.big_q_m1) | (hi << (64 - m.big_q_m1)));
        uint64_t c3 = __umul64hi(c1, m.p_barrett);
        return lo - m.p * c3;
    }
    static __device__ __forceinline__ W mul_norm(W a, W b, const Mod& m)
    {
        uint64_t prod = barrett(a, b, m);
        uint64_t q = __umul64hi(prod, m.n_inv_shoup);
        uint64_t t = prod * m.n_inv - q * m.p;
        return umin64(t, t - m.p);
    }
    static __device__ __forceinline__ W norm(W v, const Mod& m)
    {
        uint64_t q = __umul64hi(v, m.n_inv_shoup);
        uint64_t t = v * m.n_inv - q * m.p;
        return umin64(t, t - m.p);
    }
    static __device__ __forceinline__ W mul_acc(W acc, W a, W b, const Mod& m)
    {
        uint64_t prod = barrett(a, b, m);
        prod = umin64(prod, prod - m.p);
        uint64_t s = prod + acc;
        return umin64(s, s - m.p);
    }
};

// ---- 2^62 <= p < 2^63 (prime64/less_than_63bit.rs:117-133, 214-234)
struct A64L2 : A64L4 {
    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W c = umin64(z0, z0 - m.p);
        W x = shoup64(z1, t, m);
        x = umin64(x, x - m.p);
        z0 = c + x;
        z1 = c - x + m.p;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W s = z0 + z1;
        W d = z0 - z1 + m.p;
        z0 = umin64(s, s - m.p);
        W y = shoup64(d, t, m);
        z1 = umin64(y, y - m.p);
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod& m) { return umin64(x, x - m.p); }
    static __device__ __forceinline__ W canon_inv(W x, const Mod&) { return x; }
};

// ---- p = 2^64 - 2^32 + 1 (Solinas / Goldilocks), prime64/generic_solinas.rs:77-129.
//      2^64 = 2^32 - 1 =: EPS and 2^96 = -1 (mod p), so a 128-bit product (c3,c2,c1,c0) reduces to
//      (c1:c0) - c3 + c2*EPS with end-around corrections.  The class is bound by the integer pipes
//      (ncu, r01: ALU 83 %, FMA 19 % busy with an all-ALU formulation), so the arithmetic is written on
//      32-bit limbs in PTX and split across both pipes on purpose:
//        product   one mad/madc chain: ptxas keeps the carries inside IMAD.WIDE / IMAD.X / IMAD.HI (FMA pipe)
//        reduce    (c1:c0) - c3 on the ALU (5), then c2*EPS + X as ONE IMAD.WIDE with carry-out (FMA), and the
//                  end-around correction merged with the canonicalisation: "carry or r >= p" -> r += EPS mod 2^64 (5)
//        add 6 / sub 5 ALU instructions, valid when the second operand is <= p (any 64-bit first operand)
//      => 19 ALU + ~13 FMA-pipe instructions per butterfly (was 30 + 8), both pipes ~40 cycles per warp-butterfly.
//      (carry chains are never mixed: ptxas hands `subc` the raw carry predicate after an add.cc)
//      Between forward stages values are arbitrary 64-bit representatives ("lazy"); only products are
//      canonical.  Between inverse stages all values are canonical.
#ifndef CNTT_MUL128
#define CNTT_MUL128 0
#endif
struct A64S {
    typedef uint64_t W;
    typedef uint64_t Tw; // no Shoup companion
    typedef Mod64 Mod;
    static constexpr bool kShoupTable = false;
    static constexpr uint64_t P = 0xFFFFFFFF00000001ull;
    static constexpr uint64_t EPS = 0x00000000FFFFFFFFull; // 2^64 - P

    static __device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

    // x >= p  ->  x - p   (x - p = (0 : x0 - 1) because x1 must be 0xFFFFFFFF)
    static __device__ __forceinline__ W canon(W x)
    {
        uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
        asm("{\n\t"
            ".reg .pred q;\n\t"
            "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
            "setp.ne.and.u32 q, %0, 0, q;\n\t"
            "@q add.u32      %0, %0, 0xFFFFFFFF;\n\t"
            "@q mov.u32      %1, 0;\n\t"
            "}"
            : "+r"(x0), "+r"(x1));   // 5 SASS instructions; the C form of the same test compiled to 8
        return pack(x0, x1);
    }
    static __device__ __forceinline__ W mul(W a, W b) // any a, b; result canonical
    {
        uint32_t c0, c1, c2, c3, r0, r1;
#if CNTT_MUL128
        {   // the compiler's own 64 x 64 -> 128 product: 3 IMAD.WIDE + IMAD.WIDE.X + 2 adds (two instructions fewer than the chain below)
            const u128_t w = (u128_t)a * b;
            c0 = (uint32_t)w; c1 = (uint32_t)(w >> 32); c2 = (uint32_t)(w >> 64); c3 = (uint32_t)(w >> 96);
        }
#else
        asm("{\n\t"
            "mul.lo.u32      %0, %4, %6;\n\t"
            "mul.hi.u32      %1, %4, %6;\n\t"
            "mad.lo.cc.u32   %1, %4, %7, %1;\n\t"
            "madc.hi.cc.u32  %2, %4, %7, 0;\n\t"
            "addc.u32        %3, 0, 0;\n\t"
            "mad.lo.cc.u32   %1, %5, %6, %1;\n\t"
            "madc.hi.cc.u32  %2, %5, %6, %2;\n\t"
            "addc.u32        %3, %3, 0;\n\t"
            "mad.lo.cc.u32   %2, %5, %7, %2;\n\t"
            "madc.hi.u32     %3, %5, %7, %3;\n\t"
            "}"
            : "=&r"(c0), "=&r"(c1), "=&r"(c2), "=&r"(c3)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
#endif
        // X = (c1:c0) - c3 (borrow: the wrap added 2^64 = EPS too much);  r = X + c2*EPS mod 2^64, carry m.
        // m = 1: true value r + 2^64 = r + EPS, and r <= 2^64 - 2^33 so the sum does not wrap and is < p.
        // m = 0 and r >= p (r1 all ones, r0 != 0): r + EPS wraps to r - p.  Either way: r += EPS mod 2^64.
        asm("{\n\t"
            ".reg .u32 m;\n\t"
            ".reg .pred q;\n\t"
            "sub.cc.u32      %0, %2, %5;\n\t"
            "subc.cc.u32     %1, %3, 0;\n\t"
            "subc.u32        m, 0, 0;\n\t"
            "sub.cc.u32      %0, %0, m;\n\t"
            "subc.u32        %1, %1, 0;\n\t"
            "mad.lo.cc.u32   %0, %4, 0xFFFFFFFF, %0;\n\t"
            "madc.hi.cc.u32  %1, %4, 0xFFFFFFFF, %1;\n\t"
            "addc.u32        m, 0, 0;\n\t"
            "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
            "setp.ne.and.u32 q, %0, 0, q;\n\t"
            "setp.ne.or.u32  q, m, 0, q;\n\t"
            "@q add.cc.u32   %0, %0, 0xFFFFFFFF;\n\t"
            "@q addc.u32     %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"(c0), "r"(c1), "r"(c2), "r"(c3));
        return pack(r0, r1);
    }
    // z + x, z - x for any 64-bit z and x <= p; result is some 64-bit representative.
    // carry: z + x - 2^64 <= p - 1, so adding EPS cannot overflow again; borrow: z - x + 2^64 >= EPS.
    static __device__ __forceinline__ W add_lazy(W z, W x)
    {
        uint32_t r0, r1;
        asm("{\n\t"
            ".reg .u32 m;\n\t"
            "add.cc.u32   %0, %2, %4;\n\t"
            "addc.cc.u32  %1, %3, %5;\n\t"
            "addc.u32     m, 0, 0;\n\t"      // carry
            "sub.cc.u32   %0, %0, m;\n\t"    // + EPS * carry = + (m << 32) - m
            "subc.u32     %1, %1, 0;\n\t"
            "add.u32      %1, %1, m;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"((uint32_t)z), "r"((uint32_t)(z >> 32)), "r"((uint32_t)x), "r"((uint32_t)(x >> 32)));
        return pack(r0, r1);
    }
    static __device__ __forceinline__ W sub_lazy(W z, W x)
    {
        uint32_t r0, r1;
        asm("{\n\t"
            ".reg .u32 m;\n\t"
            "sub.cc.u32   %0, %2, %4;\n\t"
            "subc.cc.u32  %1, %3, %5;\n\t"
            "subc.u32     m, 0, 0;\n\t"
            "sub.cc.u32   %0, %0, m;\n\t"
            "subc.u32     %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"((uint32_t)z), "r"((uint32_t)(z >> 32)), "r"((uint32_t)x), "r"((uint32_t)(x >> 32)));
        return pack(r0, r1);
    }
    // canonical a + b for canonical a, b: "carry or sum >= p" -> sum += EPS mod 2^64 (same argument as in mul)
    static __device__ __forceinline__ W add(W a, W b)
    {
        uint32_t r0, r1;
        asm("{\n\t"
            ".reg .u32 k;\n\t"
            ".reg .pred q;\n\t"
            "add.cc.u32      %0, %2, %4;\n\t"
            "addc.cc.u32     %1, %3, %5;\n\t"
            "addc.u32        k, 0, 0;\n\t"
            "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
            "setp.ne.and.u32 q, %0, 0, q;\n\t"
            "setp.ne.or.u32  q, k, 0, q;\n\t"
            "@q add.cc.u32   %0, %0, 0xFFFFFFFF;\n\t"
            "@q addc.u32     %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
        return pack(r0, r1);
    }
    // ---- power-of-two twiddles.  2 has order 192 modulo p (2^96 = -1), so every 64-th root of unity is +-2^k, k < 96: the twiddles
    //      of the first five levels of ANY transform (level j uses primitive 2^(j+2)-th roots).  The reference's root search starts
    //      from -1 and takes square roots (roots.rs:68-91), so those entries do not depend on N: tw[h] = 2^kShiftExp[h] for the heap
    //      nodes h < 32 (levels 0..4; checked against the real table when a plan is built).  x * 2^(32a+b) with x = (x1:x0), y = x << b
    //      = (y2:y1:y0) and phi = 2^32 (phi^2 = phi - 1, phi^3 = -1):
    //        a = 0:  (y1:y0) + y2 EPS  =  (y1:y0) - [(~y2) : (y2 + 1)]  (+ p on borrow)     [(~y2):(y2+1) = p - y2 EPS]
    //        a = 1:  (y0:0) - y2 + y1 EPS                                                    the generic 128 -> 64 reduction
    //        a = 2:  y0 EPS - (y2:y1)                                                         (+ p on borrow)
    //      each result is <= p for ANY 64-bit x (what add_lazy / sub_lazy need).  B200, registers only (tools/ubench/gold_bf.cu):
    //      3.24 against 2.40 butterflies/clk/SM forward, 2.97 against 2.30 inverse -- the multiply goes, the 64-bit modular add / sub on
    //      the ALU stays (profiles/r02_experiments.txt).
    template <int B> static __device__ __forceinline__ W shl_a0(W x)
    {
        const uint32_t y2 = (uint32_t)(x >> (64 - B));
        return sub_lazy(x << B, ((uint64_t)(~y2) << 32) | (uint64_t)(y2 + 1u));
    }
    template <int B> static __device__ __forceinline__ W shl_a2(W x)
    {
        const uint32_t y0 = (uint32_t)x << B;
        return sub_lazy((uint64_t)y0 * EPS, x >> (32 - B));
    }
    template <int B> static __device__ __forceinline__ W shl_a1(W x)
    {
        const uint32_t y0 = (uint32_t)x << B, y1 = (uint32_t)(x >> (32 - B)), y2 = (uint32_t)(x >> (64 - B));
        uint32_t r0, r1;
        asm("{\n\t"
            ".reg .u32 m;\n\t"
            ".reg .pred q;\n\t"
            "sub.cc.u32      %0, 0, %4;\n\t"
            "subc.cc.u32     %1, %2, 0;\n\t"
            "subc.u32        m, 0, 0;\n\t"
            "sub.cc.u32      %0, %0, m;\n\t"
            "subc.u32        %1, %1, 0;\n\t"
            "mad.lo.cc.u32   %0, %3, 0xFFFFFFFF, %0;\n\t"
            "madc.hi.cc.u32  %1, %3, 0xFFFFFFFF, %1;\n\t"
            "addc.u32        m, 0, 0;\n\t"
            "setp.eq.u32     q, %1, 0xFFFFFFFF;\n\t"
            "setp.ne.and.u32 q, %0, 0, q;\n\t"
            "setp.ne.or.u32  q, m, 0, q;\n\t"
            "@q add.cc.u32   %0, %0, 0xFFFFFFFF;\n\t"
            "@q addc.u32     %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1) : "r"(y0), "r"(y1), "r"(y2));
        return pack(r0, r1);
    }
    template <int K> static __device__ __forceinline__ W shl_mod(W x) // |x * 2^K|: the sign of 2^K (K >= 96) is left to the caller
    {
        constexpr int k = K % 96, a = k / 32, b = k % 32;
        static_assert(b != 0 || k == 0, "limb-aligned shifts do not occur among the 64-th roots of unity");
        if constexpr (k == 0) return x;
        else if constexpr (a == 0) return shl_a0<b>(x);
        else if constexpr (a == 1) return shl_a1<b>(x);
        else return shl_a2<b>(x);
    }
    template <int K> static __device__ __forceinline__ void fwd_bf_shift(W& z0, W& z1) // fwd_bf with twiddle 2^K
    {
        const W t = shl_mod<K>(z1);
        const W a = add_lazy(z0, t), b = sub_lazy(z0, t);
        if constexpr ((K % 192) >= 96) { z0 = b; z1 = a; } else { z0 = a; z1 = b; }
    }
    template <int K> static __device__ __forceinline__ void inv_bf_shift(W& z0, W& z1) // inv_bf with twiddle 2^K, canonical in / out
    {
        const W a = add(z0, z1);
        const W d = (K % 192) >= 96 ? sub_lazy(z1, z0) : sub_lazy(z0, z1);
        z0 = a; z1 = shl_mod<K>(d); // every shl_a* result is < p (see above), d is canonical
    }
    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod&)
    {
        const W x = mul(z1, t);
        const W a = add_lazy(z0, x), b = sub_lazy(z0, x);
        z0 = a; z1 = b;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod&) // canonical in/out
    {
        const W a = add(z0, z1);
        const W d = sub_lazy(z0, z1);
        z0 = a;
        z1 = mul(d, t);
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod&) { return canon(x); }
    static __device__ __forceinline__ W canon_inv(W x, const Mod&) { return x; }
    // prime64.rs:1013-1021, 1068-1073, 1116-1120
    static __device__ __forceinline__ W mul_norm(W a, W b, const Mod& m) { return mul(mul(a, b), m.n_inv); }
    static __device__ __forceinline__ W norm(W v, const Mod& m) { return mul(v, m.n_inv); }
    static __device__ __forceinline__ W mul_acc(W acc, W a, W b, const Mod&) { return add(canon(acc), mul(a, b)); }
};

// ---- any other p >= 2^63 (prime64/generic_solinas.rs:42-75, Div64 remainder).  Device table
//      carries floor(w 2^64 / p); the product is finished in 128 bits.
struct A64G {
    typedef uint64_t W;
    typedef ulonglong2 Tw;
    typedef Mod64 Mod;
    static constexpr bool kShoupTable = true;

    static __device__ __forceinline__ W mulw(W a, Tw t, const Mod& m)
    {
        uint64_t q = __umul64hi(a, t.y);
        u128_t r = (u128_t)a * t.x - (u128_t)q * m.p; // in [0, 2p)
        if (r >= m.p) r -= m.p;
        return (W)r;
    }
    static __device__ __forceinline__ W add(W a, W b, const Mod& m)
    {
        W nb = m.p - b;
        return a >= nb ? a - nb : a + b;
    }
    static __device__ __forceinline__ W sub(W a, W b, const Mod& m) { return a >= b ? a - b : a + (m.p - b); }
    static __device__ __forceinline__ void fwd_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W x = mulw(z1, t, m);
        W a = add(z0, x, m), b = sub(z0, x, m);
        z0 = a; z1 = b;
    }
    static __device__ __forceinline__ void inv_bf(W& z0, W& z1, Tw t, const Mod& m)
    {
        W a = add(z0, z1, m), b = mulw(sub(z0, z1, m), t, m);
        z0 = a; z1 = b;
    }
    static __device__ __forceinline__ W canon_fwd(W x, const Mod&) { return x; }
    static __device__ __forceinline__ W canon_inv(W x, const Mod&) { return x; }
    static __device__ __forceinline__ W mulmod(W a, W b, const Mod& m) { return (W)(((u128_t)a * b) % m.p); }
    static __device__ __forceinline__ W mul_norm(W a, W b, const Mod& m) { return mulmod(mulmod(a, b, m), m.n_inv, m); }
    static __device__ __forceinline__ W norm(W v, const Mod& m) { return mulmod(v, m.n_inv, m); }
    static __device__ __forceinline__ W mul_acc(W acc, W a, W b, const Mod& m) { return add(acc, mulmod(a, b, m), m); }
};

// log2 of the Solinas plan's forward twiddles at heap nodes 1..31 (levels 0..4, all the 64-th roots of unity the transform uses):
// tw[h] = 2^shift_exp(h) mod p, inverse table 2^(192 - shift_exp(h)).  Derived from the reference's root chain (-1, sqrt, sqrt, ...)
// and verified per plan (capi.cu).
constexpr int kShiftNodes = 32;
__host__ __device__ constexpr int shift_exp(int h)
{
    constexpr int e[kShiftNodes] = {0,  48, 120, 168, 156, 12, 84,  132, 78,  126, 6,  54, 42, 90,  162, 18,
                                    39, 87, 159, 15,  3,   51, 123, 171, 117, 165, 45, 93, 81, 129, 9,   57};
    return e[h & (kShiftNodes - 1)];
}
template <class A> struct ShiftHead { static constexpr bool value = false; };
#ifndef CNTT_SHIFT_HEAD
#define CNTT_SHIFT_HEAD 1
#endif
template <> struct ShiftHead<A64S> { static constexpr bool value = CNTT_SHIFT_HEAD != 0; };
#ifndef CNTT_FIRST_FULL_64S
#define CNTT_FIRST_FULL_64S 1 // Solinas: full first pass (four shift levels), short last pass -- ntt_engine.cuh, Geo
#endif

} // namespace cntt
