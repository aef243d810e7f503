// fe25519.cuh -- GF(2^255-19) on 8 saturated 32-bit limbs (radix 2^32) for sm_100a.
//
// Values are kept "loosely reduced": any representative in [0, 2^256).  2^256 = 38 (mod p), so a
// 512-bit product folds as lo + 38*hi; its last fold starts at bit 255, so products are below 2^255 + 2^11.  The 8x8 limb product runs on the integer multiply pipe as
// IMAD.WIDE.U32(.X) chains: mad.lo.cc/madc.hi.cc pairs of one (a_j, b_i) fuse into a single
// 32x32+64 -> 64 multiply-add with carry-in/out predicates (checked with cuobjdump -sass), and the
// even/odd column split keeps every chain free of overlapping 64-bit slots.
//
// Replaces curve25519-dalek's FieldElement (un-vendored dependency of /root/reference, Cargo.lock:267):
// used by every point operation behind src/lib.rs:621-663 and :781-869.
//
// The same source compiles for the host (plain C++ paths, no PTX) ONLY for tests/hostsim, which
// unit-tests the arithmetic logic without a GPU; the product library never uses the host path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ACT_FN __device__ __forceinline__
#define ACT_NOINLINE __device__ __noinline__
#define ACT_CONST __device__ __constant__ const
#else
#define ACT_FN static inline
#define ACT_NOINLINE static
#define ACT_CONST static const
#endif

#if defined(__CUDA_ARCH__)
#define ACT_PTX 1
#define ACT_UNROLL _Pragma("unroll")
#define ACT_NOUNROLL _Pragma("unroll 1")
#else
#define ACT_PTX 0
#define ACT_UNROLL
#define ACT_NOUNROLL
#endif

#ifndef ACT_REDUCE_ALT
#define ACT_REDUCE_ALT 1
#endif
// ACT_TIGHT: products leave fe_fold9 "tight" (< 2^255 + 2^11: bit 255 is folded too, at no extra instruction), which lets
// the sum or the double of two PRODUCTS skip the second carry pass (fe_add_tt / fe_dbl_tt below).
#ifndef ACT_TIGHT
#define ACT_TIGHT 1
#endif
// the host build (tests/hostsim) checks the precondition of the tight operations on every call
#if !defined(__CUDA_ARCH__) && defined(ACT_HOSTSIM_CHECKS)
#include <assert.h>
#define ACT_ASSERT_TIGHT(x) assert(fe_is_tight_(x))
#else
#define ACT_ASSERT_TIGHT(x) ((void)0)
#endif

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

struct fe { u32 v[8]; };

ACT_CONST u32 FE_D_[8] = {0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu};
ACT_CONST u32 FE_D2_[8] = {0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu};
ACT_CONST u32 FE_SQRT_M1_[8] = {0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u};
ACT_CONST u32 FE_SQRT_AD_MINUS_ONE_[8] = {0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu, 0x31f5d1fdu, 0x0f3cfcc9u, 0x2b8348acu, 0x376931bfu};
ACT_CONST u32 FE_INVSQRT_A_MINUS_D_[8] = {0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u};
ACT_CONST u32 FE_ONE_MINUS_D_SQ_[8] = {0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u, 0xbe70dfe4u, 0x9994abddu, 0xb2b3e0d7u, 0x029072a8u};
ACT_CONST u32 FE_D_MINUS_ONE_SQ_[8] = {0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu, 0x529b4eebu, 0x4cdcd32fu, 0xf66c2241u, 0x5968b37au};
ACT_CONST u32 FE_BX_[8] = {0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u, 0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u};
ACT_CONST u32 FE_BY_[8] = {0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u};
ACT_CONST u32 FE_BT_[8] = {0xa5b7dda3u, 0x6dde8ab3u, 0x775152f5u, 0x20f09f80u, 0x64abe37du, 0x66ea4e8eu, 0xd78b7665u, 0x67875f0fu};

ACT_FN fe fe_const(const u32* c) {
    fe r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}
#define FE_D fe_const(FE_D_)
#define FE_D2 fe_const(FE_D2_)
#define FE_SQRT_M1 fe_const(FE_SQRT_M1_)
#define FE_SQRT_AD_MINUS_ONE fe_const(FE_SQRT_AD_MINUS_ONE_)
#define FE_INVSQRT_A_MINUS_D fe_const(FE_INVSQRT_A_MINUS_D_)
#define FE_ONE_MINUS_D_SQ fe_const(FE_ONE_MINUS_D_SQ_)
#define FE_D_MINUS_ONE_SQ fe_const(FE_D_MINUS_ONE_SQ_)

ACT_FN fe fe_zero() {
    fe r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
ACT_FN fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }

// ---- add / sub -------------------------------------------------------------------------------
// r = a + b (mod p), loosely reduced.  Carry out of 2^256 folds back as +38 (twice: the second fold
// cannot propagate because the wrapped value is then < 2^13).
ACT_FN fe fe_add(const fe& a, const fe& b) {
    fe r;
#if ACT_PTX
    u32 c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    u32 f = c * 38u;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c)
        : "r"(f));
    r.v[0] += c * 38u;
#else
    u64 c = 0;
    for (int i = 0; i < 8; i++) { c += (u64)a.v[i] + b.v[i]; r.v[i] = (u32)c; c >>= 32; }
    c *= 38;
    for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (u32)c; c >>= 32; }
    r.v[0] += (u32)c * 38u;
#endif
    return r;
}

// r = a - b (mod p).  A borrow out of 2^256 folds back as -38 (twice, same argument as fe_add).
ACT_FN fe fe_sub(const fe& a, const fe& b) {
    fe r;
#if ACT_PTX
    u32 c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    u32 f = c & 38u;  // c is 0 or 0xffffffff
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c)
        : "r"(f));
    r.v[0] -= c & 38u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - b.v[i]; r.v[i] = (u32)c; c >>= 32; }
    int64_t f = c ? 38 : 0;
    c = -f;
    for (int i = 0; i < 8; i++) { c += (int64_t)r.v[i]; r.v[i] = (u32)c; c >>= 32; }
    if (c) r.v[0] -= 38u;
#endif
    return r;
}
ACT_FN fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }

// ---- add / sub for PUBLIC data (ACT_RARE) ------------------------------------------------------------------
// The second carry pass of fe_add / fe_sub only matters when folding the wrapped 2^256 (+-38) carries (borrows) out of the
// LOW word, i.e. when the low word is within 38 of 2^32 (of 0): about 1 in 10^8 operations.  fe_add_v / fe_sub_v apply the
// fold to the low word and call an out-of-line ripple in that case: 5-6 instructions fewer on the common path (the check and
// the reconvergence pair cost 4), same result bit for bit; +1.5 % in the range kernel (profiles/r01k_variants_rare_branch.txt).
// The branch depends on the data, so these forms are for public values only (range-proof commitments, fixed-base walks over
// public scalars); everything that touches the key keeps the branch-free fe_add / fe_sub.
#ifndef ACT_RARE
#define ACT_RARE 1
#endif
#if ACT_PTX && ACT_RARE
ACT_NOINLINE fe fe_ripple_up_(fe r) {        // + 2^32, and the fold of a second wrap (then the low word is below 38)
    u32 c;
    asm("add.cc.u32 %0, %0, 1;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, 0, 0;"
        : "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c));
    r.v[0] += c * 38u;
    return r;
}
ACT_NOINLINE fe fe_ripple_down_(fe r) {      // - 2^32, and the fold of a second wrap (then the low word is above 2^32 - 39)
    u32 c;
    asm("sub.cc.u32 %0, %0, 1;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.u32 %7, 0, 0;"
        : "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c));
    r.v[0] -= c & 38u;
    return r;
}
ACT_FN fe fe_add_v(const fe& a, const fe& b) {
    fe r;
    u32 c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    u32 f = (0u - c) & 38u;
    r.v[0] += f;
    if (r.v[0] < f) r = fe_ripple_up_(r);
    return r;
}
ACT_FN fe fe_sub_v(const fe& a, const fe& b) {
    fe r;
    u32 c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    u32 f = c & 38u;  // c is 0 or 0xffffffff
    u32 lo = r.v[0];
    r.v[0] = lo - f;
    if (lo < f) r = fe_ripple_down_(r);
    return r;
}
#else
ACT_FN fe fe_add_v(const fe& a, const fe& b) { return fe_add(a, b); }
ACT_FN fe fe_sub_v(const fe& a, const fe& b) { return fe_sub(a, b); }
#endif
template <bool VT> ACT_FN fe fe_add_x(const fe& a, const fe& b) { return VT ? fe_add_v(a, b) : fe_add(a, b); }
template <bool VT> ACT_FN fe fe_sub_x(const fe& a, const fe& b) { return VT ? fe_sub_v(a, b) : fe_sub(a, b); }

// ---- sums of PRODUCTS -----------------------------------------------------------------------------
// "tight" = below 2^255 + 2^11, which is what fe_mul / fe_sq return under ACT_TIGHT (fe_fold9).  The sum of two tight values
// is below 2^256 + 2^12: if it wraps, what is left is below 2^12, so the 38 that the lost 2^256 is worth goes onto the low word
// without a second carry pass.  The results are ordinary loose values (< 2^256).  Only for operands that ARE products; the
// host build asserts it on every call (tests/hostsim, -DACT_HOSTSIM_CHECKS).
ACT_FN bool fe_is_tight_(const fe& a) { return a.v[7] < 0x80000000u || (a.v[7] == 0x80000000u && !(a.v[6] | a.v[5] | a.v[4] | a.v[3] | a.v[2] | a.v[1]) && a.v[0] < 2048u); }
#if ACT_TIGHT
ACT_FN fe fe_add_tt(const fe& a, const fe& b) {
    ACT_ASSERT_TIGHT(a); ACT_ASSERT_TIGHT(b);
    fe r;
#if ACT_PTX
    u32 c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    r.v[0] += (0u - c) & 38u;
#else
    u64 c = 0;
    for (int i = 0; i < 8; i++) { c += (u64)a.v[i] + b.v[i]; r.v[i] = (u32)c; c >>= 32; }
    r.v[0] += (u32)c * 38u;
#endif
    return r;
}
// 2a for a tight a: a one-bit shift; the bit that leaves at the top is worth 38, and then what stays is below 2^12
ACT_FN fe fe_dbl_tt(const fe& a) {
    ACT_ASSERT_TIGHT(a);
    fe r;
#if ACT_PTX
    ACT_UNROLL for (int i = 7; i >= 1; i--) r.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 1);
    r.v[0] = (a.v[0] << 1) + ((0u - (a.v[7] >> 31)) & 38u);
#else
    for (int i = 7; i >= 1; i--) r.v[i] = (a.v[i] << 1) | (a.v[i - 1] >> 31);
    r.v[0] = (a.v[0] << 1) + (a.v[7] >> 31) * 38u;
#endif
    return r;
}
#else
ACT_FN fe fe_add_tt(const fe& a, const fe& b) { return fe_add(a, b); }
ACT_FN fe fe_dbl_tt(const fe& a) { return fe_add(a, a); }
#endif

// ---- 8x8 -> 16 limb product ------------------------------------------------------------------
#if ACT_PTX
// acc[0..7] += {a0,a1,a2,a3} * bi placed in non-overlapping 64-bit slots, carry into acc[8]
ACT_FN void fe_row_chain(u32* acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 bi) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bi));
}
// the same row into words that include a still-zero last word: hi(a*b) + 0 + carry <= 2^32 - 1, the chain cannot carry out
// and the capture is dropped (which rows qualify: tools/check_fe_rows.py, checked there on extreme operands)
ACT_FN void fe_row_chain_fresh(u32* acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 bi) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bi));
}
#endif

// t[0..8] (value < 39 * 2^256) -> fe.  ACT_TIGHT: everything from bit 255 up folds as 19 * (t >> 255): the result is
// < 2^255 + 19 * 79 < 2^255 + 2^11 and the chain cannot carry out (word 7 is below 2^31 when the carry arrives); same
// instruction count as folding at 2^256 (one shift + one mask replace the carry capture and its multiply).
ACT_FN fe fe_fold9(u32* t) {
    fe r;
#if ACT_TIGHT
#if ACT_PTX
    u32 f = __funnelshift_l(t[7], t[8], 1) * 19u, t7 = t[7] & 0x7fffffffu;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, 0;\n\t"
        "addc.cc.u32 %2, %10, 0;\n\t"
        "addc.cc.u32 %3, %11, 0;\n\t"
        "addc.cc.u32 %4, %12, 0;\n\t"
        "addc.cc.u32 %5, %13, 0;\n\t"
        "addc.cc.u32 %6, %14, 0;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t7), "r"(f));
#else
    u64 c = (u64)((t[8] << 1) | (t[7] >> 31)) * 19u;
    t[7] &= 0x7fffffffu;
    for (int i = 0; i < 8; i++) { c += t[i]; r.v[i] = (u32)c; c >>= 32; }
#endif
#else
#if ACT_PTX
    u32 f = t[8] * 38u, c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, 0;\n\t"
        "addc.cc.u32 %2, %11, 0;\n\t"
        "addc.cc.u32 %3, %12, 0;\n\t"
        "addc.cc.u32 %4, %13, 0;\n\t"
        "addc.cc.u32 %5, %14, 0;\n\t"
        "addc.cc.u32 %6, %15, 0;\n\t"
        "addc.cc.u32 %7, %16, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(f));
    r.v[0] += c * 38u;
#else
    u64 c = (u64)t[8] * 38u;
    for (int i = 0; i < 8; i++) { c += t[i]; r.v[i] = (u32)c; c >>= 32; }
    r.v[0] += (u32)c * 38u;
#endif
#endif
    return r;
}

// 512-bit r[0..15] -> fe : lo + 38*hi
ACT_FN fe fe_reduce512(const u32* r) {
    u32 t[9];
#if ACT_PTX
    const u32 k38 = 38u;
    ACT_UNROLL for (int i = 0; i < 8; i++) t[i] = r[i];
    t[8] = 0;
    // even hi limbs into slots (0,1)(2,3)(4,5)(6,7); odd hi limbs into (1,2)(3,4)(5,6)(7,8)
    fe_row_chain(t, r[8], r[10], r[12], r[14], k38);
#if ACT_REDUCE_ALT
    // the odd products as free-standing 64-bit values added with one carry chain on the ALU pipe: the slots (1,2)(3,4)..
    // are not even-aligned register pairs, and as multiply-add addends they cost one IMAD.MOV each on the multiply pipe
    {
        u64 q0 = (u64)r[9] * 38u, q1 = (u64)r[11] * 38u, q2 = (u64)r[13] * 38u, q3 = (u64)r[15] * 38u;
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
            : "r"((u32)q0), "r"((u32)(q0 >> 32)), "r"((u32)q1), "r"((u32)(q1 >> 32)), "r"((u32)q2), "r"((u32)(q2 >> 32)), "r"((u32)q3), "r"((u32)(q3 >> 32)));
    }
#else
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
        : "r"(r[9]), "r"(r[11]), "r"(r[13]), "r"(r[15]), "r"(k38));
#endif
#else
    u64 c = 0;
    for (int i = 0; i < 8; i++) { c += (u64)r[i] + (u64)r[8 + i] * 38u; t[i] = (u32)c; c >>= 32; }
    t[8] = (u32)c;
#endif
    return fe_fold9(t);
}

ACT_FN fe fe_mul_inl(const fe& a, const fe& b) {
    u32 r[16];
#if ACT_PTX
    u32 ev[18], od[18];
    ACT_UNROLL for (int i = 0; i < 18; i++) { ev[i] = 0; od[i] = 0; }
    // 16 rows; the first row to reach a new last word needs no carry capture (7 captures instead of 16)
    ACT_UNROLL for (int i = 0; i < 8; i += 2) {
        if (i == 0) fe_row_chain_fresh(ev + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
        else fe_row_chain(ev + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
        fe_row_chain_fresh(od + i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
        fe_row_chain(od + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
        fe_row_chain_fresh(ev + i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
    }
    // r = ev + (od << 32), one 15-word carry chain in a single asm statement (30 operands)
    ACT_UNROLL for (int i = 0; i < 16; i++) r[i] = ev[i];
    asm("add.cc.u32 %0, %0, %15;\n\t"
        "addc.cc.u32 %1, %1, %16;\n\t"
        "addc.cc.u32 %2, %2, %17;\n\t"
        "addc.cc.u32 %3, %3, %18;\n\t"
        "addc.cc.u32 %4, %4, %19;\n\t"
        "addc.cc.u32 %5, %5, %20;\n\t"
        "addc.cc.u32 %6, %6, %21;\n\t"
        "addc.cc.u32 %7, %7, %22;\n\t"
        "addc.cc.u32 %8, %8, %23;\n\t"
        "addc.cc.u32 %9, %9, %24;\n\t"
        "addc.cc.u32 %10, %10, %25;\n\t"
        "addc.cc.u32 %11, %11, %26;\n\t"
        "addc.cc.u32 %12, %12, %27;\n\t"
        "addc.cc.u32 %13, %13, %28;\n\t"
        "addc.u32 %14, %14, %29;"
        : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
          "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]),
          "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
    for (int i = 0; i < 16; i++) r[i] = 0;
    for (int i = 0; i < 8; i++) {
        u64 c = 0;
        for (int j = 0; j < 8; j++) { c += (u64)a.v[j] * b.v[i] + r[i + j]; r[i + j] = (u32)c; c >>= 32; }
        r[i + 8] = (u32)c;
    }
#endif
    return fe_reduce512(r);
}

#if ACT_PTX
// ---- GENERATED by tools/gen_fe_sq.py (layout verified there against big-int squaring) ----
// 28 off-diagonal products in even/odd carry chains, doubled, plus the 8 diagonal squares: 36 wide
// multiply-adds instead of 64.  A chain ends with a carry capture only where its last word may already be non-zero.
ACT_FN void fe_sq_wide(u32* r, const fe& a) {
    u32 ev[16], od[16];
    ACT_UNROLL for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, %12, %6;\n\t"
        "madc.hi.u32 %7, %8, %12, %7;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "madc.lo.cc.u32 %4, %6, %9, %4;\n\t"
        "madc.hi.u32 %5, %6, %9, %5;"
        : "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
        : "r"(a.v[0]), "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %7, %8, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %7, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %7, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]), "+r"(od[8])
        : "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "madc.lo.cc.u32 %4, %6, %9, %4;\n\t"
        "madc.hi.u32 %5, %6, %9, %5;"
        : "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9])
        : "r"(a.v[1]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "madc.lo.cc.u32 %4, %6, %9, %4;\n\t"
        "madc.hi.u32 %5, %6, %9, %5;"
        : "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]), "+r"(od[8]), "+r"(od[9])
        : "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %5, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(ev[6]), "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9]), "+r"(ev[10])
        : "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %5, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(od[6]), "+r"(od[7]), "+r"(od[8]), "+r"(od[9]), "+r"(od[10])
        : "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "madc.lo.cc.u32 %2, %4, %6, %2;\n\t"
        "madc.hi.u32 %3, %4, %6, %3;"
        : "+r"(ev[8]), "+r"(ev[9]), "+r"(ev[10]), "+r"(ev[11])
        : "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "madc.lo.cc.u32 %2, %4, %6, %2;\n\t"
        "madc.hi.u32 %3, %4, %6, %3;"
        : "+r"(od[8]), "+r"(od[9]), "+r"(od[10]), "+r"(od[11])
        : "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(ev[10]), "+r"(ev[11]), "+r"(ev[12])
        : "r"(a.v[4]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(od[10]), "+r"(od[11]), "+r"(od[12])
        : "r"(a.v[5]), "r"(a.v[6]));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, %1;"
        : "+r"(ev[12]), "+r"(ev[13])
        : "r"(a.v[5]), "r"(a.v[7]));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, %1;"
        : "+r"(od[12]), "+r"(od[13])
        : "r"(a.v[6]), "r"(a.v[7]));
    ACT_UNROLL for (int i = 0; i < 16; i++) r[i] = ev[i];
    asm("add.cc.u32 %0, %0, %15;\n\t"
        "addc.cc.u32 %1, %1, %16;\n\t"
        "addc.cc.u32 %2, %2, %17;\n\t"
        "addc.cc.u32 %3, %3, %18;\n\t"
        "addc.cc.u32 %4, %4, %19;\n\t"
        "addc.cc.u32 %5, %5, %20;\n\t"
        "addc.cc.u32 %6, %6, %21;\n\t"
        "addc.cc.u32 %7, %7, %22;\n\t"
        "addc.cc.u32 %8, %8, %23;\n\t"
        "addc.cc.u32 %9, %9, %24;\n\t"
        "addc.cc.u32 %10, %10, %25;\n\t"
        "addc.cc.u32 %11, %11, %26;\n\t"
        "addc.cc.u32 %12, %12, %27;\n\t"
        "addc.cc.u32 %13, %13, %28;\n\t"
        "addc.u32 %14, %14, %29;"
        : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
    asm("add.cc.u32 %0, %0, %0;\n\t"
        "addc.cc.u32 %1, %1, %1;\n\t"
        "addc.cc.u32 %2, %2, %2;\n\t"
        "addc.cc.u32 %3, %3, %3;\n\t"
        "addc.cc.u32 %4, %4, %4;\n\t"
        "addc.cc.u32 %5, %5, %5;\n\t"
        "addc.cc.u32 %6, %6, %6;\n\t"
        "addc.cc.u32 %7, %7, %7;\n\t"
        "addc.cc.u32 %8, %8, %8;\n\t"
        "addc.cc.u32 %9, %9, %9;\n\t"
        "addc.cc.u32 %10, %10, %10;\n\t"
        "addc.cc.u32 %11, %11, %11;\n\t"
        "addc.cc.u32 %12, %12, %12;\n\t"
        "addc.cc.u32 %13, %13, %13;\n\t"
        "addc.cc.u32 %14, %14, %14;\n\t"
        "addc.u32 %15, %15, %15;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
}

// ---- end GENERATED ----
#endif

ACT_FN fe fe_sq_inl(const fe& a) {
#if ACT_PTX
    u32 r[16];
    fe_sq_wide(r, a);
    return fe_reduce512(r);
#else
    return fe_mul_inl(a, a);
#endif
}
// The multiply and the square are REAL calls (arguments and result travel in registers, checked in
// SASS: no stack traffic).  Inlining them everywhere made spend_range_kernel 460 KB of code and ncu
// showed the warps stalled on instruction fetch (stall_no_instruction 7.3 of 11.7 cycles per issue);
// as calls the whole hot loop fits the instruction cache.
#ifndef ACT_FE_CALLS
#define ACT_FE_CALLS 1
#endif
#if ACT_FE_CALLS
ACT_NOINLINE fe fe_mul(fe a, fe b) { return fe_mul_inl(a, b); }
ACT_NOINLINE fe fe_sq(fe a) { return fe_sq_inl(a); }
#else
ACT_FN fe fe_mul(const fe& a, const fe& b) { return fe_mul_inl(a, b); }
ACT_FN fe fe_sq(const fe& a) { return fe_sq_inl(a); }
#endif

ACT_FN fe fe_sqn(fe a, int n) {
    ACT_NOUNROLL for (int i = 0; i < n; i++) a = fe_sq(a);
    return a;
}

// ---- canonical form, predicates ----------------------------------------------------------------
// fully reduced representative in [0, p)
ACT_FN fe fe_canon(const fe& a) {
    fe r = a;
    // fold bit 255: x = (x mod 2^255) + 19*(x >> 255)  -> < 2^255 + 19
    u32 top = r.v[7] >> 31;
    r.v[7] &= 0x7fffffffu;
    u64 c = (u64)top * 19u;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (u32)c; c >>= 32; }
    // if x >= p then x - p = (x + 19) mod 2^255
    fe t;
    c = 19;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += r.v[i]; t.v[i] = (u32)c; c >>= 32; }
    u32 ge = t.v[7] >> 31;  // bit 255 of x+19 set <=> x >= p
    t.v[7] &= 0x7fffffffu;
    u32 m = 0u - ge;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = (r.v[i] & ~m) | (t.v[i] & m);
    return r;
}
ACT_FN u32 fe_is_negative(const fe& a) { return fe_canon(a).v[0] & 1u; }
ACT_FN u32 fe_is_zero(const fe& a) {
    fe c = fe_canon(a);
    u32 o = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) o |= c.v[i];
    return o == 0;
}
ACT_FN u32 fe_eq(const fe& a, const fe& b) { return fe_is_zero(fe_sub(a, b)); }
// r = b ? y : x   (branch-free)
ACT_FN fe fe_select(const fe& x, const fe& y, u32 b) {
    fe r;
    u32 m = 0u - (b & 1u);
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = (x.v[i] & ~m) | (y.v[i] & m);
    return r;
}
ACT_FN fe fe_cneg(const fe& x, u32 b) { return fe_select(x, fe_neg(x), b); }
ACT_FN fe fe_abs(const fe& x) { return fe_cneg(x, fe_is_negative(x)); }

// ---- bytes -------------------------------------------------------------------------------------
// little-endian words in, bit 255 ignored (dalek FieldElement::from_bytes)
ACT_FN fe fe_from_words(const u32* w) {
    fe r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = w[i];
    r.v[7] &= 0x7fffffffu;
    return r;
}
ACT_FN void fe_to_words(u32* w, const fe& a) {
    fe c = fe_canon(a);
    ACT_UNROLL for (int i = 0; i < 8; i++) w[i] = c.v[i];
}

// ---- exponentiations -----------------------------------------------------------------------------
// z^(2^250-1), z^11
ACT_NOINLINE void fe_pow22501(fe* t250, fe* z11, const fe* zp) {
    fe z = *zp;
    fe t0 = fe_sq(z);
    fe t1 = fe_sqn(t0, 2);
    t1 = fe_mul(z, t1);
    t0 = fe_mul(t0, t1);
    fe t2 = fe_sq(t0);
    t1 = fe_mul(t1, t2);
    t2 = fe_sqn(t1, 5); t1 = fe_mul(t2, t1);
    t2 = fe_sqn(t1, 10); t2 = fe_mul(t2, t1);
    fe t3 = fe_sqn(t2, 20); t2 = fe_mul(t3, t2);
    t2 = fe_sqn(t2, 10); t1 = fe_mul(t2, t1);
    t2 = fe_sqn(t1, 50); t2 = fe_mul(t2, t1);
    t3 = fe_sqn(t2, 100); t2 = fe_mul(t3, t2);
    t2 = fe_sqn(t2, 50); t1 = fe_mul(t2, t1);
    *t250 = t1; *z11 = t0;
}
ACT_FN fe fe_invert(const fe& z) {
    fe t, z11;
    fe_pow22501(&t, &z11, &z);
    t = fe_sqn(t, 5);
    return fe_mul(t, z11);
}
ACT_FN fe fe_pow22523(const fe& z) {
    fe t, z11;
    fe_pow22501(&t, &z11, &z);
    t = fe_sqn(t, 2);
    return fe_mul(t, z);
}

// RFC 9496 4.2 SQRT_RATIO_M1 / dalek FieldElement::sqrt_ratio_i.  *was_square out, returns r >= 0.
ACT_FN fe fe_sqrt_ratio_i(u32* was_square, const fe& u, const fe& v) {
    fe v3 = fe_mul(fe_sq(v), v);
    fe v7 = fe_mul(fe_sq(v3), v);
    fe r = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
    fe check = fe_mul(v, fe_sq(r));
    fe neg_u = fe_neg(u);
    fe neg_u_i = fe_mul(neg_u, FE_SQRT_M1);
    u32 correct = fe_eq(check, u);
    u32 flipped = fe_eq(check, neg_u);
    u32 flipped_i = fe_eq(check, neg_u_i);
    fe r_prime = fe_mul(FE_SQRT_M1, r);
    r = fe_select(r, r_prime, flipped | flipped_i);
    r = fe_abs(r);
    *was_square = correct | flipped;
    return r;
}
// 1/sqrt(v) when v is a nonzero square (dalek invsqrt = sqrt_ratio_i(1, v))
ACT_FN fe fe_invsqrt(u32* was_square, const fe& v) { return fe_sqrt_ratio_i(was_square, fe_one(), v); }
