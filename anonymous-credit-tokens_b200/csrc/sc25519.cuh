// sc25519.cuh -- scalars mod l = 2^252 + 27742317777372353535851937790883648493, 8x32-bit limbs.
//
// Branch-free throughout (every operation may see the issuer's secret x, the nonce alpha or
// (e+x)^-1): Montgomery CIOS multiplication, masked conditional subtraction, fixed addition chain
// for inversion.  Replaces curve25519-dalek's Scalar as used at /root/reference src/lib.rs:638,645,
// 660,795-861, src/cbor.rs:85 (from_bytes_mod_order) and src/transcript.rs:153
// (from_bytes_mod_order_wide).
#pragma once
#include "fe25519.cuh"

struct sc { u32 v[8]; };

ACT_CONST u32 SC_L_[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u};
ACT_CONST u32 SC_R_[8] = {0x8d98951du, 0xd6ec3174u, 0x737dcf70u, 0xc6ef5bf4u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x0fffffffu};   // 2^256 mod l
ACT_CONST u32 SC_RR_[8] = {0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu};  // 2^512 mod l
ACT_CONST u32 SC_LM2_[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u}; // l - 2
#define SC_LFACTOR 0x12547e1bu  // -l^-1 mod 2^32

ACT_FN sc sc_const(const u32* c) {
    sc r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}
ACT_FN sc sc_zero() {
    sc r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
ACT_FN sc sc_from_u32(u32 x) { sc r = sc_zero(); r.v[0] = x; return r; }

// r = t - l if t >= l else t, for t < 2^256 (t given with an optional 9th word `hi`), branch-free
ACT_FN sc sc_csub_l(const u32* t, u32 hi) {
    u32 d[8];
    int64_t c = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += (int64_t)t[i] - SC_L_[i]; d[i] = (u32)c; c >>= 32; }
    c += hi;
    // c == -1 : borrow -> keep t ; c >= 0 : use d
    u32 m = (u32)(c >> 63);  // all ones if borrow
    sc r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = (t[i] & m) | (d[i] & ~m);
    return r;
}
ACT_FN sc sc_add(const sc& a, const sc& b) {
    u32 t[8];
    u64 c = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += (u64)a.v[i] + b.v[i]; t[i] = (u32)c; c >>= 32; }
    return sc_csub_l(t, (u32)c);
}
ACT_FN sc sc_neg(const sc& a) {
    // l - a, mapped to 0 when a == 0
    u32 t[8], nz = 0;
    int64_t c = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += (int64_t)SC_L_[i] - a.v[i]; t[i] = (u32)c; c >>= 32; nz |= a.v[i]; }
    u32 m = (nz != 0) ? 0xffffffffu : 0u;
    sc r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = t[i] & m;
    return r;
}
ACT_FN sc sc_sub(const sc& a, const sc& b) { return sc_add(a, sc_neg(b)); }

// a*b*2^-256 mod l.  Needs a*b < l*2^256 (b < l suffices).  Output < l.
ACT_FN sc sc_montmul(const sc& a, const sc& b) {
    u32 t[10];
    ACT_UNROLL for (int i = 0; i < 10; i++) t[i] = 0;
    ACT_NOUNROLL for (int i = 0; i < 8; i++) {
        u64 c = 0;
        u32 bi = b.v[i];
        ACT_UNROLL for (int j = 0; j < 8; j++) { c += (u64)a.v[j] * bi + t[j]; t[j] = (u32)c; c >>= 32; }
        c += t[8]; t[8] = (u32)c; t[9] = (u32)(c >> 32);
        u32 m = t[0] * SC_LFACTOR;
        c = (u64)m * SC_L_[0] + t[0];
        c >>= 32;
        ACT_UNROLL for (int j = 1; j < 8; j++) { c += (u64)m * SC_L_[j] + t[j]; t[j - 1] = (u32)c; c >>= 32; }
        c += t[8]; t[7] = (u32)c; t[8] = t[9] + (u32)(c >> 32);
    }
    return sc_csub_l(t, t[8]);
}
ACT_FN sc sc_mul(const sc& a, const sc& b) { return sc_montmul(sc_montmul(a, b), sc_const(SC_RR_)); }

// Scalar::from_bytes_mod_order: any 256-bit value -> [0, l)
ACT_FN sc sc_from_words(const u32* w) {
    sc a;
    ACT_UNROLL for (int i = 0; i < 8; i++) a.v[i] = w[i];
    return sc_montmul(a, sc_const(SC_R_));
}
// Scalar::from_bytes_mod_order_wide: 512-bit little-endian -> [0, l)
ACT_FN sc sc_from_wide(const u32* w) {
    sc lo, hi;
    ACT_UNROLL for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[8 + i]; }
    return sc_add(sc_montmul(lo, sc_const(SC_R_)), sc_montmul(hi, sc_const(SC_RR_)));
}
ACT_FN u32 sc_eq(const sc& a, const sc& b) {
    u32 d = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
ACT_FN u32 sc_is_zero(const sc& a) { return sc_eq(a, sc_zero()); }

// Scalar::invert = a^(l-2) (0 -> 0).  The exponent is public; the base may be secret.
ACT_NOINLINE void sc_invert_(sc* out, const sc* in) {
    sc am = sc_montmul(*in, sc_const(SC_RR_));   // a*R
    sc acc = sc_const(SC_R_);                    // 1*R
    ACT_NOUNROLL for (int i = 252; i >= 0; i--) {
        acc = sc_montmul(acc, acc);
        u32 bit = (SC_LM2_[i >> 5] >> (i & 31)) & 1u;
        sc t = sc_montmul(acc, am);
        u32 m = 0u - bit;
        ACT_UNROLL for (int k = 0; k < 8; k++) acc.v[k] = (acc.v[k] & ~m) | (t.v[k] & m);
    }
    *out = sc_montmul(acc, sc_from_u32(1));
}
ACT_FN sc sc_invert(const sc& a) { sc r; sc_invert_(&r, &a); return r; }

// s/2 mod l (l is odd): (s + (s odd ? l : 0)) >> 1, branch-free (the parity becomes a mask), so it serves public and secret
// scalars alike: verification scalars are halved so that the batched double-and-encode of ge25519.cuh yields the encoding of
// the original point, and the signing tail halves (e+x)^-1 and alpha for the same reason (act_device.cuh bbs_sign_).
ACT_FN sc sc_half(const sc& s) {
    u32 m = 0u - (s.v[0] & 1u);
    u32 t[9];
    u64 c = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += (u64)s.v[i] + (SC_L_[i] & m); t[i] = (u32)c; c >>= 32; }
    t[8] = (u32)c;
    sc r;
    ACT_UNROLL for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
    return r;
}

// ---- signed fixed-window recoding ----------------------------------------------------------------
// digits d_i in [-2^(W-1), 2^(W-1)) with s = sum d_i 2^(W i): d_i = window_i(s + C) - 2^(W-1), where
// C has 2^(W-1) in every window.  One 256-bit addition instead of a carry-propagating digit loop;
// needs s + C < 2^256, true for s < l and W in {4, 8}.
template <int W>
ACT_FN sc sc_bias(const sc& s) {
    const u32 cw = (W == 4) ? 0x88888888u : 0x80808080u;
    sc r;
    u64 c = 0;
    ACT_UNROLL for (int i = 0; i < 8; i++) { c += (u64)s.v[i] + cw; r.v[i] = (u32)c; c >>= 32; }
    return r;
}
// window i of a biased scalar, as a signed digit
template <int W>
ACT_FN int sc_digit(const sc& biased, int i) {
    const int per = 32 / W;
    u32 w = biased.v[i / per] >> ((i % per) * W);
    return (int)(w & ((1u << W) - 1u)) - (1 << (W - 1));
}
