// ge25519.cuh -- twisted Edwards (a = -1) point arithmetic in extended coordinates and the
// ristretto255 encode / decode / one-way map (RFC 9496 4.3.1, 4.3.2, 4.3.4).
//
// Replaces curve25519-dalek's RistrettoPoint / CompressedRistretto as used by the reference at
// src/lib.rs:629-651, 787-854 (point +, -, *), src/transcript.rs:106 (compress) and
// src/cbor.rs:68-71 (decompress, with all rejection conditions).
#pragma once
#include "fe25519.cuh"
#include "sc25519.cuh"

struct ge { fe X, Y, Z, T; };             // extended: x = X/Z, y = Y/Z, xy = T/Z
struct ge_cached { fe YpX, YmX, Z, T2d; };  // "projective Niels"
struct ge_niels { fe ypx, ymx, xy2d; };     // affine Niels (Z = 1)

ACT_FN ge ge_identity() { ge r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
ACT_FN ge ge_basepoint() { ge r; r.X = fe_const(FE_BX_); r.Y = fe_const(FE_BY_); r.Z = fe_one(); r.T = fe_const(FE_BT_); return r; }
ACT_FN ge_cached ge_cached_identity() { ge_cached r; r.YpX = fe_one(); r.YmX = fe_one(); r.Z = fe_one(); r.T2d = fe_zero(); return r; }
ACT_FN ge_niels ge_niels_identity() { ge_niels r; r.ypx = fe_one(); r.ymx = fe_one(); r.xy2d = fe_zero(); return r; }

ACT_FN ge_cached ge_to_cached(const ge& p) {
    ge_cached r;
    r.YpX = fe_add(p.Y, p.X); r.YmX = fe_sub(p.Y, p.X); r.Z = p.Z; r.T2d = fe_mul(p.T, FE_D2);
    return r;
}
ACT_FN ge ge_neg(const ge& p) { ge r; r.X = fe_neg(p.X); r.Y = p.Y; r.Z = p.Z; r.T = fe_neg(p.T); return r; }

// ---- ALU-pipe copies around the field-multiply calls ---------------------------------------------------------
// ptxas marshals call arguments and results with IMAD.MOV, which runs on the integer-multiply pipe -- the pipe this
// code is bound by.  An exclusive-or with a run-time zero that ptxas cannot fold is a copy that is guaranteed to run on the
// ALU pipe (LOP3; an add would become IMAD.IADD), and the register allocator can then place its result directly in the argument registers.
// ACT_ALU_COPY bit 0: copy arguments, bit 1: copy results.  (Round 1i repeated the experiment with hand-placed one-for-one
// copies -- 0 IMAD.MOV left in the doubling, same instruction count: no change in time, profiles/r01i_variants_alu_copy_targeted.txt:
// the kernel is bound by instructions issued, not by the multiply pipe's cycles.)
#ifndef ACT_ALU_COPY
#define ACT_ALU_COPY 0
#endif
#if defined(__CUDACC__) && ACT_ALU_COPY
__device__ u32 act_zero_word = 0;
#endif
#if ACT_PTX && ACT_ALU_COPY
ACT_FN u32 act_zero() { return __ldg(&act_zero_word); }
ACT_FN fe fe_cp(const fe& a, u32 z) {
    fe r;
    ACT_UNROLL for (int i = 0; i < 8; i++) asm("xor.b32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(z));
    return r;
}
ACT_FN fe fe_mul_c(const fe& a, const fe& b, u32 z) {
    fe r = (ACT_ALU_COPY & 1) ? fe_mul(fe_cp(a, z), fe_cp(b, z)) : fe_mul(a, b);
    return (ACT_ALU_COPY & 2) ? fe_cp(r, z) : r;
}
ACT_FN fe fe_sq_c(const fe& a, u32 z) {
    fe r = (ACT_ALU_COPY & 1) ? fe_sq(fe_cp(a, z)) : fe_sq(a);
    return (ACT_ALU_COPY & 2) ? fe_cp(r, z) : r;
}
#else
ACT_FN u32 act_zero() { return 0; }
ACT_FN fe fe_mul_c(const fe& a, const fe& b, u32) { return fe_mul(a, b); }
ACT_FN fe fe_sq_c(const fe& a, u32) { return fe_sq(a); }
#endif

// The three hot point operations.  ACT_GE_CALLS=1 makes THEM the call boundary (field multiplies inlined inside:
// one round of argument moves per 7-8 multiplies).  Measured on B200 it is slower (90k vs 119k proofs/s): the
// three bodies together no longer fit the instruction cache.  Default 0 = calls at the field-multiply level.
#ifndef ACT_GE_CALLS
#define ACT_GE_CALLS 0
#endif
#if ACT_GE_CALLS
#define ACT_GE_FN ACT_NOINLINE
#define GE_MUL(a, b) fe_mul_inl(a, b)
#define GE_SQ(a) fe_sq_inl(a)
#else
#define ACT_GE_FN ACT_FN
#define GE_MUL(a, b) fe_mul_c(a, b, zc_)
#define GE_SQ(a) fe_sq_c(a, zc_)
#endif
// r = p + q, 8M
ACT_GE_FN ge ge_add_cached(ge p, ge_cached q) {
    u32 zc_ = act_zero(); (void)zc_;
    fe PP = GE_MUL(fe_add(p.Y, p.X), q.YpX);
    fe MM = GE_MUL(fe_sub(p.Y, p.X), q.YmX);
    fe E = fe_sub(PP, MM), H = fe_add_tt(PP, MM);
    fe TT = GE_MUL(p.T, q.T2d);
    fe ZZ = GE_MUL(p.Z, q.Z);
    fe ZZ2 = fe_dbl_tt(ZZ);
    fe G = fe_add(ZZ2, TT), F = fe_sub(ZZ2, TT);
    ge r;
    r.X = GE_MUL(E, F); r.Y = GE_MUL(H, G); r.Z = GE_MUL(G, F); r.T = GE_MUL(E, H);
    return r;
}
// the addition of the window loops: r = p + (neg ? -q : q) with a run-time (warp-uniform) choice of producing T -- a doubling
// (which does not read T) follows the last addition of every window: 7M instead of 8M there.  The negation of q is folded
// in: -q swaps q's Y+X / Y-X and flips the sign of the T term, which swaps G and F (no field negation needed).
template <bool VT = false>   // VT: public data, the variable-time add / sub forms (fe_add_v, fe_sub_v)
ACT_FN ge ge_add_cached_u(const ge& p, const ge_cached& q, u32 neg, bool want_t) {
    u32 zc_ = act_zero(); (void)zc_;
    // (statement and operand order chosen by the marshalling moves ptxas needs in the range kernel: 363 vs 381 instructions)
    fe TT = GE_MUL(p.T, q.T2d);
    fe ZZ = GE_MUL(p.Z, q.Z);
    fe ZZ2 = fe_dbl_tt(ZZ);                              // ZZ, PP, MM are products: tight
    fe G0 = fe_add_x<VT>(ZZ2, TT), F0 = fe_sub_x<VT>(ZZ2, TT);
    fe PP = GE_MUL(fe_select(q.YpX, q.YmX, neg), fe_add_x<VT>(p.Y, p.X));
    fe MM = GE_MUL(fe_select(q.YmX, q.YpX, neg), fe_sub_x<VT>(p.Y, p.X));
    fe E = fe_sub_x<VT>(PP, MM), H = fe_add_tt(PP, MM);
    fe G = fe_select(G0, F0, neg), F = fe_select(F0, G0, neg);
    ge r;
    r.X = GE_MUL(E, F); r.Z = GE_MUL(G, F); r.Y = GE_MUL(H, G);     // order chosen by the moves ptxas needs (369 vs 381 instructions)
    r.T = p.T;
    if (want_t) r.T = GE_MUL(E, H);
    return r;
}
// r = p + q for affine-Niels q, 7M
ACT_GE_FN ge ge_add_niels(ge p, ge_niels q) {
    u32 zc_ = act_zero(); (void)zc_;
    fe PP = GE_MUL(fe_add(p.Y, p.X), q.ypx);
    fe MM = GE_MUL(fe_sub(p.Y, p.X), q.ymx);
    fe E = fe_sub(PP, MM), H = fe_add_tt(PP, MM);
    fe TT = GE_MUL(p.T, q.xy2d);
    fe ZZ2 = fe_add(p.Z, p.Z);
    fe G = fe_add(ZZ2, TT), F = fe_sub(ZZ2, TT);
    ge r;
    r.X = GE_MUL(E, F); r.Y = GE_MUL(H, G); r.Z = GE_MUL(G, F); r.T = GE_MUL(E, H);
    return r;
}
// r = p + (neg ? -q : q) for an affine-Niels table entry, the negation folded in as in ge_add_cached_u (public data)
template <bool VT = false>
ACT_FN ge ge_add_niels_n(const ge& p, const ge_niels& q, u32 neg) {
    u32 zc_ = act_zero(); (void)zc_;
    fe PP = GE_MUL(fe_add_x<VT>(p.Y, p.X), fe_select(q.ypx, q.ymx, neg));
    fe MM = GE_MUL(fe_sub_x<VT>(p.Y, p.X), fe_select(q.ymx, q.ypx, neg));
    fe E = fe_sub_x<VT>(PP, MM), H = fe_add_tt(PP, MM);      // PP, MM are products: tight
    fe TT = GE_MUL(p.T, q.xy2d);
    fe ZZ2 = fe_add_x<VT>(p.Z, p.Z);
    fe G0 = fe_add_x<VT>(ZZ2, TT), F0 = fe_sub_x<VT>(ZZ2, TT);
    fe G = fe_select(G0, F0, neg), F = fe_select(F0, G0, neg);
    ge r;
    r.X = GE_MUL(E, F); r.Y = GE_MUL(H, G); r.Z = GE_MUL(G, F); r.T = GE_MUL(E, H);
    return r;
}
ACT_FN ge ge_add(const ge& p, const ge& q) { return ge_add_cached(p, ge_to_cached(q)); }
ACT_FN ge ge_sub(const ge& p, const ge& q) { return ge_add_cached(p, ge_to_cached(ge_neg(q))); }

// r = 2p.  T of the input is not read (4S + 4M; the T product is computed unconditionally in the call form
// so that a single instance of the code serves every doubling).
ACT_GE_FN ge ge_dbl_t(ge p) {
    u32 zc_ = act_zero(); (void)zc_;
    fe XX = GE_SQ(p.X), YY = GE_SQ(p.Y), ZZ = GE_SQ(p.Z);
    fe ZZ2 = fe_dbl_tt(ZZ);
    fe XpY2 = GE_SQ(fe_add(p.X, p.Y));
    fe Yc = fe_add_tt(YY, XX), Zc = fe_sub(YY, XX);
    fe Xc = fe_sub(XpY2, Yc), Tc = fe_sub(ZZ2, Zc);
    ge r;
    r.X = GE_MUL(Xc, Tc); r.Y = GE_MUL(Yc, Zc); r.Z = GE_MUL(Zc, Tc); r.T = GE_MUL(Xc, Yc);
    return r;
}
// doubling without the T output (4S + 3M): p.T is carried through untouched and must not be used
ACT_GE_FN ge ge_dbl_not(ge p) {
    u32 zc_ = act_zero(); (void)zc_;
    fe XX = GE_SQ(p.X), YY = GE_SQ(p.Y), ZZ = GE_SQ(p.Z);
    fe ZZ2 = fe_dbl_tt(ZZ);
    fe XpY2 = GE_SQ(fe_add(p.X, p.Y));
    fe Yc = fe_add_tt(YY, XX), Zc = fe_sub(YY, XX);
    fe Xc = fe_sub(XpY2, Yc), Tc = fe_sub(ZZ2, Zc);
    ge r;
    r.X = GE_MUL(Xc, Tc); r.Y = GE_MUL(Yc, Zc); r.Z = GE_MUL(Zc, Tc); r.T = p.T;
    return r;
}
ACT_FN ge ge_dbl(const ge& p, bool want_t) { return want_t ? ge_dbl_t(p) : ge_dbl_not(p); }

// doubling with a run-time (warp-uniform) choice of producing T: one instance of the code serves both forms, which
// keeps the window loops small.  (Inlining the field multiplications into the point operations was measured on
// B200: the 17-37 KB loop bodies miss the instruction cache and run 7-25 % slower than the call form, see DESIGN.md.)
template <bool VT = false>
ACT_FN ge ge_dbl_u(const ge& p, bool want_t) {
    u32 zc_ = act_zero(); (void)zc_;
    fe XX = GE_SQ(p.X), YY = GE_SQ(p.Y);
    fe Yc = fe_add_tt(YY, XX), Zc = fe_sub_x<VT>(YY, XX);     // XX, YY, ZZ are products: tight
    fe ZZ = GE_SQ(p.Z);
    fe ZZ2 = fe_dbl_tt(ZZ);
    fe Tc = fe_sub_x<VT>(ZZ2, Zc);
    fe XpY2 = GE_SQ(fe_add_x<VT>(p.X, p.Y));
    fe Xc = fe_sub_x<VT>(XpY2, Yc);
    ge r;
    // (call and operand order chosen by the marshalling moves ptxas needs for them in the range kernel: 262 vs 272 instructions)
    r.Y = GE_MUL(Yc, Zc); r.Z = GE_MUL(Tc, Zc); r.X = GE_MUL(Tc, Xc);
    r.T = p.T;
    if (want_t) r.T = GE_MUL(Xc, Yc);
    return r;
}

// branch-free negate-if of table entries (negation swaps y+x / y-x and flips the sign of the t term)
ACT_FN ge_cached ge_cached_cneg(const ge_cached& q, u32 neg) {
    ge_cached r;
    r.YpX = fe_select(q.YpX, q.YmX, neg); r.YmX = fe_select(q.YmX, q.YpX, neg);
    r.Z = q.Z; r.T2d = fe_select(q.T2d, fe_neg(q.T2d), neg);
    return r;
}
ACT_FN ge_niels ge_niels_cneg(const ge_niels& q, u32 neg) {
    ge_niels r;
    r.ypx = fe_select(q.ypx, q.ymx, neg); r.ymx = fe_select(q.ymx, q.ypx, neg);
    r.xy2d = fe_select(q.xy2d, fe_neg(q.xy2d), neg);
    return r;
}
// affine Niels form of a point with Z = 1
ACT_FN ge_niels ge_affine_to_niels(const fe& x, const fe& y) {
    ge_niels r;
    r.ypx = fe_add(y, x); r.ymx = fe_sub(y, x); r.xy2d = fe_mul(fe_mul(x, y), FE_D2);
    return r;
}
ACT_FN ge_niels ge_to_niels(const ge& p) {
    fe zi = fe_invert(p.Z);
    return ge_affine_to_niels(fe_mul(p.X, zi), fe_mul(p.Y, zi));
}
ACT_FN ge ge_from_niels(const ge_niels& n) {  // only for tests / table checks
    // y = (ypx+ymx)/2, x = (ypx-ymx)/2 : keep projective with Z = 2
    ge r;
    r.Y = fe_add(n.ypx, n.ymx); r.X = fe_sub(n.ypx, n.ymx); r.Z = fe_add(fe_one(), fe_one());
    fe zi = fe_invert(r.Z);
    fe x = fe_mul(r.X, zi), y = fe_mul(r.Y, zi);
    r.X = x; r.Y = y; r.Z = fe_one(); r.T = fe_mul(x, y);
    return r;
}

// ---- ristretto255 --------------------------------------------------------------------------------
// RFC 9496 4.3.1 / dalek CompressedRistretto::decompress.  w = 8 little-endian words of the wire
// bytes.  Returns 1 and the point (Z = 1) when the encoding is valid, else 0 (point = identity).
ACT_NOINLINE u32 ristretto_decode_(ge* out, const u32* w) {
    fe s = fe_from_words(w);
    fe sc_ = fe_canon(s);
    u32 canonical = ((w[7] >> 31) == 0);
    ACT_UNROLL for (int i = 0; i < 8; i++) canonical &= (sc_.v[i] == (i == 7 ? (w[7] & 0x7fffffffu) : w[i]));
    u32 negative = w[0] & 1u;
    fe ss = fe_sq(s);
    fe u1 = fe_sub(fe_one(), ss), u2 = fe_add(fe_one(), ss);
    fe u2_sqr = fe_sq(u2);
    fe v = fe_sub(fe_neg(fe_mul(FE_D, fe_sq(u1))), u2_sqr);
    u32 ok;
    fe I = fe_invsqrt(&ok, fe_mul(v, u2_sqr));
    fe Dx = fe_mul(I, u2);
    fe Dy = fe_mul(fe_mul(I, Dx), v);
    fe x = fe_abs(fe_mul(fe_add(s, s), Dx));
    fe y = fe_mul(u1, Dy);
    fe t = fe_mul(x, y);
    u32 valid = canonical & (negative ^ 1u) & ok & (fe_is_negative(t) ^ 1u) & (fe_is_zero(y) ^ 1u);
    ge id = ge_identity();
    out->X = fe_select(id.X, x, valid); out->Y = fe_select(id.Y, y, valid);
    out->Z = fe_one(); out->T = fe_select(id.T, t, valid);
    return valid;
}
// RFC 9496 4.3.2 / dalek RistrettoPoint::compress.  Writes 8 canonical little-endian words.
ACT_NOINLINE void ristretto_encode_(u32* w, const ge* pp) {
    ge p = *pp;
    fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
    fe u2 = fe_mul(p.X, p.Y);
    u32 dummy;
    fe I = fe_invsqrt(&dummy, fe_mul(u1, fe_sq(u2)));
    fe i1 = fe_mul(I, u1), i2 = fe_mul(I, u2);
    fe z_inv = fe_mul(i1, fe_mul(i2, p.T));
    fe iX = fe_mul(p.X, FE_SQRT_M1), iY = fe_mul(p.Y, FE_SQRT_M1);
    fe ench = fe_mul(i1, FE_INVSQRT_A_MINUS_D);
    u32 rotate = fe_is_negative(fe_mul(p.T, z_inv));
    fe X = fe_select(p.X, iY, rotate), Y = fe_select(p.Y, iX, rotate);
    fe den_inv = fe_select(i2, ench, rotate);
    Y = fe_cneg(Y, fe_is_negative(fe_mul(X, z_inv)));
    fe s = fe_abs(fe_mul(den_inv, fe_sub(p.Z, Y)));
    fe_to_words(w, s);
}
// ---- batched double-and-encode ----------------------------------------------------------------------
// encode(2P) without a square root (the technique of dalek's RistrettoPoint::double_and_compress_batch):
// with e = 2XY, f = Z^2 + dT^2, g = Y^2 + X^2, h = Z^2 - dT^2 the encoding of 2P needs only 1/(e g f h), and
// inverses batch (Montgomery's trick: one field inversion per batch + 3 multiplications per point).
// A caller that wants encode(Q) computes P = Q/2 by halving its scalars (sc_half).
// e*g*f*h = 0 exactly when P is 4-torsion, i.e. 2P is in the identity coset, whose encoding is 32 zero bytes.
struct ge_dbl_enc { fe e, f, g, h, eg, fh; };
ACT_FN ge_dbl_enc ge_dbl_enc_prepare(const ge& P) {
    ge_dbl_enc s;
    fe XX = fe_sq(P.X), YY = fe_sq(P.Y), ZZ = fe_sq(P.Z);
    fe dTT = fe_mul(fe_sq(P.T), FE_D);
    s.e = fe_mul(P.X, fe_add(P.Y, P.Y));
    s.f = fe_add(ZZ, dTT); s.g = fe_add(YY, XX); s.h = fe_sub(ZZ, dTT);
    s.eg = fe_mul(s.e, s.g); s.fh = fe_mul(s.f, s.h);
    return s;
}
// inv = 1/(eg*fh); writes the 8 canonical words of encode(2P)
ACT_FN void ge_dbl_enc_finish(u32* w, const ge_dbl_enc& s, const fe& inv) {
    fe Zinv = fe_mul(s.eg, inv), Tinv = fe_mul(s.fh, inv);
    u32 neg1 = fe_is_negative(fe_mul(s.eg, Zinv));
    fe e = fe_select(s.e, s.g, neg1);
    fe g = fe_select(s.g, fe_neg(s.e), neg1);
    fe h = fe_select(s.h, fe_mul(s.f, FE_SQRT_M1), neg1);
    fe magic = fe_select(FE_INVSQRT_A_MINUS_D, FE_SQRT_M1, neg1);
    u32 neg2 = fe_is_negative(fe_mul(fe_mul(h, e), Zinv));
    g = fe_cneg(g, neg2);
    fe sres = fe_abs(fe_mul(fe_sub(h, g), fe_mul(magic, fe_mul(g, Tinv))));
    fe_to_words(w, sres);
}

// RFC 9496 4.3.4 MAP / dalek elligator_ristretto_flavor
ACT_NOINLINE void ristretto_elligator_(ge* out, const fe* r0p) {
    fe r0 = *r0p;
    fe one = fe_one(), minus_one = fe_neg(fe_one());
    fe r = fe_mul(FE_SQRT_M1, fe_sq(r0));
    fe u = fe_mul(fe_add(r, one), FE_ONE_MINUS_D_SQ);
    fe v = fe_mul(fe_sub(minus_one, fe_mul(r, FE_D)), fe_add(r, FE_D));
    u32 was_square;
    fe s = fe_sqrt_ratio_i(&was_square, u, v);
    fe s_prime = fe_mul(s, r0);
    s_prime = fe_cneg(s_prime, fe_is_negative(s_prime) ^ 1u);
    s = fe_select(s_prime, s, was_square);
    fe c = fe_select(r, minus_one, was_square);
    fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), FE_D_MINUS_ONE_SQ), v);
    fe ss = fe_sq(s);
    fe w0 = fe_mul(fe_add(s, s), v);
    fe w1 = fe_mul(N, FE_SQRT_AD_MINUS_ONE);
    fe w2 = fe_sub(one, ss), w3 = fe_add(one, ss);
    out->X = fe_mul(w0, w3); out->Y = fe_mul(w2, w1); out->Z = fe_mul(w1, w3); out->T = fe_mul(w0, w2);
}
// RistrettoPoint::from_uniform_bytes: 16 little-endian words
ACT_FN ge ristretto_from_uniform(const u32* w) {
    fe r1 = fe_from_words(w), r2 = fe_from_words(w + 8);
    ge P1, P2;
    ristretto_elligator_(&P1, &r1);
    ristretto_elligator_(&P2, &r2);
    return ge_add(P1, P2);
}
// `point == RistrettoPoint::identity()` (src/lib.rs:787): X1*Y2 == Y1*X2 || X1*X2 == Y1*Y2 with (0,1)
ACT_FN u32 ristretto_is_identity(const ge& p) { return fe_is_zero(p.X) | fe_is_zero(p.Y); }
