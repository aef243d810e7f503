// blake3.cuh -- BLAKE3 compression function and the two shapes the Fiat-Shamir transcript needs:
// a single-chunk message (<= 1024 B: "request" 266 B, "respond" 466 B, "refund" 425 B, Params::new)
// and the 16-chunk "spend" transcript (15 784 B) hashed chunk-parallel.
//
// Replaces blake3::Hasher::{new, update, finalize, finalize_xof().fill(64)} as used at
// /root/reference src/transcript.rs:56-71,96-97,150-152 and src/lib.rs:299-303,333-351.
#pragma once
#include "fe25519.cuh"

#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_PARENT 4u
#define B3_ROOT 8u

ACT_CONST u32 B3_IV_[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

ACT_FN u32 b3_rotr(u32 x, int n) {
#if ACT_PTX
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

#define B3_G(a, b, c, d, x, y)                      \
    do {                                            \
        a = a + b + (x); d = b3_rotr(d ^ a, 16);    \
        c = c + d;       b = b3_rotr(b ^ c, 12);    \
        a = a + b + (y); d = b3_rotr(d ^ a, 8);     \
        c = c + d;       b = b3_rotr(b ^ c, 7);     \
    } while (0)

#define B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    B3_G(s0, s4, s8, s12, m0, m1);   B3_G(s1, s5, s9, s13, m2, m3);                     \
    B3_G(s2, s6, s10, s14, m4, m5);  B3_G(s3, s7, s11, s15, m6, m7);                    \
    B3_G(s0, s5, s10, s15, m8, m9);  B3_G(s1, s6, s11, s12, m10, m11);                  \
    B3_G(s2, s7, s8, s13, m12, m13); B3_G(s3, s4, s9, s14, m14, m15)

// out[0..15]: full 16-word output of the compression function.  The message schedule is the fixed
// BLAKE3 permutation applied 0..6 times, written out so that m[] stays in registers.
ACT_FN void b3_compress(const u32* cv, const u32* m, u32 ctr_lo, u32 ctr_hi, u32 blen, u32 flags, u32* out) {
    u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    u32 s8 = 0x6A09E667u, s9 = 0xBB67AE85u, s10 = 0x3C6EF372u, s11 = 0xA54FF53Au;
    u32 s12 = ctr_lo, s13 = ctr_hi, s14 = blen, s15 = flags;
    B3_ROUND(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
    B3_ROUND(m[2], m[6], m[3], m[10], m[7], m[0], m[4], m[13], m[1], m[11], m[12], m[5], m[9], m[14], m[15], m[8]);
    B3_ROUND(m[3], m[4], m[10], m[12], m[13], m[2], m[7], m[14], m[6], m[5], m[9], m[0], m[11], m[15], m[8], m[1]);
    B3_ROUND(m[10], m[7], m[12], m[9], m[14], m[3], m[13], m[15], m[4], m[0], m[11], m[2], m[5], m[8], m[1], m[6]);
    B3_ROUND(m[12], m[13], m[9], m[11], m[15], m[10], m[14], m[8], m[7], m[2], m[5], m[3], m[0], m[1], m[6], m[4]);
    B3_ROUND(m[9], m[14], m[11], m[5], m[8], m[12], m[15], m[1], m[13], m[3], m[0], m[10], m[2], m[6], m[4], m[7]);
    B3_ROUND(m[11], m[15], m[5], m[0], m[1], m[9], m[8], m[6], m[14], m[10], m[2], m[12], m[3], m[4], m[7], m[13]);
    out[0] = s0 ^ s8;  out[1] = s1 ^ s9;  out[2] = s2 ^ s10; out[3] = s3 ^ s11;
    out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
    out[8] = s8 ^ cv[0];  out[9] = s9 ^ cv[1];   out[10] = s10 ^ cv[2]; out[11] = s11 ^ cv[3];
    out[12] = s12 ^ cv[4]; out[13] = s13 ^ cv[5]; out[14] = s14 ^ cv[6]; out[15] = s15 ^ cv[7];
}

// Hash of a message of nbytes <= 1024 held as little-endian words (zero padded to a whole block),
// first 16 output words of the XOF (= finalize_xof().fill(64); the first 8 are finalize()).
ACT_NOINLINE void b3_hash_single_chunk(const u32* msg, u32 nbytes, u32* out16) {
    u32 cv[8], o[16];
    ACT_UNROLL for (int i = 0; i < 8; i++) cv[i] = B3_IV_[i];
    u32 nblocks = (nbytes + 63u) / 64u;
    if (nblocks == 0) nblocks = 1;
    ACT_NOUNROLL for (u32 b = 0; b < nblocks; b++) {
        u32 flags = (b == 0 ? B3_CHUNK_START : 0u);
        u32 blen = 64;
        if (b == nblocks - 1) { flags |= B3_CHUNK_END | B3_ROOT; blen = nbytes - 64u * b; }
        b3_compress(cv, msg + 16 * b, 0, 0, blen, flags, o);
        ACT_UNROLL for (int i = 0; i < 8; i++) cv[i] = o[i];
    }
    ACT_UNROLL for (int i = 0; i < 16; i++) out16[i] = o[i];
}
