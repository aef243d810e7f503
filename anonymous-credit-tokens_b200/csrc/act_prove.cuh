// act_prove.cuh -- client-side batch generators: PreIssuance::request
// (/root/reference src/lib.rs:463-487) and CreditToken::prove_spend (:972-1152, bits_of :902-915), as per-thread
// bodies of CUDA kernels.  They exist to synthesise full-size batches of VALID, UNIQUE requests and spend proofs on
// the device (a CPU cannot produce 10^6 proofs in bench time) and are bit-exact with the reference's prover for the
// same RNG bytes; they are NOT constant time (fixture grade: every scalar here is the client's own secret on the
// client's own device -- do not use them as a hardened wallet).
//
// The prover knows the openings of its commitments, so every point it needs is a fixed-base combination:
//   com_j       = h1*i_j + h3*s_j (+ h2*k* for j = 0)                                       (:1000-1004)
//   real branch = h3*s'_j (+ h2*k0')                                                          (:1026-1050)
//   simulated   = h3*z_j (+ h2*w0) - C_{j,b}*gamma_j  with  C_{j,b} = com_j - b*h1
//               = h3*(z_j - gamma_j s_j) + h1*(b - i_j)*gamma_j (+ h2*(w0 - gamma_j k*))      same group element
// and only A' = A*(r1 r2) and the A-part of A1 need a variable-base multiplication.  All 384 range-proof points are
// produced as halves and encoded by the batched square-root-free double-and-encode stage of the verifier.
//
// RNG: scalar number t of proof p is the wide reduction of 64 bytes taken either from an explicit byte stream
// (n x 524 x 64 bytes, consumed in the reference's order -- SURVEY Appendix B -- so that a run can be compared with the
// reference prover on identical bytes) or from BLAKE3-XOF(seed || u64le(first_index + p)), output block t.
#pragma once
#include "act_device.cuh"

#define ACT_PROVE_SCALARS 524
#define ACT_PROVE_PTS (3 * ACT_L)      // com[128], C'[128][2]
#define ACT_PROVE_PARTS (ACT_PROVE_PTS / ACT_ENC_BATCH)
// positions in the RNG stream (src/lib.rs:978-984, 998-999, 1010-1023, 1057-1058)
#define PR_R1 0
#define PR_R2 1
#define PR_CP 2
#define PR_RP 3
#define PR_EP 4
#define PR_R2P 5
#define PR_R3P 6
#define PR_KSTAR 7
#define PR_SI 8
#define PR_K0P 136
#define PR_SIP 137
#define PR_GI 265
#define PR_W0 393
#define PR_Z 394
#define PR_KP 522
#define PR_SP 523

struct prove_rng {
    const u32* rnd;      // explicit stream (n x 524 x 16 words) or null
    u32 seed[8];         // used when rnd is null
    u64 first_index;
};
ACT_FN sc prove_scalar(const prove_rng* R, size_t p, u32 t) {
    u32 w[16];
    if (R->rnd) {
        const u32* q = R->rnd + ((size_t)ACT_PROVE_SCALARS * p + t) * 16;
        load8(w, q); load8(w + 8, q + 8);
    } else {
        u32 m[16];
        u64 idx = R->first_index + p;
        ACT_UNROLL for (int i = 0; i < 8; i++) m[i] = R->seed[i];
        m[8] = (u32)idx; m[9] = (u32)(idx >> 32);
        ACT_UNROLL for (int i = 10; i < 16; i++) m[i] = 0;
        b3_compress(B3_IV_, m, t, 0, 40, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT, w);
    }
    return sc_from_wide(w);
}
// token record (160 B): A | e | k | r | c   (CreditToken, src/lib.rs:393-411)
ACT_FN u32 prove_bit(const u32* token, const u32* charge, int j) {
    sc m = sc_sub(load_scalar(token + 32), load_scalar(charge));      // c - s
    return (m.v[j >> 5] >> (j & 31)) & 1u;                            // bits_of (:902-915), j < 128
}
ACT_FN void store_ge(u32* q, const ge& P) { store_fe(q, P.X); store_fe(q + 8, P.Y); store_fe(q + 16, P.Z); store_fe(q + 24, P.T); }

// ---- stage 1: thread (p, j) -> halves of com_j, C'_j0, C'_j1 into cpts[p][384] -------------------------------------
ACT_FN void prove_range_thread(const act_ctx* C, const prove_rng* R, size_t p, size_t gp /* index into tokens */, int j,
                               const u32* tokens, const u32* charges, u32* cpts) {
    const u32* tk = tokens + 40 * gp;
    u32 bit = prove_bit(tk, charges + 8 * gp, j);
    sc sj = prove_scalar(R, p, PR_SI + j), sjp = prove_scalar(R, p, PR_SIP + j);
    sc gj = prove_scalar(R, p, PR_GI + j), zj = prove_scalar(R, p, PR_Z + j);
    u32* cp = cpts + (size_t)ACT_PROVE_PTS * 32 * p;
    sc one_or_zero = sc_from_u32(bit);
    // com_j / 2                                                                              (:1000-1004)
    ge Q = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H3], sc_half(sj), false);
    Q = fb_accumulate(Q, C->fb[ACT_BASE_H1], sc_half(one_or_zero), false);
    sc kstar = sc_zero();
    if (j == 0) { kstar = prove_scalar(R, p, PR_KSTAR); Q = fb_accumulate(Q, C->fb[ACT_BASE_H2], sc_half(kstar), false); }
    store_ge(cp + 32 * j, Q);
    // the real branch: h3*s'_j (+ h2*k0')                                                    (:1026-1050, second/first arm)
    ge Rl = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H3], sc_half(sjp), false);
    if (j == 0) Rl = fb_accumulate(Rl, C->fb[ACT_BASE_H2], sc_half(prove_scalar(R, p, PR_K0P)), false);
    // the simulated branch b = 1 - i_j: h3*(z_j - gamma_j s_j) + h1*(b - i_j)*gamma_j (+ h2*(w0 - gamma_j k*))
    ge Sm = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H3], sc_half(sc_sub(zj, sc_mul(gj, sj))), false);
    Sm = fb_accumulate(Sm, C->fb[ACT_BASE_H1], sc_half(gj), bit != 0);   // i_j = 0: +gamma_j h1, i_j = 1: -gamma_j h1
    if (j == 0) Sm = fb_accumulate(Sm, C->fb[ACT_BASE_H2], sc_half(sc_sub(prove_scalar(R, p, PR_W0), sc_mul(gj, kstar))), false);
    // conditional_select(a, b, i_j == 0): C'[j][0] = i_j == 0 ? real : sim, C'[j][1] = i_j == 0 ? sim : real
    store_ge(cp + 32 * (ACT_L + 2 * j + (bit ? 1 : 0)), Rl);
    store_ge(cp + 32 * (ACT_L + 2 * j + (bit ? 0 : 1)), Sm);
}

// ---- stage 2: thread p -> A', B-bar, A1, A2, C (items 1..4, 389), k (item 0), r3 = r1^-1 ---------------------------
ACT_FN void prove_head_thread(const act_ctx* C, const prove_rng* R, size_t p, size_t gp, const u32* tokens, u32* items, u32* aux, u8* status) {
    const u32* tk = tokens + 40 * gp;
    u32* it = items + (size_t)ACT_ITEM_WORDS * p;
    ge A;
    u32 valid = load_point(&A, tk);
    status[gp] = valid ? (u8)ACT_ST_OK : (u8)ACT_ST_DECODE_INVALID_POINT;
    sc k = load_scalar(tk + 16), r = load_scalar(tk + 24), c = load_scalar(tk + 32);
    sc r1 = prove_scalar(R, p, PR_R1), r2 = prove_scalar(R, p, PR_R2), cp = prove_scalar(R, p, PR_CP), rp = prove_scalar(R, p, PR_RP);
    sc ep = prove_scalar(R, p, PR_EP), r2p = prove_scalar(R, p, PR_R2P), r3p = prove_scalar(R, p, PR_R3P);
    sc r12 = sc_mul(r1, r2);
    store_scalar(it, k);                                                                    // transcript.add_scalar(&self.k)
    vb_table tA;
    vb_table_build(&tA, A);
    {
        ge Ap = vb_mul(&tA, r12, false);                                                    // a_prime = a * (r1 r2)      (:990)
        store_point(it + 8, Ap);
    }
    {
        // b_bar = (G + h1 c + h2 k + h3 r) * r1                                            (:986-991)
        ge Bb = fb_accumulate(ge_identity(), C->fb[ACT_BASE_G], r1, false);
        Bb = fb_accumulate(Bb, C->fb[ACT_BASE_H1], sc_mul(c, r1), false);
        Bb = fb_accumulate(Bb, C->fb[ACT_BASE_H2], sc_mul(k, r1), false);
        Bb = fb_accumulate(Bb, C->fb[ACT_BASE_H3], sc_mul(r, r1), false);
        store_point(it + 16, Bb);
    }
    {
        // a1 = a_prime*e' + b_bar*r2'                                                      (:993)
        sc t = sc_mul(r1, r2p);
        ge A1 = vb_mul(&tA, sc_mul(r12, ep), false);
        A1 = fb_accumulate(A1, C->fb[ACT_BASE_G], t, false);
        A1 = fb_accumulate(A1, C->fb[ACT_BASE_H1], sc_mul(c, t), false);
        A1 = fb_accumulate(A1, C->fb[ACT_BASE_H2], sc_mul(k, t), false);
        A1 = fb_accumulate(A1, C->fb[ACT_BASE_H3], sc_mul(r, t), false);
        store_point(it + 24, A1);
    }
    {
        // a2 = b_bar*r3' + h1*c' + h3*r'                                                   (:994)
        sc t = sc_mul(r1, r3p);
        ge A2 = fb_accumulate(ge_identity(), C->fb[ACT_BASE_G], t, false);
        A2 = fb_accumulate(A2, C->fb[ACT_BASE_H1], sc_add(sc_mul(c, t), cp), false);
        A2 = fb_accumulate(A2, C->fb[ACT_BASE_H2], sc_mul(k, t), false);
        A2 = fb_accumulate(A2, C->fb[ACT_BASE_H3], sc_add(sc_mul(r, t), rp), false);
        store_point(it + 32, A2);
    }
    {
        // c_ = h1*(-c') + h2*k' + h3*s'                                                    (:1059)
        ge Cc = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H1], cp, true);
        Cc = fb_accumulate(Cc, C->fb[ACT_BASE_H2], prove_scalar(R, p, PR_KP), false);
        Cc = fb_accumulate(Cc, C->fb[ACT_BASE_H3], prove_scalar(R, p, PR_SP), false);
        store_point(it + 8 * 389, Cc);
    }
    store_scalar(aux + 8 * p, sc_invert(r1));                                               // r3 = r1^-1                 (:992)
}

// ---- stage 3: fold the 16 chunk CVs into the challenge gamma (the transcript of :1061-1070 is the verifier's) --------
ACT_FN void prove_challenge_thread(size_t p, const u32* cvs, u32* gammas) {
    u32 o[16];
    spend_fold_root(cvs + (size_t)ACT_SPEND_CHUNKS * p * 8, o);
    store_scalar(gammas + 8 * p, sc_from_wide(o));
}

// ---- stage 4: thread (p, j) -> responses (:1072-1122) into the proof record; thread j = 0 also the head scalars and the
//      PreRefund (k*, r*, m) (:1124-1128) -----------------------------------------------------------------------------
ACT_FN void prove_finish_thread(const act_ctx* C, const prove_rng* R, size_t p, size_t gp, int j, const u32* tokens, const u32* charges,
                                const u32* items, const u32* aux, const u32* gammas, const u8* status, u32* proofs, u32* prerefunds) {
    (void)C;
    const u32* tk = tokens + 40 * gp;
    const u32* it = items + (size_t)ACT_ITEM_WORDS * p;
    u32* pf = proofs + (size_t)ACT_PROOF_WORDS * gp;
    if (status[gp] != ACT_ST_OK) {   // undecodable token: zero-filled outputs
        store8_zero(pf + 8 * (4 + j)); store8_zero(pf + 8 * (140 + j)); store8_zero(pf + 8 * (268 + 2 * j)); store8_zero(pf + 8 * (269 + 2 * j));
        if (j == 0) {
            const int idx[14] = {0, 1, 2, 3, 132, 133, 134, 135, 136, 137, 138, 139, 524, 525};
            ACT_NOUNROLL for (int i = 0; i < 14; i++) store8_zero(pf + 8 * idx[i]);
            ACT_NOUNROLL for (int i = 0; i < 3; i++) store8_zero(prerefunds + 24 * gp + 8 * i);
        }
        return;
    }
    u32 bit = prove_bit(tk, charges + 8 * gp, j);
    sc gamma; load8_rw(gamma.v, gammas + 8 * p);
    sc sj = prove_scalar(R, p, PR_SI + j), sjp = prove_scalar(R, p, PR_SIP + j);
    sc gj = prove_scalar(R, p, PR_GI + j), zj = prove_scalar(R, p, PR_Z + j);
    bool is0 = (bit == 0);
    sc g00 = is0 ? sc_sub(gamma, gj) : gj;                                                   // (:1077-1081, 1101-1105)
    sc g01 = sc_sub(gamma, g00);
    sc a = sc_add(sc_mul(g00, sj), sjp), b2 = sc_add(sc_mul(g01, sj), sjp);
    u32 w[8];
    load8_rw(w, it + 8 * (5 + j)); store8(pf + 8 * (4 + j), w);                              // com[j]
    store_scalar(pf + 8 * (140 + j), g00);
    store_scalar(pf + 8 * (268 + 2 * j), is0 ? a : zj);                                      // z[j][0]   (:1092-1096, 1106-1110)
    store_scalar(pf + 8 * (269 + 2 * j), is0 ? zj : b2);                                     // z[j][1]   (:1097-1100, 1111-1115)
    if (j != 0) return;
    sc e = load_scalar(tk + 8), k = load_scalar(tk + 16), r = load_scalar(tk + 24), c = load_scalar(tk + 32), s = load_scalar(charges + 8 * gp);
    sc ng = sc_neg(gamma);
    sc kstar = prove_scalar(R, p, PR_KSTAR), k0p = prove_scalar(R, p, PR_K0P), w0 = prove_scalar(R, p, PR_W0);
    sc r3; load8_rw(r3.v, aux + 8 * p);
    store_scalar(pf, k); store_scalar(pf + 8, s);
    load8_rw(w, it + 8); store8(pf + 16, w);                                                 // a_prime
    load8_rw(w, it + 16); store8(pf + 24, w);                                                // b_bar
    store_scalar(pf + 8 * 132, gamma);
    store_scalar(pf + 8 * 133, sc_add(sc_mul(ng, e), prove_scalar(R, p, PR_EP)));            // e_bar   (:1072)
    store_scalar(pf + 8 * 134, sc_add(sc_mul(gamma, prove_scalar(R, p, PR_R2)), prove_scalar(R, p, PR_R2P)));   // r2_bar (:1073)
    store_scalar(pf + 8 * 135, sc_add(sc_mul(gamma, r3), prove_scalar(R, p, PR_R3P)));       // r3_bar  (:1074)
    store_scalar(pf + 8 * 136, sc_add(sc_mul(ng, c), prove_scalar(R, p, PR_CP)));            // c_bar   (:1075)
    store_scalar(pf + 8 * 137, sc_add(sc_mul(ng, r), prove_scalar(R, p, PR_RP)));            // r_bar   (:1076)
    sc wa = sc_add(sc_mul(g00, kstar), k0p), wb = sc_add(sc_mul(g01, kstar), k0p);
    store_scalar(pf + 8 * 138, is0 ? wa : w0);                                               // w00     (:1082-1086)
    store_scalar(pf + 8 * 139, is0 ? w0 : wb);                                               // w01     (:1087-1091)
    // r* = sum s_i 2^i (:1052-1056), Horner from the top
    sc rstar = sc_zero();
    ACT_NOUNROLL for (int i = ACT_L - 1; i >= 0; i--) { rstar = sc_add(rstar, rstar); rstar = sc_add(rstar, prove_scalar(R, p, PR_SI + i)); }
    store_scalar(pf + 8 * 524, sc_add(sc_mul(gamma, kstar), prove_scalar(R, p, PR_KP)));     // k_bar   (:1121)
    store_scalar(pf + 8 * 525, sc_add(sc_mul(gamma, rstar), prove_scalar(R, p, PR_SP)));     // s_bar   (:1122)
    u32* pr = prerefunds + 24 * gp;
    store_scalar(pr, kstar); store_scalar(pr + 8, rstar); store_scalar(pr + 16, sc_sub(c, s));  // PreRefund {k, r, m}
}

// ---- PreIssuance::request (src/lib.rs:463-487).  pre: n x 16 words (r | k), rnd: n x 32 words (k'_wide | r'_wide) ------
ACT_FN void request_thread(const act_ctx* C, size_t i, const u32* pre, const u32* rnd, u32* req) {
    sc r = load_scalar(pre + 16 * i), k = load_scalar(pre + 16 * i + 8);
    u32 w[32];
    ACT_NOUNROLL for (int q = 0; q < 4; q++) load8(w + 8 * q, rnd + 32 * i + 8 * q);
    sc kp = sc_from_wide(w), rp = sc_from_wide(w + 16);
    ge K = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H2], k, false);
    K = fb_accumulate(K, C->fb[ACT_BASE_H3], r, false);
    ge K1 = fb_accumulate(ge_identity(), C->fb[ACT_BASE_H2], kp, false);
    K1 = fb_accumulate(K1, C->fb[ACT_BASE_H3], rp, false);
    tr_small tr;
    u32 kw[8], k1w[8];
    ristretto_encode_(kw, &K); ristretto_encode_(k1w, &K1);
    tr_init(&tr, C, ACT_TR_REQUEST);
    tr_add32(&tr, kw); tr_add32(&tr, k1w);
    sc gamma = tr_challenge(&tr);
    u32* out = req + 32 * i;
    store8(out, kw); store_scalar(out + 8, gamma);
    store_scalar(out + 16, sc_add(kp, sc_mul(k, gamma)));
    store_scalar(out + 24, sc_add(rp, sc_mul(r, gamma)));
}
