// hostsim.cpp -- TEST-ONLY host build of the device headers.
//
// Compiles anonymous-credit-tokens_b200/csrc/*.cuh with g++ (portable C++ paths, no PTX) and runs the
// per-thread kernel bodies in plain loops, so that the arithmetic and protocol LOGIC of the CUDA
// engine can be unit-tested against the oracle on a machine without a GPU.  It is never linked into
// libact_b200.so and is not a product path: the shipped library has no CPU fallback.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../anonymous-credit-tokens_b200/csrc/act_device.cuh"
#include "../../anonymous-credit-tokens_b200/csrc/act_aux.cuh"
#include "../../anonymous-credit-tokens_b200/csrc/act_prove.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

struct hs_ctx {
    act_ctx c;
    std::vector<ge_niels> tables;
};

static void build_prefix(act_ctx* c, int which, const char* label, const uint8_t h[96]) {
    static const char ver[] = "curve25519-ristretto anonymous-credits v1.0";
    uint8_t buf[192];
    memset(buf, 0, sizeof buf);
    size_t n = 0;
    buf[7] = 43; n = 8;
    memcpy(buf + n, ver, 43); n += 43;
    for (int i = 0; i < 3; i++) { buf[n + 7] = 32; n += 8; memcpy(buf + n, h + 32 * i, 32); n += 32; }
    size_t ll = strlen(label);
    buf[n + 7] = (uint8_t)ll; n += 8;
    memcpy(buf + n, label, ll); n += ll;
    memcpy(c->prefix[which], buf, 192);
    c->prefix_len[which] = (u32)n;
}

EXPORT hs_ctx* hs_ctx_create(const uint8_t h[96], const uint8_t x[32], const uint8_t w[32]) {
    hs_ctx* H = new hs_ctx();
    memset(&H->c, 0, sizeof H->c);
    ge bases[ACT_FB_BASES];
    bases[0] = ge_basepoint();
    u32 ok = 1, words[8];
    for (int i = 0; i < 3; i++) { memcpy(words, h + 32 * i, 32); ok &= ristretto_decode_(&bases[1 + i], words); }
    memcpy(words, w, 32);
    ok &= ristretto_decode_(&H->c.W, words);
    if (!ok) { delete H; return nullptr; }
    bases[ACT_BASE_W] = H->c.W;
    // host build: every base at the narrow width (the wide 2^16 tables of the device build would take seconds to make on a CPU);
    // base 3 gets a different width than the others so that the per-table width logic is exercised
    const u32 bits[ACT_FB_BASES] = {ACT_FB_BITS, ACT_FB_BITS, ACT_FB_BITS, 11, ACT_FB_BITS};
    size_t off[ACT_FB_BASES + 1] = {0};
    for (int b = 0; b < ACT_FB_BASES; b++) off[b + 1] = off[b] + fb_size_of(bits[b]);
    H->tables.resize(off[ACT_FB_BASES] + ACT_CT_SIZE);
    for (int b = 0; b < ACT_FB_BASES; b++)
        for (int win = 0; win < (int)fb_win_of(bits[b]); win++)
            for (int part = 0; part < (int)fb_parts_of(bits[b]); part++) build_fb_table_thread(&bases[b], bits[b], win, part, H->tables.data() + off[b]);
    for (int win = 0; win < ACT_CT_WIN; win++) build_table_thread<4, ACT_CT_ENT>(&bases[0], win, H->tables.data() + off[ACT_FB_BASES]);
    for (int b = 0; b < ACT_FB_BASES; b++) { H->c.fb[b].p = H->tables.data() + off[b]; H->c.fb[b].bits = bits[b]; H->c.fb[b].win = fb_win_of(bits[b]); H->c.fb[b].ent = fb_ent_of(bits[b]); }
    H->c.ct_g = H->tables.data() + off[ACT_FB_BASES];
    memcpy(H->c.h_enc, h, 96);
    build_prefix(&H->c, ACT_TR_REQUEST, "request", h);
    build_prefix(&H->c, ACT_TR_RESPOND, "respond", h);
    build_prefix(&H->c, ACT_TR_REFUND, "refund", h);
    build_prefix(&H->c, ACT_TR_SPEND, "spend", h);
    memcpy(H->c.x.v, x, 32);
    ctx_finalize_thread(&H->c);
    return H;
}
EXPORT void hs_ctx_destroy(hs_ctx* H) { delete H; }

EXPORT int hs_params_derive(const char* org, const char* svc, const char* dep, const char* ver, uint8_t h[96]) {
    std::string dom = std::string("ACT-v1:") + org + ":" + svc + ":" + dep + ":" + ver;
    if (dom.size() > 900) return -1;
    u32 out[24];
    params_derive_thread((const u8*)dom.data(), (u32)dom.size(), out);
    memcpy(h, out, 96);
    return 0;
}
EXPORT void hs_public_key(const hs_ctx* H, const uint8_t x[32], uint8_t w[32]) {
    u32 words[8], o[8];
    memcpy(words, x, 32);
    sc s = sc_from_words(words);
    ge W = fb_accumulate_ct(ge_identity(), H->c.ct_g, s);
    ristretto_encode_(o, &W);
    memcpy(w, o, 32);
}

// records are copied into word-aligned buffers first
static std::vector<u32> to_words(const uint8_t* p, size_t bytes) {
    std::vector<u32> v((bytes + 3) / 4 + 8);
    if (bytes) memcpy(v.data(), p, bytes);
    return v;
}

EXPORT void hs_issue(const hs_ctx* H, size_t n, const uint8_t* req, const uint8_t* cs, const uint8_t* rnd, uint8_t* resp, uint8_t* status) {
    auto r = to_words(req, n * 128), c = to_words(cs, n * 32), d = to_words(rnd, n * 128);
    std::vector<u32> o(n * 40 + 8);
    for (size_t i = 0; i < n; i++) issue_thread(&H->c, i, r.data(), c.data(), d.data(), o.data(), status);
    memcpy(resp, o.data(), n * 160);
}
EXPORT void hs_issuance_check(const hs_ctx* H, size_t n, const uint8_t* K, const uint8_t* resp, uint8_t* status) {
    auto k = to_words(K, n * 32), r = to_words(resp, n * 160);
    for (size_t i = 0; i < n; i++) issuance_check_thread(&H->c, i, k.data(), r.data(), status);
}
EXPORT void hs_refund(const hs_ctx* H, size_t n, const uint8_t* proofs, const uint8_t* rnd, uint8_t* refunds, uint8_t* nullifiers, uint8_t* status) {
    auto pf = to_words(proofs, n * (size_t)ACT_PROOF_WORDS * 4), d = to_words(rnd, n * 128);
    std::vector<u32> items(n * (size_t)ACT_ITEM_WORDS + 8), cn(n * (size_t)ACT_L * 24 + 8), kp(n * 32 + 8), flags(n + 1, 0),
        cvs(n * ACT_SPEND_CHUNKS * 8 + 8), ro(n * 32 + 8), no(n * 8 + 8);
    std::vector<vb_table> tabs(ACT_RANGE_SPLIT);
    std::vector<u32> cpts(n * 2 * ACT_L * 32 + 8);
    for (size_t p = 0; p < n; p++)
        for (int j = 0; j < ACT_L; j++) spend_range_thread(&H->c, p, j, pf.data(), items.data(), cn.data(), flags.data(), tabs.data(), cpts.data());
    for (size_t p = 0; p < n; p++) spend_head_thread(&H->c, p, pf.data(), items.data(), cn.data(), kp.data(), flags.data(), cpts.data());
    for (size_t p = 0; p < n; p++)
        for (int part = 0; part < 2 * ACT_L / ACT_ENC_BATCH; part++) spend_encode_thread(&H->c, p, part, cpts.data(), items.data());
    for (size_t p = 0; p < n; p++)
        for (int c = 0; c < ACT_SPEND_CHUNKS; c++) spend_chunk_thread(&H->c, p, c, items.data(), cvs.data());
    for (size_t p = 0; p < n; p++) spend_finish_thread(&H->c, p, pf.data(), cvs.data(), flags.data(), status);
    for (size_t p = 0; p < n; p++) refund_sign_thread(&H->c, p, pf.data(), d.data(), kp.data(), status, ro.data(), no.data());
    memcpy(refunds, ro.data(), n * 128);
    memcpy(nullifiers, no.data(), n * 32);
}
EXPORT void hs_refund_check(const hs_ctx* H, size_t n, const uint8_t* com, const uint8_t* refund, uint8_t* status) {
    auto c = to_words(com, n * 4096), r = to_words(refund, n * 128);
    for (size_t i = 0; i < n; i++) refund_check_thread(&H->c, i, c.data(), r.data(), status);
}

// ---- primitives ----
EXPORT void hs_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fe x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32);  // full 256-bit loose inputs allowed
    fe z = fe_mul(x, y); u32 w[8]; fe_to_words(w, z); memcpy(out, w, 32);
}
// raw words of a*b, a^2 (as the multiplication leaves them: must be "tight"), and of the tight sum / double of those two
EXPORT int hs_fe_tight(const uint8_t a[32], const uint8_t b[32], uint8_t prod[32], uint8_t sq[32], uint8_t sum[32], uint8_t dbl[32]) {
    fe x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    fe m = fe_mul(x, y), q = fe_sq(x);
    int tight = fe_is_tight_(m) && fe_is_tight_(q);
    fe sm = fe_add_tt(m, q), d = fe_dbl_tt(m);
    memcpy(prod, m.v, 32); memcpy(sq, q.v, 32); memcpy(sum, sm.v, 32); memcpy(dbl, d.v, 32);
    return tight;
}
EXPORT void hs_fe_addsub(const uint8_t a[32], const uint8_t b[32], uint8_t sum[32], uint8_t diff[32]) {
    fe x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    u32 w[8]; fe_to_words(w, fe_add(x, y)); memcpy(sum, w, 32); fe_to_words(w, fe_sub(x, y)); memcpy(diff, w, 32);
}
EXPORT void hs_fe_invert(const uint8_t a[32], uint8_t out[32]) {
    fe x; memcpy(x.v, a, 32); u32 w[8]; fe_to_words(w, fe_invert(x)); memcpy(out, w, 32);
}
EXPORT int hs_decode_encode(const uint8_t in[32], uint8_t out[32]) {
    u32 w[8], o[8]; memcpy(w, in, 32); ge P;
    if (!ristretto_decode_(&P, w)) return 0;
    ristretto_encode_(o, &P); memcpy(out, o, 32); return 1;
}
EXPORT void hs_from_uniform(const uint8_t in[64], uint8_t out[32]) {
    u32 w[16], o[8]; memcpy(w, in, 64); ge P = ristretto_from_uniform(w); ristretto_encode_(o, &P); memcpy(out, o, 32);
}
EXPORT int hs_scalarmult(const uint8_t s[32], const uint8_t P[32], int ct, uint8_t out[32]) {
    u32 w[8], o[8]; memcpy(w, P, 32); ge Q, R;
    if (!ristretto_decode_(&Q, w)) return 0;
    memcpy(w, s, 32); sc k = sc_from_words(w);
    if (ct) vb_mul_ct_(&R, &Q, &k);
    else { vb_table t; vb_table_build(&t, Q); R = vb_mul(&t, k, false); }
    ristretto_encode_(o, &R); memcpy(out, o, 32); return 1;
}
// base: 0..3 = G,H1,H2,H3 via the radix-256 tables; ct=1 (base 0 only) uses the constant-time table
EXPORT void hs_scalarmult_base(const hs_ctx* H, int base, const uint8_t s[32], int ct, uint8_t out[32]) {
    u32 w[8], o[8]; memcpy(w, s, 32); sc k = sc_from_words(w);
    ge R = ct ? fb_accumulate_ct(ge_identity(), H->c.ct_g, k) : fb_accumulate(ge_identity(), H->c.fb[base], k, false);
    ristretto_encode_(o, &R); memcpy(out, o, 32);
}
EXPORT void hs_sc_reduce32(const uint8_t in[32], uint8_t out[32]) { u32 w[8]; memcpy(w, in, 32); sc a = sc_from_words(w); memcpy(out, a.v, 32); }
EXPORT void hs_sc_reduce64(const uint8_t in[64], uint8_t out[32]) { u32 w[16]; memcpy(w, in, 64); sc a = sc_from_wide(w); memcpy(out, a.v, 32); }
EXPORT void hs_sc_muladd(const uint8_t a[32], const uint8_t b[32], const uint8_t c[32], uint8_t out[32]) {
    u32 w[8]; memcpy(w, a, 32); sc x = sc_from_words(w); memcpy(w, b, 32); sc y = sc_from_words(w); memcpy(w, c, 32); sc z = sc_from_words(w);
    sc r = sc_add(sc_mul(x, y), z); memcpy(out, r.v, 32);
}
EXPORT void hs_sc_invert(const uint8_t a[32], uint8_t out[32]) { u32 w[8]; memcpy(w, a, 32); sc r = sc_invert(sc_from_words(w)); memcpy(out, r.v, 32); }
EXPORT void hs_sc_negsub(const uint8_t a[32], const uint8_t b[32], uint8_t neg[32], uint8_t diff[32]) {
    u32 w[8]; memcpy(w, a, 32); sc x = sc_from_words(w); memcpy(w, b, 32); sc y = sc_from_words(w);
    sc n = sc_neg(x), d = sc_sub(x, y); memcpy(neg, n.v, 32); memcpy(diff, d.v, 32);
}
EXPORT void hs_blake3_small(const uint8_t* in, size_t n, uint8_t out[64]) {
    u32 buf[256]; memset(buf, 0, sizeof buf); memcpy(buf, in, n); u32 o[16];
    b3_hash_single_chunk(buf, (u32)n, o); memcpy(out, o, 64);
}
EXPORT void hs_recode(const uint8_t s[32], int w, int8_t* digits) {
    u32 ww[8]; memcpy(ww, s, 32); sc k = sc_from_words(ww);
    if (w == 4) { sc b = sc_bias<4>(k); for (int i = 0; i < 64; i++) digits[i] = (int8_t)sc_digit<4>(b, i); }
    else { sc b = sc_bias<8>(k); for (int i = 0; i < 32; i++) { int d = sc_digit<8>(b, i); digits[i] = (int8_t)d; } }
}

// runs the batched double-and-encode stage on 256 points given by their encodings: out[i] = encode(2 * P_i)
EXPORT int hs_encode_stage(const uint8_t* enc /* 256 x 32 */, uint8_t* out /* 256 x 32 */) {
    std::vector<u32> cpts(2 * ACT_L * 32 + 8), items(ACT_ITEM_WORDS + 8);
    for (int i = 0; i < 2 * ACT_L; i++) {
        u32 w[8]; memcpy(w, enc + 32 * i, 32); ge P;
        if (!ristretto_decode_(&P, w)) return 0;
        // randomise the projective representation a little: multiply by Z = i + 2
        fe z = fe_zero(); z.v[0] = (u32)i + 2;
        P.X = fe_mul(P.X, z); P.Y = fe_mul(P.Y, z); P.Z = fe_mul(P.Z, z); P.T = fe_mul(P.T, z);
        memcpy(&cpts[32 * i], P.X.v, 32); memcpy(&cpts[32 * i + 8], P.Y.v, 32); memcpy(&cpts[32 * i + 16], P.Z.v, 32); memcpy(&cpts[32 * i + 24], P.T.v, 32);
    }
    for (int part = 0; part < 2 * ACT_L / ACT_ENC_BATCH; part++) spend_encode_thread(nullptr, 0, part, cpts.data(), items.data());
    memcpy(out, &items[8 * 133], 256 * 32);
    return 1;
}
EXPORT void hs_sc_half(const uint8_t a[32], uint8_t out[32]) { u32 w[8]; memcpy(w, a, 32); sc r = sc_half(sc_from_words(w)); memcpy(out, r.v, 32); }

// canonical-CBOR skeletons of act_aux.cuh applied on the host (the device kernels do exactly this per byte)
EXPORT int hs_cbor_skeleton_encode(int kind, const uint8_t* rec, uint8_t* out) {
    size_t len = act_cbor_len(kind);
    std::vector<int32_t> sk(len);
    act_build_skeleton(sk.data(), kind);
    for (size_t i = 0; i < len; i++) out[i] = sk[i] >= 0 ? rec[sk[i]] : (uint8_t)(-(sk[i] + 1));
    return (int)len;
}
EXPORT int hs_cbor_skeleton_unpack(int kind, const uint8_t* cbor, uint8_t* rec) {
    size_t len = act_cbor_len(kind);
    std::vector<int32_t> sk(len);
    act_build_skeleton(sk.data(), kind);
    int bad = 0;
    for (size_t i = 0; i < len; i++) {
        if (sk[i] >= 0) rec[sk[i]] = cbor[i];
        else if (cbor[i] != (uint8_t)(-(sk[i] + 1))) bad = 1;
    }
    return bad ? 0xFF : 0;
}

// client-side generators (act_prove.cuh), the kernel bodies run in plain loops
EXPORT void hs_request(const hs_ctx* H, size_t n, const uint8_t* pre, const uint8_t* rnd, uint8_t* req) {
    auto p = to_words(pre, n * 64), r = to_words(rnd, n * 128);
    std::vector<u32> out(n * 32 + 8);
    for (size_t i = 0; i < n; i++) request_thread(&H->c, i, p.data(), r.data(), out.data());
    memcpy(req, out.data(), n * 128);
}
EXPORT void hs_prove_spend(const hs_ctx* H, size_t n, const uint8_t* tokens, const uint8_t* charges, const uint8_t* rnd, const uint8_t* seed,
                           uint64_t first_index, uint8_t* proofs, uint8_t* prerefunds, uint8_t* status) {
    auto tk = to_words(tokens, n * 160), ch = to_words(charges, n * 32);
    std::vector<u32> rw;
    if (rnd) rw = to_words(rnd, n * (size_t)ACT_PROVE_SCALARS * 64);
    prove_rng R;
    memset(&R, 0, sizeof R);
    R.rnd = rnd ? rw.data() : nullptr;
    if (seed) memcpy(R.seed, seed, 32);
    R.first_index = first_index;
    std::vector<u32> cpts(n * ACT_PROVE_PTS * 32 + 8), items(n * ACT_ITEM_WORDS + 8), cvs(n * ACT_SPEND_CHUNKS * 8 + 8), gam(n * 8 + 8), aux(n * 8 + 8);
    std::vector<u32> pf(n * ACT_PROOF_WORDS + 8), pr(n * 24 + 8);
    for (size_t p = 0; p < n; p++)
        for (int j = 0; j < ACT_L; j++) prove_range_thread(&H->c, &R, p, p, j, tk.data(), ch.data(), cpts.data());
    for (size_t p = 0; p < n; p++)
        for (int part = 0; part < ACT_PROVE_PARTS; part++) spend_encode_thread(&H->c, p, part, cpts.data(), items.data(), ACT_PROVE_PTS, 5);
    for (size_t p = 0; p < n; p++) prove_head_thread(&H->c, &R, p, p, tk.data(), items.data(), aux.data(), status);
    for (size_t p = 0; p < n; p++)
        for (int c = 0; c < ACT_SPEND_CHUNKS; c++) spend_chunk_thread(&H->c, p, c, items.data(), cvs.data());
    for (size_t p = 0; p < n; p++) prove_challenge_thread(p, cvs.data(), gam.data());
    for (size_t p = 0; p < n; p++)
        for (int j = 0; j < ACT_L; j++) prove_finish_thread(&H->c, &R, p, p, j, tk.data(), ch.data(), items.data(), aux.data(), gam.data(), status, pf.data(), pr.data());
    memcpy(proofs, pf.data(), n * (size_t)ACT_PROOF_WORDS * 4);
    memcpy(prerefunds, pr.data(), n * 96);
}
