// TEST PROGRAM for include/act.hpp (the C++ host-side mirror of the reference's interface): reads a fixture written by
// tests/test_cpp_host.py, drives the engine through act::Engine with ONE shared RNG stream, writes every output byte back.
// usage: act_hpp_check <fixture.bin> <out.bin> [device0,device1]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

#include "act.hpp"

struct StreamRng {   // the role of the reference's `impl CryptoRngCore`: hands out the bytes of a recorded stream
    const std::vector<uint8_t>& s;
    size_t pos = 0;
    void fill_bytes(uint8_t* dst, size_t n) {
        if (pos + n > s.size()) throw std::runtime_error("rng stream exhausted");
        std::memcpy(dst, s.data() + pos, n);
        pos += n;
    }
};
static uint32_t rd32(const std::vector<uint8_t>& b, size_t& o) { uint32_t v; std::memcpy(&v, b.data() + o, 4); o += 4; return v; }
static void wr(std::vector<uint8_t>& o, const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; o.insert(o.end(), q, q + n); }
static void wr32(std::vector<uint8_t>& o, uint32_t v) { wr(o, &v, 4); }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s fixture.bin out.bin [devices]\n", argv[0]); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t o = 0;
    act::Params params;
    std::memcpy(params.h.data(), b.data(), 96); o = 96;
    act::Scalar x; act::Bytes32 w;
    std::memcpy(x.data(), b.data() + o, 32); std::memcpy(w.data(), b.data() + o + 32, 32); o += 64;
    uint32_t nr = rd32(b, o);
    std::vector<act::IssuanceRequest> reqs(nr);
    std::vector<act::Scalar> cs(nr);
    for (uint32_t i = 0; i < nr; i++) { std::memcpy(reqs[i].bytes.data(), b.data() + o, ACT_REQUEST_BYTES); o += ACT_REQUEST_BYTES; }
    for (uint32_t i = 0; i < nr; i++) { std::memcpy(cs[i].data(), b.data() + o, 32); o += 32; }
    uint32_t np = rd32(b, o);
    std::vector<act::SpendProof> proofs(np);
    for (uint32_t i = 0; i < np; i++) { std::memcpy(proofs[i].bytes.data(), b.data() + o, ACT_PROOF_BYTES); o += ACT_PROOF_BYTES; }
    uint32_t nrng = rd32(b, o);
    std::vector<uint8_t> stream(b.begin() + o, b.begin() + o + nrng);

    try {
        act::PrivateKey key(x, w);
        if (act::PrivateKey::from_secret(x).public_key().w != w) throw std::runtime_error("PrivateKey::from_secret: W differs");
        std::vector<int> devices;
        if (argc > 3) { std::stringstream ss(argv[3]); std::string t; while (std::getline(ss, t, ',')) devices.push_back(std::atoi(t.c_str())); }
        act::Engine eng = devices.empty() ? act::Engine(params, key, 0) : act::Engine(params, key, devices);
        StreamRng rng{stream};
        std::vector<uint8_t> out;
        // CBOR round trip through the wire types, as a client of the crate would hand requests over
        for (auto& r : reqs) {
            std::vector<uint8_t> c = r.to_cbor();
            auto back = act::IssuanceRequest::from_cbor(c.data(), c.size());
            if (!back.ok() || back.value.bytes != r.bytes) throw std::runtime_error("IssuanceRequest CBOR round trip");
        }
        auto issued = eng.batch_issue(reqs, cs, rng);
        for (auto& r : issued) out.push_back(r.status);
        for (auto& r : issued) wr(out, r.value.bytes.data(), ACT_RESPONSE_BYTES);
        wr32(out, (uint32_t)rng.pos);
        std::vector<act::IssuanceResponse> resps(nr);
        for (uint32_t i = 0; i < nr; i++) resps[i] = issued[i].value;
        auto ic = eng.batch_issuance_check(reqs, resps);
        wr(out, ic.data(), ic.size());
        auto refunded = eng.batch_verify_spend_and_refund(proofs, rng);
        for (auto& r : refunded) out.push_back(r.status);
        for (auto& r : refunded) wr(out, r.value.nullifier.data(), 32);
        for (auto& r : refunded) wr(out, r.value.refund.bytes.data(), ACT_REFUND_BYTES);
        wr32(out, (uint32_t)rng.pos);
        std::vector<act::Refund> rf(np);
        for (uint32_t i = 0; i < np; i++) rf[i] = refunded[i].value.refund;
        auto rc = eng.batch_refund_check(proofs, rf);
        wr(out, rc.data(), rc.size());
        // Result<T, Error> behaves like the crate's: unwrap() on an Err throws, error() names the variant
        size_t errs = 0;
        for (auto& r : refunded) if (!r.ok()) { errs++; try { r.unwrap(); return 3; } catch (const std::runtime_error&) {} (void)act::to_string(r.error()); }
        std::ofstream g(argv[2], std::ios::binary);
        g.write((const char*)out.data(), (std::streamsize)out.size());
        std::printf("act_hpp_check ok: %u requests, %u proofs (%zu rejected), %zu rng bytes drawn\n", nr, np, errs, rng.pos);
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "act_hpp_check FAILED: %s\n", ex.what());
        return 1;
    }
    return 0;
}
