// cbor_host.cpp -- host-side CBOR front doors: variable-length wire items <-> fixed records.
//
// Mirrors the structural rules of the reference codec (/root/reference src/cbor.rs): integer-keyed
// maps (:105-110, :250-268), 32-byte byte strings for every scalar and point (:62-91), arrays of
// exactly L entries for keys 5/14/15 (:305-378) where a non-array value is silently skipped, unknown
// keys ignored (:134, :381), duplicate keys: last wins, missing field -> InvalidStructure
// (:138-143, :385-403).  Point *validity* is left to the device (status 0x81).
//
// Known divergences (both only on malformed or non-canonical items):
//  * an item carrying BOTH an undecodable point and a later structural defect is reported as 0x82 here, while the
//    reference returns whichever error comes first in map order;
//  * duplicate keys: the last value wins here as there, but the reference decodes EVERY occurrence (`decode_point(&val)?`
//    inside the loop over the map), so an item whose earlier, overwritten occurrence of a point key holds an undecodable
//    point is an InvalidValue there and is accepted here (only the surviving value reaches the device's validity check).
#include <stdint.h>
#include <string.h>

#include "../../include/act_engine.h"

namespace {

enum { ST_OK = 0, ST_STRUCT = 0x82, ST_PARSE = 0x83 };
const int kMaxDepth = 256;

struct Reader {
    const uint8_t* p;
    size_t n, pos;
    bool need(size_t k) const { return pos + k <= n && pos + k >= pos; }
};

// reads the head of an item: major type, additional info, argument value.  false on truncation/reserved info.
bool read_head(Reader& r, int& major, int& info, uint64_t& arg) {
    if (!r.need(1)) return false;
    uint8_t b = r.p[r.pos++];
    major = b >> 5; info = b & 31; arg = 0;
    if (info < 24) { arg = (uint64_t)info; return true; }
    if (info == 31) return major >= 2 && major <= 5 ? true : (major == 7);  // indefinite length or "break"
    if (info > 27) return false;
    size_t k = (size_t)1 << (info - 24);
    if (!r.need(k)) return false;
    for (size_t i = 0; i < k; i++) arg = (arg << 8) | r.p[r.pos++];
    return true;
}
bool skip_item(Reader& r, int depth);
bool skip_indefinite(Reader& r, int major, int depth) {
    for (;;) {
        if (!r.need(1)) return false;
        if (r.p[r.pos] == 0xff) { r.pos++; return true; }
        if (major == 2 || major == 3) {
            // chunks must be definite-length strings of the same major type
            int m, info; uint64_t arg;
            if (!read_head(r, m, info, arg) || m != major || info == 31) return false;
            if (!r.need(arg)) return false;
            r.pos += (size_t)arg;
        } else if (major == 4) {
            if (!skip_item(r, depth + 1)) return false;
        } else {  // map: key, value
            if (!skip_item(r, depth + 1)) return false;
            if (!r.need(1) || r.p[r.pos] == 0xff) return false;
            if (!skip_item(r, depth + 1)) return false;
        }
    }
}
bool skip_item(Reader& r, int depth) {
    if (depth > kMaxDepth) return false;
    int major, info; uint64_t arg;
    if (!read_head(r, major, info, arg)) return false;
    switch (major) {
    case 0: case 1: return info != 31;
    case 2: case 3:
        if (info == 31) return skip_indefinite(r, major, depth);
        if (!r.need(arg)) return false;
        r.pos += (size_t)arg; return true;
    case 4:
        if (info == 31) return skip_indefinite(r, 4, depth);
        for (uint64_t i = 0; i < arg; i++) if (!skip_item(r, depth + 1)) return false;
        return true;
    case 5:
        if (info == 31) return skip_indefinite(r, 5, depth);
        for (uint64_t i = 0; i < arg; i++) { if (!skip_item(r, depth + 1) || !skip_item(r, depth + 1)) return false; }
        return true;
    case 6: return info != 31 && skip_item(r, depth + 1);
    default:  // 7: simple / float / break (a stray break is malformed)
        return info != 31;
    }
}
// Value::Bytes of length 32 at the cursor -> out; false = "expected 32-byte array" (structure error).  The item is
// consumed either way (it was validated by the well-formedness pass).
bool take_bytes32(Reader& r, uint8_t out[32]) {
    Reader save = r;
    int major, info; uint64_t arg;
    read_head(r, major, info, arg);
    if (major != 2) { r = save; skip_item(r, 0); return false; }
    if (info != 31) {
        bool ok = (arg == 32);
        if (ok) memcpy(out, r.p + r.pos, 32);
        r.pos += (size_t)arg;
        return ok;
    }
    uint8_t buf[32]; size_t got = 0; bool ok = true;
    while (r.p[r.pos] != 0xff) {
        int m, i2; uint64_t a2;
        read_head(r, m, i2, a2);
        for (uint64_t k = 0; k < a2; k++) { if (got < 32) buf[got] = r.p[r.pos + k]; got++; if (got > 32) ok = false; }
        r.pos += (size_t)a2;
    }
    r.pos++;
    if (ok && got == 32) { memcpy(out, buf, 32); return true; }
    return false;
}
// iterates the entries of a map/array whose head has been read; definite count or indefinite
struct Iter {
    bool indefinite; uint64_t left;
    bool next(Reader& r) {
        if (indefinite) { if (r.p[r.pos] == 0xff) { r.pos++; return false; } return true; }
        if (left == 0) return false;
        left--; return true;
    }
};
// integer key value, or -1 when the key is not a (small) unsigned integer
int64_t take_key(Reader& r) {
    Reader save = r;
    int major, info; uint64_t arg;
    read_head(r, major, info, arg);
    if (major == 0 && arg < 1000) return (int64_t)arg;
    r = save; skip_item(r, 0);
    return -1;
}

// generic decoder for the flat maps {1..nf: bstr32}
int unpack_flat(const uint8_t* item, size_t len, int nf, uint8_t* rec) {
    Reader r{item, len, 0};
    { Reader chk = r; if (!skip_item(chk, 0)) return ST_PARSE; }
    int major, info; uint64_t arg;
    read_head(r, major, info, arg);
    if (major != 5) return ST_STRUCT;
    Iter it{info == 31, arg};
    uint32_t have = 0;
    while (it.next(r)) {
        int64_t k = take_key(r);
        if (k >= 1 && k <= nf) {
            if (!take_bytes32(r, rec + 32 * (k - 1))) return ST_STRUCT;
            have |= 1u << k;
        } else {
            skip_item(r, 0);
        }
    }
    for (int k = 1; k <= nf; k++) if (!(have & (1u << k))) return ST_STRUCT;
    return ST_OK;
}

const int L = 128;
// record offsets (32-byte units) of the single-value keys of SpendProof, index = key
const int kProofSlot[18] = {-1, 0, 1, 2, 3, -1, 132, 133, 134, 135, 136, 137, 138, 139, -1, -1, 524, 525};

int unpack_proof(const uint8_t* item, size_t len, uint8_t* rec) {
    Reader r{item, len, 0};
    { Reader chk = r; if (!skip_item(chk, 0)) return ST_PARSE; }
    int major, info; uint64_t arg;
    read_head(r, major, info, arg);
    if (major != 5) return ST_STRUCT;
    Iter it{info == 31, arg};
    uint32_t have = 0;
    while (it.next(r)) {
        int64_t k = take_key(r);
        if (k < 1 || k > 17) { skip_item(r, 0); continue; }
        if (k == 5 || k == 14 || k == 15) {
            Reader save = r;
            int m2, i2; uint64_t a2;
            read_head(r, m2, i2, a2);
            if (m2 != 4) { r = save; skip_item(r, 0); continue; }  // non-array: silently skipped (src/cbor.rs:306,331,347)
            Iter ait{i2 == 31, a2};
            size_t count = 0;
            size_t base = (k == 5) ? 4 : (k == 14 ? 140 : 268);
            uint8_t tmp[64];
            while (ait.next(r)) {
                bool in_range = count < (size_t)L;
                if (k == 15) {
                    Reader s2 = r;
                    int m3, i3; uint64_t a3;
                    read_head(r, m3, i3, a3);
                    if (m3 != 4) return ST_STRUCT;  // "expected array for z pair"
                    Iter pit{i3 == 31, a3};
                    size_t pc = 0; bool bad = false;
                    // the reference checks pair.len() == 2 first, then decodes both entries
                    Reader cnt = r; Iter cit = pit; size_t total = 0;
                    while (cit.next(cnt)) { skip_item(cnt, 0); total++; }
                    if (total != 2) return ST_STRUCT;  // "z pair wrong size"
                    while (pit.next(r)) { if (!take_bytes32(r, tmp + 32 * pc)) bad = true; pc++; }
                    if (bad) return ST_STRUCT;
                    if (in_range) memcpy(rec + 32 * (base + 2 * count), tmp, 64);
                    (void)s2;
                } else {
                    if (!take_bytes32(r, tmp)) return ST_STRUCT;
                    if (in_range) memcpy(rec + 32 * (base + count), tmp, 32);
                }
                count++;
            }
            if (count != (size_t)L) return ST_STRUCT;  // "... array wrong size"
            have |= 1u << k;
        } else {
            if (!take_bytes32(r, rec + 32 * kProofSlot[k])) return ST_STRUCT;
            have |= 1u << k;
        }
    }
    for (int k = 1; k <= 17; k++) if (!(have & (1u << k))) return ST_STRUCT;
    return ST_OK;
}

template <typename F>
int pack_all(size_t n, const uint8_t* const* items, const size_t* lens, uint8_t* rec, size_t rec_bytes, uint8_t* status, F f) {
    if (n && (!items || !lens || !rec || !status)) return -1;
    for (size_t i = 0; i < n; i++) {
        uint8_t* out = rec + rec_bytes * i;
        memset(out, 0, rec_bytes);
        int st = items[i] ? f(items[i], lens[i], out) : ST_PARSE;
        if (st != ST_OK) memset(out, 0, rec_bytes);
        status[i] = (uint8_t)st;
    }
    return 0;
}

uint8_t* put_bstr32(uint8_t* o, const uint8_t* v) { *o++ = 0x58; *o++ = 0x20; memcpy(o, v, 32); return o + 32; }
size_t encode_flat(const uint8_t* rec, int nf, uint8_t* out) {
    uint8_t* o = out;
    *o++ = (uint8_t)(0xA0 | nf);
    for (int k = 1; k <= nf; k++) { *o++ = (uint8_t)k; o = put_bstr32(o, rec + 32 * (k - 1)); }
    return (size_t)(o - out);
}

}  // namespace

extern "C" int act_pack_issuance_requests_cbor(size_t n, const uint8_t* const* items, const size_t* lens, uint8_t* req, uint8_t* status) {
    return pack_all(n, items, lens, req, 128, status, [](const uint8_t* p, size_t l, uint8_t* o) { return unpack_flat(p, l, 4, o); });
}
extern "C" int act_pack_issuance_responses_cbor(size_t n, const uint8_t* const* items, const size_t* lens, uint8_t* resp, uint8_t* status) {
    return pack_all(n, items, lens, resp, 160, status, [](const uint8_t* p, size_t l, uint8_t* o) { return unpack_flat(p, l, 5, o); });
}
extern "C" int act_pack_refunds_cbor(size_t n, const uint8_t* const* items, const size_t* lens, uint8_t* refunds, uint8_t* status) {
    return pack_all(n, items, lens, refunds, 128, status, [](const uint8_t* p, size_t l, uint8_t* o) { return unpack_flat(p, l, 4, o); });
}
extern "C" int act_pack_spend_proofs_cbor(size_t n, const uint8_t* const* items, const size_t* lens, uint8_t* proofs, uint8_t* status) {
    return pack_all(n, items, lens, proofs, ACT_PROOF_BYTES, status, [](const uint8_t* p, size_t l, uint8_t* o) { return unpack_proof(p, l, o); });
}
extern "C" size_t act_encode_issuance_request_cbor(const uint8_t req[128], uint8_t out[141]) { return encode_flat(req, 4, out); }
extern "C" size_t act_encode_issuance_response_cbor(const uint8_t resp[160], uint8_t out[176]) { return encode_flat(resp, 5, out); }
extern "C" size_t act_encode_refund_cbor(const uint8_t refund[128], uint8_t out[141]) { return encode_flat(refund, 4, out); }
extern "C" size_t act_encode_spend_proof_cbor(const uint8_t* pf, uint8_t* out) {
    // src/cbor.rs:236-273: map(17), keys in order, arrays of 128
    uint8_t* o = out;
    *o++ = 0xB1;
    for (int k = 1; k <= 17; k++) {
        *o++ = (uint8_t)k;
        if (k == 5 || k == 14) {
            size_t base = (k == 5) ? 4 : 140;
            *o++ = 0x98; *o++ = 0x80;
            for (int j = 0; j < L; j++) o = put_bstr32(o, pf + 32 * (base + j));
        } else if (k == 15) {
            *o++ = 0x98; *o++ = 0x80;
            for (int j = 0; j < L; j++) { *o++ = 0x82; o = put_bstr32(o, pf + 32 * (268 + 2 * j)); o = put_bstr32(o, pf + 32 * (269 + 2 * j)); }
        } else {
            o = put_bstr32(o, pf + 32 * kProofSlot[k]);
        }
    }
    return (size_t)(o - out);
}
