"""Independent restatement of the reference protocol for cross-checking the C oracle.

TEST INFRASTRUCTURE ONLY.  Stack: python `blake3` (hash), libsodium ristretto255 through ctypes
(group), python big ints (scalars mod l).  It shares no code with oracle/act_oracle.c or with the
CUDA engine, so agreement between the three is the strongest parity evidence available in an image
with no Rust toolchain.

Follows /root/reference: src/lib.rs:291-354 (Params::new), :463-487 (request), :528-562,
:621-663 (issue), :781-869 (refund), :972-1152 (prove_spend), :1217-1253; src/transcript.rs:54-155.
"""
import ctypes as C
import glob
import os

import blake3 as _b3

ELL = 2**252 + 27742317777372353535851937790883648493
L_BITS = 128
PROTOCOL_VERSION = b"curve25519-ristretto anonymous-credits v1.0"
IDENT = bytes(32)


def _find_sodium():
    import sys
    cands = []
    for sp in sys.path:
        cands += glob.glob(os.path.join(sp, "pyzmq.libs", "libsodium*.so*"))
    cands += glob.glob("/usr/lib/x86_64-linux-gnu/libsodium.so*")
    for c in cands:
        try:
            lib = C.CDLL(c)
            lib.crypto_core_ristretto255_from_hash  # noqa: B018
            return lib
        except (OSError, AttributeError):
            continue
    return None


_sodium = _find_sodium()


def available():
    return _sodium is not None


def _out():
    return C.create_string_buffer(32)


def sc_bytes(x):
    return (x % ELL).to_bytes(32, "little")


def sc_int(b):
    return int.from_bytes(b, "little") % ELL


def sc_wide(b64):
    return int.from_bytes(b64, "little") % ELL


def is_valid(p):
    return _sodium.crypto_core_ristretto255_is_valid_point(p) == 1


def mul(p, s):
    s %= ELL
    if s == 0 or p == IDENT:
        return IDENT
    q = _out()
    if _sodium.crypto_scalarmult_ristretto255(q, sc_bytes(s), p) != 0:
        return IDENT
    return q.raw


def mul_base(s):
    s %= ELL
    if s == 0:
        return IDENT
    q = _out()
    if _sodium.crypto_scalarmult_ristretto255_base(q, sc_bytes(s)) != 0:
        return IDENT
    return q.raw


def add(p, q):
    r = _out()
    assert _sodium.crypto_core_ristretto255_add(r, p, q) == 0
    return r.raw


def sub(p, q):
    r = _out()
    assert _sodium.crypto_core_ristretto255_sub(r, p, q) == 0
    return r.raw


def from_hash(b64):
    r = _out()
    _sodium.crypto_core_ristretto255_from_hash(r, b64)
    return r.raw


G = None


def gen():
    global G
    if G is None:
        G = mul_base(1)
    return G


def params_new(org, svc, dep, ver):
    dom = f"ACT-v1:{org}:{svc}:{dep}:{ver}".encode()
    seed = _b3.blake3(len(dom).to_bytes(8, "big") + dom).digest()
    hs = []
    for ctr in range(3):
        h = _b3.blake3()
        h.update(len(dom).to_bytes(8, "big") + dom)
        h.update((32).to_bytes(8, "big") + seed)
        h.update((4).to_bytes(8, "big") + ctr.to_bytes(4, "little"))
        hs.append(from_hash(h.digest(length=64)))
    return hs


class Transcript:
    def __init__(self, H, label):
        self.h = _b3.blake3()
        self.h.update(len(PROTOCOL_VERSION).to_bytes(8, "big") + PROTOCOL_VERSION)
        for p in H:
            self.add(p)
        self.h.update(len(label).to_bytes(8, "big") + label)

    def add(self, b):
        self.h.update(len(b).to_bytes(8, "big") + b)

    def add_scalar(self, s):
        self.add(sc_bytes(s))

    def challenge(self):
        return sc_wide(self.h.digest(length=64))


class Rng:
    """Byte stream; Scalar::random = next 64 bytes, wide-reduced."""

    def __init__(self, data):
        self.d, self.o = data, 0

    def scalar(self):
        s = sc_wide(self.d[self.o:self.o + 64])
        self.o += 64
        return s


def keygen(rng):
    x = rng.scalar()
    return x, mul_base(x)


def request(H, r, k, rng):
    K = add(mul(H[1], k), mul(H[2], r))
    kp, rp = rng.scalar(), rng.scalar()
    K1 = add(mul(H[1], kp), mul(H[2], rp))
    t = Transcript(H, b"request"); t.add(K); t.add(K1)
    g = t.challenge()
    return dict(K=K, gamma=g, k_bar=(kp + k * g) % ELL, r_bar=(rp + r * g) % ELL)


def issue(H, x, W, rq, c, rng):
    K1 = sub(add(mul(H[1], rq["k_bar"]), mul(H[2], rq["r_bar"])), mul(rq["K"], rq["gamma"]))
    t = Transcript(H, b"request"); t.add(rq["K"]); t.add(K1)
    if t.challenge() != rq["gamma"]:
        return None
    e = rng.scalar()
    XA = add(add(gen(), mul(H[0], c)), rq["K"])
    A = mul(XA, pow((e + x) % ELL, -1, ELL))
    XG = add(mul_base(e), W)
    alpha = rng.scalar()
    YA, YG = mul(A, alpha), mul_base(alpha)
    t = Transcript(H, b"respond"); t.add_scalar(c); t.add_scalar(e)
    for p in (A, XA, XG, YA, YG):
        t.add(p)
    g = t.challenge()
    return dict(A=A, e=e, gamma=g, z=(g * (x + e) + alpha) % ELL, c=c % ELL)


def issuance_check(H, W, K, rs):
    XA = add(add(gen(), mul(H[0], rs["c"])), K)
    XG = add(mul_base(rs["e"]), W)
    YA = add(mul(rs["A"], rs["z"]), mul(XA, -rs["gamma"]))
    YG = add(mul_base(rs["z"]), mul(XG, -rs["gamma"]))
    t = Transcript(H, b"respond"); t.add_scalar(rs["c"]); t.add_scalar(rs["e"])
    for p in (rs["A"], XA, XG, YA, YG):
        t.add(p)
    return t.challenge() == rs["gamma"]


def prove_spend(H, tok, s, rng):
    A, e, k, r, c = tok["A"], tok["e"], tok["k"], tok["r"], tok["c"]
    r1, r2, cp, rp, ep, r2p, r3p = (rng.scalar() for _ in range(7))
    B = add(add(add(gen(), mul(H[0], c)), mul(H[1], k)), mul(H[2], r))
    Ap = mul(A, r1 * r2)
    Bb = mul(B, r1)
    r3 = pow(r1, -1, ELL)
    A1 = add(mul(Ap, ep), mul(Bb, r2p))
    A2 = add(add(mul(Bb, r3p), mul(H[0], cp)), mul(H[2], rp))
    m = (c - s) % ELL
    mb = sc_bytes(m)
    bits = [(mb[i // 8] >> (i % 8)) & 1 for i in range(L_BITS)]
    kstar = rng.scalar()
    si = [rng.scalar() for _ in range(L_BITS)]
    com = []
    for j in range(L_BITS):
        p = add(mul(H[0], bits[j]), mul(H[2], si[j]))
        if j == 0:
            p = add(p, mul(H[1], kstar))
        com.append(p)
    k0p = rng.scalar()
    sip = [rng.scalar() for _ in range(L_BITS)]
    gi = [rng.scalar() for _ in range(L_BITS)]
    w0 = rng.scalar()
    z = [rng.scalar() for _ in range(L_BITS)]
    Cp = []
    for j in range(L_BITS):
        C0, C1 = com[j], sub(com[j], H[0])
        if j == 0:
            base = add(mul(H[1], w0), mul(H[2], z[0]))
            real = add(mul(H[1], k0p), mul(H[2], sip[0]))
        else:
            base = mul(H[2], z[j])
            real = mul(H[2], sip[j])
        if bits[j] == 0:
            Cp.append((real, sub(base, mul(C1, gi[j]))))
        else:
            Cp.append((sub(base, mul(C0, gi[j])), real))
    rstar = sum(si[i] << i for i in range(L_BITS)) % ELL
    kp, sp = rng.scalar(), rng.scalar()
    C_ = add(add(mul(H[0], -cp), mul(H[1], kp)), mul(H[2], sp))
    t = Transcript(H, b"spend"); t.add_scalar(k)
    for p in (Ap, Bb, A1, A2):
        t.add(p)
    for p in com:
        t.add(p)
    for a, b in Cp:
        t.add(a); t.add(b)
    t.add(C_)
    g = t.challenge()
    g00 = [(g - gi[j]) % ELL if bits[j] == 0 else gi[j] for j in range(L_BITS)]
    if bits[0] == 0:
        w00, w01 = (g00[0] * kstar + k0p) % ELL, w0
    else:
        w00, w01 = w0, ((g - g00[0]) * kstar + k0p) % ELL
    zz = []
    for j in range(L_BITS):
        if bits[j] == 0:
            zz.append(((g00[j] * si[j] + sip[j]) % ELL, z[j]))
        else:
            zz.append((z[j], ((g - g00[j]) * si[j] + sip[j]) % ELL))
    proof = dict(k=k, s=s % ELL, Ap=Ap, Bb=Bb, com=com, gamma=g, e_bar=(-g * e + ep) % ELL,
                 r2_bar=(g * r2 + r2p) % ELL, r3_bar=(g * r3 + r3p) % ELL, c_bar=(-g * c + cp) % ELL,
                 r_bar=(-g * r + rp) % ELL, w00=w00, w01=w01, gamma0=g00, z=zz,
                 k_bar=(g * kstar + kp) % ELL, s_bar=(g * rstar + sp) % ELL)
    return proof, dict(k=kstar, r=rstar, m=m)


def refund(H, x, W, pf, rng):
    if pf["Ap"] == IDENT:
        return "identity"
    g = pf["gamma"]
    Abar = mul(pf["Ap"], x)
    H1p = add(gen(), mul(H[1], pf["k"]))
    A1 = add(add(mul(pf["Ap"], pf["e_bar"]), mul(pf["Bb"], pf["r2_bar"])), mul(Abar, -g))
    A2 = add(add(add(mul(pf["Bb"], pf["r3_bar"]), mul(H[0], pf["c_bar"])), mul(H[2], pf["r_bar"])), mul(H1p, -g))
    Cp = []
    for j in range(L_BITS):
        g0 = pf["gamma0"][j]; g1 = (g - g0) % ELL
        C0, C1 = pf["com"][j], sub(pf["com"][j], H[0])
        a = mul(H[2], pf["z"][j][0]); b = mul(H[2], pf["z"][j][1])
        if j == 0:
            a = add(mul(H[1], pf["w00"]), a); b = add(mul(H[1], pf["w01"]), b)
        Cp.append((sub(a, mul(C0, g0)), sub(b, mul(C1, g1))))
    Kp = IDENT
    for i in range(L_BITS):
        Kp = add(Kp, mul(pf["com"][i], 1 << i))
    com_ = add(mul(H[0], pf["s"]), Kp)
    Cc = sub(add(add(mul(H[0], -pf["c_bar"]), mul(H[1], pf["k_bar"])), mul(H[2], pf["s_bar"])), mul(com_, g))
    t = Transcript(H, b"spend"); t.add_scalar(pf["k"])
    for p in (pf["Ap"], pf["Bb"], A1, A2):
        t.add(p)
    for p in pf["com"]:
        t.add(p)
    for a, b in Cp:
        t.add(a); t.add(b)
    t.add(Cc)
    if t.challenge() != g:
        return "invalid"
    e = rng.scalar()
    XA = add(gen(), Kp)
    A = mul(XA, pow((e + x) % ELL, -1, ELL))
    XG = add(mul_base(e), W)
    alpha = rng.scalar()
    YA, YG = mul(A, alpha), mul_base(alpha)
    t = Transcript(H, b"refund"); t.add_scalar(e)
    for p in (A, XA, XG, YA, YG):
        t.add(p)
    gr = t.challenge()
    return dict(A=A, e=e, gamma=gr, z=(gr * (x + e) + alpha) % ELL)


def refund_check(H, W, com, rf):
    Kp = IDENT
    for i in range(L_BITS):
        Kp = add(Kp, mul(com[i], 1 << i))
    XA = add(gen(), Kp)
    XG = add(mul_base(rf["e"]), W)
    YA = add(mul(rf["A"], rf["z"]), mul(XA, -rf["gamma"]))
    YG = add(mul_base(rf["z"]), mul(XG, -rf["gamma"]))
    t = Transcript(H, b"refund"); t.add_scalar(rf["e"])
    for p in (rf["A"], XA, XG, YA, YG):
        t.add(p)
    return t.challenge() == rf["gamma"]


# ---- packing to the wire-record layouts of include/act_engine.h ----
def pack_request(rq):
    return rq["K"] + sc_bytes(rq["gamma"]) + sc_bytes(rq["k_bar"]) + sc_bytes(rq["r_bar"])


def pack_response(rs):
    return rs["A"] + sc_bytes(rs["e"]) + sc_bytes(rs["gamma"]) + sc_bytes(rs["z"]) + sc_bytes(rs["c"])


def pack_proof(pf):
    out = [sc_bytes(pf["k"]), sc_bytes(pf["s"]), pf["Ap"], pf["Bb"]] + list(pf["com"])
    out += [sc_bytes(pf[n]) for n in ("gamma", "e_bar", "r2_bar", "r3_bar", "c_bar", "r_bar", "w00", "w01")]
    out += [sc_bytes(v) for v in pf["gamma0"]]
    for a, b in pf["z"]:
        out += [sc_bytes(a), sc_bytes(b)]
    out += [sc_bytes(pf["k_bar"]), sc_bytes(pf["s_bar"])]
    b = b"".join(out)
    assert len(b) == 526 * 32
    return b


def pack_refund(rf):
    return rf["A"] + sc_bytes(rf["e"]) + sc_bytes(rf["gamma"]) + sc_bytes(rf["z"])


# ---- CBOR (reference src/cbor.rs), via cbor2 ----
def cbor_request(rq):
    import cbor2
    return cbor2.dumps({1: rq["K"], 2: sc_bytes(rq["gamma"]), 3: sc_bytes(rq["k_bar"]), 4: sc_bytes(rq["r_bar"])})


def cbor_response(rs):
    import cbor2
    return cbor2.dumps({1: rs["A"], 2: sc_bytes(rs["e"]), 3: sc_bytes(rs["gamma"]), 4: sc_bytes(rs["z"]), 5: sc_bytes(rs["c"])})


def cbor_proof(pf):
    import cbor2
    return cbor2.dumps({
        1: sc_bytes(pf["k"]), 2: sc_bytes(pf["s"]), 3: pf["Ap"], 4: pf["Bb"], 5: list(pf["com"]),
        6: sc_bytes(pf["gamma"]), 7: sc_bytes(pf["e_bar"]), 8: sc_bytes(pf["r2_bar"]), 9: sc_bytes(pf["r3_bar"]),
        10: sc_bytes(pf["c_bar"]), 11: sc_bytes(pf["r_bar"]), 12: sc_bytes(pf["w00"]), 13: sc_bytes(pf["w01"]),
        14: [sc_bytes(v) for v in pf["gamma0"]], 15: [[sc_bytes(a), sc_bytes(b)] for a, b in pf["z"]],
        16: sc_bytes(pf["k_bar"]), 17: sc_bytes(pf["s_bar"])})


def cbor_refund(rf):
    import cbor2
    return cbor2.dumps({1: rf["A"], 2: sc_bytes(rf["e"]), 3: sc_bytes(rf["gamma"]), 4: sc_bytes(rf["z"])})


# ---- record-level front ends: wire bytes in, status + wire bytes out (include/act_engine.h conventions) ----------------
# Decode semantics of src/cbor.rs:62-91: a point must be a valid ristretto255 encoding (else CborError::InvalidValue ->
# 0x81, reported before anything else because decoding precedes the protocol call); scalars are reduced mod l, never rejected.
ST_OK, ST_BAD_REQUEST, ST_BAD_RESPONSE, ST_BAD_REFUND, ST_IDENTITY, ST_BAD_SPEND, ST_DECODE = 0, 1, 2, 4, 6, 7, 0x81


def unpack_proof(b):
    """SpendProof record (16 832 B, include/act_engine.h) -> dict of reduced scalars and point encodings; None if any of the
    130 points is not a valid encoding."""
    b = bytes(b)
    assert len(b) == 526 * 32
    f = lambda i: b[32 * i:32 * i + 32]
    pts = [f(2), f(3)] + [f(4 + j) for j in range(L_BITS)]
    if not all(is_valid(p) for p in pts):
        return None
    return dict(k=sc_int(f(0)), s=sc_int(f(1)), Ap=f(2), Bb=f(3), com=[f(4 + j) for j in range(L_BITS)], gamma=sc_int(f(132)),
                e_bar=sc_int(f(133)), r2_bar=sc_int(f(134)), r3_bar=sc_int(f(135)), c_bar=sc_int(f(136)), r_bar=sc_int(f(137)),
                w00=sc_int(f(138)), w01=sc_int(f(139)), gamma0=[sc_int(f(140 + j)) for j in range(L_BITS)],
                z=[(sc_int(f(268 + 2 * j)), sc_int(f(269 + 2 * j))) for j in range(L_BITS)], k_bar=sc_int(f(524)), s_bar=sc_int(f(525)))


def refund_record(H, x, W, proof, rnd128):
    """-> (status, refund 128 B, nullifier 32 B); zero-filled outputs on reject."""
    pf = unpack_proof(proof)
    if pf is None:
        return ST_DECODE, bytes(128), bytes(32)
    r = refund(H, x, W, pf, Rng(bytes(rnd128)))
    if r == "identity":
        return ST_IDENTITY, bytes(128), bytes(32)
    if r == "invalid":
        return ST_BAD_SPEND, bytes(128), bytes(32)
    return ST_OK, pack_refund(r), sc_bytes(pf["k"])


def issue_record(H, x, W, req, c32, rnd128):
    """-> (status, response 160 B)."""
    req = bytes(req)
    if not is_valid(req[:32]):
        return ST_DECODE, bytes(160)
    rq = dict(K=req[:32], gamma=sc_int(req[32:64]), k_bar=sc_int(req[64:96]), r_bar=sc_int(req[96:128]))
    r = issue(H, x, W, rq, sc_int(bytes(c32)), Rng(bytes(rnd128)))
    if r is None:
        return ST_BAD_REQUEST, bytes(160)
    return ST_OK, pack_response(r)


def issuance_check_record(H, W, K, resp):
    K, resp = bytes(K), bytes(resp)
    if not (is_valid(K) and is_valid(resp[:32])):
        return ST_DECODE
    rs = dict(A=resp[:32], e=sc_int(resp[32:64]), gamma=sc_int(resp[64:96]), z=sc_int(resp[96:128]), c=sc_int(resp[128:160]))
    return ST_OK if issuance_check(H, W, K, rs) else ST_BAD_RESPONSE


def refund_check_record(H, W, com4096, refund128):
    com4096, rf = bytes(com4096), bytes(refund128)
    com = [com4096[32 * j:32 * j + 32] for j in range(L_BITS)]
    if not (is_valid(rf[:32]) and all(is_valid(p) for p in com)):
        return ST_DECODE
    d = dict(A=rf[:32], e=sc_int(rf[32:64]), gamma=sc_int(rf[64:96]), z=sc_int(rf[96:128]))
    return ST_OK if refund_check(H, W, com, d) else ST_BAD_REFUND
