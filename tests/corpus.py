"""Deterministic synthetic corpora for parity tests and the bench: valid requests/proofs produced by the
oracle's fixture generators (request, prove_spend) plus one generator per mutation class of the
reference's own tests (/root/reference src/tests.rs).

TEST INFRASTRUCTURE ONLY (uses oracle/).
"""
import os

import numpy as np

import oracle_lib as O

ELL = 2**252 + 27742317777372353535851937790883648493
P25519 = 2**255 - 19
PROOF_BYTES = O.PROOF_BYTES
BENCH_PARAMS = ("bench-org", "bench-service", "bench-env", "2024-01-01")   # benches/benchmark.rs:10-15
TEST_PARAMS = ("test-org", "test-service", "test-env", "2024-01-01")       # src/tests.rs:59


def xof(seed: bytes, n: int) -> bytes:
    return O.blake3(seed, n)


def sc_bytes(x):
    return (x % ELL).to_bytes(32, "little")


def sc_int(b):
    return int.from_bytes(bytes(b), "little")


def make_ctx(params=TEST_PARAMS, key_seed=b"act-b200-key-0"):
    h = O.params_derive(*params)
    x, w = O.keygen(xof(key_seed, 64))
    return O.Ctx(h, x, w)


def gen_valid(ctx, n, seed=b"corpus-0", credits=(20, 1000), threads=None, want_proofs=True):
    """n independent issue->token->prove_spend trips.  credits uniform in [lo,hi), charge uniform in [1,c-1]
    (benches/benchmark.rs:60,178,194-201)."""
    threads = threads or min(os.cpu_count() or 1, 32)
    rs = np.random.RandomState(int.from_bytes(xof(seed, 4), "little"))
    seeds = np.frombuffer(xof(seed + b"/seeds", 64 * n), dtype=np.uint8).copy()
    cr = rs.randint(credits[0], credits[1], size=n).astype(np.uint64)
    ch = np.array([rs.randint(1, max(2, int(c))) for c in cr], dtype=np.uint64)
    out = ctx.generate(seeds, cr, ch, threads=threads, want_proofs=want_proofs)
    out["rnd"] = np.frombuffer(xof(seed + b"/rnd", 128 * n), dtype=np.uint8).copy()
    out["credits"], out["charges"] = cr, ch
    return out


# ---- invalid point encodings (RFC 9496 A.3 classes; src/cbor.rs:62-77) ----
def bad_point_encodings():
    out = []
    out.append((P25519).to_bytes(32, "little"))            # non-canonical field element (s = p)
    out.append((P25519 + 2).to_bytes(32, "little"))        # non-canonical
    out.append((2**255 + 4).to_bytes(32, "little"))        # top bit set
    out.append(b"\xff" * 32)
    out.append((1).to_bytes(32, "little"))                 # negative (odd) s
    # non-square / negative-t / y=0 cases from RFC 9496 A.3
    for hx in ("26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371",
               "4eac077a713c57b4f4397629a4145982c661f48044dd3f96427d40b147d9742f",
               "de6a7b00deadbeefde6a7b00deadbeefde6a7b00deadbeefde6a7b00deadbe6f",
               "bcab477be20861e01e4a0e295284146a510150d9817763caf1a6f4b422d67042",
               "2a292df7e32cababbd9de088d1d1abec9fc0440f637ed2fba145094dc14bea08",
               "f4a9e534fc0d216c44b218fa0c42d99635a0127ee2e53c712f70609649fdff22",
               "8268436f8c4126196cf64b3c7ddbda90746a378625f9813dd9b8457077256731",
               "2810e5cbc2cc4d4eece54f61c6f69758e289aa7ab440b3cbeaa21995c2f4232b",
               "3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e",
               "a45fdc55c76448c049a1ab33f17023edfb2be3581e9c7aade8a6125215e04220",
               "d483fe813c6ba647ebbfd3ec41adca1c6130c2beeee9d9bf065c8d151c5f396e",
               "8a2e1d30050198c65a54483123960ccc38aef6848e1ec8f5f780e8523769ba32",
               "32888462f8b486c68ad7dd9610be5192bbeaf3b443951ac1a8118419d9fa097b",
               "227142501b9d4355ccba290404bde41575b037693cef1f438c47f8fbf35d1165",
               "5c37cc491da847cfeb9281d407efc41e15144c876e0170b499a96a22ed31e01e",
               "445425117cb8c90edcbc7c1cc0e74f747f2c1efa5630a967c64f287792a48a4b"):
        out.append(bytes.fromhex(hx))
    return out


def mutate_requests(ctx, base, seed=b"mut-req"):
    """Returns (req, cs, rnd, expected_status, label) arrays built from valid requests `base` (dict of gen_valid)."""
    req = base["req"].reshape(-1, 128).copy(); cs = base["cs"].reshape(-1, 32).copy(); rnd = base["rnd"].reshape(-1, 128).copy()
    n = len(req)
    expect = np.zeros(n, np.uint8); labels = ["valid"] * n
    bad = bad_point_encodings()
    rs = np.random.RandomState(7)
    for i in range(n):
        kind = i % 8
        if kind == 1:      # k_bar += 1  (src/tests.rs:583-593)
            req[i, 64:96] = np.frombuffer(sc_bytes(sc_int(req[i, 64:96]) + 1), np.uint8); expect[i] = 1; labels[i] = "k_bar+1"
        elif kind == 2:    # random K and gamma (src/tests.rs:1945-1953)
            req[i, 0:32] = np.frombuffer(O.scalarmult_base(xof(seed + bytes([i & 255, 1]), 32)), np.uint8)
            req[i, 32:64] = np.frombuffer(O.sc_reduce32(xof(seed + bytes([i & 255, 2]), 32)), np.uint8); expect[i] = 1; labels[i] = "random K,gamma"
        elif kind == 3:    # malformed point
            req[i, 0:32] = np.frombuffer(bad[(i // 8) % len(bad)], np.uint8); expect[i] = 0x81; labels[i] = "bad K"
        elif kind == 4:    # non-canonical scalar encodings still accept (src/cbor.rs:80-91)
            for off in (32, 64, 96):
                v = sc_int(req[i, off:off + 32]) + ELL
                if v < 2**256:
                    req[i, off:off + 32] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
            v = sc_int(cs[i]) + ELL
            cs[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8); labels[i] = "non-canonical scalars"
        elif kind == 5:    # single random bit flip somewhere in the scalars
            b = 32 * 8 + rs.randint(0, 96 * 8)
            req[i, b // 8] ^= 1 << (b % 8); expect[i] = 255; labels[i] = "bit flip"   # 255 = whatever the oracle says
        elif kind == 6:    # gamma += t
            req[i, 32:64] = np.frombuffer(sc_bytes(sc_int(req[i, 32:64]) + 1 + rs.randint(0, 1 << 30)), np.uint8); expect[i] = 1; labels[i] = "gamma+t"
    return req.reshape(-1), cs.reshape(-1), rnd.reshape(-1), expect, labels


def _set(pf, idx, val32):
    pf[32 * idx:32 * idx + 32] = np.frombuffer(bytes(val32), np.uint8)


def _get(pf, idx):
    return bytes(pf[32 * idx:32 * idx + 32])


def mutate_proofs(ctx, base, seed=b"mut-pf"):
    """Adversarial spend corpus: (proofs, rnd, expected_status, labels).  expected 255 = defer to the oracle."""
    proofs = base["proofs"].reshape(-1, PROOF_BYTES).copy(); rnd = base["rnd"].reshape(-1, 128).copy()
    n = len(proofs)
    expect = np.zeros(n, np.uint8); labels = ["valid"] * n
    bad = bad_point_encodings()
    rs = np.random.RandomState(11)
    for i in range(n):
        pf = proofs[i]
        kind = i % 12
        if kind == 1:      # s replaced (src/tests.rs:631-638)
            _set(pf, 1, sc_bytes(sc_int(_get(pf, 1)) + 1)); expect[i] = 7; labels[i] = "s changed"
        elif kind == 2:    # gamma += t (src/tests.rs:1701-1708)
            _set(pf, 132, sc_bytes(sc_int(_get(pf, 132)) + 1 + rs.randint(0, 1 << 30))); expect[i] = 7; labels[i] = "gamma+t"
        elif kind == 3:    # a_prime = identity (src/tests.rs:868-872)
            _set(pf, 2, bytes(32)); expect[i] = 6; labels[i] = "A' identity"
        elif kind == 4:    # malformed com[j]
            _set(pf, 4 + rs.randint(0, 128), bad[(i // 12) % len(bad)]); expect[i] = 0x81; labels[i] = "bad com"
        elif kind == 5:    # malformed A' or B_bar
            _set(pf, 2 + (i // 12) % 2, bad[(i // 12 + 3) % len(bad)]); expect[i] = 0x81; labels[i] = "bad A'/B"
        elif kind == 6:    # non-canonical scalars accept; nullifier is the reduced value
            for idx in (0, 133, 140 + rs.randint(0, 128), 268 + rs.randint(0, 256), 525):
                v = sc_int(_get(pf, idx)) + ELL
                if v < 2**256:
                    _set(pf, idx, v.to_bytes(32, "little"))
            labels[i] = "non-canonical scalars"
        elif kind == 7:    # exact replay of the previous valid proof (caller's job to catch; refund still Ok)
            if i >= 7:
                proofs[i] = proofs[i - 7]; labels[i] = "replay"
        elif kind == 8:    # random bit flip in a scalar field of the range proof
            idx = 140 + rs.randint(0, 384); b = rs.randint(0, 252)
            pf[32 * idx + b // 8] ^= 1 << (b % 8); expect[i] = 7; labels[i] = "range scalar flip"
        elif kind == 9:    # valid point swapped in for a commitment (com[j] := com[j'] )
            j = rs.randint(0, 127); _set(pf, 4 + j, _get(pf, 5 + j)); expect[i] = 255; labels[i] = "com swapped"
        elif kind == 10:   # identity encoding inside com[] is a VALID point: proof fails, not a decode error
            _set(pf, 4 + rs.randint(0, 128), bytes(32)); expect[i] = 7; labels[i] = "com identity"
        elif kind == 11:   # k changed (nullifier forgery)
            _set(pf, 0, sc_bytes(sc_int(_get(pf, 0)) + 1)); expect[i] = 7; labels[i] = "k changed"
    return proofs.reshape(-1), rnd.reshape(-1), expect, labels


def edge_issue_inputs(base):
    """Valid requests with extreme credit amounts and degenerate signer randomness (src/tests.rs:642-689 large amounts,
    :876-914 zero credit, :825-848 e = 0): c in {0, 1, 2^128 - 1, l - 1, l (non-canonical zero)}; rnd all-zero (e = alpha = 0:
    Y_A and Y_G are the identity), all-ones, and e = 0 with a random alpha.  Returns (req, cs, rnd)."""
    req = base["req"].reshape(-1, 128).copy(); cs = base["cs"].reshape(-1, 32).copy(); rnd = base["rnd"].reshape(-1, 128).copy()
    n = len(req)
    cvals = [0, 1, (1 << 128) - 1, ELL - 1, ELL]
    for i in range(n):
        cs[i] = np.frombuffer(cvals[i % len(cvals)].to_bytes(32, "little"), np.uint8)
        m = (i // len(cvals)) % 4
        if m == 1: rnd[i] = 0
        elif m == 2: rnd[i] = 255
        elif m == 3: rnd[i, :64] = 0
    return req.reshape(-1), cs.reshape(-1), rnd.reshape(-1)


def mutate_proofs_head(ctx, base):
    """A second adversarial spend corpus aimed at the equations the engine computes differently from the reference
    (src/lib.rs:791-799, 806-809, 825-829): A1 without A-bar, the h2 terms of C'_00 / C'_01 added by the head stage, A2, C.
    Every class must fail the transcript comparison (status 7) unless noted; 255 = defer to the oracle.
    Returns (proofs, rnd, expected_status, labels)."""
    proofs = base["proofs"].reshape(-1, PROOF_BYTES).copy(); rnd = base["rnd"].reshape(-1, 128).copy()
    n = len(proofs)
    expect = np.zeros(n, np.uint8); labels = ["valid"] * n
    rs = np.random.RandomState(23)
    names = {133: "e_bar", 134: "r2_bar", 135: "r3_bar", 136: "c_bar", 137: "r_bar", 138: "w00", 139: "w01", 524: "k_bar", 525: "s_bar"}
    idxs = sorted(names)
    for i in range(n):
        pf = proofs[i]
        kind = i % 8
        if kind == 0:      # valid proof signed with degenerate randomness: e = alpha = 0, all-ones, or e = 0 only
            m = (i // 8) % 4
            if m == 1: rnd[i] = 0
            elif m == 2: rnd[i] = 255
            elif m == 3: rnd[i, :64] = 0
            labels[i] = "valid, rnd mode %d" % m
        if kind == 1:      # one of the response scalars of the sigma protocol += 1
            idx = idxs[(i // 8) % len(idxs)]
            _set(pf, idx, sc_bytes(sc_int(_get(pf, idx)) + 1)); expect[i] = 7; labels[i] = names[idx] + "+1"
        elif kind == 2:    # B_bar = identity: a VALID encoding (only A' is checked for identity, :787) -> transcript mismatch
            _set(pf, 3, bytes(32)); expect[i] = 7; labels[i] = "B identity"
        elif kind == 3:    # w00 / w01 = 0 or l-1: the h2 term of j = 0 vanishes or flips
            idx = 138 + (i // 8) % 2
            _set(pf, idx, sc_bytes(0 if (i // 16) % 2 else ELL - 1)); expect[i] = 7; labels[i] = names[idx] + " extreme"
        elif kind == 4:    # A' and B_bar swapped (both valid points)
            a, b = _get(pf, 2), _get(pf, 3); _set(pf, 2, b); _set(pf, 3, a); expect[i] = 7; labels[i] = "A'/B swapped"
        elif kind == 5:    # non-canonical encodings of exactly the scalars the head stage combines with x: accepted, same outputs
            for idx in (132, 133, 134, 138, 139):
                v = sc_int(_get(pf, idx)) + ELL
                if v < 2**256:
                    _set(pf, idx, v.to_bytes(32, "little"))
            labels[i] = "non-canonical head scalars"
        elif kind == 6:    # com[0] replaced by com[1] (the only commitment with h2 terms loses its own base)
            _set(pf, 4, _get(pf, 5)); expect[i] = 255; labels[i] = "com0 := com1"
        elif kind == 7:    # gamma0[0] / z[0][b] changed: the j = 0 pair that the head stage completes
            idx = (140, 268, 269)[(i // 8) % 3]
            _set(pf, idx, sc_bytes(sc_int(_get(pf, idx)) + 1 + rs.randint(0, 1 << 20))); expect[i] = 7; labels[i] = "j=0 scalar"
    return proofs.reshape(-1), rnd.reshape(-1), expect, labels


G_ENC = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")   # RFC 9496 generator


def _bump(rec, off, t=1):
    rec[off:off + 32] = np.frombuffer(sc_bytes(sc_int(rec[off:off + 32]) + t), np.uint8)


def _noncanon(rec, off):
    v = int.from_bytes(bytes(rec[off:off + 32]), "little") + ELL
    if v < 2**256:
        rec[off:off + 32] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)


def mutate_refunds(com, refunds):
    """Tampered Refund records for the client-side check PreRefund::to_credit_token (row a4): every component the
    reference's tests touch -- e + 1 (src/tests.rs:802-816), A + G, gamma + 1, z + 1 (:1176-1226) -> InvalidRefundProof --
    plus malformed / identity A, non-canonical scalar encodings (must accept) and commitments that are not the proof's.
    com: n*4096 (the proofs' com[128]), refunds: n*128 VALID refunds.  Returns (com, refunds, expected_status, labels)."""
    com = np.ascontiguousarray(com).reshape(-1, 4096).copy(); rf = np.ascontiguousarray(refunds).reshape(-1, 128).copy()
    n = len(rf)
    expect = np.zeros(n, np.uint8); labels = ["valid"] * n
    bad = bad_point_encodings()
    for i in range(n):
        kind = i % 10
        if kind == 1:
            _bump(rf[i], 32); expect[i] = 4; labels[i] = "e+1"
        elif kind == 2:
            rf[i, 0:32] = np.frombuffer(O.point_add(bytes(rf[i, 0:32]), G_ENC), np.uint8); expect[i] = 4; labels[i] = "A+G"
        elif kind == 3:
            _bump(rf[i], 64); expect[i] = 4; labels[i] = "gamma+1"
        elif kind == 4:
            _bump(rf[i], 96); expect[i] = 4; labels[i] = "z+1"
        elif kind == 5:
            rf[i, 0:32] = np.frombuffer(bad[(i // 10) % len(bad)], np.uint8); expect[i] = 0x81; labels[i] = "bad A"
        elif kind == 6:
            for off in (32, 64, 96):
                _noncanon(rf[i], off)
            labels[i] = "non-canonical scalars"
        elif kind == 7:   # the refund of ANOTHER spend (valid signature on a different K')
            rf[i] = rf[(i + 1) % n] if n > 1 else rf[i]; expect[i] = 4 if n > 1 else 0; labels[i] = "refund of another spend"
        elif kind == 8:
            j = (7 * i) % 127
            com[i, 32 * j:32 * j + 32], com[i, 32 * j + 32:32 * j + 64] = com[i, 32 * j + 32:32 * j + 64].copy(), com[i, 32 * j:32 * j + 32].copy()
            expect[i] = 4; labels[i] = "com[j], com[j+1] swapped"
        elif kind == 9:
            if (i // 10) % 2:
                com[i, 32 * 5:32 * 6] = np.frombuffer(bad[(i // 10 + 2) % len(bad)], np.uint8); expect[i] = 0x81; labels[i] = "bad com"
            else:
                rf[i, 0:32] = 0; expect[i] = 4; labels[i] = "A identity"
    return com.reshape(-1), rf.reshape(-1), expect, labels


def mutate_responses(base):
    """Tampered IssuanceResponse records for PreIssuance::to_credit_token (row a3): e + 1 (src/tests.rs:703-714), e = 0
    (:836-847), every other component the same way, a K that is not the request's, malformed points, non-canonical scalar
    encodings (must accept).  Returns (K, responses, expected_status, labels)."""
    K = base["req"].reshape(-1, 128)[:, :32].copy(); rs = base["resp"].reshape(-1, 160).copy()
    n = len(rs)
    expect = np.zeros(n, np.uint8); labels = ["valid"] * n
    bad = bad_point_encodings()
    for i in range(n):
        kind = i % 12
        if kind == 1:
            _bump(rs[i], 32); expect[i] = 2; labels[i] = "e+1"
        elif kind == 2:
            rs[i, 32:64] = 0; expect[i] = 2; labels[i] = "e=0"
        elif kind == 3:
            rs[i, 0:32] = np.frombuffer(O.point_add(bytes(rs[i, 0:32]), G_ENC), np.uint8); expect[i] = 2; labels[i] = "A+G"
        elif kind == 4:
            _bump(rs[i], 64); expect[i] = 2; labels[i] = "gamma+1"
        elif kind == 5:
            _bump(rs[i], 96); expect[i] = 2; labels[i] = "z+1"
        elif kind == 6:
            _bump(rs[i], 128); expect[i] = 2; labels[i] = "c+1"
        elif kind == 7:
            rs[i, 0:32] = np.frombuffer(bad[(i // 12) % len(bad)], np.uint8); expect[i] = 0x81; labels[i] = "bad A"
        elif kind == 8:
            for off in (32, 64, 96, 128):
                _noncanon(rs[i], off)
            labels[i] = "non-canonical scalars"
        elif kind == 9:
            K[i] = K[(i + 1) % n] if n > 1 else K[i]; expect[i] = 2 if n > 1 else 0; labels[i] = "K of another request"
        elif kind == 10:
            K[i] = np.frombuffer(bad[(i // 12 + 5) % len(bad)], np.uint8); expect[i] = 0x81; labels[i] = "bad K"
        elif kind == 11:
            rs[i, 0:32] = 0; expect[i] = 2; labels[i] = "A identity"
    return K.reshape(-1), rs.reshape(-1), expect, labels


def tampered_token_proofs(ctx, n, seed=b"token-tamper"):
    """prop_token_tampering_detection (src/tests.rs:1898-1927): a valid token whose `a` and `e` are overwritten with a random
    point and scalar still produces a SpendProof, and refund() must answer InvalidClientSpendProof; variants with only one of
    the two replaced, and with a = identity (A' = identity -> IdentityPointError, src/lib.rs:787-789).
    Returns dict(proofs, rnd, expect, labels)."""
    base = gen_valid(ctx, n, seed=seed, threads=min(os.cpu_count() or 1, 8))
    st = trip_streams(seed, n)
    tokens = tokens_from(base, st["pre"]).reshape(n, 160).copy()
    expect = np.zeros(n, np.uint8); labels = ["valid token"] * n
    for i in range(n):
        kind = i % 5
        rp = O.scalarmult_base(O.sc_reduce64(xof(seed + b"/pt/%d" % i, 64)))
        re = O.sc_reduce64(xof(seed + b"/sc/%d" % i, 64))
        if kind == 1:
            tokens[i, 0:32] = np.frombuffer(rp, np.uint8); tokens[i, 32:64] = np.frombuffer(re, np.uint8); expect[i] = 7; labels[i] = "a, e random"
        elif kind == 2:
            tokens[i, 0:32] = np.frombuffer(rp, np.uint8); expect[i] = 7; labels[i] = "a random"
        elif kind == 3:
            tokens[i, 32:64] = np.frombuffer(re, np.uint8); expect[i] = 7; labels[i] = "e random"
        elif kind == 4:
            tokens[i, 0:32] = 0; expect[i] = 6; labels[i] = "a identity"
    one = (1).to_bytes(32, "little")
    proofs = b"".join(ctx.prove_spend(tokens[i].tobytes(), one, xof(seed + b"/prove/%d" % i, O.RND_PROVE))[0] for i in range(n))
    return dict(proofs=np.frombuffer(proofs, np.uint8).copy(), rnd=np.frombuffer(xof(seed + b"/rnd", 128 * n), np.uint8).copy(),
                expect=expect, labels=labels, tokens=tokens.reshape(-1))


def overspend_proofs(ctx, n, seed=b"overspend"):
    """prove_spend run with s > c (src/tests.rs:366-374,1540-1547) -> InvalidClientSpendProof."""
    rs = np.random.RandomState(5)
    seeds = np.frombuffer(xof(seed, 64 * n), dtype=np.uint8).copy()
    cr = rs.randint(5, 500, size=n).astype(np.uint64)
    ch = (cr + rs.randint(1, 100, size=n).astype(np.uint64)).astype(np.uint64)
    out = ctx.generate(seeds, cr, ch, threads=min(os.cpu_count() or 1, 16))
    out["rnd"] = np.frombuffer(xof(seed + b"/rnd", 128 * n), dtype=np.uint8).copy()
    return out


def trip_streams(seed, n):
    """The per-trip randomness gen_valid(seed=...) hands to the oracle (oracle/act_oracle.c gen_worker): for trip i the
    stream BLAKE3-XOF(seeds_i) gives PreIssuance (r, k), the request's 128 bytes, issue's 128 and prove_spend's 33 536.
    Returns dict(pre n*64 [r|k reduced], req_rnd n*128, issue_rnd n*128, prove_rnd n*33536)."""
    seeds = xof(seed + b"/seeds", 64 * n)
    pre, rq, isr, pv = [], [], [], []
    for i in range(n):
        st = O.blake3(seeds[64 * i:64 * i + 64], 64 * 530)
        pre.append(O.sc_reduce64(st[0:64]) + O.sc_reduce64(st[64:128])); rq.append(st[128:256]); isr.append(st[256:384]); pv.append(st[384:])
    f = lambda parts: np.frombuffer(b"".join(parts), np.uint8).copy()
    return dict(pre=f(pre), req_rnd=f(rq), issue_rnd=f(isr), prove_rnd=f(pv))


def tokens_from(base, pre):
    """CreditToken records (A | e | k | r | c, 160 B) of the trips in a gen_valid() corpus (src/lib.rs:555-561);
    pre = trip_streams(...)["pre"]."""
    n = len(base["resp"]) // 160
    resp = base["resp"].reshape(n, 160); pr = pre.reshape(n, 64); cs = base["cs"].reshape(n, 32)
    return np.concatenate([resp[:, :64], pr[:, 32:64], pr[:, :32], cs], axis=1).reshape(-1).copy()


def charges_from(base):
    return np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in base["charges"]), np.uint8).copy()


def prover_stream(seed32: bytes, index: int) -> bytes:
    """The derived RNG stream of act_batch_prove_spend: BLAKE3-XOF(seed || u64le(index)), 524 x 64 bytes."""
    import blake3
    return blake3.blake3(bytes(seed32) + int(index).to_bytes(8, "little")).digest(length=524 * 64)


# ---- multi-generation token lifecycles (the scenario tests of src/tests.rs) ----
# (label, credits at issuance, amounts spent one after the other).  A spend above the balance is produced by the prover
# and rejected by the issuer (InvalidClientSpendProof); the client then keeps its old token.
LIFECYCLES = [
    ("sequential_spends", 100, [30, 20, 50, 1]),                                  # src/tests.rs:260-337 (+ one spend from the emptied token)
    ("spend_exact_balance", 50, [50, 0]),                                         # :210-257
    ("zero_spend_scenario", 100, [0, 100]),                                       # :378-426
    ("token_with_zero_credit", 0, [7, 0, 0]),                                     # :876-914
    ("exhaust_token_with_one_credit_spends", 5, [1, 1, 1, 1, 1, 1]),              # :917-1005
    ("large_amount_issuance", 2**120 + 0x1234567890abcdef, [2**119 + 99, 2**119 - 5, 2**60]),   # :642-689
    ("test_binary_decomposition_max_value", 2**128 - 1, [1, 2**128 - 2, 1]),      # :1008-1059
    ("attempt_overspend", 10, [11, 2**127, 10]),                                  # :340-375, then the honest spend still works
]


class OracleImpl:
    """The oracle behind the batch-shaped interface run_lifecycles() drives (also satisfied by hostsim_lib.Ctx and by
    the GPU engine through EngineImpl)."""

    def __init__(self, ctx):
        self.c = ctx

    def request(self, pre, rnd):
        n = len(pre) // 64
        return np.frombuffer(b"".join(self.c.request(pre[64 * i:64 * i + 64], rnd[128 * i:128 * i + 128]) for i in range(n)), np.uint8)

    def issue(self, req, cs, rnd):
        resp, st, _ = self.c.batch_issue(np.ascontiguousarray(req), np.ascontiguousarray(cs), np.ascontiguousarray(rnd))
        return resp, st

    def issuance_check(self, K, resp):
        return self.c.batch_issuance_check(np.ascontiguousarray(K), np.ascontiguousarray(resp))[0]

    def prove_spend(self, tokens, charges, rnd):
        n = len(tokens) // 160
        out = [self.c.prove_spend(tokens[160 * i:160 * i + 160], charges[32 * i:32 * i + 32], rnd[O.RND_PROVE * i:O.RND_PROVE * (i + 1)])
               for i in range(n)]
        f = lambda k: np.frombuffer(b"".join(o[k] for o in out), np.uint8)
        return f(0), f(1), np.zeros(n, np.uint8)

    def refund(self, proofs, rnd):
        ref, nul, st, _ = self.c.batch_refund(np.ascontiguousarray(proofs), np.ascontiguousarray(rnd))
        return ref, nul, st

    def refund_check(self, com, refund):
        return self.c.batch_refund_check(np.ascontiguousarray(com), np.ascontiguousarray(refund))[0]


class EngineImpl:
    """The GPU engine (host-buffer C ABI calls) behind the same interface."""

    def __init__(self, engine):
        self.e = engine

    def request(self, pre, rnd):
        return self.e.batch_request(pre, rnd)

    def issue(self, req, cs, rnd):
        return self.e.batch_issue(req, cs, rnd)

    def issuance_check(self, K, resp):
        return self.e.batch_issuance_check(K, resp)

    def prove_spend(self, tokens, charges, rnd):
        return self.e.batch_prove_spend(tokens, charges, rnd=rnd)

    def refund(self, proofs, rnd):
        return self.e.batch_verify_spend_and_refund(proofs, rnd)

    def refund_check(self, com, refund):
        return self.e.batch_refund_check(com, refund)


def run_lifecycles(impl, seed=b"lifecycle-0", scenarios=LIFECYCLES):
    """Drives every scenario through request -> issue -> issuance_check and then, generation after generation,
    prove_spend -> refund -> refund_check -> new token (A* | e* | k* | r* | m; PreRefund::to_credit_token,
    src/lib.rs:1245-1252).  All scenarios advance together, one batch per generation.  Asserts the outcomes the reference's
    tests assert (accept iff s <= balance, new balance = c - s, fresh nullifier every generation) and returns every byte the
    issuer or the client produced, so that two implementations can be compared for equality."""
    u8 = lambda b: np.frombuffer(bytes(b), np.uint8).copy()
    n = len(scenarios)
    le32 = lambda vals: u8(b"".join(int(v).to_bytes(32, "little") for v in vals))
    pre = u8(b"".join(O.sc_reduce64(xof(seed + b"/r/%d" % i, 64)) + O.sc_reduce64(xof(seed + b"/k/%d" % i, 64)) for i in range(n)))
    req = np.asarray(impl.request(pre, u8(xof(seed + b"/req", 128 * n))))
    cs = le32([c for _, c, _ in scenarios])
    resp, st = impl.issue(req, cs, u8(xof(seed + b"/issue", 128 * n)))
    resp = np.asarray(resp)
    assert (np.asarray(st) == 0).all(), st
    assert (np.asarray(impl.issuance_check(req.reshape(n, 128)[:, :32].reshape(-1).copy(), resp)) == 0).all()
    r2 = resp.reshape(n, 160); p2 = pre.reshape(n, 64)
    tokens = [bytes(r2[i, :64]) + bytes(p2[i, 32:64]) + bytes(p2[i, :32]) + bytes(cs[32 * i:32 * i + 32]) for i in range(n)]
    balance = [c for _, c, _ in scenarios]
    seen = set()
    log = dict(req=req.tobytes(), resp=resp.tobytes(), generations=[])
    for g in range(max(len(s) for _, _, s in scenarios)):
        live = [i for i in range(n) if g < len(scenarios[i][2])]
        amounts = [scenarios[i][2][g] for i in live]
        m = len(live)
        proofs, prer, pst = impl.prove_spend(u8(b"".join(tokens[i] for i in live)), le32(amounts),
                                             u8(xof(seed + b"/prove/%d" % g, O.RND_PROVE * m)))
        proofs, prer = np.asarray(proofs), np.asarray(prer)
        assert (np.asarray(pst) == 0).all()
        ref, nul, vst = impl.refund(proofs, u8(xof(seed + b"/refund/%d" % g, 128 * m)))
        ref, nul, vst = np.asarray(ref), np.asarray(nul), np.asarray(vst)
        com = proofs.reshape(m, PROOF_BYTES)[:, 128:128 + 4096].reshape(-1).copy()
        cst = np.asarray(impl.refund_check(com, ref))
        for j, i in enumerate(live):
            label, s = scenarios[i][0], amounts[j]
            if s <= balance[i]:
                assert vst[j] == 0 and cst[j] == 0, (label, g, int(vst[j]), int(cst[j]))
                k = bytes(nul[32 * j:32 * j + 32])
                assert k == tokens[i][64:96] and k not in seen, (label, g)       # nullifier() = the spent token's k, never seen before
                seen.add(k)
                q = bytes(prer[96 * j:96 * j + 96])
                balance[i] -= s
                assert int.from_bytes(q[64:96], "little") == balance[i], (label, g)
                tokens[i] = bytes(ref[128 * j:128 * j + 64]) + q
            else:
                assert vst[j] == 7, (label, g, int(vst[j]))                       # InvalidClientSpendProof
                assert not ref[128 * j:128 * j + 128].any() and not nul[32 * j:32 * j + 32].any()
                assert cst[j] != 0                                                 # an all-zero refund never verifies
        log["generations"].append(dict(proofs=proofs.tobytes(), prerefunds=prer.tobytes(), refunds=ref.tobytes(),
                                       nullifiers=nul.tobytes(), status=vst.tobytes(), check=cst.tobytes()))
    log["final_balances"] = balance
    return log
