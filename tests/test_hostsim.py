"""The CUDA engine's device headers compiled for the host (tests/hostsim) vs the oracle: same per-thread kernel
bodies, portable arithmetic paths.  Covers the engine's LOGIC without a GPU; the PTX paths and the launch plumbing
are covered by the -m gpu tests."""
import os
import random

import numpy as np

import corpus
import hostsim_lib as HS
import oracle_lib as O

ELL = corpus.ELL
P = corpus.P25519
b32 = lambda v: v.to_bytes(32, "little")


def test_field_arithmetic_loose_inputs():
    rnd = random.Random(1)
    edge = [0, 1, 19, 38, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 37, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 39]
    vals = edge + [rnd.getrandbits(256) for _ in range(60)]
    for a in vals:
        for b in vals[::5]:
            assert int.from_bytes(HS.call1("hs_fe_mul", b32(a), b32(b)), "little") == a * b % P
            keep = [HS._in(b32(a)), HS._in(b32(b))]
            s, sp = HS._out(32); d, dp = HS._out(32)
            HS.lib().hs_fe_addsub(keep[0][1], keep[1][1], sp, dp)
            assert int.from_bytes(s.tobytes(), "little") == (a + b) % P
            assert int.from_bytes(d.tobytes(), "little") == (a - b) % P
        if a % P:
            assert int.from_bytes(HS.call1("hs_fe_invert", b32(a)), "little") == pow(a, -1, P)


def test_products_are_tight_and_tight_sums_are_exact():
    """fe_mul / fe_sq leave values below 2^255 + 2^11 whatever the (loose) inputs; the one-pass sum and double of two such
    values (fe_add_tt, fe_dbl_tt) are below 2^256 and congruent to the true sum / double."""
    import ctypes as C
    rnd = random.Random(2)
    edge = [0, 1, 19, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 37, 2**255 - 1, 2**255, 2**255 + 2047, 2**256 - 1, 2**256 - 38]
    # inputs whose product is = -1, -2, ... (mod p) push the result to the top of the range: a * a^-1 * (p - k)
    near = []
    for k in (1, 2, 19, 20, 37, 38, 39):
        a = rnd.getrandbits(255) % P or 3
        near.append((a, pow(a, -1, P) * (P - k) % P))
    pairs = [(a, b) for a in edge for b in edge] + near + [(rnd.getrandbits(256), rnd.getrandbits(256)) for _ in range(400)]
    lib = HS.lib()
    lib.hs_fe_tight.restype = C.c_int
    lib.hs_fe_tight.argtypes = [C.c_void_p] * 6
    for a, b in pairs:
        ia, ib = HS._in(b32(a)), HS._in(b32(b))
        outs = [HS._out(32) for _ in range(4)]
        tight = lib.hs_fe_tight(ia[1], ib[1], *[o[1] for o in outs])
        m, q, sm, d = [int.from_bytes(o[0].tobytes(), "little") for o in outs]
        assert tight == 1 and m < 2**255 + 2**11 and q < 2**255 + 2**11
        assert m % P == a * b % P and q % P == a * a % P
        assert sm < 2**256 and sm % P == (m + q) % P
        assert d < 2**256 and d % P == 2 * m % P


def test_scalar_arithmetic():
    rnd = random.Random(2)
    for it in range(100):
        x, y, z = (rnd.getrandbits(256) for _ in range(3))
        if it == 0:
            x, y, z = 2**256 - 1, 2**256 - 1, 2**256 - 1
        if it == 1:
            x, y, z = 0, ELL, ELL - 1
        assert int.from_bytes(HS.call1("hs_sc_reduce32", b32(x)), "little") == x % ELL
        w = rnd.getrandbits(512) if it else 2**512 - 1
        assert int.from_bytes(HS.call1("hs_sc_reduce64", w.to_bytes(64, "little")), "little") == w % ELL
        assert int.from_bytes(HS.call1("hs_sc_muladd", b32(x), b32(y), b32(z)), "little") == (x * y + z) % ELL
        inv = int.from_bytes(HS.call1("hs_sc_invert", b32(x)), "little")
        assert inv == (pow(x % ELL, -1, ELL) if x % ELL else 0)
        keep = [HS._in(b32(x)), HS._in(b32(y))]
        n, np_ = HS._out(32); d, dp = HS._out(32)
        HS.lib().hs_sc_negsub(keep[0][1], keep[1][1], np_, dp)
        assert int.from_bytes(n.tobytes(), "little") == (-x) % ELL and int.from_bytes(d.tobytes(), "little") == (x - y) % ELL
        # signed window recoding reconstructs the scalar
        for wbits, nd in ((4, 64), (8, 32)):
            a, ap = HS._in(b32(x)); dg = np.zeros(64, np.int8)
            HS.lib().hs_recode(ap, wbits, dg.ctypes.data)
            assert sum(int(dg[i]) << (wbits * i) for i in range(nd)) == x % ELL
            assert all(-(1 << (wbits - 1)) <= int(v) < (1 << (wbits - 1)) for v in dg[:nd])


def test_ristretto_codec_and_mults(octx):
    rnd = random.Random(4)
    hs = HS.Ctx(octx.h, octx.x, octx.w)
    for b in corpus.bad_point_encodings():
        a, ap = HS._in(b); o, op = HS._out(32)
        assert HS.lib().hs_decode_encode(ap, op) == 0
    for it in range(12):
        u = bytes(rnd.getrandbits(8) for _ in range(64))
        Pt = O.from_uniform(u)
        assert HS.call1("hs_from_uniform", u) == Pt
        a, ap = HS._in(Pt); o, op = HS._out(32)
        assert HS.lib().hs_decode_encode(ap, op) == 1 and o.tobytes() == Pt
        s = b32(rnd.getrandbits(256) % ELL if it else ELL - 1)
        for ct in (0, 1):
            k, kp = HS._in(s); o, op = HS._out(32)
            assert HS.lib().hs_scalarmult(kp, ap, ct, op) == 1 and o.tobytes() == O.scalarmult(s, Pt)
        assert hs.scalarmult_base(0, s, ct=0) == O.scalarmult_base(s) == hs.scalarmult_base(0, s, ct=1)
        for base in (1, 2, 3):
            assert hs.scalarmult_base(base, s) == O.scalarmult(s, octx.h[32 * (base - 1):32 * base])
    z = bytes(32)
    assert hs.scalarmult_base(0, z) == z and hs.scalarmult_base(0, z, ct=1) == z    # identity
    assert hs.public_key(octx.x) == octx.w


def test_blake3_single_chunk():
    rnd = random.Random(6)
    for n in [0, 1, 63, 64, 65, 186, 266, 425, 466, 1000, 1023, 1024]:
        data = bytes(rnd.getrandbits(8) for _ in range(n))
        a, ap = HS._in(data) if n else (None, None)
        o, op = HS._out(64)
        HS.lib().hs_blake3_small(ap, n, op)
        assert o.tobytes() == O.blake3(data, 64)


def test_params_derive():
    for p in (corpus.TEST_PARAMS, corpus.BENCH_PARAMS, ("o", "s", "d", "v" * 300)):
        assert HS.params_derive(*p) == O.params_derive(*p)


def test_issue_refund_parity_with_mutations(octx):
    hs = HS.Ctx(octx.h, octx.x, octx.w)
    base = corpus.gen_valid(octx, 24, seed=b"hostsim", threads=8)
    req, cs, rnd, expect, labels = corpus.mutate_requests(octx, base)
    resp, st = hs.issue(req, cs, rnd)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    assert st.tolist() == o_st.tolist() and (resp == o_resp).all()
    K = base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1)
    r2 = base["resp"].reshape(-1, 160).copy()
    r2[1, 32:64] = 0; r2[2, 96:97] ^= 1
    st = hs.issuance_check(K, r2.reshape(-1)); o_st, _ = octx.batch_issuance_check(K, r2.reshape(-1), threads=8)
    assert st.tolist() == o_st.tolist() and st[0] == 0 and st[1] == 2 and st[2] == 2
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    ref, nul, st = hs.refund(proofs, rnd)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    assert st.tolist() == o_st.tolist() and (ref == o_ref).all() and (nul == o_nul).all()
    com = proofs.reshape(-1, corpus.PROOF_BYTES)[:, 128:128 + 4096].copy().reshape(-1)
    ref2 = ref.reshape(-1, 128).copy(); ref2[0, 64] ^= 1
    st2 = hs.refund_check(com, ref2.reshape(-1)); o_st2, _ = octx.batch_refund_check(com, ref2.reshape(-1), threads=8)
    assert st2.tolist() == o_st2.tolist() and st2[0] == 4


def test_head_stage_corpus(octx):
    """Mutations aimed at the equations the engine reformulates (A1 without A-bar, h2 terms in the head stage, A2, C):
    host build of the device code == oracle, and every class carries the reference's error."""
    base = corpus.gen_valid(octx, 24, seed=b"hostsim-head", threads=8)
    hs = HS.Ctx(octx.h, octx.x, octx.w)
    proofs, rnd, expect, labels = corpus.mutate_proofs_head(octx, base)
    ref, nul, st = hs.refund(proofs, rnd)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    assert st.tolist() == o_st.tolist()
    assert (ref == o_ref).all() and (nul == o_nul).all()
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    assert set(st.tolist()) == {0, 7}


def test_issue_edge_credits_and_degenerate_randomness(octx):
    base = corpus.gen_valid(octx, 20, seed=b"hostsim-edge", threads=8)
    hs = HS.Ctx(octx.h, octx.x, octx.w)
    req, cs, rnd = corpus.edge_issue_inputs(base)
    resp, st = hs.issue(req, cs, rnd)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    assert st.tolist() == o_st.tolist() == [0] * 20 and (resp == o_resp).all()
    K = base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1)
    assert hs.issuance_check(K, resp).tolist() == octx.batch_issuance_check(K, resp, threads=8)[0].tolist()


def test_golden_corpus_small():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corpus_small.npz"))
    hs = HS.Ctx(g["h"].tobytes(), g["x"].tobytes(), g["w"].tobytes())
    resp, st = hs.issue(g["req"], g["cs"], g["rnd_issue"])
    assert (st == g["status_issue"]).all() and (resp == g["resp"]).all()
    ref, nul, st = hs.refund(g["proofs"], g["rnd"])
    assert (st == g["status"]).all() and (ref == g["refunds"]).all() and (nul == g["nullifiers"]).all()


def test_golden_corpus_checks():
    """tests/golden/corpus_checks.npz (statuses and outputs computed by the independent stack): the tamper classes of the two
    client-side checks and the proofs made from tampered tokens, through the host build of the device code."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corpus_checks.npz"))
    hs = HS.Ctx(g["h"].tobytes(), g["x"].tobytes(), g["w"].tobytes())
    assert (hs.issuance_check(g["K"], g["responses"]) == g["status_issuance_check"]).all()
    assert (hs.refund_check(g["com"], g["refunds"]) == g["status_refund_check"]).all()
    ref, nul, st = hs.refund(g["token_proofs"], g["token_rnd"])
    assert (st == g["token_status"]).all() and (ref == g["token_refunds"]).all() and (nul == g["token_nullifiers"]).all()


def test_batched_double_and_encode_stage_with_identity_points():
    """Stage 1b (Montgomery-batched, square-root-free encode of 2P) vs the oracle's encode(2*P), including identity
    points inside a batch (e*g*f*h = 0 must not poison the other 15 points of the batch)."""
    rnd = random.Random(8)
    encs = []
    for i in range(256):
        if i in (0, 5, 16, 17, 255) or (32 <= i < 48):
            encs.append(bytes(32))                         # identity; one whole batch of identities too
        else:
            encs.append(O.scalarmult_base(b32(rnd.getrandbits(252))))
    a, ap = HS._in(b"".join(encs)); o, op = HS._out(256 * 32)
    assert HS.lib().hs_encode_stage(ap, op) == 1
    two = b32(2)
    for i in range(256):
        exp = O.scalarmult(two, encs[i])
        assert o[32 * i:32 * i + 32].tobytes() == exp, i
    for it in range(50):
        x = rnd.getrandbits(256)
        h = int.from_bytes(HS.call1("hs_sc_half", b32(x)), "little")
        assert h < ELL and (2 * h - x) % ELL == 0


def test_cbor_skeletons_match_host_encoder_and_golden(act):
    """The canonical skeletons the device fast path uses (act_aux.cuh) reproduce the host encoder byte for byte (which the
    golden trip pins to ciborium's layout), invert it, and flag anything else as not canonical."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trip.json")))
    recs = {0: bytes.fromhex(g["request"]), 1: bytes.fromhex(g["response"]), 2: bytes.fromhex(g["proof"]), 3: bytes.fromhex(g["refund"])}
    enc = {0: act.encode_issuance_request_cbor, 1: act.encode_issuance_response_cbor, 2: act.encode_spend_proof_cbor, 3: act.encode_refund_cbor}
    for kind, rec in recs.items():
        want = enc[kind](rec)
        got = HS.cbor_skeleton_encode(kind, rec, act.CBOR_BYTES[kind])
        assert got == want and len(got) == act.CBOR_BYTES[kind]
        back, st = HS.cbor_skeleton_unpack(kind, got, act.RECORD_BYTES[kind])
        assert st == 0 and back == rec
        bad = bytearray(got); bad[0] ^= 1                       # wrong map header
        assert HS.cbor_skeleton_unpack(kind, bytes(bad), act.RECORD_BYTES[kind])[1] == 0xFF
        bad = bytearray(got); bad[2] = 0x59                     # a bstr head that is not 58 20
        assert HS.cbor_skeleton_unpack(kind, bytes(bad), act.RECORD_BYTES[kind])[1] == 0xFF
    assert HS.cbor_skeleton_encode(0, recs[0], 141).hex() == g["cbor_request"]
    assert HS.cbor_skeleton_encode(3, recs[3], 141).hex() == g["cbor_refund"]


def test_client_generators_match_the_oracle_prover(octx):
    """act_prove.cuh (PreIssuance::request, CreditToken::prove_spend) on the host build: bit-exact with the oracle's
    line-by-line prover on identical RNG bytes, and the derived-stream mode equals the explicit mode on those bytes."""
    n = 3
    base = corpus.gen_valid(octx, n, seed=b"prover-parity", threads=2)
    st = corpus.trip_streams(b"prover-parity", n)
    hs = HS.Ctx(octx.h, octx.x, octx.w)
    assert (hs.request(st["pre"], st["req_rnd"]) == base["req"]).all()
    tokens, charges = corpus.tokens_from(base, st["pre"]), corpus.charges_from(base)
    proofs, prer, status = hs.prove_spend(tokens, charges, rnd=st["prove_rnd"])
    assert (status == 0).all() and (proofs == base["proofs"]).all() and (prer == base["prerefund"]).all()
    seed = corpus.xof(b"derived-seed", 32)
    rnd = np.frombuffer(b"".join(corpus.prover_stream(seed, 7 + i) for i in range(n)), np.uint8)
    p1, r1, s1 = hs.prove_spend(tokens, charges, seed=seed, first_index=7)
    p2, r2, s2 = hs.prove_spend(tokens, charges, rnd=rnd)
    assert (p1 == p2).all() and (r1 == r2).all()
    ref, nul, vst = hs.refund(p1, base["rnd"])
    assert (vst == 0).all()
    bad = tokens.copy(); bad[:32] = np.frombuffer(corpus.bad_point_encodings()[0], np.uint8)
    p3, r3, s3 = hs.prove_spend(bad, charges, seed=seed)
    assert s3.tolist() == [0x81, 0, 0] and not p3[:corpus.PROOF_BYTES].any() and (p3[corpus.PROOF_BYTES:] != 0).any()


def test_token_lifecycles_match_the_oracle(octx):
    """Multi-generation token chains (corpus.LIFECYCLES, the reference's scenario tests) on the host build of the device
    code: every proof, refund, nullifier and status equals the oracle's."""
    hs = HS.Ctx(octx.h, octx.x, octx.w)

    class Impl:
        request = staticmethod(hs.request); issue = staticmethod(hs.issue); issuance_check = staticmethod(hs.issuance_check)
        refund = staticmethod(hs.refund); refund_check = staticmethod(hs.refund_check)
        prove_spend = staticmethod(lambda tokens, charges, rnd: hs.prove_spend(tokens, charges, rnd=rnd))

    assert corpus.run_lifecycles(Impl) == corpus.run_lifecycles(corpus.OracleImpl(octx))


def test_cbor_skeleton_check_is_exact_at_every_byte(act):
    """Fast-path contract (act_aux.cuh, host build): a one-bit change at ANY structural byte of a canonical item makes the
    skeleton check hand the item back (0xFF); a change inside a payload is still canonical, and then the record equals what
    the lenient host parser (the reference's from_cbor semantics) extracts from the same bytes."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trip.json")))
    recs = {0: bytes.fromhex(g["request"]), 1: bytes.fromhex(g["response"]), 2: bytes.fromhex(g["proof"]), 3: bytes.fromhex(g["refund"])}
    enc = {0: act.encode_issuance_request_cbor, 1: act.encode_issuance_response_cbor, 2: act.encode_spend_proof_cbor, 3: act.encode_refund_cbor}
    pack = {0: act.pack_issuance_requests_cbor, 1: act.pack_issuance_responses_cbor, 2: act.pack_spend_proofs_cbor, 3: act.pack_refunds_cbor}
    rs = np.random.RandomState(5)
    for kind, rec in recs.items():
        nrec = act.RECORD_BYTES[kind]
        lo, hi = enc[kind](bytes(nrec)), enc[kind](b"\xff" * nrec)
        payload = np.frombuffer(lo, np.uint8) != np.frombuffer(hi, np.uint8)
        assert payload.sum() == nrec
        item = enc[kind](rec)
        positions = range(len(item)) if kind != 2 else sorted(set(np.nonzero(~payload)[0].tolist()) | set(rs.randint(0, len(item), 400).tolist()))
        accepted = []
        for pos in positions:
            b = bytearray(item); b[pos] ^= 1 << rs.randint(8)
            back, st = HS.cbor_skeleton_unpack(kind, bytes(b), nrec)
            assert st == (0 if payload[pos] else 0xFF), (kind, pos)
            if st == 0:
                accepted.append((bytes(b), back))
        host_rec, host_st = pack[kind]([a for a, _ in accepted])
        assert (host_st == 0).all() and host_rec.tobytes() == b"".join(r for _, r in accepted)


def test_window_form_of_the_range_pair_is_bit_exact(octx):
    """ACT_RANGE_BUCKETS=0 (the A/B alternative; the product and the default host build use the right-to-left bucket form):
    the left-to-right window evaluation of the range-proof pair gives the same refunds, nullifiers and statuses as the oracle
    on the mutation corpus."""
    import ctypes as C
    HS.lib()     # builds both host libraries
    L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(HS.__file__)), "hostsim", "libact_hostsim_windows.so"))
    vp, sz = C.c_void_p, C.c_size_t
    L.hs_ctx_create.argtypes = [vp, vp, vp]; L.hs_ctx_create.restype = vp
    L.hs_ctx_destroy.argtypes = [vp]
    L.hs_refund.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    base = corpus.gen_valid(octx, 12, seed=b"bucket-form", threads=4)
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    n = len(expect)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=4)
    buf = lambda b: np.frombuffer(bytes(b), np.uint8).copy()
    h, x, w = buf(octx.h), buf(octx.x), buf(octx.w)
    ctx = L.hs_ctx_create(h.ctypes.data, x.ctypes.data, w.ctypes.data)
    assert ctx
    ref = np.zeros(n * 128, np.uint8); nul = np.zeros(n * 32, np.uint8); st = np.zeros(n, np.uint8)
    p = np.ascontiguousarray(proofs); r = np.ascontiguousarray(rnd)
    L.hs_refund(ctx, n, p.ctypes.data, r.ctypes.data, ref.ctypes.data, nul.ctypes.data, st.ctypes.data)
    L.hs_ctx_destroy(ctx)
    assert (st == o_st).all() and (ref == o_ref).all() and (nul == o_nul).all()
    assert (st == 0).sum() >= 2 and (st != 0).sum() >= 2
