"""GPU parity tests proper: the CUDA engine through the C ABI vs the CPU oracle on identical inputs.
Bar: bit-exact (integer/byte work)."""
import numpy as np
import pytest

import corpus
import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_selftest(act):
    assert act.device_count() >= 1
    act.selftest(0)


def test_params_and_key(act, octx):
    for p in (corpus.TEST_PARAMS, corpus.BENCH_PARAMS, ("example-org", "payment-api", "production", "2024-01-15")):
        assert act.Params.new(*p).h == O.params_derive(*p)
    assert act.PrivateKey.from_secret(octx.x).w == octx.w


@pytest.fixture(scope="module")
def base(octx):
    return corpus.gen_valid(octx, 96, seed=b"gpu-parity")


def test_issue_valid_and_mutated(engine, octx, base):
    req, cs, rnd, expect, labels = corpus.mutate_requests(octx, base)
    resp, st = engine.batch_issue(req, cs, rnd)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    assert st.tolist() == o_st.tolist()
    assert (resp == o_resp).all()
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    assert (st == 0).sum() > 10 and (st != 0).sum() > 10


def test_issuance_check(engine, octx, base):
    K = base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1)
    resp = base["resp"].reshape(-1, 160).copy()
    n = len(resp)
    ell = corpus.ELL
    for i in range(n):
        if i % 4 == 1:   # e += 1 (src/tests.rs:703-714)
            resp[i, 32:64] = np.frombuffer(corpus.sc_bytes(corpus.sc_int(resp[i, 32:64]) + 1), np.uint8)
        if i % 4 == 2:   # e = 0 (src/tests.rs:836-847)
            resp[i, 32:64] = 0
        if i % 4 == 3 and i % 8 == 3:
            resp[i, 0:32] = np.frombuffer(corpus.bad_point_encodings()[i % 5], np.uint8)
    resp = resp.reshape(-1)
    st = engine.batch_issuance_check(K, resp)
    o_st, _ = octx.batch_issuance_check(K, resp, threads=8)
    assert st.tolist() == o_st.tolist()
    assert set(st.tolist()) == {0, 2, 0x81}


def test_spend_refund_valid_and_mutated(engine, octx, base):
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    ref, nul, st = engine.batch_verify_spend_and_refund(proofs, rnd)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    assert st.tolist() == o_st.tolist()
    assert (ref == o_ref).all() and (nul == o_nul).all()
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    # refund_check on what the engine produced (accepted ones) + tampering (src/tests.rs:802-816,1176-1226)
    com = proofs.reshape(-1, corpus.PROOF_BYTES)[:, 128:128 + 4096].copy().reshape(-1)
    ref2 = ref.reshape(-1, 128).copy()
    for i in range(len(ref2)):
        if st[i] == 0 and i % 3 == 1:
            ref2[i, 96:128] = np.frombuffer(corpus.sc_bytes(corpus.sc_int(ref2[i, 96:128]) + 1), np.uint8)  # z += 1
    st2 = engine.batch_refund_check(com, ref2.reshape(-1))
    o_st2, _ = octx.batch_refund_check(com, ref2.reshape(-1), threads=8)
    assert st2.tolist() == o_st2.tolist()
    assert 0 in st2.tolist() and 4 in st2.tolist()


def test_head_stage_corpus(engine, octx):
    """Mutations aimed at the equations the engine computes differently from the reference (A1 without A-bar, the h2 terms of
    C'_00 / C'_01 added by the head kernel, A2, C): statuses, refunds and nullifiers equal the oracle's bit for bit."""
    base = corpus.gen_valid(octx, 48, seed=b"gpu-head", threads=8)
    proofs, rnd, expect, labels = corpus.mutate_proofs_head(octx, base)
    ref, nul, st = engine.batch_verify_spend_and_refund(proofs, rnd)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    assert st.tolist() == o_st.tolist()
    assert (ref == o_ref).all() and (nul == o_nul).all()
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    assert set(st.tolist()) == {0, 7}


def test_issue_edge_credits_and_degenerate_randomness(engine, octx):
    """c in {0, 1, 2^128-1, l-1, l}; signer randomness all-zero (e = alpha = 0: identity Y_A, Y_G), all-ones, e = 0."""
    base = corpus.gen_valid(octx, 40, seed=b"gpu-edge", threads=8)
    req, cs, rnd = corpus.edge_issue_inputs(base)
    resp, st = engine.batch_issue(req, cs, rnd)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    assert st.tolist() == o_st.tolist() == [0] * 40 and (resp == o_resp).all()
    K = base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1)
    assert engine.batch_issuance_check(K, resp).tolist() == octx.batch_issuance_check(K, resp, threads=8)[0].tolist()


def test_overspend_rejected(engine, octx):
    bad = corpus.overspend_proofs(octx, 8)
    ref, nul, st = engine.batch_verify_spend_and_refund(bad["proofs"], bad["rnd"])
    assert (st == 7).all()
    assert not ref.any() and not nul.any()


def test_wrong_issuer_key(act, octx, base):
    x2, w2 = O.keygen(corpus.xof(b"other-key", 64))
    params = act.Params(octx.h)
    with act.Engine(params, act.PrivateKey(x2, w2)) as eng2:
        proofs = base["proofs"][:8 * corpus.PROOF_BYTES]; rnd = base["rnd"][:8 * 128]
        _, _, st = eng2.batch_verify_spend_and_refund(proofs, rnd)
        assert (st == 7).all()      # src/tests.rs:2019-2024


def test_refund_check_every_tamper_class(engine, octx, base):
    """Row a4 (PreRefund::to_credit_token, src/lib.rs:1217-1253): e + 1 (src/tests.rs:802-816), A + G, gamma + 1, z + 1
    (:1176-1226), malformed / identity A, a refund for another spend, swapped or malformed commitments, non-canonical scalar
    encodings (must accept) -- the engine's status equals the oracle's and the class's expected error."""
    ref, nul, st = engine.batch_verify_spend_and_refund(base["proofs"], base["rnd"])
    assert (st == 0).all()
    com = base["proofs"].reshape(-1, corpus.PROOF_BYTES)[:, 128:128 + 4096].copy().reshape(-1)
    c2, r2, expect, labels = corpus.mutate_refunds(com, ref)
    got = engine.batch_refund_check(c2, r2)
    o_st, _ = octx.batch_refund_check(c2, r2, threads=8)
    assert got.tolist() == o_st.tolist() == expect.tolist(), [(l, int(g), int(o)) for l, g, o in zip(labels, got, o_st) if g != o]
    assert set(got.tolist()) == {0, 4, 0x81} and len(set(labels)) >= 10


def test_issuance_check_every_tamper_class(engine, octx, base):
    """Row a3 (PreIssuance::to_credit_token, src/lib.rs:528-562): every response component tampered (src/tests.rs:703-714,
    836-847), a K that is not the request's, malformed points, non-canonical scalars (accept)."""
    K, rs, expect, labels = corpus.mutate_responses(base)
    got = engine.batch_issuance_check(K, rs)
    o_st, _ = octx.batch_issuance_check(K, rs, threads=8)
    assert got.tolist() == o_st.tolist() == expect.tolist(), [(l, int(g), int(o)) for l, g, o in zip(labels, got, o_st) if g != o]
    assert set(got.tolist()) == {0, 2, 0x81}


def test_tampered_tokens_are_rejected(engine, octx):
    """prop_token_tampering_detection (src/tests.rs:1898-1927): proofs made from tokens whose a / e were replaced are rejected
    with InvalidClientSpendProof (a = identity: IdentityPointError); outputs equal the oracle's bit for bit."""
    t = corpus.tampered_token_proofs(octx, 20)
    ref, nul, st = engine.batch_verify_spend_and_refund(t["proofs"], t["rnd"])
    o_ref, o_nul, o_st, _ = octx.batch_refund(t["proofs"], t["rnd"], threads=8)
    assert st.tolist() == o_st.tolist() == t["expect"].tolist()
    assert (ref == o_ref).all() and (nul == o_nul).all()
    # the device prover given the same tampered tokens produces proofs the engine rejects the same way
    one = np.frombuffer(b"".join((1).to_bytes(32, "little") for _ in range(20)), np.uint8)
    p2, _, s2 = engine.batch_prove_spend(t["tokens"], one, seed=corpus.xof(b"tamper-gpu", 32))
    assert (s2 == 0).all()
    _, _, st2 = engine.batch_verify_spend_and_refund(p2, t["rnd"])
    assert st2.tolist() == t["expect"].tolist()


def test_empty_and_single(engine, octx, base):
    e = np.zeros(0, np.uint8)
    resp, st = engine.batch_issue(e, e, e)
    assert resp.size == 0 and st.size == 0
    ref, nul, st = engine.batch_verify_spend_and_refund(e, e)
    assert st.size == 0
    ref, nul, st = engine.batch_verify_spend_and_refund(base["proofs"][:corpus.PROOF_BYTES], base["rnd"][:128])
    o = octx.refund(base["proofs"][:corpus.PROOF_BYTES].tobytes(), base["rnd"][:128].tobytes())
    assert st[0] == o[0] == 0 and ref.tobytes() == o[1] and nul.tobytes() == o[2]


def test_golden_trip_and_corpus_on_gpu(act):
    """BASELINE config #1 (examples/act.rs trip) and the committed golden corpus through the C ABI:
    no oracle call here -- the expected bytes are the committed fixtures."""
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "trip.json")))
    import blake3
    stream = blake3.blake3(b"act-oracle-0").digest(length=34112)
    params = act.Params.new(*g["params"])
    assert params.h.hex() == g["h"]
    key = act.PrivateKey.from_secret(bytes.fromhex(g["x"]))
    assert key.w.hex() == g["w"]
    with act.Engine(params, key) as eng:
        c40 = (40).to_bytes(32, "little")
        resp, st = eng.batch_issue(bytes.fromhex(g["request"]), c40, stream[320:448])
        assert st[0] == 0 and resp.tobytes().hex() == g["response"]
        assert eng.batch_issuance_check(bytes.fromhex(g["request"])[:32], resp)[0] == 0
        proof = bytes.fromhex(g["proof"])
        ref, nul, st = eng.batch_verify_spend_and_refund(proof, stream[448 + 524 * 64:448 + 524 * 64 + 128])
        assert st[0] == 0 and ref.tobytes().hex() == g["refund"] and nul.tobytes().hex() == g["nullifier"]
        assert eng.batch_refund_check(proof[128:128 + 4096], ref)[0] == 0
        assert act.encode_refund_cbor(ref).hex() == g["cbor_refund"]
        assert act.encode_issuance_response_cbor(resp).hex() == g["cbor_response"]
    c = np.load(os.path.join(here, "golden", "corpus_small.npz"))
    with act.Engine(act.Params(c["h"].tobytes()), act.PrivateKey(c["x"].tobytes(), c["w"].tobytes())) as eng:
        resp, st = eng.batch_issue(c["req"], c["cs"], c["rnd_issue"])
        assert (st == c["status_issue"]).all() and (resp == c["resp"]).all()
        ref, nul, st = eng.batch_verify_spend_and_refund(c["proofs"], c["rnd"])
        assert (st == c["status"]).all() and (ref == c["refunds"]).all() and (nul == c["nullifiers"]).all()
        # round-2 fixture: tampered responses / refunds for the client-side checks and proofs made from tampered tokens (same key)
        k = np.load(os.path.join(here, "golden", "corpus_checks.npz"))
        assert k["h"].tobytes() == c["h"].tobytes() and k["x"].tobytes() == c["x"].tobytes()
        assert (eng.batch_issuance_check(k["K"], k["responses"]) == k["status_issuance_check"]).all()
        assert (eng.batch_refund_check(k["com"], k["refunds"]) == k["status_refund_check"]).all()
        ref, nul, st = eng.batch_verify_spend_and_refund(k["token_proofs"], k["token_rnd"])
        assert (st == k["token_status"]).all() and (ref == k["token_refunds"]).all() and (nul == k["token_nullifiers"]).all()


def test_ragged_sizes_around_the_resident_grid(engine, octx, base):
    """The range kernel is persistent (SMs x resident blocks) and its warps draw quarter-proof units from a per-launch
    counter: sizes below, at and just above the resident grid (592 blocks on B200), and sizes that are not a multiple of
    anything, must give the tiled per-proof answers of the oracle bit for bit, call after call on the same engine."""
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    u = len(expect)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    try:
        for chunk, sizes in ((65536, (2, 3, 147, 591, 592, 593, 2369)), (4096, (4095, 4097, 16385)), (1000, (2999, 3000, 3001))):
            engine.set_spend_chunk(chunk)      # sizes just below / at / above the pipeline chunk: one, two and three chunks, two streams
            for n in sizes:
                idx = (np.arange(n) * 5 + n) % u
                P = proofs.reshape(u, -1)[idx].reshape(-1).copy(); R = rnd.reshape(u, -1)[idx].reshape(-1).copy()
                ref, nul, st = engine.batch_verify_spend_and_refund(P, R)
                assert (st == o_st[idx]).all(), n
                assert (ref.reshape(n, -1) == o_ref.reshape(u, -1)[idx]).all() and (nul.reshape(n, -1) == o_nul.reshape(u, -1)[idx]).all(), n
    finally:
        engine.set_spend_chunk(65536)


def test_large_mixed_adversarial_batch(engine, octx, base):
    """BASELINE config #5 shape at a size the GPU finishes in a second: 40 000 proofs, three quarters of them tampered
    (every mutation class of SURVEY section 4), crossing the pipeline chunks (65 536 proofs by default; this test sets 16 384) and both streams with
    a ragged tail.  Size-independent property: a tiled batch must give the tiled per-proof answers of the oracle, bit for bit;
    then the batch replay screen (the caller's nullifier check) flags every repeated accepted nullifier."""
    import importlib
    import torch
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    u = len(expect)
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    n = 40000
    idx = (np.arange(n) * 7 + 3) % u                      # a permuted tiling, so neighbours differ
    P = proofs.reshape(u, -1)[idx].reshape(-1).copy(); R = rnd.reshape(u, -1)[idx].reshape(-1).copy()
    engine.set_spend_chunk(16384)                         # ramp chunks of 2048 and 8192, then 16384 and a ragged 13376
    try:
        ref, nul, st = engine.batch_verify_spend_and_refund(P, R)
    finally:
        engine.set_spend_chunk(65536)
    assert (st == o_st[idx]).all()
    assert (ref.reshape(n, -1) == o_ref.reshape(u, -1)[idx]).all() and (nul.reshape(n, -1) == o_nul.reshape(u, -1)[idx]).all()
    assert 0.3 < (st != 0).mean() < 0.9 and set(np.unique(st)) >= {0, 6, 7, 0x81}
    # rejected proofs leave zero-filled outputs
    assert not ref.reshape(n, -1)[st != 0].any() and not nul.reshape(n, -1)[st != 0].any()
    sh = importlib.import_module("anonymous-credit-tokens_b200.sharding")
    flagged = sh.flag_replays(torch.from_numpy(st).cuda(), torch.from_numpy(nul).cuda(), engine=engine).cpu().numpy()
    seen = set(); exp = st.copy()
    for i in range(n):
        if st[i] == 0:
            k = nul[32 * i:32 * i + 32].tobytes()
            if k in seen:
                exp[i] = 3
            seen.add(k)
    assert (flagged == exp).all() and (flagged == 3).sum() > n // 8


def test_issue_large_batch_and_device_buffers(engine, octx, base):
    """batch_issue over 300 000 tiled requests (crosses the 262 144-request chunk) and the device-buffer entry points on
    a non-default stream: same bytes as the host-buffer calls and as the oracle."""
    import torch
    req, cs, rnd, expect, labels = corpus.mutate_requests(octx, base)
    u = len(expect)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    n = 300000
    idx = (np.arange(n) * 5 + 1) % u
    Rq = req.reshape(u, -1)[idx].reshape(-1).copy(); Cs = cs.reshape(u, -1)[idx].reshape(-1).copy(); Rn = rnd.reshape(u, -1)[idx].reshape(-1).copy()
    resp, st = engine.batch_issue(Rq, Cs, Rn)
    assert (st == o_st[idx]).all() and (resp.reshape(n, -1) == o_resp.reshape(u, -1)[idx]).all()
    # device buffers, caller's stream
    s = torch.cuda.Stream()
    m = 5000
    with torch.cuda.stream(s):
        d = [torch.from_numpy(a[:m * k]).cuda() for a, k in ((Rq, 128), (Cs, 32), (Rn, 128))]
        d_resp = torch.zeros(m * 160, dtype=torch.uint8, device="cuda"); d_st = torch.full((m,), 99, dtype=torch.uint8, device="cuda")
        engine.batch_issue_dev(m, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d_resp.data_ptr(), d_st.data_ptr(), s.cuda_stream)
        proofs, prnd, _, _ = corpus.mutate_proofs(octx, base)
        k = len(proofs) // corpus.PROOF_BYTES
        d_p = torch.from_numpy(proofs).cuda(); d_r = torch.from_numpy(prnd).cuda()
        d_ref = torch.zeros(k * 128, dtype=torch.uint8, device="cuda"); d_nul = torch.zeros(k * 32, dtype=torch.uint8, device="cuda")
        d_pst = torch.full((k,), 99, dtype=torch.uint8, device="cuda")
        engine.batch_verify_spend_and_refund_dev(k, d_p.data_ptr(), d_r.data_ptr(), d_ref.data_ptr(), d_nul.data_ptr(), d_pst.data_ptr(), s.cuda_stream)
    s.synchronize()
    assert (d_st.cpu().numpy() == st[:m]).all() and (d_resp.cpu().numpy() == resp[:m * 160]).all()
    h_ref, h_nul, h_st = engine.batch_verify_spend_and_refund(proofs, prnd)
    assert (d_pst.cpu().numpy() == h_st).all() and (d_ref.cpu().numpy() == h_ref).all() and (d_nul.cpu().numpy() == h_nul).all()


def test_two_gpu_sharding_matches_one_gpu(act, octx, base):
    """Shards on two devices (when the box has them) concatenate to the single-device answer."""
    if act.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import importlib
    sh = importlib.import_module("anonymous-credit-tokens_b200.sharding")
    proofs, rnd, _, _ = corpus.mutate_proofs(octx, base)
    n = len(proofs) // corpus.PROOF_BYTES
    params, key = act.Params(octx.h), act.PrivateKey(octx.x, octx.w)
    with act.Engine(params, key, device=0) as e0, act.Engine(params, key, device=1) as e1:
        full = e0.batch_verify_spend_and_refund(proofs, rnd)
        parts = []
        for r, e in enumerate((e0, e1)):
            lo, hi = sh.shard_bounds(n, r, 2)
            parts.append(e.batch_verify_spend_and_refund(proofs[lo * corpus.PROOF_BYTES:hi * corpus.PROOF_BYTES], rnd[lo * 128:hi * 128]))
        for k in range(3):
            assert (np.concatenate([p[k] for p in parts]) == full[k]).all()


def test_replay_screen_kernel(engine):
    """act_flag_replays: NullifierDb semantics (src/tests.rs:28-50) in slice order, against a python dict."""
    import importlib
    import torch
    rs = np.random.RandomState(3)
    for n, k in ((1, 0), (1000, 0), (70000, 500)):
        nul = rs.randint(0, 256, size=(n, 32)).astype(np.uint8)
        if n > 10:
            nul[rs.randint(0, n, n // 3)] = nul[rs.randint(0, n, n // 3)]       # plant duplicates
            nul[rs.randint(0, n, n // 10)] = 0                                  # one heavily repeated key
        st = (rs.rand(n) < 0.25).astype(np.uint8) * 7
        seen = nul[rs.randint(0, n, k)].copy() if k else np.zeros((0, 32), np.uint8)
        got = engine.flag_replays(st, nul.reshape(-1), seen.reshape(-1))
        db = {bytes(s) for s in seen}; exp = st.copy()
        for i in range(n):
            if st[i] == 0:
                key = bytes(nul[i])
                if key in db:
                    exp[i] = 3
                db.add(key)
        assert (got == exp).all()
        # device-pointer form through sharding.flag_replays agrees with the torch formulation
        sh = importlib.import_module("anonymous-credit-tokens_b200.sharding")
        t_st, t_nul = torch.from_numpy(st).cuda(), torch.from_numpy(nul.reshape(-1)).cuda()
        t_seen = torch.from_numpy(seen.reshape(-1)).cuda() if k else None
        a = sh.flag_replays(t_st, t_nul, t_seen, engine=engine).cpu().numpy()
        import replay_reference
        b = replay_reference.flag_replays(t_st, t_nul, t_seen).cpu().numpy()
        assert (a == exp).all() and (b == exp).all()


def test_cbor_fast_path_on_device(act, engine, octx, base):
    """Canonical-CBOR unpack/encode kernels against the host codec; non-canonical items are handed back (0xFF)."""
    proofs = base["proofs"].reshape(-1, corpus.PROOF_BYTES)[:40]
    host = b"".join(act.encode_spend_proof_cbor(p) for p in proofs)
    dev = engine.encode_cbor(act.KIND_PROOF, proofs.reshape(-1))
    assert dev.tobytes() == host
    items = np.frombuffer(host, np.uint8).reshape(40, -1).copy()
    items[3, 0] = 0xbf            # indefinite-length map: valid CBOR the host parser accepts, not the canonical skeleton
    items[7, 40] ^= 0xff          # payload byte change: still canonical
    rec, st = engine.unpack_cbor(act.KIND_PROOF, items.reshape(-1))
    assert st[3] == act.NOT_CANONICAL and (np.delete(st, 3) == 0).all()
    want = proofs.copy(); want[7] = np.frombuffer(act.pack_spend_proofs_cbor([items[7].tobytes()])[0], np.uint8)
    ok = np.ones(40, bool); ok[3] = False
    assert (rec.reshape(40, -1)[ok] == want[ok]).all()
    for kind, recs in ((act.KIND_REQUEST, base["req"]), (act.KIND_RESPONSE, base["resp"])):
        enc = engine.encode_cbor(kind, recs)
        back, st = engine.unpack_cbor(kind, enc)
        assert (st == 0).all() and (back == recs).all()
    r0 = base["resp"][:160]
    assert engine.encode_cbor(act.KIND_RESPONSE, r0).tobytes() == act.encode_issuance_response_cbor(r0)


def test_client_generators_on_gpu(act, engine, octx):
    """request / prove_spend kernels: bit-exact with the oracle prover on identical RNG bytes; then a full-device trip at a
    size no CPU prover reaches in test time: 20 000 unique tokens are requested, issued, spent and refunded on the GPU and
    every stage accepts (issue -> issuance_check -> prove_spend -> verify+refund -> refund_check), nullifiers all distinct."""
    import torch
    n = 12
    base = corpus.gen_valid(octx, n, seed=b"prover-gpu", threads=4)
    st = corpus.trip_streams(b"prover-gpu", n)
    assert (engine.batch_request(st["pre"], st["req_rnd"]) == base["req"]).all()
    tokens, charges = corpus.tokens_from(base, st["pre"]), corpus.charges_from(base)
    proofs, prer, status = engine.batch_prove_spend(tokens, charges, rnd=st["prove_rnd"])
    assert (status == 0).all() and (proofs == base["proofs"]).all() and (prer == base["prerefund"]).all()
    seed = corpus.xof(b"derived-seed", 32)
    rnd = np.frombuffer(b"".join(corpus.prover_stream(seed, 1000 + i) for i in range(n)), np.uint8)
    p1, r1, s1 = engine.batch_prove_spend(tokens, charges, seed=seed, first_index=1000)
    p2, r2, s2 = engine.batch_prove_spend(tokens, charges, rnd=rnd)
    assert (p1 == p2).all() and (r1 == r2).all() and (s1 == 0).all()
    # full-device trip: torch ops and engine launches share one (non-default) stream
    S = torch.cuda.Stream()
    with torch.cuda.stream(S):
        N = 20000
        g = torch.Generator(device="cuda"); g.manual_seed(7)
        rb = lambda k: torch.randint(0, 256, (k,), dtype=torch.uint8, device="cuda", generator=g)
        pre = rb(N * 64); pre.view(N, 64)[:, 31] &= 0x0f; pre.view(N, 64)[:, 63] &= 0x0f            # r, k < 2^252 (already reduced)
        req = torch.zeros(N * 128, dtype=torch.uint8, device="cuda")
        engine.batch_request_dev(N, pre.data_ptr(), rb(N * 128).data_ptr(), req.data_ptr(), S.cuda_stream)
        credits = torch.randint(20, 1000, (N,), device="cuda", generator=g)
        spend = (torch.rand(N, device="cuda", generator=g) * (credits - 1)).long() + 1                # 1 <= s <= c - 1
        le32 = lambda v: torch.cat([(v.view(N, 1) >> (8 * torch.arange(2, device="cuda"))).to(torch.uint8) & 0xff, torch.zeros(N, 30, dtype=torch.uint8, device="cuda")], 1).reshape(-1)
        cs, ch = le32(credits), le32(spend)
        resp = torch.zeros(N * 160, dtype=torch.uint8, device="cuda"); ist = torch.full((N,), 9, dtype=torch.uint8, device="cuda")
        engine.batch_issue_dev(N, req.data_ptr(), cs.data_ptr(), rb(N * 128).data_ptr(), resp.data_ptr(), ist.data_ptr(), S.cuda_stream)
        S.synchronize()
        assert (ist == 0).all()
        K = req.view(N, 128)[:, :32].contiguous()
        cst = torch.full((N,), 9, dtype=torch.uint8, device="cuda")
        engine.batch_issuance_check_dev(N, K.data_ptr(), resp.data_ptr(), cst.data_ptr(), S.cuda_stream)
        tok = torch.cat([resp.view(N, 160)[:, :64], pre.view(N, 64)[:, 32:], pre.view(N, 64)[:, :32], cs.view(N, 32)], 1).contiguous()
        pf = torch.zeros(N * corpus.PROOF_BYTES, dtype=torch.uint8, device="cuda"); prf = torch.zeros(N * 96, dtype=torch.uint8, device="cuda")
        pst = torch.full((N,), 9, dtype=torch.uint8, device="cuda")
        engine.batch_prove_spend_dev(N, tok.data_ptr(), ch.data_ptr(), None, seed, 0, pf.data_ptr(), prf.data_ptr(), pst.data_ptr(), S.cuda_stream)
        ref = torch.zeros(N * 128, dtype=torch.uint8, device="cuda"); nul = torch.zeros(N * 32, dtype=torch.uint8, device="cuda")
        vst = torch.full((N,), 9, dtype=torch.uint8, device="cuda")
        engine.batch_verify_spend_and_refund_dev(N, pf.data_ptr(), rb(N * 128).data_ptr(), ref.data_ptr(), nul.data_ptr(), vst.data_ptr(), S.cuda_stream)
        com = pf.view(N, corpus.PROOF_BYTES)[:, 128:128 + 4096].contiguous()
        rst = torch.full((N,), 9, dtype=torch.uint8, device="cuda")
        engine.batch_refund_check_dev(N, com.data_ptr(), ref.data_ptr(), rst.data_ptr(), S.cuda_stream)
        S.synchronize()
        assert (cst == 0).all() and (pst == 0).all() and (vst == 0).all() and (rst == 0).all()
        assert (nul.view(N, 32) == pre.view(N, 64)[:, 32:]).all()                                    # nullifier() = the token's k
        m = prf.view(N, 96)[:, 64:72].contiguous().view(torch.int64).view(N)
        assert (m == credits - spend).all()                                                          # PreRefund.m = c - s
    # a spot check of the device-generated proofs against the oracle verifier
    o_ref, o_nul, o_st, _ = octx.batch_refund(pf[:4 * corpus.PROOF_BYTES].cpu().numpy(), np.zeros(4 * 128, np.uint8), threads=4)
    assert (o_st == 0).all() and (o_nul == nul[:128].cpu().numpy()).all()
    # overspend (s > c) is produced but rejected, as in the reference (src/tests.rs:366-374)
    with torch.cuda.stream(S):
        ch_bad = le32(credits + 1)
        rb64 = rb(64 * 128)
    engine.batch_prove_spend_dev(64, tok.data_ptr(), ch_bad.data_ptr(), None, seed, 0, pf.data_ptr(), prf.data_ptr(), pst.data_ptr(), S.cuda_stream)
    engine.batch_verify_spend_and_refund_dev(64, pf.data_ptr(), rb64.data_ptr(), ref.data_ptr(), nul.data_ptr(), vst.data_ptr(), S.cuda_stream)
    S.synchronize()
    assert (vst[:64] == 7).all()


def test_same_token_spent_twice_with_different_proofs(engine, octx):
    """BASELINE config #5's "same-k different-proof replay": two honest proofs from ONE token (different charge and
    randomness) both verify -- refund() cannot know (src/lib.rs:741-745) -- and carry the same nullifier; the batch replay
    screen flags the later one.  Outputs equal the oracle's."""
    n = 6
    base = corpus.gen_valid(octx, n, seed=b"double-spend", threads=4)
    st = corpus.trip_streams(b"double-spend", n)
    tokens = corpus.tokens_from(base, st["pre"])
    ones = np.frombuffer(b"".join((1).to_bytes(32, "little") for _ in range(n)), np.uint8)
    p2, _, s2 = engine.batch_prove_spend(tokens, ones, seed=corpus.xof(b"second-proof", 32))
    assert (s2 == 0).all() and (p2 != base["proofs"]).any()
    both = np.concatenate([base["proofs"], p2]); rnd = np.frombuffer(corpus.xof(b"double-spend/rnd", 128 * 2 * n), np.uint8)
    ref, nul, vst = engine.batch_verify_spend_and_refund(both, rnd)
    o_ref, o_nul, o_st, _ = octx.batch_refund(both, rnd, threads=4)
    assert (vst == 0).all() and (o_st == 0).all() and (ref == o_ref).all() and (nul == o_nul).all()
    assert (nul[:32 * n] == nul[32 * n:]).all()
    assert engine.flag_replays(vst, nul).tolist() == [0] * n + [3] * n


def test_token_lifecycles_match_the_oracle(engine, octx):
    """Multi-generation token chains (corpus.LIFECYCLES: the reference's scenario tests, src/tests.rs:210-426,642-689,
    876-1059) through the C ABI: request, issue, issuance_check, prove_spend, verify+refund, refund_check generation
    after generation; every byte equals the oracle's."""
    assert corpus.run_lifecycles(corpus.EngineImpl(engine)) == corpus.run_lifecycles(corpus.OracleImpl(octx))


def test_sequential_rng_contract(engine, octx, base):
    """act_batch_*_seq == a loop of reference calls over ONE shared RNG: randomness is consumed only by accepted
    requests, in slice order (src/lib.rs:638-643, 842-846)."""
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    n = len(expect)
    stream = np.frombuffer(corpus.xof(b"one-shared-rng", 128 * n), np.uint8)
    ref, nul, st, used = engine.batch_verify_spend_and_refund_seq(proofs, stream)
    pos = 0
    for i in range(n):
        o_st, o_ref, o_nul = octx.refund(proofs[i * corpus.PROOF_BYTES:(i + 1) * corpus.PROOF_BYTES].tobytes(), stream[pos:pos + 128].tobytes())
        assert st[i] == o_st, (i, labels[i])
        if o_st == 0:
            pos += 128
            assert ref[128 * i:128 * i + 128].tobytes() == o_ref and nul[32 * i:32 * i + 32].tobytes() == o_nul, (i, labels[i])
        else:
            assert not ref[128 * i:128 * i + 128].any() and not nul[32 * i:32 * i + 32].any()
    assert used == pos and 0 < pos < 128 * n
    with pytest.raises(Exception):
        engine.batch_verify_spend_and_refund_seq(proofs, stream[:pos - 1])       # stream too short
    req, cs, irnd, iexp, ilab = corpus.mutate_requests(octx, base)
    resp, ist, iused = engine.batch_issue_seq(req, cs, stream)
    pos = 0
    for i in range(len(iexp)):
        o_st, o_resp = octx.issue(req[128 * i:128 * i + 128].tobytes(), cs[32 * i:32 * i + 32].tobytes(), stream[pos:pos + 128].tobytes())
        assert ist[i] == o_st, (i, ilab[i])
        if o_st == 0:
            pos += 128
            assert resp[160 * i:160 * i + 160].tobytes() == o_resp
        else:
            assert not resp[160 * i:160 * i + 160].any()
    assert iused == pos
    # a batch that crosses the pipeline chunk (two streams) still lines up with the stream positions
    engine.set_spend_chunk(8192)
    u = n; N = 20000
    idx = (np.arange(N) * 11 + 5) % u
    P = proofs.reshape(u, -1)[idx].reshape(-1).copy()
    big = np.frombuffer(corpus.xof(b"one-shared-rng-big", 128 * N), np.uint8)
    r2, n2, s2, used2 = engine.batch_verify_spend_and_refund_seq(P, big)
    acc = np.flatnonzero(s2 == 0)
    assert used2 == 128 * len(acc) and (s2 == st[idx]).all()
    packed = big[:used2].reshape(-1, 128)
    full_rnd = np.zeros((N, 128), np.uint8); full_rnd[acc] = packed
    r3, n3, s3 = engine.batch_verify_spend_and_refund(P, full_rnd.reshape(-1))
    engine.set_spend_chunk(65536)
    assert (r3 == r2).all() and (n3 == n2).all() and (s3 == s2).all()


def test_engine_rejects_inconsistent_key_and_bad_params(act, octx):
    """act_engine_create fails loudly: W that is not G*x, undecodable H or W."""
    x2, w2 = O.keygen(corpus.xof(b"other-key", 64))
    with pytest.raises(act.ActError):
        act.Engine(act.Params(octx.h), act.PrivateKey(octx.x, w2))
    bad = bytearray(octx.h); bad[0:32] = corpus.bad_point_encodings()[0]
    with pytest.raises(act.ActError):
        act.Engine(act.Params(bytes(bad)), act.PrivateKey(octx.x, octx.w))
    with pytest.raises(act.ActError):
        act.Engine(act.Params(octx.h), act.PrivateKey(octx.x, corpus.bad_point_encodings()[4]))
    # a non-canonical secret (x + l) is the same key
    xl = (corpus.sc_int(octx.x) + corpus.ELL).to_bytes(32, "little")
    with act.Engine(act.Params(octx.h), act.PrivateKey(xl, octx.w)) as e:
        assert e.launch_count > 0


def test_two_pass_forms_equal_the_one_pass_calls(engine, octx, base):
    """act_batch_issue_verify/_sign and act_batch_spend_verify / act_batch_refund_sign: the verify pass gives the statuses (and
    nullifiers) of the one-pass call, and signing with 128 bytes per ACCEPTED request in slice order gives the outputs a loop
    of reference calls over one RNG gives (src/lib.rs:638-643, 842-846) -- checked against the oracle called one by one."""
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    n = len(expect)
    nul, st, kp = engine.batch_spend_verify(proofs)
    ref1, nul1, st1 = engine.batch_verify_spend_and_refund(proofs, rnd)
    assert (st == st1).all() and (nul == nul1).all()
    acc = int((st == 0).sum())
    stream = np.frombuffer(corpus.xof(b"two-pass", 128 * acc), np.uint8)
    ref = engine.batch_refund_sign(kp, st, stream)
    pos = 0
    for i in range(n):
        if st[i] == 0:
            o_st, o_ref, o_nul = octx.refund(proofs[i * corpus.PROOF_BYTES:(i + 1) * corpus.PROOF_BYTES].tobytes(), stream[pos:pos + 128].tobytes())
            pos += 128
            assert o_st == 0 and ref[128 * i:128 * i + 128].tobytes() == o_ref and nul[32 * i:32 * i + 32].tobytes() == o_nul, (i, labels[i])
        else:
            assert not ref[128 * i:128 * i + 128].any() and not nul[32 * i:32 * i + 32].any()
    with pytest.raises(Exception):
        engine.batch_refund_sign(kp, st, stream[:-1])            # one byte short
    req, cs, irnd, iexp, ilab = corpus.mutate_requests(octx, base)
    ist = engine.batch_issue_verify(req)
    resp1, ist1 = engine.batch_issue(req, cs, irnd)
    assert (ist == ist1).all()
    acc = int((ist == 0).sum())
    stream = np.frombuffer(corpus.xof(b"two-pass-issue", 128 * acc), np.uint8)
    resp = engine.batch_issue_sign(req, cs, ist, stream)
    pos = 0
    for i in range(len(iexp)):
        if ist[i] == 0:
            o_st, o_resp = octx.issue(req[128 * i:128 * i + 128].tobytes(), cs[32 * i:32 * i + 32].tobytes(), stream[pos:pos + 128].tobytes())
            pos += 128
            assert o_st == 0 and resp[160 * i:160 * i + 160].tobytes() == o_resp, (i, ilab[i])
        else:
            assert not resp[160 * i:160 * i + 160].any()


def _multi_devices(act):
    """Two replicas: on two GPUs when the box has them, else both on GPU 0 (same sharded code path, same-device "peer" copies)."""
    return [0, 1] if act.device_count() >= 2 else [0, 0]


def test_multi_device_engine_matches_single_device(act, engine, octx, base):
    """act_engine_create_multi (SURVEY 8b `devices[], n_devices`; 8e: contiguous shards, no cross-GPU arithmetic): every
    host-buffer call on the multi-device engine returns the single-device bytes -- ragged shard sizes included."""
    params, key = act.Params(octx.h), act.PrivateKey(octx.x, octx.w)
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    u = len(expect)
    req, cs, irnd, iexp, _ = corpus.mutate_requests(octx, base)
    with act.Engine(params, key, devices=_multi_devices(act)) as m:
        assert m.replica_count == 2
        for n in (1, 2, 95, u):
            P, R = proofs[:n * corpus.PROOF_BYTES], rnd[:n * 128]
            a, b = m.batch_verify_spend_and_refund(P, R), engine.batch_verify_spend_and_refund(P, R)
            assert all((x == y).all() for x, y in zip(a, b)), n
        a, b = m.batch_issue(req, cs, irnd), engine.batch_issue(req, cs, irnd)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        K = base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1)
        assert (m.batch_issuance_check(K, base["resp"]) == engine.batch_issuance_check(K, base["resp"])).all()
        com = proofs.reshape(u, -1)[:, 128:128 + 4096].copy().reshape(-1)
        ref = engine.batch_verify_spend_and_refund(proofs, rnd)[0]
        assert (m.batch_refund_check(com, ref) == engine.batch_refund_check(com, ref)).all()
        # the sequential-RNG contract across shards: replica g starts at the stream position the accepted requests of the
        # shards before it consumed
        stream = np.frombuffer(corpus.xof(b"multi-seq", 128 * u), np.uint8)
        a, b = m.batch_verify_spend_and_refund_seq(proofs, stream), engine.batch_verify_spend_and_refund_seq(proofs, stream)
        assert all((np.asarray(x) == np.asarray(y)).all() for x, y in zip(a, b))
        a, b = m.batch_issue_seq(req, cs, stream), engine.batch_issue_seq(req, cs, stream)
        assert all((np.asarray(x) == np.asarray(y)).all() for x, y in zip(a, b))
        # a bigger tiled batch crossing the chunk size on every replica
        m.set_spend_chunk(2048)
        n = 9001
        idx = (np.arange(n) * 7 + 1) % u
        P = proofs.reshape(u, -1)[idx].reshape(-1).copy(); R = rnd.reshape(u, -1)[idx].reshape(-1).copy()
        o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
        r2, n2, s2 = m.batch_verify_spend_and_refund(P, R)
        assert (s2 == o_st[idx]).all() and (r2.reshape(n, -1) == o_ref.reshape(u, -1)[idx]).all() and (n2.reshape(n, -1) == o_nul.reshape(u, -1)[idx]).all()
        with pytest.raises(act.ActError):      # device-buffer calls take one replica, not the multi-device handle
            m.batch_issue_dev(1, 0, 0, 0, 0, 0)


def test_screened_batch_call(act, engine, octx, base):
    """act_batch_verify_spend_and_refund_screened = verify + refund + the caller's nullifier check (examples/act.rs:60-77) over a
    slice: first valid occurrence wins, later ones and members of `seen` get DoubleSpendError and NO refund; on the
    multi-device engine the status + nullifier gather crosses devices.  Expected result built from the oracle + a python dict."""
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    u = len(expect)
    n = 5000
    idx = (np.arange(n) * 13 + 2) % u                       # heavy repetition: most accepted proofs are replays
    P = proofs.reshape(u, -1)[idx].reshape(-1).copy(); R = rnd.reshape(u, -1)[idx].reshape(-1).copy()
    o_ref, o_nul, o_st, _ = octx.batch_refund(proofs, rnd, threads=8)
    e_ref = o_ref.reshape(u, -1)[idx].copy(); e_nul = o_nul.reshape(u, -1)[idx].copy(); e_st = o_st[idx].copy()
    seen = o_nul.reshape(u, -1)[[i for i in range(u) if o_st[i] == 0][:3]].copy()     # three tokens spent before this batch
    db = {bytes(s) for s in seen}
    for i in range(n):
        if e_st[i] == 0:
            k = bytes(e_nul[i])
            if k in db:
                e_st[i] = 3; e_ref[i] = 0
            db.add(k)
    assert (e_st == 3).sum() > n // 5 and (e_st == 0).sum() > 10
    params, key = act.Params(octx.h), act.PrivateKey(octx.x, octx.w)
    with act.Engine(params, key, devices=_multi_devices(act)) as m:
        m.set_spend_chunk(1024)
        for eng in (engine, m):
            ref, nul, st = eng.batch_verify_spend_and_refund_screened(P, R, seen=seen.reshape(-1))
            assert (st == e_st).all()
            assert (ref.reshape(n, -1) == e_ref).all() and (nul.reshape(n, -1) == e_nul).all()


def test_differential_fuzz_against_the_oracle(engine, octx, base):
    """prop_invalid_proofs_rejected (src/tests.rs:1681-1713) turned into a differential fuzz: 600 records with 1-3 random bit flips
    anywhere in the 16 832 bytes (points, scalars, any field), 400 requests likewise, plus whole-field random replacements --
    status, refund / response and nullifier bytes equal the oracle's on every record, whatever the oracle says."""
    rs = np.random.RandomState(20261017)
    u = len(base["proofs"]) // corpus.PROOF_BYTES
    n = 600
    P = base["proofs"].reshape(u, -1)[rs.randint(0, u, n)].copy(); R = base["rnd"].reshape(u, -1)[rs.randint(0, u, n)].copy()
    for i in range(n):
        if i % 5 == 4:      # a whole 32-byte item replaced by random bytes (a random point encoding decodes about one time in four)
            item = rs.randint(0, 526)
            P[i, 32 * item:32 * item + 32] = rs.randint(0, 256, 32)
        else:
            for _ in range(1 + i % 3):
                b = rs.randint(0, corpus.PROOF_BYTES * 8)
                P[i, b // 8] ^= 1 << (b % 8)
    P = P.reshape(-1); R = R.reshape(-1)
    ref, nul, st = engine.batch_verify_spend_and_refund(P, R)
    o_ref, o_nul, o_st, _ = octx.batch_refund(P, R, threads=8)
    assert st.tolist() == o_st.tolist()
    assert (ref == o_ref).all() and (nul == o_nul).all()
    assert {7, 0x81} <= set(st.tolist())
    m = 400
    Q = base["req"].reshape(u, -1)[rs.randint(0, u, m)].copy(); C = base["cs"].reshape(u, -1)[rs.randint(0, u, m)].copy()
    IR = base["rnd"].reshape(u, -1)[rs.randint(0, u, m)].copy()
    for i in range(m):
        b = rs.randint(0, 128 * 8)
        Q[i, b // 8] ^= 1 << (b % 8)
        if i % 7 == 0:
            C[i] = rs.randint(0, 256, 32)          # any 32 bytes are a credit amount (reduced mod l)
    Q = Q.reshape(-1); C = C.reshape(-1); IR = IR.reshape(-1)
    resp, ist = engine.batch_issue(Q, C, IR)
    o_resp, o_ist, _ = octx.batch_issue(Q, C, IR, threads=8)
    assert ist.tolist() == o_ist.tolist() and (resp == o_resp).all()


def test_device_entry_points_reject_misaligned_and_null_buffers(act, engine):
    """The _dev calls move records with 16-byte vector accesses: a misaligned or null device pointer is refused up front
    (a whole-call error), never dereferenced."""
    import torch
    buf = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    p = buf.data_ptr()
    with pytest.raises(act.ActError, match="16-byte aligned"):
        engine.batch_issue_dev(1, p + 4, p + 1024, p + 2048, p + 4096, p + 8192)
    with pytest.raises(act.ActError, match="16-byte aligned"):
        engine.batch_verify_spend_and_refund_dev(1, p, p + 32768 + 8, p + 40000 - 40000 % 16, p + 50000 - 50000 % 16, p + 60000)
    with pytest.raises(act.ActError, match="null buffer"):
        engine.batch_issue_dev(1, p, None, p + 2048, p + 4096, p + 8192)
