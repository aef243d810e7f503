"""Pins the CPU oracle (oracle/act_oracle.c): RFC 9496 vectors, BLAKE3 via the independent python package,
libsodium ristretto255, python big ints, and the committed golden trip.  No GPU."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import corpus
import oracle_lib as O
import refstack as R

HERE = os.path.dirname(os.path.abspath(__file__))
ELL = corpus.ELL
P = corpus.P25519

RFC9496_MULTIPLES = [  # RFC 9496 Appendix A.1: encodings of 0*B .. 10*B
    "0000000000000000000000000000000000000000000000000000000000000000",
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
    "da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57",
    "e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e",
    "f64746d3c92b13050ed8d80236a7f0007c3b3f962f5ba793d19a601ebb1df403",
    "44f53520926ec81fbd5a387845beb7df85a96a24ece18738bdcfa6a7822a176d",
    "903293d8f2287ebe10e2374dc1a53e0bc887e592699f02d077d5263cdd55601c",
    "02622ace8f7303a31cafc63f8fc48fdc16e1c8c8d234b2f0d6685282a9076031",
    "20706fd788b2720a1ed2a5dad4952b01f413bcf0e7564de8cdc816689e2db95f",
]


def test_rfc9496_generator_multiples():
    for k, hx in enumerate(RFC9496_MULTIPLES):
        s = k.to_bytes(32, "little")
        assert O.scalarmult_base(s).hex() == hx
        if k:
            assert O.scalarmult(s, bytes.fromhex(RFC9496_MULTIPLES[1])).hex() == hx
        assert O.decode_encode(bytes.fromhex(hx)).hex() == hx


def test_rfc9496_invalid_encodings_rejected():
    for b in corpus.bad_point_encodings():
        assert O.decode_encode(b) is None
        if R.available():
            assert not R.is_valid(b)


def test_rfc9496_hash_to_group_vectors():
    """RFC 9496 Appendix A.3 (one-way map applied to SHA-512 of the label): three of its vectors.  This is the map behind
    RistrettoPoint::from_uniform_bytes, i.e. behind Params::new (src/lib.rs:320-353)."""
    vectors = [("Ristretto is traditionally a short shot of espresso coffee", "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"),
               ("This produces a concentrated shot of coffee per volume.", "ae81e7dedf20a497e10c304a765c1767a42d6e06029758d2d7e8ef7cc4c41179"),
               ("Just pulling a normal shot short will produce a weaker shot", "e2705652ff9f5e44d3e841bf1c251cf7dddb77d140870d1ab2ed64f1a9ce8628")]
    for label, hx in vectors:
        u = hashlib.sha512(label.encode()).digest()
        assert O.from_uniform(u).hex() == hx
        if R.available():
            assert R.from_hash(u).hex() == hx


def test_field_and_scalar_arithmetic_vs_python_ints():
    rnd = random.Random(3)
    for it in range(200):
        a, b, c = (rnd.getrandbits(255) for _ in range(3))
        ab = lambda v: v.to_bytes(32, "little")
        assert int.from_bytes(O.fe_mul(ab(a), ab(b)), "little") == a * b % P
        if a % P:
            assert int.from_bytes(O.fe_invert(ab(a)), "little") == pow(a, -1, P)
        x, y, z = (rnd.getrandbits(256) for _ in range(3))
        assert int.from_bytes(O.sc_reduce32(ab(x)), "little") == x % ELL
        w = rnd.getrandbits(512)
        assert int.from_bytes(O.sc_reduce64(w.to_bytes(64, "little")), "little") == w % ELL
        assert int.from_bytes(O.sc_muladd(ab(x), ab(y), ab(z)), "little") == (x * y + z) % ELL
        if x % ELL:
            assert int.from_bytes(O.sc_invert(ab(x)), "little") == pow(x % ELL, -1, ELL)
    assert O.sc_invert(bytes(32)) == bytes(32)       # Scalar::invert(0) = 0
    assert O.sc_reduce32(b"\xff" * 32) == ((2**256 - 1) % ELL).to_bytes(32, "little")


def test_blake3_vs_python_package():
    import blake3
    rnd = random.Random(5)
    for n in [0, 1, 3, 63, 64, 65, 127, 128, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 4097, 8192, 15784, 16384, 16385, 40000]:
        data = bytes(rnd.getrandbits(8) for _ in range(n))
        assert O.blake3(data, 32) == blake3.blake3(data).digest()
        assert O.blake3(data, 131) == blake3.blake3(data).digest(length=131)
    assert O.blake3(b"abc", 32).hex().startswith("6437b3ac38465133ffb63b75273a8db5")


@pytest.mark.skipif(not R.available(), reason="libsodium with ristretto255 not found")
def test_group_ops_vs_libsodium():
    rnd = random.Random(9)
    for it in range(40):
        s = rnd.getrandbits(256) % ELL
        u = bytes(rnd.getrandbits(8) for _ in range(64))
        Pt = R.from_hash(u)
        assert O.from_uniform(u) == Pt
        assert O.scalarmult(s.to_bytes(32, "little"), Pt) == R.mul(Pt, s)
        assert O.scalarmult_base(s.to_bytes(32, "little")) == R.mul_base(s)
        Q = R.mul_base(rnd.getrandbits(200))
        assert O.point_add(Pt, Q) == R.add(Pt, Q)
    # non-canonical scalar l+5 acts like 5
    assert O.scalarmult_base((ELL + 5).to_bytes(32, "little")).hex() == RFC9496_MULTIPLES[5]


def test_params_known_values():
    # computed by an independent stack (blake3-py + libsodium from_hash)
    exp = {
        ("example-org", "payment-api", "production", "2024-01-15"):
            "1e2015fd2f2d25c3fb25b0998a6daf6f6b85e0f8f578ff22ae54eeeadd47b15d8c95a42b7d7684cf3bc8e46be4d8301e2cf63408dd209bc9136d45b9d4baff653af77a3fd3e13c3142dbab7aca37f366e73b3fd607744c542a4e3626be89ac5e",
        corpus.TEST_PARAMS:
            "e84286ef61fec5910cd7909ec3c2a269ca4132415e5ae45636687926208c414a363f0f854026f8f82660a597ebcfea5149492a20f00d63433949a0075e142d3ef66eda327af6078d092390739de7df0aa3967fc80aad615decbf982bd305ff67",
        corpus.BENCH_PARAMS:
            "0e9f46fc532ca519709bb687c84a728e05dbf3535c76e5bc4553b08fc14ea12c58bd86e98af5dcbde663aedc02e5adbd20e75ebebf2da947dd0bad4b1afcb75e3c61933acaae64fa8062855ee095c48ea2c54a95b8ff86b8a205546560ff8d28",
    }
    for p, hx in exp.items():
        assert O.params_derive(*p).hex() == hx
        if R.available():
            assert b"".join(R.params_new(*p)).hex() == hx
    # determinism / separation (src/tests.rs:723-748)
    assert O.params_derive("a", "b", "c", "d") == O.params_derive("a", "b", "c", "d")
    assert O.params_derive("a", "b", "c", "d") != O.params_derive("a", "b", "c", "e")


def test_empty_spend_transcript_challenge():
    # SURVEY.md 8(c) provisional cross-check value
    h = O.params_derive("example-org", "payment-api", "production", "2024-01-15")
    x, w = O.keygen(bytes(64))
    ctx = O.Ctx(h, x, w)
    assert ctx.transcript_challenge("spend", b"").hex() == "47312d20b4c2103ce733c8fc4f0ec79ddb4380ac77a87eeac8bb6294f145ea05"
    assert ctx.transcript_challenge("spend", b"") != ctx.transcript_challenge("refund", b"")   # label separation


def test_golden_trip_config1():
    """BASELINE config #1: examples/act.rs issue 40 -> spend 20 -> refund; committed bytes from the independent stack."""
    g = json.load(open(os.path.join(HERE, "golden", "trip.json")))
    import blake3
    stream = blake3.blake3(b"act-oracle-0").digest(length=34112)
    # SURVEY Appendix C table
    assert g["x"] == "2fab32e93c4d6d8c368770ca3735644fac2518840056fe8588194a167f97c600"
    assert g["w"] == "2e4a7e5f5bdcc7cf7619717dde2f8200d61e95a73bdcd0f0d183f2bbd39b4a3c"
    assert g["sha256"]["request"] == "a3a15088d06cbc0d43bea764fcddde0f27830adff9e1a6ed17caebac150542bd"
    assert g["sha256"]["response"] == "01d75369132e12e51f5bb7be1f865290ccda5099ae590285cdb7db7a0a384027"
    assert g["sha256"]["proof"] == "af72874bfd910e4db4ceaef532af49b1b4211465663d6070b1e3121e13904259"
    assert g["sha256"]["refund"] == "5088c48f232723256a64cecb91b83265275628f8ef74f1bc85e6de190d664501"
    assert g["cbor_proof_sha256"] == "b22e03774d259d43152a79185d7248cdbf3285ca26d415f6bc7864f0135a7a23"
    assert hashlib.sha256(bytes.fromhex(g["cbor_request"])).hexdigest() == "8cca2536a45754010a1557749d9aebbbc9bb739fc0105a9e8c512c3a4b2e0a38"
    assert hashlib.sha256(bytes.fromhex(g["cbor_response"])).hexdigest() == "e98aa2bff59dc9c4f2a34c3369476d6167b95152a03d1b032e1507d5524f11d3"
    assert hashlib.sha256(bytes.fromhex(g["cbor_refund"])).hexdigest() == "418d5f0790afb1f929fdcf36b8b3116cbb37cfa73ae989726e5cff1516544e14"
    # the oracle reproduces every byte
    x, w = O.keygen(stream[:64])
    assert x.hex() == g["x"] and w.hex() == g["w"]
    h = O.params_derive(*g["params"])
    assert h.hex() == g["h"]
    ctx = O.Ctx(h, x, w)
    pre = O.sc_reduce64(stream[64:128]) + O.sc_reduce64(stream[128:192])
    req = ctx.request(pre, stream[192:320])
    assert req.hex() == g["request"]
    c40 = (40).to_bytes(32, "little")
    st, resp = ctx.issue(req, c40, stream[320:448])
    assert st == 0 and resp.hex() == g["response"]
    assert ctx.issuance_check(req[:32], resp) == 0
    token = resp[:64] + pre[32:] + pre[:32] + c40
    proof, prer = ctx.prove_spend(token, (20).to_bytes(32, "little"), stream[448:448 + 524 * 64])
    assert proof.hex() == g["proof"] and prer.hex() == g["prerefund"]
    st, refund, nul = ctx.refund(proof, stream[448 + 524 * 64:])
    assert st == 0 and refund.hex() == g["refund"] and nul.hex() == g["nullifier"]
    assert ctx.refund_check(proof[128:128 + 4096], refund) == 0
    # credits 40 then 20 (examples/act.rs:57,78): remaining balance m = 20
    assert int.from_bytes(prer[64:96], "little") == 20
    # mutations of Appendix C
    bad = bytearray(proof); bad[32:64] = (21).to_bytes(32, "little")
    assert ctx.refund(bytes(bad), bytes(128))[0] == 7
    bad = bytearray(proof); bad[64:96] = bytes(32)
    assert ctx.refund(bytes(bad), bytes(128))[0] == 6
    bad = bytearray(req); bad[64:96] = corpus.sc_bytes(corpus.sc_int(req[64:96]) + 1)
    assert ctx.issue(bytes(bad), c40, bytes(128))[0] == 1


def test_mutation_corpus_statuses(octx):
    """Every mutation class of the reference's tests gets the reference's error from the oracle."""
    base = corpus.gen_valid(octx, 24, seed=b"oracle-mut", threads=8)
    req, cs, rnd, expect, labels = corpus.mutate_requests(octx, base)
    resp, st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    # non-canonical scalar encodings give the same response as the canonical ones
    r0, s0, _ = octx.batch_issue(base["req"], base["cs"], base["rnd"], threads=8)
    for i, l in enumerate(labels):
        if l == "non-canonical scalars":
            assert (resp.reshape(-1, 160)[i] == r0.reshape(-1, 160)[i]).all()
    # rejected requests produce zero output and consume no randomness-dependent state
    assert not resp.reshape(-1, 160)[st != 0].any()
    proofs, rnd, expect, labels = corpus.mutate_proofs(octx, base)
    ref, nul, st, _ = octx.batch_refund(proofs, rnd, threads=8)
    for i, e in enumerate(expect):
        if e != 255:
            assert st[i] == e, (i, labels[i], st[i])
    assert not ref.reshape(-1, 128)[st != 0].any() and not nul.reshape(-1, 32)[st != 0].any()
    over = corpus.overspend_proofs(octx, 4)
    _, _, st, _ = octx.batch_refund(over["proofs"], over["rnd"], threads=4)
    assert (st == 7).all()


def _pool_map(fn, jobs):
    """The libsodium/big-int stack costs ~0.16 s per SpendProof: fan the corpus out over the host cores (fork; read-only state)."""
    import multiprocessing as mp
    workers = min(os.cpu_count() or 1, 8, max(1, len(jobs)))
    if workers == 1:
        return [fn(j) for j in jobs]
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(fn, jobs)


_XS = {}


def _x_refund(job):
    pf, rnd = job
    return R.refund_record(_XS["H"], _XS["x"], _XS["w"], pf, rnd)


def _x_issue(job):
    return R.issue_record(_XS["H"], _XS["x"], _XS["w"], *job)


def _x_icheck(job):
    return R.issuance_check_record(_XS["H"], _XS["w"], *job)


def _x_rcheck(job):
    return R.refund_check_record(_XS["H"], _XS["w"], *job)


@pytest.mark.skipif(not R.available(), reason="libsodium with ristretto255 not found")
def test_whole_mutation_corpus_oracle_vs_independent_stack(octx):
    """EVERY item of every mutation corpus the GPU tests use (mutate_requests, mutate_proofs, mutate_proofs_head,
    tampered_token_proofs, overspend_proofs, mutate_responses, mutate_refunds) goes through the independent stack
    (tests/refstack.py: libsodium ristretto255 + python big ints + python blake3; no code shared with the oracle or the
    engine): status AND output bytes must equal the oracle's, item by item.  With no reference-held vectors and no Rust
    toolchain this is the pin of the oracle's accept/reject behaviour and of its outputs on adversarial inputs."""
    _XS.update(H=[octx.h[0:32], octx.h[32:64], octx.h[64:96]], x=int.from_bytes(octx.x, "little"), w=octx.w)
    PB = O.PROOF_BYTES
    base = corpus.gen_valid(octx, 24, seed=b"xcheck-corpus", threads=8)
    # --- issue (row a1)
    req, cs, rnd, expect, labels = corpus.mutate_requests(octx, base)
    o_resp, o_st, _ = octx.batch_issue(req, cs, rnd, threads=8)
    n = len(expect)
    got = _pool_map(_x_issue, [(req[128 * i:128 * i + 128].tobytes(), cs[32 * i:32 * i + 32].tobytes(), rnd[128 * i:128 * i + 128].tobytes()) for i in range(n)])
    for i, (st, resp) in enumerate(got):
        assert st == o_st[i], ("issue", i, labels[i], st, int(o_st[i]))
        assert resp == o_resp[160 * i:160 * i + 160].tobytes(), ("issue bytes", i, labels[i])
    assert {0, 1, 0x81} <= set(o_st.tolist())
    # --- spend verification + refund (row a2): the three proof corpora
    t = corpus.tampered_token_proofs(octx, 10)
    over = corpus.overspend_proofs(octx, 3)
    head_base = corpus.gen_valid(octx, 16, seed=b"xcheck-head", threads=8)
    sets = [("mutate_proofs",) + corpus.mutate_proofs(octx, base)[:2] + (corpus.mutate_proofs(octx, base)[3],),
            ("mutate_proofs_head",) + corpus.mutate_proofs_head(octx, head_base)[:2] + (corpus.mutate_proofs_head(octx, head_base)[3],),
            ("tampered_token", t["proofs"], t["rnd"], t["labels"]),
            ("overspend", over["proofs"], over["rnd"], ["overspend"] * 3)]
    seen_status = set()
    accepted = []
    for name, proofs, prnd, plabels in sets:
        m = len(proofs) // PB
        o_ref, o_nul, o_pst, _ = octx.batch_refund(proofs, prnd, threads=8)
        got = _pool_map(_x_refund, [(proofs[PB * i:PB * (i + 1)].tobytes(), prnd[128 * i:128 * i + 128].tobytes()) for i in range(m)])
        for i, (st, ref, nul) in enumerate(got):
            assert st == o_pst[i], (name, i, plabels[i], st, int(o_pst[i]))
            assert ref == o_ref[128 * i:128 * i + 128].tobytes() and nul == o_nul[32 * i:32 * i + 32].tobytes(), (name, "bytes", i, plabels[i])
            if st == 0 and name == "mutate_proofs":
                accepted.append((proofs[PB * i + 128:PB * i + 128 + 4096].tobytes(), ref))
        seen_status |= set(o_pst.tolist())
    assert seen_status == {0, 6, 7, 0x81}
    # --- client-side checks (rows a3, a4)
    K, rs, rexp, rlab = corpus.mutate_responses(base)
    o_cst, _ = octx.batch_issuance_check(K, rs, threads=8)
    got = _pool_map(_x_icheck, [(K[32 * i:32 * i + 32].tobytes(), rs[160 * i:160 * i + 160].tobytes()) for i in range(len(rexp))])
    assert got == o_cst.tolist() == rexp.tolist(), list(zip(rlab, got, o_cst.tolist()))
    com = np.frombuffer(b"".join(c for c, _ in accepted), np.uint8).copy(); rfs = np.frombuffer(b"".join(r for _, r in accepted), np.uint8).copy()
    assert len(accepted) >= 4
    # tile the accepted (com, refund) pairs so that every tamper class of mutate_refunds occurs
    reps = (20 + len(accepted) - 1) // len(accepted)
    com = np.tile(com.reshape(-1, 4096), (reps, 1)).reshape(-1); rfs = np.tile(rfs.reshape(-1, 128), (reps, 1)).reshape(-1)
    c2, r2, fexp, flab = corpus.mutate_refunds(com, rfs)
    o_fst, _ = octx.batch_refund_check(c2, r2, threads=8)
    got = _pool_map(_x_rcheck, [(c2[4096 * i:4096 * i + 4096].tobytes(), r2[128 * i:128 * i + 128].tobytes()) for i in range(len(fexp))])
    assert got == o_fst.tolist() == fexp.tolist(), list(zip(flab, got, o_fst.tolist()))
    assert set(got) == {0, 4, 0x81}


@pytest.mark.skipif(not R.available(), reason="libsodium with ristretto255 not found")
def test_differential_fuzz_oracle_vs_independent_stack(octx):
    """Random bit flips and whole-item replacements anywhere in SpendProof and IssuanceRequest records: the oracle and the
    independent stack agree on status and output bytes for every record (the GPU suite runs the same fuzz, engine vs oracle)."""
    _XS.update(H=[octx.h[0:32], octx.h[32:64], octx.h[64:96]], x=int.from_bytes(octx.x, "little"), w=octx.w)
    PB = O.PROOF_BYTES
    base = corpus.gen_valid(octx, 8, seed=b"fuzz-xcheck", threads=8)
    rs = np.random.RandomState(77)
    n = 96
    P = base["proofs"].reshape(8, -1)[rs.randint(0, 8, n)].copy(); Rn = base["rnd"].reshape(8, -1)[rs.randint(0, 8, n)].copy()
    for i in range(n):
        if i % 4 == 3:
            item = rs.randint(0, 526); P[i, 32 * item:32 * item + 32] = rs.randint(0, 256, 32)
        elif i % 4:
            b = rs.randint(0, PB * 8); P[i, b // 8] ^= 1 << (b % 8)
    o_ref, o_nul, o_st, _ = octx.batch_refund(P.reshape(-1), Rn.reshape(-1), threads=8)
    got = _pool_map(_x_refund, [(P[i].tobytes(), Rn[i].tobytes()) for i in range(n)])
    for i, (st, ref, nul) in enumerate(got):
        assert st == o_st[i] and ref == o_ref[128 * i:128 * i + 128].tobytes() and nul == o_nul[32 * i:32 * i + 32].tobytes(), i
    assert {0, 7, 0x81} <= set(o_st.tolist())
    Q = base["req"].reshape(8, -1)[rs.randint(0, 8, n)].copy(); C = base["cs"].reshape(8, -1)[rs.randint(0, 8, n)].copy()
    for i in range(n):
        if i % 3:
            b = rs.randint(0, 128 * 8); Q[i, b // 8] ^= 1 << (b % 8)
        if i % 5 == 0:
            C[i] = rs.randint(0, 256, 32)
    o_resp, o_ist, _ = octx.batch_issue(Q.reshape(-1), C.reshape(-1), Rn.reshape(-1), threads=8)
    got = _pool_map(_x_issue, [(Q[i].tobytes(), C[i].tobytes(), Rn[i].tobytes()) for i in range(n)])
    for i, (st, resp) in enumerate(got):
        assert st == o_ist[i] and resp == o_resp[160 * i:160 * i + 160].tobytes(), i


def test_golden_corpus_checks_on_the_oracle():
    """The committed round-2 fixture (independent-stack statuses and outputs for tampered responses, refunds and tokens)."""
    g = np.load(os.path.join(HERE, "golden", "corpus_checks.npz"))
    ctx = O.Ctx(g["h"].tobytes(), g["x"].tobytes(), g["w"].tobytes())
    assert (ctx.batch_issuance_check(g["K"], g["responses"], threads=4)[0] == g["status_issuance_check"]).all()
    assert (ctx.batch_refund_check(g["com"], g["refunds"], threads=4)[0] == g["status_refund_check"]).all()
    ref, nul, st, _ = ctx.batch_refund(g["token_proofs"], g["token_rnd"], threads=4)
    assert (st == g["token_status"]).all() and (ref == g["token_refunds"]).all() and (nul == g["token_nullifiers"]).all()
    assert set(g["status_issuance_check"].tolist()) == {0, 2, 0x81} and set(g["status_refund_check"].tolist()) == {0, 4, 0x81}


def test_token_lifecycles(octx):
    """The reference's scenario tests (sequential spends, exact balance, zero spend, zero-credit token, one-credit
    exhaustion, 2^120 and 2^128-1 credit tokens, overspend; src/tests.rs:210-426,642-689,876-1059) as multi-generation
    chains on the oracle: every generation's refund becomes the next token."""
    log = corpus.run_lifecycles(corpus.OracleImpl(octx))
    assert log["final_balances"] == [0, 0, 0, 0, 0, 0x1234567890abcdef - 94 - 2**60, 0, 0]
    # determinism: same seeds, same bytes
    again = corpus.run_lifecycles(corpus.OracleImpl(octx), scenarios=corpus.LIFECYCLES[:2])
    assert again["generations"][0]["refunds"][:128] == log["generations"][0]["refunds"][:128]
