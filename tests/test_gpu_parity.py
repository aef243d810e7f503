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


def test_empty_and_single(engine, octx, base):
    e = np.zeros(0, np.uint8)
    resp, st = engine.batch_issue(e, e, e)
    assert resp.size == 0 and st.size == 0
    ref, nul, st = engine.batch_verify_spend_and_refund(e, e)
    assert st.size == 0
    ref, nul, st = engine.batch_verify_spend_and_refund(base["proofs"][:corpus.PROOF_BYTES], base["rnd"][:128])
    o = octx.refund(base["proofs"][:corpus.PROOF_BYTES].tobytes(), base["rnd"][:128].tobytes())
    assert st[0] == o[0] == 0 and ref.tobytes() == o[1] and nul.tobytes() == o[2]
