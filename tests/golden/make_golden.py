"""Generates the committed golden fixtures under tests/golden/.

  trip.json          BASELINE config #1 (examples/act.rs: issue 40 -> spend 20 -> refund) with RNG = BLAKE3-XOF("act-oracle-0"),
                     produced by the INDEPENDENT python stack tests/refstack.py (blake3-py + libsodium ristretto255 + big ints),
                     i.e. not by the oracle and not by the CUDA engine.  Values equal SURVEY.md Appendix C.
  corpus_small.npz   a small adversarial batch (valid + mutated requests/proofs) with the oracle's outputs, so the GPU
                     parity tests can also be checked against committed bytes.

  corpus_checks.npz  (round 2) tampered IssuanceResponses / Refunds for the client-side checks and proofs made from tampered tokens,
                     with statuses and outputs computed by the independent stack and asserted equal to the oracle's.

Run from the repo root:  python tests/golden/make_golden.py [trip] [corpus_small] [corpus_checks]
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import blake3  # noqa: E402
import corpus  # noqa: E402
import oracle_lib as O  # noqa: E402
import refstack as R  # noqa: E402


def trip():
    params = ("example-org", "payment-api", "production", "2024-01-15")
    stream = blake3.blake3(b"act-oracle-0").digest(length=34112)
    H = R.params_new(*params)
    rng = R.Rng(stream)
    x, W = R.keygen(rng)
    r, k = rng.scalar(), rng.scalar()
    rq = R.request(H, r, k, rng)
    rs = R.issue(H, x, W, rq, 40, rng)
    assert R.issuance_check(H, W, rq["K"], rs)
    pf, pre = R.prove_spend(H, dict(A=rs["A"], e=rs["e"], k=k, r=r, c=40), 20, rng)
    rf = R.refund(H, x, W, pf, rng)
    assert R.refund_check(H, W, pf["com"], rf) and rng.o == 34112
    d = dict(params=list(params), rng="BLAKE3-XOF of b'act-oracle-0', 34112 bytes, consumed in reference order (SURVEY Appendix B)",
             h=b"".join(H).hex(), x=R.sc_bytes(x).hex(), w=W.hex(), pre_r=R.sc_bytes(r).hex(), pre_k=R.sc_bytes(k).hex(),
             request=R.pack_request(rq).hex(), response=R.pack_response(rs).hex(), proof=R.pack_proof(pf).hex(),
             refund=R.pack_refund(rf).hex(), nullifier=R.sc_bytes(k).hex(),
             prerefund=(R.sc_bytes(pre["k"]) + R.sc_bytes(pre["r"]) + R.sc_bytes(pre["m"])).hex(),
             cbor_request=R.cbor_request(rq).hex(), cbor_response=R.cbor_response(rs).hex(), cbor_refund=R.cbor_refund(rf).hex(),
             cbor_proof_sha256=hashlib.sha256(R.cbor_proof(pf)).hexdigest(),
             sha256=dict(request=hashlib.sha256(R.pack_request(rq)).hexdigest(), response=hashlib.sha256(R.pack_response(rs)).hexdigest(),
                         proof=hashlib.sha256(R.pack_proof(pf)).hexdigest(), refund=hashlib.sha256(R.pack_refund(rf)).hexdigest()))
    json.dump(d, open(os.path.join(HERE, "trip.json"), "w"), indent=1)
    print("trip.json", d["sha256"])


def corpus_small():
    ctx = corpus.make_ctx(corpus.TEST_PARAMS)
    base = corpus.gen_valid(ctx, 12, seed=b"golden-small", threads=4)
    req, cs, rnd_i, exp_i, _ = corpus.mutate_requests(ctx, base)
    resp, st_i, _ = ctx.batch_issue(req, cs, rnd_i, threads=4)
    proofs, rnd, exp, _ = corpus.mutate_proofs(ctx, base)
    ref, nul, st, _ = ctx.batch_refund(proofs, rnd, threads=4)
    np.savez_compressed(os.path.join(HERE, "corpus_small.npz"), h=np.frombuffer(ctx.h, np.uint8), x=np.frombuffer(ctx.x, np.uint8),
                        w=np.frombuffer(ctx.w, np.uint8), req=req, cs=cs, rnd_issue=rnd_i, resp=resp, status_issue=st_i,
                        proofs=proofs, rnd=rnd, refunds=ref, nullifiers=nul, status=st)
    print("corpus_small.npz", st_i.tolist(), st.tolist())


def corpus_checks():
    """Round 2: the tamper classes of the client-side checks (rows a3, a4) and the tampered-token proofs, with statuses computed by
    the INDEPENDENT stack (tests/refstack.py record-level front ends) and asserted equal to the oracle's."""
    ctx = corpus.make_ctx(corpus.TEST_PARAMS)
    H = [ctx.h[0:32], ctx.h[32:64], ctx.h[64:96]]
    x = int.from_bytes(ctx.x, "little")
    base = corpus.gen_valid(ctx, 20, seed=b"golden-checks", threads=4)
    K, rs, _, _ = corpus.mutate_responses(base)
    st_ic = np.array([R.issuance_check_record(H, ctx.w, K[32 * i:32 * i + 32], rs[160 * i:160 * i + 160]) for i in range(20)], np.uint8)
    assert st_ic.tolist() == ctx.batch_issuance_check(K, rs, threads=4)[0].tolist()
    ref, nul, st, _ = ctx.batch_refund(base["proofs"], base["rnd"], threads=4)
    assert (st == 0).all()
    com = base["proofs"].reshape(20, -1)[:, 128:128 + 4096].copy().reshape(-1)
    c2, r2, _, _ = corpus.mutate_refunds(com, ref)
    st_rc = np.array([R.refund_check_record(H, ctx.w, c2[4096 * i:4096 * i + 4096], r2[128 * i:128 * i + 128]) for i in range(20)], np.uint8)
    assert st_rc.tolist() == ctx.batch_refund_check(c2, r2, threads=4)[0].tolist()
    t = corpus.tampered_token_proofs(ctx, 10)
    outs = [R.refund_record(H, x, ctx.w, t["proofs"][O.PROOF_BYTES * i:O.PROOF_BYTES * (i + 1)], t["rnd"][128 * i:128 * i + 128]) for i in range(10)]
    o_ref, o_nul, o_st, _ = ctx.batch_refund(t["proofs"], t["rnd"], threads=4)
    assert [o[0] for o in outs] == o_st.tolist() and b"".join(o[1] for o in outs) == o_ref.tobytes() and b"".join(o[2] for o in outs) == o_nul.tobytes()
    np.savez_compressed(os.path.join(HERE, "corpus_checks.npz"), h=np.frombuffer(ctx.h, np.uint8), x=np.frombuffer(ctx.x, np.uint8),
                        w=np.frombuffer(ctx.w, np.uint8), K=K, responses=rs, status_issuance_check=st_ic, com=c2, refunds=r2,
                        status_refund_check=st_rc, token_proofs=t["proofs"], token_rnd=t["rnd"], token_status=np.array([o[0] for o in outs], np.uint8),
                        token_refunds=np.frombuffer(b"".join(o[1] for o in outs), np.uint8), token_nullifiers=np.frombuffer(b"".join(o[2] for o in outs), np.uint8))
    print("corpus_checks.npz", st_ic.tolist(), st_rc.tolist(), [o[0] for o in outs])


if __name__ == "__main__":
    which = sys.argv[1:] or ["trip", "corpus_small", "corpus_checks"]
    for w in which:
        {"trip": trip, "corpus_small": corpus_small, "corpus_checks": corpus_checks}[w]()
