"""Dev tool: every entry point once on a tiny batch, meant to run under compute-sanitizer (memcheck / initcheck)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
act = importlib.import_module("anonymous-credit-tokens_b200")
ctx = corpus.make_ctx(corpus.TEST_PARAMS)
n = 5
base = corpus.gen_valid(ctx, n, seed=b"sanitize", threads=4)
st = corpus.trip_streams(b"sanitize", n)
params = act.Params.new(*corpus.TEST_PARAMS)
with act.Engine(params, act.PrivateKey.from_secret(ctx.x)) as eng:
    proofs, rnd, expect, _ = corpus.mutate_proofs(ctx, base)
    ref, nul, s = eng.batch_verify_spend_and_refund(proofs, rnd)
    req, cs, irnd, _, _ = corpus.mutate_requests(ctx, base)
    resp, ist = eng.batch_issue(req, cs, irnd)
    eng.batch_issuance_check(base["req"].reshape(-1, 128)[:, :32].copy().reshape(-1), base["resp"])
    eng.batch_refund_check(proofs.reshape(n, -1)[:, 128:128 + 4096].copy().reshape(-1), ref)
    eng.flag_replays(s, nul, nul[:64])
    enc = eng.encode_cbor(act.KIND_PROOF, proofs); eng.unpack_cbor(act.KIND_PROOF, enc)
    eng.batch_request(st["pre"], st["req_rnd"])
    tokens, charges = corpus.tokens_from(base, st["pre"]), corpus.charges_from(base)
    eng.batch_prove_spend(tokens, charges, rnd=st["prove_rnd"])
    eng.batch_prove_spend(tokens, charges, seed=bytes(32))
    stream = np.zeros(128 * n, np.uint8)
    eng.batch_verify_spend_and_refund_seq(proofs, stream)
    eng.batch_issue_seq(req, cs, stream)
    # round 2: two-pass forms, the screened call, a chunk-crossing batch with ramp chunks, and a two-replica engine
    nul2, st2, kp = eng.batch_spend_verify(proofs)
    eng.batch_refund_sign(kp, st2, stream)
    eng.batch_issue_sign(req, cs, eng.batch_issue_verify(req), stream)
    eng.batch_verify_spend_and_refund_screened(proofs, rnd, seen=nul[:64])
    eng.set_spend_chunk(4096)
    m = 4096 + 700
    u = len(proofs) // corpus.PROOF_BYTES
    idx = np.arange(m) % u
    eng.batch_verify_spend_and_refund(proofs.reshape(u, -1)[idx].reshape(-1).copy(), rnd.reshape(u, -1)[idx].reshape(-1).copy())
with act.Engine(params, act.PrivateKey.from_secret(ctx.x), devices=[0, 0]) as meng:
    meng.batch_verify_spend_and_refund_screened(proofs, rnd)
    meng.batch_issue(req, cs, irnd)
print("sanitize_run ok", s.tolist(), ist.tolist())
