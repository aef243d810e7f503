"""Driver for ncu: one pass over every kernel family on small batches (spend pipeline, issue, checks, prover, aux)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
act = importlib.import_module("anonymous-credit-tokens_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
ctx = corpus.make_ctx(corpus.BENCH_PARAMS)
u = 64
base = corpus.gen_valid(ctx, u, seed=b"prof-all", threads=os.cpu_count())
st = corpus.trip_streams(b"prof-all", u)
rep = (n + u - 1) // u
tile = lambda a, rec: np.tile(a.reshape(u, rec), (rep, 1))[:n].reshape(-1).copy()
eng = act.Engine(act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w))
tokens, charges = tile(corpus.tokens_from(base, st["pre"]), 160), tile(corpus.charges_from(base), 32)
proofs, prer, pst = eng.batch_prove_spend(tokens, charges, seed=bytes(range(32)))
assert (pst == 0).all()
ref, nul, s = eng.batch_verify_spend_and_refund(proofs, tile(base["rnd"], 128))
assert (s == 0).all()
ni = 8 * n
req = eng.batch_request(np.tile(st["pre"].reshape(u, 64), (ni // u, 1)).reshape(-1).copy(), np.tile(st["req_rnd"].reshape(u, 128), (ni // u, 1)).reshape(-1).copy())
resp, ist = eng.batch_issue(req, np.tile(base["cs"].reshape(u, 32), (ni // u, 1)).reshape(-1).copy(), np.tile(base["rnd"].reshape(u, 128), (ni // u, 1)).reshape(-1).copy())
assert (ist == 0).all()
eng.batch_issuance_check(req.reshape(-1, 128)[:, :32].copy().reshape(-1), resp)
eng.batch_refund_check(proofs.reshape(n, -1)[:, 128:128 + 4096].copy().reshape(-1), ref)
eng.flag_replays(s, nul)
eng.unpack_cbor(act.KIND_PROOF, eng.encode_cbor(act.KIND_PROOF, proofs))
print("ok", n)
