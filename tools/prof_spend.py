"""Minimal driver for ncu: runs the spend pipeline (and optionally issue) once on N synthetic proofs."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
act = importlib.import_module("anonymous-credit-tokens_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = corpus.make_ctx(corpus.BENCH_PARAMS)
u = min(n, 256)
base = corpus.gen_valid(ctx, u, seed=b"bench-spend", threads=os.cpu_count())
proofs = np.tile(base["proofs"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
rnd = np.tile(base["rnd"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
eng = act.Engine(act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w))
for _ in range(reps):
    ref, nul, st = eng.batch_verify_spend_and_refund(proofs, rnd)
assert (st == 0).all()
ni = n * 8
resp, st = eng.batch_issue(np.tile(base["req"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy(), np.tile(base["cs"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy(),
                           np.tile(base["rnd"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy())
assert (st == 0).all()
print("ok", n)
