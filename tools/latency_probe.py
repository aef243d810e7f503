"""Dev tool: latency of small host-buffer calls (n = 1, 32, 1024) through the C ABI, pinned and pageable buffers alike (numpy)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
act = importlib.import_module("anonymous-credit-tokens_b200")
ctx = corpus.make_ctx(corpus.BENCH_PARAMS)
base = corpus.gen_valid(ctx, 64, seed=b"latency", threads=8)
u = 64
with act.Engine(act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w)) as eng:
    for n in (1, 32, 1024, 4096):
        idx = np.arange(n) % u
        P = base["proofs"].reshape(u, -1)[idx].reshape(-1).copy(); R = base["rnd"].reshape(u, -1)[idx].reshape(-1).copy()
        Q = base["req"].reshape(u, -1)[idx].reshape(-1).copy(); C = base["cs"].reshape(u, -1)[idx].reshape(-1).copy()
        for name, fn in (("verify_spend_and_refund", lambda: eng.batch_verify_spend_and_refund(P, R)), ("issue", lambda: eng.batch_issue(Q, C, R))):
            fn(); fn()
            ts = []
            for _ in range(10):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            print(f"{name:26s} n={n:5d}: median {1e3 * sorted(ts)[5]:8.3f} ms  min {1e3 * min(ts):8.3f} ms", flush=True)
