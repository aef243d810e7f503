"""Dev tool: time the spend pipeline of several library variants (tools/bin/libact_<name>.so) on the GPU.
usage: python tools/variant_bench.py N name1 name2 ...   ("default" = the in-tree library)
Each variant runs in its own process; prints per-kernel device ms and proofs/s, and checks outputs against the oracle."""
import importlib, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
KINDS = ["range", "head", "hash", "finish", "sign", "issue", "icheck", "rcheck", "encode"]


def fixtures(u=256):
    import corpus
    path = "/tmp/variant_fixtures2.npz"
    ctx = corpus.make_ctx(corpus.BENCH_PARAMS)
    if os.path.exists(path):
        d = np.load(path)
        return ctx, {k: d[k] for k in d.files}
    base = corpus.gen_valid(ctx, u, seed=b"variant-bench", threads=os.cpu_count())
    o_ref, o_nul, o_st, _ = ctx.batch_refund(base["proofs"], base["rnd"], threads=os.cpu_count())
    o_resp, o_ist, _ = ctx.batch_issue(base["req"], base["cs"], base["rnd"])
    d = {"proofs": base["proofs"], "rnd": base["rnd"], "o_ref": o_ref, "o_nul": o_nul, "o_st": o_st,
         "req": base["req"], "cs": base["cs"], "o_resp": o_resp, "o_ist": o_ist}
    np.savez(path, **d)
    return ctx, d


def child(n, name):
    import ctypes as C
    act = importlib.import_module("anonymous-credit-tokens_b200")
    if name != "default":
        act.LIB_PATH = os.path.join(ROOT, "tools", "bin", f"libact_{name}.so")
    ctx, d = fixtures()
    u = len(d["o_st"])
    proofs = np.tile(d["proofs"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
    rnd = np.tile(d["rnd"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
    eng = act.Engine(act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w))
    ref, nul, st = eng.batch_verify_spend_and_refund(proofs, rnd)   # warm-up + correctness
    full = (n // u) * u
    ok = bool((st == 0).all() and (ref[:full * 128].reshape(n // u, u * 128) == d["o_ref"].reshape(1, -1)).all()
              and (nul[:full * 32].reshape(n // u, u * 32) == d["o_nul"].reshape(1, -1)).all())
    lib = act.load_library()
    lib.act_engine_set_timing(eng._h, 1)
    import time
    t0 = time.time()
    reps = 2
    for _ in range(reps):
        eng.batch_verify_spend_and_refund(proofs, rnd)
    wall = (time.time() - t0) / reps
    ms = (C.c_double * 9)(); cnt = (C.c_uint64 * 9)()
    lib.act_engine_get_timing(eng._h, ms, cnt)
    k = {KINDS[i]: round(ms[i] / reps, 3) for i in range(9) if cnt[i]}
    tot = sum(k.values())
    # batch_issue of 16 n requests (tiled), device time of the issue kernel
    ni = 16 * n
    req = np.tile(d["req"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy(); cs = np.tile(d["cs"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy()
    irnd = np.tile(d["rnd"].reshape(u, -1), (ni // u, 1)).reshape(-1).copy()
    resp, ist = eng.batch_issue(req, cs, irnd)
    ok = ok and bool((ist == 0).all() and (resp.reshape(ni // u, -1) == d["o_resp"].reshape(1, -1)).all())
    lib.act_engine_get_timing(eng._h, ms, cnt)
    for _ in range(reps):
        eng.batch_issue(req, cs, irnd)
    lib.act_engine_get_timing(eng._h, ms, cnt)
    k["issue"] = round(ms[5] / reps, 3)
    print(json.dumps({"variant": name, "n": n, "ok": ok, "issues_per_s": round(ni / k["issue"] * 1e3), "kernel_ms": k, "sum_ms": round(tot, 2), "proofs_per_s_kernels": round(n / tot * 1e3), "range_proofs_per_s": round(n / k["range"] * 1e3), "wall_e2e_proofs_per_s": round(n / wall)}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
    else:
        n = int(sys.argv[1])
        fixtures()
        for name in sys.argv[2:]:
            r = subprocess.run([sys.executable, __file__, "--child", str(n), name], capture_output=True, text=True)
            print(r.stdout.strip() or ("FAILED " + name + ": " + r.stderr[-800:]), flush=True)
