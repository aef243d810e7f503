"""Constant-time evidence for the signing paths, by counting (dev tool; run by tools/run_ct_counts.sh on the GPU box).

A kernel with no secret-dependent branch, predicate or address executes EXACTLY the same warp instructions, the same
thread instructions (no secret-dependent predication or divergence) and touches the same number of memory sectors
(no secret-dependent address, at 32-byte granularity) whatever the secret is.  `run` drives the three kernels that
handle secrets -- the refund signing pass, the issue signing pass and the head stage of spend verification -- over
FIXED public inputs while the secrets vary: four issuer keys x (1, l-1 and two random ones; W follows), and three
settings of the alpha half of the signer randomness (all-zero, all-ones, random) with the e half fixed (e is public:
it is part of the response).  `ncu` records the counters of every launch; `summarise` groups the launches by kernel
and checks that every counter is identical across the variants.

    ncu --metrics ... --csv --log-file counts.csv python tools/ct_counts.py run
    python tools/ct_counts.py summarise counts.csv profiles/r02_ct_counts.txt
"""
import csv
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
N = 2048                                   # requests / proofs per launch (a multiple of every block size)
KERNELS = ("refund_sign_seq_kernel", "issue_mode_kernel", "spend_head_kernel", "finalize_ctx_kernel")


def run():
    import corpus
    import oracle_lib as O
    act = importlib.import_module("anonymous-credit-tokens_b200")
    ctx = corpus.make_ctx(corpus.TEST_PARAMS)
    base = corpus.gen_valid(ctx, 64, seed=b"ct-counts", threads=8)
    u = 64
    proofs = np.tile(base["proofs"].reshape(u, -1), (N // u, 1)).reshape(-1).copy()
    req = np.tile(base["req"].reshape(u, -1), (N // u, 1)).reshape(-1).copy()
    cs = np.tile(base["cs"].reshape(u, -1), (N // u, 1)).reshape(-1).copy()
    ell = corpus.ELL
    keys = [(1).to_bytes(32, "little"), (ell - 1).to_bytes(32, "little"), O.sc_reduce64(corpus.xof(b"ct-key-a", 64)), O.sc_reduce64(corpus.xof(b"ct-key-b", 64))]
    e_half = np.frombuffer(corpus.xof(b"ct-e-half", 64 * N), np.uint8).reshape(N, 64)
    alphas = [np.zeros((N, 64), np.uint8), np.full((N, 64), 255, np.uint8), np.frombuffer(corpus.xof(b"ct-alpha", 64 * N), np.uint8).reshape(N, 64)]
    all_ok = np.zeros(N, np.uint8)
    params = act.Params(ctx.h)
    kp = None
    for ki, x in enumerate(keys):
        key = act.PrivateKey.from_secret(x)
        with act.Engine(params, key) as eng:                  # finalize_ctx_kernel: W = G*x with the constant-time table walk
            nul, st, kprime = eng.batch_spend_verify(proofs)   # spend_head_kernel: the A' term's scalar depends on x
            if kp is None:
                kp = kprime                                    # K' is a function of the proofs only: the same for every key
            assert (kprime == kp).all()
            for ai, al in enumerate(alphas):
                rnd = np.concatenate([e_half, al], axis=1).reshape(-1).copy()
                eng.batch_refund_sign(kp, all_ok, rnd)         # refund_sign_seq_kernel on identical public inputs
                eng.batch_issue_sign(req, cs, all_ok, rnd)     # issue_mode_kernel (sign mode) on identical public inputs
        print(f"key {ki} done", file=sys.stderr)


def summarise(csv_path, out_path):
    rows = list(csv.reader(open(csv_path, errors="ignore")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    per = {}
    for r in rows[start + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        if not any(k in name for k in KERNELS):
            continue
        per.setdefault(name, {}).setdefault(d["ID"], {})[d["Metric Name"]] = d["Metric Value"]
    lines = ["# constant-time evidence by counting: tools/ct_counts.py (see its docstring); every launch of a kernel below ran on the SAME public",
             "# inputs with a DIFFERENT secret (4 issuer keys x; for the signing kernels also 3 settings of alpha).  A counter that differs",
             "# between launches would mean a secret-dependent instruction stream, predicate or address.", ""]
    ok_all = True
    for name, launches in sorted(per.items()):
        if "issue_mode_kernel" in name:   # the verify-mode launches of the same kernel (act_batch_issue_verify is not run here) -- none expected
            pass
        ids = sorted(launches, key=int)
        metrics = sorted({m for l in launches.values() for m in l})
        lines.append(f"## {name}: {len(ids)} launches")
        for m in metrics:
            vals = [launches[i].get(m, "?") for i in ids]
            same = len(set(vals)) == 1
            ok_all &= same
            lines.append(f"  {'IDENTICAL' if same else 'DIFFERS  '}  {m} = {vals[0] if same else vals}")
        lines.append("")
    lines.append("RESULT: " + ("PASS -- every counter of every secret-handling kernel is identical across all secrets" if ok_all and per else "FAIL"))
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return 0 if ok_all and per else 1


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        sys.exit(summarise(sys.argv[2], sys.argv[3]))
