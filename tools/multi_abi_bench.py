"""Dev tool: one process, one multi-device handle over G GPUs -- times act_batch_verify_spend_and_refund[_screened] against G
single-device engines driven from G python threads.  usage: python tools/multi_abi_bench.py G per_gpu [portable|default|torch]"""
import ctypes as C
import importlib
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
act = importlib.import_module("anonymous-credit-tokens_b200")
G = int(sys.argv[1]); per = int(sys.argv[2]); mode = sys.argv[3] if len(sys.argv) > 3 else "torch"
PB = act.PROOF_BYTES
ctx = corpus.make_ctx(corpus.BENCH_PARAMS)
params, key = act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w)
n = G * per
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
rb = lambda k: torch.randint(0, 256, (k,), dtype=torch.uint8, device=dev, generator=gen)
with act.Engine(params, key, device=0) as e0:
    S = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(S):
        pre = rb(n * 64); pre.view(n, 64)[:, 31] &= 0x0f; pre.view(n, 64)[:, 63] &= 0x0f
        req = torch.empty(n * 128, dtype=torch.uint8, device=dev)
        e0.batch_request_dev(n, pre.data_ptr(), rb(n * 128).data_ptr(), req.data_ptr(), S.cuda_stream)
        cs = torch.zeros(n, 32, dtype=torch.uint8, device=dev); cs[:, 0] = 200
        resp = torch.empty(n * 160, dtype=torch.uint8, device=dev); ist = torch.empty(n, dtype=torch.uint8, device=dev)
        e0.batch_issue_dev(n, req.data_ptr(), cs.data_ptr(), rb(n * 128).data_ptr(), resp.data_ptr(), ist.data_ptr(), S.cuda_stream)
        tok = torch.cat([resp.view(n, 160)[:, :64], pre.view(n, 64)[:, 32:], pre.view(n, 64)[:, :32], cs], 1).contiguous()
        ch = torch.zeros(n, 32, dtype=torch.uint8, device=dev); ch[:, 0] = 7
        pf = torch.empty(n * PB, dtype=torch.uint8, device=dev); pr = torch.empty(n * 96, dtype=torch.uint8, device=dev); ps = torch.empty(n, dtype=torch.uint8, device=dev)
        e0.batch_prove_spend_dev(n, tok.data_ptr(), ch.data_ptr(), None, bytes(range(32)), 0, pf.data_ptr(), pr.data_ptr(), ps.data_ptr(), S.cuda_stream)
        rnd = rb(n * 128)
    S.synchronize()
    assert bool((ps == 0).all())
lib = act.load_library()


def host(nbytes):
    if mode == "torch":
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return t, t.data_ptr()
    p = lib.act_host_alloc(nbytes)
    assert p
    return (C.c_uint8 * nbytes).from_address(p), p


hp, hp_p = host(n * PB); hr, hr_p = host(n * 128); href, href_p = host(n * 128); hnul, hnul_p = host(n * 32); hst, hst_p = host(n)
torch.frombuffer(hp, dtype=torch.uint8).copy_(pf) if mode != "torch" else hp.copy_(pf)
torch.frombuffer(hr, dtype=torch.uint8).copy_(rnd) if mode != "torch" else hr.copy_(rnd)
del pf, tok, resp, req, pre
torch.cuda.empty_cache()
print(f"G={G} per_gpu={per} host memory: {mode}", flush=True)
with act.Engine(params, key, devices=list(range(G))) as m:
    for name, fn in (("screened", lambda: m.batch_verify_spend_and_refund_screened_ptr(n, hp_p, hr_p, 0, None, href_p, hnul_p, hst_p)),
                     ("plain", lambda: m.batch_verify_spend_and_refund_ptr(n, hp_p, hr_p, href_p, hnul_p, hst_p))):
        fn()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        print(f"multi handle, {name}: steps {[round(t, 3) for t in ts]} s -> {n / min(ts):.0f} proofs/s (best), {3 * n / sum(ts):.0f} (mean)", flush=True)
engs = [act.Engine(params, key, device=g) for g in range(G)]


def one(g):
    engs[g].batch_verify_spend_and_refund_ptr(per, hp_p + g * per * PB, hr_p + g * per * 128, href_p + g * per * 128, hnul_p + g * per * 32, hst_p + g * per)


def allg():
    th = [threading.Thread(target=one, args=(g,)) for g in range(G)]
    [t.start() for t in th]; [t.join() for t in th]


allg()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); allg(); ts.append(time.perf_counter() - t0)
print(f"{G} single-device engines from {G} python threads: steps {[round(t, 3) for t in ts]} s -> {n / min(ts):.0f} proofs/s (best)", flush=True)
for g in range(G):
    t0 = time.perf_counter(); one(g); print(f"  GPU {g} alone: {per / (time.perf_counter() - t0):.0f} proofs/s", flush=True)
[e.close() for e in engs]
