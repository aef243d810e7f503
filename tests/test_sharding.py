"""N>1 path on CPU: world_size-2 gloo processes shard a batch by contiguous ranges, compute their shard (the engine's
logic through tests/hostsim, standing in for the GPU), gather status+nullifiers, and must equal the single-process result."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import corpus
    import hostsim_lib as HS
    import replay_reference
    sh = importlib.import_module("anonymous-credit-tokens_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ctx = corpus.make_ctx(corpus.TEST_PARAMS)
    base = corpus.gen_valid(ctx, n, seed=b"shard", threads=2)
    proofs, rnd, _, _ = corpus.mutate_proofs(ctx, base)
    lo, hi = sh.shard_bounds(n, rank, world)
    hs = HS.Ctx(ctx.h, ctx.x, ctx.w)
    ref, nul, st = hs.refund(proofs[lo * corpus.PROOF_BYTES:hi * corpus.PROOF_BYTES], rnd[lo * 128:hi * 128])
    g_st, g_nul = sh.all_gather_results(torch.from_numpy(st), torch.from_numpy(nul), n)
    if rank == 0:
        o_ref, o_nul, o_st, _ = ctx.batch_refund(proofs, rnd, threads=2)
        q.put((g_st.numpy().tolist() == o_st.tolist(), bool((g_nul.numpy() == o_nul).all()), replay_reference.flag_replays(g_st, g_nul).tolist(), o_st.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    n, world = 9, 2   # ragged: 5 + 4
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    ps = [ctxm.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    ok_st, ok_nul, flagged, o_st = q.get(timeout=300)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_st and ok_nul
    # proof 7 is an exact replay of proof 0 (corpus kind 7): both accepted by refund, second flagged by the screen
    assert o_st[0] == 0 and o_st[7] == 0 and flagged[0] == 0 and flagged[7] == 3


def test_shard_bounds_and_replay_flags():
    sh = importlib.import_module("anonymous-credit-tokens_b200.sharding")
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [sh.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    rs = np.random.RandomState(1)
    n = 500
    nul = rs.randint(0, 256, size=(n, 32)).astype(np.uint8)
    nul[rs.randint(0, n, 100)] = nul[rs.randint(0, n, 100)]          # plant duplicates
    st = (rs.rand(n) < 0.2).astype(np.uint8) * 7
    seen = nul[rs.randint(0, n, 20)].copy()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import replay_reference
    got = replay_reference.flag_replays(torch.from_numpy(st), torch.from_numpy(nul.reshape(-1)), torch.from_numpy(seen.reshape(-1))).numpy()
    db = {bytes(s) for s in seen}; exp = st.copy()
    for i in range(n):                                               # NullifierDb semantics of src/tests.rs:28-50
        if st[i] == 0:
            k = bytes(nul[i])
            if k in db:
                exp[i] = 3
            db.add(k)
    assert got.tolist() == exp.tolist()
