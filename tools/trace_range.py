"""Dev tool: per-warp residency trace of spend_range_kernel (library built with -DACT_RANGE_TRACE=1).
usage: python tools/trace_range.py <variant name> <n proofs>
Prints how the warps of the launch were spread over SMs and in time (globaltimer), and the units each one processed."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import variant_bench as vb
act = importlib.import_module("anonymous-credit-tokens_b200")
name, n = sys.argv[1], int(sys.argv[2])
act.LIB_PATH = os.path.join(ROOT, "tools", "bin", f"libact_{name}.so")
ctx, d = vb.fixtures()
u = len(d["o_st"])
proofs = np.tile(d["proofs"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
rnd = np.tile(d["rnd"].reshape(u, -1), ((n + u - 1) // u, 1))[:n].reshape(-1).copy()
eng = act.Engine(act.Params(ctx.h), act.PrivateKey(ctx.x, ctx.w))
for rep in range(2):
    ref, nul, st = eng.batch_verify_spend_and_refund(proofs, rnd)
assert (st == 0).all()
lib = C.CDLL(act.LIB_PATH)
W = 4096
buf = (C.c_ulonglong * (4 * W))()
rc = lib.act_debug_read_trace(buf, C.c_size_t(4 * W))
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.uint64).reshape(W, 4).astype(np.int64)
smid, t0, t1, units = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
live = t1 > 0
smid, t0, t1, units = smid[live], t0[live], t1[live], units[live]
base = t0.min()
dur = (t1 - t0) / 1e6
print(f"variant {name} n={n}: warps traced {live.sum()}, distinct SMs {len(set(smid.tolist()))}")
cnt = np.bincount(smid)
print("  warps per SM: min %d max %d  (histogram %s)" % (cnt[cnt > 0].min(), cnt.max(), np.bincount(cnt[cnt > 0]).tolist()))
print("  start spread  %.3f ms; end: first %.3f ms, median %.3f ms, last %.3f ms after launch" % ((t0.max() - base) / 1e6, (t1.min() - base) / 1e6, (np.median(t1) - base) / 1e6, (t1.max() - base) / 1e6))
print("  warp lifetime ms: min %.3f median %.3f max %.3f" % (dur.min(), np.median(dur), dur.max()))
total = (t1.max() - base) / 1e6
print("  mean resident warps / launched warps over the launch: %.3f" % (dur.sum() / (total * len(dur))))
print("  units per warp: min %d median %d max %d sum %d" % (units.min(), np.median(units), units.max(), units.sum()))
# per-SM finish time
fin = np.zeros(cnt.size); 
for s in set(smid.tolist()): fin[s] = (t1[smid == s].max() - base) / 1e6
f = fin[fin > 0]
print("  per-SM finish ms: min %.3f median %.3f max %.3f" % (f.min(), np.median(f), f.max()))
w0 = dur[0::4] if len(dur) == 592 * 4 else None
if w0 is not None:
    print("  lifetime of warp 0 of each block: median %.3f; warps 1-3: median %.3f" % (np.median(dur[0::4]), np.median(np.concatenate([dur[1::4], dur[2::4], dur[3::4]]))))
