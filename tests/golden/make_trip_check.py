import sys, hashlib, time
sys.path.insert(0, "tests")
import blake3
import oracle_lib as O, refstack as R

stream = blake3.blake3(b"act-oracle-0").digest(length=34112)
H = R.params_new("example-org","payment-api","production","2024-01-15")
h_o = O.params_derive("example-org","payment-api","production","2024-01-15")
print("params match:", b"".join(H) == h_o, H[0].hex())
assert H[0].hex()=="1e2015fd2f2d25c3fb25b0998a6daf6f6b85e0f8f578ff22ae54eeeadd47b15d"
# python stack
t0=time.time()
rng = R.Rng(stream)
x, W = R.keygen(rng)
r, k = rng.scalar(), rng.scalar()
rq = R.request(H, r, k, rng)
rs = R.issue(H, x, W, rq, 40, rng)
assert R.issuance_check(H, W, rq["K"], rs)
tok = dict(A=rs["A"], e=rs["e"], k=k, r=r, c=40)
pf, pre = R.prove_spend(H, tok, 20, rng)
rf = R.refund(H, x, W, pf, rng)
assert isinstance(rf, dict), rf
assert R.refund_check(H, W, pf["com"], rf)
assert rng.o == 34112
print("python trip %.1fs"%(time.time()-t0))
print("x", R.sc_bytes(x).hex()); print("W", W.hex())
print("req sha", hashlib.sha256(R.pack_request(rq)).hexdigest())
print("resp sha", hashlib.sha256(R.pack_response(rs)).hexdigest())
print("proof sha", hashlib.sha256(R.pack_proof(pf)).hexdigest())
print("refund sha", hashlib.sha256(R.pack_refund(rf)).hexdigest())
print("cbor proof sha", hashlib.sha256(R.cbor_proof(pf)).hexdigest(), len(R.cbor_proof(pf)))
# oracle
t0=time.time()
o=0
xo, wo = O.keygen(stream[0:64]); o=64
assert xo==R.sc_bytes(x) and wo==W, "keygen"
ctx = O.Ctx(h_o, xo, wo)
pre64 = O.sc_reduce64(stream[64:128]) + O.sc_reduce64(stream[128:192]); o=192
req = ctx.request(pre64, stream[o:o+128]); o+=128
assert req==R.pack_request(rq), "request"
c40 = (40).to_bytes(32,'little')
st, resp = ctx.issue(req, c40, stream[o:o+128]); o+=128
assert st==0 and resp==R.pack_response(rs), "issue"
assert ctx.issuance_check(req[:32], resp)==0
token = resp[:64] + pre64[32:] + pre64[:32] + c40
proof, prer = ctx.prove_spend(token, (20).to_bytes(32,'little'), stream[o:o+524*64]); o+=524*64
assert proof==R.pack_proof(pf), "proof"
assert prer == R.sc_bytes(pre["k"])+R.sc_bytes(pre["r"])+R.sc_bytes(pre["m"])
st, refund, nul = ctx.refund(proof, stream[o:o+128]); o+=128
assert st==0 and refund==R.pack_refund(rf), "refund"
assert nul == R.sc_bytes(k)
assert ctx.refund_check(proof[128:128+4096], refund)==0
print("oracle trip %.3fs, consumed %d"%(time.time()-t0, o))
print("ALL MATCH")
