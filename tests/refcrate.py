"""TEST INFRASTRUCTURE ONLY: consumer of tests/golden/ref_crate.json -- known-answer vectors dumped from the REAL
anonymous-credit-tokens crate by rust/golden-dump (every call's inputs as the crate's CBOR, the exact RNG bytes the call
drew, and its outcome).  `check(doc, impl)` replays every call through an implementation (the CPU oracle, or the CUDA
engine) and compares statuses and output bytes; `make_like_dumper()` writes a file of the same schema from the
independent libsodium/big-int stack (tests/refstack.py) so that the consumer itself is tested while no Rust toolchain
is available to produce the real file."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_crate.json")


def load(path=GOLDEN):
    return json.load(open(path)) if os.path.exists(path) else None


def _flat(cbor, nfields):
    """payloads of a canonical flat map {1: bstr32, ...} as one record"""
    b = bytes.fromhex(cbor) if isinstance(cbor, str) else bytes(cbor)
    assert len(b) == 1 + 35 * nfields and b[0] == 0xA0 | nfields, "not the canonical flat encoding"
    return b"".join(b[1 + 35 * i + 3:1 + 35 * i + 35] for i in range(nfields))


class OracleBackend:
    """the CPU oracle behind the calls check() makes"""
    name = "oracle"

    def __init__(self, act, h, x, w):
        import oracle_lib as O
        self.act, self.c = act, O.Ctx(h, x, w)

    def issue(self, req, c32, rnd):
        st, resp = self.c.issue(req, c32, rnd)
        return st, resp

    def refund(self, proof, rnd):
        return self.c.refund(proof, rnd)

    def issuance_check(self, K, resp):
        return self.c.issuance_check(K, resp)

    def refund_check(self, com, refund):
        return self.c.refund_check(com, refund)

    def request(self, pre, rnd):
        return self.c.request(pre, rnd)

    def prove_spend(self, token, charge32, rnd):
        return self.c.prove_spend(token, charge32, rnd)


class EngineBackend:
    """the CUDA engine through the C ABI"""
    name = "engine"

    def __init__(self, act, h, x, w):
        self.act = act
        self.e = act.Engine(act.Params(h), act.PrivateKey(x, w))

    def issue(self, req, c32, rnd):
        resp, st = self.e.batch_issue(req, c32, rnd)
        return int(st[0]), resp.tobytes()

    def refund(self, proof, rnd):
        ref, nul, st = self.e.batch_verify_spend_and_refund(proof, rnd)
        return int(st[0]), ref.tobytes(), nul.tobytes()

    def issuance_check(self, K, resp):
        return int(self.e.batch_issuance_check(K, resp)[0])

    def refund_check(self, com, refund):
        return int(self.e.batch_refund_check(com, refund)[0])

    def request(self, pre, rnd):
        return self.e.batch_request(pre, rnd).tobytes()

    def prove_spend(self, token, charge32, rnd):
        p, q, st = self.e.batch_prove_spend(token, charge32, rnd=rnd)
        assert st[0] == 0
        return p.tobytes(), q.tobytes()


def check(doc, act, backend_cls):
    """Replays every call of the golden document; returns the number of calls checked per op.  Raises on the first mismatch."""
    key = _flat(doc["private_key_cbor"], 2)
    x, w = key[:32], key[32:]
    assert bytes.fromhex(doc["public_key_cbor"]) == b"\x58\x20" + w
    h = act.Params.new(*doc["params"]).h if backend_cls is EngineBackend else __import__("oracle_lib").params_derive(*doc["params"])
    be = backend_cls(act, h, x, w)
    counts = {}
    pad = lambda hexs, n: (bytes.fromhex(hexs) + bytes(n))[:n]
    for i, c in enumerate(doc["calls"]):
        op = c["op"]; where = (i, op, c.get("label"), c.get("trip"))
        counts[op] = counts.get(op, 0) + 1
        if op == "issue":
            rec, pst = act.pack_issuance_requests_cbor([bytes.fromhex(c["request_cbor"])])
            assert pst[0] == 0, where
            assert len(c["rng"]) == (256 if c["status"] == 0 else 0), ("the reference draws 128 bytes on accept, none on reject", where)
            st, resp = be.issue(rec.tobytes(), bytes.fromhex(c["c"]), pad(c["rng"], 128))
            assert st == c["status"], where
            if st == 0:
                assert act.encode_issuance_response_cbor(resp).hex() == c["out_cbor"], where
        elif op == "refund":
            rec, pst = act.pack_spend_proofs_cbor([bytes.fromhex(c["proof_cbor"])])
            assert pst[0] == 0, where
            assert len(c["rng"]) == (256 if c["status"] == 0 else 0), where
            st, ref, nul = be.refund(rec.tobytes(), pad(c["rng"], 128))
            assert st == c["status"], where
            if st == 0:
                assert act.encode_refund_cbor(ref).hex() == c["out_cbor"] and nul.hex() == c["nullifier"], where
        elif op == "issuance_check":
            rq, _ = act.pack_issuance_requests_cbor([bytes.fromhex(c["request_cbor"])])
            rs, pst = act.pack_issuance_responses_cbor([bytes.fromhex(c["response_cbor"])])
            assert pst[0] == 0 and be.issuance_check(rq.tobytes()[:32], rs.tobytes()) == c["status"], where
        elif op == "refund_check":
            pf, _ = act.pack_spend_proofs_cbor([bytes.fromhex(c["proof_cbor"])])
            rf, pst = act.pack_refunds_cbor([bytes.fromhex(c["refund_cbor"])])
            assert pst[0] == 0 and be.refund_check(pf.tobytes()[128:128 + 4096], rf.tobytes()) == c["status"], where
        elif op == "request":
            pre = _flat(c["preissuance_cbor"], 2)        # PreIssuance = {1: r, 2: k} (src/cbor.rs:537-556) = the `pre` record r | k
            assert len(c["rng"]) == 256, where
            assert act.encode_issuance_request_cbor(be.request(pre, bytes.fromhex(c["rng"]))).hex() == c["out_cbor"], where
        elif op == "prove_spend":
            token = _flat(c["token_cbor"], 5)            # CreditToken = {1: a, 2: e, 3: k, 4: r, 5: c} (src/cbor.rs:585-603) = the token record
            assert len(c["rng"]) == 2 * 524 * 64, where
            proof, prer = be.prove_spend(token, int(c["charge"]).to_bytes(32, "little"), bytes.fromhex(c["rng"]))
            assert act.encode_spend_proof_cbor(proof).hex() == c["out_cbor"], where
            if "prerefund_cbor" in c:                    # PreRefund = {1: r, 2: k, 3: m}; the record is k* | r* | m
                q = _flat(c["prerefund_cbor"], 3)
                assert prer == q[32:64] + q[0:32] + q[64:96], where
        else:
            raise AssertionError(f"unknown op {op}")
    return counts


def make_like_dumper(path, trips=3):
    """A document with the dumper's schema, produced by the independent stack (libsodium + big ints + python blake3 + cbor2):
    exercises check() end to end.  It is NOT a pin to the reference crate and is never written to tests/golden/."""
    import cbor2
    import blake3
    import refstack as R
    stream = blake3.blake3(b"refcrate-like").digest(length=64 * (600 * trips + 64))
    pos = [0]

    class Rec(R.Rng):
        def __init__(self):
            self.log = b""

        def scalar(self):
            b = stream[pos[0]:pos[0] + 64]; pos[0] += 64
            self.log += b
            return R.sc_wide(b)

        def take(self):
            s = self.log.hex(); self.log = b""
            return s

    rng = Rec()
    params = ["test-org", "test-service", "test-env", "2024-01-01"]
    H = R.params_new(*params)
    x = rng.scalar(); W = R.mul_base(x)
    sb = R.sc_bytes
    doc = {"source": "tests/refstack.py (independent stack) -- schema self-test, NOT the reference crate", "params": params,
           "private_key_cbor": cbor2.dumps({1: sb(x), 2: W}).hex(), "public_key_cbor": cbor2.dumps(W).hex(), "private_key_rng": rng.take(), "calls": []}
    calls = doc["calls"]
    for trip in range(trips):
        credits = 20 + 61 * trip; charge = 1 + (7 * trip) % credits
        r, k = rng.scalar(), rng.scalar(); pre_rng = rng.take()
        rq = R.request(H, r, k, rng)
        calls.append({"op": "request", "trip": trip, "preissuance_cbor": cbor2.dumps({1: sb(r), 2: sb(k)}).hex(), "preissuance_rng": pre_rng,
                      "rng": rng.take(), "out_cbor": R.cbor_request(rq).hex()})
        for label, mut in (("valid", None), ("k_bar+1", "k_bar"), ("gamma+1", "gamma")):
            q = dict(rq)
            if mut:
                q[mut] = (q[mut] + 1) % R.ELL
            rs = R.issue(H, x, W, q, credits, rng)
            calls.append({"op": "issue", "trip": trip, "label": label, "request_cbor": R.cbor_request(q).hex(), "c": sb(credits).hex(), "rng": rng.take(),
                          "status": 0 if rs else 1, "out_cbor": R.cbor_response(rs).hex() if rs else ""})
            if not mut:
                resp = rs
        for label, mut in (("valid", None), ("e+1", "e"), ("z+1", "z"), ("c+1", "c")):
            q = dict(resp)
            if mut:
                q[mut] = (q[mut] + 1) % R.ELL
            calls.append({"op": "issuance_check", "trip": trip, "label": label, "request_cbor": R.cbor_request(rq).hex(), "response_cbor": R.cbor_response(q).hex(),
                          "status": 0 if R.issuance_check(H, W, rq["K"], q) else 2})
        tok = dict(A=resp["A"], e=resp["e"], k=k, r=r, c=credits)
        tok_cbor = cbor2.dumps({1: tok["A"], 2: sb(tok["e"]), 3: sb(k), 4: sb(r), 5: sb(credits)}).hex()
        pf, prer = R.prove_spend(H, tok, charge, rng)
        calls.append({"op": "prove_spend", "trip": trip, "token_cbor": tok_cbor, "charge": charge, "rng": rng.take(), "out_cbor": R.cbor_proof(pf).hex(),
                      "prerefund_cbor": cbor2.dumps({1: sb(prer["r"]), 2: sb(prer["k"]), 3: sb(prer["m"])}).hex()})
        for label, mut in (("valid", None), ("s+1", "s"), ("e_bar+1", "e_bar"), ("A' identity", "Ap")):
            q = dict(pf)
            if mut == "Ap":
                q["Ap"] = bytes(32)
            elif mut:
                q[mut] = (q[mut] + 1) % R.ELL
            rf = R.refund(H, x, W, q, rng)
            st = 0 if isinstance(rf, dict) else (6 if rf == "identity" else 7)
            calls.append({"op": "refund", "trip": trip, "label": label, "proof_cbor": R.cbor_proof(q).hex(), "rng": rng.take(), "status": st,
                          "nullifier": sb(q["k"]).hex(), "out_cbor": R.cbor_refund(rf).hex() if st == 0 else ""})
            if not mut:
                refund = rf
        for label, mut in (("valid", None), ("e+1", "e"), ("gamma+1", "gamma"), ("z+1", "z")):
            q = dict(refund)
            if mut:
                q[mut] = (q[mut] + 1) % R.ELL
            calls.append({"op": "refund_check", "trip": trip, "label": label, "proof_cbor": R.cbor_proof(pf).hex(), "refund_cbor": R.cbor_refund(q).hex(),
                          "status": 0 if R.refund_check(H, W, pf["com"], q) else 4})
    json.dump(doc, open(path, "w"))
    return doc
