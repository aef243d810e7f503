"""ctypes binding of tests/hostsim/libact_hostsim.so: the device headers compiled for the host.

TEST INFRASTRUCTURE ONLY -- lets `-m "not gpu"` tests exercise the engine's arithmetic/protocol logic
against the oracle without a GPU.  The product (libact_b200.so) has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_SO = os.path.join(_DIR, "libact_hostsim.so")
PROOF_BYTES = 526 * 32
_lib = None


def build():
    subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, cp, sz, i32 = C.c_void_p, C.c_char_p, C.c_size_t, C.c_int
        L.hs_ctx_create.argtypes = [vp, vp, vp]; L.hs_ctx_create.restype = vp
        L.hs_ctx_destroy.argtypes = [vp]
        L.hs_params_derive.argtypes = [cp, cp, cp, cp, vp]; L.hs_params_derive.restype = i32
        L.hs_public_key.argtypes = [vp, vp, vp]
        L.hs_issue.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.hs_issuance_check.argtypes = [vp, sz, vp, vp, vp]
        L.hs_refund.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.hs_refund_check.argtypes = [vp, sz, vp, vp, vp]
        L.hs_fe_mul.argtypes = [vp, vp, vp]
        L.hs_fe_addsub.argtypes = [vp, vp, vp, vp]
        L.hs_fe_invert.argtypes = [vp, vp]
        L.hs_decode_encode.argtypes = [vp, vp]; L.hs_decode_encode.restype = i32
        L.hs_from_uniform.argtypes = [vp, vp]
        L.hs_scalarmult.argtypes = [vp, vp, i32, vp]; L.hs_scalarmult.restype = i32
        L.hs_scalarmult_base.argtypes = [vp, i32, vp, i32, vp]
        L.hs_sc_reduce32.argtypes = [vp, vp]; L.hs_sc_reduce64.argtypes = [vp, vp]
        L.hs_sc_muladd.argtypes = [vp, vp, vp, vp]; L.hs_sc_invert.argtypes = [vp, vp]
        L.hs_sc_negsub.argtypes = [vp, vp, vp, vp]
        L.hs_blake3_small.argtypes = [vp, sz, vp]
        L.hs_recode.argtypes = [vp, i32, vp]
        L.hs_encode_stage.argtypes = [vp, vp]; L.hs_encode_stage.restype = i32
        L.hs_sc_half.argtypes = [vp, vp]
        L.hs_request.argtypes = [vp, sz, vp, vp, vp]
        L.hs_prove_spend.argtypes = [vp, sz, vp, vp, vp, vp, C.c_uint64, vp, vp, vp]
        L.hs_cbor_skeleton_encode.argtypes = [i32, vp, vp]; L.hs_cbor_skeleton_encode.restype = i32
        L.hs_cbor_skeleton_unpack.argtypes = [i32, vp, vp]; L.hs_cbor_skeleton_unpack.restype = i32
        _lib = L
    return _lib


def _in(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b)
    return a, a.ctypes.data


def _out(n, dt=np.uint8):
    a = np.zeros(n, dtype=dt)
    return a, a.ctypes.data


def call1(name, *ins, out=32):
    """call lib().<name>(in..., out) with one output buffer"""
    keep = [_in(x) for x in ins]
    o, op = _out(out)
    getattr(lib(), name)(*[k[1] for k in keep], op)
    return o.tobytes()


def params_derive(org, svc, dep, ver):
    o, op = _out(96)
    assert lib().hs_params_derive(org.encode(), svc.encode(), dep.encode(), ver.encode(), op) == 0
    return o.tobytes()


class Ctx:
    def __init__(self, h, x, w):
        a, ap = _in(h); b, bp = _in(x); c, cp = _in(w)
        self.p = lib().hs_ctx_create(ap, bp, cp)
        if not self.p:
            raise ValueError("invalid point")

    def __del__(self):
        try:
            if self.p:
                lib().hs_ctx_destroy(self.p); self.p = None
        except Exception:
            pass

    def public_key(self, x):
        a, ap = _in(x); o, op = _out(32)
        lib().hs_public_key(self.p, ap, op)
        return o.tobytes()

    def issue(self, req, cs, rnd):
        n = len(req) // 128
        a, ap = _in(req); b, bp = _in(cs); c, cp = _in(rnd)
        o, op = _out(n * 160); s, sp = _out(n)
        lib().hs_issue(self.p, n, ap, bp, cp, op, sp)
        return o, s

    def issuance_check(self, K, resp):
        n = len(K) // 32
        a, ap = _in(K); b, bp = _in(resp); s, sp = _out(n)
        lib().hs_issuance_check(self.p, n, ap, bp, sp)
        return s

    def refund(self, proofs, rnd):
        n = len(proofs) // PROOF_BYTES
        a, ap = _in(proofs); b, bp = _in(rnd)
        o, op = _out(n * 128); u, up = _out(n * 32); s, sp = _out(n)
        lib().hs_refund(self.p, n, ap, bp, op, up, sp)
        return o, u, s

    def refund_check(self, com, refund):
        n = len(refund) // 128
        a, ap = _in(com); b, bp = _in(refund); s, sp = _out(n)
        lib().hs_refund_check(self.p, n, ap, bp, sp)
        return s

    def request(self, pre, rnd):
        n = len(pre) // 64
        a, ap = _in(pre); b, bp = _in(rnd); o, op = _out(n * 128)
        lib().hs_request(self.p, n, ap, bp, op)
        return o

    def prove_spend(self, tokens, charges, rnd=None, seed=None, first_index=0):
        n = len(tokens) // 160
        a, ap = _in(tokens); b, bp = _in(charges)
        r, rp = _in(rnd) if rnd is not None else (None, None)
        sd, sdp = _in(seed) if seed is not None else (None, None)
        o, op = _out(n * PROOF_BYTES); q, qp = _out(n * 96); s, sp = _out(n)
        lib().hs_prove_spend(self.p, n, ap, bp, rp, sdp, first_index, op, qp, sp)
        return o, q, s

    def scalarmult_base(self, base, s, ct=0):
        a, ap = _in(s); o, op = _out(32)
        lib().hs_scalarmult_base(self.p, base, ap, ct, op)
        return o.tobytes()


def cbor_skeleton_encode(kind, rec, out_len):
    a, ap = _in(rec); o, op = _out(out_len)
    n = lib().hs_cbor_skeleton_encode(kind, ap, op)
    return o[:n].tobytes()


def cbor_skeleton_unpack(kind, cbor, rec_len):
    a, ap = _in(cbor); o, op = _out(rec_len)
    st = lib().hs_cbor_skeleton_unpack(kind, ap, op)
    return o.tobytes(), st
