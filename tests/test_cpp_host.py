"""include/act.hpp -- the C++ host-side mirror of the reference crate's interface (the Rust wrapper of rust/ cannot be compiled in
this image).  CPU: the header and its check program compile and link against the product library.  GPU: the program drives
act::Engine::batch_issue / batch_verify_spend_and_refund with ONE shared RNG stream and every output byte equals a loop of oracle
calls over that stream (randomness consumed only by accepted requests, in slice order: src/lib.rs:638-643, 842-846)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import corpus

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")
EXE = os.path.join(CPP, "act_hpp_check")


def _build(act):
    assert os.path.exists(act.LIB_PATH)
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return EXE


def test_cpp_host_mirror_compiles_and_links(act):
    exe = _build(act)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr          # no arguments: prints its usage, touches no device


def _run(act, octx, tmp_path, devices=None):
    exe = _build(act)
    base = corpus.gen_valid(octx, 24, seed=b"cpp-host", threads=8)
    req, cs, _, iexp, ilab = corpus.mutate_requests(octx, base)
    proofs, _, pexp, plab = corpus.mutate_proofs(octx, base)
    nr, npf = len(iexp), len(pexp)
    stream = corpus.xof(b"cpp-host/one-rng", 128 * (nr + npf))
    fix = tmp_path / "fixture.bin"; out = tmp_path / "out.bin"
    with open(fix, "wb") as f:
        f.write(octx.h + octx.x + octx.w)
        f.write(struct.pack("<I", nr) + req.tobytes() + cs.tobytes())
        f.write(struct.pack("<I", npf) + proofs.tobytes())
        f.write(struct.pack("<I", len(stream)) + stream)
    cmd = [exe, str(fix), str(out)] + ([",".join(map(str, devices))] if devices else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    b = open(out, "rb").read()
    o = 0

    def take(n):
        nonlocal o
        v = b[o:o + n]; o += n
        return v
    ist = list(take(nr)); resp = take(nr * 160); pos1 = struct.unpack("<I", take(4))[0]; ic = list(take(nr))
    pst = list(take(npf)); nul = take(npf * 32); ref = take(npf * 128); pos2 = struct.unpack("<I", take(4))[0]; rc = list(take(npf))
    assert o == len(b)
    # expected: a loop of oracle calls over the same stream
    pos = 0
    for i in range(nr):
        st, e_resp = octx.issue(req[128 * i:128 * i + 128].tobytes(), cs[32 * i:32 * i + 32].tobytes(), stream[pos:pos + 128])
        assert ist[i] == st, (i, ilab[i])
        if st == 0:
            pos += 128
            assert resp[160 * i:160 * i + 160] == e_resp, (i, ilab[i])
            assert ic[i] == 0
        else:
            assert not any(resp[160 * i:160 * i + 160]) and ic[i] != 0
    assert pos1 == pos
    for i in range(npf):
        st, e_ref, e_nul = octx.refund(proofs[i * corpus.PROOF_BYTES:(i + 1) * corpus.PROOF_BYTES].tobytes(), stream[pos:pos + 128])
        assert pst[i] == st, (i, plab[i])
        if st == 0:
            pos += 128
            assert ref[128 * i:128 * i + 128] == e_ref and nul[32 * i:32 * i + 32] == e_nul, (i, plab[i])
            assert rc[i] == 0
        else:
            assert not any(ref[128 * i:128 * i + 128]) and not any(nul[32 * i:32 * i + 32])
    assert pos2 == pos and 0 < pos < len(stream)
    assert {0, 1, 0x81} <= set(ist) and {0, 6, 7, 0x81} <= set(pst)


@pytest.mark.gpu
def test_cpp_host_mirror_matches_a_loop_of_reference_calls(act, octx, tmp_path):
    _run(act, octx, tmp_path)


@pytest.mark.gpu
def test_cpp_host_mirror_on_a_multi_device_handle(act, octx, tmp_path):
    _run(act, octx, tmp_path, devices=[0, 1] if act.device_count() >= 2 else [0, 0])
