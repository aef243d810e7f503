"""The PTX paths of fe25519.cuh cannot run without a GPU; their carry-chain LAYOUT can: tools/check_fe_rows.py and
tools/gen_fe_sq.py emulate the row chains word by word (including which carry captures are dropped) against python big ints,
and the squaring in the header must be exactly what the generator emits."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_pruned_carry_captures_never_fire():
    m = _load("check_fe_rows")
    rnd = random.Random(11)
    mr, sr = m.mul_rows(), m.sq_rows()
    mf, sf = m.capture_flags(mr), m.capture_flags(sr)
    assert sum(mf) == 7 and len(mf) == 16 and sum(sf) == 5 and len(sf) == 13
    for it in range(3000):
        a, b = m.operands(rnd, it), m.operands(rnd, it * 3 + 1)
        A = sum(x << (32 * i) for i, x in enumerate(a)); B = sum(x << (32 * i) for i, x in enumerate(b))
        assert m.run(mr, mf, a, b) == A * B
        off = m.run(sr, sf, a, a)
        assert 2 * off + sum((a[i] * a[i]) << (64 * i) for i in range(8)) == A * A


def test_header_squaring_is_the_generated_one():
    g = _load("gen_fe_sq")
    g.check()
    src = open(os.path.join(ROOT, "anonymous-credit-tokens_b200", "csrc", "fe25519.cuh")).read()
    a = src.index("ACT_FN void fe_sq_wide(u32* r, const fe& a) {")
    b = src.index("// ---- end GENERATED ----")
    assert src[a:b].strip() == g.emit().strip()
    # the multiplication rows in the header follow the capture flags of the emulator
    m = _load("check_fe_rows")
    flags = m.capture_flags(m.mul_rows())
    assert flags == [False, False, True, False] + [True, False, True, False] * 3
