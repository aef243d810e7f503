"""Generator + bit-level emulator for the PTX squaring path of fe25519.cuh.

The off-diagonal products a_i*a_j (i<j) are laid out as carry chains over two accumulators (even / odd
32-bit column parity) so that every chain touches non-overlapping 64-bit slots; the result is doubled,
the diagonal squares are added in one more chain, and the 512-bit value is folded by fe_reduce512.
Running this file checks the layout against python big ints and prints the CUDA source of
fe_sq_wide() (paste between the GENERATED markers in fe25519.cuh).
"""
import random

M32 = 0xffffffff


def chains():
    out = []
    for i in range(7):
        odd = [(i, j) for j in range(i + 1, 8) if (i + j) % 2 == 1]
        even = [(i, j) for j in range(i + 1, 8) if (i + j) % 2 == 0]
        if odd:
            out.append(("od", odd[0][0] + odd[0][1] - 1, odd))
        if even:
            out.append(("ev", even[0][0] + even[0][1], even))
    return out


def capture_flags():
    """A chain needs its final `addc top, top, 0` only if its last word may be non-zero before the chain runs: into a word
    that is still zero, hi(a*b) + carry <= 2^32 - 1 cannot carry out (tools/check_fe_rows.py checks this exhaustively on
    extreme operands)."""
    written = {"ev": set(), "od": set()}
    flags = []
    for acc, start, prods in chains():
        top = start + 2 * len(prods) - 1
        need = top in written[acc]
        flags.append(need)
        written[acc].update(range(start, top + 1))
        if need:
            written[acc].add(top + 1)
    return flags


def emulate(a):
    ev = [0] * 18; od = [0] * 18
    for (acc_name, start, prods), need in zip(chains(), capture_flags()):
        acc = ev if acc_name == "ev" else od
        carry = 0
        w = start
        for (i, j) in prods:
            p = a[i] * a[j]
            t = acc[w] + (p & M32) + carry; acc[w] = t & M32; carry = t >> 32
            t = acc[w + 1] + (p >> 32) + carry; acc[w + 1] = t & M32; carry = t >> 32
            w += 2
        if need:
            t = acc[w] + carry; acc[w] = t & M32
            assert t >> 32 == 0
        else:
            assert carry == 0
    # r = ev + (od << 32)
    r = [0] * 16
    r[0] = ev[0]; carry = 0
    for k in range(1, 16):
        t = ev[k] + od[k - 1] + carry; r[k] = t & M32; carry = t >> 32
    assert carry == 0 and ev[16] == 0 and od[15] == 0 and od[16] == 0
    # double
    carry = 0
    for k in range(16):
        t = 2 * r[k] + carry; r[k] = t & M32; carry = t >> 32
    assert carry == 0
    # diagonal
    carry = 0
    for i in range(8):
        p = a[i] * a[i]
        t = r[2 * i] + (p & M32) + carry; r[2 * i] = t & M32; carry = t >> 32
        t = r[2 * i + 1] + (p >> 32) + carry; r[2 * i + 1] = t & M32; carry = t >> 32
    assert carry == 0
    return r


def check():
    rnd = random.Random(1)
    for it in range(2000):
        a = [rnd.getrandbits(32) for _ in range(8)]
        if it % 5 == 0:
            a = [M32] * 8
        if it % 7 == 0:
            a = [rnd.choice([0, M32, 1]) for _ in range(8)]
        r = emulate(a)
        A = sum(x << (32 * i) for i, x in enumerate(a))
        R = sum(x << (32 * i) for i, x in enumerate(r))
        assert R == A * A, it
    print("// layout verified against big-int squaring on 2000 vectors")


def emit():
    L = []
    L.append("ACT_FN void fe_sq_wide(u32* r, const fe& a) {")
    L.append("    u32 ev[16], od[16];")
    L.append("    ACT_UNROLL for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }")
    for (acc, start, prods), need in zip(chains(), capture_flags()):
        n = len(prods)
        words = list(range(start, start + 2 * n + (1 if need else 0)))
        ops = []
        # operand numbering: outputs first (2n+1), then a-limb inputs
        limbs = []
        for (i, j) in prods:
            for x in (i, j):
                if x not in limbs:
                    limbs.append(x)
        base = len(words)
        def opn(x):
            return "%%%d" % (base + limbs.index(x))
        lines = []
        for k, (i, j) in enumerate(prods):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            lines.append('%s %%%d, %s, %s, %%%d;' % (lo, 2 * k, opn(i), opn(j), 2 * k))
            hi = "madc.hi.cc.u32" if (need or k < n - 1) else "madc.hi.u32"
            lines.append('%s %%%d, %s, %s, %%%d;' % (hi, 2 * k + 1, opn(i), opn(j), 2 * k + 1))
        if need:
            lines.append('addc.u32 %%%d, %%%d, 0;' % (2 * n, 2 * n))
        body = '\\n\\t"\n        "'.join(lines)
        outs = ", ".join('"+r"(%s[%d])' % (acc, w) for w in words)
        ins = ", ".join('"r"(a.v[%d])' % x for x in limbs)
        L.append('    asm("%s"\n        : %s\n        : %s);' % (body, outs, ins))
    # combine ev + od<<32 into r (15-word chain), double, add diagonal
    L.append("    ACT_UNROLL for (int i = 0; i < 16; i++) r[i] = ev[i];")
    lines = ['add.cc.u32 %0, %0, %15;'] + ['addc.cc.u32 %%%d, %%%d, %%%d;' % (k, k, 15 + k) for k in range(1, 14)] + ['addc.u32 %14, %14, %29;']
    body = '\\n\\t"\n        "'.join(lines)
    outs = ", ".join('"+r"(r[%d])' % k for k in range(1, 16))
    ins = ", ".join('"r"(od[%d])' % k for k in range(0, 15))
    L.append('    asm("%s"\n        : %s\n        : %s);' % (body, outs, ins))
    lines = ['add.cc.u32 %0, %0, %0;'] + ['addc.cc.u32 %%%d, %%%d, %%%d;' % (k, k, k) for k in range(1, 15)] + ['addc.u32 %15, %15, %15;']
    body = '\\n\\t"\n        "'.join(lines)
    outs = ", ".join('"+r"(r[%d])' % k for k in range(16))
    L.append('    asm("%s"\n        : %s);' % (body, outs))
    lines = []
    for i in range(8):
        lo = "mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32"
        hi = "madc.hi.cc.u32" if i < 7 else "madc.hi.u32"
        lines.append('%s %%%d, %%%d, %%%d, %%%d;' % (lo, 2 * i, 16 + i, 16 + i, 2 * i))
        lines.append('%s %%%d, %%%d, %%%d, %%%d;' % (hi, 2 * i + 1, 16 + i, 16 + i, 2 * i + 1))
    body = '\\n\\t"\n        "'.join(lines)
    ins = ", ".join('"r"(a.v[%d])' % i for i in range(8))
    L.append('    asm("%s"\n        : %s\n        : %s);' % (body, outs, ins))
    L.append("}")
    return "\n".join(L)


if __name__ == "__main__":
    check()
    print(emit())
